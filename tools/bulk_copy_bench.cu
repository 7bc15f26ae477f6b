// Microbenchmark: L2 -> shared-memory fill rate per SM for cp.async.bulk (1D TMA) vs ld.global+st.shared, one CTA per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bulk_copy_bench tools/bulk_copy_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint64_t* b, uint32_t par) {
  asm volatile("{\n.reg .pred P1;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DN;\nbra LW;\nDN:\n}" ::"r"(s32(b)), "r"(par) : "memory");
}
// mode 0: bulk copies, `split` requests per stage issued by `split` lanes of warp 0; consumer = warp 1 just waits + arrives
__global__ void __launch_bounds__(256, 1) k_bulk(const float* src, size_t span_floats, int stage_bytes, int stages, int split,
                                                 int iters, int shared_src, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 8;
  uint8_t* buf = sm + 1024;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t stage_floats = stage_bytes / 4;
  const size_t base = shared_src ? 0 : (size_t(blockIdx.x) * 7919u * stage_floats) % span_floats;
  long long t0 = clock64();
  if (warp == 0) {
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      if (lane == 0) {
        mwait(&empty[s], ((i / stages) & 1) ^ 1);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(stage_bytes) : "memory");
      }
      __syncwarp();
      if (lane < split) {
        const int piece = stage_bytes / split;
        const size_t off = (base + size_t(i) * stage_floats) % (span_floats - stage_floats);
        const float* g = src + (off & ~size_t(3)) + size_t(lane) * (piece / 4);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(buf + size_t(s) * stage_bytes + size_t(lane) * piece)), "l"(g), "r"(piece), "r"(s32(&full[s])) : "memory");
      }
    }
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      mwait(&full[s], (i / stages) & 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
// mode 1: all 256 threads ld.global.v4 -> st.shared.v4, unrolled x8
__global__ void __launch_bounds__(256, 1) k_ldg(const float4* src, size_t span_vec, int bytes_per_iter, int iters, int shared_src,
                                                long long* cycles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  float4* buf = reinterpret_cast<float4*>(sm);
  const int vec_per_iter = bytes_per_iter / 16;
  const size_t base = shared_src ? 0 : (size_t(blockIdx.x) * 7919u * vec_per_iter) % span_vec;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    const size_t off = (base + size_t(i) * vec_per_iter) % (span_vec - vec_per_iter);
    for (int v = threadIdx.x; v < vec_per_iter; v += 256 * 8) {
      float4 r[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) r[u] = __ldcg(src + off + v + u * 256);
#pragma unroll
      for (int u = 0; u < 8; ++u) buf[(v + u * 256) % 8192] = r[u];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
int main() {
  const size_t span = size_t(16) << 20;   // 16 Mi floats = 64 MB: L2 resident
  float* src; cudaMalloc(&src, span * 4); cudaMemset(src, 0, span * 4);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k_ldg, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int iters = 400;
  for (int ctas : {148, 64, 16}) for (int shared_src : {0, 1}) {
    for (int stage_kb : {16, 48}) for (int stages : {2, 4}) for (int split : {1, 4, 16}) {
      if (stage_kb * stages > 192) continue;
      for (int rep = 0; rep < 2; ++rep) k_bulk<<<ctas, 256, 1024 + stage_kb * 1024 * stages>>>(src, span, stage_kb * 1024, stages, split, iters, shared_src, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
      double mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("bulk ctas=%3d shared_src=%d stage=%2dKB stages=%d split=%2d : %.1f B/clk/SM (chip %.2f KB/clk)\n", ctas, shared_src, stage_kb,
             stages, split, double(stage_kb) * 1024 * iters / mx, double(stage_kb) * iters * ctas / mx);
    }
    for (int rep = 0; rep < 2; ++rep) k_ldg<<<ctas, 256, 132 * 1024>>>(reinterpret_cast<float4*>(src), span / 4, 64 * 1024, iters, shared_src, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("ldg  ctas=%3d shared_src=%d 64KB/iter 256 thr x8 unroll : %.1f B/clk/SM\n", ctas, shared_src, 64.0 * 1024 * iters / mx);
  }
  return 0;
}
