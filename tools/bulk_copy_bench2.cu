// Microbenchmark 2: cp.async.bulk L2->SMEM fill rate per SM vs request size, stages in flight and number of issuing warps.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint64_t* b, uint32_t par) {
  asm volatile("{\n.reg .pred P1;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DN;\nbra LW;\nDN:\n}" ::"r"(s32(b)), "r"(par) : "memory");
}
// P producer warps (warps 0..P-1), each owns stages s with s % P == w.  Each stage = `nreq` requests of req_bytes issued
// back-to-back by lane 0 of the owning warp.  Consumer warp (warp 8) waits each stage in order and frees it.
__global__ void __launch_bounds__(288, 1) k(const float* src, size_t span_floats, int req_bytes, int nreq, int stages, int P,
                                            int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 16;
  uint8_t* buf = sm + 1024;
  const int stage_bytes = req_bytes * nreq;
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
  const size_t base = (size_t(blockIdx.x) * 7919u * stage_floats) % span_floats;
  long long t0 = clock64();
  if (warp < P && lane == 0) {
    for (int i = warp; i < iters; i += P) {
      const int s = i % stages;
      mwait(&empty[s], ((i / stages) & 1) ^ 1);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(stage_bytes) : "memory");
      const size_t off = ((base + size_t(i) * stage_floats) % (span_floats - stage_floats)) & ~size_t(3);
      for (int r = 0; r < nreq; ++r)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(buf + size_t(s) * stage_bytes + size_t(r) * req_bytes)), "l"(src + off + size_t(r) * (req_bytes / 4)),
                     "r"(req_bytes), "r"(s32(&full[s])) : "memory");
    }
  } else if (warp == 8 && lane == 0) {
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      mwait(&full[s], (i / stages) & 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}
int main() {
  const size_t span = size_t(16) << 20;
  float* src; cudaMalloc(&src, span * 4); cudaMemset(src, 0, span * 4);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  const int iters = 480;
  printf("req_KB nreq stages P : B/clk/SM  (clk per stage)\n");
  int cfgs[][4] = {{4,1,8,1},{8,1,8,1},{16,1,8,1},{16,1,12,1},{32,1,6,1},{48,1,4,1},{64,1,3,1},{96,1,2,1},
                   {16,1,8,2},{16,1,8,4},{16,1,12,4},{32,1,6,2},{48,1,4,2},{48,1,4,4},
                   {16,3,4,1},{16,3,4,2},{16,3,4,4},{8,6,4,1},{8,6,4,4},{4,12,4,1},{4,12,4,4},{24,2,4,1},{24,2,4,2},
                   {2,24,4,1},{2,24,4,4}};
  for (auto& c : cfgs) {
    int req = c[0] * 1024, nreq = c[1], stages = c[2], P = c[3];
    for (int rep = 0; rep < 2; ++rep) k<<<148, 288, 1024 + req * nreq * stages>>>(src, span, req, nreq, stages, P, iters, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%3d %2d %2d %d : %.1f  (%.0f)\n", c[0], nreq, stages, P, double(req) * nreq * iters / mx, mx / iters);
  }
  return 0;
}
