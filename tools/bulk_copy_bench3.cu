// Microbenchmark 3: what costs ~735 cycles per producer iteration?  One producer thread, 16 KB requests, 8 stages.
// variant 0: try_wait(empty) ; arrive.expect_tx ; cp.async.bulk            (baseline)
// variant 1: arrive.expect_tx.relaxed
// variant 2: cp.async.bulk first, then arrive.expect_tx
// variant 3: no empty wait at all for the first `stages` iterations only (iters == stages): pure issue cost
// variant 4: baseline but consumer/empty handshake removed: barriers pre-armed per iteration by the CONSUMER thread
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint64_t* b, uint32_t par) {
  asm volatile("{\n.reg .pred P1;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DN;\nbra LW;\nDN:\n}" ::"r"(s32(b)), "r"(par) : "memory");
}
__global__ void __launch_bounds__(64, 1) k(const float* src, size_t span_floats, int req_bytes, int stages, int iters, int variant,
                                           long long* cycles, long long* per_iter) {
  extern __shared__ __align__(1024) uint8_t sm[];
  uint64_t* full = reinterpret_cast<uint64_t*>(sm);
  uint64_t* empty = full + 16;
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
  const size_t stage_floats = req_bytes / 4;
  const size_t base = (size_t(blockIdx.x) * 7919u * stage_floats) % span_floats;
  long long t0 = clock64();
  if (warp == 0 && lane == 0) {
    long long tp = clock64();
    for (int i = 0; i < iters; ++i) {
      const int s = i % stages;
      mwait(&empty[s], ((i / stages) & 1) ^ 1);
      const size_t off = ((base + size_t(i) * stage_floats) % (span_floats - stage_floats)) & ~size_t(3);
      if (variant == 2)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(buf + size_t(s) * req_bytes)), "l"(src + off), "r"(req_bytes), "r"(s32(&full[s])) : "memory");
      if (variant == 1)
        asm volatile("mbarrier.arrive.expect_tx.relaxed.cta.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(req_bytes) : "memory");
      else
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(req_bytes) : "memory");
      if (variant != 2)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(buf + size_t(s) * req_bytes)), "l"(src + off), "r"(req_bytes), "r"(s32(&full[s])) : "memory");
      if (blockIdx.x == 0 && i < 16) { long long t = clock64(); per_iter[i] = t - tp; tp = t; }
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
int main() {
  const size_t span = size_t(16) << 20;
  float* src; cudaMalloc(&src, span * 4); cudaMemset(src, 0, span * 4);
  long long *cyc, *pi; cudaMalloc(&cyc, 148 * 8); cudaMalloc(&pi, 16 * 8);
  long long h[148], hp[16];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024);
  for (int ctas : {1, 148}) for (int variant : {0, 1, 2}) for (int stages : {8}) for (int iters : {8, 480}) {
    const int req = 16 * 1024;
    for (int rep = 0; rep < 2; ++rep) k<<<ctas, 64, 1024 + req * stages>>>(src, span, req, stages, iters, variant, cyc, pi);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost); cudaMemcpy(hp, pi, 16 * 8, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("ctas=%3d variant=%d stages=%d iters=%3d : %.0f clk/iter total; producer per-iter issue gaps:", ctas, variant, stages, iters, mx / iters);
    for (int i = 0; i < 12; ++i) printf(" %lld", hp[i]);
    printf("\n");
  }
  return 0;
}
