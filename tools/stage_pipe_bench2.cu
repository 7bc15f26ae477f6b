// Microbenchmark 5: the operand pipeline with cta_group::2 (CTA pairs).  Each CTA stages its own 128-row A block (16 KB)
// and HALF of the 256-column B block (16 KB) per K = 32 stage; the leader CTA issues 6 x tcgen05.mma.cta_group::2
// (M = 256 over the pair, N = 256) + one multicast commit per stage; the peer forwards "my stage landed" to the leader
// with a remote mbarrier arrive.  Prints cycles per stage (all 74 clusters running).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint64_t* b, uint32_t par) {
  asm volatile("{\n.reg .pred P1;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DN;\nbra LW;\nDN:\n}" ::"r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return uint64_t((saddr >> 4) & 0x3FFFu) | (uint64_t((lbo >> 4) & 0x3FFFu) << 16) | (uint64_t((sbo >> 4) & 0x3FFFu) << 32) | (uint64_t(1) << 46);
}
__device__ __forceinline__ void mma2(uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
  constexpr uint32_t idesc = (1u << 4) | (uint32_t(256 >> 3) << 17) | (uint32_t(256 >> 4) << 24);   // M = 256, N = 256, f16 -> f32
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
constexpr int kBlk = 16384;
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
k(const char* src, size_t span, int S, int iters, long long* cycles) {
  extern __shared__ uint8_t sm_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + size_t(S) * 2 * kBlk);   // own copies landed
  uint64_t* empty = full + 8;                                                 // MMAs have read the stage (multicast commit)
  uint64_t* pfull = empty + 8;                                                // leader only: the peer's copies landed
  uint32_t* slot = reinterpret_cast<uint32_t*>(pfull + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_rank();
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&pfull[s])), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  const size_t base = (size_t(blockIdx.x) * 7919u * kBlk) % (span - 2 * kBlk);
  const long long t0 = clock64();
  if (warp == 2) {                       // producer (both CTAs): A block + B half
    if (lane < 2) {
      for (int i = 0; i < iters; ++i) {
        const int s = i % S;
        if (lane == 0) {
          mwait(&empty[s], ((i / S) & 1) ^ 1);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(2 * kBlk) : "memory");
        }
        __syncwarp(0x3);
        const size_t off = ((base + size_t(i) * 2 * kBlk + lane * kBlk) % (span - kBlk)) & ~size_t(15);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         s32(sm + size_t(s) * 2 * kBlk + lane * kBlk)), "l"(src + off), "r"(kBlk), "r"(s32(&full[s])) : "memory");
      }
    }
    __syncwarp();
  } else if (warp == 1 && rank == 1) {   // peer: forward "landed" to the leader's pfull[s]
    if (lane == 0) {
      for (int i = 0; i < iters; ++i) {
        const int s = i % S;
        mwait(&full[s], (i / S) & 1);
        uint32_t remote;
        asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(s32(&pfull[s])), "r"(0));
        asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
      }
    }
    __syncwarp();
  } else if (warp == 0 && rank == 0) {   // leader: MMA issue for the pair
    const uint64_t a_hi = desc(s32(sm), 2048, 128), a_lo = a_hi + (8192 >> 4);
    const uint64_t b_hi = desc(s32(sm) + kBlk, 4096, 128), b_lo = b_hi + (2048 >> 4);   // [chunk][hi|lo][128 rows][16 B]
    for (int i = 0; i < iters; ++i) {
      const int s = i % S;
      mwait(&full[s], (i / S) & 1);
      mwait(&pfull[s], (i / S) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint64_t so = uint64_t(s) * (2 * kBlk >> 4);
        mma2(tmem, a_hi + so, b_hi + so, i != 0);
        mma2(tmem + 256, a_lo + so, b_hi + so, i != 0);
        mma2(tmem + 256, a_hi + so, b_lo + so, 1);
        mma2(tmem, a_hi + so + 256, b_hi + so + 512, 1);
        mma2(tmem + 256, a_lo + so + 256, b_hi + so + 512, 1);
        mma2(tmem + 256, a_hi + so + 256, b_lo + so + 512, 1);
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         s32(&empty[s])), "h"((uint16_t)3) : "memory");
      }
      __syncwarp();
    }
  }
  // drain: everyone waits until the last stage has been released (the commit of the last MMAs has fired in both CTAs)
  if (warp == 3 && lane == 0) { const int i = iters - 1; mwait(&empty[i % S], (i / S) & 1); }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  cluster_sync();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}
int main() {
  const size_t span = size_t(64) << 20;
  char* src; cudaMalloc(&src, span); cudaMemset(src, 0, span);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int iters = 640;
  for (int S : {4, 6}) {
    for (int ctas : {2, 148}) {
      for (int rep = 0; rep < 2; ++rep) k<<<ctas, 128, S * 2 * kBlk + 512 + 1024>>>(src, span, S, iters, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
      double mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("cta_group::2  S=%d ctas=%3d : %6.0f clk/stage (M = 256 x N = 256 x K = 32 per stage and pair; %.1f B/clk/SM)\n", S, ctas,
             mx / iters, 2.0 * kBlk / (mx / iters));
    }
  }
  return 0;
}
