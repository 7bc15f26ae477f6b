// Microbenchmark 4: the operand pipeline of rollout_persist_kernel in isolation (no epilogue, no dependencies):
// two producer warps (lanes 0/1 = 16 KB activation block + 16 KB weight block per stage) -> ring of S stages ->
// one consumer warp that (mode 0) frees the stage at once, (mode 1) issues the stage's four tcgen05.mma + commit,
// (mode 2) issues only the commit.  Prints cycles per stage with all 148 SMs running.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mwait(uint64_t* b, uint32_t par) {
  asm volatile("{\n.reg .pred P1;\nLW:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DN;\nbra LW;\nDN:\n}" ::"r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ uint64_t desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return uint64_t((saddr >> 4) & 0x3FFFu) | (uint64_t((lbo >> 4) & 0x3FFFu) << 16) | (uint64_t((sbo >> 4) & 0x3FFFu) << 32) | (uint64_t(1) << 46);
}
template <int N>
__device__ __forceinline__ void mma(uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
  constexpr uint32_t idesc = (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(128 >> 4) << 24);
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
constexpr int kBlk = 16384;
__global__ void __launch_bounds__(128, 1) k(const char* src, size_t span, int S, int iters, int mode, int a_bytes, int b_bytes,
                                            int hold, long long* cycles) {
  extern __shared__ uint8_t sm_raw[];
  uint8_t* sm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sm_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(sm + size_t(S) * 2 * kBlk);
  uint64_t* empty = full + 8;
  uint32_t* slot = reinterpret_cast<uint32_t*>(empty + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&full[s])), "r"(1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(1));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(slot)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *slot;
  const size_t base = (size_t(blockIdx.x) * 7919u * kBlk) % (span - 2 * kBlk);
  const long long t0 = clock64();
  if (warp >= 2) {                       // producers
    if (lane < 2) {
      for (int i = warp - 2; i < iters; i += 2) {
        const int s = i % S;
        if (lane == 0) {
          mwait(&empty[s], ((i / S) & 1) ^ 1);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(a_bytes + b_bytes) : "memory");
        }
        __syncwarp(0x3);
        const size_t off = ((base + size_t(i) * 2 * kBlk + lane * kBlk) % (span - kBlk)) & ~size_t(15);
        const int bytes = lane == 0 ? a_bytes : b_bytes;
        if (bytes)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           s32(sm + size_t(s) * 2 * kBlk + lane * kBlk)), "l"(src + off), "r"(bytes), "r"(s32(&full[s])) : "memory");
      }
    }
    __syncwarp();
  } else if (warp == 0) {                // consumer
    const uint64_t a_hi = desc(s32(sm), 2048, 128), a_lo = a_hi + (8192 >> 4), b_all = desc(s32(sm) + kBlk, 4096, 128);
    for (int i = 0; i < iters; ++i) {
      const int s = i % S;
      mwait(&full[s], (i / S) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint64_t so = uint64_t(s) * (2 * kBlk >> 4);
        if (mode == 0) {
          if (hold) { const long long h0 = clock64(); while (clock64() - h0 < hold) { } }
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
        } else {
          if (mode == 3) {              // N = 256 tile: 3 x N = 256 per k-step (main 0..255, correction 256..511)
            mma<256>(tmem, a_hi + so, b_all + so, i != 0);
            mma<256>(tmem + 256, a_lo + so, b_all + so, i != 0);
            mma<256>(tmem + 256, a_hi + so, b_all + so, 1);
            mma<256>(tmem, a_hi + so + 256, b_all + so + 512, 1);
            mma<256>(tmem + 256, a_lo + so + 256, b_all + so + 512, 1);
            mma<256>(tmem + 256, a_hi + so + 256, b_all + so + 512, 1);
          }
          if (mode == 4) {              // two k-steps of 2 x N = 256 (issue-rate probe)
            mma<256>(tmem, a_hi + so, b_all + so, i != 0);
            mma<256>(tmem + 256, a_lo + so, b_all + so, i != 0);
            mma<256>(tmem, a_hi + so + 256, b_all + so + 512, 1);
            mma<256>(tmem + 256, a_lo + so + 256, b_all + so + 512, 1);
          }
          if (mode == 1) {
            mma<256>(tmem, a_hi + so, b_all + so, i != 0);
            mma<128>(tmem + 128, a_lo + so, b_all + so, 1);
            mma<256>(tmem, a_hi + so + 256, b_all + so + 512, 1);
            mma<128>(tmem + 128, a_lo + so + 256, b_all + so + 512, 1);
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&empty[s])) : "memory");
        }
      }
      __syncwarp();
    }
    // drain: wait until the last stage's commit has fired
    if (mode != 0) { const int i = iters - 1; mwait(&empty[i % S], (i / S) & 1); }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}
int main() {
  const size_t span = size_t(64) << 20;
  char* src; cudaMalloc(&src, span); cudaMemset(src, 0, span);
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  long long h[148];
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  const int iters = 640;
  struct Cfg { int S, mode, a, b, hold; const char* what; };
  const Cfg cfgs[] = {
      {4, 0, 16384, 16384, 0, "free at once"},       {6, 0, 16384, 16384, 0, "free at once"},
      {4, 0, 16384, 16384, 400, "hold 400 clk"},     {6, 0, 16384, 16384, 400, "hold 400 clk"},
      {4, 2, 16384, 16384, 0, "commit only"},        {6, 2, 16384, 16384, 0, "commit only"},
      {4, 1, 16384, 16384, 0, "4 MMAs + commit"},    {6, 1, 16384, 16384, 0, "4 MMAs + commit"},
      {6, 1, 16384, 16, 0, "4 MMAs + commit, token weight copy"}, {6, 1, 8192, 0, 0, "4 MMAs + commit, 8 KB per stage"},
      {6, 1, 0, 0, 0, "4 MMAs + commit, no copies (expect_tx 0)"},
      {4, 3, 0, 0, 0, "6 x N=256 MMAs + commit, no copies"}, {4, 4, 0, 0, 0, "4 x N=256 MMAs + commit, no copies"},
      {4, 3, 16384, 16384, 0, "6 x N=256 MMAs + commit, 32 KB per stage"},
  };
  for (const Cfg& c : cfgs) {
    for (int ctas : {1, 148}) {
      for (int rep = 0; rep < 2; ++rep) k<<<ctas, 128, c.S * 2 * kBlk + 256 + 1024>>>(src, span, c.S, iters, c.mode, c.a, c.b, c.hold, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, cyc, ctas * 8, cudaMemcpyDeviceToHost);
      double mx = 0; for (int i = 0; i < ctas; ++i) mx = h[i] > mx ? h[i] : mx;
      printf("S=%d ctas=%3d %-44s : %6.0f clk/stage  (%.1f B/clk/SM)\n", c.S, ctas, c.what, mx / iters, (c.a + c.b) / (mx / iters));
    }
  }
  return 0;
}
