// kbs_net_tc.cu -- tcgen05 (5th-gen tensor core) datapath of the LSTM trunk, fp32-accurate by 3xTF32.
//
// One kernel = one LSTM layer-step for a 128-env tile x 64 hidden units (all four gates):
//     gates[128][256] = [x | h_prev][128][2H] . Wcat[256][2H]^T          (tcgen05.mma kind::tf32, fp32 accum in TMEM)
//     epilogue: + bias, c' = s(f) c + s(i) tanh(g), h' = s(o) tanh(c')   (eqx LSTMCell, train.py:889-893, 919)
// Every fp32 operand x is carried as a pair of TF32-representable planes hi = rn_tf32(x), lo = rn_tf32(x - hi) and the
// product is accumulated as hi.hi + hi.lo + lo.hi (the lo.lo term is below 2^-22 relative): three MMAs per K-step,
// |error| ~ fp32 rounding, which is what the 1e-5 parity bound needs.
//
// Operand layout ("SB" = split-blocked): 128-row panels (A) / 256-column gate tiles (B), K cut into blocks of 16; a
// block holds [hi|lo][4 k-chunks of 16 B][rows][4 floats] -- exactly the UMMA canonical K-major SWIZZLE_NONE layout
// (8-row x 16-byte core matrices, SBO = 128 B between 8-row groups, LBO = rows*16 B between k-chunks), so one
// cp.async.bulk per operand per stage lands it in shared memory ready for the tensor core, no tensor map needed.
// Weights are packed once (kbs_weights_pack); activations are written in SB form by the producing epilogue.
#include <math.h>

#include "kbs_common.cuh"

namespace {

constexpr int kPanelRows = 128;          // UMMA M
constexpr int kTileCols = 256;           // UMMA N: 64 hidden units x 4 gates (gate-interleaved)
constexpr int kUnitsPerTile = 64;
constexpr int kBlkK = 16;                // K per pipeline stage (2 MMA k-steps of 8)
constexpr int kStages = 4;
constexpr int kABlockFloats = 2 * 4 * kPanelRows * 4;   // 4096 floats = 16 KB
constexpr int kBBlockFloats = 2 * 4 * kTileCols * 4;    // 8192 floats = 32 KB
constexpr int kStageBytes = (kABlockFloats + kBBlockFloats) * 4;   // 48 KB
constexpr int kEpiWarps = 8;
constexpr int kThreadsTC = 64 + 32 * kEpiWarps;         // warp 0 = bulk-copy producer, warp 1 = MMA issuer
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*bias*/ + 256 /*barriers*/ + 1024 /*align*/;

// ---- PTX wrappers -----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor: addr>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type=0 [61,64)).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return uint64_t((saddr >> 4) & 0x3FFFu) | (uint64_t((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         (uint64_t((sbo_bytes >> 4) & 0x3FFFu) << 32) | (uint64_t(1) << 46);
}
// instruction descriptor: D = F32 (1<<4), A = B = TF32 (2<<7, 2<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(kTileCols >> 3) << 17) |
                            (uint32_t(kPanelRows >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- fp32 -> (hi, lo) TF32 planes -------------------------------------------------------------------------------
__device__ __forceinline__ float tf32_rn(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ void split4(const float (&x)[4], float4& hi, float4& lo) {
  float h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { h[i] = tf32_rn(x[i]); l[i] = tf32_rn(x[i] - h[i]); }
  hi = make_float4(h[0], h[1], h[2], h[3]);
  lo = make_float4(l[0], l[1], l[2], l[3]);
}
// SB address of (row, k) chunk start (k % 4 == 0) in a buffer of `kblocks` K-blocks per panel, rows per panel R.
template <int R>
__device__ __forceinline__ size_t sb_index(int64_t row, int k, int kblocks, int part) {
  const int64_t panel = row / R;
  const int r = int(row - panel * R);
  const int b = k >> 4, kc = (k >> 2) & 3;
  return (((size_t(panel) * kblocks + b) * 2 + part) * 4 + kc) * (size_t(R) * 4) + size_t(r) * 4;
}
template <int R>
__device__ __forceinline__ void sb_store4(float* __restrict__ sb, int64_t row, int k, int kblocks, const float (&x)[4]) {
  float4 hi, lo;
  split4(x, hi, lo);
  *reinterpret_cast<float4*>(sb + sb_index<R>(row, k, kblocks, 0)) = hi;
  *reinterpret_cast<float4*>(sb + sb_index<R>(row, k, kblocks, 1)) = lo;
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// ---- the layer kernel -----------------------------------------------------------------------------------------------
struct LayerArgs {
  const float* x_sb;      // [panels][H/16] A blocks: layer input
  const float* h_sb_in;   // [panels][H/16] A blocks: recurrent input (h_{t-1}, already reset where done_{t-1})
  const float* w_sb;      // [H/64 tiles][2H/16] B blocks
  const float* bias_t;    // [H/64][256] gate-interleaved bias
  float* c;               // [n][H] row-major cell carry in/out (reset where done)
  float* h_carry;         // [n][H] row-major hidden carry out (reset where done)
  float* h_sb_out;        // SB recurrent state out (reset where done); must NOT alias h_sb_in
  float* x_next_sb;       // SB input of the next layer (un-reset) or nullptr
  float* h_next_rm;       // [n][H] row-major un-reset output (last layer) or nullptr
  float* raw_gates;       // debug: [n][4H] pre-activation gates in eqx order (i,f,g,o), nullptr in production
  const uint8_t* done;    // [n] or nullptr
  int64_t n;
  int H;
};

__global__ void __launch_bounds__(kThreadsTC, 1) lstm_layer_tc_kernel(const LayerArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* stage_base = reinterpret_cast<float*>(smem);
  float* bias_s = reinterpret_cast<float*>(smem + kStages * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + 1024);
  uint64_t* full = bars;                 // [kStages]
  uint64_t* empty = bars + kStages;      // [kStages]
  uint64_t* acc_full = bars + 2 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int panel = blockIdx.x, tile = blockIdx.y;
  const int H = a.H;
  const int kb_half = H / kBlkK;         // K blocks of the x part (= of the h part)
  const int kb_total = 2 * kb_half;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {   // TMEM: 256 fp32 accumulator columns x 128 lanes
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(uint32_t(kTileCols)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < kTileCols; i += 32 * kEpiWarps) bias_s[i] = a.bias_t[size_t(tile) * kTileCols + i];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== producer: one bulk copy per operand per stage =====
      const float* xa = a.x_sb + size_t(panel) * kb_half * kABlockFloats;
      const float* ha = a.h_sb_in + size_t(panel) * kb_half * kABlockFloats;
      const float* wb = a.w_sb + size_t(tile) * kb_total * kBBlockFloats;
      for (int b = 0; b < kb_total; ++b) {
        const int s = b % kStages;
        mbar_wait(&empty[s], ((b / kStages) & 1) ^ 1);
        float* sa = stage_base + size_t(s) * (kStageBytes / 4);
        float* sb = sa + kABlockFloats;
        mbar_expect_tx(&full[s], kStageBytes);
        const float* asrc = (b < kb_half) ? xa + size_t(b) * kABlockFloats : ha + size_t(b - kb_half) * kABlockFloats;
        bulk_g2s(sa, asrc, kABlockFloats * 4, &full[s]);
        bulk_g2s(sb, wb + size_t(b) * kBBlockFloats, kBBlockFloats * 4, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer: 2 k-steps x (hi.hi + hi.lo + lo.hi) per stage =====
      for (int b = 0; b < kb_total; ++b) {
        const int s = b % kStages;
        mbar_wait(&full[s], (b / kStages) & 1);
        tc_fence_after();
        const uint32_t sa = smem_u32(stage_base + size_t(s) * (kStageBytes / 4));
        const uint32_t sb = sa + kABlockFloats * 4;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          // A: [part][kc][128 rows][16 B]: part stride 8 KB, kc stride 2 KB.   B: part 16 KB, kc 4 KB.
          const uint64_t a_hi = umma_desc(sa + ks * 4096, 2048, 128);
          const uint64_t a_lo = umma_desc(sa + 8192 + ks * 4096, 2048, 128);
          const uint64_t b_hi = umma_desc(sb + ks * 8192, 4096, 128);
          const uint64_t b_lo = umma_desc(sb + 16384 + ks * 8192, 4096, 128);
          umma_tf32(tmem_base, a_lo, b_hi, (b | ks) != 0);
          umma_tf32(tmem_base, a_hi, b_lo, 1);
          umma_tf32(tmem_base, a_hi, b_hi, 1);
        }
        umma_commit(&empty[s]);          // frees the stage when these MMAs have read it
      }
      umma_commit(acc_full);             // accumulator complete
    }
  } else {
    // ===== epilogue: 8 warps; warp%4 selects the TMEM lane quarter, (warp-2)/4 the half of the 64 units =====
    const int q4 = warp & 3, half = (warp - 2) >> 2;
    const int r = q4 * 32 + lane;
    const int64_t R = int64_t(panel) * kPanelRows + r;
    const bool live = R < a.n;
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const bool rst = live && a.done && a.done[R];
#pragma unroll 1
    for (int ch = 2 * half; ch < 2 * half + 2; ++ch) {
      const int uo = ch * 16;                              // unit offset inside the tile
      const uint32_t t0 = tmem_base + (uint32_t(q4 * 32) << 16) + uint32_t(uo);
      float gi[16], gf[16], gg[16], go[16];
      tmem_ld16(t0, gi);
      tmem_ld16(t0 + 64, gf);
      tmem_ld16(t0 + 128, gg);
      tmem_ld16(t0 + 192, go);
      tmem_ld_wait();
      const int u0 = tile * kUnitsPerTile + uo;            // first hidden unit of this chunk
      if (!live) continue;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        gi[i] += bias_s[uo + i]; gf[i] += bias_s[64 + uo + i]; gg[i] += bias_s[128 + uo + i]; go[i] += bias_s[192 + uo + i];
      }
      if (a.raw_gates) {
        float* g = a.raw_gates + R * 4 * H + u0;
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          *reinterpret_cast<float4*>(g + i) = make_float4(gi[i], gi[i + 1], gi[i + 2], gi[i + 3]);
          *reinterpret_cast<float4*>(g + H + i) = make_float4(gf[i], gf[i + 1], gf[i + 2], gf[i + 3]);
          *reinterpret_cast<float4*>(g + 2 * H + i) = make_float4(gg[i], gg[i + 1], gg[i + 2], gg[i + 3]);
          *reinterpret_cast<float4*>(g + 3 * H + i) = make_float4(go[i], go[i + 1], go[i + 2], go[i + 3]);
        }
        continue;
      }
      float* cp = a.c + R * H + u0;
      float* hp = a.h_carry + R * H + u0;
#pragma unroll
      for (int i = 0; i < 16; i += 4) {
        const float4 c4 = *reinterpret_cast<const float4*>(cp + i);
        const float cv[4] = {c4.x, c4.y, c4.z, c4.w};
        float hn[4], cn[4], hr[4], cr[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          cn[l] = sigmoidf_(gf[i + l]) * cv[l] + sigmoidf_(gi[i + l]) * tanhf(gg[i + l]);
          hn[l] = sigmoidf_(go[i + l]) * tanhf(cn[l]);
          hr[l] = rst ? 0.0f : hn[l];
          cr[l] = rst ? 0.0f : cn[l];
        }
        *reinterpret_cast<float4*>(cp + i) = make_float4(cr[0], cr[1], cr[2], cr[3]);
        *reinterpret_cast<float4*>(hp + i) = make_float4(hr[0], hr[1], hr[2], hr[3]);
        sb_store4<kPanelRows>(a.h_sb_out, R, u0 + i, kb_half, hr);
        if (a.x_next_sb) sb_store4<kPanelRows>(a.x_next_sb, R, u0 + i, kb_half, hn);
        if (a.h_next_rm) *reinterpret_cast<float4*>(a.h_next_rm + R * H + u0 + i) = make_float4(hn[0], hn[1], hn[2], hn[3]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(kTileCols)));
  }
}

// ---- packing kernels ------------------------------------------------------------------------------------------------
// eqx LSTMCell weights [4H][H] x2 + bias [4H]  ->  gate-interleaved SB tiles (hi/lo) + interleaved bias.
__global__ void __launch_bounds__(256)
pack_lstm_weights_kernel(const float* __restrict__ w_ih, const float* __restrict__ w_hh, const float* __restrict__ b,
                         float* __restrict__ w_sb, float* __restrict__ bias_t, int H) {
  const int kq = 2 * H / 4;                               // 16-byte chunks along K
  const int64_t idx = int64_t(blockIdx.x) * 256 + threadIdx.x;
  const int64_t total = int64_t(4 * H) * kq;
  if (idx >= total) return;
  const int col_g = int(idx / kq);                        // global packed column: tile * 256 + c
  const int k = int(idx % kq) * 4;
  const int tile = col_g / kTileCols, c = col_g % kTileCols;
  const int gate = c / kUnitsPerTile, u = tile * kUnitsPerTile + c % kUnitsPerTile;
  const int row = gate * H + u;                           // eqx row (i,f,g,o blocks of H)
  const float* src = (k < H) ? w_ih + size_t(row) * H + k : w_hh + size_t(row) * H + (k - H);
  const float x[4] = {src[0], src[1], src[2], src[3]};
  sb_store4<kTileCols>(w_sb, col_g, k, 2 * H / kBlkK, x);
  if (k == 0) bias_t[col_g] = b[row];
}

// row-major [n][K] fp32 -> SB (A operand, 128-row panels).  Rows >= n of the last panel are zero-filled.
__global__ void __launch_bounds__(256)
pack_rows_sb_kernel(const float* __restrict__ src, int64_t ld, float* __restrict__ sb, int64_t n, int64_t n_pad, int K) {
  const int kq = K / 4;
  const int64_t idx = int64_t(blockIdx.x) * 256 + threadIdx.x;
  if (idx >= n_pad * kq) return;
  // consecutive threads -> consecutive rows of the same k-chunk: coalesced SB stores
  const int64_t panel = idx / (int64_t(kPanelRows) * kq);
  const int rem = int(idx - panel * int64_t(kPanelRows) * kq);
  const int kc = rem / kPanelRows, r = rem % kPanelRows;
  const int64_t row = panel * kPanelRows + r;
  float x[4] = {0.f, 0.f, 0.f, 0.f};
  if (row < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + row * ld + kc * 4);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
  }
  sb_store4<kPanelRows>(sb, row, kc * 4, K / kBlkK, x);
}

inline int64_t pad_rows(int64_t n) { return (n + kPanelRows - 1) / kPanelRows * kPanelRows; }

}  // namespace

// ---- host side --------------------------------------------------------------------------------------------------------
// tc_image of one net: per layer [w_sb (4H x 2H x 2 floats) | bias_t (4H floats)]
static size_t layer_image_floats(int H) { return size_t(4 * H) * (2 * H) * 2 + size_t(4 * H); }

int kbs_tc_pack(kbs_handle* h, int net, cudaStream_t st) {
  KbsNet& N = h->net[net];
  const int H = h->p.hidden_size;
  if (H % kUnitsPerTile) return KBS_E_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    KBS_CUDA_TRY(cudaFuncSetAttribute(lstm_layer_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_set = true;
  }
  const size_t per_layer = layer_image_floats(H);
  if (N.tc_image) { KBS_CUDA_TRY(cudaFree(N.tc_image)); N.tc_image = nullptr; }
  N.tc_image_floats = per_layer * h->p.depth;
  KBS_CUDA_TRY(cudaMalloc(&N.tc_image, N.tc_image_floats * sizeof(float)));
  for (int l = 0; l < h->p.depth; ++l) {
    float* w_sb = N.tc_image + per_layer * l;
    float* bias_t = w_sb + size_t(4 * H) * (2 * H) * 2;
    const int64_t total = int64_t(4 * H) * (2 * H / 4);
    KBS_LAUNCH(h, KBS_K_PACK, st,
               (pack_lstm_weights_kernel<<<unsigned((total + 255) / 256), 256, 0, st>>>(N.w_ih[l], N.w_hh[l], N.b[l], w_sb,
                                                                                      bias_t, H)));
  }
  KBS_LAUNCH_CHECK();
  return KBS_OK;
}

// scratch (floats) the TC trunk needs for n envs: x_sb ping/pong + h_sb per layer + x row-major + hbuf row-major
size_t kbs_tc_scratch_floats(const kbs_handle* h, int64_t n) {
  const size_t H = size_t(h->p.hidden_size), np = size_t(pad_rows(n));
  return np * H * 2 * (2 + 2 * size_t(h->p.depth)) + 2 * size_t(n) * H + 256;
}

// Runs depth LSTM layers on the tensor cores.  x_rm: [n][H] row-major layer-0 input (input_proj output);
// carry: ABI layout [depth][2][n][H]; out_h_rm: [n][H] row-major un-reset top-layer output.  ws: kbs_tc_scratch_floats.
int kbs_tc_lstm_stack(kbs_handle* h, int net, const float* x_rm, float* carry, const uint8_t* done, float* out_h_rm,
                      float* ws, int64_t n, bool carry_sb_valid, cudaStream_t st) {
  const KbsNet& N = h->net[net];
  if (!N.packed || !N.tc_image) return KBS_E_STATE;
  const int H = h->p.hidden_size, depth = h->p.depth;
  const int64_t np = pad_rows(n);
  const size_t sbf = size_t(np) * H * 2;
  float* x_sb[2] = {ws, ws + sbf};
  float* h_sb = ws + 2 * sbf;   // [depth][2 (in, out)] x sbf: the 4 column-tile CTAs of a panel all read h_{t-1}
                                // while their epilogues write h_t, so in and out must be distinct buffers
  const int64_t chunks = np * (H / 4);
  KBS_LAUNCH(h, KBS_K_PACK, st,
             (pack_rows_sb_kernel<<<unsigned((chunks + 255) / 256), 256, 0, st>>>(x_rm, H, x_sb[0], n, np, H)));
  if (!carry_sb_valid) {
    for (int l = 0; l < depth; ++l) {
      const float* ch = carry + (size_t(l) * 2 + 0) * size_t(n) * H;
      KBS_LAUNCH(h, KBS_K_PACK, st,
                 (pack_rows_sb_kernel<<<unsigned((chunks + 255) / 256), 256, 0, st>>>(ch, H, h_sb + sbf * (2 * l), n, np, H)));
    }
  }
  const size_t per_layer = layer_image_floats(H);
  for (int l = 0; l < depth; ++l) {
    LayerArgs a{};
    a.x_sb = x_sb[l & 1];
    a.h_sb_in = h_sb + sbf * (2 * l);
    a.w_sb = N.tc_image + per_layer * l;
    a.bias_t = a.w_sb + size_t(4 * H) * (2 * H) * 2;
    a.c = carry + (size_t(l) * 2 + 1) * size_t(n) * H;
    a.h_carry = carry + (size_t(l) * 2 + 0) * size_t(n) * H;
    a.h_sb_out = h_sb + sbf * (2 * l + 1);
    a.x_next_sb = (l + 1 < depth) ? x_sb[(l + 1) & 1] : nullptr;
    a.h_next_rm = (l + 1 < depth) ? nullptr : out_h_rm;
    a.raw_gates = nullptr;
    a.done = done;
    a.n = n;
    a.H = H;
    dim3 grid(unsigned(np / kPanelRows), unsigned(H / kUnitsPerTile));
    KBS_LAUNCH(h, KBS_K_LSTM_TC, st, (lstm_layer_tc_kernel<<<grid, kThreadsTC, kSmemBytes, st>>>(a)));
  }
  KBS_LAUNCH_CHECK();
  return KBS_OK;
}

// Debug / test entry: raw pre-activation gates (eqx order i,f,g,o; bias added) of one layer from row-major inputs.
int kbs_tc_debug_gates(kbs_handle* h, int net, int layer, const float* x_rm, const float* h_rm, float* gates_out,
                       float* ws, int64_t n, cudaStream_t st) {
  const KbsNet& N = h->net[net];
  if (!N.packed || !N.tc_image) return KBS_E_STATE;
  const int H = h->p.hidden_size;
  const int64_t np = pad_rows(n);
  const size_t sbf = size_t(np) * H * 2;
  const int64_t chunks = np * (H / 4);
  KBS_LAUNCH(h, KBS_K_PACK, st, (pack_rows_sb_kernel<<<unsigned((chunks + 255) / 256), 256, 0, st>>>(x_rm, H, ws, n, np, H)));
  KBS_LAUNCH(h, KBS_K_PACK, st,
             (pack_rows_sb_kernel<<<unsigned((chunks + 255) / 256), 256, 0, st>>>(h_rm, H, ws + sbf, n, np, H)));
  const size_t per_layer = layer_image_floats(H);
  LayerArgs a{};
  a.x_sb = ws;
  a.h_sb_in = ws + sbf;
  a.w_sb = N.tc_image + per_layer * layer;
  a.bias_t = a.w_sb + size_t(4 * H) * (2 * H) * 2;
  a.raw_gates = gates_out;
  a.n = n;
  a.H = H;
  dim3 grid(unsigned(np / kPanelRows), unsigned(H / kUnitsPerTile));
  KBS_LAUNCH(h, KBS_K_LSTM_TC, st, (lstm_layer_tc_kernel<<<grid, kThreadsTC, kSmemBytes, st>>>(a)));
  KBS_LAUNCH_CHECK();
  return KBS_OK;
}
