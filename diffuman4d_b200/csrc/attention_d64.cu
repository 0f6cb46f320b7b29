// tcgen05 flash-attention forward for head_dim 64 on sm_100a: the kernel behind every attention call of the SD-2.1
// channel layout (5/10/20/20 heads of 64), i.e. the 3-D attention over all F frames of a CFG half (reference
// src/diffusers/models/attention.py:68-83, F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0) and the
// per-image 2-D attention of level 0.  Same operand contract as attention_umma.cu (Q/K/V are column slices of the fused
// QKV-GEMM output, read in place through strided TMA boxes).
//
// Why a second kernel: at head_dim 64 every exponential carries only 4*64 FLOP, so the softmax - not the tensor pipe -
// bounds the kernel (MUFU.EX2 runs at 16/clk/SM against 32/clk needed; profiles/README.md).  The generic kernel spends
// about half of its time on per-tile fixed costs and keeps only two softmax warps per sub-partition.  This kernel is
// organised around the softmax instead:
//
//   * one CTA per SM owning NQ Q tiles of 128 rows and walking K/V tiles of KV keys (configurations <NQ, KV> = <2, 128>
//     and <3, 64>); every K/V tile is staged once for all Q tiles;
//   * 4 softmax warps per Q tile (thread = query row = TMEM lane), i.e. NQ warps per sub-partition.  The KV scores of a
//     tile are pulled into registers chunk by chunk while the previous tile's P is being stored, which frees S for
//     Q.K^T(j+1) at once: registers are the second S buffer, so the tensor pipe runs one tile ahead of the softmax;
//   * the MMA warp serves the Q tiles independently, whichever hands a tile over first, so that their softmax warps do
//     not run in lock-step;
//   * exp2 is split between the MUFU pipe and a Cody-Waite / degree-3 polynomial evaluated with packed FFMA2 / FADD2 on
//     the FMA pipe (round-down magic-number split, exponent patched in with one IMAD); the share is a compile-time
//     pattern over pairs of columns, 8.8e-5 relative error, far below the bf16 rounding of P;
//   * the running max is lazy: it only moves when the row max grows by more than 8 (log2 units), so O is rescaled (by
//     the softmax warps themselves, in TMEM) a handful of times per row.
//
// TMEM (512 columns), per Q tile q at column q * (KV + KV/2 + 64): S_q fp32 [+0, KV) | P_q bf16 [+KV, KV/2) | O_q fp32 [.., 64).
// Warps: 4q .. 4q+3 softmax of Q tile q, 4 NQ = MMA issuer (+ TMEM owner), 4 NQ + 1 = TMA producer.  No call is made
// while a softmax thread holds its S row (trap-only barrier wait), which keeps the row in registers.
// Hand-offs: tcgen05.commit -> mbarrier (s_full[q], pv_done[q]) towards the softmax warps; one mbarrier per Q tile
// ("P(j) stored and S(j+1) in registers") towards the MMA warp, which then issues P.V(j) and Q.K^T(j+2) for that Q tile.
#include <math.h>

#include "kernels.h"

namespace d4d {

namespace {

constexpr int QT_ROWS = 128;  // rows of one Q tile (= TMEM lanes)
constexpr int TMEM_COLS = 512;

// kernel shape: NQ Q tiles per CTA, KV keys per K/V tile
template <int NQ_, int KV_>
struct Shape {
  static constexpr int NQ = NQ_, KV = KV_;
  static constexpr int CTA_ROWS = NQ * QT_ROWS;
  static constexpr int Q_BYTES = QT_ROWS * 64 * 2;    // one [128 rows][64 ch] 128B-swizzled Q box
  static constexpr int TILE_BYTES = KV * 64 * 2;      // one K or V tile
  static constexpr int SLOTS = 128 * 1024 / TILE_BYTES;  // K/V ring: 128 KB
  static constexpr int THREADS = (4 * NQ + 2) * 32;
  static constexpr int SMEM = NQ * Q_BYTES + SLOTS * TILE_BYTES + 1024 + 512;
  static constexpr int Q_STRIDE = KV + KV / 2 + 64;   // TMEM columns per Q tile
  static constexpr int COL_P = KV, COL_O = KV + KV / 2;
  static constexpr int CHUNKS = KV / 32;
  static_assert(NQ * Q_STRIDE <= TMEM_COLS, "TMEM budget");
  static_assert(KV % 32 == 0 && SLOTS >= 6, "tile shape");
};

struct A64Args {
  int seq_q, seq_kv, heads, n_kv_tiles;
  float scale_log2;
  bf16* out;
  int ld_out;
  unsigned long long* trace;  // tools build: phase timestamps of the softmax warps of CTA (0, 0) (tools/trace_attention.py)
};

#ifdef D4D_ATTN_TRACE  // phase timeline build (tools/trace_attention.py): costs registers, so not even in the tools build
#define D4D_TRACE(ev) do { if (tr) tr[(j * 8 + (ev))] = clock64(); } while (0)
#else
#define D4D_TRACE(ev) do { } while (0)
#endif

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2_rm(uint64_t a, uint64_t b) {  // packed add, round towards -inf
  uint64_t d;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// position of a K/V tile in the ring: the producer loads K0 K1 V0 K2 V1 K3 ... (the order a Q tile consumes them)
// = K0 K1, then for every t: V(t), K(t+2) while K(t+2) exists: the last V follows its predecessor directly, and a single-tile
// problem has no K1
__device__ __forceinline__ int ord_k(int t) { return t < 2 ? t : 2 * t - 1; }
__device__ __forceinline__ int ord_v(int t, int n_tiles) { return n_tiles == 1 ? 1 : 2 + t + min(t, n_tiles - 2); }

// ---- exp2 / packing variants (tools build instantiates all of them for A/B timing; the product uses the defaults) ----
// poly_mask: bit p = the p-th pair (of every 8 consecutive column pairs) takes the FMA-pipe exp2.
// trunc:     P is packed to bf16x2 by TRUNCATION (one PRMT on the ALU pipe) instead of F2FP.  Truncation alone would bias
//            P low by E[frac ulp] = 2^-8 E[1/mantissa] = 0.28 %; the exponent argument is therefore shifted by
//            log2(1.0028) so that E[trunc(P')] = P, and the row sum (taken from the un-truncated P') is divided by the
//            same factor in the epilogue.  |error| <= 0.28 % per element against 0.2 % of round-to-nearest.
// (Measured and dropped, profiles/README.md: integer round-half-up packing on the ALU pipe or with IMAD on the FMA pipe,
//  a fused round-down FFMA2 for the polynomial's floor, keeping P in registers until chunk 1 / the end of the tile,
//  shared-memory hand-off counters and overlapped barrier probes in the MMA warp, a head start for one Q tile.)
struct Variant { uint32_t poly_mask; bool trunc; };
__host__ __device__ constexpr Variant variant_of(int v) {
  return v == 0 ? Variant{0x00, false}
       : v == 1 ? Variant{0x10, false}   // 1/8 poly
       : v == 2 ? Variant{0x44, false}   // 2/8 poly
       : v == 3 ? Variant{0x00, true}
       : v == 4 ? Variant{0x10, true}
       : v == 5 ? Variant{0x44, true}
       :          Variant{0x92, true};   // 6: 3/8 poly
}
constexpr int kNumVariants = 7;
#ifndef D4D_ATTN64_VARIANT
#define D4D_ATTN64_VARIANT 1
#endif
#ifndef D4D_ATTN64_SHAPE   // 0: <2, 128>   1: <3, 64>
#define D4D_ATTN64_SHAPE 0
#endif

// mean relative truncation error of a bf16 with log-uniform mantissa = 0.5 ulp * E[1/mantissa] = 2^-8 / (2 ln 2)
__host__ __device__ constexpr float trunc_delta(const Variant& v) { return v.trunc ? 0.0028180f : 0.f; }

// exp2 of one pair of columns of this thread's row -> packed bf16x2 (lo = first column), row-sum share into lsum.
//   x = s * scale_log2 - m  (one FFMA2).  MUFU path: ex2.approx.  Polynomial path (x <= ~100 by construction of m):
//   xc = max(x, -126); t = xc + 1.5*2^23 rounded DOWN, so the low mantissa bits of t hold floor(xc); r = xc - floor(xc)
//   in [0, 1); 2^r ~ 1 + r (c1 + r (c2 + r c3)); result bits = (t << 23) + bits(2^r)  (exponent += floor(xc)).
template <bool kPoly, bool kTrunc, bool kMasked>
__device__ __forceinline__ uint32_t exp_pair(uint32_t s0, uint32_t s1, uint64_t sc2, uint64_t nm2, uint64_t& lsum, int col,
                                             int valid) {
  float x0, x1, e0, e1;
  f2_unpack(f2_fma(f2_pack(__uint_as_float(s0), __uint_as_float(s1)), sc2, nm2), x0, x1);
  if (!kPoly) {
    e0 = ex2f(x0);
    e1 = ex2f(x1);
  } else {
    const uint64_t xc = f2_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
    const uint64_t t = add2_rm(xc, f2_splat(12582912.f));
    const uint64_t r = sub2(xc, f2_add(t, f2_splat(-12582912.f)));  // xc - floor(xc); the inner add is exact
    uint64_t p = f2_fma(f2_splat(0.077119089663028717041015625f), r, f2_splat(0.227564394474029541015625f));
    p = f2_fma(p, r, f2_splat(0.695146143436431884765625f));
    p = f2_fma(p, r, f2_splat(1.0f));
    float t0, t1, p0, p1;
    f2_unpack(t, t0, t1);
    f2_unpack(p, p0, p1);
    e0 = __uint_as_float((__float_as_uint(t0) << 23) + __float_as_uint(p0));
    e1 = __uint_as_float((__float_as_uint(t1) << 23) + __float_as_uint(p1));
  }
  if (kMasked) {
    if (col >= valid) e0 = 0.f;
    if (col + 1 >= valid) e1 = 0.f;
  }
  lsum = f2_add(lsum, f2_pack(e0, e1));
  if (kTrunc) return __byte_perm(__float_as_uint(e0), __float_as_uint(e1), 0x7632);
  return pack_bf16x2(e0, e1);
}

// compile-time loop over the 16 column pairs of chunk C (the exp2 flavour of a pair is a constant of its index)
template <int VAR, bool kMasked, int KV, int C, int P = 0>
struct ExpChunk {
  static __device__ __forceinline__ void run(const uint32_t (&sv)[KV], uint32_t (&pk)[16], uint64_t sc2, uint64_t nm2,
                                             uint64_t (&lsum)[2], int valid) {
    constexpr Variant V = variant_of(VAR);
    constexpr int gp = C * 16 + P;
    constexpr bool poly = !kMasked && ((V.poly_mask >> (gp & 7)) & 1u);
    pk[P] = exp_pair<poly, V.trunc, kMasked>(sv[2 * gp], sv[2 * gp + 1], sc2, nm2, lsum[P & 1], 2 * gp, valid);
    if constexpr (P + 1 < 16) ExpChunk<VAR, kMasked, KV, C, P + 1>::run(sv, pk, sc2, nm2, lsum, valid);
  }
};

template <int VAR, int NQ, int KV>
__global__ void __launch_bounds__(Shape<NQ, KV>::THREADS, 1)
attn64_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                  const __grid_constant__ CUtensorMap tmap_v, const A64Args a) {
  using S = Shape<NQ, KV>;
  constexpr Variant V = variant_of(VAR);
  constexpr int SLOTS = S::SLOTS, TILE_BYTES = S::TILE_BYTES;
  constexpr int MMA_WARP = 4 * NQ, TMA_WARP = 4 * NQ + 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                        // NQ Q tiles
  uint8_t* sRing = smem + NQ * S::Q_BYTES;   // K/V ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + SLOTS * TILE_BYTES);
  uint64_t* ring_full = bars;           // [SLOTS]
  uint64_t* ring_empty = bars + SLOTS;  // [SLOTS]
  uint64_t* q_full = bars + 2 * SLOTS;
  uint64_t* s_full = q_full + 1;           // [NQ]  Q.K^T(q, j) complete (also: every MMA issued before it)
  uint64_t* pv_done = s_full + NQ;         // [NQ]  P.V(q, j) complete
  uint64_t* tile_done = pv_done + NQ;      // [NQ]  softmax warps of Q tile q: "P_q(j) stored, S_q(j+1) in registers" (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tile_done + NQ);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // role branches are warp-uniform
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / a.heads;
  const int hd = bh - b * a.heads;
  const int q_row0 = b * a.seq_q + blockIdx.x * S::CTA_ROWS;  // first query-token row of this CTA
  const int kv_row0 = b * a.seq_kv;                           // first key/value-token row of this batch entry
  const int col0 = hd * 64;                                   // first column of this head inside the q/k/v slice
  const int n_tiles = a.n_kv_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SLOTS; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], 1);
    }
    mbar_init(q_full, 1);
    for (int i = 0; i < NQ; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&pv_done[i], 1);
      mbar_init(&tile_done[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();
  pdl_launch_dependents();

  if (warp == TMA_WARP) {
    // ============================ TMA producer ============================
    if (elect_one()) {
      tma_prefetch_desc(&tmap_q);
      tma_prefetch_desc(&tmap_k);
      tma_prefetch_desc(&tmap_v);
      mbar_expect_tx(q_full, NQ * S::Q_BYTES);
#pragma unroll
      for (int q = 0; q < NQ; ++q) tma_load_2d(sQ + q * S::Q_BYTES, &tmap_q, q_full, col0, q_row0 + q * QT_ROWS);
    }
    __syncwarp();
    int slot = 0;
    uint32_t phase = 0;
    auto load_tile = [&](const CUtensorMap* tm, int j) {
      mbar_wait(&ring_empty[slot], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&ring_full[slot], TILE_BYTES);
        tma_load_2d(sRing + slot * TILE_BYTES, tm, &ring_full[slot], col0, kv_row0 + j * KV);
      }
      __syncwarp();
      if (++slot == SLOTS) { slot = 0; phase ^= 1; }
    };
    // the order a Q tile consumes them: K0 K1, then for every j: V(j), K(j+2)
    load_tile(&tmap_k, 0);
    if (n_tiles > 1) load_tile(&tmap_k, 1);
    for (int j = 0; j < n_tiles; ++j) {
      load_tile(&tmap_v, j);
      if (j + 2 < n_tiles) load_tile(&tmap_k, j + 2);
    }
  } else if (warp == MMA_WARP) {
    // ============================ MMA issuer ============================
    // whole warp in uniform control flow, one elected lane around the asynchronous instructions (see gemm_umma.cu)
    const uint32_t idesc_qk = make_idesc_bf16(QT_ROWS, KV, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(QT_ROWS, 64, 0, 1);
    const uint32_t q_addr = smem_u32(sQ);
    const uint32_t ring_addr = smem_u32(sRing);
    auto wait_tile = [&](int ord) {  // K/V tile at ring position `ord` has landed
      mbar_wait(&ring_full[ord % SLOTS], (ord / SLOTS) & 1);
      return ord % SLOTS;
    };
    auto issue_qk = [&](int q, int sl) {  // S_q = Q_q K^T : 4 x (M128 N=KV K16), both operands K-major SW128
      const uint32_t kaddr = ring_addr + sl * TILE_BYTES;
      const uint32_t s_tmem = tmem + q * S::Q_STRIDE;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = make_smem_desc(q_addr + q * S::Q_BYTES + k * 32, 0, 1024, 2);
        const uint64_t bd = make_smem_desc(kaddr + k * 32, 0, 1024, 2);
        umma_ss(s_tmem, ad, bd, idesc_qk, k != 0 ? 1u : 0u);
      }
    };
    auto issue_pv = [&](int q, int sl, int j) {  // O_q += P_q V : KV/16 x (M128 N64 K16), A = P bf16 in TMEM, B = V MN-major
      const uint32_t vaddr = ring_addr + sl * TILE_BYTES;
      const uint32_t p_tmem = tmem + q * S::Q_STRIDE + S::COL_P;
      const uint32_t o_tmem = tmem + q * S::Q_STRIDE + S::COL_O;
#pragma unroll
      for (int k = 0; k < KV / 16; ++k) {
        // 16 keys = two 8-row swizzle atoms = 2048 bytes; SBO = 1024 (next 8 keys); LBO unused (one 64-wide d block)
        const uint64_t bd = make_smem_desc(vaddr + k * 2048, TILE_BYTES, 1024, 2);
        umma_ts(o_tmem, p_tmem + k * 8, bd, idesc_pv, (j | k) != 0 ? 1u : 0u);
      }
    };
    mbar_wait(q_full, 0);
    {  // S_q(0) for every Q tile
      const int k_slot = wait_tile(ord_k(0));
      if (elect_one()) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          issue_qk(q, k_slot);
          umma_commit(&s_full[q]);
        }
        umma_commit(&ring_empty[k_slot]);
      }
      __syncwarp();
    }
    // The Q tiles are served INDEPENDENTLY, whichever hands a tile over first (a fixed order would lock their softmax
    // warps - which share the sub-partitions - into the same phase: all in the exp2 phase fighting for the MUFU pipe, then
    // all in the max / TMEM phase leaving it idle).  Step j of Q tile q (j = -1 .. n-1): issue P.V(q, j) (j >= 0) and
    // Q.K^T(q, j+2) (j+2 < n).  A K/V ring slot is released by whichever Q tile uses it last.
    int next[NQ];  // next step of each Q tile
#pragma unroll
    for (int q = 0; q < NQ; ++q) next[q] = -1;
    auto tile_landed = [&](int ord) {  // non-blocking
      return __all_sync(0xffffffffu, mbar_test(&ring_full[ord % SLOTS], (ord / SLOTS) & 1));
    };
    for (;;) {
      bool progress = false, all_done = true;
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int j = next[q];
        if (j >= n_tiles) continue;
        all_done = false;
        const bool has_pv = j >= 0, has_qk = j + 2 < n_tiles;
        // never block on the ring here: a slot this Q tile waits for may only be refilled after ANOTHER Q tile has been
        // served (it releases the slot's previous occupant)
        if (!__all_sync(0xffffffffu, mbar_test(&tile_done[q], (j + 1) & 1))) continue;
        if (has_pv && !tile_landed(ord_v(j, n_tiles))) continue;
        if (has_qk && !tile_landed(ord_k(j + 2))) continue;
        progress = true;
        tc_fence_after();  // P_q(j) was written / S_q(j+1) was read with tcgen05.st / .ld by the softmax warps
        bool last = true;  // every other Q tile has already taken this step
#pragma unroll
        for (int o = 0; o < NQ; ++o) last = last && (o == q || next[o] > j);
        const int v_slot = has_pv ? ord_v(j, n_tiles) % SLOTS : 0;
        const int k_slot = has_qk ? ord_k(j + 2) % SLOTS : 0;
        if (elect_one()) {
          if (has_pv) {
            issue_pv(q, v_slot, j);
            umma_commit(&pv_done[q]);
          }
          if (has_qk) {
            issue_qk(q, k_slot);
            umma_commit(&s_full[q]);  // covers P.V(q, j) too: "S(j+2) ready" implies "P buffer free"
          }
          if (last) {
            if (has_pv) umma_commit(&ring_empty[v_slot]);
            if (has_qk) umma_commit(&ring_empty[k_slot]);
          }
        }
        __syncwarp();
        next[q] = j + 1;
      }
      if (all_done) break;
      if (!progress) __nanosleep(40);
    }
  } else {
    // ============================ softmax / correction / epilogue ============================
    const int q = warp >> 2;
    const int r = (warp & 3) * 32 + lane;  // row within the Q tile = TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t s_tmem = tmem + q * S::Q_STRIDE + lane_sel;
    const uint32_t p_tmem = s_tmem + S::COL_P;
    const uint32_t o_tmem = s_tmem + S::COL_O;
    const int qrow = blockIdx.x * S::CTA_ROWS + q * QT_ROWS + r;  // row within the sequence
    constexpr float kDelta = trunc_delta(V);
    constexpr float kBiasLog2 = kDelta * 1.4426950408889634f * (1.0f - 0.5f * kDelta);  // log2(1 + delta)
    const uint64_t sc2 = f2_splat(a.scale_log2);
    uint64_t nm2 = 0ull;
    float m = -INFINITY;  // reference point of the exponentials (scaled log2 units); lags the true row max by <= 8
    float l = 0.f;        // running row sum

#ifdef D4D_ATTN_TRACE
    unsigned long long* tr = (a.trace && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && (warp & 3) == 0)
                                 ? a.trace + static_cast<size_t>(q) * 64 * 8 : nullptr;
#endif
    uint32_t sv[KV];
    auto rescale_o = [&](float alpha) {  // O[row, :] *= alpha (warp-collective; alpha is per row)
#pragma unroll 1
      for (int c = 0; c < 64; c += 16) {
        uint32_t v[16];
        tmem_ld16(o_tmem + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
        tmem_st16(o_tmem + c, v);
      }
    };

    mbar_wait_quiet(&s_full[q], 0);
    tc_fence_after();
#pragma unroll
    for (int c = 0; c < S::CHUNKS; ++c) tmem_ld32(s_tmem + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
    tmem_ld_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&tile_done[q]);  // S_q(0) is in registers: Q.K^T(q, 1) may overwrite S_q

    // one K/V tile: sv (S_q(j)) -> P_q(j) in TMEM, l, m; S_q(j+1) -> sv
    auto tile = [&](auto masked_tag, int j, int valid) {
      constexpr bool kMasked = decltype(masked_tag)::value;
      const bool has_next = j + 1 < n_tiles;
      // P_q is free once P.V(q, j-1) has completed; S_q(j+1) complete implies that (commit order in the MMA warp).
      // Looking at an mbarrier costs ~250 cycles even when its phase is complete, so it is PROBED twice without
      // consuming the answer (here and after the row max) and the answers are consumed right before the first P store:
      // the probes' latencies hide behind the row max and the exponentials of chunk 0.
      uint64_t* free_bar = has_next ? &s_full[q] : &pv_done[q];
      const uint32_t free_par = has_next ? ((j + 1) & 1) : ((j - 1) & 1);
      const bool need_free = has_next || j > 0;
      const bool probe0 = !need_free || mbar_test(free_bar, free_par);
#ifdef D4D_ATTN_TRACE
      if (j >= 64) tr = nullptr;
#endif
      D4D_TRACE(0);

      // ---- row max of this tile ----
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < KV; i += 8) {
        if (!kMasked) {
          mx0 = fmax3(mx0, __uint_as_float(sv[i]), __uint_as_float(sv[i + 1]));
          mx1 = fmax3(mx1, __uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3]));
          mx2 = fmax3(mx2, __uint_as_float(sv[i + 4]), __uint_as_float(sv[i + 5]));
          mx3 = fmax3(mx3, __uint_as_float(sv[i + 6]), __uint_as_float(sv[i + 7]));
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (i + k < valid) mx0 = fmaxf(mx0, __uint_as_float(sv[i + k]));
        }
      }
      const float tmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * a.scale_log2;
      // lazy reference point: move it only when the row max grew by more than 8 (P <= 2^8 otherwise; bf16 and the fp32
      // sums have the range for it).  Always taken for tile 0 (m = -inf).
      float alpha = 1.f;
      if (tmax > m + 8.f) {
        alpha = ex2f(m - tmax);  // 0 for tile 0
        m = tmax;
        nm2 = f2_splat(kBiasLog2 - m);
      }
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        mbar_wait_quiet(&pv_done[q], (j - 1) & 1);  // O_q holds P.V(0..j-1); P.V(j) is not issued before this tile is done
        tc_fence_after();
        rescale_o(alpha);
      }
      l *= alpha;  // l == 0 for tile 0
      const bool probe1 = probe0 || mbar_test(free_bar, free_par);
      D4D_TRACE(1);

      // ---- P = exp2(S * scale - m), chunks of 32 columns ----
      uint64_t lsum[2] = {0ull, 0ull};
      auto chunk = [&](auto c_tag) {
        constexpr int c = decltype(c_tag)::value;
        if constexpr (c < S::CHUNKS) {
          uint32_t pk[16];
          ExpChunk<VAR, kMasked, KV, c>::run(sv, pk, sc2, nm2, lsum, valid);
          if (c == 0) D4D_TRACE(2);
          if (c == 0 && need_free) {
            if (!__all_sync(0xffffffffu, probe1)) mbar_wait_quiet(free_bar, free_par);
            tc_fence_after();
          }
          D4D_TRACE(3 + c);
          tmem_st16(p_tmem + c * 16, pk);
          // the 32 columns of this chunk are dead: pull the same columns of S_q(j+1) in now, so that only the last
          // chunk's TMEM round trip is exposed at the end of the tile (s_full(j+1) was waited for before the first P store)
          if (has_next) tmem_ld32(s_tmem + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[c * 32]));
        }
      };
      chunk(std::integral_constant<int, 0>{});
      chunk(std::integral_constant<int, 1>{});
      chunk(std::integral_constant<int, 2>{});
      chunk(std::integral_constant<int, 3>{});
      {
        float s0, s1, s2, s3;
        f2_unpack(lsum[0], s0, s1);
        f2_unpack(lsum[1], s2, s3);
        l += (s0 + s1) + (s2 + s3);
      }
      tmem_st_wait();
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tile_done[q]);
      D4D_TRACE(7);
    };

    for (int j = 0; j < n_tiles; ++j) {
      const int valid = a.seq_kv - j * KV;
      if (valid >= KV) tile(std::false_type{}, j, KV);
      else tile(std::true_type{}, j, valid);
    }

    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait_quiet(&pv_done[q], (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv_l = (1.0f + kDelta) / l;  // trunc variants: l was summed from the (1 + delta)-shifted exponentials
    bf16* orow = a.out + static_cast<size_t>(b * a.seq_q + qrow) * a.ld_out + col0;
#pragma unroll 1
    for (int c = 0; c < 64; c += 16) {
      uint32_t v[16];
      tmem_ld16(o_tmem + c, v);
      tmem_ld_wait();
      if (qrow < a.seq_q) {
        uint4 o0, o1;
        o0.x = pack_bf16x2(__uint_as_float(v[0]) * inv_l, __uint_as_float(v[1]) * inv_l);
        o0.y = pack_bf16x2(__uint_as_float(v[2]) * inv_l, __uint_as_float(v[3]) * inv_l);
        o0.z = pack_bf16x2(__uint_as_float(v[4]) * inv_l, __uint_as_float(v[5]) * inv_l);
        o0.w = pack_bf16x2(__uint_as_float(v[6]) * inv_l, __uint_as_float(v[7]) * inv_l);
        o1.x = pack_bf16x2(__uint_as_float(v[8]) * inv_l, __uint_as_float(v[9]) * inv_l);
        o1.y = pack_bf16x2(__uint_as_float(v[10]) * inv_l, __uint_as_float(v[11]) * inv_l);
        o1.z = pack_bf16x2(__uint_as_float(v[12]) * inv_l, __uint_as_float(v[13]) * inv_l);
        o1.w = pack_bf16x2(__uint_as_float(v[14]) * inv_l, __uint_as_float(v[15]) * inv_l);
        uint4* op = reinterpret_cast<uint4*>(orow + c);
        op[0] = o0;
        op[1] = o1;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

template <int VAR, int NQ, int KV>
int launch_attn64(const AttnLaunch& L, cudaStream_t stream) {
  using S = Shape<NQ, KV>;
  static PerDeviceOnce attr_once;
  if (int rc = ensure_dyn_smem(attn64_fwd_kernel<VAR, NQ, KV>, S::SMEM, attr_once)) return rc;
  A64Args a;
  a.seq_q = L.d.seq;
  a.seq_kv = L.d.seq_kv > 0 ? L.d.seq_kv : L.d.seq;
  a.heads = L.d.heads;
  a.n_kv_tiles = (a.seq_kv + KV - 1) / KV;
  a.scale_log2 = L.d.scale * 1.4426950408889634f;
  a.out = L.d.out;
  a.ld_out = L.d.ld_out;
  a.trace = nullptr;
#ifdef D4D_ABLATE
  if (const char* e = getenv("D4D_ATTN_TRACE")) a.trace = reinterpret_cast<unsigned long long*>(strtoull(e, nullptr, 0));
#endif
  dim3 grid((L.d.seq + S::CTA_ROWS - 1) / S::CTA_ROWS, L.d.batch * L.d.heads);
  D4D_CUDA_OK(launch_pdl(attn64_fwd_kernel<VAR, NQ, KV>, grid, dim3(S::THREADS), S::SMEM, stream, L.tmap_q, L.tmap_k, L.tmap_v, a));
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int NQ, int KV>
int launch_variant(int v, const AttnLaunch& L, cudaStream_t stream) {
#ifdef D4D_ABLATE  // tools build: every variant (tools/bench_attention.py)
  switch (v) {
    case 0: return launch_attn64<0, NQ, KV>(L, stream);
    case 1: return launch_attn64<1, NQ, KV>(L, stream);
    case 2: return launch_attn64<2, NQ, KV>(L, stream);
    case 3: return launch_attn64<3, NQ, KV>(L, stream);
    case 4: return launch_attn64<4, NQ, KV>(L, stream);
    case 5: return launch_attn64<5, NQ, KV>(L, stream);
    case 6: return launch_attn64<6, NQ, KV>(L, stream);
  }
  set_error("attention: unknown D4D_ATTN_VARIANT");
  return 1;
#else
  return launch_attn64<D4D_ATTN64_VARIANT, NQ, KV>(L, stream);
#endif
}

}  // namespace

// kernel shape of a launch: L->variant = 0 encodes <2, 128>, -1 encodes <3, 64>
static int attn64_shape() {
#ifdef D4D_ABLATE
  if (const char* e = getenv("D4D_ATTN_SHAPE")) return atoi(e);
#endif
  return D4D_ATTN64_SHAPE;
}

// tensor maps + grid of the head_dim-64 kernel (called by attn_prepare)
int attn64_prepare(const AttnDesc& d, AttnLaunch* L) {
  const int shape = attn64_shape();
  const int kv_rows = shape == 1 ? 64 : 128;
  const int cta_rows = shape == 1 ? 3 * QT_ROWS : 2 * QT_ROWS;
  const int seq_kv = d.seq_kv > 0 ? d.seq_kv : d.seq;
  const int ld_kv = d.ld_kv > 0 ? d.ld_kv : d.ld_qkv;
  const uint64_t q_tokens = static_cast<uint64_t>(d.batch) * d.seq;
  const uint64_t kv_tokens = static_cast<uint64_t>(d.batch) * seq_kv;
  const uint64_t width = static_cast<uint64_t>(d.heads) * 64;
  if (int rc = make_tmap_2d(&L->tmap_q, d.q, q_tokens, width, d.ld_qkv, 64, QT_ROWS, 128)) return rc;
  if (int rc = make_tmap_2d(&L->tmap_k, d.k, kv_tokens, width, ld_kv, 64, kv_rows, 128)) return rc;
  if (int rc = make_tmap_2d(&L->tmap_v, d.v, kv_tokens, width, ld_kv, 64, kv_rows, 128)) return rc;
  L->variant = shape == 1 ? -1 : 0;
  L->grid_x = (d.seq + cta_rows - 1) / cta_rows;
  L->grid_y = d.batch * d.heads;
  return 0;
}

int attn64_run(const AttnLaunch& L, cudaStream_t stream) {
  int v = D4D_ATTN64_VARIANT;
#ifdef D4D_ABLATE
  if (const char* e = getenv("D4D_ATTN_VARIANT")) v = atoi(e);
#endif
  if (L.variant == -1) return launch_variant<3, 64>(v, L, stream);
  return launch_variant<2, 128>(v, L, stream);
}

}  // namespace d4d
