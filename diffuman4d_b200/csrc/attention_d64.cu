// tcgen05 flash-attention forward for head_dim 64 on sm_100a: the kernel behind every attention call of the SD-2.1
// channel layout (5/10/20/20 heads of 64), i.e. the 3-D attention over all F frames of a CFG half (reference
// src/diffusers/models/attention.py:68-83, F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0) and the
// per-image 2-D attention of level 0.  Same operand contract as attention_umma.cu (Q/K/V are column slices of the fused
// QKV-GEMM output, read in place through strided TMA boxes).
//
// Why a second kernel: at head_dim 64 every exponential carries only 4*64 FLOP, so the softmax - not the tensor pipe -
// bounds the kernel (MUFU.EX2 runs at 16/clk/SM against 32/clk needed; profiles/README.md).  The generic kernel spends
// about half of its time on per-tile fixed costs (64-key tiles: one barrier round trip, one TMEM round trip and an
// issue-bound N = 64 MMA group per 64 keys).  This kernel is organised around the softmax instead:
//
//   * one CTA per SM, 256 query rows (two 128-row Q tiles) x 128-key K/V tiles: half the per-key fixed costs, every
//     K/V tile staged once for both Q tiles, Q.K^T issued as full-rate N = 128 MMAs;
//   * 8 softmax warps (two per sub-partition, one Q tile each; thread = query row = TMEM lane).  The 128 scores of a
//     tile are pulled into registers in one TMEM round trip at the END of the previous tile, which frees S for
//     Q.K^T(j+1) at once: registers are the second S buffer, so the tensor pipe always runs one tile ahead of the
//     softmax and the softmax warps never wait for it;
//   * exp2 is split between the MUFU pipe and a Cody-Waite / degree-3 polynomial evaluated with packed FFMA2 / FADD2 on
//     the FMA pipe (round-down magic-number split, exponent patched in with one IMAD); the share is a compile-time
//     pattern over pairs of columns (kPolyMask), 8.8e-5 relative error, far below the bf16 rounding of P;
//   * the running max is lazy: it only moves when the row max grows by more than 8 (log2 units), so O is rescaled (by
//     the softmax warps themselves, in TMEM) a handful of times per row.
//
// TMEM (512 columns): per Q tile q: S_q fp32 [q*256, +128) | P_q bf16 [+128, +64) | O_q fp32 [+192, +64).
// Warps: 0-3 softmax of Q tile 0, 4-7 softmax of Q tile 1, 8 MMA issuer (+ TMEM owner), 9 TMA producer.
// Hand-offs: tcgen05.commit -> mbarrier (s_full[q], pv_done[q]) towards the softmax warps; one named barrier per Q tile
// ("P(j) stored and S(j+1) in registers") towards the MMA warp, which then issues P.V(j) and Q.K^T(j+2).
#include <math.h>

#include "kernels.h"

namespace d4d {

namespace {

constexpr int QT_ROWS = 128;               // rows of one Q tile (= TMEM lanes)
constexpr int CTA_ROWS = 2 * QT_ROWS;      // query rows per CTA
constexpr int KV_ROWS = 128;               // keys per K/V tile
constexpr int TILE_BYTES = 128 * 64 * 2;   // one [128 rows][64 ch] 128B-swizzled box (Q, K or V tile)
constexpr int SLOTS = 8;                   // K/V ring
constexpr int A64_THREADS = 320;
constexpr int A64_SMEM = 2 * TILE_BYTES + SLOTS * TILE_BYTES + 1024 + 256;
constexpr int TMEM_COLS = 512;

struct A64Args {
  int seq_q, seq_kv, heads, n_kv_tiles;
  float scale_log2;
  bf16* out;
  int ld_out;
};

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

// "tile done" hand-off to the MMA warp: 128 softmax threads of Q tile q arrive, the MMA warp syncs (ids 1, 2)
__device__ __forceinline__ void tile_done_arrive(int q) { asm volatile("bar.arrive %0, 160;" ::"r"(q + 1) : "memory"); }
__device__ __forceinline__ void tile_done_sync(int q) { asm volatile("bar.sync %0, 160;" ::"r"(q + 1) : "memory"); }

// ---- kernel variants (tools build instantiates all of them for A/B timing; the product uses kDefaultVariant) ----
// poly_mask: bit p = the p-th pair (of every 8 consecutive column pairs) takes the FMA-pipe exp2; int_mufu / int_poly:
// pack P with integer rounding (IADD + PRMT on the ALU pipe) instead of F2FP (XU pipe, shared with MUFU)
struct Variant { uint32_t poly_mask; bool int_mufu, int_poly; };
__host__ __device__ constexpr Variant variant_of(int v) {
  return v == 0 ? Variant{0x00, false, false}
       : v == 1 ? Variant{0x44, false, false}   // 2/8 poly
       : v == 2 ? Variant{0x92, false, false}   // 3/8 poly
       : v == 3 ? Variant{0xaa, false, false}   // 4/8 poly
       : v == 4 ? Variant{0x92, false, true}
       : v == 5 ? Variant{0xaa, false, true}
       : v == 6 ? Variant{0xaa, true, true}
       :          Variant{0x00, true, false};
}
constexpr int kNumVariants = 8;
#ifndef D4D_ATTN64_DEFAULT
#define D4D_ATTN64_DEFAULT 2
#endif

// exp2 of one pair of columns of this thread's row -> packed bf16x2 (lo = first column), row-sum share into lsum.
//   x = s * scale_log2 - m  (one FFMA2).  MUFU path: ex2.approx.  Polynomial path (x <= ~100 by construction of m):
//   xc = max(x, -126); t = xc + 1.5*2^23 rounded DOWN, so the low mantissa bits of t hold floor(xc); r = xc - floor(xc)
//   in [0, 1); 2^r ~ 1 + r (c1 + r (c2 + r c3)); result bits = (t << 23) + bits(2^r)  (exponent += floor(xc)).
template <bool kPoly, bool kIntPack, bool kMasked>
__device__ __forceinline__ uint32_t exp_pair(uint32_t s0, uint32_t s1, uint64_t sc2, uint64_t nm2, uint64_t& lsum, int col,
                                             int valid) {
  const uint64_t x = f2_fma(f2_pack(__uint_as_float(s0), __uint_as_float(s1)), sc2, nm2);
  float x0, x1, e0, e1;
  f2_unpack(x, x0, x1);
  if (!kPoly) {
    e0 = ex2f(x0);
    e1 = ex2f(x1);
  } else {
    const uint64_t xc = f2_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
    const uint64_t t = add2_rm(xc, f2_splat(12582912.f));
    const uint64_t fl = f2_add(t, f2_splat(-12582912.f));  // exact
    const uint64_t r = sub2(xc, fl);
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
  if (kIntPack) return __byte_perm(__float_as_uint(e0) + 0x8000u, __float_as_uint(e1) + 0x8000u, 0x7632);  // e >= 0, finite
  return pack_bf16x2(e0, e1);
}

template <int VAR>
__global__ void __launch_bounds__(A64_THREADS, 1)
attn64_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                  const __grid_constant__ CUtensorMap tmap_v, const A64Args a) {
  constexpr Variant V = variant_of(VAR);
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                      // two Q tiles
  uint8_t* sRing = smem + 2 * TILE_BYTES;  // K/V ring
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + SLOTS * TILE_BYTES);
  uint64_t* ring_full = bars;           // [SLOTS]
  uint64_t* ring_empty = bars + SLOTS;  // [SLOTS]
  uint64_t* q_full = bars + 2 * SLOTS;
  uint64_t* s_full = q_full + 1;   // [2]  Q.K^T(q, j) complete (also: every MMA issued before it)
  uint64_t* pv_done = q_full + 3;  // [2]  P.V(q, j) complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 5);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // role branches are warp-uniform
  const int lane = threadIdx.x & 31;
  const int bh = blockIdx.y;
  const int b = bh / a.heads;
  const int hd = bh - b * a.heads;
  const int q_row0 = b * a.seq_q + blockIdx.x * CTA_ROWS;  // first query-token row of this CTA
  const int kv_row0 = b * a.seq_kv;                       // first key/value-token row of this batch entry
  const int col0 = hd * 64;                               // first column of this head inside the q/k/v slice
  const int n_tiles = a.n_kv_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < SLOTS; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], 1);
    }
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();
  pdl_launch_dependents();

  if (warp == 9) {
    // ============================ TMA producer ============================
    if (elect_one()) {
      tma_prefetch_desc(&tmap_q);
      tma_prefetch_desc(&tmap_k);
      tma_prefetch_desc(&tmap_v);
      mbar_expect_tx(q_full, 2 * TILE_BYTES);
      tma_load_2d(sQ, &tmap_q, q_full, col0, q_row0);
      tma_load_2d(sQ + TILE_BYTES, &tmap_q, q_full, col0, q_row0 + QT_ROWS);
    }
    __syncwarp();
    int slot = 0;
    uint32_t phase = 0;
    auto load_tile = [&](const CUtensorMap* tm, int j) {
      mbar_wait(&ring_empty[slot], phase ^ 1);
      if (elect_one()) {
        mbar_expect_tx(&ring_full[slot], TILE_BYTES);
        tma_load_2d(sRing + slot * TILE_BYTES, tm, &ring_full[slot], col0, kv_row0 + j * KV_ROWS);
      }
      __syncwarp();
      if (++slot == SLOTS) { slot = 0; phase ^= 1; }
    };
    // the order the MMA warp consumes: K0 K1, then for every j: V(j), K(j+2)
    load_tile(&tmap_k, 0);
    if (n_tiles > 1) load_tile(&tmap_k, 1);
    for (int j = 0; j < n_tiles; ++j) {
      load_tile(&tmap_v, j);
      if (j + 2 < n_tiles) load_tile(&tmap_k, j + 2);
    }
  } else if (warp == 8) {
    // ============================ MMA issuer ============================
    // whole warp in uniform control flow, one elected lane around the asynchronous instructions (see gemm_umma.cu)
    const uint32_t idesc_qk = make_idesc_bf16(QT_ROWS, KV_ROWS, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(QT_ROWS, 64, 0, 1);
    const uint32_t q_addr = smem_u32(sQ);
    const uint32_t ring_addr = smem_u32(sRing);
    int slot = 0;
    uint32_t phase = 0;
    auto wait_ahead = [&](int k) {  // wait for the k-th next ring slot without consuming it
      int sl = slot + k;
      uint32_t ph = phase;
      if (sl >= SLOTS) { sl -= SLOTS; ph ^= 1; }
      mbar_wait(&ring_full[sl], ph);
    };
    auto advance = [&]() {
      const int used = slot;
      if (++slot == SLOTS) { slot = 0; phase ^= 1; }
      return used;
    };
    auto issue_qk = [&](int q, int sl) {  // S_q = Q_q K^T : 4 x (M128 N128 K16), both operands K-major SW128
      const uint32_t kaddr = ring_addr + sl * TILE_BYTES;
      const uint32_t s_tmem = tmem + q * 256;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = make_smem_desc(q_addr + q * TILE_BYTES + k * 32, 0, 1024, 2);
        const uint64_t bd = make_smem_desc(kaddr + k * 32, 0, 1024, 2);
        umma_ss(s_tmem, ad, bd, idesc_qk, k != 0 ? 1u : 0u);
      }
    };
    auto issue_pv = [&](int q, int sl, int j) {  // O_q += P_q V : 8 x (M128 N64 K16), A = P bf16 in TMEM, B = V MN-major
      const uint32_t vaddr = ring_addr + sl * TILE_BYTES;
      const uint32_t p_tmem = tmem + q * 256 + 128;
      const uint32_t o_tmem = tmem + q * 256 + 192;
#pragma unroll
      for (int k = 0; k < KV_ROWS / 16; ++k) {
        // 16 keys = two 8-row swizzle atoms = 2048 bytes; SBO = 1024 (next 8 keys); LBO unused (one 64-wide d block)
        const uint64_t bd = make_smem_desc(vaddr + k * 2048, TILE_BYTES, 1024, 2);
        umma_ts(o_tmem, p_tmem + k * 8, bd, idesc_pv, (j | k) != 0 ? 1u : 0u);
      }
    };
    mbar_wait(q_full, 0);
    {  // S_q(0) for both Q tiles
      wait_ahead(0);
      const int k_slot = advance();
      if (elect_one()) {
        issue_qk(0, k_slot);
        umma_commit(&s_full[0]);
        issue_qk(1, k_slot);
        umma_commit(&s_full[1]);
        umma_commit(&ring_empty[k_slot]);
      }
      __syncwarp();
    }
    {  // S_q(1) as soon as the softmax warps hold S_q(0) in registers
      const bool more = n_tiles > 1;
      int k_slot = 0;
      if (more) {
        wait_ahead(0);
        k_slot = advance();
      }
      for (int q = 0; q < 2; ++q) {
        tile_done_sync(q);
        tc_fence_after();
        if (more && elect_one()) {
          issue_qk(q, k_slot);
          umma_commit(&s_full[q]);
          if (q == 1) umma_commit(&ring_empty[k_slot]);
        }
        __syncwarp();
      }
    }
    for (int j = 0; j < n_tiles; ++j) {
      const bool more = j + 2 < n_tiles;
      wait_ahead(0);            // V(j)
      if (more) wait_ahead(1);  // K(j+2)
      const int v_slot = advance();
      const int k_slot = more ? advance() : 0;
      for (int q = 0; q < 2; ++q) {
        tile_done_sync(q);  // P_q(j) is in TMEM and S_q(j+1) has been read out
        tc_fence_after();
        if (elect_one()) {
          issue_pv(q, v_slot, j);
          umma_commit(&pv_done[q]);
          if (more) {
            issue_qk(q, k_slot);
            umma_commit(&s_full[q]);  // covers P.V(q, j) too: "S(j+2) ready" implies "P buffer free"
          }
          if (q == 1) {
            umma_commit(&ring_empty[v_slot]);
            if (more) umma_commit(&ring_empty[k_slot]);
          }
        }
        __syncwarp();
      }
    }
  } else {
    // ============================ softmax / correction / epilogue ============================
    const int q = warp >> 2;
    const int r = (warp & 3) * 32 + lane;  // row within the Q tile = TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t s_tmem = tmem + q * 256 + lane_sel;
    const uint32_t p_tmem = s_tmem + 128;
    const uint32_t o_tmem = s_tmem + 192;
    const int qrow = blockIdx.x * CTA_ROWS + q * QT_ROWS + r;  // row within the sequence
    const uint64_t sc2 = f2_splat(a.scale_log2);
    float m = -INFINITY;  // reference point of the exponentials (scaled log2 units); lags the true row max by <= 8
    float l = 0.f;        // running row sum

    uint32_t sv[128];
    auto load_s = [&]() {
      tmem_ld32(s_tmem, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
      tmem_ld32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
      tmem_ld32(s_tmem + 64, *reinterpret_cast<uint32_t(*)[32]>(&sv[64]));
      tmem_ld32(s_tmem + 96, *reinterpret_cast<uint32_t(*)[32]>(&sv[96]));
    };
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

    mbar_wait(&s_full[q], 0);
    tc_fence_after();
    load_s();
    tmem_ld_wait();
    tc_fence_before();
    tile_done_arrive(q);  // S_q(0) is in registers: Q.K^T(q, 1) may overwrite S_q

    // one K/V tile: sv (S_q(j)) -> P_q(j) in TMEM, l, m; then S_q(j+1) -> sv
    auto tile = [&](auto masked_tag, int j, int valid) {
      constexpr bool kMasked = decltype(masked_tag)::value;
      const bool has_next = j + 1 < n_tiles;
      // P_q is free once P.V(q, j-1) has completed; S_q(j+1) complete implies that (commit order in the MMA warp).
      // Probe now, wait (if at all) right before the first P store.
      uint64_t* free_bar = has_next ? &s_full[q] : &pv_done[q];
      const uint32_t free_par = has_next ? ((j + 1) & 1) : ((j - 1) & 1);
      const bool need_free = has_next || j > 0;
      bool p_free = !need_free || __all_sync(0xffffffffu, mbar_test(free_bar, free_par));

      // ---- row max of this tile ----
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 128; i += 8) {
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
      }
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        mbar_wait(&pv_done[q], (j - 1) & 1);  // O_q holds P.V(0..j-1); P.V(j) is not issued before this tile is done
        tc_fence_after();
        rescale_o(alpha);
      }
      l *= alpha;  // l == 0 for tile 0

      // ---- P = exp2(S * scale - m), 4 chunks of 32 columns ----
      const uint64_t nm2 = f2_splat(-m);
      uint64_t lsum[2] = {0ull, 0ull};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int p = 0; p < 16; ++p) {
          const int gp = c * 16 + p;
          constexpr uint32_t mask = V.poly_mask;
          const bool poly = !kMasked && ((mask >> (gp & 7)) & 1u);
          if (poly) {
            if (V.int_poly) pk[p] = exp_pair<true, true, false>(sv[2 * gp], sv[2 * gp + 1], sc2, nm2, lsum[p & 1], 2 * gp, valid);
            else pk[p] = exp_pair<true, false, false>(sv[2 * gp], sv[2 * gp + 1], sc2, nm2, lsum[p & 1], 2 * gp, valid);
          } else {
            if (V.int_mufu) pk[p] = exp_pair<false, true, kMasked>(sv[2 * gp], sv[2 * gp + 1], sc2, nm2, lsum[p & 1], 2 * gp, valid);
            else pk[p] = exp_pair<false, false, kMasked>(sv[2 * gp], sv[2 * gp + 1], sc2, nm2, lsum[p & 1], 2 * gp, valid);
          }
        }
        if (c == 0 && need_free) {
          if (!p_free) mbar_wait(free_bar, free_par);
          tc_fence_after();
        }
        tmem_st16(p_tmem + c * 16, pk);
      }
      {
        float s0, s1, s2, s3;
        f2_unpack(lsum[0], s0, s1);
        f2_unpack(lsum[1], s2, s3);
        l += (s0 + s1) + (s2 + s3);
      }
      if (has_next) load_s();  // S_q(j+1): s_full was waited for above (it is the "P free" barrier of this tile)
      tmem_st_wait();
      tmem_ld_wait();
      tc_fence_before();
      tile_done_arrive(q);
    };

    for (int j = 0; j < n_tiles; ++j) {
      const int valid = a.seq_kv - j * KV_ROWS;
      if (valid >= KV_ROWS) tile(std::false_type{}, j, KV_ROWS);
      else tile(std::true_type{}, j, valid);
    }

    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait(&pv_done[q], (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
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
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

template <int VAR>
int launch_attn64(const AttnLaunch& L, cudaStream_t stream) {
  static PerDeviceOnce attr_once;
  if (int rc = ensure_dyn_smem(attn64_fwd_kernel<VAR>, A64_SMEM, attr_once)) return rc;
  A64Args a;
  a.seq_q = L.d.seq;
  a.seq_kv = L.d.seq_kv > 0 ? L.d.seq_kv : L.d.seq;
  a.heads = L.d.heads;
  a.n_kv_tiles = (a.seq_kv + KV_ROWS - 1) / KV_ROWS;
  a.scale_log2 = L.d.scale * 1.4426950408889634f;
  a.out = L.d.out;
  a.ld_out = L.d.ld_out;
  dim3 grid(L.grid_x, L.grid_y);
  D4D_CUDA_OK(launch_pdl(attn64_fwd_kernel<VAR>, grid, dim3(A64_THREADS), A64_SMEM, stream, L.tmap_q, L.tmap_k, L.tmap_v, a));
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

// tensor maps + grid of the head_dim-64 kernel (called by attn_prepare)
int attn64_prepare(const AttnDesc& d, AttnLaunch* L) {
  const int seq_kv = d.seq_kv > 0 ? d.seq_kv : d.seq;
  const int ld_kv = d.ld_kv > 0 ? d.ld_kv : d.ld_qkv;
  const uint64_t q_tokens = static_cast<uint64_t>(d.batch) * d.seq;
  const uint64_t kv_tokens = static_cast<uint64_t>(d.batch) * seq_kv;
  const uint64_t width = static_cast<uint64_t>(d.heads) * 64;
  if (int rc = make_tmap_2d(&L->tmap_q, d.q, q_tokens, width, d.ld_qkv, 64, QT_ROWS, 128)) return rc;
  if (int rc = make_tmap_2d(&L->tmap_k, d.k, kv_tokens, width, ld_kv, 64, KV_ROWS, 128)) return rc;
  if (int rc = make_tmap_2d(&L->tmap_v, d.v, kv_tokens, width, ld_kv, 64, KV_ROWS, 128)) return rc;
  L->grid_x = (d.seq + CTA_ROWS - 1) / CTA_ROWS;
  L->grid_y = d.batch * d.heads;
  return 0;
}

int attn64_run(const AttnLaunch& L, cudaStream_t stream) {
#ifdef D4D_ABLATE  // tools build: every variant, chosen by D4D_ATTN_VARIANT (tools/bench_attention.py)
  const char* e = getenv("D4D_ATTN_VARIANT");
  const int v = e ? atoi(e) : D4D_ATTN64_DEFAULT;
  switch (v) {
    case 0: return launch_attn64<0>(L, stream);
    case 1: return launch_attn64<1>(L, stream);
    case 2: return launch_attn64<2>(L, stream);
    case 3: return launch_attn64<3>(L, stream);
    case 4: return launch_attn64<4>(L, stream);
    case 5: return launch_attn64<5>(L, stream);
    case 6: return launch_attn64<6>(L, stream);
    case 7: return launch_attn64<7>(L, stream);
  }
  set_error("attention: unknown D4D_ATTN_VARIANT");
  return 1;
#else
  return launch_attn64<D4D_ATTN64_DEFAULT>(L, stream);
#endif
}

}  // namespace d4d
