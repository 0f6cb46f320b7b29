// tcgen05 flash-attention forward for sm_100a (no mask, non-causal; the only attention on the path).
//
//   O[b, i, h, :] = softmax_j(scale * Q[b,i,h,:] . K[b,j,h,:]) V[b,j,h,:]
//
// Replaces F.scaled_dot_product_attention as called by diffusers' AttnProcessor2_0 for attn1 (3-D: the
// tokens of all F frames of one CFG half form ONE sequence, reference src/diffusers/models/attention.py:68-83)
// and attn2 (per-image).  Q/K/V are read in place from the fused QKV-GEMM output [tokens, 3C] through
// strided TMA boxes, so "(b t) hw c -> b (t hw) c" and the head split are address arithmetic only.
//
// One CTA = 128 query rows of one (batch, head), K/V tiles of 64 keys; two CTAs are co-resident per SM at head_dim 64
// (256 TMEM columns each: S0 S1 | P0 P1 | O).  S and P are DOUBLE-BUFFERED in TMEM so that the three stages
//     S(j+2) = Q K^T   |   softmax(j): S -> P   |   O += P(j-1) V
// run concurrently and the softmax warps never wait for the tensor core in steady state.
// Warp roles (192 threads):
//   warps 0-3  softmax: thread t owns query row t (TMEM lane t).  The 64 S values of the tile are pulled into
//              registers with one exposed TMEM round trip and S is released at once.  P = exp2(S*scale - m) uses the
//              running max m of the previous tiles; if the row max grows by more than 8 (log2 units) O and l are
//              rescaled before the next tile, and only if it would overflow (> 100; always for tile 0) P is recomputed
//              from the registers with the new max.  FMNMX3 / FFMA2 / FADD2 packed math; the masked tail tile is a
//              separate instantiation.
//   warp  4    TMA producer: Q once, then K0 K1 V0 K2 V1 K3 ... through a ring of 8 KB*NB slots
//   warp  5    MMA issuer:   S = Q K_j^T (SS, both K-major), O += P V_j (A = P bf16 from TMEM, B = V MN-major)
#include <math.h>
#include <stdlib.h>

#include "kernels.h"

namespace d4d {

namespace {

constexpr int BLOCK_Q = 128;
constexpr int BLOCK_KV = 64;
constexpr int ATT_THREADS = 192;
constexpr int QTILE_BYTES = 128 * 64 * 2;  // one [128 rows][64 ch] swizzled box
constexpr int KTILE_BYTES = 64 * 64 * 2;   // one [64 keys][64 ch] swizzled box

template <int NB>
struct AttCfg {
  static constexpr int D = 64 * NB;
  static constexpr int SLOT_BYTES = KTILE_BYTES * NB;  // one K or V tile
  static constexpr int SLOTS = NB == 1 ? 8 : 6;
  static constexpr int Q_BYTES = QTILE_BYTES * NB;
  static constexpr int SMEM_BYTES = Q_BYTES + SLOTS * SLOT_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = NB == 1 ? 256 : 512;
  static constexpr int COL_S = 0, COL_P = 128, COL_O = 192;  // S0 S1 (64 each) | P0 P1 (32 each) | O (D)
};

struct AttKernelArgs {
  int seq_q, seq_kv, heads, n_kv_tiles;
  float scale_log2;
  bf16* out;
  int ld_out;
  int dbg;  // ablation switches for tools/ablate_attention.py (0 in production): 1 no ex2, 2 no P store, 4 no P.V, 8 no Q.K
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sm_100a packed-fp32 / 3-input helpers (SASS: FMNMX3, FFMA2, FADD2) -- halve the issue slots of the softmax
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// named barriers 1/2 ("P(j) ready", even/odd tiles): 128 softmax threads arrive, the MMA warp syncs -- a hardware barrier
// wake-up is several times faster than an mbarrier try_wait round trip, and this hop is on the per-tile critical path
__device__ __forceinline__ void pready_arrive(int buf) { asm volatile("bar.arrive %0, 160;" ::"r"(buf + 1) : "memory"); }
__device__ __forceinline__ void pready_sync(int buf) { asm volatile("bar.sync %0, 160;" ::"r"(buf + 1) : "memory"); }

// Row max over 32 S columns held in registers (four independent FMNMX3 chains); kMasked: only columns < valid count
template <bool kMasked>
__device__ __forceinline__ void max32(const uint32_t* v, float (&mx)[4], int valid_cols) {
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    if (!kMasked) {
      mx[0] = max3(mx[0], __uint_as_float(v[i]), __uint_as_float(v[i + 1]));
      mx[1] = max3(mx[1], __uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
      mx[2] = max3(mx[2], __uint_as_float(v[i + 4]), __uint_as_float(v[i + 5]));
      mx[3] = max3(mx[3], __uint_as_float(v[i + 6]), __uint_as_float(v[i + 7]));
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k)
        if (i + k < valid_cols) mx[k & 3] = fmaxf(mx[k & 3], __uint_as_float(v[i + k]));
    }
  }
}
// e = exp2(s*scale - m) for 32 columns -> 16 packed bf16x2 P columns in TMEM, row-sum share in two f32x2 accumulators
template <bool kMasked>
__device__ __forceinline__ void exp32(const uint32_t* v, uint64_t sc2, uint64_t nm2, uint64_t (&lsum)[2], uint32_t p_addr,
                                      int valid_cols, int dbg = 0) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t pk[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int i = h * 8 + k;
      const uint64_t x = fma2(pack2(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), sc2, nm2);
      float x0, x1;
      unpack2(x, x0, x1);
      float e0, e1;
      if (dbg & 1) { e0 = x0; e1 = x1; }
      else { e0 = ex2_approx(x0); e1 = ex2_approx(x1); }
      if (kMasked) {
        if (2 * i >= valid_cols) e0 = 0.f;
        if (2 * i + 1 >= valid_cols) e1 = 0.f;
      }
      pk[k] = pack_bf16x2(e0, e1);
      lsum[k & 1] = add2(lsum[k & 1], pack2(e0, e1));
    }
    if (!(dbg & 2)) tmem_st8(p_addr + h * 8, pk);
  }
}

template <int NB>
__global__ void __launch_bounds__(ATT_THREADS, NB == 1 ? 2 : 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttKernelArgs a) {
  using C = AttCfg<NB>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sRing = smem + C::Q_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + C::SLOTS * C::SLOT_BYTES);
  uint64_t* ring_full = bars;              // [SLOTS]
  uint64_t* ring_empty = bars + C::SLOTS;  // [SLOTS]
  uint64_t* q_full = bars + 2 * C::SLOTS;
  uint64_t* s_full = q_full + 1;   // [2]
  uint64_t* s_free = q_full + 3;   // [2]
  uint64_t* p_ready = q_full + 5;  // [2]
  uint64_t* pv_done = q_full + 7;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 9);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // role branches are warp-uniform
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int bh = blockIdx.y;
  const int b = bh / a.heads;
  const int hd = bh - b * a.heads;
  const int q_row0 = b * a.seq_q;    // first query-token row of this batch
  const int kv_row0 = b * a.seq_kv;  // first key/value-token row of this batch
  const int col0 = hd * C::D;        // first column of this head inside the q/k/v slice
  const int n_tiles = a.n_kv_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::SLOTS; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], 1);
    }
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_ready[i], 4);
      mbar_init(&pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();
  pdl_launch_dependents();

  // Producer and MMA issuer run with the whole warp in uniform control flow and elect one lane around the asynchronous
  // instructions only (single UTMALDG / UTCHMMA / UTCBAR instructions with uniform-register operands; see gemm_umma.cu).
  if (warp == 4) {
    // ============================ TMA producer ============================
    if (elect_one()) {
      tma_prefetch_desc(&tmap_q);
      tma_prefetch_desc(&tmap_k);
      tma_prefetch_desc(&tmap_v);
      mbar_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
        tma_load_2d(sQ + nb * QTILE_BYTES, &tmap_q, q_full, col0 + nb * 64, q_row0 + q_tile * BLOCK_Q);
    }
    __syncwarp();
    int slot = 0;
    uint32_t phase = 0;
    auto load_tile = [&](bool is_v, int j) {
      mbar_wait(&ring_empty[slot], phase ^ 1);
      uint8_t* dst = sRing + slot * C::SLOT_BYTES;
      if (elect_one()) {
        if (D4D_DBG(a, 16)) {  // ablation: no K/V traffic
          mbar_arrive(&ring_full[slot]);
        } else {
          mbar_expect_tx(&ring_full[slot], C::SLOT_BYTES);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            if (is_v) tma_load_2d(dst + nb * KTILE_BYTES, &tmap_v, &ring_full[slot], col0 + nb * 64, kv_row0 + j * BLOCK_KV);
            else tma_load_2d(dst + nb * KTILE_BYTES, &tmap_k, &ring_full[slot], col0 + nb * 64, kv_row0 + j * BLOCK_KV);
          }
        }
      }
      __syncwarp();
      if (++slot == C::SLOTS) { slot = 0; phase ^= 1; }
    };
    // same order as the MMA warp consumes: K0 K1, then for every j: V(j), K(j+2)
    load_tile(false, 0);
    if (n_tiles > 1) load_tile(false, 1);
    for (int j = 0; j < n_tiles; ++j) {
      load_tile(true, j);
      if (j + 2 < n_tiles) load_tile(false, j + 2);
    }
  } else if (warp == 5) {
    // ============================ MMA issuer ============================
    const uint32_t idesc_qk = make_idesc_bf16(BLOCK_Q, BLOCK_KV, 0, 0);
    const uint32_t idesc_pv = make_idesc_bf16(BLOCK_Q, C::D, 0, 1);
    const uint32_t o_tmem = tmem + C::COL_O;
    const uint32_t q_addr = smem_u32(sQ);
    const uint32_t ring_addr = smem_u32(sRing);
    int slot = 0;
    uint32_t phase = 0;
    auto wait_ahead = [&](int k) {  // wait for the k-th next ring slot without consuming it
      int sl = slot + k;
      uint32_t ph = phase;
      if (sl >= C::SLOTS) { sl -= C::SLOTS; ph ^= 1; }
      mbar_wait(&ring_full[sl], ph);
    };
    auto issue_qk = [&](int j, int sl) {  // S[j&1] = Q K_j^T   (operands already waited for; elected lane only)
      const uint32_t kaddr = ring_addr + sl * C::SLOT_BYTES;
      const uint32_t s_tmem = tmem + C::COL_S + (j & 1) * 64;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = make_smem_desc(q_addr + nb * QTILE_BYTES + k * 32, 0, 1024, 2);
          const uint64_t bd = make_smem_desc(kaddr + nb * KTILE_BYTES + k * 32, 0, 1024, 2);
          if (!D4D_DBG(a, 8)) umma_ss(s_tmem, ad, bd, idesc_qk, (nb | k) != 0 ? 1u : 0u);
        }
      }
    };
    auto issue_pv = [&](int j, int sl) {  // O += P[j&1] V_j     (operands already waited for; elected lane only)
      const uint32_t vaddr = ring_addr + sl * C::SLOT_BYTES;
      const uint32_t p_tmem = tmem + C::COL_P + (j & 1) * 32;
#pragma unroll
      for (int k = 0; k < BLOCK_KV / 16; ++k) {
        // V tile = NB boxes of [64 keys][64 d] (d contiguous): MN-major B operand.
        // 16 keys = two 8-row swizzle atoms = 2048 bytes; SBO = 1024 (next 8 keys), LBO = next 64-wide d block
        const uint64_t bd = make_smem_desc(vaddr + k * 2048, KTILE_BYTES, 1024, 2);
        if (!D4D_DBG(a, 4)) umma_ts(o_tmem, p_tmem + k * 8, bd, idesc_pv, (j | k) != 0 ? 1u : 0u);
      }
    };
    auto advance = [&]() {
      const int used = slot;
      if (++slot == C::SLOTS) { slot = 0; phase ^= 1; }
      return used;
    };
    mbar_wait(q_full, 0);
    for (int j0 = 0; j0 < 2 && j0 < n_tiles; ++j0) {  // prologue: S(0), S(1)
      wait_ahead(0);
      const int k_slot = advance();
      if (elect_one()) {
        issue_qk(j0, k_slot);
        umma_commit(&s_full[j0]);
        umma_commit(&ring_empty[k_slot]);
      }
      __syncwarp();
    }
    for (int j = 0; j < n_tiles; ++j) {
      // operands of this iteration, waited for BEFORE the critical-path barrier
      const bool more = j + 2 < n_tiles;
      wait_ahead(0);            // V(j)
      if (more) wait_ahead(1);  // K(j+2)
      const int v_slot = advance();
      const int k_slot = more ? advance() : 0;
      // P(j) ready also means S(j) has been consumed.  P.V(j) is issued BEFORE Q.K(j+2): tcgen05.commit covers all
      // earlier MMAs of this thread, so s_full(j+2) doubles as "P.V(j) done" and the softmax warps need no separate
      // wait before reusing the P buffer two tiles later.
      pready_sync(j & 1);
      tc_fence_after();  // P was written with tcgen05.st by the softmax warps
      if (elect_one()) {
        issue_pv(j, v_slot);
        if (more) {
          issue_qk(j + 2, k_slot);
          umma_commit(&s_full[j & 1]);  // the commit the softmax warps wait for goes first; it covers P.V(j) too
          umma_commit(&pv_done[j & 1]);
          umma_commit(&ring_empty[v_slot]);
          umma_commit(&ring_empty[k_slot]);
        } else {
          umma_commit(&pv_done[j & 1]);
          umma_commit(&ring_empty[v_slot]);
        }
      }
      __syncwarp();
    }
  } else {
    // ============================ softmax / correction / epilogue ============================
    const int r = warp * 32 + lane;  // query row of this thread = TMEM lane
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t o_tmem = tmem + C::COL_O + lane_sel;
    const int qrow = q_tile * BLOCK_Q + r;  // row within the sequence
    const uint64_t sc2 = pack2(a.scale_log2, a.scale_log2);
    float m = -INFINITY;     // running max (scaled log2 units) used for the exponentials of the next tile
    float l = 0.f;           // running row sum
    float alpha_pend = 1.f;  // pending rescale of O and l (applied once P.V of the previous tile has landed)

    auto rescale_o = [&](float alpha) {  // O[row, :] *= alpha (warp-collective; alpha is per lane/row)
#pragma unroll 1
      for (int c = 0; c < C::D; c += 16) {
        uint32_t v[16];
        tmem_ld16(o_tmem + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
        tmem_st16(o_tmem + c, v);
      }
    };

    // S(j) lives in `sv`; the TMEM loads of S(j+1) are issued at the end of iteration j (after the last use of sv) so
    // that their latency overlaps the P-store drain and the barrier traffic of tile j (software pipelining, no extra regs)
    uint32_t sv[64];
    mbar_wait(&s_full[0], 0);
    tc_fence_after();
    tmem_ld32(tmem + C::COL_S + lane_sel, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
    tmem_ld32(tmem + C::COL_S + lane_sel + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));

    for (int j = 0; j < n_tiles; ++j) {
      const int valid = min(BLOCK_KV, a.seq_kv - j * BLOCK_KV);
      const bool full = valid == BLOCK_KV;
      const int buf = j & 1;
      const uint32_t par = (j >> 1) & 1;
      const uint32_t p_tmem = tmem + C::COL_P + buf * 32 + lane_sel;
      tmem_ld_wait();  // S(j) is in registers
      // Q.K runs two tiles ahead, so S(j+1) is normally complete by now: probe it here (non-blocking) and skip the
      // ~200-cycle try_wait at the end of the iteration when the probe succeeded
      const bool next_ready = (j + 1 < n_tiles) && __all_sync(0xffffffffu, mbar_test(&s_full[buf ^ 1], ((j + 1) >> 1) & 1));

      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      if (D4D_DBG(a, 64)) {  // ablation: barriers only
        tc_fence_before();
        pready_arrive(buf);
        if (j + 1 < n_tiles) {
          mbar_wait(&s_full[buf ^ 1], ((j + 1) >> 1) & 1);
          tc_fence_after();
        }
        continue;
      }
      if (full) { max32<false>(sv, mx, 32); max32<false>(sv + 32, mx, 32); }
      else { max32<true>(sv, mx, valid); max32<true>(sv + 32, mx, valid - 32); }
      const float tmax = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])) * a.scale_log2;

      if (__any_sync(0xffffffffu, alpha_pend != 1.f)) {  // rescale decided at the end of tile j-1
        mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);  // O holds P.V(0..j-1)
        tc_fence_after();
        rescale_o(alpha_pend);
        l *= alpha_pend;
        alpha_pend = 1.f;
      }
      // P[buf] was last read by P.V(j-2), which completed before s_full(j) fired (MMA issue order + commit semantics)
      uint64_t lsum[2] = {0ull, 0ull};
      if (j > 0) {  // speculative P with the running max of the previous tiles
        const uint64_t nm2 = pack2(-m, -m);
        if (full) { exp32<false>(sv, sc2, nm2, lsum, p_tmem, 32, D4D_DBGV(a)); exp32<false>(sv + 32, sc2, nm2, lsum, p_tmem + 16, 32, D4D_DBGV(a)); }
        else { exp32<true>(sv, sc2, nm2, lsum, p_tmem, valid); exp32<true>(sv + 32, sc2, nm2, lsum, p_tmem + 16, valid - 32); }
      }
      const bool ovf = !(tmax <= m + 100.f);  // would overflow with the old max; always true for tile 0 (m = -inf)
      if (__any_sync(0xffffffffu, ovf)) {
        // exact path: new max >= every logit of this tile; rescale history, recompute P from the registers
        const float m_new = fmaxf(m, tmax);
        if (j > 0) {
          const float alpha = ex2_approx(m - m_new);  // m is finite for j > 0
          if (__any_sync(0xffffffffu, alpha != 1.f)) {
            mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
            tc_fence_after();
            rescale_o(alpha);
          }
          l *= alpha;
        }
        m = m_new;
        const uint64_t nm2 = pack2(-m, -m);
        lsum[0] = 0ull;
        lsum[1] = 0ull;
        if (full) { exp32<false>(sv, sc2, nm2, lsum, p_tmem, 32, D4D_DBGV(a)); exp32<false>(sv + 32, sc2, nm2, lsum, p_tmem + 16, 32, D4D_DBGV(a)); }
        else { exp32<true>(sv, sc2, nm2, lsum, p_tmem, valid); exp32<true>(sv + 32, sc2, nm2, lsum, p_tmem + 16, valid - 32); }
      } else if (tmax > m + 8.f) {
        // lazy rescale: this tile used the old max; fold the change into O and l before the next tile
        alpha_pend = ex2_approx(m - tmax);
        m = tmax;
      }
      {
        float s0, s1, s2, s3;
        unpack2(lsum[0], s0, s1);
        unpack2(lsum[1], s2, s3);
        l += (s0 + s1) + (s2 + s3);
      }
      tmem_st_wait();
      tc_fence_before();
      pready_arrive(buf);  // hand P(j) to the MMA warp FIRST: the wait below must not delay P.V(j) / Q.K(j+2)
      if (j + 1 < n_tiles) {  // then fetch S(j+1) (Q.K runs two tiles ahead); its latency overlaps the hand-off
        if (!next_ready) mbar_wait(&s_full[buf ^ 1], ((j + 1) >> 1) & 1);
        tc_fence_after();
        const uint32_t s_next = tmem + C::COL_S + (buf ^ 1) * 64 + lane_sel;
        if (!D4D_DBG(a, 32)) {
          tmem_ld32(s_next, *reinterpret_cast<uint32_t(*)[32]>(&sv[0]));
          tmem_ld32(s_next + 32, *reinterpret_cast<uint32_t(*)[32]>(&sv[32]));
        }
      }
    }
    tmem_ld_wait();
    // ---- epilogue: O / l -> bf16 -> global (a pending rescale multiplies O and l alike: skipped) ----
    mbar_wait(&pv_done[(n_tiles - 1) & 1], ((n_tiles - 1) >> 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    bf16* orow = a.out + static_cast<size_t>(q_row0 + qrow) * a.ld_out + col0;
#pragma unroll 1
    for (int c = 0; c < C::D; c += 16) {
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
  if (warp == 5) {
    tc_fence_after();
    tmem_dealloc(tmem, C::TMEM_COLS);
  }
}

template <int NB>
int launch_attn(const AttnLaunch& L, cudaStream_t stream) {
  using C = AttCfg<NB>;
  static PerDeviceOnce attr_once;
  if (int rc = ensure_dyn_smem(attn_fwd_kernel<NB>, C::SMEM_BYTES, attr_once)) return rc;
  AttKernelArgs a;
  a.seq_q = L.d.seq;
  a.seq_kv = L.d.seq_kv > 0 ? L.d.seq_kv : L.d.seq;
  a.heads = L.d.heads;
  a.n_kv_tiles = (a.seq_kv + BLOCK_KV - 1) / BLOCK_KV;
  a.scale_log2 = L.d.scale * 1.4426950408889634f;
  a.out = L.d.out;
  a.ld_out = L.d.ld_out;
  a.dbg = 0;
#ifdef D4D_ABLATE
  a.dbg = ablate_env("D4D_ATTN_ABLATE");
#endif
  dim3 grid(L.grid_x, L.grid_y);
  D4D_CUDA_OK(launch_pdl(attn_fwd_kernel<NB>, grid, dim3(ATT_THREADS), C::SMEM_BYTES, stream, L.tmap_q, L.tmap_k, L.tmap_v, a));
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

// tools build only: D4D_ATTN_GENERIC=1 routes head_dim 64 through the generic kernel (A/B timing, tools/bench_attention.py)
static inline bool attn_generic_forced() {
#ifdef D4D_ABLATE
  const char* e = getenv("D4D_ATTN_GENERIC");
  return e && e[0] == '1';
#else
  return false;
#endif
}

// head_dim 64 has its own kernel (attention_d64.cu): 256-row CTAs, 128-key tiles, split MUFU / FMA-pipe exp2
int attn64_prepare(const AttnDesc& d, AttnLaunch* L);
int attn64_run(const AttnLaunch& L, cudaStream_t stream);

int attn_prepare(const AttnDesc& d, AttnLaunch* L) {
  D4D_REQUIRE(d.head_dim == 64 || d.head_dim == 128 || d.head_dim == 192,
              "attention head_dim (after padding) must be 64, 128 or 192");
  D4D_REQUIRE(d.batch > 0 && d.seq > 0 && d.heads > 0 && d.seq_kv >= 0, "empty attention problem");
  D4D_REQUIRE(d.ld_qkv % 8 == 0 && d.ld_out % 8 == 0 && d.ld_kv % 8 == 0, "leading dimensions must be multiples of 8");
  D4D_REQUIRE(d.scale > 0.f, "softmax scale must be positive");
  L->d = d;
  if (d.head_dim == 64 && !attn_generic_forced()) {
    if (int rc = attn64_prepare(d, L)) return rc;  // sets variant 0 / -1 (kernel shape)
    D4D_REQUIRE(L->grid_y <= 65535, "batch*heads exceeds grid.y limit");
    return 0;
  }
  L->variant = d.head_dim / 64;
  const int seq_kv = d.seq_kv > 0 ? d.seq_kv : d.seq;
  const int ld_kv = d.ld_kv > 0 ? d.ld_kv : d.ld_qkv;
  const uint64_t q_tokens = static_cast<uint64_t>(d.batch) * d.seq;
  const uint64_t kv_tokens = static_cast<uint64_t>(d.batch) * seq_kv;
  const uint64_t width = static_cast<uint64_t>(d.heads) * d.head_dim;
  if (int rc = make_tmap_2d(&L->tmap_q, d.q, q_tokens, width, d.ld_qkv, 64, BLOCK_Q, 128)) return rc;
  if (int rc = make_tmap_2d(&L->tmap_k, d.k, kv_tokens, width, ld_kv, 64, BLOCK_KV, 128)) return rc;
  if (int rc = make_tmap_2d(&L->tmap_v, d.v, kv_tokens, width, ld_kv, 64, BLOCK_KV, 128)) return rc;
  L->grid_x = (d.seq + BLOCK_Q - 1) / BLOCK_Q;
  L->grid_y = d.batch * d.heads;
  D4D_REQUIRE(L->grid_y <= 65535, "batch*heads exceeds grid.y limit");
  return 0;
}

int attn_run(const AttnLaunch& L, cudaStream_t stream) {
  switch (L.variant) {
    case 0:
    case -1: return attn64_run(L, stream);
    case 1: return launch_attn<1>(L, stream);
    case 2: return launch_attn<2>(L, stream);
    case 3: return launch_attn<3>(L, stream);
  }
  set_error("attention: unsupported head_dim variant");
  return 1;
}

double attn_flops(const AttnDesc& d) {
  const double skv = d.seq_kv > 0 ? d.seq_kv : d.seq;
  return 4.0 * d.batch * d.heads * static_cast<double>(d.seq) * skv * d.head_dim;
}

}  // namespace d4d
