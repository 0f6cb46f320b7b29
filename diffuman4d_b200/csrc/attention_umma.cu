// tcgen05 flash-attention forward for sm_100a (no mask, non-causal; the only attention on the path).
//
//   O[b, i, h, :] = softmax_j(scale * Q[b,i,h,:] . K[b,j,h,:]) V[b,j,h,:]
//
// Replaces F.scaled_dot_product_attention as called by diffusers' AttnProcessor2_0 for attn1 (3-D: the
// tokens of all F frames of one CFG half form ONE sequence, reference src/diffusers/models/attention.py:68-83)
// and attn2 (per-image).  Q/K/V are read in place from the fused QKV-GEMM output [tokens, 3C] through
// strided TMA boxes, so "(b t) hw c -> b (t hw) c" and the head split are address arithmetic only.
//
// One CTA = 128 query rows of one (batch, head).  Warp roles (192 threads):
//   warps 0-3  softmax: thread t owns query row t (TMEM lane t): S -> max/exp2/sum -> P (bf16) back to TMEM,
//              lazy O rescale (only when the running max grows by > 8 in log2 units), final O/l store
//   warp  4    TMA producer: Q once, then K_0, K_1, V_0, K_2, V_1, ... through a ring of 16 KB*NB slots
//   warp  5    MMA issuer:   S = Q K_j^T (SS, both K-major), O += P V_j (A = P from TMEM, B = V MN-major)
// S(j+1) is issued as soon as the softmax warps have pulled S(j) into registers, so the tensor core runs
// under the exp phase; with head_dim 64 two CTAs are co-resident per SM (256 TMEM columns, ~97 KB smem each).
#include <math.h>

#include "kernels.h"

namespace d4d {

namespace {

constexpr int BLOCK_Q = 128;
constexpr int BLOCK_KV = 128;
constexpr int ATT_THREADS = 192;
constexpr int TILE_BYTES = 128 * 64 * 2;  // one [128 rows][64 ch] swizzled box

template <int NB>
struct AttCfg {
  static constexpr int D = 64 * NB;
  static constexpr int SLOT_BYTES = TILE_BYTES * NB;           // one K or V tile
  static constexpr int SLOTS = NB == 3 ? 3 : 5;
  static constexpr int Q_BYTES = TILE_BYTES * NB;
  static constexpr int SMEM_BYTES = Q_BYTES + SLOTS * SLOT_BYTES + 1024 + 256;
  static constexpr int TMEM_COLS = NB == 1 ? 256 : 512;
  static constexpr int COL_S = 0, COL_P = 128, COL_O = 192;
};

struct AttKernelArgs {
  int seq, heads, n_kv_tiles;
  float scale_log2;
  bf16* out;
  int ld_out;
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
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
  uint64_t* ring_full = bars;                  // [SLOTS]
  uint64_t* ring_empty = bars + C::SLOTS;      // [SLOTS]
  uint64_t* q_full = bars + 2 * C::SLOTS;
  uint64_t* s_full = q_full + 1;
  uint64_t* s_free = q_full + 2;
  uint64_t* p_ready = q_full + 3;
  uint64_t* pv_done = q_full + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_full + 5);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int bh = blockIdx.y;
  const int b = bh / a.heads;
  const int hd = bh - b * a.heads;
  const int row0 = b * a.seq;               // first token row of this batch in the [tokens, ld] matrix
  const int col0 = hd * C::D;               // first column of this head inside the q/k/v slice
  const int n_tiles = a.n_kv_tiles;

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::SLOTS; ++i) {
      mbar_init(&ring_full[i], 1);
      mbar_init(&ring_empty[i], 1);
    }
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_ready, 4);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 5) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 4) {
    // ============================ TMA producer ============================
    if (lane == 0) {
      tma_prefetch_desc(&tmap_q);
      tma_prefetch_desc(&tmap_k);
      tma_prefetch_desc(&tmap_v);
      mbar_expect_tx(q_full, C::Q_BYTES);
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
        tma_load_2d(sQ + nb * TILE_BYTES, &tmap_q, q_full, col0 + nb * 64, row0 + q_tile * BLOCK_Q);
      int slot = 0;
      uint32_t phase = 0;
      // order of tiles through the ring: K0, K1, V0, K2, V1, ..., K(n-1), V(n-2), V(n-1)
      for (int step = 0; step < 2 * n_tiles; ++step) {
        int is_v, j;
        if (step == 0) { is_v = 0; j = 0; }
        else if (step == 2 * n_tiles - 1) { is_v = 1; j = n_tiles - 1; }
        else { is_v = (step & 1) ? 0 : 1; j = is_v ? (step / 2 - 1) : ((step + 1) / 2); }
        mbar_wait(&ring_empty[slot], phase ^ 1);
        uint8_t* dst = sRing + slot * C::SLOT_BYTES;
        mbar_expect_tx(&ring_full[slot], C::SLOT_BYTES);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
          tma_load_2d(dst + nb * TILE_BYTES, is_v ? &tmap_v : &tmap_k, &ring_full[slot], col0 + nb * 64,
                      row0 + j * BLOCK_KV);
        if (++slot == C::SLOTS) { slot = 0; phase ^= 1; }
      }
    }
  } else if (warp == 5) {
    // ============================ MMA issuer ============================
    if (lane == 0) {
      const uint32_t idesc_qk = make_idesc_bf16(BLOCK_Q, BLOCK_KV, 0, 0);
      const uint32_t idesc_pv = make_idesc_bf16(BLOCK_Q, C::D, 0, 1);
      const uint32_t s_tmem = tmem + C::COL_S, p_tmem = tmem + C::COL_P, o_tmem = tmem + C::COL_O;
      int slot = 0;
      uint32_t phase = 0;
      auto issue_qk = [&]() {
        mbar_wait(&ring_full[slot], phase);
        tc_fence_after();
        const uint32_t kaddr = smem_u32(sRing + slot * C::SLOT_BYTES);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_smem_desc(smem_u32(sQ) + nb * TILE_BYTES + k * 32, 0, 1024, 2);
            const uint64_t bd = make_smem_desc(kaddr + nb * TILE_BYTES + k * 32, 0, 1024, 2);
            umma_ss(s_tmem, ad, bd, idesc_qk, (nb | k) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&ring_empty[slot]);
        umma_commit(s_full);
        if (++slot == C::SLOTS) { slot = 0; phase ^= 1; }
      };
      auto issue_pv = [&](int j) {
        mbar_wait(&ring_full[slot], phase);
        tc_fence_after();
        const uint32_t vaddr = smem_u32(sRing + slot * C::SLOT_BYTES);
#pragma unroll
        for (int k = 0; k < BLOCK_KV / 16; ++k) {
          // V tile = NB boxes of [128 keys][64 d] (d contiguous): MN-major B operand.
          // 16 keys = two 8-row swizzle atoms = 2048 bytes; SBO = 1024 (next 8 keys), LBO = next 64-wide d block
          const uint64_t bd = make_smem_desc(vaddr + k * 2048, TILE_BYTES, 1024, 2);
          umma_ts(o_tmem, p_tmem + k * 8, bd, idesc_pv, (j | k) != 0 ? 1u : 0u);
        }
        umma_commit(&ring_empty[slot]);
        umma_commit(pv_done);
        if (++slot == C::SLOTS) { slot = 0; phase ^= 1; }
      };
      mbar_wait(q_full, 0);
      tc_fence_after();
      issue_qk();  // S(0)
      for (int j = 0; j < n_tiles; ++j) {
        if (j + 1 < n_tiles) {
          mbar_wait(s_free, j & 1);  // softmax has S(j) in registers
          tc_fence_after();
          issue_qk();                // S(j+1)
        }
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        issue_pv(j);
      }
    }
  } else {
    // ============================ softmax / correction / epilogue ============================
    const uint32_t lane_sel = static_cast<uint32_t>(warp * 32) << 16;
    const uint32_t s_tmem = tmem + C::COL_S + lane_sel;
    const uint32_t p_tmem = tmem + C::COL_P + lane_sel;
    const uint32_t o_tmem = tmem + C::COL_O + lane_sel;
    const int qrow = q_tile * BLOCK_Q + threadIdx.x;  // row within the sequence
    float m = -INFINITY;  // running max, in scaled log2 units
    float l = 0.f;
    for (int j = 0; j < n_tiles; ++j) {
      const int valid = min(BLOCK_KV, a.seq - j * BLOCK_KV);
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      // ---- pass 1: row max ----
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(s_tmem + c * 32, v);
        tmem_ld_wait();
        if (valid == BLOCK_KV) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i < valid) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      mx *= a.scale_log2;
      const bool grow = mx > m + 8.0f;
      const float m_new = grow ? mx : m;
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // P and O are free to touch
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {
          const float alpha = grow ? ex2_approx(m - m_new) : 1.0f;
          l *= alpha;
#pragma unroll 1
          for (int c = 0; c < C::D / 16; ++c) {
            uint32_t v[16];
            tmem_ld16(o_tmem + c * 16, v);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
            tmem_st16(o_tmem + c * 16, v);
          }
        }
      }
      m = m_new;
      // ---- pass 2: P = exp2(S*scale - m), row sum, bf16 pack -> TMEM ----
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        uint32_t v[32];
        tmem_ld32(s_tmem + c * 32, v);
        tmem_ld_wait();
        if (c == 3) {
          // S(j) is fully in registers: let the MMA warp overwrite it with S(j+1)
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);
        }
        float p[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float e = ex2_approx(fmaf(__uint_as_float(v[i]), a.scale_log2, -m));
          if (valid != BLOCK_KV && c * 32 + i >= valid) e = 0.f;
          p[i] = e;
          l += e;
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[i] = pack_bf16x2(p[2 * i], p[2 * i + 1]);
        tmem_st16(p_tmem + c * 16, pk);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    // ---- epilogue: O / l -> bf16 -> global ----
    mbar_wait(pv_done, (n_tiles - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    bf16* orow = a.out + static_cast<size_t>(row0 + qrow) * a.ld_out + col0;
#pragma unroll 1
    for (int c = 0; c < C::D / 16; ++c) {
      uint32_t v[16];
      tmem_ld16(o_tmem + c * 16, v);
      tmem_ld_wait();
      if (qrow < a.seq) {
        uint4 o0, o1;
        o0.x = pack_bf16x2(__uint_as_float(v[0]) * inv_l, __uint_as_float(v[1]) * inv_l);
        o0.y = pack_bf16x2(__uint_as_float(v[2]) * inv_l, __uint_as_float(v[3]) * inv_l);
        o0.z = pack_bf16x2(__uint_as_float(v[4]) * inv_l, __uint_as_float(v[5]) * inv_l);
        o0.w = pack_bf16x2(__uint_as_float(v[6]) * inv_l, __uint_as_float(v[7]) * inv_l);
        o1.x = pack_bf16x2(__uint_as_float(v[8]) * inv_l, __uint_as_float(v[9]) * inv_l);
        o1.y = pack_bf16x2(__uint_as_float(v[10]) * inv_l, __uint_as_float(v[11]) * inv_l);
        o1.z = pack_bf16x2(__uint_as_float(v[12]) * inv_l, __uint_as_float(v[13]) * inv_l);
        o1.w = pack_bf16x2(__uint_as_float(v[14]) * inv_l, __uint_as_float(v[15]) * inv_l);
        uint4* op = reinterpret_cast<uint4*>(orow + c * 16);
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
  static bool attr_set[64] = {};
  if (int rc = ensure_dyn_smem(attn_fwd_kernel<NB>, C::SMEM_BYTES, attr_set)) return rc;
  AttKernelArgs a;
  a.seq = L.d.seq;
  a.heads = L.d.heads;
  a.n_kv_tiles = (L.d.seq + BLOCK_KV - 1) / BLOCK_KV;
  a.scale_log2 = L.d.scale * 1.4426950408889634f;
  a.out = L.d.out;
  a.ld_out = L.d.ld_out;
  dim3 grid(L.grid_x, L.grid_y);
  attn_fwd_kernel<NB><<<grid, ATT_THREADS, C::SMEM_BYTES, stream>>>(L.tmap_q, L.tmap_k, L.tmap_v, a);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace

int attn_prepare(const AttnDesc& d, AttnLaunch* L) {
  D4D_REQUIRE(d.head_dim == 64 || d.head_dim == 128 || d.head_dim == 192,
              "attention head_dim (after padding) must be 64, 128 or 192");
  D4D_REQUIRE(d.batch > 0 && d.seq > 0 && d.heads > 0, "empty attention problem");
  D4D_REQUIRE(d.ld_qkv % 8 == 0 && d.ld_out % 8 == 0, "leading dimensions must be multiples of 8");
  D4D_REQUIRE(d.scale > 0.f, "softmax scale must be positive");
  L->d = d;
  L->variant = d.head_dim / 64;
  const uint64_t tokens = static_cast<uint64_t>(d.batch) * d.seq;
  const uint64_t width = static_cast<uint64_t>(d.heads) * d.head_dim;
  if (int rc = make_tmap_2d(&L->tmap_q, d.q, tokens, width, d.ld_qkv, 64, BLOCK_Q, 128)) return rc;
  if (int rc = make_tmap_2d(&L->tmap_k, d.k, tokens, width, d.ld_qkv, 64, BLOCK_KV, 128)) return rc;
  if (int rc = make_tmap_2d(&L->tmap_v, d.v, tokens, width, d.ld_qkv, 64, BLOCK_KV, 128)) return rc;
  L->grid_x = (d.seq + BLOCK_Q - 1) / BLOCK_Q;
  L->grid_y = d.batch * d.heads;
  D4D_REQUIRE(L->grid_y <= 65535, "batch*heads exceeds grid.y limit");
  return 0;
}

int attn_run(const AttnLaunch& L, cudaStream_t stream) {
  switch (L.variant) {
    case 1: return launch_attn<1>(L, stream);
    case 2: return launch_attn<2>(L, stream);
    case 3: return launch_attn<3>(L, stream);
  }
  set_error("attention: unsupported head_dim variant");
  return 1;
}

double attn_flops(const AttnDesc& d) {
  return 4.0 * d.batch * d.heads * static_cast<double>(d.seq) * d.seq * d.head_dim;
}

}  // namespace d4d
