// tcgen05 (UMMA) GEMM / implicit-GEMM 3x3 convolution for sm_100a.
//
//   D[M, N] = A[M, K] * W[N, K]^T  (+bias[N]) (+rowvec[image(m), N]) (+residual[M, N])   -> bf16
//
// * operands are bf16, K-major, staged by TMA into 128B-swizzled shared memory (4-stage mbarrier ring)
// * one elected thread issues tcgen05.mma (M=128, N=block_n<=256, K=16) with fp32 accumulators in TMEM
// * accumulators are double-buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the
//   main loop of tile i+1; the kernel is persistent (grid = #SMs), tiles are walked n-fastest so CTAs
//   that run concurrently share the same A rows through L2
// * "conv" mode turns the A loader into an implicit-GEMM gather: the A tile for k-block (tap, c0) is a
//   4-D TMA box {64 ch, BW, BH, BN} of the NHWC activation at spatial offset (ky-1, kx-1); out-of-bounds
//   rows/cols are zero-filled by TMA, which is exactly the conv's zero padding.  K = 9*Cin, weights are
//   pre-laid-out as [Cout][tap][Cin].
// * "two-source" mode reads the first kb_split k-blocks from A and the rest from A2 (a channel concat
//   that is never materialised: resnet shortcut 1x1 conv over [hidden | skip]).
// * epilogue options: +bias (fp32), +per-image row vector (time-embedding projection), +residual,
//   GEGLU (a * gelu_erf(g), weights pre-interleaved so one N tile holds matching a/g columns).
//
// Replaces, on the reference path: F.linear / 1x1 conv / 3x3 conv calls of diffusers ResnetBlock2D,
// Attention, FeedForward, Transformer2DModel (reference call sites: unet_multiview_blocks.py:274,423,585,
// transformer_multiview.py:46-77, attention.py:73,116,142).
#include "kernels.h"

namespace d4d {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;
constexpr int EPI_WARPS_PER_QUARTER = 3;
constexpr int MAX_STAGES = 8;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;     // 16 KB
constexpr int BAR_BYTES = 1024;                    // barrier block in front of the ring keeps the stages 1024-B aligned
constexpr int SMEM_BYTES = 227 * 1024;             // everything an SM has: the ring gets as many stages as fit
// Output staging: every epilogue warp owns 32 rows x 64 B.  A thread holds one ROW of the accumulator (TMEM lane), so a
// direct store makes each warp-level st.global touch 32 different lines (measured: the stores were 50% of the K = 320
// GEMMs); staged through shared memory the warp stores 8 rows x 64 contiguous bytes per instruction.
// With TMA stores the staged unit IS the source box of a cp.async.bulk.tensor store (the staging swizzle below is the
// 64-byte TMA swizzle), double-buffered per warp so that one store can still be reading while the next unit is staged.
// Only the shortest main loops (K = 320: k_blocks <= 5) get the second staging buffer: it costs a ring stage, and from
// K = 640 on the ring depth matters more than the overlapped store (measured per shape with D4D_GEMM_STG_BUFS in the tools
// build: K = 320 -1..-3 %, K = 640 +5 %, K = 1280 +8 % with two buffers; profiles/README.md).
constexpr int STG_WARP_BYTES = 32 * 64;
constexpr int STG_BYTES1 = 4 * EPI_WARPS_PER_QUARTER * STG_WARP_BYTES;  // 24 KB per set of staging buffers
inline int stg_bufs_for(int block_n, int k_blocks) {
#ifdef D4D_ABLATE
  if (const char* e = getenv("D4D_GEMM_STG_BUFS")) return atoi(e) == 2 ? 2 : 1;  // tools build: A/B the rule below
#endif
  return k_blocks <= 5 ? 2 : 1;
}
__host__ __device__ inline int ring_bytes_for(int stg_bufs) {
  return SMEM_BYTES - 1024 /*align*/ - BAR_BYTES - STG_BYTES1 * stg_bufs;
}
// The main loop is bound by the latency of the TMA loads in flight (measured: ~650-700 cycles per k-block whatever block_n
// is, with 4 x 48 KB stages), so the ring depth follows the tile width: 4 stages at block_n 256 ... 8 at block_n <= 96.
__host__ __device__ inline int stage_bytes_for(int block_n) { return A_BYTES + block_n * BLOCK_K * 2; }
__host__ __device__ inline int stages_for(int block_n, int stg_bufs) {
  const int s = ring_bytes_for(stg_bufs) / stage_bytes_for(block_n);
  return s > MAX_STAGES ? MAX_STAGES : s;
}
constexpr int NUM_THREADS = 64 + 128 * EPI_WARPS_PER_QUARTER;  // warp0 TMA, warp1 MMA(+TMEM alloc), then the epilogue warps
constexpr int TMEM_COLS = 512;
constexpr int ACC_STRIDE = 256;

struct TileCoord {
  int m0;          // plain: first row.  conv: unused
  int n_img0, y0, x0;
  int pa, pb;      // conv, n_phases == 4: sub-pixel phase of this tile (0 otherwise)
};

template <bool kConv>
__device__ __forceinline__ void tile_coords(const GemmKernelArgs& a, int m_tile, TileCoord& t) {
  if (!kConv) {
    t.m0 = m_tile * BLOCK_M;
    t.n_img0 = t.y0 = t.x0 = 0;
    t.pa = t.pb = 0;
  } else {
    const int ph = a.n_phases > 1 ? m_tile / a.tiles_per_phase : 0;  // phases outermost: concurrent CTAs share a weight slab
    m_tile -= ph * a.tiles_per_phase;
    t.pa = ph >> 1;
    t.pb = ph & 1;
    int tx = m_tile % a.tiles_x;
    int r = m_tile / a.tiles_x;
    int ty = r % a.tiles_y;
    int tn = r / a.tiles_y;
    t.m0 = 0;
    t.x0 = tx * a.BW;
    t.y0 = ty * a.BH;
    t.n_img0 = tn * a.BN;
  }
}

__device__ __forceinline__ void add8_bf16(float* f, const uint4& u) {
  float2 p;
  p = unpack_bf16x2(u.x); f[0] += p.x; f[1] += p.y;
  p = unpack_bf16x2(u.y); f[2] += p.x; f[3] += p.y;
  p = unpack_bf16x2(u.z); f[4] += p.x; f[5] += p.y;
  p = unpack_bf16x2(u.w); f[6] += p.x; f[7] += p.y;
}

// Column sums over the 32 rows of a warp for 16 columns held one row per lane: a butterfly that halves the number of
// columns a lane keeps at every step (8 + 4 + 2 + 1 + 1 shuffles).  Returns, in every lane, the total of column
// (lane >> 1) & 15 ... precisely: bit 4 of the lane selects columns 8-15, bit 3 the upper 4 of those, bit 2, bit 1.
__device__ __forceinline__ float colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool up = lane & 16;
    const float send = up ? v[i] : v[i + 8], keep = up ? v[i + 8] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const bool up = lane & 8;
    const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const bool up = lane & 4;
    const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  {
    const bool up = lane & 2;
    const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Epilogue feature bits of the kernel template: a CLEAR bit compiles the feature out, a set bit is still checked at run
// time.  The epilogue runs 12 warps x (block_n / 16) chunks per tile and, for the short-K projections, is what bounds the
// kernel; with every feature behind a run-time flag a 16-column chunk cost ~140 instructions (uniform loads, tests and
// branches around ~30 of real work; ncu source counters, profiles/README.md).  gemm_run picks the instantiation whose bits
// equal the launch's features, or the E_ALL one.
enum : int { E_BIAS = 1, E_ROWVEC = 2, E_ACT = 4, E_RES = 8, E_STATS = 16, E_KV = 32, E_ALL = 63 };

// kConv: implicit-GEMM conv loader / output mapping (a.mode == 1); kTma: TMA-store epilogue (a.tma_store)
template <bool kGeglu, bool kConv, bool kTma, int kEpi>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_a2,
                 const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_out,
                 const GemmKernelArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint8_t* staging = smem + BAR_BYTES;
  const int STG_BUFS = a.stg_bufs;
  smem += BAR_BYTES + STG_BYTES1 * STG_BUFS;
  uint64_t* full = bars;                   // [MAX_STAGES]
  uint64_t* empty = bars + MAX_STAGES;     // [MAX_STAGES]
  uint64_t* tfull = bars + 2 * MAX_STAGES; // [2]
  const int STAGES = stages_for(a.block_n, a.stg_bufs);
  const int STAGE_BYTES = stage_bytes_for(a.block_n);
  uint64_t* tempty = tfull + 2;        // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // tells the compiler the role branches are warp-uniform
  const int lane = threadIdx.x & 31;
  const int total_tiles = a.m_tiles * a.n_tiles;
  const uint32_t b_bytes = static_cast<uint32_t>(a.block_n) * BLOCK_K * 2;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (a.kb_split < a.k_blocks) tma_prefetch_desc(&tmap_a2);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4 * EPI_WARPS_PER_QUARTER);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  pdl_wait();  // barriers / TMEM are set up while the previous kernel drains
  pdl_launch_dependents();

  // The producer and the MMA issuer run their loops with the WHOLE warp in uniform control flow and elect one lane only
  // around the asynchronous instructions: descriptors / coordinates then live in uniform registers and every
  // UTMALDG / UTCHMMA is a single instruction.  (With the loop inside `if (lane == 0)` the compiler wraps each of them
  // in an ELECT + 5x R2UR + branch sequence, and the issuing thread - not the tensor pipe - bounds the main loop:
  // measured ~650-800 cycles per k-block for every block_n, profiles/README.md.)
  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const int m_tile = tile / a.n_tiles;
      const int n_tile = tile % a.n_tiles;
      TileCoord tc;
      tile_coords<kConv>(a, m_tile, tc);
      const int n0 = n_tile * a.block_n;
      for (int kb = 0; kb < a.k_blocks; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = smem + stage * STAGE_BYTES;
        uint8_t* sb = sa + A_BYTES;
        if (D4D_DBG(a, 1)) {  // ablation: no loads (tools build only)
          if (elect_one()) mbar_arrive(&full[stage]);
        } else if (!kConv) {
          const bool first = kb < a.kb_split;
          const int ka = first ? kb * BLOCK_K : (kb - a.kb_split) * BLOCK_K;
          if (elect_one()) {
            mbar_expect_tx(&full[stage], A_BYTES + b_bytes);
            if (first) tma_load_2d(sa, &tmap_a, &full[stage], ka, tc.m0);
            else tma_load_2d(sa, &tmap_a2, &full[stage], ka, tc.m0);
            tma_load_2d(sb, &tmap_b, &full[stage], kb * BLOCK_K, n0);
          }
        } else {
          const int tap = kb / a.cin_blocks;
          const int cb = kb - tap * a.cin_blocks;
          if (elect_one()) {
            mbar_expect_tx(&full[stage], A_BYTES + b_bytes);
            tma_load_4d(sa, &tmap_a, &full[stage], cb * BLOCK_K, tc.x0 * a.in_stride + a.tap_dx[tap] + tc.pb,
                        tc.y0 * a.in_stride + a.tap_dy[tap] + tc.pa, tc.n_img0);
            tma_load_2d(sb, &tmap_b, &full[stage], tap * a.Cin + cb * BLOCK_K, n0 + (tc.pa * 2 + tc.pb) * a.N);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = make_idesc_bf16(BLOCK_M, a.block_n, 0, 0);
    const uint32_t ring = smem_u32(smem);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    bool ready = false;  // full[stage] already seen complete by the probe of the previous k-block
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tempty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
      for (int kb = 0; kb < a.k_blocks; ++kb) {
        // operands were written by TMA: the mbarrier orders them, no tcgen05 fence needed
        if (!ready) mbar_wait(&full[stage], phase);
        const uint32_t sa = ring + stage * STAGE_BYTES;
        const uint64_t adesc = make_smem_desc(sa, 0, 1024, 2);
        const uint64_t bdesc = make_smem_desc(sa + A_BYTES, 0, 1024, 2);
        {  // probe the next stage now; the answer is needed only after this k-block's MMAs have been issued
          const int ns = stage + 1 == STAGES ? 0 : stage + 1;
          const uint32_t np = stage + 1 == STAGES ? phase ^ 1 : phase;
          ready = __all_sync(0xffffffffu, mbar_test(&full[ns], np));
        }
        if (elect_one()) {
          if (!D4D_DBG(a, 2)) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / 16; ++k) {
              // advancing 16 bf16 along K inside the 128B swizzle atom = +32 bytes = +2 in the (addr>>4) field
              umma_ss(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elect_one()) umma_commit(&tfull[acc]);
      __syncwarp();
    }
  } else {
    // ===== epilogue warps; TMEM lane quarter = warp%4; the 3 warps of a quarter interleave 32-column units (two 16-column
    // TMEM chunks), which are staged in shared memory and written out as 64-byte row segments =====
    const int q = warp & 3;
    const int unit0 = (warp - 2) >> 2;
    const int r = q * 32 + lane;  // row within the 128-row tile
    uint8_t* stg_base = staging + (warp - 2) * STG_WARP_BYTES * STG_BUFS;
    uint8_t* stg = stg_base;  // current staging buffer of this warp (alternates per unit)
    int stg_sel = 0;
    // staging layout: row-in-warp * 64 B + (16-byte piece ^ swizzle(row)) = CU_TENSOR_MAP_SWIZZLE_64B; conflict-free for the
    // row-wise writes and for the transposed reads (lane -> row = 8 i + lane / 4, piece = lane % 4)
    uint32_t stg_w = smem_u32(stg) + lane * 64;
    const int sw_w = (lane >> 1) & 3;
    const int t_piece = lane & 3;
    auto chunk_at = [&](int k) { return 2 * (unit0 + EPI_WARPS_PER_QUARTER * (k >> 1)) + (k & 1); };
    auto row_of = [&](const TileCoord& tc, int rr, long long& row, int& img, bool& valid) {
      if (!kConv) {
        row = static_cast<long long>(tc.m0) + rr;
        valid = row < a.M;
        img = a.rows_per_image > 0 ? static_cast<int>(row / a.rows_per_image) : 0;
      } else {
        const int bx = rr % a.BW;
        const int r2 = rr / a.BW;
        const int by = r2 % a.BH;
        const int bn = r2 / a.BH;
        const int x = tc.x0 + bx, y = tc.y0 + by, n = tc.n_img0 + bn;
        valid = (x < a.W) && (y < a.H) && (n < a.n_img);
        row = (static_cast<long long>(n) * a.out_H + (y * a.out_sy + a.out_oy + tc.pa)) * a.out_W + (x * a.out_sx + a.out_ox + tc.pb);
        img = n;
      }
    };
    auto stage_half = [&](int h, const uint4& o0, const uint4& o1) {  // 16 columns (32 B) of this thread's row
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_w + (((2 * h) ^ sw_w) << 4)), "r"(o0.x), "r"(o0.y),
                   "r"(o0.z), "r"(o0.w) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_w + (((2 * h + 1) ^ sw_w) << 4)), "r"(o1.x), "r"(o1.y),
                   "r"(o1.z), "r"(o1.w) : "memory");
    };
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m_tile = tile / a.n_tiles;
      const int n_tile = tile % a.n_tiles;
      TileCoord tc;
      tile_coords<kConv>(a, m_tile, tc);
      const int n0 = n_tile * a.block_n;

      long long row;  // output row (pixel/token index) of this thread's accumulator row
      int img;        // image index for the per-image row vector
      bool valid;
      row_of(tc, r, row, img, valid);
      // rows this lane writes in the transposed (coalesced) store
      long long t_row[4];
      bool t_valid[4];
      auto calc_t_rows = [&]() {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int img_unused;
          row_of(tc, q * 32 + i * 8 + (lane >> 2), t_row[i], img_unused, t_valid[i]);
        }
      };
      if (!kTma) calc_t_rows();  // (TMA-store kernels need them only for an odd last 16-column chunk: computed there)
      // flush one staged unit: `pieces` = 2 (one 16-column chunk) or 4; ocol = first output column of the unit
      auto flush = [&](int pieces, int ocol) {
        if (kTma && pieces == 4) {
          // the staged [32 rows][32 columns] unit is the source box of one TMA store (rows / columns outside the tensor
          // are clipped by the tensor map); the other staging buffer takes the next unit while this one is being read
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && !D4D_DBG(a, 32)) {
            if (!kConv) {
              tma_store_2d(&tmap_out, stg, ocol, tc.m0 + q * 32);
            } else {
              const int r2 = (q * 32) / a.BW;
              tma_store_4d(&tmap_out, stg, ocol, tc.x0, tc.y0 + r2 % a.BH, tc.n_img0 + r2 / a.BH);
            }
            bulk_commit_group();
          }
          if (STG_BUFS == 2) {
            stg_sel ^= 1;
            stg = stg_base + stg_sel * STG_WARP_BYTES;
            stg_w = smem_u32(stg) + lane * 64;
            if (lane == 0) bulk_wait_group_read<1>();  // the store that last read the buffer we switch to has drained it
          } else {
            if (lane == 0) bulk_wait_group_read<0>();  // single buffer: wait until this store has read it
          }
          __syncwarp();
          return;
        }
        __syncwarp();
        if (kTma) calc_t_rows();
        uint4 t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rw = i * 8 + (lane >> 2);
          const uint32_t addr = smem_u32(stg) + rw * 64 + ((t_piece ^ ((rw >> 1) & 3)) << 4);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t[i].x), "=r"(t[i].y), "=r"(t[i].z), "=r"(t[i].w) : "r"(addr) : "memory");
        }
        __syncwarp();  // the next unit may overwrite the staging rows
        if (t_piece < pieces && !D4D_DBG(a, 32)) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (t_valid[i]) *reinterpret_cast<uint4*>(a.out + static_cast<size_t>(t_row[i]) * a.ldo + ocol + t_piece * 8) = t[i];
        }
      };

      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * ACC_STRIDE + (static_cast<uint32_t>(q * 32) << 16);

      // Residual / row-vector loads of the next chunk are issued before the TMEM wait of the current one so that their
      // L2/HBM latency overlaps the TMEM load and the math; bias (L1-resident broadcast) is loaded under the wait.
      // (Fetching the residual two chunks ahead measured no better: 74 vs 69 us on the L0 projection.)
      if (!kGeglu) {
        const int chunks = a.block_n / 16;
        const bool direct = (kEpi & E_KV) && a.kv_world > 0;  // fused K/V scatter keeps the direct row-wise stores
        const bool has_bias = (kEpi & E_BIAS) && a.bias != nullptr, has_rv = (kEpi & E_ROWVEC) && a.rowvec != nullptr;
        const bool has_res = (kEpi & E_RES) && a.residual != nullptr, has_stats = (kEpi & E_STATS) && a.stats != nullptr;
        const bool has_act = (kEpi & E_ACT) && a.act == 1, has_scale = (kEpi & E_ACT) && a.out_scale != 1.0f;
        const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
        uint4 r0 = z4, r1 = z4, v0 = z4, v1 = z4, nr0 = z4, nr1 = z4, nv0 = z4, nv1 = z4;
        const bf16* res_row = has_res ? a.residual + static_cast<size_t>(row) * a.ld_res + n0 : nullptr;
        const bf16* rv_row = has_rv ? a.rowvec + static_cast<size_t>(img) * a.ld_rowvec + n0 : nullptr;
        const bool ld_res = valid && has_res && !D4D_DBG(a, 64), ld_rv = valid && has_rv;
        if (chunk_at(0) < chunks) {
          const int c0 = chunk_at(0);
          if (ld_res) {  // plain loads: the residual may alias the output (in-place add)
            r0 = *reinterpret_cast<const uint4*>(res_row + c0 * 16);
            r1 = *reinterpret_cast<const uint4*>(res_row + c0 * 16 + 8);
          }
          if (ld_rv) {
            v0 = __ldg(reinterpret_cast<const uint4*>(rv_row + c0 * 16));
            v1 = __ldg(reinterpret_cast<const uint4*>(rv_row + c0 * 16 + 8));
          }
        }
        for (int k = 0;; ++k) {
          const int c = chunk_at(k);
          if (c >= chunks) break;
          const int cn = chunk_at(k + 1);
          uint32_t v[16];
          tmem_ld16(taddr + c * 16, v);
          const int col = n0 + c * 16;
          if (cn < chunks) {
            if (ld_res) {
              nr0 = *reinterpret_cast<const uint4*>(res_row + cn * 16);
              nr1 = *reinterpret_cast<const uint4*>(res_row + cn * 16 + 8);
            }
            if (ld_rv) {
              nv0 = __ldg(reinterpret_cast<const uint4*>(rv_row + cn * 16));
              nv1 = __ldg(reinterpret_cast<const uint4*>(rv_row + cn * 16 + 8));
            }
          }
          float4 b4[4];
          if (has_bias) {
            const float4* bp = reinterpret_cast<const float4*>(a.bias + col);
#pragma unroll
            for (int i = 0; i < 4; ++i) b4[i] = __ldg(bp + i);
          }
          tmem_ld_wait();
          float sv[16];  // stats: the rounded outputs of this thread's row (zero for rows outside the tensor)
          if (has_stats) {
#pragma unroll
            for (int i = 0; i < 16; ++i) sv[i] = 0.f;
          }
          if (valid) {
            float f[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]);
            if (has_bias) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                f[4 * i] += b4[i].x; f[4 * i + 1] += b4[i].y; f[4 * i + 2] += b4[i].z; f[4 * i + 3] += b4[i].w;
              }
            }
            if (has_rv) { add8_bf16(f, v0); add8_bf16(f + 8, v1); }
            if (has_act) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] = silu_f(f[i]);
            }
            if (has_scale) {
#pragma unroll
              for (int i = 0; i < 16; ++i) f[i] *= a.out_scale;
            }
            if (has_res) { add8_bf16(f, r0); add8_bf16(f + 8, r1); }
            if (has_stats) {  // statistics of what is stored: round to bf16 first
#pragma unroll
              for (int i = 0; i < 16; ++i) sv[i] = f[i] = __bfloat162float(__float2bfloat16_rn(f[i]));
            }
            uint4 o0, o1;
            o0.x = pack_bf16x2(f[0], f[1]);   o0.y = pack_bf16x2(f[2], f[3]);
            o0.z = pack_bf16x2(f[4], f[5]);   o0.w = pack_bf16x2(f[6], f[7]);
            o1.x = pack_bf16x2(f[8], f[9]);   o1.y = pack_bf16x2(f[10], f[11]);
            o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
            if (!direct) {
              stage_half(k & 1, o0, o1);
            } else if (col >= a.kv_col0) {
              // fused all-gather: the K|V columns of the QKV projection go straight into every rank's gathered
              // K/V buffer (peer memory over NVLink; own rank included) at this rank's global token rows
              const long long half = row / a.kv_rows_local;
              const long long grow = half * a.kv_rows_global + a.kv_row_offset + (row - half * a.kv_rows_local);
              const size_t off = static_cast<size_t>(grow) * a.kv_ld + (col - a.kv_col0);
#pragma unroll 1
              for (int rk = 0; rk < a.kv_world; ++rk) {
                uint4* dp = reinterpret_cast<uint4*>(a.kv_dst[rk] + off);
                dp[0] = o0;
                dp[1] = o1;
              }
            } else {
              uint4* op = reinterpret_cast<uint4*>(a.out + static_cast<size_t>(row) * a.ldo + col);
              op[0] = o0;
              op[1] = o1;
            }
          }
          if (has_stats) {
            // GroupNorm statistics of the output: column sums over this warp's 32 rows (they belong to ONE image:
            // gemm_prepare checks it), one red.global per (column, moment) and warp.  Lane 0 holds the first row / the
            // tile's minimal (x, y): when it is outside the tensor the whole warp is.
            float sq[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) sq[i] = sv[i] * sv[i];
            const float cs = colsum16(sv, lane), cq = colsum16(sq, lane);
            const int cc = ((lane & 16) ? 8 : 0) + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
            const int img0 = __shfl_sync(0xffffffffu, !kConv ? static_cast<int>(row / a.stats_rows) : img, 0);
            if (__shfl_sync(0xffffffffu, valid ? 1 : 0, 0)) {
              const long long fx = __float2ll_rn((lane & 1) ? cq * kGnSqScale : cs * kGnSumScale);  // fixed point: order-free
              atomicAdd(reinterpret_cast<unsigned long long*>(a.stats) + (static_cast<size_t>(img0) * a.N + col + cc) * 2 + (lane & 1),
                        static_cast<unsigned long long>(fx));
            }
          }
          if (!direct && ((k & 1) || c + 1 >= chunks)) flush((k & 1) ? 4 : 2, n0 + (c & ~1) * 16);
          r0 = nr0; r1 = nr1; v0 = nv0; v1 = nv1;
        }
      } else {
        // GEGLU: this N tile holds [a (block_n/2 cols) | g (block_n/2 cols)] for output cols
        // n_tile*block_n/2 .. +block_n/2
        const int half = a.block_n / 2;
        const int chunks = half / 16;
        const bool has_bias = (kEpi & E_BIAS) && a.bias != nullptr;
        for (int k = 0;; ++k) {
          const int c = chunk_at(k);
          if (c >= chunks) break;
          uint32_t va[16], vg[16];
          tmem_ld16(taddr + c * 16, va);
          tmem_ld16(taddr + half + c * 16, vg);
          float4 ba[4], bg[4];
          if (has_bias) {
            const float4* pa = reinterpret_cast<const float4*>(a.bias + n0 + c * 16);
            const float4* pg = reinterpret_cast<const float4*>(a.bias + n0 + half + c * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) { ba[i] = __ldg(pa + i); bg[i] = __ldg(pg + i); }
          }
          tmem_ld_wait();
          if (valid) {
            float o[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint64_t a01 = f2_pack(__uint_as_float(va[4 * i]), __uint_as_float(va[4 * i + 1]));
              uint64_t a23 = f2_pack(__uint_as_float(va[4 * i + 2]), __uint_as_float(va[4 * i + 3]));
              uint64_t g01 = f2_pack(__uint_as_float(vg[4 * i]), __uint_as_float(vg[4 * i + 1]));
              uint64_t g23 = f2_pack(__uint_as_float(vg[4 * i + 2]), __uint_as_float(vg[4 * i + 3]));
              if (has_bias) {
                a01 = f2_add(a01, f2_pack(ba[i].x, ba[i].y));
                a23 = f2_add(a23, f2_pack(ba[i].z, ba[i].w));
                g01 = f2_add(g01, f2_pack(bg[i].x, bg[i].y));
                g23 = f2_add(g23, f2_pack(bg[i].z, bg[i].w));
              }
              f2_unpack(geglu2(a01, g01), o[4 * i], o[4 * i + 1]);
              f2_unpack(geglu2(a23, g23), o[4 * i + 2], o[4 * i + 3]);
            }
            uint4 o0, o1;
            o0.x = pack_bf16x2(o[0], o[1]);   o0.y = pack_bf16x2(o[2], o[3]);
            o0.z = pack_bf16x2(o[4], o[5]);   o0.w = pack_bf16x2(o[6], o[7]);
            o1.x = pack_bf16x2(o[8], o[9]);   o1.y = pack_bf16x2(o[10], o[11]);
            o1.z = pack_bf16x2(o[12], o[13]); o1.w = pack_bf16x2(o[14], o[15]);
            if (kTma) {
              stage_half(k & 1, o0, o1);  // staged unit -> one TMA store (the one-MUFU GELU left the stores as the bound)
            } else if (!D4D_DBG(a, 32)) {  // direct row-wise stores
              uint4* op = reinterpret_cast<uint4*>(a.out + static_cast<size_t>(row) * a.ldo + n_tile * half + c * 16);
              op[0] = o0;
              op[1] = o1;
            }
          }
          if (kTma && ((k & 1) || c + 1 >= chunks)) flush((k & 1) ? 4 : 2, n_tile * half + (c & ~1) * 16);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
    if (lane == 0) bulk_wait_group<0>();  // TMA stores of this warp have been written before the CTA retires
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace

int gemm_prepare(const GemmDesc& d, GemmLaunch* L) {
  D4D_REQUIRE(d.N % 16 == 0, "GEMM N must be a multiple of 16");
  D4D_REQUIRE(d.out != nullptr && d.A != nullptr && d.Wt != nullptr, "null operand");
  GemmKernelArgs& a = L->args;
  memset(&a, 0, sizeof(a));
  int bn = d.block_n;
  if (bn <= 0) {
    // widest tile that divides N, unless a narrower one wastes fewer SM-waves (persistent grid = #SMs):
    // cost ~ waves * (block_n + fixed per-tile work)
    int dev0 = 0, sms0 = 148;
    cudaGetDevice(&dev0);
    cudaDeviceGetAttribute(&sms0, cudaDevAttrMultiProcessorCount, dev0);
    const long long rows = d.conv ? static_cast<long long>(d.n_img) * d.H * d.W : d.M;
    const long long m_tiles_est = (rows + BLOCK_M - 1) / BLOCK_M;
    const int mult = d.geglu ? 32 : 16;
    bn = gemm_pick_block_n(d.N, mult);
    // K <= 320 (5 k-blocks): the epilogue, not the main loop, paces the tile, and its length is the number of 16-column
    // chunks of the busiest of the 3 warps of a TMEM lane quarter (32-column units dealt round-robin): 6 at block_n 240,
    // 4 at 192 or 160 (measured: L0 qkv 97 us at 240, 94 us at 192, profiles/r02b_tune_block_n.txt)
    const int kb_est = d.conv ? 9 * ((d.Cin + BLOCK_K - 1) / BLOCK_K) : (d.K1 + (d.A2 ? d.K2 : 0) + BLOCK_K - 1) / BLOCK_K;
    const bool epilogue_paced = !d.geglu && kb_est <= 5;
    long long best_cost = -1;
    for (int c = mult; c <= 256; c += mult) {
      if (d.N % c != 0 || (c < 64 && c != bn)) continue;
      const long long tiles = m_tiles_est * (d.N / c);
      const long long waves = (tiles + sms0 - 1) / sms0;
      const int chunks = c / 16, units = (chunks + 1) / 2;
      const int busiest = 2 * ((units + EPI_WARPS_PER_QUARTER - 1) / EPI_WARPS_PER_QUARTER);  // chunks of warp 0 (upper bound)
      const long long cost = waves * ((epilogue_paced ? 40 * (busiest < chunks ? busiest : chunks) : c) + 48);
      if (best_cost < 0 || cost < best_cost || (cost == best_cost && c > bn)) { best_cost = cost; bn = c; }
    }
  }
  D4D_REQUIRE(bn >= 16 && d.N % bn == 0 && bn % (d.geglu ? 32 : 16) == 0, "no valid block_n");
  a.block_n = bn;
#ifdef D4D_ABLATE
  a.dbg = ablate_env("D4D_GEMM_ABLATE");
#endif
  a.n_tiles = d.N / bn;
  a.N = d.N;
  a.bias = d.bias;
  a.rowvec = d.rowvec;
  a.ld_rowvec = d.ld_rowvec;
  a.rows_per_image = d.rows_per_image;
  a.residual = d.residual;
  a.ld_res = d.ld_res;
  a.out = d.out;
  a.ldo = d.ldo;
  a.geglu = d.geglu;
  a.act = d.act;
  a.out_scale = d.out_scale;
  a.stats = d.stats;
  a.stats_rows = d.stats_rows > 0 ? d.stats_rows : 1;
  D4D_REQUIRE(d.stats == nullptr || (!d.geglu && d.kv_world == 0), "GroupNorm statistics: plain / conv epilogue only");
  D4D_REQUIRE(d.stats == nullptr || d.conv || (d.stats_rows > 0 && d.stats_rows % 32 == 0), "statistics need rows-per-image % 32 == 0");
  a.kv_world = d.kv_world;
  a.kv_col0 = d.kv_col0;
  a.kv_ld = d.kv_ld;
  a.kv_rows_local = d.kv_rows_local > 0 ? d.kv_rows_local : 1;
  a.kv_rows_global = d.kv_rows_global;
  a.kv_row_offset = d.kv_row_offset;
  for (int i = 0; i < 8; ++i) a.kv_dst[i] = d.kv_dst[i];
  D4D_REQUIRE(d.kv_world >= 0 && d.kv_world <= 8 && (d.kv_world == 0 || (d.kv_col0 % 16 == 0 && d.kv_ld % 8 == 0 && !d.geglu)), "K/V scatter arguments");
  D4D_REQUIRE(d.ldo % 8 == 0 && (d.residual == nullptr || d.ld_res % 8 == 0) &&
              (d.rowvec == nullptr || d.ld_rowvec % 8 == 0), "leading dimensions must be multiples of 8");

  if (!d.conv) {
    D4D_REQUIRE(d.K1 > 0 && d.K1 % 8 == 0 && d.K2 % 8 == 0, "K must be a multiple of 8");
    D4D_REQUIRE(d.A2 == nullptr || d.K1 % BLOCK_K == 0, "two-source GEMM needs K1 % 64 == 0");
    a.mode = 0;
    a.M = d.M;
    a.m_tiles = (d.M + BLOCK_M - 1) / BLOCK_M;
    a.kb_split = (d.K1 + BLOCK_K - 1) / BLOCK_K;
    const int kb2 = d.A2 ? (d.K2 + BLOCK_K - 1) / BLOCK_K : 0;
    a.k_blocks = a.kb_split + kb2;
    const int K = d.K1 + (d.A2 ? d.K2 : 0);
    if (int rc = make_tmap_2d(&L->tmap_a, d.A, d.M, d.K1, d.lda, BLOCK_K, BLOCK_M, 128)) return rc;
    if (d.A2) {
      if (int rc = make_tmap_2d(&L->tmap_a2, d.A2, d.M, d.K2, d.lda2, BLOCK_K, BLOCK_M, 128)) return rc;
    } else {
      L->tmap_a2 = L->tmap_a;
    }
    if (int rc = make_tmap_2d(&L->tmap_b, d.Wt, d.N, K, K, BLOCK_K, bn, 128)) return rc;
  } else {
    D4D_REQUIRE(d.Cin % 8 == 0, "conv Cin must be a multiple of 8");
    D4D_REQUIRE(d.conv_kind >= 0 && d.conv_kind <= 3, "conv_kind");
    D4D_REQUIRE(d.conv_kind != 1 || (d.H % 2 == 0 && d.W % 2 == 0), "stride-2 conv needs even H, W");
    a.mode = 1;
    a.n_img = d.n_img; a.Cin = d.Cin;
    // tap table, grid of output positions, output pixel mapping (GemmKernelArgs)
    a.in_stride = 1; a.out_sy = a.out_sx = 1; a.out_oy = a.out_ox = 0;
    a.n_phases = 1;
    if (d.conv_kind >= 2) {  // sub-pixel phase (a, b): rows {-1, 0} for a = 0, {0, +1} for a = 1 (same for columns)
      const int pa = d.conv_kind == 3 ? 0 : d.up_a, pb = d.conv_kind == 3 ? 0 : d.up_b;  // kind 3: the tile adds its phase
      a.n_taps = 4;
      for (int t = 0; t < 4; ++t) {
        a.tap_dy[t] = static_cast<signed char>((t >> 1) + pa - 1);
        a.tap_dx[t] = static_cast<signed char>((t & 1) + pb - 1);
      }
      a.H = d.H; a.W = d.W; a.out_H = 2 * d.H; a.out_W = 2 * d.W;
      a.out_sy = a.out_sx = 2; a.out_oy = pa; a.out_ox = pb;
      if (d.conv_kind == 3) a.n_phases = 4;
    } else {
      a.n_taps = 9;
      for (int t = 0; t < 9; ++t) {
        a.tap_dy[t] = static_cast<signed char>(t / 3 - 1);
        a.tap_dx[t] = static_cast<signed char>(t % 3 - 1);
      }
      if (d.conv_kind == 1) { a.in_stride = 2; a.H = d.H / 2; a.W = d.W / 2; }
      else { a.H = d.H; a.W = d.W; }
      a.out_H = a.H; a.out_W = a.W;
    }
    a.M = d.n_img * a.H * a.W;
    a.cin_blocks = (d.Cin + BLOCK_K - 1) / BLOCK_K;
    a.k_blocks = a.n_taps * a.cin_blocks;
    a.kb_split = a.k_blocks;
    // spatial tile: BW x BH x BN = 128 output positions
    int bw = 16; while (bw > a.W) bw >>= 1;
    int bh = 128 / bw; while (bh > a.H && bh > 1) bh >>= 1;
    // H, W need not be powers of two: the tile may overhang, TMA zero-fills and the epilogue masks
    int bnimg = 128 / (bw * bh);
    D4D_REQUIRE(bw * bh * bnimg == 128 && bnimg <= 256, "conv tile shape");
    D4D_REQUIRE(d.stats == nullptr || (bw * bh) % 32 == 0, "statistics need 32-row warps inside one image");
    a.BW = bw; a.BH = bh; a.BN = bnimg;
    a.tiles_x = (a.W + bw - 1) / bw;
    a.tiles_y = (a.H + bh - 1) / bh;
    const int tiles_n = (d.n_img + bnimg - 1) / bnimg;
    a.tiles_per_phase = a.tiles_x * a.tiles_y * tiles_n;
    a.m_tiles = a.tiles_per_phase * a.n_phases;
    a.M *= a.n_phases;  // output positions of the launch (FLOP count)
    if (int rc = make_tmap_nhwc(&L->tmap_a, d.A, d.n_img, d.H, d.W, d.Cin, BLOCK_K, bw, bh, bnimg, 128, a.in_stride)) return rc;
    L->tmap_a2 = L->tmap_a;
    const uint64_t kw = static_cast<uint64_t>(a.n_taps) * d.Cin;
    if (int rc = make_tmap_2d(&L->tmap_b, d.Wt, static_cast<uint64_t>(d.N) * a.n_phases, kw, kw, BLOCK_K, bn, 128)) return rc;
  }
  // TMA-store epilogue: plain GEMMs and unit-stride convs whose rows are 16-byte aligned (not the fused K/V scatter, not
  // the sub-pixel phase convs whose output pixels are strided)
  a.tma_store = 0;
  L->tmap_out = L->tmap_a;
  // ... and only short main loops (k_blocks <= 24), where the epilogue is what bounds the kernel: for long-K tiles the
  // row-segment stores hide under the main loop and the TMA path measured slightly slower
  if (d.kv_world == 0 && a.k_blocks <= 24 && (reinterpret_cast<uintptr_t>(d.out) & 15) == 0 && d.ldo % 8 == 0) {
    if (!d.conv) {
      const int out_cols = d.geglu ? d.N / 2 : d.N;
      // GEGLU: every unit must be a full 32 columns (the fallback of an odd last chunk is the row-segment path, which the
      // GEGLU branch does not carry)
      if (d.M >= 32 && out_cols >= 32 && (!d.geglu || (bn / 2) % 32 == 0)) {  // (boxes never exceed the tensor)
        if (int rc = make_tmap_2d(&L->tmap_out, d.out, d.M, out_cols, d.ldo, 32, 32, 64)) return rc;
        a.tma_store = 1;
      }
    } else if (d.conv_kind <= 1 && d.ldo == d.N && d.N >= 32) {
      const int bhq = a.BH < 32 / a.BW ? a.BH : 32 / a.BW;
      const int bnq = 32 / (a.BW * bhq);
      if (d.n_img >= bnq && a.out_H >= bhq && a.out_W >= a.BW) {
        if (int rc = make_tmap_nhwc_store(&L->tmap_out, d.out, d.n_img, a.out_H, a.out_W, d.N, a.BW, bhq, bnq)) return rc;
        a.tma_store = 1;
      }
    }
  }
  a.stg_bufs = stg_bufs_for(bn, a.k_blocks);
  int dev = 0, sms = 0;
  D4D_CUDA_OK(cudaGetDevice(&dev));
  D4D_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int total = a.m_tiles * a.n_tiles;
  L->grid = total < sms ? total : sms;
  return 0;
}

namespace {

using GemmKernelFn = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmKernelArgs);
struct GemmVariant {
  bool geglu, conv, tma;
  int epi;
  GemmKernelFn fn;
};
#define D4D_GV(G, C, T, E) {G, C, T, E, gemm_umma_kernel<G, C, T, E>}
// the feature sets the UNet plan launches (csrc/unet.cu), plus one E_ALL kernel per (conv, tma) pair for everything else
const GemmVariant kGemmVariants[] = {
    // plain GEMM, TMA-store epilogue (k_blocks <= 24): qkv | proj_in, shortcut | attn out, ff2 | proj_out, conv_in
    D4D_GV(false, false, true, 0), D4D_GV(false, false, true, E_BIAS), D4D_GV(false, false, true, E_BIAS | E_RES),
    D4D_GV(false, false, true, E_BIAS | E_RES | E_STATS), D4D_GV(false, false, true, E_ALL),
    // plain GEMM, row-segment stores (long K): ff2
    D4D_GV(false, false, false, E_BIAS | E_RES), D4D_GV(false, false, false, E_ALL),
    // 3x3 convs (long K: row-segment stores): resnet conv1 | conv2 | down / up sampling
    D4D_GV(false, true, false, E_BIAS | E_ROWVEC | E_STATS), D4D_GV(false, true, false, E_BIAS | E_RES | E_STATS),
    D4D_GV(false, true, false, E_BIAS | E_STATS), D4D_GV(false, true, false, E_ALL), D4D_GV(false, true, true, E_ALL),
    // GEGLU projection (only the bias bit matters)
    D4D_GV(true, false, true, E_BIAS), D4D_GV(true, false, false, E_BIAS),
};
#undef D4D_GV
constexpr int kNumGemmVariants = sizeof(kGemmVariants) / sizeof(kGemmVariants[0]);

int gemm_variant_of(const GemmKernelArgs& a) {
  const bool conv = a.mode != 0, tma = a.tma_store != 0, geglu = a.geglu != 0;
  int need = 0;
  if (a.bias) need |= E_BIAS;
  if (!geglu) {
    if (a.rowvec) need |= E_ROWVEC;
    if (a.act == 1 || a.out_scale != 1.0f) need |= E_ACT;
    if (a.residual) need |= E_RES;
    if (a.stats) need |= E_STATS;
    if (a.kv_world > 0) need |= E_KV;
  }
  int generic = -1;
  for (int i = 0; i < kNumGemmVariants; ++i) {
    const GemmVariant& v = kGemmVariants[i];
    if (v.geglu != geglu || v.conv != conv || v.tma != tma) continue;
    if (v.epi == need) return i;
    if ((v.epi & need) == need && (generic < 0 || v.epi == E_ALL)) generic = i;  // geglu: E_BIAS covers {} as well
  }
  return generic;
}

}  // namespace

int gemm_run(const GemmLaunch& L, cudaStream_t stream) {
  static PerDeviceOnce attr_once[kNumGemmVariants];
  const int vi = gemm_variant_of(L.args);
  D4D_REQUIRE(vi >= 0, "no GEMM kernel instantiation covers this launch");
  const GemmVariant& v = kGemmVariants[vi];
  if (int rc = ensure_dyn_smem(v.fn, SMEM_BYTES, attr_once[vi])) return rc;
  D4D_CUDA_OK(launch_pdl(v.fn, dim3(L.grid), dim3(NUM_THREADS), SMEM_BYTES, stream, L.tmap_a, L.tmap_a2, L.tmap_b, L.tmap_out, L.args));
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

double gemm_flops(const GemmLaunch& L) {
  return 2.0 * static_cast<double>(L.args.M) * L.args.N * (static_cast<double>(L.args.k_blocks) * BLOCK_K);
}

}  // namespace d4d
