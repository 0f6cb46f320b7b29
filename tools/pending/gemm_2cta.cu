// DRAFT for round 2 -- NOT part of libd4d.so, NOT validated on hardware (written after the round-1 GPU budget was spent).
//
// out[M, N] = A[M, K] . W[N, K]^T (+bias) with CTA PAIRS: `tcgen05.mma.cta_group::2` (M = 256 per pair, N = 256, K = 16).
// Why: the single-CTA kernel (csrc/gemm_umma.cu) sits on the shared-memory bandwidth of one SM -- every operand byte crosses
// shared memory twice (TMA write + MMA read): 96 KB per 64-wide k-block at block_n 256 = 768 cycles against 512 of MMA time
// (profiles/r01d_ablate_gemm.txt).  In a pair each CTA stages its 128 rows of A and only HALF of B
// (128 of the 256 N rows); the instruction reads the other half from the peer's shared memory, so a CTA moves 32 + 32 KB per
// k-block = 512 cycles: the main loop becomes MMA-bound.
//
// Protocol (PTX forms and barrier ownership as in the CUTLASS sm100 headers vendored in this image:
// cute/arch/copy_sm100_tma.hpp SM100_TMA_2SM_LOAD_2D, cutlass/arch/barrier.h umma_arrive_multicast_2x1SM,
// cute/arch/tmem_allocator_sm100.hpp Allocator2Sm, cutlass/pipeline/sm100_pipeline.hpp PipelineTmaUmmaAsync):
//   * cluster (2,1,1); rank 0 = leader.  Both CTAs run a TMA producer warp and 12 epilogue warps; only the leader's
//     warp 1 issues MMAs.
//   * full[stage] lives in the LEADER only: its producer arms it with the bytes of BOTH CTAs; the peer's loads use the
//     `.cta_group::2` TMA form with the barrier address masked to the even CTA (0xFEFFFFFF).
//   * empty[stage] / tfull[acc] exist in both CTAs at the same offsets; the leader's commits are multicast (mask 0b11).
//   * tempty[acc] lives in the leader; the peer's epilogue warps arrive remotely (mapa + mbarrier.arrive.shared::cluster).
//   * TMEM: both CTAs allocate with cta_group::2 from the warp with the same index, 2 x 256 columns (double buffer).
// Build + run next round:  tools/pending/run_gemm_2cta.py
#include "../../diffuman4d_b200/csrc/kernels.h"

namespace d4d {
namespace {

constexpr int BM = 128;        // rows of A per CTA (256 per pair)
constexpr int BN_HALF = 128;   // rows of W per CTA (N tile 256 per pair)
constexpr int BK = 64;
constexpr int A_BYTES2 = BM * BK * 2;        // 16 KB
constexpr int B_BYTES2 = BN_HALF * BK * 2;   // 16 KB
constexpr int STAGE2 = A_BYTES2 + B_BYTES2;  // 32 KB per CTA per stage
constexpr int STAGES2 = 6;
constexpr int EPI_Q = 3;                     // epilogue warps per TMEM lane quarter
constexpr int THREADS2 = 64 + 128 * EPI_Q;
constexpr int BAR2_BYTES = 1024;
constexpr int SMEM2 = 1024 + BAR2_BYTES + STAGES2 * STAGE2;
constexpr uint32_t PEER_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address -> even CTA of the pair

struct Gemm2Args {
  int M, N, K;
  int m_tiles;   // ceil(M / 256)
  int n_tiles;   // N / 256
  int k_blocks;  // ceil(K / 64)
  const float* bias;
  bf16* out;
  int ldo;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* t, uint64_t* leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(t)), "r"(smem_u32(leader_bar) & PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {  // arrive on the same barrier of CTA `cta`
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS2, 1)
gemm_2cta_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const Gemm2Args a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
  uint64_t* full = bars;                  // [STAGES2]  used in the leader only
  uint64_t* empty = bars + STAGES2;       // [STAGES2]  both CTAs (multicast commit)
  uint64_t* tfull = bars + 2 * STAGES2;   // [2]        both CTAs (multicast commit)
  uint64_t* tempty = tfull + 2;           // [2]        leader only; 2 x 4 x EPI_Q arrivals (both CTAs' epilogue warps)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* ring = smem + BAR2_BYTES;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  const int total_tiles = a.m_tiles * a.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < STAGES2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * 4 * EPI_Q);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2cta(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer's barriers / TMEM must exist before any remote arrive, multicast commit or 2-CTA MMA
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs): own 128 rows of A, own 128 rows of W =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs) {
      const int m_tile = tile / a.n_tiles;
      const int n_tile = tile % a.n_tiles;
      const int m0 = m_tile * 2 * BM + static_cast<int>(rank) * BM;
      const int n0 = n_tile * 2 * BN_HALF + static_cast<int>(rank) * BN_HALF;
      for (int kb = 0; kb < a.k_blocks; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1);
        uint8_t* sa = ring + stage * STAGE2;
        if (elect_one()) {
          if (leader) mbar_expect_tx(&full[stage], 2 * STAGE2);  // bytes of BOTH CTAs land on the leader's barrier
          tma_load_2d_2sm(sa, &tmap_a, &full[stage], kb * BK, m0);
          tma_load_2d_2sm(sa + A_BYTES2, &tmap_b, &full[stage], kb * BK, n0);
        }
        __syncwarp();
        if (++stage == STAGES2) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (leader) {
      const uint32_t idesc = make_idesc_bf16(2 * BM, 2 * BN_HALF, 0, 0);
      const uint32_t ring_u32 = smem_u32(ring);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * 256;
        for (int kb = 0; kb < a.k_blocks; ++kb) {
          mbar_wait(&full[stage], phase);
          const uint32_t sa = ring_u32 + stage * STAGE2;
          const uint64_t adesc = make_smem_desc(sa, 0, 1024, 2);
          const uint64_t bdesc = make_smem_desc(sa + A_BYTES2, 0, 1024, 2);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_ss_2cta(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2cta(&empty[stage]);
          }
          __syncwarp();
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit_2cta(&tfull[acc]);
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue (both CTAs): own 128 accumulator rows x 256 columns =====================
    const int q = warp & 3;
    const int chunk0 = (warp - 2) >> 2;
    const int r = q * 32 + lane;
    int it = 0;
    for (int tile = pair; tile < total_tiles; tile += n_pairs, ++it) {
      const int acc = it & 1;
      const int m_tile = tile / a.n_tiles;
      const int n_tile = tile % a.n_tiles;
      const long long row = static_cast<long long>(m_tile) * 2 * BM + rank * BM + r;
      const int n0 = n_tile * 2 * BN_HALF;
      mbar_wait(&tfull[acc], (it >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * 256 + (static_cast<uint32_t>(q * 32) << 16);
      for (int c = chunk0; c < 16; c += EPI_Q) {
        uint32_t v[16];
        tmem_ld16(taddr + c * 16, v);
        tmem_ld_wait();
        if (row < a.M) {
          float f[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) f[i] = __uint_as_float(v[i]) + (a.bias ? __ldg(a.bias + n0 + c * 16 + i) : 0.f);
          uint4 o0, o1;
          o0.x = pack_bf16x2(f[0], f[1]);   o0.y = pack_bf16x2(f[2], f[3]);
          o0.z = pack_bf16x2(f[4], f[5]);   o0.w = pack_bf16x2(f[6], f[7]);
          o1.x = pack_bf16x2(f[8], f[9]);   o1.y = pack_bf16x2(f[10], f[11]);
          o1.z = pack_bf16x2(f[12], f[13]); o1.w = pack_bf16x2(f[14], f[15]);
          uint4* op = reinterpret_cast<uint4*>(a.out + static_cast<size_t>(row) * a.ldo + n0 + c * 16);
          op[0] = o0;
          op[1] = o1;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&tempty[acc], 0);  // the accumulator buffer of THIS CTA is drained -> tell the leader
    }
  }

  // the leader's MMAs read the peer's shared memory and both CTAs' TMEM: nobody leaves before everything has drained
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, 512);
  }
}

}  // namespace
}  // namespace d4d

// out = A . W^T (+bias);  M any, N % 256 == 0, K % 8 == 0;  row-major bf16, fp32 bias.  Returns 0 / d4d status code.
extern "C" int gemm_2cta_run(const void* A, int lda, const void* W, int ldw, int M, int N, int K, const float* bias, void* out,
                             int ldo, void* stream) {
  using namespace d4d;
  D4D_REQUIRE(N % 256 == 0 && K % 8 == 0 && M > 0, "gemm_2cta: N must be a multiple of 256, K of 8");
  CUtensorMap ta, tb;
  if (int rc = make_tmap_2d(&ta, A, M, K, lda, BK, BM, 128)) return rc;
  if (int rc = make_tmap_2d(&tb, W, N, K, ldw, BK, BN_HALF, 128)) return rc;
  Gemm2Args a;
  a.M = M; a.N = N; a.K = K;
  a.m_tiles = (M + 2 * BM - 1) / (2 * BM);
  a.n_tiles = N / 256;
  a.k_blocks = (K + BK - 1) / BK;
  a.bias = bias;
  a.out = static_cast<bf16*>(out);
  a.ldo = ldo;
  int dev = 0, sms = 148;
  D4D_CUDA_OK(cudaGetDevice(&dev));
  D4D_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  D4D_CUDA_OK(cudaFuncSetAttribute(gemm_2cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2));
  int pairs = sms / 2;
  if (pairs > a.m_tiles * a.n_tiles) pairs = a.m_tiles * a.n_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs);
  cfg.blockDim = dim3(THREADS2);
  cfg.dynamicSmemBytes = SMEM2;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  D4D_CUDA_OK(cudaLaunchKernelEx(&cfg, gemm_2cta_kernel, ta, tb, a));
  return 0;
}
