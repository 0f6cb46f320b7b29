// UMMA operand-encoding probe (bring-up/test utility, exported as d4d_probe_umma).
//
// One CTA computes D[128, N] = A[128, K] * B^T with a single accumulator tile, where the operand
// sourcing is selected at run time:
//   a_src   0: A from shared memory (K-major, 128B swizzle, TMA-loaded)   1: A from TMEM (bf16 pairs
//              written with tcgen05.st, lane = row, column j = elements 2j, 2j+1)
//   b_major 0: B given as [N, K] row-major (K-major operand)   1: B given as [K, N] row-major (MN-major
//              operand -- the layout of V in attention)
// and the shared-memory descriptor fields of B (LBO, SBO, per-16-K start-address advance) are passed
// in, so the test-suite can pin the encoding the attention kernel relies on against a torch matmul.
#include "kernels.h"

namespace d4d {

namespace {

struct ProbeArgs {
  int N, K;
  int a_src, b_major;
  uint32_t b_lbo, b_sbo, b_kadv;
  const bf16* A;  // [128, K]
  float* D;       // [128, N]
};

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const ProbeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;              // up to 2 boxes of [128 rows][64] = 32 KB
  uint8_t* sB = smem + 32768;      // up to 32 KB
  uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + 65536);
  uint64_t* bar_mma = bar_load + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar_load + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_load, 1);
    mbar_init(bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t d_tmem = tmem;        // columns [0, N)
  const uint32_t a_tmem = tmem + 128;  // columns [128, 128 + K/2)

  if (threadIdx.x == 0) {
    const int a_boxes = p.K / 64;
    uint32_t bytes = 128 * p.K * 2;  // A
    if (p.b_major == 0) bytes += p.N * p.K * 2;
    else bytes += p.K * p.N * 2;
    mbar_expect_tx(bar_load, bytes);
    for (int i = 0; i < a_boxes; ++i) tma_load_2d(sA + i * 16384, &tmap_a, bar_load, i * 64, 0);
    if (p.b_major == 0) {
      // [N, K]: box {64 k, N rows} per 64-wide k block
      for (int i = 0; i < a_boxes; ++i) tma_load_2d(sB + i * (p.N * 128), &tmap_b, bar_load, i * 64, 0);
    } else {
      // [K, N]: box {64 n, K rows} per 64-wide n block
      for (int i = 0; i < p.N / 64; ++i) tma_load_2d(sB + i * (p.K * 128), &tmap_b, bar_load, i * 64, 0);
    }
  }
  if (p.a_src == 1) {
    // stage A into TMEM: thread t <-> row t
    const int row = threadIdx.x;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(p.A + static_cast<size_t>(row) * p.K);
    for (int c = 0; c < p.K / 2; c += 16) {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = src[c + i];
      tmem_st16(a_tmem + c + (static_cast<uint32_t>(warp * 32) << 16), v);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (threadIdx.x == 0) {
    mbar_wait(bar_load, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, p.N, 0, p.b_major);
    for (int k = 0; k < p.K / 16; ++k) {
      // A (smem, K-major SW128): 64-wide k blocks are 16 KB apart, +32 B per 16 elements inside a block
      const uint32_t a_addr = smem_u32(sA) + (k / 4) * 16384 + (k % 4) * 32;
      const uint64_t adesc = make_smem_desc(a_addr, 0, 1024, 2);
      uint64_t bdesc;
      if (p.b_major == 0) {
        const uint32_t b_addr = smem_u32(sB) + (k / 4) * (p.N * 128) + (k % 4) * 32;
        bdesc = make_smem_desc(b_addr, 0, 1024, 2);
      } else {
        bdesc = make_smem_desc(smem_u32(sB) + k * p.b_kadv, p.b_lbo, p.b_sbo, 2);
      }
      if (p.a_src == 0) umma_ss(d_tmem, adesc, bdesc, idesc, k > 0);
      else umma_ts(d_tmem, a_tmem + k * 8, bdesc, idesc, k > 0);
    }
    umma_commit(bar_mma);
  }
  __syncwarp();
  mbar_wait(bar_mma, 0);
  tc_fence_after();
  {
    const int row = threadIdx.x;
    for (int c = 0; c < p.N; c += 16) {
      uint32_t v[16];
      tmem_ld16(d_tmem + c + (static_cast<uint32_t>(warp * 32) << 16), v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) p.D[static_cast<size_t>(row) * p.N + c + i] = __uint_as_float(v[i]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 256);
  }
}

}  // namespace

int probe_umma_run(const bf16* A, const bf16* B, float* D, int N, int K, int a_src, int b_major, uint32_t b_lbo,
                   uint32_t b_sbo, uint32_t b_kadv, cudaStream_t stream) {
  D4D_REQUIRE((N == 64 || N == 128) && (K == 64 || K == 128), "probe supports N,K in {64,128}");
  CUtensorMap ta, tb;
  if (int rc = make_tmap_2d(&ta, A, 128, K, K, 64, 128, 128)) return rc;
  if (b_major == 0) {
    if (int rc = make_tmap_2d(&tb, B, N, K, K, 64, N, 128)) return rc;
  } else {
    if (int rc = make_tmap_2d(&tb, B, K, N, N, 64, K, 128)) return rc;
  }
  ProbeArgs p;
  p.N = N; p.K = K; p.a_src = a_src; p.b_major = b_major;
  p.b_lbo = b_lbo; p.b_sbo = b_sbo; p.b_kadv = b_kadv;
  p.A = A; p.D = D;
  static PerDeviceOnce attr_once;
  const int smem = 65536 + 1024 + 64;
  if (int rc = ensure_dyn_smem(probe_kernel, smem, attr_once)) return rc;
  probe_kernel<<<1, 128, smem, stream>>>(ta, tb, p);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace d4d
