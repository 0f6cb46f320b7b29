// Device microbenchmarks used to size the attention kernel (exported as d4d_microbench, test/bring-up utility):
// per-SM throughput of tcgen05.ld / tcgen05.st (TMEM <-> registers), MUFU.EX2 and the bf16 pack conversion.
#include "kernels.h"

namespace d4d {
namespace {

// kind 0: tcgen05.ld 32x32b.x32   1: tcgen05.ld 32x32b.x16   2: tcgen05.st 32x32b.x16   3: ex2.approx   4: cvt.bf16x2
__global__ void microbench_kernel(int kind, int iters, unsigned long long* cycles, float* sink) {
  __shared__ uint32_t slot;
  __shared__ uint64_t bars[9];
  __shared__ __align__(1024) uint8_t tile[2][128 * 128];
  uint64_t& bar = bars[0];
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot;
  const uint32_t taddr = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + ((warp >> 2) & 3) * 64;
  float acc = 0.f;
  __syncthreads();
  const unsigned long long t0 = clock64();
  if (kind == 0) {
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32];
      tmem_ld32(taddr + (i & 1) * 32, v);
      tmem_ld_wait();
      acc += __uint_as_float(v[0] ^ v[13] ^ v[31]);
    }
  } else if (kind == 1) {
    for (int i = 0; i < iters; ++i) {
      uint32_t v[16];
      tmem_ld16(taddr + (i & 3) * 16, v);
      tmem_ld_wait();
      acc += __uint_as_float(v[0] ^ v[7] ^ v[15]);
    }
  } else if (kind == 5) {
    for (int i = 0; i < iters; ++i) {
      uint32_t v[16], w[16];
      tmem_ld16(taddr, v);
      tmem_ld16(taddr + 16, w);
      tmem_ld_wait();
      acc += __uint_as_float(v[0] ^ v[15] ^ w[0] ^ w[15]);
    }
  } else if (kind == 6) {
    for (int i = 0; i < iters; ++i) {
      uint32_t v[32], w[32];
      tmem_ld32(taddr, v);
      tmem_ld32(taddr + 32, w);
      tmem_ld_wait();
      acc += __uint_as_float(v[0] ^ v[31] ^ w[0] ^ w[31]);
    }
  } else if (kind == 7 || kind == 8) {
    // latency of tcgen05.commit -> mbarrier completion -> try_wait return (kind 8: preceded by one 128x64x16 MMA)
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      fence_mbar_init();
      uint32_t ph = 0;
      const uint32_t idesc = make_idesc_bf16(128, 64, 0, 0);
      for (int i = 0; i < iters; ++i) {
        if (kind == 8) {
          const uint64_t ad = make_smem_desc(smem_u32(tile[0]), 0, 1024, 2);
          const uint64_t bd = make_smem_desc(smem_u32(tile[1]), 0, 1024, 2);
          umma_ss(tmem, ad, bd, idesc, 0);
        }
        umma_commit(&bar);
        mbar_wait(&bar, ph);
        ph ^= 1;
      }
    }
  } else if (kind >= 9 && kind <= 16) {
    // tensor-pipe throughput of back-to-back 128xNx16 MMAs with / without interleaved tcgen05.commit (nobody waits on the
    // barriers except for the last one):
    //   9: N=256, no commits   10: N=256, commit every 4 MMAs   11: N=64, no commits   12: N=64, commit every 4 MMAs
    //  13: commits only        14: N=256, commit every 16 MMAs  15: N=160, no commits  16: N=160, commit every 4 MMAs
    if (threadIdx.x == 0) {
      for (int i = 0; i < 9; ++i) mbar_init(&bars[i], 1);
      fence_mbar_init();
      const int N = (kind == 11 || kind == 12) ? 64 : (kind >= 15 ? 160 : 256);
      const int every = (kind == 10 || kind == 12 || kind == 16) ? 4 : (kind == 14 ? 16 : (kind == 13 ? 1 : 0));
      const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      const uint64_t ad = make_smem_desc(smem_u32(tile[0]), 0, 1024, 2);
      const uint64_t bd = make_smem_desc(smem_u32(tile[0]), 0, 1024, 2);
      int b = 0;
      for (int i = 0; i < iters; ++i) {
        if (kind != 13) umma_ss(tmem + (i & 1) * 256, ad + 2 * (i & 3), bd + 2 * (i & 3), idesc, 1);
        if (every && (i % every) == every - 1) {
          umma_commit(&bars[b]);
          b = (b + 1) & 7;
        }
      }
      umma_commit(&bars[8]);
      mbar_wait(&bars[8], 0);
    }
  } else if (kind == 17 || kind == 18) {
    // issue cost of tcgen05.fence::after_thread_sync (17) / of a try_wait on an already completed barrier phase (18)
    if (threadIdx.x == 0) {
      mbar_init(&bars[0], 1);
      fence_mbar_init();
      mbar_arrive(&bars[0]);  // phase 0 complete
      for (int i = 0; i < iters; ++i) {
        if (kind == 17) tc_fence_after();
        else mbar_wait(&bars[0], 0);
      }
    }
  } else if (kind >= 19 && kind <= 24) {
    // 19: N=64 MMAs issued from two threads of different warps (is the 114-cycle floor the pipe or the issuing thread?)
    // 20..24: one thread, no commits, N = 128 / 192 / 224 / 32 / 240
    const int N = kind == 19 ? 64 : (kind == 20 ? 128 : (kind == 21 ? 192 : (kind == 22 ? 224 : (kind == 23 ? 32 : 240))));
    const int issuers = kind == 19 ? 2 : 1;
    if (threadIdx.x == 0) {
      for (int i = 0; i < 9; ++i) mbar_init(&bars[i], 1);
      fence_mbar_init();
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && warp < issuers) {
      const uint32_t idesc = make_idesc_bf16(128, N, 0, 0);
      const uint64_t ad = make_smem_desc(smem_u32(tile[0]), 0, 1024, 2);
      for (int i = 0; i < iters / issuers; ++i) umma_ss(tmem + warp * 256, ad + 2 * (i & 3), ad + 2 * (i & 3), idesc, 1);
      umma_commit(&bars[warp]);
      mbar_wait(&bars[warp], 0);
    }
  } else if (kind == 2) {
    uint32_t v[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) v[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
      tmem_st16(taddr + (i & 3) * 16, v);
      v[0] += 1;
    }
    tmem_st_wait();
  } else if (kind == 3) {
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = -0.001f * (threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; ++k) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += x[k];
  } else {
    float x[8];
    uint32_t u = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = 0.5f * (threadIdx.x + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        uint32_t p;
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(x[k + 1]), "f"(x[k]));
        u ^= p;
        x[k] += 1.0f;
      }
    }
    acc += __uint_as_float(u);
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  if (acc == 12345.678f) sink[0] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// Issue-rate probes of the CUDA-core pipes the attention softmax leans on (kinds 100+): every thread runs 8 independent
// dependency chains of one instruction type; cycles / (8 * iters * warps per sub-partition) = cycles per warp instruction.
template <int kind>
__global__ void pipe_bench_kernel(int iters, unsigned long long* cycles, float* sink, unsigned one) {
  float x[8];
  uint64_t y[8];
  uint32_t u[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    x[k] = -0.001f * (threadIdx.x + k) - 0.5f;
    y[k] = f2_pack(0.5f + 0.001f * k, 0.25f + 0.001f * threadIdx.x);
    u[k] = threadIdx.x * 7 + k;
  }
  const uint64_t ca = f2_pack(0.999f, 1.001f), cb = f2_pack(0.0001f, -0.0001f);
  __syncthreads();
  const unsigned long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      switch (kind) {
        case 100: asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k])); break;
        case 101: asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[k]) : "f"(x[k]), "f"(x[(k + 1) & 7])); break;
        case 102: asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[k]) : "l"(ca), "l"(cb)); break;
        case 103: asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(y[k]) : "l"(cb)); break;
        case 104: asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(x[k]) : "f"(x[(k + 1) & 7]), "f"(x[(k + 2) & 7])); break;
        case 105: asm volatile("mad.lo.u32 %0, %0, %1, 0x8000;" : "+r"(u[k]) : "r"(one)); break;
        case 106: asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u[k]) : "r"(u[(k + 1) & 7])); break;
        case 107: asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[k]) : "f"(0.999f), "f"(0.0001f)); break;
        case 108:
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[k]) : "l"(ca), "l"(cb));
          break;
        case 109: asm volatile("max.f32 %0, %0, %1;" : "+f"(x[k]) : "f"(x[(k + 1) & 7])); break;
        case 110: asm volatile("add.rm.f32x2 %0, %0, %1;" : "+l"(y[k]) : "l"(cb)); break;
        case 111:  // the v0 pair: FFMA2, 2 MUFU, FADD2, F2FP
          asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(y[k]) : "l"(ca), "l"(cb));
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[k]));
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[(k + 4) & 7]));
          asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(y[(k + 4) & 7]) : "l"(cb));
          asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u[k]) : "f"(x[k]), "f"(x[(k + 1) & 7]));
          break;
        case 112: asm volatile("add.u32 %0, %0, 0x8000;" : "+r"(u[k])); break;
        default: break;
      }
    }
  }
  __syncthreads();
  const unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  float acc = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    float a, b;
    f2_unpack(y[k], a, b);
    acc += x[k] + a + b + __uint_as_float(u[k]);
  }
  if (acc == 12345.678f) sink[0] = acc;
}

}  // namespace

int microbench_run(int kind, int warps, int iters, int blocks, unsigned long long* cycles_dev, float* sink_dev, cudaStream_t s) {
  if (kind >= 100) {
    D4D_REQUIRE(kind <= 112 && warps >= 1 && warps <= 32 && iters > 0 && blocks > 0, "microbench arguments");
#define D4D_PB(K) case K: pipe_bench_kernel<K><<<blocks, warps * 32, 0, s>>>(iters, cycles_dev, sink_dev, 1u); break;
    switch (kind) {
      D4D_PB(100) D4D_PB(101) D4D_PB(102) D4D_PB(103) D4D_PB(104) D4D_PB(105) D4D_PB(106) D4D_PB(107) D4D_PB(108) D4D_PB(109)
      D4D_PB(110) D4D_PB(111) D4D_PB(112)
    }
#undef D4D_PB
    D4D_CUDA_OK(cudaGetLastError());
    return 0;
  }
  D4D_REQUIRE(kind >= 0 && kind <= 24 && warps >= 1 && warps <= 16 && iters > 0 && blocks > 0, "microbench arguments");
  microbench_kernel<<<blocks, warps * 32, 0, s>>>(kind, iters, cycles_dev, sink_dev);
  D4D_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace d4d
