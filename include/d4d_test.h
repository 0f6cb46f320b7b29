/* d4d_test.h -- extra entry points of libd4d_test.so (NOT part of the product library libd4d.so).
 *
 * libd4d_test.so is the same sources built with -DD4D_TEST_KERNELS -DD4D_ABLATE (diffuman4d_b200/build.py): it exports
 * everything include/d4d.h declares plus the measurement / probe kernels below, and honours the ablation switches
 * D4D_GEMM_ABLATE / D4D_ATTN_ABLATE used by tools/ablate_*.py.  tests/ use it only to pin the UMMA operand encodings. */
#ifndef D4D_TEST_H_
#define D4D_TEST_H_
#include "d4d.h"
#ifdef __cplusplus
extern "C" {
#endif
/* UMMA operand-encoding probe (tests pin the shared-memory descriptor conventions against torch.matmul) */
int d4d_op_probe_umma(const void* A, const void* B, float* D, int N, int K, int a_src, int b_major, uint32_t b_lbo,
                      uint32_t b_sbo, uint32_t b_kadv, void* stream);
/* Device microbenchmarks that size the attention / GEMM kernels (tools/microbench.py); `warps` per CTA, `blocks` CTAs;
 * cycles_dev[blocks]. */
int d4d_microbench(int kind, int warps, int iters, int blocks, uint64_t* cycles_dev, float* sink_dev, void* stream);
#ifdef __cplusplus
}
#endif
#endif /* D4D_TEST_H_ */
