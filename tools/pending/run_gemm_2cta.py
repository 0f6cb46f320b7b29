"""Round-2 bring-up of the CTA-pair GEMM draft (tools/pending/gemm_2cta.cu): build it against libd4d.so, check it against
torch, time it next to the production single-CTA kernel.  Run on a B200 under a timeout (a protocol slip can hang a CTA pair):
    timeout 120 python tools/pending/run_gemm_2cta.py
"""
import ctypes as C
import os
import subprocess
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from diffuman4d_b200 import ops  # noqa: E402  (loads libd4d.so)
from diffuman4d_b200._lib import lib  # noqa: E402

src = os.path.join(ROOT, "tools", "pending", "gemm_2cta.cu")
so = os.path.join(ROOT, "tools", "pending", "libgemm2cta.so")
pkg = os.path.join(ROOT, "diffuman4d_b200")
if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
                    "-shared", "-o", so, src, f"-L{pkg}", "-ld4d", f"-Xlinker=-rpath={pkg}"], check=True)
g2 = C.CDLL(so)
g2.gemm_2cta_run.restype = C.c_int
lib().d4d_last_error.restype = C.c_char_p


def gemm2(a, w, bias=None):
    M, K = a.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16)
    rc = g2.gemm_2cta_run(C.c_void_p(a.data_ptr()), a.stride(0), C.c_void_p(w.data_ptr()), w.stride(0), M, N, K,
                          C.c_void_p(bias.data_ptr() if bias is not None else 0), C.c_void_p(out.data_ptr()), out.stride(0),
                          C.c_void_p(torch.cuda.current_stream().cuda_stream))
    if rc:
        raise RuntimeError(lib().d4d_last_error().decode())
    return out


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for M, N, K in [(256, 256, 64), (512, 512, 256), (1000, 768, 320), (8192, 1280, 5120), (32768, 5120, 640), (131072, 2560, 320)]:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    out = gemm2(a, w, bias)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    err = (out.float() - ref).abs().max().item() / ref.abs().max().item()
    t2 = timeit(lambda: gemm2(a, w, bias))
    t1 = timeit(lambda: ops.gemm(a, w, bias, block_n=256))
    fl = 2.0 * M * N * K
    print(f"M={M:6d} N={N:5d} K={K:5d}  rel err {err:.2e}  2-CTA {t2:8.1f} us ({fl / t2 / 1e6:6.0f} TF/s)   1-CTA bn256 {t1:8.1f} us "
          f"({fl / t1 / 1e6:6.0f} TF/s)", flush=True)
    assert err < 1e-2
