"""GEMM / conv main-loop ablations (D4D_GEMM_ABLATE bits: 1 no TMA loads, 2 no MMAs, 4 no A loads, 8 no B loads).
Results are garbage by construction; only the device time matters.  One process per setting (the switch is read once).
    for a in 0 1 2 3 4 8; do D4D_GEMM_ABLATE=$a python tools/ablate_gemm.py; done
"""
import os
import sys

os.environ["D4D_USE_TEST_LIB"] = "1"   # the ablation switches exist only in the tools build (libd4d_test.so)

import torch

sys.path.insert(0, ".")
from diffuman4d_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")


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


def conv(hw, cin, cout, bn):
    x = torch.randn(32, hw, hw, cin, device=dev).bfloat16()
    w = (torch.randn(cout, 9, cin, device=dev) * 0.02).bfloat16()
    b = torch.randn(cout, device=dev)
    t = timeit(lambda: ops.conv3x3(x, w, b, block_n=bn))
    tiles = (32 * hw * hw // 128) * (cout // bn)
    waves = -(-tiles // 148)
    kb = 9 * cin // 64
    return t, t * 1e-6 / waves / kb * 1.965e9


def gemm(M, N, K, bn, geglu=False, residual=False):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    b = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).bfloat16() if residual else None
    t = timeit(lambda: ops.gemm(a, w, b, block_n=bn, geglu=geglu, residual=res))
    tiles = (M // 128) * (N // bn)
    waves = -(-tiles // 148)
    kb = K // 64
    return t, t * 1e-6 / waves / kb * 1.965e9


abl = os.environ.get("D4D_GEMM_ABLATE", "0")
out = [f"ablate={abl:>2s}"]
small = [("gemm L0 qkv K320 bn240", lambda: gemm(131072, 960, 320, 240)), ("gemm L0 proj+res K320 bn160", lambda: gemm(131072, 320, 320, 160, residual=True)),
         ("gemm L0 ff1 geglu K320 bn256", lambda: gemm(131072, 2560, 320, 256, geglu=True)), ("gemm L0 ff2+res K1280 bn160", lambda: gemm(131072, 320, 1280, 160, residual=True)),
         ("gemm L1 qkv K640 bn240", lambda: gemm(32768, 1920, 640, 240)), ("gemm L1 ff1 geglu K640 bn256", lambda: gemm(32768, 5120, 640, 256, geglu=True))]
for name, fn in small if "--small" in sys.argv else [("conv L0 320 bn160", lambda: conv(64, 320, 320, 160)), ("conv L0 960 bn64", lambda: conv(64, 960, 320, 64)),
                 ("conv L0 960 bn160", lambda: conv(64, 960, 320, 160)), ("conv L2 2560 bn256", lambda: conv(16, 2560, 1280, 256)),
                 ("gemm L1 ff2 bn160", lambda: gemm(32768, 640, 2560, 160)), ("gemm L1 ff2 bn64", lambda: gemm(32768, 640, 2560, 64)),
                 ("gemm L2 ff2 bn256", lambda: gemm(8192, 1280, 5120, 256))]:
    t, cyc = fn()
    out.append(f"{name}: {t:7.1f}us {cyc:5.0f}cyc/kb")
print(" | ".join(out), flush=True)
