"""Sweep the GEMM / conv tile width (block_n) over the shapes of the W16 @ 64x64 window step and print the device time of
every candidate next to the one the built-in heuristic (block_n = 0) picks.  Run on the GPU box:
    python tools/tune_block_n.py > gpurun_out/tune_block_n.txt
"""
import sys

import torch

sys.path.insert(0, ".")
from diffuman4d_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B = 32
LEVELS = [(320, 64), (640, 32), (1280, 16), (1280, 8)]


def timeit(fn, iters=12):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us


def cands(N, mult):
    return [c for c in range(mult, 257, mult) if N % c == 0 and c >= 64]


def sweep_gemm(name, M, N, K, geglu=False, residual=False):
    a = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
    bias = torch.randn(N, device=dev)
    res = torch.randn(M, N, device=dev).bfloat16() if residual else None
    flops = 2.0 * M * N * K
    t0 = timeit(lambda: ops.gemm(a, w, bias, geglu=geglu, residual=res))
    row = [f"{name:34s} M={M:6d} N={N:5d} K={K:5d} auto={t0:7.1f}us ({flops / t0 / 1e6:6.0f} TF/s) |"]
    for c in cands(N, 32 if geglu else 16):
        t = timeit(lambda: ops.gemm(a, w, bias, geglu=geglu, residual=res, block_n=c))
        row.append(f"{c}:{t:7.1f}")
    print(" ".join(row), flush=True)


def sweep_conv(name, hw, cin, cout):
    x = torch.randn(B, hw, hw, cin, device=dev).bfloat16()
    w = (torch.randn(cout, 9, cin, device=dev) * 0.02).bfloat16()
    bias = torch.randn(cout, device=dev)
    flops = 2.0 * B * hw * hw * cout * 9 * cin
    t0 = timeit(lambda: ops.conv3x3(x, w, bias))
    row = [f"{name:34s} hw={hw:3d} cin={cin:5d} cout={cout:5d} auto={t0:7.1f}us ({flops / t0 / 1e6:6.0f} TF/s) |"]
    for c in cands(cout, 16):
        t = timeit(lambda: ops.conv3x3(x, w, bias, block_n=c))
        row.append(f"{c}:{t:7.1f}")
    print(" ".join(row), flush=True)


print("== GEMM")
for li, (C, hw) in enumerate(LEVELS[:3]):
    M = B * hw * hw
    sweep_gemm(f"L{li} qkv", M, 3 * C, C)
    sweep_gemm(f"L{li} proj (+residual)", M, C, C, residual=True)
    sweep_gemm(f"L{li} ff1 geglu", M, 8 * C, C, geglu=True)
    sweep_gemm(f"L{li} ff2 (+residual)", M, C, 4 * C, residual=True)
for li, (C, hw) in enumerate(LEVELS):
    M = B * hw * hw
    for cin in sorted({C + LEVELS[max(li - 1, 0)][0], 2 * C, C + LEVELS[min(li + 1, 3)][0]}):
        sweep_gemm(f"L{li} shortcut 1x1", M, C, cin)
print("== conv3x3")
sweep_conv("L0", 64, 320, 320)
sweep_conv("L0 up", 64, 640, 320)
sweep_conv("L0 up", 64, 960, 320)
sweep_conv("L1 down", 32, 320, 640)
sweep_conv("L1", 32, 640, 640)
sweep_conv("L1 up", 32, 1280, 640)
sweep_conv("L1 up", 32, 960, 640)
sweep_conv("L1 up", 32, 1920, 640)
sweep_conv("L2 down", 16, 640, 1280)
sweep_conv("L2", 16, 1280, 1280)
sweep_conv("L2 up", 16, 2560, 1280)
sweep_conv("L2 up", 16, 1920, 1280)
sweep_conv("L3", 8, 1280, 1280)
sweep_conv("L3 up", 8, 2560, 1280)
