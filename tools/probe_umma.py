"""Sweep UMMA operand encodings on a B200 and print which ones reproduce torch.matmul.
Each variant runs in this process; a failing variant gives wrong numbers, not a hang (the probe kernel
traps on barrier timeouts).  Usage: python tools/probe_umma.py > gpurun_out/probe.txt"""
import itertools
import json
import sys

import torch

sys.path.insert(0, ".")
from diffuman4d_b200 import ops  # noqa: E402


def run(N, K, a_src, b_major, lbo, sbo, kadv):
    g = torch.Generator(device="cpu").manual_seed(N * 1000 + K)
    A = torch.randn(128, K, generator=g).to(torch.bfloat16).cuda()
    if b_major == 0:
        B = torch.randn(N, K, generator=g).to(torch.bfloat16).cuda()
        ref = A.float() @ B.float().t()
    else:
        B = torch.randn(K, N, generator=g).to(torch.bfloat16).cuda()
        ref = A.float() @ B.float()
    D = ops.probe_umma(A, B, N, K, a_src, b_major, lbo, sbo, kadv)
    torch.cuda.synchronize()
    err = (D - ref).abs().max().item()
    return err, ref.abs().max().item()


def main():
    results = []
    for N, K in [(64, 64), (128, 64), (64, 128), (128, 128)]:
        for a_src in (0, 1):
            err, scale = run(N, K, a_src, 0, 0, 1024, 32)
            results.append(dict(N=N, K=K, a_src=a_src, b_major=0, err=err, scale=scale))
            print(f"KK   N={N:3d} K={K:3d} a_src={a_src} err={err:.4g} (scale {scale:.3g})", flush=True)
    # MN-major B ([K, N] row-major, the V operand).  Candidates for (LBO, SBO, per-16-K advance)
    cands = []
    for lbo, sbo, kadv in itertools.product([0, 1024, 8192, 16384], [1024, 2048, 128, 8192, 16384], [2048, 256, 32, 4096]):
        cands.append((lbo, sbo, kadv))
    for N, K in [(64, 128), (128, 128), (64, 64), (128, 64)]:
        good = []
        for a_src in (0, 1):
            for lbo, sbo, kadv in cands:
                try:
                    err, scale = run(N, K, a_src, 1, lbo, sbo, kadv)
                except Exception as e:  # noqa: BLE001
                    print("EXC", N, K, a_src, lbo, sbo, kadv, repr(e), flush=True)
                    raise
                ok = err < 1e-2 * max(scale, 1.0)
                results.append(dict(N=N, K=K, a_src=a_src, b_major=1, lbo=lbo, sbo=sbo, kadv=kadv, err=err, ok=ok))
                if ok:
                    good.append((a_src, lbo, sbo, kadv))
        print(f"MN   N={N:3d} K={K:3d} good (a_src,lbo,sbo,kadv): {good}", flush=True)
    json.dump(results, open("gpurun_out/probe.json", "w"))


if __name__ == "__main__":
    main()
