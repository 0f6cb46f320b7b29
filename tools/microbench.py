"""Per-SM throughput of the units that bound the attention softmax (run on a B200)."""
import sys

import torch

sys.path.insert(0, ".")
from diffuman4d_b200._lib import check, test_lib as lib  # noqa: E402  (tools build: libd4d_test.so)

cyc = torch.zeros(148, dtype=torch.int64, device="cuda")
sink = torch.zeros(4, device="cuda")
names = {0: "tcgen05.ld 32x32b.x32 (4 KB/warp-instr)", 1: "tcgen05.ld 32x32b.x16 (2 KB)", 2: "tcgen05.st 32x32b.x16 (2 KB)",
         3: "ex2.approx (8 / iter)", 4: "cvt.rn.bf16x2.f32 (4 / iter)", 5: "2 x tcgen05.ld x16, one wait (4 KB)",
         6: "2 x tcgen05.ld x32, one wait (8 KB)"}
per_iter_bytes = {0: 4096, 1: 2048, 2: 2048, 5: 4096, 6: 8192}
for kind in (7, 8):
    iters = 2000
    for _ in range(2):
        check(lib().d4d_microbench(kind, 4, iters, 148, cyc.data_ptr(), sink.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    print(f"tcgen05.commit -> mbarrier -> try_wait round trip ({'after one MMA' if kind == 8 else 'empty pipe'}): {cyc.float().mean().item() / iters:8.1f} cycles", flush=True)
pipe = {9: ("N=256, no commits", 256), 10: ("N=256, commit / 4 MMAs", 256), 14: ("N=256, commit / 16 MMAs", 256),
        15: ("N=160, no commits", 160), 16: ("N=160, commit / 4 MMAs", 160), 11: ("N=64, no commits", 64),
        12: ("N=64, commit / 4 MMAs", 64), 13: ("commits only", 0)}
for kind, (label, n) in pipe.items():
    iters = 4096
    for _ in range(2):
        check(lib().d4d_microbench(kind, 4, iters, 148, cyc.data_ptr(), sink.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    c = cyc.float().mean().item() / iters
    print(f"tensor pipe, 128xNx16 MMAs back to back, {label:26s}: {c:7.1f} cycles / MMA" + (f"  ({128 * n * 16 * 2 / c:7.0f} FLOP/clk/SM)" if n else ""), flush=True)
extra = {20: ("N=128, no commits", 128), 21: ("N=192, no commits", 192), 22: ("N=224, no commits", 224), 24: ("N=240, no commits", 240),
         23: ("N=32, no commits", 32), 19: ("N=64, issued by 2 threads", 64), 17: ("tcgen05.fence::after only", 0),
         18: ("successful try_wait only", 0)}
for kind, (label, n) in extra.items():
    iters = 4096
    for _ in range(2):
        check(lib().d4d_microbench(kind, 4, iters, 148, cyc.data_ptr(), sink.data_ptr(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    c = cyc.float().mean().item() / iters
    print(f"{label:34s}: {c:7.1f} cycles / iteration" + (f"  ({128 * n * 16 * 2 / c:7.0f} FLOP/clk/SM)" if n else ""), flush=True)
if "--pipe-only" in sys.argv:
    sys.exit(0)
for kind in (0, 1, 5, 6, 2, 3, 4):
    for warps in (4, 8, 16):
        iters = 4000
        for _ in range(2):
            check(lib().d4d_microbench(kind, warps, iters, 148, cyc.data_ptr(), sink.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        c = cyc.float().mean().item()
        if kind in per_iter_bytes:
            print(f"{names[kind]:42s} warps/CTA={warps:2d}: {c / iters:8.1f} cyc/iter/warp-set, {per_iter_bytes[kind] * warps * iters / c:8.1f} B/clk/SM")
        else:
            n = {3: 8, 4: 4}[kind]
            print(f"{names[kind]:42s} warps/CTA={warps:2d}: {n * warps * 32 * iters / c:8.2f} lanes/clk/SM")
