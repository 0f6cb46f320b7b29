"""Time the attention kernel on the two dominant shapes of the W16@64x64 step; with D4D_ATTN_ABLATE=<bits> parts of the
kernel are switched off (results are then wrong) to see what the time is sensitive to."""
import os
import sys

os.environ["D4D_USE_TEST_LIB"] = "1"   # the ablation switches exist only in the tools build (libd4d_test.so)

import torch

sys.path.insert(0, ".")
from diffuman4d_b200 import ops  # noqa: E402

shapes = [("L0 2-D  b32 s4096 h5", 32, 4096, 5), ("L1 3-D  b2 s16384 h10", 2, 16384, 10), ("L2 3-D  b2 s4096 h20", 2, 4096, 20)]
for name, b, s, h in shapes:
    qkv = (torch.randn(b * s, 3 * h * 64, device="cuda") * 1.0).to(torch.bfloat16)
    for _ in range(2):
        ops.attention(qkv, b, s, h, 64, 0.125)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    n = 5
    for _ in range(n):
        ops.attention(qkv, b, s, h, 64, 0.125)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    fl = 4.0 * b * h * s * s * 64
    print(f"ablate={os.environ.get('D4D_ATTN_ABLATE', '0'):>2s} {name}: {ms:8.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s", flush=True)
