"""Phase timeline of the softmax warps of one CTA of the head_dim-64 attention kernel (tools build): SM-clock stamps at
tile start / after the row max / after chunk 0 / after the P-free wait / after chunks 1..3 / after the hand-off, for warp 0
of every Q tile (they share sub-partition 0).  D4D_ATTN_SHAPE / D4D_ATTN_VARIANT choose the kernel.
Needs a tools build with -DD4D_ATTN_TRACE (add it to TEST_DEFS in diffuman4d_b200/build.py): the stamps cost registers."""
import os
import sys

os.environ["D4D_USE_TEST_LIB"] = "1"
import torch

sys.path.insert(0, ".")
from diffuman4d_b200 import ops  # noqa: E402

b, s, h = 2, 16384, 10
nq = 3 if os.environ.get("D4D_ATTN_SHAPE", "0") == "1" else 2
qkv = torch.randn(b * s, 3 * h * 64, device="cuda").to(torch.bfloat16)
ops.attention(qkv, b, s, h, 64, 0.125)
buf = torch.zeros(nq * 64 * 8, dtype=torch.int64, device="cuda")
os.environ["D4D_ATTN_TRACE"] = str(buf.data_ptr())
ops.attention(qkv, b, s, h, 64, 0.125)
torch.cuda.synchronize()
os.environ.pop("D4D_ATTN_TRACE")
t = buf.cpu().view(nq, 64, 8)
t0 = int(t[:, 0, 0].min())
names = ["start", "max", "c0", "pfree", "c1", "c2", "c3", "done"]
print("tile  " + "   ".join(f"q{q}: " + " ".join(f"{n:>6s}" for n in names) for q in range(nq)))
for j in range(8, 28):
    print(f"{j:3d}   " + "   ".join("    " + " ".join(f"{int(t[q, j, e]) - t0:6d}" for e in range(8)) for q in range(nq)))
for q in range(nq):
    d = (t[q, 10:60, :] - t[q, 10:60, 0:1]).float().mean(0)
    per = (t[q, 11:61, 0] - t[q, 10:60, 0]).float().mean().item()
    print(f"q{q}: mean offsets within a tile {[int(x) for x in d]}  period {per:.0f}")
if nq >= 2:
    print("q1 start - q0 start:", [(int(t[1, j, 0]) - int(t[0, j, 0])) for j in range(10, 40)])
