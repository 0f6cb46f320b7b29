"""Issue rate of the CUDA-core pipes the attention softmax uses (csrc/microbench.cu kinds 100+): cycles per warp instruction
per sub-partition at 1, 2, 4 and 8 warps per sub-partition."""
import os
import sys

os.environ["D4D_USE_TEST_LIB"] = "1"
import torch

sys.path.insert(0, ".")
from diffuman4d_b200._lib import check, lib  # noqa: E402

KINDS = {100: "MUFU.EX2", 101: "F2FP.BF16.PACK_AB", 102: "FFMA2", 103: "FADD2", 104: "FMNMX3", 105: "IMAD (reg mult, imm add)", 106: "PRMT",
         107: "FFMA (imm operands)", 108: "MUFU.EX2 + FFMA2 pair", 109: "FMNMX", 110: "FADD2.RM",
         111: "softmax pair: FFMA2, 2 MUFU, FADD2, F2FP", 112: "IADD imm"}
cyc = torch.zeros(148, dtype=torch.int64, device="cuda")
sink = torch.zeros(4, dtype=torch.float32, device="cuda")
iters = 2000
for kind, name in KINDS.items():
    line = f"{name:44s}"
    for warps in (4, 8, 16, 32):
        check(lib().d4d_microbench(kind, warps, iters, 148, cyc.data_ptr(), sink.data_ptr(), torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        c = cyc.float().mean().item()
        per = c / (iters * 8 * (warps / 4))
        line += f"  {warps // 4}w/smsp: {per:6.2f}"
    print(line + "   cycles per loop-body instance per sub-partition", flush=True)
