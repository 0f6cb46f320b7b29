"""Generate golden fixtures by RUNNING code from /root/reference (only possible in the build
container -- /root/reference does not exist on the GPU box, so the outputs are committed).

Pinned here:
  pose_encoder.pt   -- the reference ``PoseEncoder`` (src/diffusers/models/unets/pose_encoder.py), imported
                       directly by file path (it only needs torch/numpy/einops).
  mv_block.pt       -- the reference ``MultiviewTransformerBlock.forward`` (src/diffusers/models/attention.py:22-153)
                       running ITS OWN code on top of a stub of the un-installed upstream base class
                       ``diffusers.models.attention.BasicTransformerBlock`` (the stub = our restatement of
                       upstream 0.33.1; the 3-D token reshape and residual wiring are the reference's).
  plucker.pt        -- ``calc_plucker_embeds`` from src/data/utils/ray_utils.py on a small camera ring (realistic
                       value ranges for the bench's synthetic Plucker channels).

Run:  python tests/golden/gen_golden.py      (from the repo root, inside the build container)
"""
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def load_by_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def gen_pose_encoder():
    pe_mod = load_by_path("ref_pose_encoder", f"{REF}/src/diffusers/models/unets/pose_encoder.py")
    torch.manual_seed(11)
    pe = pe_mod.PoseEncoder(out_channels=32)
    with torch.no_grad():  # un-zero the zero-init projection so the output is informative
        pe.final_proj.weight.normal_(0, 0.1)
        pe.final_proj.bias.normal_(0, 0.1)
        for m in pe.conv_layers:
            if isinstance(m, nn.Conv2d):
                m.bias.normal_(0, 0.1)
    x = torch.rand(2, 3, 32, 40) * 2 - 1
    with torch.no_grad():
        y = pe(x)
    torch.save({"state_dict": pe.state_dict(), "x": x, "y": y}, f"{HERE}/pose_encoder.pt")
    print("pose_encoder", tuple(y.shape), float(y.abs().mean()))


def gen_mv_block():
    from oracle import unet_oracle as O

    class StubAttention(O.Attention):
        def forward(self, x, encoder_hidden_states=None, attention_mask=None, **kw):
            assert encoder_hidden_states is None and attention_mask is None
            return super().forward(x)

    class StubBasicTransformerBlock(nn.Module):
        """minimal upstream BasicTransformerBlock (layer_norm, geglu) for the reference subclass"""

        def __init__(self, dim, heads, attn2):
            super().__init__()
            self.norm_type = "layer_norm"
            self.pos_embed = None
            self.only_cross_attention = False
            self._chunk_size = None
            self._chunk_dim = 0
            self.norm1 = nn.LayerNorm(dim, eps=1e-5)
            self.attn1 = StubAttention(dim, heads)
            self.norm2 = nn.LayerNorm(dim, eps=1e-5) if attn2 else None
            self.attn2 = StubAttention(dim, heads) if attn2 else None
            self.norm3 = nn.LayerNorm(dim, eps=1e-5)
            self.ff = O.FeedForward(dim)

    # stub the import surface of src/diffusers/models/attention.py:1-10
    d = types.ModuleType("diffusers")
    du = types.ModuleType("diffusers.utils")
    dut = types.ModuleType("diffusers.utils.torch_utils")
    dm = types.ModuleType("diffusers.models")
    dma = types.ModuleType("diffusers.models.attention")
    import logging as _logging
    du.logging = types.SimpleNamespace(get_logger=_logging.getLogger)
    dut.maybe_allow_in_graph = lambda c: c
    dma._chunked_feed_forward = None
    dma.BasicTransformerBlock = StubBasicTransformerBlock
    for name, mod in [("diffusers", d), ("diffusers.utils", du), ("diffusers.utils.torch_utils", dut),
                      ("diffusers.models", dm), ("diffusers.models.attention", dma)]:
        sys.modules[name] = mod
    ref_att = load_by_path("ref_attention", f"{REF}/src/diffusers/models/attention.py")

    out = {}
    for tag, attn2 in (("no_attn2", False), ("attn2", True)):
        torch.manual_seed(5 if attn2 else 4)
        blk = ref_att.MultiviewTransformerBlock(64, 2, attn2)
        with torch.no_grad():
            for p in blk.parameters():
                p.normal_(0, 0.15)
        x = torch.randn(6, 16, 64)  # (b t) hw c with b=2, t=3
        with torch.no_grad():
            y3 = blk(x, num_frames=3)
            y1 = blk(x, num_frames=1)
        out[tag] = {"state_dict": blk.state_dict(), "x": x, "y_3d": y3, "y_2d": y1}
        print("mv_block", tag, float(y3.abs().mean()), float((y3 - y1).abs().mean()))
    torch.save(out, f"{HERE}/mv_block.pt")


def gen_plucker():
    ray = load_by_path("ref_ray_utils", f"{REF}/src/data/utils/ray_utils.py")
    import math
    n, h, w = 6, 16, 16
    Ks, c2ws = [], []
    for i in range(n):
        a = 2 * math.pi * i / n
        eye = torch.tensor([2.5 * math.cos(a), 0.3, 2.5 * math.sin(a)])
        fwd = -eye / eye.norm()
        up = torch.tensor([0.0, 1.0, 0.0])
        right = torch.linalg.cross(fwd, up)
        right = right / right.norm()
        up2 = torch.linalg.cross(right, fwd)
        c2w = torch.eye(4)
        c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, -up2, fwd, eye
        c2ws.append(c2w)
        K = torch.tensor([[1.2 * w, 0, w / 2], [0, 1.2 * w, h / 2], [0, 0, 1.0]])
        Ks.append(K)
    Ks, c2ws = torch.stack(Ks), torch.stack(c2ws)
    try:
        pl = ray.calc_plucker_embeds(h, w, Ks, c2ws)
    except TypeError:
        import inspect
        print("calc_plucker_embeds signature:", inspect.signature(ray.calc_plucker_embeds))
        raise
    torch.save({"K": Ks, "c2w": c2ws, "plucker": pl}, f"{HERE}/plucker.pt")
    print("plucker", tuple(pl.shape), float(pl.min()), float(pl.max()))


if __name__ == "__main__":
    gen_pose_encoder()
    gen_mv_block()
    try:
        gen_plucker()
    except Exception as e:  # value-range fixture only; not a parity gate
        print("plucker fixture skipped:", repr(e))
