"""Weight-key contract of the UNet (diffusers layout) and seeded random weights.

``state_dict_spec`` lists every parameter of ``unet/diffusion_pytorch_model.safetensors`` for a given config with its
diffusers shape; the C++ loader (csrc/unet.cu ``declare_keys``) and the CPU oracle must agree with it (tested).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Tuple

import torch

from .config import UNetConfig

POSE_SPEC = [(3, 3, 3), (3, 16, 4), (16, 16, 3), (16, 32, 4), (32, 32, 3), (32, 64, 4), (64, 64, 3), (64, 128, 3)]


def state_dict_spec(cfg: UNetConfig) -> "OrderedDict[str, Tuple[int, ...]]":
    spec: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    ch = cfg.block_out_channels
    C0, TE, L = ch[0], cfg.time_embed_dim, cfg.layers_per_block

    def lin(p, out, inp, bias=True):
        spec[p + ".weight"] = (out, inp)
        if bias:
            spec[p + ".bias"] = (out,)

    def conv(p, out, inp, k):
        spec[p + ".weight"] = (out, inp, k, k)
        spec[p + ".bias"] = (out,)

    def norm(p, c):
        spec[p + ".weight"] = (c,)
        spec[p + ".bias"] = (c,)

    def resnet(p, cin, cout):
        norm(p + ".norm1", cin)
        conv(p + ".conv1", cout, cin, 3)
        lin(p + ".time_emb_proj", cout, TE)
        norm(p + ".norm2", cout)
        conv(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(p + ".conv_shortcut", cout, cin, 1)

    def xf(p, C, attn2):
        norm(p + ".norm", C)
        if cfg.use_linear_projection:
            lin(p + ".proj_in", C, C)
        else:
            conv(p + ".proj_in", C, C, 1)
        b = p + ".transformer_blocks.0"
        norm(b + ".norm1", C)
        for n in ("to_q", "to_k", "to_v"):
            lin(f"{b}.attn1.{n}", C, C, bias=False)
        lin(b + ".attn1.to_out.0", C, C)
        if attn2:
            norm(b + ".norm2", C)
            for n in ("to_q", "to_k", "to_v"):
                lin(f"{b}.attn2.{n}", C, C, bias=False)
            lin(b + ".attn2.to_out.0", C, C)
        norm(b + ".norm3", C)
        lin(b + ".ff.net.0.proj", 8 * C, C)
        lin(b + ".ff.net.2", C, 4 * C)
        if cfg.use_linear_projection:
            lin(p + ".proj_out", C, C)
        else:
            conv(p + ".proj_out", C, C, 1)

    conv("conv_in", C0, cfg.in_channels, 3)
    lin("time_embedding.linear_1", TE, C0)
    lin("time_embedding.linear_2", TE, TE)
    if cfg.enable_tem_embeds:
        lin("temporal_pos_embed.linear_1", TE, C0)
        lin("temporal_pos_embed.linear_2", TE, TE)
    if cfg.enable_pose_encoder:
        for i, (ci, co, k) in enumerate(POSE_SPEC):
            conv(f"pose_encoder.conv_layers.{2 * i}", co, ci, k)
        conv("pose_encoder.final_proj", C0, 128, 1)
        spec["pose_encoder.scale"] = (1,)
    cout = C0
    for i in range(4):
        cin, cout = cout, ch[i]
        for j in range(L):
            resnet(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
            if i < 3:
                xf(f"down_blocks.{i}.attentions.{j}", cout, cfg.has_attn2(i))
        if i < 3:
            conv(f"down_blocks.{i}.downsamplers.0.conv", cout, cout, 3)
    resnet("mid_block.resnets.0", ch[3], ch[3])
    xf("mid_block.attentions.0", ch[3], cfg.has_attn2(3))
    resnet("mid_block.resnets.1", ch[3], ch[3])
    cout = ch[3]
    for i in range(4):
        cprev, cout = cout, ch[3 - i]
        cin = ch[3 - min(i + 1, 3)]
        for j in range(L + 1):
            skip = cin if j == L else cout
            rin = cprev if j == 0 else cout
            resnet(f"up_blocks.{i}.resnets.{j}", rin + skip, cout)
            if i > 0:
                xf(f"up_blocks.{i}.attentions.{j}", cout, cfg.has_attn2(3 - i))
        if i < 3:
            conv(f"up_blocks.{i}.upsamplers.0.conv", cout, cout, 3)
    norm("conv_norm_out", C0)
    conv("conv_out", cfg.out_channels, C0, 3)
    return spec


def random_state_dict(cfg: UNetConfig, seed: int = 1, dtype=torch.bfloat16) -> Dict[str, torch.Tensor]:
    """Seeded random weights (no network for the real checkpoint).  Fan-in-scaled normal weights, small biases,
    randomised norm affines and NON-zero values for the reference's zero-initialised branches
    (pose_encoder.final_proj, temporal_pos_embed.linear_2) so those paths carry signal."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for key, shape in state_dict_spec(cfg).items():
        if key.endswith("scale"):
            t = torch.full(shape, 2.0)
        else:
            is_norm = ".norm" in key or key.startswith("conv_norm_out")
            if is_norm and key.endswith("weight"):
                t = 1.0 + 0.2 * torch.randn(shape, generator=g)
            elif is_norm and key.endswith("bias"):
                t = 0.1 * torch.randn(shape, generator=g)
            elif key.endswith("bias"):
                t = 0.05 * torch.randn(shape, generator=g)
            else:
                fan_in = 1
                for s in shape[1:]:
                    fan_in *= s
                t = torch.randn(shape, generator=g) / math.sqrt(fan_in)
        sd[key] = t.to(dtype)
    return sd
