"""Plücker ray embeddings on the device (input pipeline).

Closed-form restatement of ``calc_plucker_embeds`` / ``calc_relative_poses`` (reference: src/data/utils/ray_utils.py:101-118,
which builds per-pixel 3x3 matrix stacks through ``get_rays``): for camera-to-world pose ``[R_c | t]`` and intrinsics K the ray
of pixel (i, j) is ``d = normalize(R_c K^-1 (j + 0.5, i + 0.5, 1)^T)``, origin ``o = t`` and the embedding is ``(d, o x d)``.
Runs on whatever device the inputs live on (a few fused elementwise torch ops over [F, h, w, 3]); the reference computes it on
the CPU per task (src/data/spatem_dataset.py:165).  Pinned against the reference function in tests/test_host.py.
"""
from __future__ import annotations

import torch


def relative_poses(poses: torch.Tensor) -> torch.Tensor:
    """ray_utils.py:114-118: poses relative to the first one."""
    return torch.linalg.inv(poses[0]) @ poses


def plucker_embeds(h: int, w: int, K: torch.Tensor, pose: torch.Tensor, eps: float = 1e-8) -> torch.Tensor:
    """K [F,3,3], pose (camera-to-world) [F,4,4] -> [F, 6, h, w] (ray direction | origin x direction)."""
    dt, dev = pose.dtype, pose.device
    R, t = pose[:, :3, :3], pose[:, :3, 3]
    i = torch.arange(h, dtype=dt, device=dev) + 0.5
    j = torch.arange(w, dtype=dt, device=dev) + 0.5
    xy1 = torch.stack([j[None, :].expand(h, w), i[:, None].expand(h, w), torch.ones(h, w, dtype=dt, device=dev)], dim=-1)
    M = R @ torch.linalg.inv(K.float()).to(dt)                      # [F,3,3]: pixel -> world direction
    d = torch.einsum("fab,hwb->fhwa", M, xy1)
    d = d / (torch.linalg.norm(d, dim=-1, keepdim=True) + eps)
    o = t[:, None, None, :].expand_as(d)
    return torch.cat([d, torch.linalg.cross(o, d, dim=-1)], dim=-1).permute(0, 3, 1, 2).contiguous()
