"""Synthetic stand-in for the reference's file-based ``SpaTemDataset`` (src/data/spatem_dataset.py:76-212) with the same
``get_item`` contract: label ordering of spatial / temporal samples, nearest-input-camera choice for temporal samples, the
hard-coded initial cond mask.  Tensors are derived from the labels alone, so the golden generator (which drives the
reference's ``SlidingIterativeSampler``) and the tests (which drive ``diffuman4d_b200.sampler``) see identical data."""
import math

import torch


class SyntheticSpaTemDataset:
    scene_label = "scene0"

    def __init__(self, n_cams: int, h: int = 8, w: int = 8):
        self.n_cams, self.h, self.w = n_cams, h, w

    def _pos(self, spa_label: str) -> torch.Tensor:
        a = 2 * math.pi * int(spa_label) / self.n_cams
        return torch.tensor([2.5 * math.cos(a), 0.3, 2.5 * math.sin(a)])

    def _frame(self, spa_label: str, tem_label: str):
        g = torch.Generator().manual_seed(1000 * int(spa_label) + int(tem_label) + 17)
        H, W = 8 * self.h, 8 * self.w
        pix = torch.rand(3, H, W, generator=g) * 2 - 1
        skel = torch.rand(3, H, W, generator=g) * 2 - 1
        plk = (torch.rand(6, self.h, self.w, generator=g) * 2 - 1)
        return pix, skel, plk

    def get_item(self, scene_label, spa_labels, tem_labels, input_spa_labels):
        if len(spa_labels) > 1 and len(tem_labels) == 1:          # DATA:83-88
            domain = "spatial"
            labels = [(scene_label, s, tem_labels[0]) for s in spa_labels]
        elif len(spa_labels) == 1 and len(tem_labels) > 1:
            domain = "temporal"
            d = torch.stack([self._pos(s) for s in input_spa_labels]) - self._pos(spa_labels[0])   # DATA:98-104
            cond = input_spa_labels[int(torch.argmin(torch.norm(d, dim=1)))]
            labels = [(scene_label, s, t) for s in [cond] + list(spa_labels) for t in tem_labels]
        else:
            raise ValueError(f"Error: invalid spa_labels and tem_labels: {spa_labels} and {tem_labels}")
        frames = [self._frame(s, t) for _, s, t in labels]
        pixel_values = torch.stack([f[0] for f in frames])
        cond_masks = torch.ones_like(pixel_values)[:, :1, ...]
        cond_masks[len(pixel_values) // 2:, ...] = 0.0            # DATA:171 ("hard code"; the sampler overwrites it)
        return {"domain": domain, "labels": labels, "pixel_values": pixel_values,
                "plucker_embeds": torch.stack([f[2] for f in frames]), "skeletons": torch.stack([f[1] for f in frames]),
                "cond_masks": cond_masks}
