"""GPU smoke of the device-resident sampler driving the B200 pipeline end to end: three alternation rounds
over a 6-camera x 4-frame grid with the tiny UNet.  Parity of the sampler logic itself is pinned on the CPU
(tests/test_sampler.py); here the grid must live on the GPU and the reference's bookkeeping invariants must hold."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu  # not collected: lives outside tests/ until it has run green on a B200

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))


class PoolVAE:
    """8x average pooling stand-in for AutoencoderKL (the VAE is outside the hot path)"""

    def encode_latents(self, x):
        z = F.avg_pool2d(x.float(), 8)
        return torch.cat([z, z.mean(dim=1, keepdim=True)], dim=1).to(torch.bfloat16)

    def decode_latents(self, latents):
        return latents


def test_sampler_drives_b200_pipeline(cuda):
    from synthetic_dataset import SyntheticSpaTemDataset
    from diffuman4d_b200.config import SchedulerConfig, UNetConfig
    from diffuman4d_b200.pipeline import B200Diffuman4DPipeline
    from diffuman4d_b200.sampler import B200SlidingIterativeSampler
    from diffuman4d_b200.unet import B200MultiviewUNet
    from diffuman4d_b200.weights import random_state_dict

    cfg = UNetConfig.tiny()
    unet = B200MultiviewUNet(cfg, 0).load_state_dict(random_state_dict(cfg, seed=1))
    vae = PoolVAE()
    pipe = B200Diffuman4DPipeline(unet, SchedulerConfig(), vae=vae)
    ds = SyntheticSpaTemDataset(8, h=16, w=16)
    s = B200SlidingIterativeSampler(ds, [pipe], output_dir=None, spa_label_range=[0, 6, 1], tem_label_range=[0, 4, 1],
                                    input_spa_labels=[1, 4], window_size=2, sliding_stride=1, bidirectional=True,
                                    alternation_rounds=3, guidance_scale=2.0)
    s.execute_tasks()
    torch.cuda.synchronize()
    assert s.grid_latents.is_cuda and s.grid_latents.dtype == torch.bfloat16 and s.grid_latents.shape == (6, 4, 4, 16, 16)
    assert torch.isfinite(s.grid_latents.float()).all()
    ti = s.grid_timestep_indices.cpu()
    n_inf = 2 * 1 // 1 * 2 * 3                                   # window * steps / stride, bidirectional, 3 rounds
    for v, spa in enumerate(s.spa_labels):
        expect = 0 if spa in s.input_spa_labels else n_inf      # every target cell fully denoised, inputs untouched
        assert (ti[v] == expect).all(), (spa, ti[v])
    # cond cells come back as the encoded image latents (the reference's aliasing of latent_model_input, PIPE:375-379)
    for spa in s.input_spa_labels:
        for tem in s.tem_labels:
            pix = ds.get_item(ds.scene_label, [spa, s.target_spa_labels[0]], [tem], s.input_spa_labels)["pixel_values"][:1]
            ref = vae.encode_latents(pix.to(torch.bfloat16))[0]      # the pipeline encodes the bf16 image
            torch.testing.assert_close(s.latent(spa, tem).cpu().float(), ref.float(), rtol=2e-2, atol=2e-2)
