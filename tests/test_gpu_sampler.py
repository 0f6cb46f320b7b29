"""GPU tests of the two outer seams driving the CUDA path end to end:

* B-1 ``loader.load_pipelines`` on a diffusers-layout checkpoint directory (unet/config.json + safetensors + scheduler
  config), then the reference sampler's exact call pattern (src/samplers/sliding_iterative_sampler.py:155-190:
  ``pipeline.sliding_iterative_denoise(pixel_values=..., latents=None, ...)`` followed by ``result["images"].float().cpu()``);
* ``B200SlidingIterativeSampler`` with the V x T grid resident on the device over three alternation rounds
  (sampler logic itself is pinned against the reference sampler on the CPU, tests/test_sampler.py).
"""
import json
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)


def _tiny_checkpoint(tmp_path):
    from safetensors.torch import save_file
    from diffuman4d_b200.config import UNetConfig
    from diffuman4d_b200.weights import random_state_dict
    cfg = UNetConfig.tiny()
    os.makedirs(tmp_path / "unet")
    os.makedirs(tmp_path / "scheduler")
    json.dump(dict(in_channels=11, out_channels=4, block_out_channels=[64, 128, 256, 256], attention_head_dim=[1, 2, 4, 4],
                   cross_attention_dim=None, use_linear_projection=True, enable_pose_encoder=True, enable_tem_embeds=True,
                   layers_per_block=2, num_3d_attn_blocks=3), open(tmp_path / "unet" / "config.json", "w"))
    json.dump({"_class_name": "DDIMScheduler", "beta_schedule": "scaled_linear", "beta_start": 0.00085, "beta_end": 0.012,
               "clip_sample": False, "set_alpha_to_one": False, "steps_offset": 1, "prediction_type": "epsilon"},
              open(tmp_path / "scheduler" / "scheduler_config.json", "w"))
    save_file({k: v.contiguous() for k, v in random_state_dict(cfg, seed=1).items()},
              str(tmp_path / "unet" / "diffusion_pytorch_model.safetensors"))
    return cfg


def test_load_pipelines_and_reference_sampler_call_pattern(cuda, tmp_path):
    from diffuman4d_b200.loader import load_pipelines
    from synthetic_dataset import SyntheticSpaTemDataset
    cfg = _tiny_checkpoint(tmp_path)
    pipes = load_pipelines(model_dir=str(tmp_path), torch_dtype="bf16", gpu_ids=[0], vae_factory="pool_vae.make")
    assert len(pipes) == 1 and pipes[0].unet.config == cfg and pipes[0].vae is not None
    with pytest.raises(ValueError, match="Unsupported torch_dtype"):
        load_pipelines(model_dir=str(tmp_path), torch_dtype="fp16", gpu_ids=[0])
    pipe = pipes[0]
    pipe.to("cuda:0")                                               # SUTIL:47
    pipe.set_progress_bar_config(disable=True)                      # SUTIL:48
    ds = SyntheticSpaTemDataset(6, h=8, w=8)
    sample = ds.get_item(ds.scene_label, [f"{i:02d}" for i in range(6)], ["000000"], ["01", "04"])
    inp = torch.tensor([1, 4])
    sample["cond_masks"][...] = 1.0
    sample["cond_masks"][inp] = 0.0
    seen = []
    result = pipe.sliding_iterative_denoise(
        pixel_values=sample["pixel_values"], plucker_embeds=sample["plucker_embeds"], skeletons=sample["skeletons"],
        cond_masks=sample["cond_masks"], latents=None, domain="spatial", timestep_indices=torch.tensor([0] * 6),
        window_size=2, sliding_stride=1, sliding_shift=0, bidirectional=True, num_denoising_steps=1, alternation_rounds=3,
        guidance_scale=2.0, tqdm=lambda it, **kw: (seen.append(kw.get("total")), it)[1])
    images = result["images"].float().cpu()                         # SAMP:187
    assert images.shape == (6, 3, 64, 64) and torch.isfinite(images).all() and 0 <= float(images.min()) <= float(images.max()) <= 1
    ti = result["timestep_indices"].cpu()                           # SAMP:188
    assert ti[inp].eq(0).all() and ti[[0, 2, 3, 5]].eq(4).all() and not result["fully_denoised"].cpu().any()
    assert seen == [8]                                              # 4 targets / stride 1, both directions
    assert result["latents"].shape == (6, 4, 8, 8)
    # cond frames come back as the encoded image latents (PIPE:375-379 aliasing)
    z = pipe.vae.encode_latents(sample["pixel_values"].to("cuda", torch.bfloat16))
    assert torch.equal(result["latents"][inp.cuda()], z[inp.cuda()])
    # Diffuman4DPipeline.__call__ with latents=None draws the initial noise (PIPE:172-183)
    out = pipe(pixel_values_latents=z, plucker_embeds_latents=sample["plucker_embeds"], skeletons_latents=sample["skeletons"],
               cond_masks_latents=sample["cond_masks"][..., ::8, ::8].contiguous(), latents=None, domains=["spatial"],
               num_inference_steps=2, guidance_scale=2.0, generator=torch.Generator(device="cuda").manual_seed(3))
    assert out.shape == (6, 4, 8, 8) and torch.isfinite(out.float()).all()


def test_sampler_drives_b200_pipeline(cuda):
    from pool_vae import PoolVAE
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
