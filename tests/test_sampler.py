"""CPU tests of ``diffuman4d_b200.sampler`` (device-resident V x T grid) against the reference's own
``SlidingIterativeSampler`` run end to end on stubs (tests/golden/gen_golden.py::gen_sampler -> golden/sampler_ref.pt),
plus the 2-rank (gloo) round exchange."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

from diffuman4d_b200.config import SchedulerConfig
from diffuman4d_b200.sampler import B200SlidingIterativeSampler
from oracle.pipeline_oracle import DDIMOracle, sliding_iterative_denoise_oracle

GOLD = os.path.join(os.path.dirname(__file__), "golden")
sys.path.insert(0, GOLD)
from fake_unet import make_fake_unet  # noqa: E402
from synthetic_dataset import SyntheticSpaTemDataset  # noqa: E402


class OraclePipeline:
    """``sliding_iterative_denoise`` (PIPE:439-559) on the oracle, with the stand-ins the golden generator gave the
    reference pipeline: fake UNet, 8x-average-pooling "VAE", identity decode, k-th noise draw from Generator(9000 + k)."""

    def __init__(self, noise: str = "counter"):
        self.noise, self.calls = noise, 0
        self.unet = make_fake_unet(11)

    def sliding_iterative_denoise(self, pixel_values, plucker_embeds, skeletons, cond_masks, latents, domain,
                                  timestep_indices, window_size, sliding_stride, sliding_shift, bidirectional,
                                  num_denoising_steps, alternation_rounds, guidance_scale, **kw):
        z = F.avg_pool2d(pixel_values, 8)
        pix = torch.cat([z, z.mean(dim=1, keepdim=True)], dim=1)
        h, w = pix.shape[-2:]
        mask = F.interpolate(cond_masks, size=(h, w), mode="nearest")
        if latents is None:
            seed = 9000 + self.calls if self.noise == "counter" else int(pixel_values.abs().sum().item() * 1000) % (2 ** 31)
            self.calls += 1
            latents = torch.randn((len(pix), 4, h, w), generator=torch.Generator().manual_seed(seed))
        out = sliding_iterative_denoise_oracle(
            self.unet, DDIMOracle(SchedulerConfig()), pixel_latents=pix, plucker=plucker_embeds, skeletons=skeletons,
            cond_mask=mask, latents=latents, domain=domain, timestep_indices=timestep_indices.cpu(), window_size=window_size,
            sliding_stride=sliding_stride, sliding_shift=sliding_shift, bidirectional=bidirectional,
            num_denoising_steps=num_denoising_steps, alternation_rounds=alternation_rounds, guidance_scale=guidance_scale,
            enable_pose_encoder=True)
        out["images"] = out["latents"]
        return out


def _make(kwargs, n_cams, pipe, save_fn=None):
    return B200SlidingIterativeSampler(dataset=SyntheticSpaTemDataset(n_cams), pipelines=[pipe], output_dir=None,
                                       num_denoising_steps=1, guidance_scale=2.0, sliding_shift=0, save_fn=save_fn, **kwargs)


@pytest.mark.parametrize("tag", ["v6_t4_stride1", "v5_t2_stride2_unidir"])
def test_sampler_matches_reference_sampler_golden(tag):
    c = torch.load(os.path.join(GOLD, "sampler_ref.pt"))["cases"][tag]
    saved = []
    s = _make(c["kwargs"], c["n_cams"], OraclePipeline(), save_fn=lambda sample, out_dir: saved.append(sample))
    assert s.all_tasks == c["all_tasks"]                                  # SAMP:192-199
    s.execute_tasks()
    for (spa, tem), ref in c["grid_latents"].items():                     # final grid == the reference's dict of latents
        torch.testing.assert_close(s.latent(spa, tem), ref, rtol=1e-5, atol=1e-5)
        assert s.timestep_index(spa, tem) == c["grid_timestep_indices"][(spa, tem)]
    assert len(saved) == len(c["saved"])
    for got, ref in zip(saved, c["saved"]):                               # per task: order, labels, bookkeeping
        assert (got["alt"], got["domain"], got["domain_label"]) == (ref["alt"], ref["domain"], ref["domain_label"])
        assert [tuple(x) for x in got["labels"]] == [tuple(x) for x in ref["labels"]]
        assert torch.equal(got["timestep_indices"].cpu(), ref["timestep_indices"])
        assert torch.equal(got["fully_denoised"].cpu(), ref["fully_denoised"])


def test_sampler_argument_errors_match_reference_messages():
    errs = torch.load(os.path.join(GOLD, "sampler_ref.pt"))["errors"]
    cases = {"window_gt_targets": dict(spa_label_range=[0, 4, 1], input_spa_labels=[1], window_size=4),
             "targets_mod_stride": dict(spa_label_range=[0, 6, 1], input_spa_labels=[1], window_size=2, sliding_stride=2),
             "tems_mod_stride": dict(spa_label_range=[0, 6, 1], input_spa_labels=[1, 4], tem_label_range=[0, 3, 1],
                                     window_size=2, sliding_stride=2),
             "window_gt_tems": dict(spa_label_range=[0, 6, 1], input_spa_labels=[1, 4], tem_label_range=[0, 1, 1],
                                    window_size=2, alternation_rounds=2),
             "no_spa": dict(spa_label_range=None, spa_labels=None)}
    for tag, kw in cases.items():
        assert errs[tag] is not None
        with pytest.raises(ValueError) as e:
            B200SlidingIterativeSampler(dataset=None, pipelines=[], **{"tem_label_range": [0, 4, 1], **kw})
        assert str(e.value) == errs[tag]


def _rank_main(rank, world, port, kwargs, n_cams, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    s = _make(kwargs, n_cams, OraclePipeline(noise="content"))
    s.execute_tasks(rank=rank, world=world)
    q.put((rank, s.grid_latents.tolist(), s.grid_timestep_indices.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_sampler_two_ranks_equal_single_process():
    """Round-sharded tasks + one all-gather of the updated cells per round (gloo, world_size 2) reproduce the
    single-process grid on every rank."""
    import torch.multiprocessing as mp
    kwargs = dict(spa_label_range=[0, 6, 1], tem_label_range=[0, 4, 1], input_spa_labels=[1, 4], window_size=2,
                  sliding_stride=1, bidirectional=False, alternation_rounds=2)
    single = _make(kwargs, 8, OraclePipeline(noise="content"))
    single.execute_tasks()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 400
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, kwargs, 8, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, lat, ti in results:
        torch.testing.assert_close(torch.tensor(lat), single.grid_latents, rtol=1e-5, atol=1e-5)
        assert torch.tensor(ti).equal(single.grid_timestep_indices)


# ---- dataset prefetch and asynchronous save (host logic) ---------------------------------------
def _grid_of(s):
    return s.grid_latents.clone(), s.grid_timestep_indices.clone()


@pytest.mark.parametrize("tag", ["v6_t4_stride1", "v5_t2_stride2_unidir"])
def test_prefetch_and_async_save_change_nothing_but_the_overlap(tag):
    """Same grid, same samples handed to ``save_fn`` in the same order as the sequential run (and therefore as the
    reference sampler: the sequential run is pinned against it above)."""
    c = torch.load(os.path.join(GOLD, "sampler_ref.pt"))["cases"][tag]
    seq, ovl = [], []
    a = _make(c["kwargs"], c["n_cams"], OraclePipeline(), save_fn=lambda sample, out_dir: seq.append(sample))
    a.execute_tasks()
    b = B200SlidingIterativeSampler(dataset=SyntheticSpaTemDataset(c["n_cams"]), pipelines=[OraclePipeline()], output_dir=None,
                                    num_denoising_steps=1, guidance_scale=2.0, sliding_shift=0, prefetch=True, async_save=True,
                                    save_fn=lambda sample, out_dir: ovl.append(sample), **c["kwargs"])
    b.execute_tasks()
    ga, gb = _grid_of(a), _grid_of(b)
    assert torch.equal(ga[0], gb[0]) and torch.equal(ga[1], gb[1])
    assert len(seq) == len(ovl) == sum(len(t) for t in a.all_tasks)
    for x, y in zip(seq, ovl):
        assert (x["alt"], x["domain"], x["domain_label"]) == (y["alt"], y["domain"], y["domain_label"])
        assert torch.equal(x["result_latents"], y["result_latents"])
        assert torch.equal(x["timestep_indices"], y["timestep_indices"])


def test_prefetch_and_async_save_overlap_host_work_with_the_denoise():
    """Dataset load, denoise and save take 60 ms each (sleeps = no GIL held, like PIL decode / CUDA waits / file writes):
    sequentially 3 x 60 ms per task, overlapped about 60 ms per task."""
    import time

    class SlowDataset(SyntheticSpaTemDataset):
        def get_item(self, **kw):
            time.sleep(0.06)
            return super().get_item(**kw)

    class SlowPipeline(OraclePipeline):
        def sliding_iterative_denoise(self, **kw):
            time.sleep(0.06)
            return super().sliding_iterative_denoise(**kw)

    def slow_save(sample, out_dir):
        time.sleep(0.06)

    kwargs = dict(window_size=2, sliding_stride=1, bidirectional=True, alternation_rounds=1, spa_label_range=(0, 6, 1),
                  tem_label_range=(0, 8, 1), input_spa_labels=(1, 4))

    def run(overlap):
        s = B200SlidingIterativeSampler(dataset=SlowDataset(6), pipelines=[SlowPipeline()], output_dir=None,
                                        num_denoising_steps=1, guidance_scale=2.0, sliding_shift=0, save_fn=slow_save,
                                        prefetch=overlap, async_save=overlap, **kwargs)
        t0 = time.perf_counter()
        s.execute_tasks()
        return time.perf_counter() - t0, sum(len(t) for t in s.all_tasks)

    t_seq, n = run(False)
    t_ovl, _ = run(True)
    assert n == 8
    assert t_seq > n * 0.17                      # three 60 ms phases back to back
    assert t_ovl < t_seq - 0.5 * n * 0.06, (t_seq, t_ovl)   # at least half of one phase per task hidden (ideal: two)


def test_async_save_and_prefetch_errors_surface_on_the_caller():
    c = torch.load(os.path.join(GOLD, "sampler_ref.pt"))["cases"]["v6_t4_stride1"]

    def bad_save(sample, out_dir):
        raise OSError("disk full")
    s = B200SlidingIterativeSampler(dataset=SyntheticSpaTemDataset(c["n_cams"]), pipelines=[OraclePipeline()], output_dir=None,
                                    num_denoising_steps=1, guidance_scale=2.0, sliding_shift=0, async_save=True,
                                    save_fn=bad_save, **c["kwargs"])
    with pytest.raises(OSError, match="disk full"):
        s.execute_tasks()

    class BadDataset(SyntheticSpaTemDataset):
        def get_item(self, **kw):
            if kw["tem_labels"] == ["000002"]:
                raise FileNotFoundError("missing frame 000002")
            return super().get_item(**kw)
    s = B200SlidingIterativeSampler(dataset=BadDataset(c["n_cams"]), pipelines=[OraclePipeline()], output_dir=None,
                                    num_denoising_steps=1, guidance_scale=2.0, sliding_shift=0, prefetch=True, **c["kwargs"])
    with pytest.raises(FileNotFoundError, match="000002"):
        s.execute_tasks()


# ---- SURVEY 8f row 1, first step: encoded-image cache across tasks -----------------------------------------------------
class _CountingVAE:
    def __init__(self):
        self.images = 0

    def encode_latents(self, x):
        self.images += len(x)
        z = F.avg_pool2d(x.float(), 8)
        return torch.cat([z, z.mean(dim=1, keepdim=True)], dim=1)


class _VaePipeline(OraclePipeline):
    """OraclePipeline with the ``vae`` / ``device`` attributes and the ``pixel_values_latents`` argument of
    ``B200Diffuman4DPipeline.sliding_iterative_denoise`` (images are rounded to bf16 before the encoder, as there)."""

    def __init__(self):
        super().__init__()
        self.vae, self.device = _CountingVAE(), torch.device("cpu")

    def sliding_iterative_denoise(self, pixel_values=None, pixel_values_latents=None, **kw):
        if pixel_values_latents is None:
            pixel_values_latents = self.vae.encode_latents(pixel_values.to(torch.bfloat16))
        h, w = pixel_values_latents.shape[-2:]
        mask = F.interpolate(kw["cond_masks"], size=(h, w), mode="nearest")
        latents = kw["latents"]
        if latents is None:
            latents = torch.randn((len(pixel_values_latents), 4, h, w), generator=torch.Generator().manual_seed(9000 + self.calls))
            self.calls += 1
        out = sliding_iterative_denoise_oracle(
            self.unet, DDIMOracle(SchedulerConfig()), pixel_latents=pixel_values_latents, plucker=kw["plucker_embeds"],
            skeletons=kw["skeletons"], cond_mask=mask, latents=latents, domain=kw["domain"],
            timestep_indices=kw["timestep_indices"].cpu(), window_size=kw["window_size"], sliding_stride=kw["sliding_stride"],
            sliding_shift=kw["sliding_shift"], bidirectional=kw["bidirectional"], num_denoising_steps=kw["num_denoising_steps"],
            alternation_rounds=kw["alternation_rounds"], guidance_scale=kw["guidance_scale"], enable_pose_encoder=True)
        out["images"] = out["latents"]
        return out


def test_pixel_latent_cache_encodes_every_grid_cell_once():
    c = torch.load(os.path.join(GOLD, "sampler_ref.pt"))["cases"]["v6_t4_stride1"]

    def run(cache):
        pipe = _VaePipeline()
        s = B200SlidingIterativeSampler(dataset=SyntheticSpaTemDataset(c["n_cams"]), pipelines=[pipe], output_dir=None,
                                        num_denoising_steps=1, guidance_scale=2.0, sliding_shift=0, cache_pixel_latents=cache,
                                        **c["kwargs"])
        s.execute_tasks()
        return s, pipe
    a, pa = run(False)
    b, pb = run(True)
    assert torch.equal(a.grid_latents, b.grid_latents) and torch.equal(a.grid_timestep_indices, b.grid_timestep_indices)
    per_task = sum(len(a._fetch(**t)["labels"]) for tasks in a.all_tasks for t in tasks)
    cells = len(a.spa_labels) * len(a.tem_labels)
    assert pa.vae.images == per_task                       # the reference's behaviour: every frame of every task
    assert pb.vae.images == b.vae_images_encoded == cells  # cached: every image of the grid once
    assert cells < per_task
