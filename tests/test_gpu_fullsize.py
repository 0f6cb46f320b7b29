"""GPU parity tests AT THE BENCHMARKED SIZE: the CUDA path (through the C ABI) against the fp32 CPU oracle for the SD-2.1
channel layout (320/640/1280/1280) -- the configuration bench.py times -- instead of the tiny-channel configs of
tests/test_gpu_unet.py.  Covered: UNet forward at F=4 and F=16 (the W16 spatial window, 32 images @ 64x64 latents), F=24
temporal (W24, 48 images), one full window denoise step at W16, the reference constructor's head layout (8 heads per level
=> head_dim 40/80/160, padded to 64/128/192 in the kernels) at full channel widths, and one case at the reference's default
128x128 latents.  These exercise what the small configs cannot: K = 11 520 / 23 040 convs, two-source K 2560+1280
shortcuts, N = 10 240 GEGLU, block_n 256 tiles, 16 384-token 3-D attention and the 20 160-wide time-embedding GEMM.

Criteria (floating point; BASELINE's rtol 1e-3 is below bf16 resolution, see tests/test_gpu_unet.py):
  global       e_ours <= 1.5 * e_eager + 2e-3,  e = max|y - oracle_fp32| / max|oracle_fp32|, e_eager from the SAME torch graph
               run in bf16 on the GPU (what the reference executes)
  elementwise  |y - oracle| <= 3 * e_eager * max|oracle| + 2e-2 * |oracle|   for EVERY element (a single bad channel or tile
               cannot hide under the global maximum)
and a per-level drift table (max / rms error of every block output, ours next to bf16 eager) is printed.
"""
import pytest
import torch

from diffuman4d_b200.config import SchedulerConfig, UNetConfig
from diffuman4d_b200.weights import random_state_dict

pytestmark = pytest.mark.gpu


def _inputs(cfg, F, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    B = 2 * F
    x = torch.randn(B, cfg.in_channels, h, w, generator=g).to(torch.bfloat16)
    t = torch.randint(0, 1000, (B,), generator=g)
    sk = (torch.rand(B, 3, 8 * h, 8 * w, generator=g) * 2 - 1).to(torch.bfloat16) if cfg.enable_pose_encoder else None
    return x, t, sk


def _build(cfg, seed=1):
    from diffuman4d_b200.unet import B200MultiviewUNet
    from oracle.unet_oracle import OracleUNet
    sd = random_state_dict(cfg, seed=seed, dtype=torch.bfloat16)
    ours = B200MultiviewUNet(cfg, device=0).load_state_dict(sd)
    ref = OracleUNet(cfg).eval()
    ref.load_state_dict({k: v.float() for k, v in sd.items()})
    return ours, ref, sd


def _eager_bf16(cfg, sd):
    from oracle.unet_oracle import OracleUNet
    m = OracleUNet(cfg).eval()
    m.load_state_dict({k: v.float() for k, v in sd.items()})
    return m.to("cuda").to(torch.bfloat16)


class _Taps:
    """Forward hooks on the oracle that record the block outputs the product exposes through d4d_debug_tap."""

    def __init__(self, net):
        self.out, self._h, self._parts = {}, [], {}
        grab = lambda name: (lambda m, i, o: self.out.__setitem__(name, (o[0] if isinstance(o, tuple) else o).detach().float().cpu()))
        for i, b in enumerate(net.down_blocks):
            self._h.append(b.register_forward_hook(grab(f"down_blocks.{i}")))
        self._h.append(net.mid_block.register_forward_hook(grab("mid_block")))
        for i, b in enumerate(net.up_blocks):
            self._h.append(b.register_forward_hook(grab(f"up_blocks.{i}")))
        part = lambda name: (lambda m, i, o: self._parts.__setitem__(name, o.detach().float().cpu()))
        self._h.append(net.conv_in.register_forward_hook(part("conv_in")))
        if hasattr(net, "pose_encoder"):
            self._h.append(net.pose_encoder.register_forward_hook(part("pose")))

    def close(self):
        for h in self._h:
            h.remove()
        if "conv_in" in self._parts:
            self.out["conv_in"] = self._parts["conv_in"] + self._parts.get("pose", 0.0)
        return self.out


def _drift_table(title, ours_taps, ref_taps, eager_taps):
    print(f"\n  per-level drift [{title}]   (max|err| / max|ref|, rms err / rms ref)")
    print(f"  {'block':<16}{'ours max':>11}{'ours rms':>11}{'eager max':>11}{'eager rms':>11}")
    worst = 0.0
    for name, r in ref_taps.items():
        if name not in ours_taps:
            continue
        def stats(y):
            d = y.float().cpu() - r
            return d.abs().max().item() / r.abs().max().item(), d.pow(2).mean().sqrt().item() / r.pow(2).mean().sqrt().item()
        om, orr = stats(ours_taps[name])
        em, er = stats(eager_taps[name]) if name in eager_taps else (float("nan"), float("nan"))
        worst = max(worst, orr / max(er, 1e-9))
        print(f"  {name:<16}{om:>11.3e}{orr:>11.3e}{em:>11.3e}{er:>11.3e}")
    return worst


def _check(title, y, y_ref, y_eager):
    y, y_ref, y_eager = y.float().cpu(), y_ref.float().cpu(), y_eager.float().cpu()
    assert y.shape == y_ref.shape and torch.isfinite(y).all()
    scale = y_ref.abs().max().item()
    err, err_e = (y - y_ref).abs(), (y_eager - y_ref).abs()
    e_ours, e_eager = err.max().item() / scale, err_e.max().item() / scale
    rms = lambda d: d.pow(2).mean().sqrt().item() / y_ref.pow(2).mean().sqrt().item()
    bound = 3.0 * e_eager * scale + 2e-2 * y_ref.abs()
    bad = int((err > bound).sum())
    print(f"\n[{title}] e_ours={e_ours:.3e} e_eager_bf16={e_eager:.3e}  rms: ours={rms(err):.3e} eager={rms(err_e):.3e}  "
          f"elementwise violations={bad}/{err.numel()}")
    assert e_ours <= 1.5 * e_eager + 2e-3, (e_ours, e_eager)
    assert bad == 0, f"{bad} elements exceed 3*e_eager*max + 2e-2*|ref| (worst {(err - bound).max().item():.3e} over)"
    return e_ours, e_eager


def _unet_case(cfg, F, h, w, domain, title, drift=True):
    ours, ref, sd = _build(cfg)
    x, t, sk = _inputs(cfg, F, h, w)
    doms = [domain, domain]
    taps = _Taps(ref) if drift else None
    with torch.no_grad():
        y_ref = ref(x.float(), t, None if sk is None else sk.float(), doms, F)          # fp32 oracle on the host cores
    ref_taps = taps.close() if drift else {}
    ref16 = _eager_bf16(cfg, sd)
    taps16 = _Taps(ref16) if drift else None
    with torch.no_grad():
        y_eager = ref16(x.cuda(), t.cuda(), None if sk is None else sk.cuda(), doms, F)
    eager_taps = taps16.close() if drift else {}
    del ref16
    torch.cuda.empty_cache()
    y = ours(x.cuda(), t.cuda(), None if sk is None else sk.cuda(), doms, F, return_dict=False)[0]
    torch.cuda.synchronize()
    if drift:
        ours_taps = ours.debug_taps(x.cuda(), t.cuda(), None if sk is None else sk.cuda(), doms, F)
        torch.cuda.synchronize()
        assert set(ours_taps) == set(ref_taps), (sorted(ours_taps), sorted(ref_taps))
        for k in ours_taps:
            assert ours_taps[k].shape == ref_taps[k].shape, (k, ours_taps[k].shape, ref_taps[k].shape)
        worst = _drift_table(title, ours_taps, ref_taps, eager_taps)
        assert worst <= 2.0, f"a block drifts {worst:.2f}x further (rms) than bf16 eager"
        # the taps must not disturb the plan: a plain forward afterwards reproduces the result bit for bit
        y2 = ours(x.cuda(), t.cuda(), None if sk is None else sk.cuda(), doms, F, return_dict=False)[0]
        assert torch.equal(y, y2)
    _check(title, y, y_ref, y_eager)


def test_sd21_unet_f4_64(cuda):
    _unet_case(UNetConfig.sd21(), 4, 64, 64, "spatial", "sd21 F=4 64x64 spatial")


def test_sd21_unet_w24_temporal_64(cuda):
    """W24: 12 cond + 12 target frames, CFG => 48 images, frame-index embedding arange(12).repeat(2), 24 576-token 3-D attention."""
    _unet_case(UNetConfig.sd21(), 24, 64, 64, "temporal", "sd21 W24 64x64 temporal", drift=False)


def test_ctor_default_heads_full_width(cuda):
    """Reference constructor defaults: 8 heads per level => head_dim 40/80/160 (zero-padded to 64/128/192 in the fused QKV /
    to_out weights), 1x1-conv projections, skeleton latents concatenated (in_channels 15), no frame embedding."""
    _unet_case(UNetConfig.ctor_default(), 4, 32, 32, "spatial", "ctor_default heads 8/8/8/8 F=4 32x32")


def test_sd21_unet_f2_128(cuda):
    """Reference default latent size (1024^2 px => 128x128): level-0 2-D attention over 16 384 tokens, level-1 3-D over 8 192."""
    _unet_case(UNetConfig.sd21(), 2, 128, 128, "temporal", "sd21 F=2 128x128 temporal", drift=False)


def test_sd21_w16_unet_and_window_step(cuda):
    """THE benchmarked configuration (bench.py WORKLOAD): spatial window W16 = 4 cond + 12 target frames, CFG 2.0, 64x64
    latents.  One fp32 oracle window step on the host cores yields both references: the UNet's noise prediction on the
    assembled 32-image batch (compared with d4d_unet_forward on the same batch) and the updated latents / timestep indices
    (compared with ONE d4d_denoise_window call)."""
    from diffuman4d_b200.pipeline import B200Diffuman4DPipeline
    from oracle.pipeline_oracle import DDIMOracle, denoise_window_oracle
    cfg = UNetConfig.sd21()
    ours, ref, sd = _build(cfg)
    F, n_cond, h, w = 16, 4, 64, 64
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g).to(torch.bfloat16)
    lat, pix = r(F, 4, h, w), r(F, 4, h, w)
    plk = (torch.rand(F, 6, h, w, generator=g) * 2 - 1).to(torch.bfloat16)
    skel = (torch.rand(F, 3, 8 * h, 8 * w, generator=g) * 2 - 1).to(torch.bfloat16)
    mask = torch.ones(F, 1, h, w, dtype=torch.bfloat16)
    mask[:n_cond] = 0
    ti = torch.zeros(F, dtype=torch.int64)
    ti[n_cond:] = torch.tensor([min(17, (F - n_cond - 1 - i) // 2) for i in range(F - n_cond)])   # staggered like PIPE:503-543
    sc = SchedulerConfig()
    orc = DDIMOracle(sc)
    orc.set_timesteps(18)
    seen = {}
    taps = _Taps(ref)

    def unet_fp32(x, t, sk, doms, nf):
        with torch.no_grad():
            y = ref(x.float(), t, sk.float(), doms, nf)
        seen.update(x=x.clone(), t=t.clone(), sk=sk.clone(), doms=list(doms), nf=nf, y=y)
        return y

    ref_lat, ref_ti = denoise_window_oracle(unet_fp32, orc, latents=lat.float(), pixel_latents=pix.float(), plucker=plk.float(),
                                            skeletons=skel.float(), cond_mask=mask.float(), timestep_indices=ti,
                                            domain="spatial", guidance_scale=2.0)
    ref_taps = taps.close()
    # the same graph in bf16 on the GPU (what the reference runs) for the error scale
    ref16 = _eager_bf16(cfg, sd)
    taps16 = _Taps(ref16)

    def unet_bf16(x, t, sk, doms, nf):
        with torch.no_grad():
            return ref16(x.cuda().to(torch.bfloat16), t.cuda(), sk.cuda().to(torch.bfloat16), doms, nf).float().cpu()

    seen16 = {}
    eag_lat, _ = denoise_window_oracle(lambda *a: seen16.setdefault("y", unet_bf16(*a)), orc, latents=lat.float(),
                                       pixel_latents=pix.float(), plucker=plk.float(), skeletons=skel.float(),
                                       cond_mask=mask.float(), timestep_indices=ti, domain="spatial", guidance_scale=2.0)
    eager_taps = taps16.close()
    del ref16
    torch.cuda.empty_cache()

    # ---- (1) UNet forward on the assembled batch ----
    xb, tb, skb = seen["x"].to(torch.bfloat16).cuda(), seen["t"].cuda(), seen["sk"].to(torch.bfloat16).cuda()
    assert xb.shape == (2 * F, cfg.in_channels, h, w) and seen["nf"] == F
    y = ours(xb, tb, skb, seen["doms"], F, return_dict=False)[0]
    ours_taps = ours.debug_taps(xb, tb, skb, seen["doms"], F)
    torch.cuda.synchronize()
    worst = _drift_table("sd21 W16 64x64 spatial (bench workload)", ours_taps, ref_taps, eager_taps)
    assert worst <= 2.0
    _check("sd21 W16 UNet (bench workload)", y, seen["y"], seen16["y"])

    # ---- (2) one window step through the C ABI ----
    pipe = B200Diffuman4DPipeline(ours, sc)
    pipe.parepare_schedulers(18, F)
    l_d, t_d = lat.clone().cuda(), ti.clone().cuda()
    pipe.denoise_window(latents=l_d, pixel_values_latents=pix, plucker_embeds_latents=plk, skeletons_latents=skel,
                        cond_masks_latents=mask, timestep_indices=t_d, domain="spatial", guidance_scale=2.0)
    torch.cuda.synchronize()
    assert torch.equal(t_d.cpu(), ref_ti)
    assert torch.equal(l_d.cpu()[:n_cond], pix[:n_cond])                    # cond frames hold the image latents (PIPE:375-379)
    _check("sd21 W16 window step latents", l_d[n_cond:], ref_lat[n_cond:], eag_lat[n_cond:])
