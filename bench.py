#!/usr/bin/env python
"""bench.py -- Diffuman4D window denoise-steps/sec on B200 (BASELINE.json metric).

A "step" is ONE window denoise step of the reference's sliding-window sampler
(pipeline_diffuman4d.py:369-425): input assembly -> UNet forward on 2F images (CFG) -> CFG combine -> F per-frame
DDIM updates.  Workload at N=1: the spatial window W16 of `demo_4d_tiny` (4 cond + 12 target frames, CFG => 32
images, SD-2.1 layout UNet, latents 64x64 -- BASELINE's synthetic size).  Synthetic seeded inputs, random-init
weights of the real architecture (no network for the checkpoint).

    python bench.py --gpus 1 --steps 10 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...   (weak scaling: every
        rank denoises its own window; windows of one sampler round are independent units)
    python bench.py --impl reference ...   (the CPU oracle -- the reference's torch graph restated -- on host cores)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(name="demo_4d_tiny spatial window W16 (4 cond + 12 target frames), CFG 2.0, latents 64x64",
                F=16, n_cond=4, h=64, w=64, guidance=2.0, domain="spatial", n_steps=18)
# bounded samples of the same window step for the CPU arm (largest that fits the time budget is used)
CPU_SAMPLES = [dict(F=4, n_cond=1, h=64, w=64), dict(F=2, n_cond=1, h=64, w=64), dict(F=2, n_cond=1, h=32, w=32)]
METRIC = "unet_window_denoise_steps_per_sec"
UNIT = "window-steps/s"


def synth_inputs(F, n_cond, h, w, pose=True, seed=0):
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    d = dict(latents=r(F, 4, h, w), pixel=r(F, 4, h, w), plucker=(torch.rand(F, 6, h, w, generator=g) * 2 - 1),
             mask=torch.ones(F, 1, h, w))
    d["mask"][:n_cond] = 0
    d["skel"] = (torch.rand(F, 3, 8 * h, 8 * w, generator=g) * 2 - 1) if pose else r(F, 4, h, w)
    ti = torch.zeros(F, dtype=torch.int64)
    tgt = F - n_cond
    ti[n_cond:] = torch.tensor([min(17, (tgt - 1 - i) // 2) for i in range(tgt)])   # staggered like PIPE:503-543
    d["ts"] = ti
    return d


class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                p = [x.strip() for x in o.strip().split(",")]
                if len(p) >= 6:
                    self.samples.append(p)
            except Exception:  # noqa: BLE001
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=3)
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------------- CPU arm
def calibrate_cpu_threads():
    """Pick the torch thread count that maximises fp32 conv throughput on this host (a cgroup-limited box can be much
    slower with one thread per visible core).  Returns (threads, conv GFLOP/s)."""
    import torch.nn.functional as Fn
    try:
        usable = len(os.sched_getaffinity(0))
    except AttributeError:
        usable = os.cpu_count() or 1
    x = torch.randn(4, 320, 64, 64)
    wgt = torch.randn(320, 320, 3, 3)
    fl = 2.0 * 4 * 64 * 64 * 320 * 320 * 9
    best = (1, 0.0)
    for nt in sorted({usable, max(1, usable // 2), max(1, usable // 4), min(usable, 16), min(usable, 8)}):
        torch.set_num_threads(nt)
        Fn.conv2d(x, wgt, padding=1)
        t0 = time.perf_counter()
        for _ in range(2):
            Fn.conv2d(x, wgt, padding=1)
        g = 2 * fl / (time.perf_counter() - t0) / 1e9
        if g > best[1]:
            best = (nt, g)
    torch.set_num_threads(best[0])
    return best


def cpu_oracle_window_seconds(repeats: int, warmup: int, budget_s: float):
    """Times the oracle (test infrastructure, used here ONLY as the reported CPU baseline) on a bounded sample of the
    window step chosen so that (warmup + repeats) steps fit in ``budget_s``."""
    from diffuman4d_b200.config import SchedulerConfig, UNetConfig
    from diffuman4d_b200.flops import unet_flops
    from diffuman4d_b200.weights import random_state_dict
    from oracle.pipeline_oracle import DDIMOracle, denoise_window_oracle
    from oracle.unet_oracle import OracleUNet
    threads, conv_gflops = calibrate_cpu_threads()
    cfg = UNetConfig.sd21()
    per_step = budget_s / max(1, warmup + repeats)
    s = CPU_SAMPLES[-1]
    for cand in CPU_SAMPLES:
        est = unet_flops(cfg, 2 * cand["F"], cand["F"], cand["h"], cand["w"])["total"] / (0.6 * conv_gflops * 1e9)
        if est <= per_step:
            s = cand
            break
    net = OracleUNet(cfg).eval()
    net.load_state_dict({k: v.float() for k, v in random_state_dict(cfg, seed=1).items()})
    d = synth_inputs(s["F"], s["n_cond"], s["h"], s["w"])
    sched = DDIMOracle(SchedulerConfig())
    sched.set_timesteps(WORKLOAD["n_steps"])

    def unet(x, t, sk, doms, nf):
        with torch.no_grad():
            return net(x, t, sk, doms, nf)

    times = []
    for i in range(warmup + repeats):
        t0 = time.perf_counter()
        denoise_window_oracle(unet, sched, latents=d["latents"].clone(), pixel_latents=d["pixel"], plucker=d["plucker"],
                              skeletons=d["skel"], cond_mask=d["mask"], timestep_indices=d["ts"], domain="spatial",
                              guidance_scale=WORKLOAD["guidance"])
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    fl_s = unet_flops(cfg, 2 * s["F"], s["F"], s["h"], s["w"])["total"]
    fl_w = unet_flops(cfg, 2 * WORKLOAD["F"], WORKLOAD["F"], WORKLOAD["h"], WORKLOAD["w"])["total"]
    return times, fl_s, fl_w, threads, s


def cpu_baseline_obj(times, fl_s, fl_w, threads, s):
    sec = sum(times) / len(times)
    return {"value": (fl_s / sec) / fl_w, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"oracle fp32, one window step with F={s['F']} frames ({s['n_cond']} cond) CFG at "
                       f"{s['h']}x{s['w']} latents = {fl_s / 1e12:.2f} TFLOP in {sec:.2f} s ({threads} torch threads, calibrated); "
                       f"scaled to the W16 step ({fl_w / 1e12:.2f} TFLOP) by FLOP ratio"),
            "gflops": fl_s / sec / 1e9}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t0 = time.perf_counter()
    times, fl_s, fl_w, threads, smp = cpu_oracle_window_seconds(args.steps, args.warmup, budget_s=150.0)
    cb = cpu_baseline_obj(times, fl_s, fl_w, threads, smp)
    val = cb["value"]
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD["name"], "note": "CPU torch restatement of the reference graph (diffusers is not installable here)"},
        "cpu_baseline": cb, "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": time.perf_counter() - t0}))


# ------------------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch.distributed as dist
    from diffuman4d_b200 import build as d4d_build
    from diffuman4d_b200._lib import check, lib
    from diffuman4d_b200.config import SchedulerConfig, UNetConfig
    from diffuman4d_b200.flops import unet_flops
    from diffuman4d_b200.pipeline import B200Diffuman4DPipeline
    from diffuman4d_b200.unet import B200MultiviewUNet
    from diffuman4d_b200.weights import random_state_dict

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.path.exists(os.path.join(ROOT, "diffuman4d_b200", "libd4d.so")):
        d4d_build.build()
    dev = torch.device("cuda", local)
    cfg = UNetConfig.sd21()
    wl = WORKLOAD
    F, h, w = wl["F"], wl["h"], wl["w"]
    unet = B200MultiviewUNet(cfg, local).load_state_dict(random_state_dict(cfg, seed=1))
    pipe = B200Diffuman4DPipeline(unet, SchedulerConfig())
    pipe.parepare_schedulers(wl["n_steps"], F)
    sharded = args.mode == "sharded" and world > 1
    full_inputs = synth_inputs(F, wl["n_cond"], h, w, seed=0 if sharded else rank)
    F_total = F
    sh = None
    if sharded:   # one window, frames split over the ranks, fused K/V exchange over peer memory
        from diffuman4d_b200.sharded import FrameShardedPipeline
        sh = FrameShardedPipeline(pipe, max_frames=F_total, h=h, w=w)
        lo, hi = sh.frames(F_total)
        full_inputs = {k: v[lo:hi].contiguous() for k, v in full_inputs.items()}
        F = hi - lo
    host = {k: (v.to(torch.bfloat16) if v.dtype.is_floating_point else v).pin_memory()
            for k, v in full_inputs.items()}
    devt = {k: v.to(dev) for k, v in host.items()}
    lat, ts = devt["latents"].clone(), devt["ts"].clone()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def window_step(latents, pixel, plucker, skel, mask, tsi):
        if sh is not None:
            sh.denoise_window(latents=latents, pixel_values_latents=pixel, plucker_embeds_latents=plucker,
                              skeletons_latents=skel, cond_masks_latents=mask, timestep_indices=tsi, domain=wl["domain"],
                              guidance_scale=wl["guidance"], F_total=F_total)
        else:
            pipe.denoise_window(latents=latents, pixel_values_latents=pixel, plucker_embeds_latents=plucker,
                                skeletons_latents=skel, cond_masks_latents=mask, timestep_indices=tsi,
                                domain=wl["domain"], guidance_scale=wl["guidance"])

    def step_resident():
        lat.copy_(devt["latents"])
        ts.copy_(devt["ts"])
        window_step(lat, devt["pixel"], devt["plucker"], devt["skel"], devt["mask"], ts)

    out_host = torch.empty_like(host["latents"]).pin_memory()
    ts_host = torch.empty_like(host["ts"]).pin_memory()
    # e2e = the public call fed from pinned HOST buffers: every step uploads its own inputs (H2D) and reads its result
    # back (D2H).  The upload of step i+1 is issued on a copy stream while step i computes (double-buffered staging), as
    # a caller streaming windows through the pipeline would do; each step still ends with a stream synchronisation.
    stages = [{k: torch.empty_like(v) for k, v in devt.items()} for _ in range(2)]
    copy_stream = torch.cuda.Stream(device=dev)
    h2d_done = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"i": 0, "primed": False}
    KEYS = ("latents", "pixel", "plucker", "skel", "mask", "ts")

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            for k in KEYS:
                stages[slot][k].copy_(host[k], non_blocking=True)
            h2d_done[slot].record(copy_stream)

    def step_e2e():
        i = e2e_state["i"]
        cur = i & 1
        if not e2e_state["primed"]:
            upload(cur)
            e2e_state["primed"] = True
        upload(cur ^ 1)                                   # inputs of the NEXT step, overlapped with this step's compute
        torch.cuda.current_stream().wait_event(h2d_done[cur])
        st = stages[cur]
        window_step(st["latents"], st["pixel"], st["plucker"], st["skel"], st["mask"], st["ts"])
        out_host.copy_(st["latents"], non_blocking=True)
        ts_host.copy_(st["ts"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        e2e_state["i"] = i + 1

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        barrier()
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    sampler = ClockSampler(local)
    sampler.start()
    ms_res = timed(step_resident, args.steps, max(args.warmup, 3))
    clocks = sampler.stop()
    ms_e2e = timed(step_e2e, args.steps, 1)
    assert torch.isfinite(out_host.float()).all(), "non-finite latents out of the window step"

    # ---- live per-kernel-kind device times of the UNet forward (CUDA events around every launch) ----
    F = F_total        # the profile below always runs the full single-GPU window
    B = 2 * F
    x = torch.randn(B, cfg.in_channels, h, w, device=dev).to(torch.bfloat16)
    tt = torch.randint(0, 1000, (B,), device=dev)
    skp = torch.rand(F, 3, 8 * h, 8 * w, device=dev).to(torch.bfloat16) * 2 - 1
    sk = torch.cat([-torch.ones_like(skp), skp])
    y = torch.empty(B, 4, h, w, device=dev, dtype=torch.bfloat16)
    doms = (C.c_int32 * 2)(0, 0)
    ms_k, n_k, fl_k = (C.c_float * 6)(), (C.c_int32 * 6)(), (C.c_double * 6)()
    acc = [0.0] * 6
    reps = 3
    for i in range(reps + 1):
        check(lib().d4d_profile_forward(unet._h, x.data_ptr(), tt.data_ptr(), sk.data_ptr(), doms, 2, B, F, h, w,
                                        y.data_ptr(), torch.cuda.current_stream().cuda_stream, ms_k, n_k, fl_k))
        if i > 0:
            acc = [a + m for a, m in zip(acc, ms_k)]
    kind_ms = [a / reps for a in acc]
    fl = unet_flops(cfg, B, F, h, w)
    alg = {0: fl["linear"] + fl["ff"], 1: fl["conv3x3"], 2: fl["attn3d"] + fl["attn2d"]}
    names = ["gemm", "conv3x3", "attention", "groupnorm", "layernorm", "other"]
    top = max((0, 1, 2), key=lambda k: kind_ms[k])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved = alg[top] / (kind_ms[top] * 1e-3) / 1e12
    launches_fwd = unet.forward_launches(2, B, F, h, w)
    launches_step = launches_fwd + 3   # + assemble, cfg-skeleton, cfg+ddim kernels (2 memcpys not counted)
    roofline = {"bound": "tensor", "kernel": names[top], "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf,
                # dram__bytes_read.sum + dram__bytes_write.sum of the captured attention launch (level-2 3-D layer, grid 1280,
                # profiles/r01d_ncu_attention.txt); its algorithmic Q/K/V/O bytes are 67.1 MB => no wasted HBM re-reads
                "traffic": 68148480 if names[top] == "attention" else None,
                "traffic_note": "ncu --set full capture of one level-2 3-D attention launch: 68.1 MB DRAM vs 67.1 MB algorithmic",
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained",
                "launches_per_forward": int(n_k[top]), "ms_per_forward": kind_ms[top],
                "by_kind_ms": {names[k]: round(kind_ms[k], 4) for k in range(6)},
                "by_kind_tflops": {names[k]: round(alg[k] / (kind_ms[k] * 1e-3) / 1e12, 1) for k in (0, 1, 2)},
                "unet_forward_ms_sum": round(sum(kind_ms), 3)}

    jobs = 1 if sharded else world   # sharded: all ranks cooperate on ONE window per step (strong scaling)
    value = jobs * args.steps / (ms_res * 1e-3)
    e2e_val = jobs * args.steps / (ms_e2e * 1e-3)
    h2d = sum(host[k].numel() * host[k].element_size() for k in ("latents", "pixel", "plucker", "skel", "mask", "ts"))
    d2h = out_host.numel() * 2 + ts_host.numel() * 8
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "strong" if sharded else "weak",
        "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": wl["name"], "unet": "SD-2.1 layout 320/640/1280/1280, heads 5/10/20/20, in_channels 11 "
                   "(pose encoder + frame-index embedding), no attn2", "images_per_step": B,
                   "parallelism": (f"frame-sharded window x{world} (fused K/V exchange over NVLink peer memory)" if sharded else
                                   f"replicas x{world} (independent windows, no collective)" if world > 1 else "single GPU"),
                   "l2": f"no explicit flush: one step streams ~{unet.workspace_bytes(2, B, F, h, w) / 2**30:.1f} GiB of "
                         "activations + 1.6 GB of weights through a 126 MB L2"},
        "tflops_per_step": fl["total"] / 1e12,
        "unet_tflops_achieved": jobs * fl["total"] / (ms_res / args.steps * 1e-3) / 1e12,
        "unet_roofline_frac": jobs * fl["total"] / (ms_res / args.steps * 1e-3) / 1e12 / peak_tf / world,
        "roofline": roofline, "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches_step * args.steps,
    }
    if world == 1 and not args.no_cpu_baseline:
        times, fl_s, fl_w, threads, smp = cpu_oracle_window_seconds(1, 0, budget_s=25.0)
        line["cpu_baseline"] = cpu_baseline_obj(times, fl_s, fl_w, threads, smp)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="replicas", choices=["replicas", "sharded"],
                    help="N>1: independent windows per GPU (weak) or one frame-sharded window over all GPUs (strong)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
