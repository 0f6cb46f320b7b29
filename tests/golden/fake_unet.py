"""A cheap deterministic stand-in for the UNet, used ONLY to pin the pipeline-level logic (input assembly, CFG, per-frame
scheduler steps, window schedule) of the oracle against the reference's own ``Diffuman4DPipeline`` code: the generator
script runs the reference pipeline with this function as ``pipeline.unet`` and the CPU tests run the oracle with the
same function.  Its output depends on every argument the reference passes (all input channels, the per-image timestep,
the skeletons, the domain list, num_frames and - through the cross-frame mean - the frame grouping of each CFG half),
so an assembly or ordering mistake changes the result.
"""
import torch


def make_fake_unet(in_channels: int, seed: int = 123):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(4, in_channels, generator=g) * 0.5
    b = torch.randn(4, generator=g) * 0.1

    def fake_unet(sample, timestep, skeletons=None, domains=None, num_frames=1):
        x = sample.float()
        B = x.shape[0]
        y = torch.tanh(torch.einsum("oc,bchw->bohw", w, x) + b.view(1, 4, 1, 1))
        y = y + 0.3 * torch.sin(timestep.float().view(B, 1, 1, 1) / 200.0)
        if skeletons is not None:
            sk = skeletons.float().reshape(B, -1).mean(dim=1)
            y = y + 0.2 * sk.view(B, 1, 1, 1)
        # 3-D coupling: every image sees the mean of its group of num_frames images (one group per CFG half / domain entry)
        groups = B // num_frames
        assert groups == len(domains), (B, num_frames, domains)
        yg = y.view(groups, num_frames, *y.shape[1:])
        frame_pos = torch.arange(num_frames, dtype=torch.float32).view(1, num_frames, 1, 1, 1)
        for gi, d in enumerate(domains):
            if d not in ("spatial", "temporal"):
                raise ValueError(d)
        dom = torch.tensor([0.05 if d == "spatial" else -0.05 for d in domains]).view(groups, 1, 1, 1, 1)
        yg = yg + 0.25 * yg.mean(dim=1, keepdim=True) + dom + 0.01 * frame_pos
        return yg.reshape(B, *y.shape[1:]).to(sample.dtype)

    return fake_unet
