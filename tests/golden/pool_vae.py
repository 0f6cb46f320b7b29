"""8x average-pooling stand-in for the stock AutoencoderKL (SURVEY 8f row 1 is outside the CUDA hot path): the object
contract ``B200Diffuman4DPipeline(vae=...)`` expects -- ``encode_latents(images)`` / ``decode_latents(latents)`` -- and a
factory with the ``vae_factory(model_dir, gpu_id)`` signature of ``diffuman4d_b200.loader.load_pipelines`` so tests can name it
by dotted path like a Hydra yaml would."""
import torch
import torch.nn.functional as F


class PoolVAE:
    def encode_latents(self, x):
        z = F.avg_pool2d(x.float(), 8)
        return torch.cat([z, z.mean(dim=1, keepdim=True)], dim=1).to(torch.bfloat16)

    def decode_latents(self, latents):
        return latents[:, :3].float().repeat_interleave(8, 2).repeat_interleave(8, 3).clamp(-1, 1) / 2 + 0.5


def make(model_dir, gpu_id):
    return PoolVAE()
