"""Stand-in for the unpublished Swin fundus encoder: ``build_model()`` returns a module mapping
``[B, 3, 384, 384] -> (tokens [B, 144, 1024], pooled [B, 1024])`` (contract read off code/fusion_net.py:884,898 and
the EPRL/DILR input widths).  A 32 x 32 patch embedding + LayerNorm + one MLP block, run in bf16 under autocast on CUDA
(BASELINE configs[2]: "bf16 encoders"); the hot-path modules receive fp32 tokens."""
import torch
import torch.nn as nn


class PatchEncoder2D(nn.Module):
    def __init__(self, dim=1024, patch=32):
        super().__init__()
        self.embed = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)
        self.norm = nn.LayerNorm(dim)
        self.mlp = nn.Sequential(nn.Linear(dim, dim), nn.GELU(), nn.Linear(dim, dim))

    def forward(self, x):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=x.is_cuda):
            t = self.embed(x).flatten(2).transpose(1, 2)            # [B, 144, dim]
            t = self.norm(t)
            t = t + self.mlp(t)
        t = t.float()
        return t, t.mean(dim=1)


def build_model(*args, **kwargs):
    return PatchEncoder2D()
