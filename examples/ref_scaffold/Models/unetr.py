"""Stand-in for the unpublished UNETR OCT encoder: ``UNETR_base_3DNet(num_classes=...)`` returns a module mapping
``[B, 1, 96, 96, 96] -> (tokens [B, 216, 768], pooled [B, 768])`` (contract read off code/fusion_net.py:885,899).
A 16^3 patch embedding + LayerNorm + one MLP block, bf16 under autocast on CUDA."""
import torch
import torch.nn as nn


class PatchEncoder3D(nn.Module):
    def __init__(self, dim=768, patch=16):
        super().__init__()
        self.embed = nn.Conv3d(1, dim, kernel_size=patch, stride=patch)
        self.norm = nn.LayerNorm(dim)
        self.mlp = nn.Sequential(nn.Linear(dim, dim), nn.GELU(), nn.Linear(dim, dim))

    def forward(self, x):
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=x.is_cuda):
            t = self.embed(x).flatten(2).transpose(1, 2)            # [B, 216, dim]
            t = self.norm(t)
            t = t + self.mlp(t)
        t = t.float()
        return t, t.mean(dim=1)


def UNETR_base_3DNet(num_classes=2, **kwargs):
    return PatchEncoder3D()
