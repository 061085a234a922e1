"""A synthetic two-view EDRL training step on one GPU (BASELINE configs[2] shape: batch 64, bf16 encoders) with the
hot path swapped between this package's kernels and the reference's torch op sequence.

The published `MedFusion` cannot run (unpublished encoders, two runtime bugs -- SURVEY.md F3/F6), so the caller
here is a STAND-IN with the same data contract, not a re-implementation: random-init patch-embedding encoders
producing `[B,144,1024]` (fundus, from `[B,3,384,384]`) and `[B,216,768]` (OCT, from `[B,1,96,96,96]`) tokens under
bf16 autocast, two Essence-Point modules, a linear head emitting `combined_features [B,3072]`, label-smoothed CE, and
`MK_MMD` between the clean and the noisy view (code/fusion_train.py:176-224).  What is measured is the step time and
how much of it the hot path is; everything outside the hot path is identical in both arms.

    python examples/edrl_step_synthetic.py [--batch 64] [--steps 10] [--noise device|reference]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch
import torch.nn as nn
import torch.nn.functional as F

import edrl_b200


class StandInEncoders(nn.Module):
    """Patch embeddings with the reference encoders' output contract."""

    def __init__(self):
        super().__init__()
        self.fundus = nn.Conv2d(3, 1024, kernel_size=32, stride=32)        # 384/32 = 12 -> 144 tokens
        self.oct = nn.Conv3d(1, 768, kernel_size=16, stride=16)            # 96/16 = 6 -> 216 tokens

    def forward(self, fundus, oct_):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            a = self.fundus(fundus).flatten(2).transpose(1, 2)             # [B,144,1024]
            b = self.oct(oct_).flatten(2).transpose(1, 2)                  # [B,216,768]
        return a.float(), b.float()


class TorchEPRL(edrl_b200.EPRL):
    """Same module, but the score/select/loss block runs as the reference's torch op sequence."""

    def forward(self, x, y=None):
        from oracle import cpu_port   # baseline arm only
        z = self.encoder_result(x)
        mu, sigma = self.encoder_proxies()
        eps = self.gaussian_noise(samples=([self.num_classes, self.sample_num]), K=self.z_dim, seed=self.seed)
        loss = cpu_port.eprl_train_loss_graph(z, self.proxies, eps, y, self.z_dim)
        B = x.shape[0]
        return mu.repeat(B, 1, 1), sigma.repeat(B, 1, 1), loss, z


class StandInFusion(nn.Module):
    def __init__(self, batch, eprl_cls, noise):
        super().__init__()
        self.enc = StandInEncoders()
        kw = dict(num_classes=2, sample_num=800, batch_size=batch, noise=noise, validate_labels=False)
        self.eprl_f = eprl_cls(1024, **kw)
        self.eprl_o = eprl_cls(768, **kw)
        self.head = nn.Linear(512, 3072)
        self.fc = nn.Linear(3072, 2)

    def forward(self, fundus, oct_, y):
        a, b = self.enc(fundus, oct_)
        _, _, pl_f, zf = self.eprl_f(a, y)
        _, _, pl_o, zo = self.eprl_o(b, y)
        combined = self.head(torch.cat([zf.mean(1), zo.mean(1)], dim=1))   # [B,3072] like DILR's output
        ce = F.cross_entropy(self.fc(combined), y, label_smoothing=0.1)
        return ce + 0.3 * (pl_f + pl_o), combined


def run(batch, steps, noise, arm):
    torch.manual_seed(0)
    dev = "cuda"
    if arm == "ours":
        model = StandInFusion(batch, edrl_b200.EPRL, noise).to(dev)
        mmd = edrl_b200.MK_MMD
    else:
        from oracle import cpu_port
        model = StandInFusion(batch, TorchEPRL, noise).to(dev)
        mmd = cpu_port.mk_mmd_graph
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-6)
    g = torch.Generator(device=dev).manual_seed(1)
    fundus = torch.rand(batch, 3, 384, 384, device=dev, generator=g)
    oct_ = torch.rand(batch, 1, 96, 96, 96, device=dev, generator=g)
    y = torch.randint(0, 2, (batch,), device=dev, generator=g)
    fundus2 = (fundus + 0.5 * torch.randn_like(fundus)).clamp(0, 1)        # the sigma = 0.5 noise view
    oct2 = (oct_ + 0.5 * torch.randn_like(oct_)).clamp(0, 1)

    def step():
        opt.zero_grad(set_to_none=True)
        l1, c1 = model(fundus, oct_, y)
        l2, c2 = model(fundus2, oct2, y)
        loss = l1 + mmd(c1, c2)
        loss.backward()
        opt.step()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    return ms, float(loss.detach())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--noise", default="device", choices=["device", "reference"])
    a = ap.parse_args()
    out = {}
    for arm in ("ours", "torch_ops"):
        ms, loss = run(a.batch, a.steps, a.noise, arm)
        out[arm] = {"ms_per_step": ms, "samples_per_s": a.batch / ms * 1e3, "last_loss": loss}
    out["speedup"] = out["torch_ops"]["ms_per_step"] / out["ours"]["ms_per_step"]
    out["config"] = {"batch": a.batch, "noise": a.noise, "encoders": "stand-in patch embeddings, bf16 autocast"}
    import json
    print(json.dumps(out))


if __name__ == "__main__":
    main()
