"""A synthetic two-view EDRL training step on one GPU (BASELINE configs[2] shape: batch 64, bf16 encoders) with the hot
path -- and the rows SURVEY.md 8(f) puts next to it -- swapped between this package's kernels and the reference's torch op
sequences, and the whole step captured in a CUDA graph.

The published `MedFusion` cannot run as published; `examples/run_reference_driver.py` runs it through
the reference's own drivers, where the numpy loader sets the step time.  This harness is the GPU-resident view of the same
step: a STAND-IN caller with the reference's data contract -- random-init patch-embedding encoders producing
`[B,144,1024]` (fundus, from `[B,3,384,384]`) and `[B,216,768]` (OCT, from `[B,1,96,96,96]`) tokens under bf16 autocast,
two Essence-Point modules, 2048-wide projections feeding the Barlow-Twins cross-correlation loss of DILR
(code/fusion_net.py:656-677), a head emitting `combined_features [B,3072]`, the label-smoothed CE + the two KL terms
(:929-942), and `MK_MMD` between the clean and the noisy view (code/fusion_train.py:176-224); the two views come from the
preprocessed batch with the loader's `clip(x + N(0, 0.5), 0, 1)` (code/data_harvard.py:769-783).

Arms: `ours` (edrl_b200 kernels: EPRL, MK_MMD, bt_loss_cross, head_losses, noise_views), `ours_graph` (the same step, one
CUDA graph launch), `torch_ops` (the reference's op sequences for all of those on the same GPU).  Everything else is
identical in the three arms.

    python examples/edrl_step_synthetic.py [--batch 64] [--steps 10] [--noise device|reference]
"""
import argparse
import json
import os
import sys
import time
import types

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


def torch_bt_loss_cross(self, z1, z2, common_dim):
    """code/fusion_net.py:656-677 as written (torch ops)."""
    c = self.bn1(z1).T @ self.bn2(z2)
    c.div_(self.args.batch_size * 4)
    dc = int(common_dim)
    c_c, c_u = c[:dc, :dc], c[dc:, dc:]

    def off_diagonal(x):
        n = x.shape[0]
        return x.flatten()[:-1].view(n - 1, n + 1)[:, 1:].flatten()

    on_c = torch.diagonal(c_c).add(-1).pow(2).sum()
    off_c = off_diagonal(c_c).pow(2).sum()
    on_u = torch.diagonal(c_u).pow(2).sum()
    off_u = off_diagonal(c_u).pow(2).sum()
    return on_c + 0.0051 * off_c, on_c, off_c, on_u + 0.0051 * off_u, on_u, off_u


def torch_head_losses(pred, y, mf, sf, mo, so, smoothing=0.1):
    """code/fusion_net.py:929-942 with KL_between_normals (:390-402) as written."""
    p = pred[:, :2]
    with torch.no_grad():
        t = torch.zeros_like(p)
        t.fill_(smoothing / 1)
        t.scatter_(1, y.unsqueeze(1), 1.0 - smoothing)
    loss1 = torch.sum(-t * F.log_softmax(p, dim=-1), dim=-1).mean()

    def kl(mu_q, sigma_q):
        mu_p, sigma_p = torch.zeros_like(mu_q), torch.ones_like(sigma_q)
        k = mu_q.size(1)
        mu_diff = mu_p - mu_q
        ldq = torch.sum(2 * torch.log(torch.clamp(sigma_q, min=1e-8)), dim=1)
        ldp = torch.sum(2 * torch.log(torch.clamp(sigma_p, min=1e-8)), dim=1)
        fs = torch.sum(torch.div(sigma_q ** 2, sigma_p ** 2), dim=1) + torch.sum(torch.div(mu_diff * mu_diff, sigma_p ** 2), dim=1)
        return torch.mean(torch.mean((fs - k + ldp - ldq) * 0.5))

    return loss1, kl(mf, sf), kl(mo, so)


class StandInFusion(nn.Module):
    def __init__(self, batch, ours, noise):
        super().__init__()
        self.ours = ours
        self.enc = StandInEncoders()
        kw = dict(num_classes=2, sample_num=800, batch_size=batch, noise=noise, validate_labels=False)
        cls = edrl_b200.EPRL if ours else TorchEPRL
        self.eprl_f = cls(1024, **kw)
        self.eprl_o = cls(768, **kw)
        self.proj1 = nn.Linear(256, 2048)
        self.proj2 = nn.Linear(256, 2048)
        self.bn1 = nn.BatchNorm1d(2048, affine=False)
        self.bn2 = nn.BatchNorm1d(2048, affine=False)
        self.args = types.SimpleNamespace(batch_size=batch)
        self.head = nn.Linear(4096, 3072)
        self.fc = nn.Linear(3072, 2)

    def forward(self, fundus, oct_, y):
        a, b = self.enc(fundus, oct_)
        mu_f, sig_f, pl_f, zf = self.eprl_f(a, y)
        mu_o, sig_o, pl_o, zo = self.eprl_o(b, y)
        y1, y2 = self.proj1(zf.mean(1)), self.proj2(zo.mean(1))                 # [B, 2048] each, like DILR's y1 / y2
        bt = edrl_b200.bt_loss_cross if self.ours else torch_bt_loss_cross
        loss_c, _, _, loss_u, _, _ = bt(self, y1, y2, 1024)
        combined = self.head(torch.cat([y1, y2], dim=1))                        # [B, 3072] like DILR's output
        pred = self.fc(combined)
        hl = edrl_b200.head_losses if self.ours else torch_head_losses
        loss1, kl_f, kl_o = hl(pred, y, mu_f, sig_f, mu_o, sig_o)
        loss = loss1 + 0.01 * kl_f + 0.01 * kl_o + 0.3 * (pl_f + pl_o) + 0.001 * (loss_c + loss_u) / 2.0   # :942-948
        return loss, combined


def run(batch, steps, noise, arm, same_views=False):
    torch.manual_seed(0)
    dev = "cuda"
    ours = arm != "torch_ops"
    model = StandInFusion(batch, ours, noise).to(dev)
    if ours:
        mmd = edrl_b200.MK_MMD
    else:
        from oracle import cpu_port
        mmd = cpu_port.mk_mmd_graph
    graph = arm == "ours_graph"
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, weight_decay=1e-6, capturable=graph)
    g = torch.Generator(device=dev).manual_seed(1)
    fundus = torch.rand(batch, 3, 384, 384, device=dev, generator=g)
    oct_ = torch.rand(batch, 1, 96, 96, 96, device=dev, generator=g)
    y = torch.randint(0, 2, (batch,), device=dev, generator=g)
    loss_out = torch.zeros((), device=dev)

    def views():
        """the loader's clean / noisy views (code/data_harvard.py:722-731, 769-783)"""
        if ours and not same_views:
            f_lo, f_hi = edrl_b200.noise_views(fundus, sigma=0.5, seed=11)
            o_lo, o_hi = edrl_b200.noise_views(oct_, sigma=0.5, seed=11)
            return f_lo, o_lo, f_hi, o_hi
        return (fundus.clamp(0, 1), oct_.clamp(0, 1), (fundus + 0.5 * torch.randn_like(fundus)).clamp(0, 1),
                (oct_ + 0.5 * torch.randn_like(oct_)).clamp(0, 1))

    def step():
        opt.zero_grad(set_to_none=not graph)
        f_lo, o_lo, f_hi, o_hi = views()
        l1, c1 = model(f_lo, o_lo, y)
        l2, c2 = model(f_hi, o_hi, y)
        loss = l1 + mmd(c1, c2)
        loss.backward()
        opt.step()
        loss_out.copy_(loss.detach())

    n0 = edrl_b200.launch_count()
    if graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        n0 = edrl_b200.launch_count()
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            step()
        per_step = edrl_b200.launch_count() - n0
        run_step = cg.replay
    else:
        for _ in range(3):
            step()
        n0 = edrl_b200.launch_count()
        step()
        per_step = edrl_b200.launch_count() - n0
        run_step = step
    for _ in range(2):
        run_step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        run_step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / steps * 1e3
    return ms, float(loss_out), int(per_step)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--noise", default="device", choices=["device", "reference"])
    a = ap.parse_args()
    print(json.dumps(compare(a.batch, a.steps, a.noise)))


def compare(batch=64, steps=10, noise="device"):
    out = {}
    for arm in ("ours", "ours_graph", "torch_ops"):
        if arm == "ours_graph" and noise == "reference":
            continue                                         # the reference's CPU-drawn proxy noise cannot be captured
        try:
            ms, loss, launches = run(batch, steps, noise, arm)
            out[arm] = {"ms_per_step": ms, "samples_per_s": batch / ms * 1e3, "last_loss": loss}
            if arm != "torch_ops":
                out[arm]["edrl_kernel_launches_per_step"] = launches
        except Exception as exc:
            out[arm] = {"error": repr(exc)}
    if "ms_per_step" in out.get("ours", {}) and "ms_per_step" in out.get("torch_ops", {}):
        out["speedup_eager"] = out["torch_ops"]["ms_per_step"] / out["ours"]["ms_per_step"]
        if "ms_per_step" in out.get("ours_graph", {}):
            out["speedup_graph"] = out["torch_ops"]["ms_per_step"] / out["ours_graph"]["ms_per_step"]
    out["config"] = {"batch": batch, "noise": noise, "encoders": "stand-in patch embeddings, bf16 autocast",
                     "step": "views -> 2 x (encoders, 2 EPRL, DILR bt loss, head losses) -> MK_MMD -> backward -> Adam"}
    return out


if __name__ == "__main__":
    main()
