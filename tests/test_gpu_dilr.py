"""GPU parity, SURVEY.md 8f-1: the fused DILR Barlow-Twins cross-correlation loss (csrc/dilr.cu) against the numpy oracle
(oracle/edrl_oracle.dilr_bt_loss_cross, itself pinned to the reference's own method in tests/test_oracle_golden.py) and
against the committed outputs of the unmodified reference (tests/golden/dilr_reference.npz).  Tolerance: the reference
is fp32 and so is the kernel -- values rtol 2e-5, gradients 2e-5 |g|_inf against the fp64 oracle on the same fp32 inputs."""
import os
import types

import numpy as np
import pytest
import torch

from gpu_util import dev, have_gpu
from oracle import edrl_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="needs a CUDA device")]


@pytest.mark.parametrize("key", ["tiny_f32", "odd_f32", "ref2048_f32"])
def test_bt_loss_cross_matches_reference_golden(key, golden_dir):
    import edrl_b200
    g = np.load(os.path.join(golden_dir, "dilr_reference.npz"))
    b, d, dc, bs, _ = (int(v) for v in g[key + "_cfg"])
    z1 = dev(g[key + "_z1"]).requires_grad_(True)
    z2 = dev(g[key + "_z2"]).requires_grad_(True)
    holder = types.SimpleNamespace(args=types.SimpleNamespace(batch_size=bs),
                                   bn1=torch.nn.BatchNorm1d(d, affine=False).cuda().train(),
                                   bn2=torch.nn.BatchNorm1d(d, affine=False).cuda().train())
    vals = edrl_b200.bt_loss_cross(holder, z1, z2, dc)
    assert len(vals) == 6 and all(v.dim() == 0 for v in vals)
    out = np.array([v.item() for v in vals])
    np.testing.assert_allclose(out, g[key + "_out"], rtol=2e-5)
    w = g["weights"]
    sum(float(wi) * v for wi, v in zip(w, vals)).backward()
    for mine, ref in ((z1.grad, g[key + "_dz1"]), (z2.grad, g[key + "_dz2"])):
        assert np.abs(mine.cpu().numpy() - ref).max() <= 3e-5 * np.abs(ref).max()
    # the BatchNorm side effects of the reference call: running statistics and the batch counter advance once
    np.testing.assert_allclose(holder.bn1.running_mean.cpu().numpy(), g[key + "_rm1"], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(holder.bn1.running_var.cpu().numpy(), g[key + "_rv1"], rtol=1e-5)
    assert int(holder.bn1.num_batches_tracked) == 1 and int(holder.bn2.num_batches_tracked) == 1


@pytest.mark.parametrize("B,D,dc,bs", [(64, 2048, 1024, 64), (32, 2048, 1024, 32), (7, 130, 50, 4), (100, 96, 0, 16),
                                       (300, 192, 192, 64), (1, 64, 32, 1)])
def test_bt_loss_cross_vs_oracle_shapes(B, D, dc, bs):
    """The reference's training shapes (batch 32 / 64, D = 2048, common half) and ragged ones: tiles that straddle the block
    boundary, an empty common or unique block, batches above one register chunk, a single row."""
    import edrl_b200
    rng = np.random.default_rng(B * 31 + D)
    z1 = (rng.standard_normal((B, D)) * 1.3 + 0.2).astype(np.float32)
    z2 = (0.5 * z1 + rng.standard_normal((B, D)) * 0.9 - 0.1).astype(np.float32)
    w = np.array([1.0, -0.4, 0.2, 0.6, 0.3, -0.1])
    ref, d1, d2 = O.dilr_bt_loss_cross(z1, z2, dc, bs, grad_w=w)
    a, b = dev(z1).requires_grad_(True), dev(z2).requires_grad_(True)
    out = edrl_b200.bt_loss_cross_values(a, b, dc, bs)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref, rtol=3e-5, atol=1e-6)
    (out * dev(w)).sum().backward()
    for mine, r in ((a.grad, d1), (b.grad, d2)):
        assert np.abs(mine.cpu().numpy() - r).max() <= 5e-5 * max(np.abs(r).max(), 1e-12), (B, D, dc)


def test_bt_loss_cross_eval_mode_and_torch_restatement():
    """Eval mode normalises with the running statistics; and the whole thing against the reference formula written with
    torch ops on the same GPU (train mode), gradients included."""
    import edrl_b200
    torch.manual_seed(0)
    B, D, dc, bs = 48, 512, 256, 48
    z1 = torch.randn(B, D, device="cuda") * 1.2 + 0.1
    z2 = 0.7 * z1 + 0.7 * torch.randn(B, D, device="cuda")
    bn1 = torch.nn.BatchNorm1d(D, affine=False).cuda()
    bn2 = torch.nn.BatchNorm1d(D, affine=False).cuda()

    def ref(x1, x2):
        c = bn1(x1).T @ bn2(x2) / (bs * 4)
        cc, cu = c[:dc, :dc], c[dc:, dc:]
        on_c = (torch.diagonal(cc) - 1).pow(2).sum()
        off_c = cc.pow(2).sum() - torch.diagonal(cc).pow(2).sum()
        on_u = torch.diagonal(cu).pow(2).sum()
        off_u = cu.pow(2).sum() - torch.diagonal(cu).pow(2).sum()
        return torch.stack([on_c + 0.0051 * off_c, on_c, off_c, on_u + 0.0051 * off_u, on_u, off_u])

    holder = types.SimpleNamespace(args=types.SimpleNamespace(batch_size=bs), bn1=bn1, bn2=bn2)
    for training in (True, False):
        bn1.train(training)
        bn2.train(training)
        state = [t.clone() for t in (bn1.running_mean, bn1.running_var, bn2.running_mean, bn2.running_var)]
        a, b = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
        r = ref(a, b)
        r.sum().backward()
        for t, s in zip((bn1.running_mean, bn1.running_var, bn2.running_mean, bn2.running_var), state):
            t.copy_(s)                                   # same starting statistics for our call
        a2, b2 = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
        o = torch.stack(edrl_b200.bt_loss_cross(holder, a2, b2, dc))
        o.sum().backward()
        assert torch.allclose(o, r, rtol=1e-4, atol=1e-6), training
        gm = a.grad.abs().max().item()
        assert (a2.grad - a.grad).abs().max().item() <= 1e-4 * gm and (b2.grad - b.grad).abs().max().item() <= 1e-4 * gm
