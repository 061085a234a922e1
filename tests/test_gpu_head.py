"""GPU parity, SURVEY.md 8f-4: the fused head losses (csrc/head.cu) against the reference's own op sequence
(code/fusion_net.py:929-942 and KL_between_normals :390-402, restated with torch ops; fp64 on the GPU as the truth)."""
import numpy as np
import pytest
import torch

from gpu_util import have_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="needs a CUDA device")]


def _kl_between_normals(mu_q, sigma_q):
    """code/fusion_net.py:390-402 with the prior of get_KL_loss (:838-850): zeros / ones."""
    mu_p, sigma_p = torch.zeros_like(mu_q), torch.ones_like(sigma_q)
    k = mu_q.size(1)
    mu_diff = mu_p - mu_q
    logdet_q = torch.sum(2 * torch.log(torch.clamp(sigma_q, min=1e-8)), dim=1)
    logdet_p = torch.sum(2 * torch.log(torch.clamp(sigma_p, min=1e-8)), dim=1)
    fs = torch.sum(sigma_q ** 2 / sigma_p ** 2, dim=1) + torch.sum(mu_diff * mu_diff / sigma_p ** 2, dim=1)
    return torch.mean(torch.mean((fs - k + logdet_p - logdet_q) * 0.5))


def _reference(pred, y, mf, sf, mo, so, smoothing=0.1, C=2):
    p = pred[:, :C]                                                           # :930
    with torch.no_grad():
        t = torch.zeros_like(p)
        t.fill_(smoothing / (C - 1))
        t.scatter_(1, y.unsqueeze(1), 1.0 - smoothing)
    loss1 = torch.sum(-t * torch.log_softmax(p, dim=-1), dim=-1).mean()      # :939
    return loss1, _kl_between_normals(mf, sf), _kl_between_normals(mo, so)


@pytest.mark.parametrize("B,width,C,Cm,F", [(64, 2, 2, 2, 256), (32, 2, 2, 2, 256), (5, 4, 2, 3, 17), (7, 3, 3, 2, 64)])
def test_head_losses_match_reference_ops(B, width, C, Cm, F):
    import edrl_b200
    g = torch.Generator(device="cuda").manual_seed(B + F)
    pred = torch.randn(B, width, device="cuda", generator=g) * 2
    y = torch.randint(0, C, (B,), device="cuda", generator=g)
    mf, mo = (torch.randn(B, Cm, F, device="cuda", generator=g) * 0.3 for _ in range(2))
    sf, so = (torch.rand(B, Cm, F, device="cuda", generator=g) + 0.05 for _ in range(2))
    sf[0, 0, 0] = 1e-9                                                        # below the clamp: zero log-gradient
    w = torch.tensor([1.0, 0.01, 0.02], device="cuda")
    ins = [t.clone().requires_grad_(True) for t in (pred, mf, sf, mo, so)]
    out = torch.stack(edrl_b200.head_losses(ins[0], y, ins[1], ins[2], ins[3], ins[4], 0.1, C))
    (out * w).sum().backward()
    ref_in = [t.double().clone().requires_grad_(True) for t in (pred, mf, sf, mo, so)]
    ref = torch.stack(_reference(ref_in[0], y, *ref_in[1:], C=C))
    (ref * w.double()).sum().backward()
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref.detach().cpu().numpy(), rtol=2e-5)
    for a, b in zip(ins, ref_in):
        gm = b.grad.abs().max().item()
        assert (a.grad.double() - b.grad).abs().max().item() <= 2e-5 * gm + 1e-12
    assert ins[0].grad.shape == pred.shape and (width == C or ins[0].grad[:, C:].abs().max().item() == 0.0)
