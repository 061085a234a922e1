"""Multi-threaded CPU port of the reference MK_MMD arithmetic (TEST / BASELINE INFRASTRUCTURE ONLY).

The reference's own implementation of this path is PyTorch on the CPU (code/MMD.py imports only
torch) and cannot travel to the GPU box, so ``bench.py``'s ``cpu_baseline`` leg and its
``--impl reference`` arm time this port instead (kind = "port").  It issues the same tensor
operations the reference issues -- Gram by ``matmul``, broadcast distance, clamp, one ``exp`` per
bandwidth, block sums; backward through autograd -- with all host threads torch can use.

``mk_mmd_fwd_bwd`` is the whole algorithm (code/MMD.py:16-72 + autograd).  ``rowblock_fwd_bwd`` is a
bounded sample of the same workload: rows [r0, r0+m) of the n x n problem against all n columns,
forward and backward; a full step costs n/m such blocks.  Checked against the numpy oracle in
tests/test_cpu_port.py.
"""
from __future__ import annotations

import torch


def mk_mmd_fwd_bwd(x: torch.Tensor, y: torch.Tensor, kernel_mul: float = 2.0, kernel_num: int = 5):
    """loss, dX, dY of the reference algorithm (code/MMD.py:16-44, 60-72) on CPU tensors."""
    x = x.detach().clone().requires_grad_(True)
    y = y.detach().clone().requires_grad_(True)
    ns, nt = x.shape[0], y.shape[0]
    n = ns + nt
    z = torch.cat([x, y], dim=0)                                   # :21
    sq = (z ** 2).sum(dim=1, keepdim=True)                         # :25
    dist2 = (sq + sq.t() - 2 * (z @ z.t())).clamp(min=0.0)         # :26-27
    bw = dist2.sum() / (n * n - n)                                 # :31
    bw = bw / kernel_mul ** (kernel_num // 2)                      # :34
    kmat = sum(torch.exp(-dist2 / (bw * kernel_mul ** i)) for i in range(kernel_num))   # :37-42
    loss = (kmat[:ns, :ns].sum() / ns ** 2 + kmat[ns:, ns:].sum() / nt ** 2
            - kmat[:ns, ns:].sum() / (ns * nt) - kmat[ns:, :ns].sum() / (ns * nt)).abs()   # :66-72
    loss.backward()
    return loss.detach(), x.grad, y.grad


def rowblock_fwd_bwd(z: torch.Tensor, ns: int, r0: int, m: int, kernel_mul: float = 2.0, kernel_num: int = 5):
    """Rows [r0, r0+m) of the same computation against all n columns, forward + backward for those rows.

    The bandwidth comes from the O(n d) closed form (the full matrix is not available in a sample);
    everything else -- the m x n Gram block, distance, clamp, ``kernel_num`` exponentials, weighted
    block sums and autograd through all of it -- is the reference's operation sequence on the block.
    Returns (partial signed block sum, gradient w.r.t. the m rows from this block's terms).
    """
    n = z.shape[0]
    nt = n - ns
    zi = z[r0:r0 + m].detach().clone().requires_grad_(True)
    zall = z.detach()
    mean = zall.mean(dim=0, keepdim=True)
    bw = 2.0 * n * ((zall - mean) ** 2).sum() / (n * n - n) / kernel_mul ** (kernel_num // 2)
    sq_i = (zi ** 2).sum(dim=1, keepdim=True)
    sq_j = (zall ** 2).sum(dim=1, keepdim=True)
    dist2 = (sq_i + sq_j.t() - 2 * (zi @ zall.t())).clamp(min=0.0)
    kblk = sum(torch.exp(-dist2 / (bw * kernel_mul ** i)) for i in range(kernel_num))
    a = torch.cat([torch.full((ns,), 1.0 / ns, dtype=z.dtype), torch.full((nt,), -1.0 / nt, dtype=z.dtype)])
    part = (a[r0:r0 + m, None] * a[None, :] * kblk).sum()
    part.backward()
    return part.detach(), zi.grad


def eprl_train_fwd_bwd(z, proxies, eps, y, z_dim, k=100):
    """The reference's EPRL train-branch arithmetic after the encoder (code/fusion_net.py:137-150, 220-243) as
    the same torch op sequence (normalize over tokens / samples, expanded batched matmul, permute + mean,
    masked_select split, two top-k, exp/mean), forward + backward.  Device agnostic; baseline leg only."""
    import torch.nn.functional as F
    z = z.detach().clone().requires_grad_(True)
    proxies = proxies.detach().clone().requires_grad_(True)
    b = z.shape[0]
    mu = proxies[:, :z_dim]                                                   # :116-119
    sigma = F.softplus(proxies[:, z_dim:])
    z_proxy = mu.unsqueeze(1) + sigma.unsqueeze(1) * eps                      # :143-146
    z_norm = F.normalize(z, dim=1)                                            # :149
    zp_norm = F.normalize(z_proxy)                                            # :150
    zp_exp = zp_norm.unsqueeze(0).expand(b, -1, -1, -1)                       # :221
    att = torch.matmul(z_norm.unsqueeze(1), zp_exp.transpose(2, 3))           # :223
    att = att.permute(0, 2, 1, 3).mean(dim=1)                                 # :224-225
    mask = torch.zeros(b, att.shape[1], dtype=torch.bool, device=z.device)    # :230-231
    mask[torch.arange(b, device=z.device), y] = True
    pos = torch.masked_select(att, mask.unsqueeze(-1)).view(b, -1)            # :233-234
    neg = torch.masked_select(att, ~mask.unsqueeze(-1)).view(b, -1)
    tp, _ = torch.topk(pos, k, dim=1)                                         # :236-238
    tn, _ = torch.topk(neg, k, dim=1)
    loss = torch.mean(torch.exp(-tp.mean(dim=1) + tn.mean(dim=1)))            # :240-243
    loss.backward()
    return loss.detach(), z.grad, proxies.grad


def mk_mmd_graph(source, target, kernel_mul=2.0, kernel_num=5):
    """code/MMD.py:46-74 as a differentiable torch expression (keeps the autograd graph; baseline arm of
    examples/edrl_step_synthetic.py)."""
    ns, nt = source.shape[0], target.shape[0]
    n = ns + nt
    z = torch.cat([source, target], dim=0)
    sq = (z ** 2).sum(dim=1, keepdim=True)
    dist2 = (sq + sq.t() - 2 * (z @ z.t())).clamp(min=0.0)
    bw = dist2.sum() / (n * n - n)
    bw = bw / kernel_mul ** (kernel_num // 2)
    kmat = sum(torch.exp(-dist2 / (bw * kernel_mul ** i)) for i in range(kernel_num))
    return (kmat[:ns, :ns].sum() / ns ** 2 + kmat[ns:, ns:].sum() / nt ** 2
            - kmat[:ns, ns:].sum() / (ns * nt) - kmat[ns:, :ns].sum() / (ns * nt)).abs()


def eprl_train_loss_graph(z, proxies, eps, y, z_dim, k=100):
    """The train-branch proxy loss (code/fusion_net.py:137-150, 220-243) as a differentiable torch expression."""
    import torch.nn.functional as F
    b = z.shape[0]
    mu = proxies[:, :z_dim]
    sigma = F.softplus(proxies[:, z_dim:])
    z_proxy = mu.unsqueeze(1) + sigma.unsqueeze(1) * eps
    z_norm = F.normalize(z, dim=1)
    zp_norm = F.normalize(z_proxy)
    att = torch.matmul(z_norm.unsqueeze(1), zp_norm.unsqueeze(0).expand(b, -1, -1, -1).transpose(2, 3))
    att = att.permute(0, 2, 1, 3).mean(dim=1)
    mask = torch.zeros(b, att.shape[1], dtype=torch.bool, device=z.device)
    mask[torch.arange(b, device=z.device), y] = True
    pos = torch.masked_select(att, mask.unsqueeze(-1)).view(b, -1)
    neg = torch.masked_select(att, ~mask.unsqueeze(-1)).view(b, -1)
    tp, _ = torch.topk(pos, k, dim=1)
    tn, _ = torch.topk(neg, k, dim=1)
    return torch.mean(torch.exp(-tp.mean(dim=1) + tn.mean(dim=1)))
