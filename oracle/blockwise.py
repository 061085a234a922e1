"""Row-blocked fp64 restatement of the reference MK_MMD for sizes whose n x n matrices fit nowhere
(TEST INFRASTRUCTURE ONLY -- imported by ``tests/`` and by ``bench.py``'s parity leg, never by the product).

``oracle/edrl_oracle.py::mk_mmd_grad`` follows code/MMD.py literally on full n x n numpy matrices, which stops at
n of a few thousand.  BASELINE configs[3] is n = 131072: one fp64 n x n matrix is 137 GB.  This module evaluates the
SAME formulas (code/MMD.py:16-44, 60-72 and the autograd of them) one row block at a time with
torch fp64 tensors on whatever device the inputs live on (CPU in the ``not gpu`` tests, the B200 in the ``gpu`` tests):

  pass 1  sum of the clamped distances  -> sigma_0 = sum(L) / (n^2 - n) / mul^(num // 2)        (code/MMD.py:25-34,
          the reference's n x n reduction, not the O(nd) closed form the CUDA path uses)
  pass 2  M = sum_ij a_i a_j K_ij (block means of code/MMD.py:66-72 written as one weighted sum: a_i = 1/n_s on source
          rows, -1/n_t on target rows), D = sum_ij a_i a_j sum_k e^{-L/s_k} L / (s_k s_0)
  pass 3  for the requested rows only: G = (a_i a_j A_ij + c) [L_raw >= 0],  dZ_i = g sign(M) 4 (rowsum(G)_i z_i - (G Z)_i)

Pinned against ``edrl_oracle.mk_mmd_grad`` (itself pinned to reference-generated goldens) in
``tests/test_oracle_blockwise.py``.
"""
from __future__ import annotations

import torch


def _blocks(n, block):
    for r0 in range(0, n, block):
        yield r0, min(n, r0 + block)


@torch.no_grad()
def mk_mmd_blockwise(x, y, rows=None, kernel_mul=2.0, kernel_num=5, grad_out=1.0, block=1024):
    """(loss, M, sigma_0, dZ[rows]) of code/MMD.py:46-74 in fp64, never holding more than ``block`` x n entries.

    ``rows``: 1-D int64 tensor of row indices into Z = [X; Y] whose gradient rows are wanted (None: no gradient).
    """
    z = torch.cat([x, y], dim=0).to(torch.float64)                                   # code/MMD.py:21
    ns, nt = x.shape[0], y.shape[0]
    n = ns + nt
    dev = z.device
    sq = (z * z).sum(dim=1)                                                          # :25
    a = torch.cat([torch.full((ns,), 1.0 / ns, dtype=torch.float64, device=dev),
                   torch.full((nt,), -1.0 / nt, dtype=torch.float64, device=dev)])
    half = kernel_mul ** (kernel_num // 2)

    def dist_block(r0, r1):
        return sq[r0:r1, None] + sq[None, :] - 2.0 * (z[r0:r1] @ z.t())              # :26

    total = torch.zeros((), dtype=torch.float64, device=dev)
    for r0, r1 in _blocks(n, block):
        total += dist_block(r0, r1).clamp_(min=0.0).sum()                            # :27, :31
    sigma0 = total / (n * n - n) / half                                              # :31-34
    m_sum = torch.zeros((), dtype=torch.float64, device=dev)
    d_sum = torch.zeros((), dtype=torch.float64, device=dev)
    for r0, r1 in _blocks(n, block):
        l2 = dist_block(r0, r1).clamp_(min=0.0)
        w = a[r0:r1, None] * a[None, :]
        for k in range(kernel_num):                                                  # :37-42
            sk = sigma0 * kernel_mul ** k
            e = torch.exp(-l2 / sk)
            m_sum += (w * e).sum()                                                   # :66-69 as one weighted sum
            d_sum += (w * e * l2).sum() / (sk * sigma0)
    loss = m_sum.abs()                                                               # :72
    grad = None
    if rows is not None:
        rows = rows.to(dev)
        c = d_sum / ((n * n - n) * half)
        sgn = torch.sign(m_sum)
        grad = torch.empty(rows.numel(), z.shape[1], dtype=torch.float64, device=dev)
        for b0, b1 in _blocks(rows.numel(), block):
            ri = rows[b0:b1]
            l_raw = sq[ri, None] + sq[None, :] - 2.0 * (z[ri] @ z.t())
            l2 = l_raw.clamp(min=0.0)
            amat = torch.zeros_like(l2)
            for k in range(kernel_num):
                sk = sigma0 * kernel_mul ** k
                amat -= torch.exp(-l2 / sk) / sk
            g = (a[ri, None] * a[None, :] * amat + c) * (l_raw >= 0.0)
            grad[b0:b1] = grad_out * sgn * 4.0 * (g.sum(dim=1, keepdim=True) * z[ri] - g @ z)
    return loss, m_sum, sigma0, grad
