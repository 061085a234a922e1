"""The row-blocked fp64 oracle (oracle/blockwise.py, used for the N = 8192 and N = 65536 parity tests on the GPU)
against the literal numpy oracle, which is pinned to the reference-generated golden vectors."""
import numpy as np
import pytest
import torch

from oracle import blockwise
from oracle import edrl_oracle as O
from oracle.gen_golden import mmd_inputs


@pytest.mark.parametrize("case", [(104, 37, 53, 24, 0.2, 1.3), (105, 256, 256, 64, 0.1, 1.25), (7, 130, 200, 100, 0.0, 1.0)])
@pytest.mark.parametrize("block", [64, 1024])
def test_blockwise_matches_literal_oracle(case, block):
    x, y = mmd_inputs(*case)
    ns, nt = x.shape[0], y.shape[0]
    rows = torch.tensor([0, 1, ns - 1, ns, ns + nt - 1, (ns + nt) // 2])
    loss, m, s0, g = blockwise.mk_mmd_blockwise(x, y, rows=rows, grad_out=2.0, block=block)
    ref, mref, dx, dy = O.mk_mmd_grad(x.numpy(), y.numpy(), grad_out=2.0)
    dz = np.concatenate([dx, dy])
    assert np.isclose(loss.item(), ref, rtol=1e-11)
    assert np.isclose(m.item(), mref, rtol=1e-11)
    assert np.isclose(s0.item(), O.mmd_bandwidth(np.concatenate([x.numpy(), y.numpy()]), 2.0, 5), rtol=1e-12)
    np.testing.assert_allclose(g.numpy(), dz[rows.numpy()], rtol=1e-9, atol=1e-13)


def test_blockwise_duplicate_rows_and_other_hyperparameters():
    g = torch.Generator().manual_seed(3)
    x = torch.randn(40, 16, generator=g, dtype=torch.float64)
    y = torch.cat([x[:10], torch.randn(30, 16, generator=g, dtype=torch.float64) + 0.3])      # duplicates: clamp at 0
    rows = torch.arange(80)
    loss, _, _, gr = blockwise.mk_mmd_blockwise(x, y, rows=rows, kernel_mul=3.0, kernel_num=4, block=32)
    ref, _, dx, dy = O.mk_mmd_grad(x.numpy(), y.numpy(), 3.0, 4)
    assert np.isclose(loss.item(), ref, rtol=1e-11)
    np.testing.assert_allclose(gr.numpy(), np.concatenate([dx, dy]), rtol=1e-8, atol=1e-13)
