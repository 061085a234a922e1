"""The torch CPU port used as bench.py's cpu_baseline / reference arm follows the oracle."""
import numpy as np
import torch

from oracle import cpu_port
from oracle import edrl_oracle as O
from oracle.gen_golden import MMD_CASES, mmd_inputs


def test_cpu_port_full_matches_oracle():
    x, y = mmd_inputs(*MMD_CASES[3])
    loss, dx, dy = cpu_port.mk_mmd_fwd_bwd(x, y)
    ref, _, rx, ry = O.mk_mmd_grad(x.numpy(), y.numpy())
    assert np.isclose(loss.item(), ref, rtol=1e-12)
    np.testing.assert_allclose(dx.numpy(), rx, rtol=1e-8, atol=1e-15)
    np.testing.assert_allclose(dy.numpy(), ry, rtol=1e-8, atol=1e-15)


def test_cpu_port_rowblocks_sum_to_signed_mean():
    x, y = mmd_inputs(*MMD_CASES[3])
    z = torch.cat([x, y])
    ns = x.shape[0]
    tot = 0.0
    for r0 in range(0, z.shape[0], 30):
        p, g = cpu_port.rowblock_fwd_bwd(z, ns, r0, min(30, z.shape[0] - r0))
        tot += p.item()
        assert g.shape == (min(30, z.shape[0] - r0), z.shape[1])
    _, m, _, _ = O.mk_mmd_grad(x.numpy(), y.numpy())
    assert np.isclose(tot, m, rtol=1e-9)
