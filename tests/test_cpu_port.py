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


def test_cpu_port_eprl_matches_oracle():
    rng = np.random.default_rng(4)
    b, t, f, s = 5, 12, 16, 130
    z = rng.standard_normal((b, t, f))
    prox = rng.standard_normal((2, 2 * f)) * 0.3
    eps = rng.standard_normal((2, s, f))
    y = np.array([0, 1, 1, 0, 1])
    loss, dz, dprox = cpu_port.eprl_train_fwd_bwd(torch.tensor(z), torch.tensor(prox), torch.tensor(eps),
                                                  torch.tensor(y), f)
    sp = np.log1p(np.exp(-np.abs(prox[:, f:]))) + np.maximum(prox[:, f:], 0)
    bw = O.eprl_train_backward(z, prox[:, :f], sp, eps, y, k=100)
    assert np.isclose(loss.item(), bw["loss"], rtol=1e-10)
    np.testing.assert_allclose(dz.numpy(), bw["dz"], rtol=1e-7, atol=1e-13)
    sig = 1.0 / (1.0 + np.exp(-prox[:, f:]))
    np.testing.assert_allclose(dprox.numpy(), np.concatenate([bw["dmu"], bw["dsigma"] * sig], 1), rtol=1e-7, atol=1e-13)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs next to ours) prints one JSON line with the contract's
    keys, the same metric / unit / config as our arm, and never touches a GPU."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--ref-budget-s", "1"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["impl"] == "reference" and j["unit"] == "samples/s" and j["higher_is_better"] is True
    assert j["config"]["N_per_side"] == 8192 and j["config"]["d"] == 512 and j["vs_baseline"] is None
    # the reference's own MMD.py from the oracle/_ref copy when build() made one, else the torch port; whole steps
    assert j["cpu_baseline"]["kind"] in ("reference", "port") and j["cpu_baseline"]["cores"] >= 1 and j["value"] > 0
    assert j["steps_measured"] == 1 and j["extrapolated"] is False
    assert set(j["config"]) == {"workload", "N_per_side", "d", "kernel_mul", "kernel_num", "parallelism", "step"}
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
