"""GPU parity, SURVEY.md 8f-3: the device-side clean / noisy views (csrc/views.cu) against the reference's numpy formula
(code/data_harvard.py:722-731, 769-783).  With an injected noise tensor: bit-exact in fp32.  With the device generator
(Philox + Box-Muller; numpy's MT19937 stream cannot be replayed by a counter-based generator): moments, determinism,
the shared-field semantics of the reference's per-item reseeding, clipping statistics."""
import numpy as np
import pytest
import torch

from gpu_util import have_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_gpu(), reason="needs a CUDA device")]


@pytest.mark.parametrize("shape", [(3, 1, 24, 24, 24), (2, 3, 37, 41), (5, 7)])
def test_injected_noise_is_the_reference_formula(shape):
    import edrl_b200
    rng = np.random.default_rng(sum(shape))
    x = rng.random(shape).astype(np.float32)
    nz = (rng.standard_normal(shape) * 0.5).astype(np.float32)
    lo, hi = edrl_b200.noise_views(torch.tensor(x).cuda(), noise=torch.tensor(nz).cuda())
    np.testing.assert_array_equal(lo.cpu().numpy(), np.clip(x, 0.0, 1.0))
    np.testing.assert_array_equal(hi.cpu().numpy(), np.clip(x + nz, np.float32(0.0), np.float32(1.0)))
    # uint8 input: the / 255 of the reference's loader is fused
    xb = rng.integers(0, 256, size=shape, dtype=np.uint8)
    lo, hi = edrl_b200.noise_views(torch.tensor(xb).cuda(), noise=torch.tensor(nz).cuda())
    xf = xb.astype(np.float32) / np.float32(255.0)
    np.testing.assert_array_equal(lo.cpu().numpy(), xf)
    np.testing.assert_array_equal(hi.cpu().numpy(), np.clip(xf + nz, np.float32(0.0), np.float32(1.0)))


def test_device_generator_statistics_and_semantics():
    import edrl_b200
    B, per = 4, 96 * 96 * 96
    x = torch.full((B, 1, 96, 96, 96), 0.5, device="cuda")
    lo, hi = edrl_b200.noise_views(x, sigma=0.05, seed=11)                 # sigma small: nothing clips at x = 0.5
    assert torch.equal(lo, x)
    nz = (hi - x).double()
    assert abs(nz.mean().item()) < 5e-5 and abs(nz.std().item() - 0.05) < 1e-4
    assert abs((nz ** 3).mean().item()) < 1e-6                             # symmetric
    assert abs((nz ** 4).mean().item() / 0.05 ** 4 - 3.0) < 0.02           # Gaussian kurtosis
    # the reference reseeds per item with the same seed: every item carries the same field
    assert torch.equal(hi[0], hi[1]) and torch.equal(hi[0], hi[3])
    lo2, hi2 = edrl_b200.noise_views(x, sigma=0.05, seed=11)
    assert torch.equal(hi, hi2)                                            # a pure function of (seed, index)
    _, hi3 = edrl_b200.noise_views(x, sigma=0.05, seed=12)
    assert not torch.equal(hi, hi3)
    _, hi4 = edrl_b200.noise_views(x, sigma=0.05, seed=11, shared_field=False)
    assert not torch.equal(hi4[0], hi4[1])
    c = np.corrcoef(hi4[0].flatten().cpu().numpy()[:100000], hi4[1].flatten().cpu().numpy()[:100000])[0, 1]
    assert abs(c) < 0.02
    # the reference's sigma = 0.5 on uniform data: the share of clipped values matches the closed form
    xu = torch.rand(2, 3, 384, 384, device="cuda")
    _, hu = edrl_b200.noise_views(xu, sigma=0.5, seed=11)
    from math import erf, exp, pi, sqrt
    # P(x + n <= 0), x ~ U(0,1), n ~ N(0, s): integral_0^1 Phi(-x / s) dx
    s = 0.5
    xs = np.linspace(0, 1, 20001)
    p0 = np.trapezoid([0.5 * (1 + erf(-v / s / sqrt(2))) for v in xs], xs)
    assert abs((hu == 0).float().mean().item() - p0) < 3e-3
    assert abs((hu == 1).float().mean().item() - p0) < 3e-3
    # odd per-item size (not a multiple of 4) keeps the shared field aligned per item
    xo = torch.full((3, 1001), 0.5, device="cuda")
    _, ho = edrl_b200.noise_views(xo, sigma=0.05, seed=5)
    assert torch.equal(ho[0], ho[1]) and torch.equal(ho[0], ho[2])
