"""BASELINE configs[2] and [4] through the reference's own drivers (unmodified fusion_train.py / fusion_test.py from
the oracle/_ref copy) on the GPU: the swapped arm (this package's MK_MMD / EPRL bound under the reference's names) and
the reference arm (the reference's own PyTorch ops on the same GPU) both train for a few steps on the same synthetic data,
and the missing-modality evaluation runs from the checkpoint the training run saved."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]


def test_fusion_train_and_fusion_test_run_on_the_swapped_path(tmp_path):
    import run_reference_driver as R
    if R.reference_dir() is None:
        pytest.skip("no reference source on this box (oracle/_ref is built where /root/reference exists)")
    import edrl_b200
    sw = R.run(arm="swapped", driver="fusion_train", device="cuda", batch=16, steps=4, keep_dir=str(tmp_path))
    assert "error" not in sw, sw
    assert sw["train"]["steps"] >= 4
    # per training step: 4 EPRL calls (2 views x 2 modalities) x 2 C-ABI passes + MK_MMD forward/backward kernels
    assert sw["edrl_kernel_launches"] >= sw["train"]["steps"] * 10
    ref = R.run(arm="reference", driver="fusion_train", device="cuda", batch=16, steps=4)
    assert "error" not in ref, ref
    # same data, same seeds up to the proxy-noise stream position: the first epoch's mean training loss agrees loosely
    def loss_of(r):
        ln = [l for l in r["driver_output_tail"] if "Train Epoch" in l][0]
        return float(ln.split("Loss:")[1].split()[0])
    assert abs(loss_of(sw) - loss_of(ref)) <= 0.25 * abs(loss_of(ref)), (loss_of(sw), loss_of(ref))
    ck = sw.get("checkpoint")
    assert ck and os.path.isfile(ck)
    for missing in ("oct", "fundus"):
        ev = R.run(arm="swapped", driver="fusion_test", device="cuda", batch=16, steps=4, missing=missing, checkpoint=ck)
        assert "error" not in ev, ev
        assert ev["eval"]["volumes_per_s"] > 0 and ev["eval"]["missing_modality"] == missing
