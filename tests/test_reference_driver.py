"""The reference's own training driver (fusion_train.py, unmodified, from the oracle/_ref copy or the mounted
checkout) runs end to end on the synthetic scaffolding -- BASELINE configs[0]: the whole model, MMD and Essence-Point
losses included, one CPU batch of 4.  The GPU arms of the same runner (configs[2], [4]) are in test_gpu_reference_driver.py."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "examples"))


def test_reference_training_driver_runs_on_cpu():
    import run_reference_driver as R
    if R.reference_dir() is None:
        pytest.skip("no reference checkout and no oracle/_ref copy")
    res = R.run(arm="reference", driver="fusion_train", device="cpu", batch=4, steps=2)
    assert "error" not in res, res
    assert res["train"]["steps"] >= 2 and res["train"]["step_period_ms_median"] > 0
    assert any("Train Epoch: 1" in ln for ln in res["driver_output_tail"])


def test_medfusion_patch_is_exact():
    """The two neutralised statements and nothing else differ from the reference source."""
    import difflib
    import importlib.util
    import run_reference_driver as R
    refdir = R.reference_dir()
    if refdir is None:
        pytest.skip("no reference source")
    spec = importlib.util.spec_from_file_location("_dropin_fn_src", os.path.join(R.PKG, "dropin", "fusion_net.py"))
    src = open(os.path.join(refdir, "fusion_net.py")).read()
    text = open(spec.origin).read()
    ns = {}
    start = text.index("_PATCHES = (")
    end = text.index("if _ref is not None:")
    exec(text[start:end], ns)
    patched = ns["_patch_medfusion"](src)
    diff = [l for l in difflib.unified_diff(src.splitlines(), patched.splitlines(), lineterm="", n=0)
            if l[:1] in "+-" and not l.startswith(("+++", "---"))]
    assert len(diff) == 6, diff            # 2 removed lines (the dead call) + 2 x (one line changed)
