#!/usr/bin/env python
"""Run the reference's OWN drivers -- ``fusion_train.py`` / ``fusion_test.py``, unmodified -- on synthetic
Harvard-30K-shaped data, with the hot path either swapped for this package's sm_100a kernels or left as the reference's
PyTorch code (BASELINE configs[0], [2], [4]).  Test / benchmark scaffolding, not product.

    python examples/run_reference_driver.py --arm swapped   --batch 64 --steps 12          # configs[2]
    python examples/run_reference_driver.py --arm reference --batch 64 --steps 12          # same step, reference ops on the GPU
    python examples/run_reference_driver.py --arm reference --device cpu --batch 4 --steps 3   # configs[0]
    python examples/run_reference_driver.py --arm swapped --driver fusion_test --missing oct   # configs[4]

What runs: the driver file (from ``EDRL_REFERENCE_ROOT/code`` or the ``oracle/_ref`` copy) is executed with ``runpy`` as
``__main__`` with ``--dataset dr2 --model_name MedFusion --mode train&test``; ``sys.path`` = ``<pkg>/dropin`` :
``examples/ref_scaffold`` : the driver's directory.  ``from MMD import MK_MMD`` (code/fusion_train.py:11) therefore binds
``<pkg>/dropin/MMD.py`` in the swapped arm (the reference's ``MMD.py`` is pre-loaded under that name in the reference arm)
and ``DR_2.fusion_net.MedFusion`` is the reference's ``MedFusion`` with SURVEY.md F6's two statements neutralised and
``EPRL`` swapped or not.  Unpublished encoders / dataset are the stand-ins of ``examples/ref_scaffold`` (random init,
synthetic pixels: throughput only).  ``--device cpu`` makes ``.cuda()`` the identity (the drivers hard-code it).

Prints ONE JSON line: step period through the driver (host clock between the first forward calls of consecutive training
steps, loader and bookkeeping included), the model-forward time on the device, kernel launches of libedrl_b200.so, and
for ``fusion_test`` the evaluation throughput in volumes/s.
"""
from __future__ import annotations

import argparse
import contextlib
import importlib.util
import io
import json
import os
import runpy
import shutil
import statistics
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "robust-multimodal-learning-for-ophthalmic-disease-grading-via-disentangled-representation_b200")
SCAFFOLD = os.path.join(ROOT, "examples", "ref_scaffold")


def reference_dir():
    cands = []
    if os.environ.get("EDRL_REFERENCE_ROOT"):
        cands.append(os.path.join(os.environ["EDRL_REFERENCE_ROOT"], "code"))
    cands += [os.path.join(ROOT, "oracle", "_ref"), "/root/reference/code"]
    for c in cands:
        if os.path.isfile(os.path.join(c, "fusion_train.py")) and os.path.isfile(os.path.join(c, "MMD.py")):
            return c
    return None


def run(arm="swapped", driver="fusion_train", device="cuda", batch=64, steps=12, missing="", checkpoint=None,
        keep_dir=None, quiet=True, pool=4, workdir=None):
    """Returns the result dict (also what ``main`` prints)."""
    import torch
    refdir = reference_dir()
    if refdir is None:
        return {"unavailable": "no reference checkout and no oracle/_ref copy (run oracle/build_ref.py where /root/reference exists)"}
    if device == "cuda" and not torch.cuda.is_available():
        return {"unavailable": "no CUDA device"}
    n_train = batch * steps
    # KFold(5): 80 % train; the val / test loader wants at least one full batch of 16 (drop_last, code/fusion_train.py:593)
    n_files = max((n_train * 5 + 3) // 4 + 5, 85)
    if driver == "fusion_test":
        n_files = max(n_files, 5 * 16 * max(steps, 4) + 5)     # `steps` full evaluation batches of 16
    work = workdir or tempfile.mkdtemp(prefix="edrl_driver_")
    os.makedirs(os.path.join(work, "Your_train_path"), exist_ok=True)
    for i in range(n_files):
        open(os.path.join(work, "Your_train_path", f"{i:05d}"), "w").close()
    for sub in ("log/train_log", "log/val_log"):
        os.makedirs(os.path.join(work, sub), exist_ok=True)

    saved = dict(path=list(sys.path), argv=list(sys.argv), cwd=os.getcwd(), env=dict(os.environ),
                 modules=set(sys.modules))
    patched = []
    res = {"arm": arm, "driver": driver + ".py", "device": device, "batch": batch, "reference_dir": refdir}
    try:
        os.environ["EDRL_PATCH_MEDFUSION"] = "1"
        os.environ["EDRL_SWAP_EPRL"] = "1" if arm == "swapped" else "0"
        os.environ["EDRL_SYNTH_MISSING"] = missing
        os.environ["EDRL_SYNTH_POOL"] = str(pool)
        sys.path[:0] = [os.path.join(PKG, "dropin"), SCAFFOLD, refdir, ROOT]
        for name in ("MMD", "fusion_net", "DR_2", "DR_2.fusion_net", "DR_2.data_harvard", "baseline_models", "metrics",
                     "metrics2"):
            sys.modules.pop(name, None)
        if arm == "reference":
            spec = importlib.util.spec_from_file_location("MMD", os.path.join(refdir, "MMD.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            sys.modules["MMD"] = mod                            # the reference's own MK_MMD
        if device == "cpu":
            import torch.nn as nn
            patched = [(torch.Tensor, "cuda", torch.Tensor.cuda), (nn.Module, "cuda", nn.Module.cuda)]
            torch.Tensor.cuda = lambda self, *a, **k: self
            nn.Module.cuda = lambda self, *a, **k: self
        os.chdir(work)
        argv = [os.path.join(refdir, driver + ".py"), "--dataset", "dr2", "--model_name", "MedFusion", "--mode", "train&test",
                "--batch_size", str(batch), "--condition", "noise", "--condition_name", "Gaussian", "--name", "synthetic"]
        if driver == "fusion_test":
            argv += ["--start_epoch", "1", "--end_epochs", "0", "--checkpoint", checkpoint]
        else:
            argv += ["--start_epoch", "1", "--end_epochs", "1"]
        sys.argv = argv
        launches0 = None
        if arm == "swapped" and device == "cuda":
            import edrl_b200
            launches0 = edrl_b200.launch_count()
        out = io.StringIO()
        t0 = time.perf_counter()
        err = None
        with contextlib.redirect_stdout(out if quiet else sys.stdout), contextlib.redirect_stderr(out if quiet else sys.stderr):
            stdin = sys.stdin
            sys.stdin = open(os.devnull)
            import pdb
            set_trace = pdb.set_trace
            # fusion_test.py ends in a stray `import pdb; pdb.set_trace()` (code/fusion_test.py:759-760) after test() has
            # returned: a debugger prompt has no place in a batch run (under pytest it would wait for a terminal forever)
            pdb.set_trace = lambda *a, **k: None
            try:
                runpy.run_path(argv[0], run_name="__main__")
            except SystemExit:
                pass
            except Exception as exc:                            # bdb.BdbQuit from the stray pdb.set_trace() included
                if type(exc).__name__ != "BdbQuit":
                    err = f"{type(exc).__name__}: {exc}"
            finally:
                pdb.set_trace = set_trace
                sys.stdin.close()
                sys.stdin = stdin
        if device == "cuda":
            torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if err:
            res["error"] = err
            res["tail"] = out.getvalue()[-1500:]
            return res
        st = sys.modules["DR_2.fusion_net"].STATS
        train_t = [t for t, tr in zip(st["forward_host_t"], st["training"]) if tr]
        eval_t = [t for t, tr in zip(st["forward_host_t"], st["training"]) if not tr]
        fwd_ms = [a.elapsed_time(b) for a, b in st["events"]] if st["events"] else []
        fwd_train = [m for m, tr in zip(fwd_ms, st["training"]) if tr]
        fwd_eval = [m for m, tr in zip(fwd_ms, st["training"]) if not tr]
        if len(train_t) >= 6:
            firsts = train_t[0::2]                              # two forwards per training step (clean + noisy view)
            periods = [b - a for a, b in zip(firsts[:-1], firsts[1:])][1:]      # drop the first (warm-up) step
            res["train"] = {"steps": len(firsts), "step_period_ms_median": statistics.median(periods) * 1e3,
                            "step_period_ms_min": min(periods) * 1e3,
                            "samples_per_s": batch / statistics.median(periods),
                            "model_forward_ms_median": statistics.median(fwd_train[2:]) if len(fwd_train) > 2 else None}
        if len(eval_t) >= 3:
            ebatch = 16                                         # code/fusion_train.py:593 (val / test loader batch size)
            per = [b - a for a, b in zip(eval_t[:-1], eval_t[1:])][1:]
            res["eval"] = {"batches": len(eval_t), "batch": ebatch, "batch_period_ms_median": statistics.median(per) * 1e3,
                           "volumes_per_s": ebatch / statistics.median(per),
                           "model_forward_ms_median": statistics.median(fwd_eval[1:]) if len(fwd_eval) > 1 else None,
                           "missing_modality": missing or None}
        if launches0 is not None:
            import edrl_b200
            res["edrl_kernel_launches"] = int(edrl_b200.launch_count() - launches0)
        res["wall_s"] = wall
        ck = []
        for dp, _, fs in os.walk(os.path.join(work, "checkpoint")):
            ck += [os.path.join(dp, f) for f in fs if f.endswith(".pth")]
        if driver == "fusion_train":
            if not ck and st["models"]:                         # accuracy 0.0 on the synthetic val split: save one ourselves
                os.makedirs(os.path.join(work, "checkpoint"), exist_ok=True)
                ck = [os.path.join(work, "checkpoint", "synthetic.pth")]
                torch.save({"epoch": 1, "state_dict": st["models"][-1].state_dict()}, ck[0])
            if ck and keep_dir:
                os.makedirs(keep_dir, exist_ok=True)
                dst = os.path.join(keep_dir, "medfusion_synthetic.pth")
                shutil.copyfile(ck[0], dst)
                res["checkpoint"] = dst
        loss_lines = [ln for ln in out.getvalue().splitlines() if "Loss:" in ln]
        res["driver_output_tail"] = loss_lines[-2:]
        return res
    finally:
        for obj, name, fn in patched:
            setattr(obj, name, fn)
        os.chdir(saved["cwd"])
        sys.path[:] = saved["path"]
        sys.argv = saved["argv"]
        for k in list(os.environ):
            if k not in saved["env"]:
                del os.environ[k]
        os.environ.update(saved["env"])
        for name in list(sys.modules):
            if name not in saved["modules"] and (name in ("MMD", "fusion_net", "baseline_models", "metrics", "metrics2", "ot")
                                                 or name.split(".")[0] in ("DR_2", "glu2", "Models", "matplotlib")):
                del sys.modules[name]
        if workdir is None:
            shutil.rmtree(work, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", default="swapped", choices=["swapped", "reference"])
    ap.add_argument("--driver", default="fusion_train", choices=["fusion_train", "fusion_test"])
    ap.add_argument("--device", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--missing", default="", choices=["", "oct", "fundus"])
    ap.add_argument("--checkpoint", default=None)
    ap.add_argument("--keep-dir", default=None, help="copy the checkpoint the training run saved here")
    ap.add_argument("--pool", type=int, default=4)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(json.dumps(run(a.arm, a.driver, a.device, a.batch, a.steps, a.missing, a.checkpoint, a.keep_dir,
                         quiet=not a.verbose, pool=a.pool)))


if __name__ == "__main__":
    main()
