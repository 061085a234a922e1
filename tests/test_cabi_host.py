"""CPU-side checks: the C-ABI library loads and exports every symbol include/edrl_b200.h declares,
the ctypes prototypes cover the header, and the product path has no CPU fallback."""
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "edrl_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(edrl_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as ge
    ge.build()
    import edrl_b200
    return edrl_b200


def test_header_symbols_exported_and_bound(pkg):
    names = _declared()
    assert len(names) >= 22
    lib = pkg._lib.load()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert sorted(pkg._lib.PROTOTYPES) == names
    assert lib.edrl_abi_version() == 1
    out = subprocess.run(["nm", "-D", "--defined-only", pkg._lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (edrl_[a-z0-9_]+)", out))
    assert exported == set(names)


def test_workspace_bytes_is_pure_host_logic(pkg):
    lib = pkg._lib.load()
    assert lib.edrl_mmd_workspace_bytes(0, 4, 8, 0) == 0
    small = lib.edrl_mmd_workspace_bytes(4, 4, 8, 0)
    assert small >= 2 * 128 * 32 * 4
    assert lib.edrl_mmd_workspace_bytes(4, 4, 8, 1) > small                     # hi + lo operands
    big = lib.edrl_mmd_workspace_bytes(8192, 8192, 512, 0)
    assert 2 * 16384 * 512 * 4 <= big <= 2 * 16384 * 512 * 4 + (1 << 20)


@pytest.mark.parametrize("ns,nt,d,rc,rc2,sms", [
    (8192, 8192, 512, 16384, 0, 148),      # headline: 128 panels on 74 pairs -> 74 whole + 54 x 4 slabs
    (65536, 65536, 1024, 131072, 0, 148),  # configs[3] on one GPU: 2 feature passes
    (65536, 65536, 1024, 8192, 8192, 148), # one of 8 ranks: its source rows + its target rows
    (64, 64, 3072, 128, 0, 148),           # the reference's training shape: 1 panel x 6 feature passes
    (300, 212, 700, 512, 0, 148), (5000, 4600, 700, 9600, 0, 148), (1000, 24, 96, 1024, 0, 132),
    (37, 53, 24, 17, 40, 148), (8192, 8192, 512, 16384, 0, 2)])
@pytest.mark.parametrize("flags", [0, 1, 2, 4])
def test_sweep_work_list_covers_every_tile_once(pkg, ns, nt, d, rc, rc2, sms, flags):
    """Host logic of the fused sweep (csrc/mmd.cu make_plan / sweep_item): every (virtual panel, column group) is swept
    by exactly one work item, slabs are non-empty, the slab count is what edrl_mmd_grad_slabs reports, and splitting
    never makes the critical path (items per CTA pair x groups) longer."""
    import ctypes
    lib = pkg._lib.load()
    plan = (ctypes.c_int * 10)()
    assert lib.edrl_mmd_sweep_plan(ns, nt, d, flags, rc, rc2, sms, plan) == 0
    panels, vpanels, full, split, items, pairs, groups, d_pad, quad, pass_feats = list(plan)
    assert panels == -(-rc // 128) + -(-rc2 // 128)
    assert quad == (1 if d_pad > 768 and flags != 1 else 0) and pass_feats == (1024 if quad else 512)   # (3xTF32: pairs only)
    assert vpanels == panels * -(-d_pad // pass_feats) and d_pad >= d and d_pad % (128 if flags == 4 else 64) == 0
    assert split in (1, 2, 4, 8) and split <= max(groups, 1)
    assert items == full + (vpanels - full) * split and 1 <= pairs <= max(sms // (4 if quad else 2), 1) and pairs <= items
    seen = {}
    load = [0] * pairs
    for it in range(items):
        if it < full:
            vp, g0, g1 = it, 0, groups
        else:
            q = it - full
            vp, s = full + q // split, q % split
            g0, g1 = s * groups // split, (s + 1) * groups // split
        assert 0 <= vp < vpanels and g0 < g1
        for g in range(g0, g1):
            assert (vp, g) not in seen
            seen[(vp, g)] = it
        load[it % pairs] += g1 - g0
    assert len(seen) == vpanels * groups
    unsplit = [0] * pairs
    for vp in range(vpanels):
        unsplit[vp % pairs] += groups
    assert max(load) <= max(unsplit)
    if sms == 148 and rc2 == 0:
        assert split == lib.edrl_mmd_grad_slabs(ns, nt, d, flags, rc, rc2) or not __import__("torch").cuda.is_available()


def test_sass_is_blackwell_native(pkg):
    """tcgen05 MMA / TMEM loads / TMA in the SASS of the shipped library (B200_PROFILING.md table)."""
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.isfile(cuobjdump):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", pkg._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCMMA" in sass or re.search(r"UTC\w*MMA", sass)
    assert "LDTM" in sass
    assert "UTMALDG" in sass
    assert "HMMA." not in sass.replace("UTCHMMA", "")          # no legacy mma.sync path


def test_no_cpu_fallback(pkg):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.MK_MMD(torch.randn(4, 8), torch.randn(4, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.topk_rows(torch.randn(4, 8), 2)
    m = pkg.EPRL(16, z_dim=8, sample_num=120, num_classes=2, batch_size=2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.randn(2, 5, 16), torch.tensor([0, 1]))
    # the product never imports the oracle
    for mod in list(sys.modules):
        assert not (mod.startswith("edrl_b200") and "oracle" in mod)
    pkg_dir = os.path.dirname(pkg.__file__)
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg_dir, fn)).read().replace("the oracle", "")


def test_missing_library_fails_loudly(pkg, monkeypatch, tmp_path):
    monkeypatch.setattr(pkg._lib, "_lib", None)
    monkeypatch.setattr(pkg._lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="missing"):
        pkg._lib.load()


def test_state_dict_names_match_reference(pkg, golden_dir):
    import numpy as np
    ge = np.load(os.path.join(golden_dir, "eprl_reference.npz"))
    m = pkg.EPRL(1024, num_classes=2, sample_num=800, batch_size=4)
    sd = m.state_dict()
    assert sorted(sd.keys()) == sorted(ge["state_dict_keys"].tolist())
    shapes = dict(zip(ge["state_dict_keys"].tolist(), ge["state_dict_shapes"].tolist()))
    for k, v in sd.items():
        assert str(tuple(v.shape)) == shapes[k]


def test_dropin_modules_shadow_the_reference_names(pkg, monkeypatch):
    import importlib
    import types
    dropin = os.path.join(os.path.dirname(pkg.__file__), "dropin")
    monkeypatch.syspath_prepend(dropin)
    for name in ("MMD", "fusion_net"):
        sys.modules.pop(name, None)
    mmd = importlib.import_module("MMD")
    assert mmd.MK_MMD is pkg.MK_MMD and mmd.gaussian_kernel is pkg.gaussian_kernel
    assert callable(mmd.compute_js_divergence) and callable(mmd.compute_kl_divergence)
    # with the reference mounted and its unpublished imports stubbed, the caller code (MedFusion) is the
    # reference's own and only EPRL is rebound
    from oracle import ref_loader
    if ref_loader.reference_available():
        stubs = {}
        for name in ("ot", "matplotlib", "matplotlib.pyplot", "Models", "Models.fundus_swin_network", "Models.unetr"):
            if name not in sys.modules:
                stubs[name] = types.ModuleType(name)
        if "Models.fundus_swin_network" in stubs:
            stubs["Models.fundus_swin_network"].build_model = lambda *a, **k: None
        if "Models.unetr" in stubs:
            stubs["Models.unetr"].UNETR_base_3DNet = lambda *a, **k: None
        for k, v in stubs.items():
            monkeypatch.setitem(sys.modules, k, v)
        fn = importlib.import_module("fusion_net")
        assert fn.REFERENCE_LOADED and hasattr(fn, "MedFusion") and hasattr(fn, "PoE")
        assert fn.EPRL is pkg.EPRL
    else:
        fn = importlib.import_module("fusion_net")
        assert fn.EPRL is pkg.EPRL
    for name in ("MMD", "fusion_net"):
        sys.modules.pop(name, None)
