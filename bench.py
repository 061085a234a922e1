#!/usr/bin/env python
"""Benchmark of the EDRL hot path on B200 -- metric and configs from BASELINE.json.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one MK_MMD forward + backward over one synthetic batch (Part A of the path; Part B -- the Essence-Point
select -- is HBM-bound and is measured in the same run on the scaled sweep of SURVEY.md 8d: `roofline.part_b_select`).
  N = 1 : BASELINE configs[1] at its largest size: N = 8192 samples per side, d = 512, fp32 inputs.
  N > 1 : BASELINE configs[3]: N = 65536 per side, d = 1024, row-block sharded over the ranks
          (NCCL all-gather of the feature rows + all-reduce of two partial sums); fixed total work.  The N = 1 line also
          carries `scale_anchor`: the same configs[3] workload on one GPU through the same public API.
`value` = samples per side / step time, inputs resident in HBM, timed with CUDA events per step
(L2 flushed between steps, untimed), max over ranks.  `e2e` = the same step through the public API
from pinned HOST buffers (H2D of X and Y, D2H of the loss and both gradients inside the timed region).
`roofline` = the dominant kernel (the fused forward+gradient Gram sweep) alone, algorithmic FLOPs / its duration
(`roofline.part_a` repeats it; `roofline.part_b_select` is the select kernel against measured HBM bandwidth).
`cpu_baseline` / `--impl reference` = the reference's OWN MMD.py (the verbatim copy `oracle/build_ref.py` makes under
oracle/_ref/, kind "reference"; the torch-CPU port oracle/cpu_port.py when that copy is absent, kind "port") on the
box's host cores: whole steps of configs[1] for as many of --steps as a time budget allows (`steps_measured`); a
row-block sample, labelled `extrapolated`, for configs[3], which no single device can run (773 GB of activations).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "MMD+Essence-Point fwd+bwd samples/sec"
UNIT = "samples/s"
WORKLOADS = {
    "sweep8192": dict(N=8192, d=512, seed=1013,
                      name="MK_MMD fwd+bwd, N=8192 per side, d=512, fp32 inputs (BASELINE configs[1], largest size)"),
    "sharded65536": dict(N=65536, d=1024, seed=2000,
                         name="MK_MMD fwd+bwd, N=65536 per side, d=1024, row-block sharded (BASELINE configs[3])"),
}


# ----------------------------------------------------------------------------------------- helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm_gbs=float(p["hbm_gbs"]), bf16_burst=float(p["bf16_tflops"]),
                    bf16_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm)}


def profiled_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
    (profiles/traffic.json, written by tools/summarize_profiles.py); None when no capture is committed."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(path) as f:
            t = json.load(f)["dram_bytes_per_launch"]
        for k, v in t.items():
            if kernel_substr in k:
                return float(v)
    except Exception:
        pass
    return None


def bind_to_gpu_cpus(index):
    """Pin this process to the CPUs NVML reports as local to the GPU (its NUMA node) before any pinned host buffer is
    allocated: the e2e leg streams 64 MiB per step over PCIe, and a cross-socket hop costs a third of the bandwidth.
    Returns the number of CPUs bound to (0 = left alone)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = {c for c in cpus if c in allowed}
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def make_inputs(N, d, seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn(N, d, device=device, generator=g)
    y = torch.randn(N, d, device=device, generator=g) * 1.25 + 0.1
    return x, y


class L2Flush:
    def __init__(self, device, mib=256):
        self.buf = torch.empty(mib << 20, dtype=torch.uint8, device=device)

    def __call__(self):
        self.buf.zero_()


def timed_steps(step_fn, steps, warmup, flush, world):
    """W untimed warm-ups, then exactly K steps each bracketed by CUDA events on the current stream
    (L2 flush between, untimed).  Returns total ms (max over ranks when world > 1)."""
    import torch.distributed as dist
    for _ in range(warmup):
        step_fn()
        flush()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
    total = sum(a.elapsed_time(b) for a, b in evs)
    if world > 1:
        t = torch.tensor([total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total = float(t.item())
    return total


def config_of(wl, world, prec):
    """The `config` object -- the same dict in both arms (the driver compares them)."""
    N, d = wl["N"], wl["d"]
    return {"workload": wl["name"], "N_per_side": N, "d": d, "kernel_mul": 2.0, "kernel_num": 5,
            "parallelism": f"row-block x{world}" if world > 1 else "single GPU",
            "step": "MK_MMD forward + backward (Part A); Part B select measured on the scaled sweep, see roofline.part_b_select"}


def reference_mk_mmd():
    """(callable(x, y) -> loss tensor with autograd graph, kind).  The reference's own MK_MMD from the oracle/_ref copy
    (baseline legs only), else the port."""
    from oracle import build_ref, cpu_port
    if build_ref.available():
        return build_ref.load_mmd().MK_MMD, "reference"
    return cpu_port.mk_mmd_graph, "port"


def cpu_whole_steps(N, d, seed, max_steps, budget_s):
    """Whole fwd+bwd steps of the reference MK_MMD on the host cores (all threads): median seconds per step, steps done."""
    fn, kind = reference_mk_mmd()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(N, d, generator=g)
    y = torch.randn(N, d, generator=g) * 1.25 + 0.1
    xs, ys = x[:512].clone().requires_grad_(True), y[:512].clone().requires_grad_(True)
    fn(xs, ys).backward()                                   # thread pool / allocator warm-up on a small problem
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < max_steps and (not times or time.perf_counter() + statistics.median(times) < t_end):
        a = x.clone().requires_grad_(True)
        b = y.clone().requires_grad_(True)
        t0 = time.perf_counter()
        fn(a, b).backward()
        times.append(time.perf_counter() - t0)
        del a, b
    return statistics.median(times), len(times), cores, kind


# ----------------------------------------------------------------------------------------- CPU arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = WORKLOADS["sweep8192" if world == 1 else "sharded65536"] if args.workload == "auto" else WORKLOADS[args.workload]
    N, d = wl["N"], wl["d"]
    n = 2 * N
    if n * d <= 16384 * 512:
        # configs[1]: WHOLE steps of the reference's MK_MMD (code/MMD.py:46-74 + autograd), as many of --steps as fit
        t_step, done, cores, kind = cpu_whole_steps(N, d, wl["seed"], args.steps, args.ref_budget_s)
        sample = (f"{done} whole fwd+bwd steps of the reference MK_MMD at N={N} per side, d={d} on {cores} host threads "
                  f"(time budget {args.ref_budget_s:.0f} s; --steps asked for {args.steps})")
        extra = {"steps_measured": done, "extrapolated": False}
    else:
        # configs[3]: no device can hold the reference's n x n activations (773 GB): a row-block sample, extrapolated
        from oracle import cpu_port
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        g = torch.Generator().manual_seed(wl["seed"])
        z = torch.cat([torch.randn(N, d, generator=g), torch.randn(N, d, generator=g) * 1.25 + 0.1])
        m = 128
        for _ in range(2):
            cpu_port.rowblock_fwd_bwd(z, N, 0, m)
        times, r0 = [], 0
        t_end = time.perf_counter() + args.ref_budget_s
        while len(times) < args.steps and time.perf_counter() < t_end:
            t0 = time.perf_counter()
            cpu_port.rowblock_fwd_bwd(z, N, r0, m)
            times.append(time.perf_counter() - t0)
            r0 = (r0 + m) % (n - m)
        t_step = statistics.median(times) * (n / m)
        done, kind = len(times), "port"
        sample = (f"rows [r0, r0+{m}) of the {n}x{n} problem (d={d}) fwd+bwd, torch CPU ops ({done} blocks timed), "
                  f"x{n // m} = one step: EXTRAPOLATED -- the reference cannot run this size on any single device")
        extra = {"steps_measured": 0, "blocks_measured": done, "extrapolated": True}
    value = N / t_step
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_step * 1e3, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_of(wl, world, None),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    line.update(extra)
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- extras
def essence_point_extras(peaks):
    """Essence-Point select / gather on the scaled sweep of SURVEY.md 8d (HBM-bound), outside the timed region."""
    import edrl_b200
    out = {}
    try:
        R, W, k = 1 << 18, 800, 100
        x = torch.randn(R, W, device="cuda")
        for _ in range(2):
            edrl_b200.topk_rows(x, k)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            edrl_b200.topk_rows(x, k)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        byts = R * W * 4 + R * k * 8
        out["select_topk"] = {"rows": R, "width": W, "k": k, "ms": ms, "GB/s": byts / ms / 1e6,
                              "frac_hbm": byts / ms / 1e6 / peaks["hbm_gbs"], "rows_per_s": R / ms * 1e3,
                              "order": "descending (torch.topk parity)"}
        for _ in range(2):
            edrl_b200.topk_rows(x, k, sorted=False)        # (first launch of this instantiation: module load)
        torch.cuda.synchronize()
        a.record()
        for _ in range(5):
            edrl_b200.topk_rows(x, k, sorted=False)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 5
        out["select_topk_unsorted"] = {"rows": R, "width": W, "k": k, "ms": ms, "GB/s": byts / ms / 1e6,
                                       "frac_hbm": byts / ms / 1e6 / peaks["hbm_gbs"],
                                       "order": "selection set only (what the Essence-Point loss consumes)"}
        del x
        # the rest of the scaled select sweep of SURVEY.md 8d (R x W x 4 bytes >= 256 MB each)
        sweep = []
        for (R2, W2) in ((1 << 17, 1600), (1 << 15, 8192), (1 << 20, 800)):
            x2 = torch.randn(R2, W2, device="cuda")
            row = {"rows": R2, "width": W2, "k": k}
            for srt in (False, True):
                for _ in range(2):
                    edrl_b200.topk_rows(x2, k, sorted=srt)
                torch.cuda.synchronize()
                a.record()
                for _ in range(5):
                    edrl_b200.topk_rows(x2, k, sorted=srt)
                b.record()
                torch.cuda.synchronize()
                ms2 = a.elapsed_time(b) / 5
                byts2 = R2 * W2 * 4 + R2 * k * 8
                row["sorted" if srt else "unsorted"] = {"ms": ms2, "GB/s": byts2 / ms2 / 1e6,
                                                        "frac_hbm": byts2 / ms2 / 1e6 / peaks["hbm_gbs"]}
            sweep.append(row)
            del x2
        out["select_sweep"] = sweep
        out["select_sweep_note"] = ("512 <= W <= 2048 (W % 4 == 0): one warp per row, topk_sift_kernel; 2048 < W <= 8192: "
                                    "topk_sift_stream_kernel (one warp per row, the row streamed) + a marked-row pass of "
                                    "topk_vecblock_kernel; other widths: topk_vec_kernel / radix kernels")
        B, T, D, kk = 4096, 216, 768, 32
        feat = torch.randn(B, T, D, device="cuda")
        idx = torch.stack([torch.randperm(T, device="cuda")[:kk] for _ in range(64)]).repeat(B // 64, 1).int()
        for _ in range(2):
            edrl_b200.gather_rows(feat, idx)
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            edrl_b200.gather_rows(feat, idx)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        byts = 2 * B * kk * D * 4
        out["gather_rows"] = {"B": B, "T": T, "D": D, "k": kk, "ms": ms, "GB/s": byts / ms / 1e6,
                              "frac_hbm": byts / ms / 1e6 / peaks["hbm_gbs"]}
        # the north-star's formulation: score per slice, top-k over the T slices, gather the k feature rows
        scores = torch.randn(B, T, device="cuda")
        for _ in range(2):
            edrl_b200.select_gather(feat, scores, kk)
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            edrl_b200.select_gather(feat, scores, kk)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        byts = B * T * 4 + B * kk * 8 + 2 * B * kk * D * 4
        out["select_gather"] = {"B": B, "T": T, "D": D, "k": kk, "ms": ms, "GB/s": byts / ms / 1e6,
                                "frac_hbm": byts / ms / 1e6 / peaks["hbm_gbs"],
                                "what": "topk_rows over [B,T] scores (sorted, indices bit-exact vs torch.topk) + "
                                        "gather_rows of [B,k,D] features; bytes = scores + values/indices + 2 B k D 4"}
    except Exception as exc:   # extras never take the headline down
        out["error"] = repr(exc)
    return out


def select_roofline(peaks):
    """Part B of the path: the Essence-Point select (top-k of 100 over rows of 800 proxy-sample scores,
    code/fusion_net.py:233-238) on the scaled sweep of SURVEY.md 8d -- 2^18 rows x 800, R W 4 bytes = 839 MB, far larger
    than L2 -- against the measured HBM copy bandwidth.  Algorithmic bytes: R W 4 read + R k 8 written.  `unsorted` is what
    the path consumes (the reference averages the selected values and discards the indices; edrl_essence_train_fwd calls
    the select with sorted = 0); `sorted` is torch.topk's output order."""
    import edrl_b200
    out = {"bound": "hbm", "peak": peaks["hbm_gbs"], "unit": "GB/s",
           "peak_source": f"{peaks['source']} HBM copy bandwidth (MEASURED_PEAKS.json)",
           "kernel": "topk_sift_kernel<6, true, 3, 8, SORTED, Rows, 10> (one warp per row: sample pivot -> survivors -> exact select)",
           "traffic": profiled_traffic("topk_sift_kernel<6, 1, 3, 8, 0"),
           "traffic_note": "DRAM bytes per launch of the unsorted 2^18 x 800 select (ncu --set full, profiles/)"}
    R, W, k = 1 << 18, 800, 100
    x = torch.randn(R, W, device="cuda")
    byts = R * W * 4 + R * k * 8
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, srt in (("unsorted", False), ("sorted", True)):
        for _ in range(3):
            edrl_b200.topk_rows(x, k, sorted=srt)
        torch.cuda.synchronize()
        a.record()
        for _ in range(10):
            edrl_b200.topk_rows(x, k, sorted=srt)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        ent = {"ms": ms, "achieved": byts / ms / 1e6, "frac": byts / ms / 1e6 / peaks["hbm_gbs"]}
        if srt:
            out["sorted"] = ent
        else:
            out.update(ent)
            out["algorithmic_bytes"] = byts
            out["workload"] = f"{R} rows x {W}, k = {k}, selection set (unsorted)"
    del x
    return out


def reference_driver_runs(steps=8):
    """BASELINE configs[0], [2], [4] through the reference's OWN drivers (fusion_train.py / fusion_test.py, unmodified, from
    the oracle/_ref copy) on the synthetic scaffolding of examples/ref_scaffold: see examples/run_reference_driver.py."""
    sys.path.insert(0, os.path.join(ROOT, "examples"))
    out = {}
    try:
        import tempfile
        import run_reference_driver as RD
        if RD.reference_dir() is None:
            return {"unavailable": "no oracle/_ref copy of the reference drivers on this box"}
        keep = tempfile.mkdtemp(prefix="edrl_ckpt_")
        out["configs2_full_step_batch64"] = {
            "swapped": RD.run(arm="swapped", driver="fusion_train", device="cuda", batch=64, steps=steps, keep_dir=keep),
            "reference_ops_same_gpu": RD.run(arm="reference", driver="fusion_train", device="cuda", batch=64, steps=steps)}
        ck = out["configs2_full_step_batch64"]["swapped"].get("checkpoint")
        if ck:
            out["configs4_missing_modality"] = {
                m: {arm: RD.run(arm=arm, driver="fusion_test", device="cuda", batch=64, steps=6, missing=m, checkpoint=ck)
                    for arm in ("swapped", "reference")} for m in ("oct", "fundus")}
        out["configs0_cpu_whole_model_batch4"] = RD.run(arm="reference", driver="fusion_train", device="cpu", batch=4, steps=3)
        out["note"] = ("the drivers are data-loader bound at these sizes (numpy noise views and pageable H2D copies in the "
                       "reference's own loop): `step_period_ms` is the whole loop, `model_forward_ms` the device time of one "
                       "MedFusion.forward; stand-in encoders, synthetic data, random init")
    except Exception as exc:
        out["error"] = repr(exc)
    return out


def next_rows_extras(peaks):
    """the rows next to the path -- DILR Barlow-Twins loss (8f-1), device-side noise views (8f-3), fused head
    losses + the whole stand-in step as one CUDA graph (8f-4) -- timed against the reference's torch op sequences."""
    import importlib.util
    import types
    import edrl_b200
    out = {}
    try:
        spec = importlib.util.spec_from_file_location("edrl_step_synthetic", os.path.join(ROOT, "examples", "edrl_step_synthetic.py"))
        syn = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(syn)

        def timeit(fn, reps=20):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                fn()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            return statistics.median(ts)

        rows = []
        for B in (32, 64, 256):
            holder = types.SimpleNamespace(args=types.SimpleNamespace(batch_size=B),
                                           bn1=torch.nn.BatchNorm1d(2048, affine=False).cuda(),
                                           bn2=torch.nn.BatchNorm1d(2048, affine=False).cuda())
            z1 = torch.randn(B, 2048, device="cuda")
            z2 = 0.6 * z1 + 0.8 * torch.randn(B, 2048, device="cuda")

            def run(fn):
                a, b = z1.clone().requires_grad_(True), z2.clone().requires_grad_(True)
                v = fn(holder, a, b, 1024)
                ((v[0] + v[3]) / 2).backward()

            ours = timeit(lambda: run(edrl_b200.bt_loss_cross))
            ref = timeit(lambda: run(syn.torch_bt_loss_cross))
            rows.append({"B": B, "D": 2048, "ours_ms": ours, "torch_eager_gpu_ms": ref, "speedup": ref / ours})
        out["dilr_bt_loss_cross_fwd_bwd"] = {"rows": rows, "bound": "launch latency at the reference's batch sizes (4 kernels "
                                             "against ~60 launches); fp32 CUDA cores, 2 x 1024^2 x B MACs each way"}
        # noise views: batch 64 of the reference's shapes, HBM-bound (4 B read + 8 B written per element)
        x = torch.rand(64, 1, 96, 96, 96, device="cuda")
        ms = timeit(lambda: edrl_b200.noise_views(x, 0.5, 11), 10)
        byts = x.numel() * 12
        ref = timeit(lambda: (x.clamp(0, 1), (x + 0.5 * torch.randn_like(x)).clamp(0, 1)), 10)
        out["noise_views_oct_batch64"] = {"ms": ms, "GB/s": byts / ms / 1e6, "frac_hbm": byts / ms / 1e6 / peaks["hbm_gbs"],
                                          "torch_eager_gpu_ms": ref, "bytes": byts,
                                          "cpu_reference": "numpy on the loader's workers: ~20 ms per SAMPLE (reference_drivers)"}
        out["synthetic_step_batch64"] = syn.compare(64, 10, "device")
    except Exception as exc:
        out["error"] = repr(exc)
    return out


def essence_path_vs_torch_gpu():
    """Essence-Point score + select + loss, forward + backward after the encoder, at the
    reference's shapes (C=2, S=800, F=256, k=100) against the reference's torch op sequence on the same GPU,
    plus the token-statistics stream on a scaled batch as an HBM figure."""
    import edrl_b200
    from oracle import cpu_port
    out = []
    try:
        for (B, T) in ((64, 216), (64, 144), (512, 216)):
            Fd, S, C = 256, 800, 2
            z = torch.randn(B, T, Fd, device="cuda")
            prox = torch.randn(C, 2 * Fd, device="cuda") * 0.1
            eps = torch.randn(C, S, Fd, device="cuda")
            y = torch.randint(0, 2, (B,), device="cuda")

            def ours():
                zz = z.detach().requires_grad_(True)
                pp = prox.detach().requires_grad_(True)
                edrl_b200.essence_train_loss(zz, pp, eps, y, 100).backward()

            def ref():
                cpu_port.eprl_train_fwd_bwd(z, prox, eps, y, Fd)

            res = {"B": B, "T": T}
            for name, fn in (("ours_ms", ours), ("torch_eager_gpu_ms", ref)):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                ts = []
                for _ in range(10):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    fn()
                    b.record()
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b))
                res[name] = statistics.median(ts)
            res["speedup"] = res["torch_eager_gpu_ms"] / res["ours_ms"]
            res["samples_per_s"] = B / res["ours_ms"] * 1e3
            out.append(res)
        # token statistics (normalize over tokens + token mean), forward and backward, B = 4096: 906 MB of z
        lib = edrl_b200._lib.load()
        B, T, Fd = 4096, 216, 256
        z = torch.randn(B, T, Fd, device="cuda")
        zbar, cs, cn = (torch.empty(B, Fd, device="cuda") for _ in range(3))
        dz = torch.empty_like(z)
        st = edrl_b200._lib.stream_and_device(z)
        for name, fn, byts in (
                ("token_stats_fwd", lambda: lib.edrl_token_stats_fwd(z.data_ptr(), B, T, Fd, zbar.data_ptr(), cs.data_ptr(),
                                                                   cn.data_ptr(), st), B * T * Fd * 4),
                ("token_stats_bwd", lambda: lib.edrl_token_stats_bwd(z.data_ptr(), cs.data_ptr(), cn.data_ptr(),
                                                                   zbar.data_ptr(), B, T, Fd, dz.data_ptr(), st),
                 2 * B * T * Fd * 4)):
            for _ in range(2):
                fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                fn()
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 5
            out.append({"kernel": name, "B": B, "T": T, "F": Fd, "ms": ms, "GB/s": byts / ms / 1e6,
                        "frac_hbm": byts / ms / 1e6 / measured_peaks()["hbm_gbs"]})
    except Exception as exc:
        out.append({"error": repr(exc)})
    return out


def eval_missing_modality():
    """BASELINE configs[4]: Essence-Point eval branch (pseudo-labels, no gradients) when one modality is missing,
    i.e. its tokens come from a zero-filled input (the reference has no modality switch; zero filling is the closest
    published semantics).  Volumes (samples) per second through EPRL.eval()."""
    import edrl_b200
    out = []
    try:
        for (name, T, xd) in (("oct_missing(zero OCT tokens)", 216, 768), ("fundus_missing(zero fundus tokens)", 144, 1024)):
            for noise in ("reference", "device"):
                for B in (16, 64):
                    torch.manual_seed(0)
                    m = edrl_b200.EPRL(xd, num_classes=2, sample_num=800, batch_size=B, noise=noise).cuda().eval()
                    x = torch.zeros(B, T, xd, device="cuda")
                    with torch.no_grad():
                        for _ in range(3):
                            m(x)
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        for _ in range(20):
                            m(x)
                        torch.cuda.synchronize()
                    ms = (time.perf_counter() - t0) / 20 * 1e3
                    out.append({"case": name, "noise": noise, "B": B, "ms": ms, "volumes_per_s": B / ms * 1e3})
    except Exception as exc:
        out.append({"error": repr(exc)})
    return out


def sweep_vs_torch_gpu(prec):
    """BASELINE configs[1]: MK_MMD fwd+bwd sweep, d=512, against the reference's torch op sequence run eagerly on
    the SAME GPU (oracle/cpu_port.mk_mmd_fwd_bwd is device agnostic) -- plus the reference's own training shape
    (N=64/side, d=3072).  Baseline leg only; median of a few runs, untimed warm-up."""
    import edrl_b200
    from oracle import cpu_port
    out = []
    try:
        for (N, d) in ((64, 3072), (256, 512), (512, 512), (1024, 512), (2048, 512), (4096, 512), (8192, 512)):
            x, y = make_inputs(N, d, 1000 + int(math.log2(N)), "cuda")

            def ours():
                a = x.detach().requires_grad_(True)
                b = y.detach().requires_grad_(True)
                edrl_b200.MK_MMD(a, b, precision=prec).backward()

            def ref():
                cpu_port.mk_mmd_fwd_bwd(x, y)

            res = {"N_per_side": N, "d": d}
            for name, fn, reps in (("ours_ms", ours, 10), ("torch_eager_gpu_ms", ref, 3 if N >= 4096 else 10)):
                for _ in range(2):
                    fn()
                torch.cuda.synchronize()
                ts = []
                for _ in range(reps):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    fn()
                    b.record()
                    torch.cuda.synchronize()
                    ts.append(a.elapsed_time(b))
                res[name] = statistics.median(ts)
            res["speedup"] = res["torch_eager_gpu_ms"] / res["ours_ms"]
            out.append(res)
            torch.cuda.empty_cache()
    except Exception as exc:
        out.append({"error": repr(exc)})
    return out


# ----------------------------------------------------------------------------------------- our arm
def run_ours(args, rank, local_rank, world):
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    bound_cpus = bind_to_gpu_cpus(local_rank)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import edrl_b200
    from edrl_b200 import _lib
    from edrl_b200.mmd import Workspace, _flags
    lib = _lib.load()
    peaks = measured_peaks()
    prec = args.precision
    wl_key = ("sweep8192" if world == 1 else "sharded65536") if args.workload == "auto" else args.workload
    wl = WORKLOADS[wl_key]
    N, d = wl["N"], wl["d"]
    flush = L2Flush(dev)
    sharded = world > 1

    if sharded:
        # every rank draws the whole problem from the one seed and keeps its row slice, so that rank 0 can evaluate the
        # same inputs unsharded (anchor + parity)
        assert N % world == 0
        nl = N // world
        xf, yf = make_inputs(N, d, wl["seed"], dev)
        x, y = xf[rank * nl:(rank + 1) * nl].clone(), yf[rank * nl:(rank + 1) * nl].clone()
        del xf, yf
        torch.cuda.empty_cache()
    else:
        nl = N
        x, y = make_inputs(N, d, wl["seed"], dev)
    x.requires_grad_(True)
    y.requires_grad_(True)

    def loss_fn(a, b):
        if sharded:
            return edrl_b200.sharded_MK_MMD(a, b, precision=prec)
        return edrl_b200.MK_MMD(a, b, precision=prec)

    def step():
        x.grad = None
        y.grad = None
        loss_fn(x, y).backward()

    # ---- device-resident timing (value)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed_steps(step, args.steps, args.warmup, flush, world)
    l0 = _lib.launch_count()
    step()
    launches = (_lib.launch_count() - l0) * args.steps      # our kernels per step x timed steps
    ms_per_step = total_ms / args.steps
    value = N / (ms_per_step * 1e-3)

    # ---- end to end from pinned host buffers (e2e)
    # Every step copies its X, Y from pinned host memory, runs forward+backward through the public API and copies
    # the loss and both gradients back.  `e2e_serial` runs the three phases back to back on one stream; `e2e`
    # software-pipelines them over the steps (copy-in stream / compute stream / copy-out stream, two device input
    # buffers) the way a training loop prefetches batches: same bytes and same kernels per step, PCIe both ways
    # overlapped with compute.
    hx = torch.empty(nl, d, pin_memory=True).copy_(x.detach())
    hy = torch.empty(nl, d, pin_memory=True).copy_(y.detach())
    hgx = [torch.empty(nl, d, pin_memory=True) for _ in range(2)]
    hgy = [torch.empty(nl, d, pin_memory=True) for _ in range(2)]
    hloss = torch.empty((), pin_memory=True)
    dxb = [torch.empty(nl, d, device=dev) for _ in range(2)]
    dyb = [torch.empty(nl, d, device=dev) for _ in range(2)]

    def e2e_step():
        dxb[0].copy_(hx, non_blocking=True)
        dyb[0].copy_(hy, non_blocking=True)
        a = dxb[0].detach().requires_grad_(True)
        b = dyb[0].detach().requires_grad_(True)
        loss = loss_fn(a, b)
        loss.backward()
        hloss.copy_(loss.detach(), non_blocking=True)
        hgx[0].copy_(a.grad, non_blocking=True)
        hgy[0].copy_(b.grad, non_blocking=True)

    e2e_serial_ms = timed_steps(e2e_step, args.steps, 3, flush, world) / args.steps

    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    s_cmp = torch.cuda.current_stream(dev)

    def pipelined(nsteps):
        ev_cmp = [None, None]
        for i in range(nsteps):
            k = i & 1
            with torch.cuda.stream(s_in):
                if ev_cmp[k] is not None:
                    s_in.wait_event(ev_cmp[k])          # input buffer k is free once step i-2 has computed
                dxb[k].copy_(hx, non_blocking=True)
                dyb[k].copy_(hy, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(s_in)
            s_cmp.wait_event(ev_in)
            a = dxb[k].detach().requires_grad_(True)
            b = dyb[k].detach().requires_grad_(True)
            loss = loss_fn(a, b)
            loss.backward()
            ev_cmp[k] = torch.cuda.Event()
            ev_cmp[k].record(s_cmp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_cmp[k])
                hloss.copy_(loss.detach(), non_blocking=True)
                hgx[k].copy_(a.grad, non_blocking=True)
                hgy[k].copy_(b.grad, non_blocking=True)
                for t in (loss, a.grad, b.grad):
                    t.record_stream(s_out)
        s_cmp.wait_stream(s_out)

    pipelined(max(10, args.warmup))        # also lets the caching allocator settle its cross-stream blocks
    # K pipelined steps, three times; the median is reported (which DMA engines the driver gives the two copy
    # directions varies from run to run: the same command measures 0.91 or 1.47 ms per step on the same box)
    e2e_runs = []
    for _ in range(3):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pipelined(args.steps)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e2e_total = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([e2e_total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_total = float(t.item())
        e2e_runs.append(e2e_total / args.steps)
    e2e_ms = statistics.median(e2e_runs)
    clocks = sampler.stop() if rank == 0 else None
    h2d = 2 * nl * d * 4
    d2h = 2 * nl * d * 4 + 4

    # ---- roofline of the dominant kernel (backward), timed alone through the C-ABI; single GPU shape only
    roof = None
    fwd_info = None
    bwd_info = None
    base = None
    if rank == 0:
        flags = _flags(prec)
        if sharded:
            xa, ya = make_inputs(N, d, wl["seed"], dev)
        else:
            xa, ya = x.detach(), y.detach()
        n = 2 * N
        ws = Workspace(N, N, d, flags, dev)
        loss_t = torch.empty((), device=dev)
        stats = torch.empty(8, device=dev)
        gout = torch.ones((), device=dev)
        dz = torch.empty(int(lib.edrl_mmd_grad_slabs(N, N, d, flags, n, 0)) * n, d, device=dev)
        st = _lib.stream_and_device(xa)

        def fwd_only():
            _lib.check(lib.edrl_mmd_forward(xa.data_ptr(), ya.data_ptr(), N, N, d, 2.0, 5, flags, 0, 1,
                                            loss_t.data_ptr(), stats.data_ptr(), None, ws.ptr, ws.nbytes, st))

        def bwd_only():
            _lib.check(lib.edrl_mmd_backward(N, N, d, 2.0, 5, flags, stats.data_ptr(), gout.data_ptr(), 0, n,
                                             dz.data_ptr(), ws.ptr, ws.nbytes, st))

        def fused_only():
            _lib.check(lib.edrl_mmd_forward_grad(xa.data_ptr(), ya.data_ptr(), N, N, d, 2.0, 5, flags, 0, n, 0, 0, 1,
                                                 loss_t.data_ptr(), stats.data_ptr(), None, dz.data_ptr(), ws.ptr,
                                                 ws.nbytes, st))

        reps = 3 if sharded else max(5, args.steps)
        fwd_only()
        f_ms = timed_steps(fwd_only, reps, 2, flush, 1) / reps
        b_ms = timed_steps(bwd_only, reps, 2, flush, 1) / reps
        mma_per_product = 3 if prec == "3xtf32" else 1
        peak = peaks["bf16_burst"] / 2.0          # kind::tf32 issues at half the bf16 rate
        flops_b = 2.0 * n * n * d                 # G.Z: the algorithmic backward contraction
        flops_f = 1.0 * n * n * d                 # unique Gram entries n(n+1)/2 x 2d
        # Gram sweeps per step: one per 512-column feature pass (pair kernel), per 1024-column pass for d > 768 (quad kernel)
        passes = math.ceil(d / 1024) if (d > 768 and prec != "3xtf32") else math.ceil(d / 512)
        ach_b = flops_b / (b_ms * 1e-3) / 1e12
        ach_f = flops_f / (f_ms * 1e-3) / 1e12
        fwd_info = {"kernel": "prep + mmd_fwd_pair_kernel (loss only, e.g. under no_grad)", "ms": f_ms,
                    "achieved": ach_f, "frac": ach_f / peak, "algorithmic_flops": flops_f}
        bwd_info = {"kernel": "edrl_mmd_backward: the sweep kernel + apply_grad in place", "ms": b_ms,
                    "achieved": ach_b, "frac": ach_b / peak, "algorithmic_flops": flops_b}
        if prec in ("tf32", "tf32h", "f16s", "3xtf32"):
            # the training step's dominant launch: forward sums + gradient in one sweep over the Gram tiles
            g_ms = timed_steps(fused_only, reps, 2, flush, 1) / reps
            flops_g = flops_f + flops_b
            ach = flops_g / (g_ms * 1e-3) / 1e12
            # which measured peak: the burst figure for a kernel timed alone for a millisecond, the sustained one (cuBLAS
            # back to back for seconds: the power cap has pulled the clocks down) when one launch runs for tens of ms
            peak_kind = "sustained" if g_ms >= 20.0 else "burst"
            bf16_peak = peaks["bf16_" + peak_kind]
            peak = bf16_peak / 2.0
            if prec == "tf32h":
                # Gram at the kind::tf32 rate, G.Z at the kind::f16 rate (twice as fast): blended peak for 1 : 2 work
                peak = flops_g / (flops_f / peak + flops_b / (2.0 * peak))
            elif prec == "f16s":
                peak = bf16_peak                      # both contractions issue kind::f16 MMAs
            if prec == "3xtf32":
                peak = peak / 3.0                     # three TF32 MMAs per product
            mode_id = {"tf32": 0, "tf32h": 1, "f16s": 2, "3xtf32": 3}[prec]
            roof = {"bound": "tensor",
                    "kernel": (f"mmd_sweep_quad_kernel<FAST, MODE={mode_id}>" if (d > 768 and prec != "3xtf32") else f"mmd_sweep256_kernel<FAST, MODE={mode_id}>")
                              + " (forward sums + gradient, one persistent Gram sweep)",
                    "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": (profiled_traffic(f"mmd_sweep256_kernel<1, {mode_id}>") if (N, d) == (8192, 512) else None),
                    "traffic_note": "DRAM bytes per launch (ncu, profiles/); algorithmic HBM bytes are 3 n d 4 = 96 MiB "
                                    "(read Z and Z^T, write U) -- the kernel is bound by L2 -> SM traffic, not DRAM",
                    "ms": g_ms,
                    "ms_includes": "the two prep kernels (~0.06 ms at N=8192) launched by the same C-ABI call",
                    "peak_source": {
                        "tf32": f"{peaks['source']} bf16 {peak_kind} {bf16_peak} TF/s / 2 (TF32 rate)",
                        "tf32h": f"{peaks['source']} bf16 {peak_kind} {bf16_peak} TF/s: Gram (1/3 of the work) at the "
                                 "TF32 rate (/2), G.Z (2/3) at the f16 rate",
                        "f16s": f"{peaks['source']} bf16 {peak_kind} {bf16_peak} TF/s (kind::f16 MMAs)",
                        "3xtf32": f"{peaks['source']} bf16 {peak_kind} {bf16_peak} TF/s / 2 (TF32 rate) / 3 (MMAs per product)"}[prec]
                                   + ("; sustained because one launch runs %.0f ms (>= 20 ms)" % g_ms if peak_kind == "sustained" else ""),
                    "frac_of_burst_peak": ach / (peak * peaks["bf16_burst"] / bf16_peak),
                    "algorithmic_flops": flops_g, "mma_per_product": 1,
                    "executed_tensor_flops": (2.0 * n * n * d * passes + 2.0 * n * n * d)}
        else:
            roof = {"bound": "tensor", "kernel": "edrl_mmd_backward (sweep + apply_grad in place)", "achieved": ach_b, "peak": peak,
                    "unit": "TFLOP/s", "frac": ach_b / peak, "traffic": None, "ms": b_ms,
                    "peak_source": f"{peaks['source']} bf16 burst {peaks['bf16_burst']} TF/s / 2 (TF32 rate)",
                    "algorithmic_flops": flops_b, "mma_per_product": mma_per_product,
                    "executed_tensor_flops": (2.0 * n * n * d * math.ceil(d / 256) + 2.0 * n * n * d) * 3}
    # ---- N > 1: the same workload on ONE GPU through the same public API (the anchor of the scaling curve), and the
    #      sharded result checked against it in this very run (the NCCL parity test needs >= 2 GPUs, the test box has 1)
    parity = None
    if sharded:
        gr = torch.Generator().manual_seed(7)
        rows_l = torch.randint(0, nl, (256,), generator=gr).to(dev)        # the same local row numbers on every rank
        x.grad = None
        y.grad = None
        l_sh = loss_fn(x, y)
        l_sh.backward()
        torch.cuda.synchronize()
        mine = torch.cat([x.grad[rows_l], y.grad[rows_l]]).contiguous()     # [512, d]
        my_loss = l_sh.detach()
        gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, gathered, dst=0)
        if rank == 0:
            xa.requires_grad_(True)
            ya.requires_grad_(True)

            def single_step():
                xa.grad = None
                ya.grad = None
                edrl_b200.MK_MMD(xa, ya, precision=prec).backward()

            single_step()
            one = timed_steps(single_step, 3, 1, flush, 1) / 3
            with torch.no_grad():
                single_loss = edrl_b200.MK_MMD(xa, ya, precision=prec).item()
            base = {"n_gpus": 1, "ms_per_step": one, "value": N / (one * 1e-3),
                    "note": "same workload, unsharded, on rank 0 alone through the same public API "
                            "(edrl_b200.MK_MMD + backward, apply_grad included)"}
            gmax = max(xa.grad.abs().max().item(), ya.grad.abs().max().item())
            gerr = 0.0
            for rk in range(world):
                ref_rows = torch.cat([xa.grad[rk * nl + rows_l], ya.grad[rk * nl + rows_l]])
                gerr = max(gerr, (gathered[rk] - ref_rows).abs().max().item())
            parity = {"loss_sharded": my_loss.item(), "loss_single_gpu": single_loss,
                      "loss_rel_err": abs(my_loss.item() - single_loss) / abs(single_loss),
                      "grad_max_err_over_gmax": gerr / gmax, "rows_sampled_per_rank": 512,
                      "what": "sharded loss and 512 sampled gradient rows of every rank against the single-GPU evaluation of "
                              "the same inputs in the same run (same kernels, different work lists: agreement to fp32 "
                              "summation order); the single-GPU kernel itself is checked against the fp64 row-blocked "
                              "oracle at this size in tests/test_gpu_mmd.py",
                      "ok": bool(abs(my_loss.item() - single_loss) <= 1e-5 * abs(single_loss) and gerr <= 1e-4 * gmax)}
            xa.requires_grad_(False)
            ya.requires_grad_(False)
    if world > 1:
        dist.barrier()

    if rank == 0:
        if args.no_cpu_baseline:
            cpu_v, cores, sample, cpu_kind = None, 0, "skipped (--no-cpu-baseline, profiling run)", "port"
        elif not sharded:
            t_cpu, done, cores, cpu_kind = cpu_whole_steps(N, d, wl["seed"], 2, 30.0)
            cpu_v = N / t_cpu
            sample = (f"{done} whole fwd+bwd step(s) of the reference MK_MMD (code/MMD.py:46-74 + autograd) at N={N} per side, "
                      f"d={d}: {t_cpu:.2f} s per step on {cores} host threads")
        else:
            from oracle import cpu_port
            cores, cpu_kind = os.cpu_count() or 1, "port"
            torch.set_num_threads(cores)
            g = torch.Generator().manual_seed(wl["seed"])
            z = torch.cat([torch.randn(N, d, generator=g), torch.randn(N, d, generator=g) * 1.25 + 0.1])
            cpu_port.rowblock_fwd_bwd(z, N, 0, 128)
            ts = []
            for i in range(5):
                t0 = time.perf_counter()
                cpu_port.rowblock_fwd_bwd(z, N, 128 * i, 128)
                ts.append(time.perf_counter() - t0)
            cpu_v = N / (statistics.median(ts) * (2 * N / 128))
            sample = (f"EXTRAPOLATED: rows [r0, r0+128) of the {2 * N}x{2 * N} problem (d={d}) fwd+bwd with torch CPU ops, median "
                      f"of 5 blocks x{2 * N // 128}; the reference cannot run this size on any single device")
        extras = {} if (args.no_extras or sharded) else essence_point_extras(peaks)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": prec, "data": "synthetic",
            "config": config_of(wl, world, prec),
            "timing": {"precision": prec, "l2": "256 MiB memset between steps (untimed); per-step CUDA events",
                       "host_cpus_bound": bound_cpus},
            "clocks": clocks,
            "e2e": {"value": N / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "how": "edrl_b200.MK_MMD + backward per step from pinned host buffers; H2D(X,Y) / compute / "
                           "D2H(loss,dX,dY) software-pipelined over the steps on three streams; one pair of CUDA "
                           "events around all K steps; median of three such runs",
                    "ms_per_step_runs": e2e_runs},
            "e2e_serial": {"value": N / (e2e_serial_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_serial_ms,
                           "how": "same step, copies and compute back to back on one stream, per-step events"},
            "gpu_launches": int(launches),
            "roofline": roof, "roofline_fwd": fwd_info, "roofline_bwd_separate": bwd_info,
            "cpu_baseline": {"value": cpu_v, "unit": UNIT, "cores": cores, "kind": cpu_kind, "sample": sample},
        }
        if roof is not None:
            roof["part_a"] = {k: roof[k] for k in ("bound", "kernel", "achieved", "peak", "unit", "frac", "traffic", "ms")
                              if k in roof}
            if not sharded:
                try:
                    roof["part_b_select"] = select_roofline(peaks)
                except Exception as exc:
                    roof["part_b_select"] = {"error": repr(exc)}
        if parity:
            line["parity"] = parity
        if base:
            line["strong_scaling_base"] = base
            line["note"] = ("N>1 runs configs[3] (fixed total work, strong scaling); the 1-GPU time of the SAME workload through "
                            "the same API is strong_scaling_base (and `scale_anchor` of the N=1 line), while `bench.py --gpus 1` "
                            "runs configs[1] (N=8192, d=512) as the contract asks: per-N values compare with "
                            "strong_scaling_base, not with the N=1 line's value")
        if extras:
            line["essence_point"] = extras
            line["sweep_vs_torch_gpu"] = sweep_vs_torch_gpu(prec)
            # the same step in the other precision modes, same timing recipe
            def mode_roofline(o):
                """fused C-ABI call (prep + sweep) alone in precision mode o: ms, achieved TFLOP/s, its own peak"""
                fl = _flags(o)
                ws_o = Workspace(N, N, d, fl, dev)
                nn = 2 * N
                u_o = torch.empty(int(lib.edrl_mmd_grad_slabs(N, N, d, fl, nn, 0)) * nn, d, device=dev)
                lo, so = torch.empty((), device=dev), torch.empty(8, device=dev)
                xo, yo = x.detach(), y.detach()
                sto = _lib.stream_and_device(xo)

                def call():
                    _lib.check(lib.edrl_mmd_forward_grad(xo.data_ptr(), yo.data_ptr(), N, N, d, 2.0, 5, fl, 0, nn, 0, 0, 1,
                                                         lo.data_ptr(), so.data_ptr(), None, u_o.data_ptr(), ws_o.ptr,
                                                         ws_o.nbytes, sto))
                t = timed_steps(call, 10, 3, flush, 1) / 10
                tf32_peak = peaks["bf16_burst"] / 2.0
                fa = 3.0 * nn * nn * d
                pk = {"tf32": tf32_peak, "tf32h": 3.0 / (1.0 / tf32_peak + 2.0 / peaks["bf16_burst"]),
                      "f16s": peaks["bf16_burst"], "3xtf32": tf32_peak / 3.0}[o]      # three TF32 MMAs per product
                return {"fused_call_ms": t, "achieved_tflops": fa / (t * 1e-3) / 1e12, "peak_tflops": pk,
                        "frac": fa / (t * 1e-3) / 1e12 / pk}

            line["precision_modes"] = {prec: dict(ms_per_step=ms_per_step, value=value, **mode_roofline(prec))}
            for other in ("tf32", "tf32h", "f16s", "3xtf32"):
                if other == prec:
                    continue

                def other_step(o=other):
                    x.grad = None
                    y.grad = None
                    edrl_b200.MK_MMD(x, y, precision=o).backward()

                t = timed_steps(other_step, 5, 3, flush, 1) / 5
                line["precision_modes"][other] = dict(ms_per_step=t, value=N / (t * 1e-3), **mode_roofline(other))
            line["precision_modes"]["note"] = (
                "tf32: TF32 Gram and TF32 G.Z (headline). tf32h: same TF32 Gram, G.Z operands stored as scaled "
                "binary16 with the same 11-bit significands (gradients agree with tf32 to 2e-5 |g|_inf). f16s: the Gram "
                "too reads a scaled binary16 copy of the TF32-rounded operand (identical significands, exact products, "
                "fp32 accumulation; agrees with tf32 to 2e-6 on the loss and 5e-5 |g|_inf on gradients). 3xtf32: hi/lo "
                "split, three MMAs per product, fp32-level accuracy, fused sweep (MODE 3; beyond d = 768 one Gram pass per 512 columns).")
            line["essence_path_vs_torch_gpu"] = essence_path_vs_torch_gpu()
            line["eval_missing_modality"] = eval_missing_modality()
            # the configs[3] workload (N = 65536 per side, d = 1024) on this one GPU through the same public API: the anchor
            # the N > 1 lines' values compare with (the driver's own curve divides by THIS line's `value`, which is the
            # configs[1] workload, as the contract asks)
            try:
                wa = WORKLOADS["sharded65536"]
                xs_, ys_ = make_inputs(wa["N"], wa["d"], wa["seed"], dev)
                xs_.requires_grad_(True)
                ys_.requires_grad_(True)

                def anchor_step():
                    xs_.grad = None
                    ys_.grad = None
                    edrl_b200.MK_MMD(xs_, ys_, precision=prec).backward()

                anchor_step()
                t = timed_steps(anchor_step, 3, 1, flush, 1) / 3
                line["scale_anchor"] = {"workload": wa["name"], "n_gpus": 1, "ms_per_step": t, "value": wa["N"] / (t * 1e-3),
                                        "note": "N > 1 lines run this workload sharded: compare their `value` with this one"}
                del xs_, ys_
                torch.cuda.empty_cache()
            except Exception as exc:
                line["scale_anchor"] = {"error": repr(exc)}
            if not args.no_drivers:
                line["reference_drivers"] = reference_driver_runs()
            line["next_rows"] = next_rows_extras(peaks)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + list(WORKLOADS))
    ap.add_argument("--precision", default="tf32", choices=["tf32", "tf32h", "f16s", "3xtf32"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-drivers", action="store_true", help="skip the reference-driver runs (configs[0], [2], [4]) of the extras")
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="--impl reference: wall-clock budget of the timed steps")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="profiling runs only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank, local_rank, world = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
