"""Turn the ncu reports / launch list under gpurun_out/ into the tracked summaries under profiles/
(run in the build container: ncu can read reports without a GPU)."""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r01"

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum.per_second", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__cluster_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__cycles_elapsed.max", "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second"]


def raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for i, h in enumerate(hdr):
            if h in WANT:
                d[h] = f"{r[i]} {units[i]}".strip()
        out.append(d)
    return txt, out


def main():
    os.makedirs(OUT, exist_ok=True)
    summary = {}
    traffic = {}
    for name in ("prof_mmd_final", "prof_sweep_f16s", "prof_sweep_quad", "prof_topk_final"):
        rep = os.path.join(SRC, name + ".ncu-rep")
        if not os.path.isfile(rep):
            continue
        txt, ks = raw(rep)
        with open(os.path.join(OUT, f"{TAG}_{name}_raw.csv"), "w") as f:
            f.write(txt)
        summary[name] = ks
        for k in ks:
            try:
                rd = float(k["dram__bytes_read.sum"].split()[0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[k["dram__bytes_read.sum"].split()[1]]
                wr = float(k["dram__bytes_write.sum"].split()[0]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[k["dram__bytes_write.sum"].split()[1]]
                short = k["kernel"].split("(")[0].replace("void ", "")
                traffic[short] = rd + wr
            except Exception:
                pass
    ll = os.path.join(SRC, "launches.csv")
    if os.path.isfile(ll):
        lines = [l for l in open(ll) if not l.startswith("==")]
        with open(os.path.join(OUT, f"{TAG}_launches_bench.csv"), "w") as f:
            f.writelines(lines)
        agg = collections.OrderedDict()
        for row in csv.DictReader(lines):
            nm = row["Kernel Name"].split("(")[0][:70]
            v = float(row["Metric Value"].replace(",", ""))
            a = agg.setdefault(nm, [0, 0.0])
            a[0] += 1
            a[1] += v
        tot = sum(a[1] for a in agg.values())
        summary["launch_list"] = [{"kernel": k, "launches": c, "avg_us": t / c / 1e3, "share_pct": 100 * t / tot}
                                  for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]
    with open(os.path.join(OUT, f"{TAG}_summary.json"), "w") as f:
        json.dump(summary, f, indent=1)
    with open(os.path.join(OUT, "traffic.json"), "w") as f:
        json.dump({"source": f"ncu --set full, {TAG}", "dram_bytes_per_launch": traffic}, f, indent=1)
    print(json.dumps(summary, indent=1)[:6000])


if __name__ == "__main__":
    main()
