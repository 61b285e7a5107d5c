#!/usr/bin/env python
"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/<tag>_*.ncu-rep, <tag>_launches.csv) into the small
text summaries committed under profiles/.  Runs on the CPU box.

    python scripts/summarize_profile.py r01
"""
import collections
import csv
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
OUT = ROOT / "profiles"
SRC = ROOT / "gpurun_out"

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__sass_average_branch_targets_threads_uniform.pct", "smsp__cycles_active.avg",
]


def launches(tag):
    f = SRC / f"{tag}_launches.csv"
    if not f.exists():
        return ""
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    rows = rows[rows.index(hdr) + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        name = r[ki].split("(")[0].replace("void ", "").replace("sffg::<unnamed>::", "")[:70]
        v = float(r[vi].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    out = [f"# {tag}: ncu launch list of `python bench.py --steps 5 --warmup 3 --no-cpu --no-extra` (gpu__time_duration.sum, --clock-control none)",
           "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
           "# the timed region of `value` launches collide_poses_kernel only (1 launch per step = 100 % of the step);",
           "# the remaining collide launches are the chunked e2e leg (16 chunks per host call), the check_edges launches are the",
           "# edge e2e leg (4 chunks per host call); the secondary `extra` block (edges, k-NN, planner solves) is switched off", ""]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"{t / 1e6:10.3f} ms  {n:4d} launches  {100 * t / tot:5.1f} %  {k}")
    return "\n".join(out) + "\n"


def raw(rep):
    txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    out = [f"kernel: {d.get('Kernel Name', ('?', ''))[0]}"]
    for k in KEYS:
        if k in d:
            out.append(f"{k:75s} {d[k][0]} {d[k][1]}")
    stalls = sorted(((float(v[0]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")) for h, v in d.items()
                     if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued") and v[0]),
                    reverse=True)
    tot = sum(s for s, _ in stalls) or 1
    out.append("stall samples: " + ", ".join(f"{n} {100 * s / tot:.1f}%" for s, n in stalls[:8]))
    return "\n".join(out) + "\n"


def source_sha16(files):
    """hash of the kernel sources a counter file belongs to: bench.py refuses counters of an older kernel"""
    import hashlib
    h = hashlib.sha256()
    for f in files:
        h.update((ROOT / "space_filling_forest_star_b200" / "csrc" / f).read_bytes())
    return h.hexdigest()[:16]


COLLIDE_SOURCES = ["collide_kernels.cu", "collide_kernels.cuh", "common.h"]
KNN_SOURCES = ["knn_pruned.cu", "knn_common.cuh", "knn_kernels.cu"]


def counters(tag):
    """profiles/traffic.json (collide) and profiles/knn_counters.json: the per-launch ncu counters bench.py turns into the
    issue-rate roofline, stamped with the hash of the kernel sources they were captured from"""
    import json
    for name, out, files, extra in (("collide", "traffic.json", COLLIDE_SOURCES, {"kernel": "collide_poses_kernel<f32>", "poses_per_launch": 1 << 24}),
                                    ("knn", "knn_counters.json", KNN_SOURCES, {"kernel": "knn_pruned_kernel<6,8,1>", "config": "N=1000000 Q=100000 k=16"})):
        rep = SRC / f"{tag}_{name}.ncu-rep"
        if not rep.exists():
            continue
        txt = subprocess.run(["ncu", "-i", str(rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        d = dict(zip(rows[0], rows[2]))
        unit = dict(zip(rows[0], rows[1]))
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rec = dict(extra)
        rec.update({
            "dram_bytes_read": float(d["dram__bytes_read.sum"]) * scale.get(unit["dram__bytes_read.sum"], 1.0),
            "dram_bytes_write": float(d["dram__bytes_write.sum"]) * scale.get(unit["dram__bytes_write.sum"], 1.0),
            "warp_instructions": float(d["smsp__inst_executed.sum"]),
            "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
            "active_lanes_per_inst": float(d["smsp__thread_inst_executed_per_inst_executed.ratio"]),
            "source": f"profiles/{tag}_{name}.txt (ncu --set full --clock-control none, one launch)",
            "source_sha16": source_sha16(files),
        })
        (OUT / out).write_text(json.dumps(rec, indent=1))
        print("wrote", OUT / out)


def main():
    tag = sys.argv[1]
    OUT.mkdir(exist_ok=True)
    txt = launches(tag)
    if txt:
        (OUT / f"{tag}_launches.txt").write_text(txt)
    kernels = {"collide": "collide_poses_kernelILi0ELb0", "knn": "knn_pruned_kernel", "edges": "check_edges_kernelILb0"}
    for name, key in kernels.items():
        rep = SRC / f"{tag}_{name}.ncu-rep"
        if not rep.exists():
            continue
        body = f"# {tag}: ncu --set full --clock-control none, one launch of the {name} kernel\n\n" + raw(rep)
        lines = subprocess.run([sys.executable, str(ROOT / "scripts" / "ncu_lines.py"), str(rep), key, "--top", "25"],
                               capture_output=True, text=True).stdout
        body += "\n# hottest source lines (share of stall samples / of executed warp instructions / avg active lanes)\n" + lines
        (OUT / f"{tag}_{name}.txt").write_text(body)
        print("wrote", OUT / f"{tag}_{name}.txt")
    counters(tag)


if __name__ == "__main__":
    main()
