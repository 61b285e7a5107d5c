#!/usr/bin/env python
"""Attribute an ncu capture to CUDA source lines (runs on the CPU box, no GPU needed).

    python scripts/ncu_lines.py gpurun_out/r01_collide.ncu-rep collide_poses_kernelILb0ELb0 [--top 40]

ncu's CSV source page is SASS-only, so the SASS offsets are joined with the line table nvdisasm prints for the
cubin embedded in libsffg.so (built with -lineinfo).  Output: per source line, share of warp-stall samples and of
executed warp instructions, average active threads, and the dominant stall reasons.
"""
import argparse
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
LIB = ROOT / "space_filling_forest_star_b200" / "libsffg.so"


def line_table(kernel_key: str):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", str(LIB)], cwd=tmp, capture_output=True)
    table = {}
    for cubin in Path(tmp).glob("*.cubin"):
        out = subprocess.run(["nvdisasm", "-gi", str(cubin)], capture_output=True, text=True).stdout
        cur_fn, cur_line, active, pending = None, None, False, []
        for ln in out.splitlines():
            m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
            if m:
                cur_fn = m.group(1)
                active = kernel_key in cur_fn
                continue
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                pending.append((m.group(1), int(m.group(2))))   # innermost location first, call sites after it
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
            if m:
                if pending:
                    own = [p for p in pending if "/csrc/" in p[0]]
                    pick = own[0] if own else pending[0]
                    cur_line = (Path(pick[0]).name, pick[1])
                    pending = []
                if cur_line:
                    table[int(m.group(1), 16)] = cur_line
        if table:
            break
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("kernel_key")
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    table = line_table(a.kernel_key)
    if not table:
        sys.exit("kernel not found in cubin line tables")
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Address")
    body = rows[rows.index(hdr) + 1:]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_")]
    base = min(int(r[0], 16) for r in body if r and r[0].startswith("0x"))
    agg = defaultdict(lambda: defaultdict(float))
    for r in body:
        if not r or not r[0].startswith("0x"):
            continue
        off = int(r[0], 16) - base
        key = table.get(off, ("?", 0))
        g = agg[key]
        g["samples"] += float(r[col["# Samples"]] or 0)
        g["inst"] += float(r[col["Instructions Executed"]] or 0)
        g["tinst"] += float(r[col["Thread Instructions Executed"]] or 0)
        for s in stall_cols:
            g[s] += float(r[col[s]] or 0)
    ts = sum(g["samples"] for g in agg.values()) or 1
    ti = sum(g["inst"] for g in agg.values()) or 1
    src_cache = {}

    def src(fn, n):
        if fn not in src_cache:
            cand = list((ROOT / "space_filling_forest_star_b200" / "csrc").glob(fn))
            src_cache[fn] = cand[0].read_text().splitlines() if cand else []
        L = src_cache[fn]
        return L[n - 1].strip()[:100] if 0 < n <= len(L) else ""

    print(f"total samples {ts:.0f}, warp instructions {ti:.0f}")
    for key, g in sorted(agg.items(), key=lambda kv: -kv[1]["samples"])[: a.top]:
        stalls = sorted(((g[s], s[6:]) for s in stall_cols if g[s] > 0), reverse=True)[:3]
        st = " ".join(f"{n}:{100 * v / max(g['samples'], 1):.0f}%" for v, n in stalls)
        act = g["tinst"] / g["inst"] if g["inst"] else 0
        print(f"{key[0]}:{key[1]:<5d} smp {100 * g['samples'] / ts:5.1f}%  inst {100 * g['inst'] / ti:5.1f}%  act {act:4.1f}  [{st}]  {src(*key)}")


if __name__ == "__main__":
    main()
