#!/usr/bin/env python
"""profiles/rNN_traffic.json from one `ncu --set full` capture of the translate pass.

    python tools/ncu_traffic.py gpurun_out/chat_r02.ncu-rep --bodies 200000 --round 2

Reads dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum / smsp__inst_executed.sum per launch from the report
(`ncu -i … --page raw --csv`), groups the launches by kernel, and writes the per-body DRAM bytes of the stages of ONE translate
pass.  bench.py multiplies bytes_per_body by the bodies of a step for roofline.traffic (the capture itself runs on a smaller
batch: replaying 1 M bodies under --set full takes too long), which is valid because every stage streams: none of the three
re-reads across bodies, so bytes per body do not depend on the batch once it exceeds L2 (200 k bodies = 190 MB in).
"""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    for r in rd[2:]:
        yield {h: (v, u) for h, v, u in zip(hdr, r, units)}


SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1, "s": 1, "inst": 1, "": 1}


def num(cell):
    v, u = cell
    v = float(v.replace(",", ""))
    return v * SCALE.get(u, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--bodies", type=int, required=True, help="bodies per translate pass in the captured command")
    ap.add_argument("--round", type=int, default=2)
    ap.add_argument("--command", default="")
    a = ap.parse_args()
    per = {}
    for r in rows(a.report):
        m = re.search(r"\b([A-Za-z_]\w*)\s*(<|\()", re.sub(r"^void\s+", "", r["Kernel Name"][0]))
        name = m.group(1) if m else r["Kernel Name"][0]
        k = per.setdefault(name, {"launches": 0, "dram_read": 0.0, "dram_write": 0.0, "time_s": 0.0, "inst": 0.0})
        k["launches"] += 1
        k["dram_read"] += num(r["dram__bytes_read.sum"])
        k["dram_write"] += num(r["dram__bytes_write.sum"])
        k["time_s"] += num(r["gpu__time_duration.sum"])
        k["inst"] += num(r["smsp__inst_executed.sum"])
    stages = {}
    passes = None
    for name, k in per.items():
        if not name.startswith("chat_"):
            continue
        n = k["launches"]
        passes = n if passes is None else min(passes, n)
        stages[name] = {"launches_captured": n, "dram_read_per_body": k["dram_read"] / n / a.bodies, "dram_write_per_body": k["dram_write"] / n / a.bodies,
                        "time_ms_per_launch_under_ncu": k["time_s"] / n * 1e3, "warp_inst_per_body": k["inst"] / n / a.bodies}
    total = sum(s["dram_read_per_body"] + s["dram_write_per_body"] for s in stages.values())
    out = {"bytes_per_body": total, "bodies_in_capture": a.bodies, "stages": stages,
           "source": f"profiles/r{a.round:02d}_traffic.json <- ncu --set full --clock-control none, {os.path.basename(a.report)} ({a.bodies} bodies per pass"
                     + (f"; {a.command}" if a.command else "") + "); dram__bytes_read.sum + dram__bytes_write.sum per launch, per body"}
    p = os.path.join(ROOT, "profiles", f"r{a.round:02d}_traffic.json")
    json.dump(out, open(p, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    sys.exit(main())
