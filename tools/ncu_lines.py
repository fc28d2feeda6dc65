#!/usr/bin/env python3
"""Join an ncu SASS source page (csv) with nvdisasm -g line info: per-source-line instruction and sample shares.
usage: tools/ncu_lines.py <src.csv from `ncu --page source --csv`> <nvdisasm -g -c output> <kernel substring> [top N]"""
import csv, re, sys, collections
srccsv, dis, want = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
fn = None; cur = None; mapping = {}
for l in open(dis):
    m = re.match(r'\s*\.section\s+\.text\.(\S+),', l)
    if m: fn = m.group(1); mapping[fn] = {}; cur = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);', l)
    if m and fn: mapping[fn][int(m.group(1), 16)] = cur
rows = list(csv.reader(open(srccsv)))
# split per kernel section; take the first section whose name contains `want`
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sec = next(i for i in secs if want in rows[i][1])
end = next((j for j in secs if j > sec), len(rows))
h = rows[sec + 1]; data = rows[sec + 2:end]
iI, iS, iA, iT = h.index('Instructions Executed'), h.index('# Samples'), h.index('Address'), h.index('Thread Instructions Executed')
key = next(k for k in mapping if want.replace("aigw::", "") .split("<")[0] in k and ("5120" in k or "walk" in k or len(mapping) == 1))
offs = mapping[key]
base = int(data[0][iA], 16)
agg = collections.defaultdict(lambda: [0, 0, 0]); tot = [0, 0, 0]
for r in data:
    try: off = int(r[iA], 16) - base; n = int(r[iI]); s = int(r[iS]); t = int(r[iT])
    except Exception: continue
    c = offs.get(off) or ('?', 0)
    agg[c][0] += n; agg[c][1] += s; agg[c][2] += t
    tot[0] += n; tot[1] += s; tot[2] += t
src = {}
import os
for f in ('chat_kernel.cu', 'chat_kernel.cuh', 'chat_walk.cu', 'chat_walk_impl.cuh', 'classify.cuh', 'chat_internal.cuh', 'bulk.cuh', 'sse_kernel.cu', 'tjson.cuh', 'stream_kernel.cu'):
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'aigw_b200', 'csrc', f)
    if os.path.exists(p): src[f] = open(p).read().split('\n')
print(f"{want}: {tot[0]} warp instructions, {tot[1]} samples, avg active threads {tot[2] / max(1, tot[0]):.1f}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    line = src.get(k[0], [''])[k[1] - 1][:95].strip() if k[0] in src and k[1] > 0 else ''
    print(f"{100 * v[1] / tot[1]:5.1f}% time {100 * v[0] / tot[0]:5.1f}% inst thr {v[2] / max(1, v[0]):4.1f}  {k[0]}:{k[1]:<5} {line}")
