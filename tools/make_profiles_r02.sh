#!/bin/bash
# profiles/r02_*: summaries of the final captures of the round (run here, no GPU needed: reads gpurun_out/*.ncu-rep, *.csv, bench JSONs)
set -e
cd "$(dirname "$0")/.."
REP=gpurun_out/chat_r02_final.ncu-rep
T=/tmp/aigw_prof; rm -rf $T; mkdir -p $T/elf
(cd $T/elf && cuobjdump -xelf all $OLDPWD/aigw_b200/libaigw_b200.so > /dev/null && nvdisasm -g -c chat_kernel.sm_100a.cubin > ck.dis 2>/dev/null && nvdisasm -g -c chat_walk_g0.sm_100a.cubin > w0.dis 2>/dev/null)
ncu -i $REP --page source --csv > $T/src.csv 2>/dev/null
python tools/ncu_traffic.py $REP --bodies 200000 --round 2 --command "python bench.py --steps 1 --warmup 1 --bodies 200000 --skip-e2e" > /dev/null
{
  echo "# Round 2 — chat translate pass (index / sort / walk / emit), final state: ncu --set full --clock-control none"
  echo
  echo "Command: \`ncu --set full --clock-control none --import-source on -k regex:chat_ -s 4 -c 4 python bench.py --steps 1 --warmup 1 --bodies 200000 --skip-e2e\` (tools/run_gpu_r03f.sh; 200 k bodies of the configs[1] corpus, resident inputs)."
  echo "Times under ncu are cold-cache and serialised; the bench numbers (CUDA events) are in profiles/r02_bench_*.json.  Per-body DRAM traffic: profiles/r02_traffic.json."
  echo
  python tools/ncu_summary.py $REP
  echo; echo "## hottest lines (SASS joined to source lines, tools/ncu_lines.py)"; echo; echo '```'
  python tools/ncu_lines.py $T/src.csv $T/elf/ck.dis chat_index_kernel 25; echo
  python tools/ncu_lines.py $T/src.csv $T/elf/w0.dis chat_walk_kernel_g0 25; echo
  python tools/ncu_lines.py $T/src.csv $T/elf/ck.dis chat_emit_kernel 25
  echo '```'
} > profiles/r02_chat.md 2>/dev/null
cp gpurun_out/launches_r02_final.csv profiles/r02_launches_chat.csv
if [ -f gpurun_out/small_r02_final.ncu-rep ]; then
  { echo "# Round 2 — fused small-batch kernel (chat_small_kernel_g0), ncu --set full, single 4 KB body (tools/lat_probe.py)"; echo; python tools/ncu_summary.py gpurun_out/small_r02_final.ncu-rep; } > profiles/r02_small_kernel.md 2>/dev/null
fi
for f in final final_ref final_c1 final_c4 final_c5; do [ -s gpurun_out/bench_r02_$f.json ] && grep '^{' gpurun_out/bench_r02_$f.json | tail -1 > profiles/r02_bench_$f.json; done
[ -s gpurun_out/bench_r02_c3.json ] && grep '^{' gpurun_out/bench_r02_c3.json | tail -1 > profiles/r02_bench_c3.json
[ -s gpurun_out/bench_r02_n2.json ] && grep '^{' gpurun_out/bench_r02_n2.json | tail -1 > profiles/r02_bench_n2.json
ls -la profiles/
