for v in 20_1024 32_512; do AIGW_B200_SO=/root/repo/variants/libbpe_$v.so timeout 600 python bench.py --config 3 --steps 3 --warmup 3 --skip-e2e 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'): d=json.loads(l); print('$v', d['value'], d['ms_per_step'])"; done
