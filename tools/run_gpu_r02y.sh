AIGW_SMALL_DBG=1 timeout 300 python - <<'PY' 2>&1 | tail -12
import sys, os
sys.path.insert(0, "tests")
import aigw_b200 as A, _workload as W
ctx = A.Context(0); cfg = ctx.cfg("aws-bedrock")
arena, offs, lens = W.chat_corpus(2, 0, 512)
for n in (1, 1, 1, 1, 64, 64, 512, 512, 1, 1):
    ctx.chat_translate_host(cfg, arena[: int(offs[n]) + 16], offs[: n + 1].copy(), lens[:n].copy())
PY
