#!/usr/bin/env python
"""p50 / p99 of one chat_translate_host call for small batches (1, 8, 64, 256, 512 bodies) through the fused small-batch
kernel and through the throughput pipeline.  usage: python tools/lat_probe.py"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import aigw_b200 as A
import _workload as W

ctx = A.Context(0)
cfg = ctx.cfg("aws-bedrock")
arena, offs, lens = W.chat_corpus(2, 0, 512)
out = {}
for mode, sm in (("small", 1 << 20), ("pipeline", 0)):
    ctx.chat_set_small_batch(sm)
    for n in (1, 8, 64, 256, 512):
        a, o, l = arena[: int(offs[n]) + 16], offs[: n + 1].copy(), lens[:n].copy()
        lat = []
        import ctypes as C
        from aigw_b200.capi import _BatchOut
        bo = _BatchOut(); fn = ctx.L.aigw_chat_translate_host; args = (ctx.h, C.byref(cfg), a.ctypes.data, o.ctypes.data, l.ctypes.data, n, C.byref(bo))
        kms = []
        for k in range(400):
            t = time.perf_counter(); rc = fn(*args); lat.append(time.perf_counter() - t); kms.append(bo.kernel_ms)
            assert rc == 0
        lat = np.array(lat[100:]) * 1e6
        out[f"{mode}_n{n}"] = {"p50_us": float(np.median(lat)), "p99_us": float(np.percentile(lat, 99)), "bodies_per_s": n / float(np.median(lat)) * 1e6, "kernel_us_p50": float(np.median(kms[100:])) * 1e3}
print(json.dumps(out, indent=1))
