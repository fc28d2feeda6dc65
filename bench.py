#!/usr/bin/env python3
"""bench.py — BASELINE.json's metric on its config #2:
    "1M OpenAI→AWS Bedrock Converse body translate, 4 KB bodies, 1×B200"

One step = one pass of the hot path (parse → validate → translate → emit) over a batch of synthetic
ChatCompletion bodies.  Per GPU the batch is --bodies (default 1,000,000; weak scaling: every rank gets
its own 1 M shard, seeded by rank).  Printed (rank 0, one JSON line):

  value      bodies/s, whole job, inputs already resident in HBM (device API; CUDA-event time)
  e2e        bodies/s through the host-buffer C-ABI call: pinned host arena → H2D → kernel → D2H, all in
             the timed region (the call the cgo shim makes)
  roofline   HBM roofline of the translate kernel: algorithmic bytes (input body + output record)
             ÷ CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle (CPU restatement of the reference Go path) on this box's host cores,
             bounded sample, rank 0 / N=1 only
  p50_added_us  submit→complete latency of a single-body call through the host API

--impl reference times the CPU restatement alone (the reference is Go; no Go toolchain exists in the
image, so the restatement — pinned by the reference's goldens — stands in as kind "port").
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "extproc bodies/sec, 4 KB ChatCompletion translate (OpenAI→AWS Bedrock Converse)"
SEED = 2


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        busy = sorted(sm)[len(sm) // 2:] if sm else []  # upper half ≈ samples under load
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(arena, offs, lens, n_sample, threads):
    import _oracle as O
    n = min(n_sample, len(lens))
    status = np.zeros(n, dtype=np.int32); olen = np.zeros(n, dtype=np.uint32); tot = C.c_uint64(0)
    L = O.lib()
    sec = L.oracle_chat_translate_batch(1, arena.ctypes.data, offs.ctypes.data, n, 0, threads, status.ctypes.data, olen.ctypes.data, C.byref(tot))
    assert (status == 0).all()
    return n / sec, n, int(tot.value)


# DRAM traffic per body of the three stages, from the ncu --set full captures named in profiles/r01_chat_final.md
NCU_DRAM_BYTES_PER_BODY = 16860


def host_cores():
    """threads the CPU legs may really use: the scheduler affinity mask, further limited by a cgroup CPU quota when one is set
    (a box can show 128 logical CPUs and still be capped; oversubscribing a quota only adds throttling)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q = open("/sys/fs/cgroup/cpu.max").read().split()
        if q[0] != "max":
            n = max(1, min(n, int(float(q[0]) / float(q[1]) + 0.5)))
    except Exception:
        try:
            quota = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read()); period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if quota > 0:
                n = max(1, min(n, int(quota / period + 0.5)))
        except Exception:
            pass
    return n


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bodies", type=int, default=int(os.environ.get("AIGW_BENCH_BODIES", 1_000_000)), help="bodies per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=400_000)
    ap.add_argument("--skip-e2e", action="store_true")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
    ncpu = host_cores()
    import _workload as W
    import __graft_entry__ as entry

    if a.impl == "reference":
        # reference arm: the reference's own CPU implementation of the path.  It is Go and cannot be built here,
        # so the restatement (oracle, kind "port") is timed with every host thread.  Rank 0 only.
        if rank != 0:
            return
        entry.build()
        n = min(a.bodies, 400_000)
        arena, offs, lens = W.chat_corpus(SEED, 0, n, threads=ncpu)
        times = []
        for s in range(a.warmup + a.steps):
            rate, nn, tot = cpu_oracle_rate(arena, offs, lens, n, ncpu)
            if s >= a.warmup:
                times.append(nn / rate)
        t = float(np.sum(times))
        val = n * a.steps / t
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": val, "unit": "bodies/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                          "ms_per_step": 1e3 * t / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                          "config": {"workload": "configs[1]: OpenAI→AWS Bedrock Converse translate, 4 KB bodies", "bodies_per_step": n, "seed": SEED,
                                     "note": "CPU restatement of the reference Go path (oracle); bounded sample of the 1M-body workload"},
                          "cpu_baseline": {"value": val, "unit": "bodies/s", "cores": ncpu, "kind": "port", "sample": f"{n} bodies/step x {a.steps} steps, {ncpu} threads"},
                          "e2e": {"value": val, "unit": "bodies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    entry.build()
    import aigw_b200 as A
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    ctx = A.Context(local)
    n = a.bodies
    # ---- synthetic shard for this rank, generated straight into a pinned arena
    lens = W.chat_lens(SEED, rank * n, n, threads=ncpu)
    offs = W.chat_offsets(lens)
    in_bytes = int(offs[-1]) + 16
    arena, arena_ptr = ctx.host_array(in_bytes)
    arena[-16:] = 0x20
    W.chat_fill(SEED, rank * n, n, arena, offs, lens, threads=ncpu)
    max_len = int(lens.max())
    cfg = ctx.cfg("aws-bedrock")

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload, all host threads and 1 thread
    cpu = None
    if rank == 0 and world == 1:
        r1, n1, _ = cpu_oracle_rate(arena, offs, lens, max(2000, min(a.cpu_sample // 8, 50_000)), 1)
        rN, nN, _ = cpu_oracle_rate(arena, offs, lens, a.cpu_sample, ncpu)
        rN2, _, _ = cpu_oracle_rate(arena, offs, lens, a.cpu_sample, ncpu)
        cpu = {"value": max(rN, rN2), "unit": "bodies/s", "cores": ncpu, "kind": "port", "sample": f"first {nN} bodies of the workload, {ncpu} threads (1 thread: {r1:.0f} bodies/s on {n1})",
               "p50_us_per_body_1thread": 1e6 / r1}

    # ---- device-resident pass ("value")
    d_in = ctx.dalloc(in_bytes); d_off = ctx.dalloc(offs.nbytes); d_len = ctx.dalloc(lens.nbytes)
    out_cap = int(in_bytes * 1.25) + n * 64 + 4096
    d_out = ctx.dalloc(out_cap); d_res = ctx.dalloc(n * 32); d_used = ctx.dalloc(8)
    ctx.h2d(d_in, arena); ctx.h2d(d_off, offs); ctx.h2d(d_len, lens)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
        ctx.sync()

    for _ in range(a.warmup):
        ctx.memset(d_used, 0, 8)
        ctx.chat_translate_device(cfg, d_in, d_off, d_len, n, max_len, d_out, out_cap, d_res, d_used)
    sampler = ClockSampler(local); sampler.start()
    barrier()
    t0 = time.perf_counter(); kms = []; stage = np.zeros(3); dev_launches = 0
    for _ in range(a.steps):
        ctx.memset(d_used, 0, 8)
        kms.append(ctx.chat_translate_device(cfg, d_in, d_off, d_len, n, max_len, d_out, out_cap, d_res, d_used))
        dev_launches += ctx.chat_last_profile()["launches"]
    barrier()
    wall = time.perf_counter() - t0
    dev_s = float(np.sum(kms)) / 1e3
    # per-stage times: one extra pass outside the timed region with the stages serialised on one stream (the timed passes
    # overlap the walk of one sub-batch with the index / emit of the next, so stage times do not exist there)
    ctx.chat_set_profile(True); ctx.memset(d_used, 0, 8)
    serial_ms = ctx.chat_translate_device(cfg, d_in, d_off, d_len, n, max_len, d_out, out_cap, d_res, d_used)
    pr = ctx.chat_last_profile(); stage += np.array([pr["index_ms"], pr["walk_ms"], pr["emit_ms"]]) * a.steps
    ctx.chat_set_profile(False); ctx.memset(d_used, 0, 8)
    ctx.chat_translate_device(cfg, d_in, d_off, d_len, n, max_len, d_out, out_cap, d_res, d_used)
    used = np.zeros(1, dtype=np.uint64); ctx.d2h(used, d_used)
    res = np.zeros(n, dtype=A.DocResult); ctx.d2h(res, d_res)
    n_ok = int((res["status"] == A.AIGW_OK).sum())
    out_bytes = int(res["body_len"].astype(np.uint64).sum() + res["path_len"].astype(np.uint64).sum())
    alg_bytes = int(lens.astype(np.uint64).sum()) + out_bytes  # bytes_in + bytes_out per body
    for p in (d_in, d_off, d_len, d_out, d_res, d_used):
        ctx.dfree(p)

    # ---- end-to-end through the host-buffer C-ABI call
    e2e = None; p50 = None; launches = dev_launches
    if not a.skip_e2e:
        for _ in range(max(1, min(a.warmup, 2))):
            _, _, st = ctx.chat_translate_host(cfg, arena, offs, lens)
        barrier()
        t1 = time.perf_counter(); e_launch = 0
        for _ in range(a.steps):
            r2, o2, st = ctx.chat_translate_host(cfg, arena, offs, lens)
            e_launch += st["gpu_launches"]
        barrier()
        e_wall = time.perf_counter() - t1
        assert int((r2["status"] == A.AIGW_OK).sum()) == n_ok
        e2e = {"wall_s": e_wall, "h2d": st["h2d_bytes"], "d2h": st["d2h_bytes"], "launches_per_step": e_launch // a.steps}
        launches += e_launch
        # p50 added latency: one 4 KB body, submit → complete
        one_a, one_o, one_l = arena[: int(offs[1]) + 16], offs[:2].copy(), lens[:1].copy()
        lat = []
        for k in range(300):
            t = time.perf_counter(); ctx.chat_translate_host(cfg, one_a, one_o, one_l); lat.append(time.perf_counter() - t)
        p50 = float(np.median(lat[50:])) * 1e6
    clocks = sampler.stop()

    # ---- reduce over ranks: MAX time, SUM bodies
    tot_bodies = n * world
    dev_s_max, wall_max, e_wall_max = dev_s, wall, (e2e["wall_s"] if e2e else 0.0)
    if dist is not None:
        import torch
        t = torch.tensor([dev_s, wall, e_wall_max], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_s_max, wall_max, e_wall_max = [float(x) for x in t.tolist()]
    if rank == 0:
        peak, peak_src = hbm_peak()
        step_s = dev_s_max / a.steps
        achieved = alg_bytes / (dev_s / a.steps) / 1e9
        line = {"metric": METRIC, "value": tot_bodies / step_s, "unit": "bodies/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": "configs[1]: OpenAI→AWS Bedrock Converse translate, 4 KB bodies (4096±32 B, 4–8 messages, 10% tool use), seed 2",
                           "bodies_per_gpu_per_step": n, "mean_in_bytes": float(lens.mean()), "mean_out_bytes": out_bytes / max(1, n_ok), "accepted": n_ok, "declined": n - n_ok,
                           "l2": f"inputs larger than L2 ({in_bytes / 1e6:.0f} MB in + {out_bytes / 1e6:.0f} MB out per step vs 126 MB L2)", "sharding": "hash(request-id)→device, no collective",
                           "wall_ms_per_step_incl_launch": wall_max / a.steps * 1e3},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": NCU_DRAM_BYTES_PER_BODY * n, "peak_source": peak_src,
                             "traffic_source": "ncu --set full captures of one 131072-body launch of each stage (profiles/r01_chat_final.md): dram read+write per body "
                                               "index 4.50 KB + walk 3.03 KB + emit 9.33 KB = 16.86 KB, i.e. 2.0x the algorithmic 8.25 KB (the body is read by index and again by emit)",
                             "kernel": "chat_index_kernel + chat_walk_kernel + chat_emit_kernel (the three stages of one translate pass)",
                             "algorithmic_bytes_per_step": alg_bytes, "avg_step_ms": dev_s / a.steps * 1e3,
                             "stage_ms_serialised": {"index": stage[0] / a.steps, "walk": stage[1] / a.steps, "emit": stage[2] / a.steps, "sum_one_stream": serial_ms},
                             "launches_per_step": dev_launches // max(1, a.steps)},
                "gpu_launches": launches, "clocks": clocks}
        if e2e:
            line["e2e"] = {"value": tot_bodies * a.steps / e_wall_max, "unit": "bodies/s", "h2d_bytes_per_step": int(e2e["h2d"]), "d2h_bytes_per_step": int(e2e["d2h"]),
                           "launches_per_step": e2e["launches_per_step"], "pcie_gbs_each_way": [e2e["h2d"] * a.steps / e_wall_max / 1e9, e2e["d2h"] * a.steps / e_wall_max / 1e9]}
            line["p50_added_us"] = p50
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    ctx.host_free(arena_ptr)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
