"""bench.py's reference arm needs no GPU (it times the CPU restatement of the path on the host cores): run it small and check the JSON line's
contract — the keys the driver reads, the same metric / unit / config wording as the product arm, zero copy bytes, no product library loaded."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return json.loads([l for l in out.stdout.strip().split("\n") if l.startswith("{")][-1])


def test_reference_arm_line_config2():
    d = _line(["--impl", "reference", "--steps", "2", "--warmup", "1", "--bodies", "20000", "--cpu-sample", "20000"])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 2 and d["warmup"] == 1 and d["higher_is_better"] is True
    assert d["unit"] == "bodies/s" and d["value"] > 0 and d["ms_per_step"] > 0 and d["dtype"] == "u8" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "OpenAI" in d["metric"] and d["config"]["workload"].startswith("configs[1]") and d["config"]["bodies_per_gpu_per_step"] == 20000
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "bodies/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_does_not_load_the_product_library():
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--bodies', '4000', '--cpu-sample', '4000'];\n"
            "runpy.run_path(%r, run_name='__main__')\n"
            "maps = open('/proc/self/maps').read(); assert 'libaigw_b200' not in maps, 'the reference arm mapped the product library'; assert 'liboracle' in maps\n") % os.path.join(ROOT, "bench.py")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
