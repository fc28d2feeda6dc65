"""aigw_b200/csrc/shortest_f64.cuh decides whether a 16- / 17-digit decimal is the text Go prints for the float64 it parses to (the request
translators re-encode numbers with strconv's shortest round-trip formatting).  The header is plain C++: tools/shortest_f64_check.cpp builds it
for the host and this test checks it on random doubles against Python's repr (the same shortest-digits algorithm; 'f' form in the tested
range): it must NEVER accept a text that is not its own float's repr (that would change bytes); declining a valid one is allowed."""
import os
import random
import shutil
import struct
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_never_accepts_a_text_go_would_respell(tmp_path):
    exe = str(tmp_path / "sf64")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tools", "shortest_f64_check.cpp")])
    rng = random.Random(1)
    cases = ["0.30000000000000004", "0.7000000000000001", "0.10000000000000001", "0.1", "1.0000000000000002", "0.30000000000000005", "12345.678901234567", "0.000123456789012345678"]
    for _ in range(60000):
        mode = rng.random()
        if mode < 0.4: x = rng.random() * 10 ** rng.randint(-3, 2)
        elif mode < 0.7: x = rng.uniform(0, 2)
        elif mode < 0.8: x = (rng.randint(1, 20) / 10.0) + (rng.randint(1, 20) / 10.0) * rng.choice([1, 0.1, 0.01])   # float-arithmetic artefacts
        elif mode < 0.9: x = struct.unpack("<d", struct.pack("<Q", rng.randint(0x3F00000000000000, 0x4340000000000000)))[0]
        else: x = 2.0 ** rng.randint(-12, 40) * (1 + rng.choice([0, 2 ** -52, -2 ** -53, 2 ** -51]))
        if rng.random() < 0.3: x = -x
        s = repr(x)
        if "e" in s or "." not in s:
            continue
        g17 = "%.17g" % x
        cases += [s, s[:-1] + str((int(s[-1]) + rng.choice([1, 9])) % 10), g17 if "e" not in g17 else s, s + str(rng.randint(1, 9))]
    out = subprocess.run([exe], input=("\n".join(cases) + "\n").encode(), capture_output=True, check=True).stdout.decode().split()
    assert len(out) == len(cases)
    accepted = missed = 0
    for s, o in zip(cases, out):
        nsig = len(s.lstrip("-").replace(".", "").lstrip("0"))
        own = repr(float(s)) == s
        if o == "1":
            assert own and nsig in (16, 17), (s, repr(float(s)))
            accepted += 1
        elif own and nsig in (16, 17) and not s.endswith("0") and len(s.split(".")[1]) <= 19:
            missed += 1
    assert out[0] == "1" and out[1] == "1" and out[2] == "0" and out[3] == "0" and out[5] == "0"
    assert accepted > 50000 and missed < accepted // 100, (accepted, missed)
