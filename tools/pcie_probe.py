"""PCIe probe on the GPU box: pinned H2D / D2H copy rates through the library's own memcpy entry points."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import aigw_b200 as A
ctx = A.Context(0)
n = 1 << 30
h, hp = ctx.host_array(n)
h[:] = 1
d = ctx.dalloc(n)
for name, fn in (("h2d", lambda: ctx.h2d(d, h)), ("d2h", lambda: ctx.d2h(h, d))):
    fn()
    t = time.perf_counter()
    for _ in range(5):
        fn()
    dt = time.perf_counter() - t
    print(name, "pinned 1 GiB x5: %.1f GB/s" % (5 * n / dt / 1e9))
import torch
x = torch.empty(n, dtype=torch.uint8).pin_memory(); y = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
x2 = torch.empty(n, dtype=torch.uint8).pin_memory(); y2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1): y.copy_(x, non_blocking=True)
    with torch.cuda.stream(s2): x2.copy_(y2, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("torch concurrent h2d+d2h: %.1f GB/s each way" % (5 * n / dt / 1e9))
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(5):
    y.copy_(x, non_blocking=True)
torch.cuda.synchronize(); dt = time.perf_counter() - t
print("torch h2d alone: %.1f GB/s" % (5 * n / dt / 1e9))
