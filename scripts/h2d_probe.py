"""H2D bandwidth of pinned host buffers: does it depend on how the buffer was made, or on link warm-up?"""
import os, time, subprocess
import torch
dev = torch.device("cuda", 0)
torch.cuda.init()
n = 17825792 // 4
d = torch.empty(n, device=dev)
def bw(h, reps=20, warm=3):
    for _ in range(warm):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    return n * 4 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9
def link():
    return subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv,noheader", shell=True, capture_output=True, text=True).stdout.strip()
print("link before", link())
a = torch.randn(n).pin_memory()
print("A randn().pin_memory(): " + " ".join("%.1f" % bw(a, warm=0) for _ in range(6)), "| link", link())
b = torch.empty(n).pin_memory(); b.normal_()
print("B empty().pin_memory()+normal_: " + " ".join("%.1f" % bw(b) for _ in range(3)))
print("A again: " + " ".join("%.1f" % bw(a) for _ in range(3)))
c = torch.randn(n).pin_memory()
print("C second randn().pin_memory(): " + " ".join("%.1f" % bw(c) for _ in range(3)))
big = torch.randn(10, n).pin_memory()
print("D slices of one 10x pinned block: " + " ".join("%.1f" % bw(big[i]) for i in (0, 3, 9)))
e = torch.empty(n, pin_memory=True); e.normal_()
print("E empty(pin_memory=True): " + " ".join("%.1f" % bw(e) for _ in range(2)))
time.sleep(2.0)
print("after 2 s idle, A: " + " ".join("%.1f" % bw(a, warm=0, reps=2) for _ in range(6)), "| link", link())
# single-copy latency in a loop with host sync each time (the serial e2e pattern)
torch.cuda.synchronize()
for h, name in ((a, "A"), (b, "B"), (big[5], "D5")):
    ts = []
    for _ in range(10):
        t = time.perf_counter(); d.copy_(h, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    print(name, "one copy + sync, us:", " ".join("%.0f" % (x * 1e6) for x in ts))
