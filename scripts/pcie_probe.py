"""Host<->device copy bandwidth on the GPU box (pinned memory), to put the e2e number in context."""
import time
import torch
n = 1 << 28  # 256 Mi int64 = 2 GiB
h = torch.empty(n, dtype=torch.int64).pin_memory()
d = torch.empty(n, dtype=torch.int64, device="cuda")
h2 = torch.empty(n, dtype=torch.int64).pin_memory()
d2 = torch.empty(n, dtype=torch.int64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for name, fn in [("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))]:
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    print(name, "GB/s", 3 * n * 8 / (time.perf_counter() - t0) / 1e9)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize()
print("bidir GB/s each way", 3 * n * 8 / (time.perf_counter() - t0) / 1e9)
import os
print("cpus", len(os.sched_getaffinity(0)), open("/proc/cpuinfo").read().split("model name")[1].split("\n")[0])
print(os.popen("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current,pcie.link.gen.max --format=csv").read())
print(os.popen("free -g | head -2").read())
