import torch, time
x = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for name, f in (("H2D pinned 1GiB", lambda: d.copy_(x, non_blocking=True)), ("D2H pinned 1GiB", lambda: x.copy_(d, non_blocking=True))):
    f(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): f()
    torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (5 * (1 << 30) / (time.perf_counter() - t) / 1e9))
p = torch.empty(1 << 28, dtype=torch.uint8)
t = time.perf_counter(); d[: 1 << 28].copy_(p); torch.cuda.synchronize(); print("H2D pageable 256MiB %.1f GB/s" % ((1 << 28) / (time.perf_counter() - t) / 1e9))
t = time.perf_counter(); q = d[: 1 << 28].cpu(); print("D2H pageable 256MiB %.1f GB/s" % ((1 << 28) / (time.perf_counter() - t) / 1e9))
import os; print("cpus", os.cpu_count())
# chunked copies of 256 MB like aggregate_host
s2 = torch.cuda.Stream()
t = time.perf_counter()
with torch.cuda.stream(s2):
    for i in range(4):
        d[i * (1 << 28):(i + 1) * (1 << 28)].copy_(x[i * (1 << 28):(i + 1) * (1 << 28)], non_blocking=True)
torch.cuda.synchronize(); print("H2D pinned 4x256MiB on side stream %.1f GB/s" % ((1 << 30) / (time.perf_counter() - t) / 1e9))
