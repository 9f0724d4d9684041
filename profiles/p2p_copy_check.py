import torch, time
n = 1 << 30
a = torch.empty(n, dtype=torch.uint8, device="cuda:0")
b = torch.empty(n, dtype=torch.uint8, device="cuda:1")
for _ in range(3): b.copy_(a)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
t0 = time.perf_counter()
for _ in range(10): b.copy_(a)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
dt = (time.perf_counter() - t0) / 10
print(f"peer DMA copy 0->1: {n / dt / 1e9:.0f} GB/s")
# both directions at once
c = torch.empty(n, dtype=torch.uint8, device="cuda:0"); d = torch.empty(n, dtype=torch.uint8, device="cuda:1")
s0, s1 = torch.cuda.Stream(0), torch.cuda.Stream(1)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s0): b.copy_(a, non_blocking=True)
    with torch.cuda.stream(s1): c.copy_(d, non_blocking=True)
torch.cuda.synchronize(0); torch.cuda.synchronize(1)
dt = (time.perf_counter() - t0) / 10
print(f"bidirectional: {n / dt / 1e9:.0f} GB/s per direction")
