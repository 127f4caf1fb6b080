import torch, time
x = torch.empty(256 << 20, dtype=torch.uint8).pin_memory()
y = torch.empty_like(x, device="cuda")
for _ in range(2): y.copy_(x, non_blocking=True); torch.cuda.synchronize()
t = time.time(); y.copy_(x, non_blocking=True); torch.cuda.synchronize(); dt = time.time() - t
print("H2D pinned 256 MiB: %.2f GB/s" % (x.numel() / dt / 1e9))
t = time.time(); x.copy_(y, non_blocking=True); torch.cuda.synchronize(); dt = time.time() - t
print("D2H pinned 256 MiB: %.2f GB/s" % (x.numel() / dt / 1e9))
