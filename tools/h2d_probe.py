import torch, time
x = torch.empty(1812070404 // 4, dtype=torch.float32, pin_memory=True)
d = torch.empty_like(x, device="cuda")
for n in (1, 4, 16):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    chunks = x.chunk(n); dch = d.chunk(n)
    e0.record()
    for a, b in zip(chunks, dch): b.copy_(a, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(n, "chunks: H2D GB/s", x.numel() * 4 / e0.elapsed_time(e1) / 1e6)
y = torch.empty(1 << 28, dtype=torch.float32, pin_memory=True)
e0.record(); y.copy_(d[: 1 << 28], non_blocking=True); e1.record(); torch.cuda.synchronize()
print("D2H GB/s", (1 << 30) / e0.elapsed_time(e1) / 1e6)
