import torch, time
for mb in (64, 512, 2048):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    t0 = time.perf_counter()
    for _ in range(5): h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dt2 = (time.perf_counter() - t0) / 5
    print(f"{mb} MB: H2D {n/dt/1e9:.1f} GB/s, D2H {n/dt2/1e9:.1f} GB/s")
