"""dev tool (GPU box): raw pinned-memory PCIe rates, to know what the host-buffer path can reach."""
import torch, time
dev = torch.device("cuda", 0)
for mb in (2, 9, 18, 70):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
    h2 = torch.empty(8 << 20, dtype=torch.uint8).pin_memory(); d2 = torch.empty(8 << 20, dtype=torch.uint8, device=dev)
    s2 = torch.cuda.Stream()
    for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(20): fn()
        torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
        print("%s %3d MB: %.3f ms  %.1f GB/s" % (name, mb, dt * 1e3, n / dt / 1e9))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20):
        h.copy_(d, non_blocking=True)
        with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    print("d2h %d MB + h2d 8 MB concurrently: %.3f ms" % (mb, dt * 1e3))
