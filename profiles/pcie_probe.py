"""Host<->device copy rates of the box for the sizes the pipeline moves (pinned memory, CUDA events): is the 13.8 GB/s seen for the
result copy a property of the platform?"""
import torch
dev = torch.device("cuda", 0)
for mb in (1, 4, 17, 64, 256):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, fn in (("h2d", lambda: d.copy_(h, non_blocking=True)), ("d2h", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): fn()
        b.record(); torch.cuda.synchronize()
        print(f"{name} {mb:4d} MiB: {n * 10 / (a.elapsed_time(b) * 1e-3) / 1e9:6.1f} GB/s", flush=True)
