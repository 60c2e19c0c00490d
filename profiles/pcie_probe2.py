"""Does a device->host copy slow down when the pinned destination is dirty in the CPU caches (the host stages edit the
result buffers in place, and the next step's copy lands on the same lines)?"""
import torch
dev = torch.device("cuda", 0)
torch.set_num_threads(16)
for mb in (4, 12, 24):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
    for mode in ("untouched", "cpu read before each copy", "cpu wrote before each copy"):
        ms = 0.0
        for it in range(8):
            if mode.startswith("cpu read"): _ = int(h.sum())
            elif mode.startswith("cpu wrote"): h.add_(1)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); h.copy_(d, non_blocking=True); b.record(); torch.cuda.synchronize()
            if it >= 2: ms += a.elapsed_time(b)
        print(f"d2h {mb:3d} MiB, destination {mode:28s}: {n * 6 / (ms * 1e-3) / 1e9:6.1f} GB/s", flush=True)
