"""Pinned host -> device copy bandwidth (one stream), the ceiling of the host-buffer entry point."""
import torch, time
for mb in (8, 48, 300):
    h = torch.empty(mb << 20, dtype=torch.uint8).pin_memory()
    d = torch.empty(mb << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3): d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10): d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"H2D {mb} MiB: {ms:.3f} ms  {(mb << 20) / ms / 1e6:.1f} GB/s", flush=True)
    e0.record()
    for _ in range(10): h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"D2H {mb} MiB: {ms:.3f} ms  {(mb << 20) / ms / 1e6:.1f} GB/s", flush=True)
