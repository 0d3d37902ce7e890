#!/usr/bin/env python
"""Pinned host<->device copy bandwidth of the box (sets the floor of the e2e leg: ~17 MB in + ~19 MB out per step)."""
import torch
dev = torch.device("cuda", 0)
for mb in (4, 18, 64):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device=dev)
    for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): fn()
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name} {mb} MB: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s")
# both directions at once
h2 = torch.empty(18 << 20, dtype=torch.uint8).pin_memory(); d2 = torch.empty(18 << 20, dtype=torch.uint8, device=dev)
h3 = torch.empty(18 << 20, dtype=torch.uint8).pin_memory(); d3 = torch.empty(18 << 20, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    with torch.cuda.stream(s1): d2.copy_(h2, non_blocking=True)
    with torch.cuda.stream(s2): h3.copy_(d3, non_blocking=True)
torch.cuda.synchronize()
e1.record(); e1.synchronize()
print("bidirectional 18+18 MB: %.3f ms per pair" % (e0.elapsed_time(e1) / 10))
