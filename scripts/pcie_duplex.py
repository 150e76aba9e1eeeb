#!/usr/bin/env python
"""PCIe copy bandwidth, one direction at a time and both at once (pinned host memory); GPU box only."""
import torch

n = 1 << 28
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.zeros(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=10):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)
    e1.record()
    e1.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


run(True, True, 2)
print("H2D alone   %.1f GB/s" % run(True, False))
print("D2H alone   %.1f GB/s" % run(False, True))
print("both at once %.1f GB/s per direction" % run(True, True))
