#!/usr/bin/env python
"""PCIe roofline for the host-array (e2e) path: 262 MB pinned H2D alone, D2H alone, and both at once on two
streams.  The e2e apply can not be faster than the simultaneous figure."""
import json
import torch

n = 32768000  # doubles = 262 MB, the bench's coefficient array
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.zeros(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def both():
    s1.wait_stream(torch.cuda.current_stream())
    s2.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)


t_h2d = timed(lambda: d_in.copy_(h_in, non_blocking=True))
t_d2h = timed(lambda: h_out.copy_(d_out, non_blocking=True))
t_both = timed(both)
gb = n * 8 / 1e9
print(json.dumps({"bytes_each_way": n * 8, "h2d_ms": t_h2d, "h2d_gbs": gb / t_h2d * 1e3, "d2h_ms": t_d2h,
                  "d2h_gbs": gb / t_d2h * 1e3, "simultaneous_ms": t_both,
                  "simultaneous_gbs_each_way": gb / t_both * 1e3}))
