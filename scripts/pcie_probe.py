#!/usr/bin/env python
"""Host<->device copy ceilings of this box: contiguous cudaMemcpyAsync of pinned buffers, one direction at a time and both
at once (two streams) -- the bound of bench.py's `e2e` numbers (16 bytes per parameter per chain must cross the bus)."""
import json
import time

import torch

dev = torch.device("cuda", 0)
n = (2 << 30) // 8
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device=dev)
d_out = torch.empty(n, dtype=torch.float64, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=6):
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps):
        if h2d:
            with torch.cuda.stream(s1):
                d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    return n * 8 / dt / 1e9


run(True, True, 1)
out = {"h2d_alone_gbs": run(True, False), "d2h_alone_gbs": run(False, True), "both_each_way_gbs": run(True, True),
       "bytes_per_copy": n * 8}
# chunked: 256 MB pieces alternating on the two streams (what a pipelined host path issues)
k = 8
torch.cuda.synchronize()
t = time.perf_counter()
for rep in range(4):
    for q in range(k):
        a, b = q * n // k, (q + 1) * n // k
        with torch.cuda.stream(s1):
            d_in[a:b].copy_(h_in[a:b], non_blocking=True)
        with torch.cuda.stream(s2):
            h_out[a:b].copy_(d_out[a:b], non_blocking=True)
torch.cuda.synchronize()
out["both_each_way_256MB_chunks_gbs"] = n * 8 / ((time.perf_counter() - t) / 4) / 1e9
# 2-D strided rows (a slab of 256 chains out of a wider batch): 2 KB rows
rows, width, ld = 1 << 19, 256, 1024
h2 = torch.empty((rows, ld), dtype=torch.float64).pin_memory()
d2 = torch.empty((rows, width), dtype=torch.float64, device=dev)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(4):
    d2.copy_(h2[:, :width], non_blocking=True)
torch.cuda.synchronize()
out["h2d_2d_2KB_rows_gbs"] = rows * width * 8 / ((time.perf_counter() - t) / 4) / 1e9
print(json.dumps(out))
