"""Concurrent pinned-host -> device copy bandwidth for subsets of the box's GPUs: the ceiling of the end-to-end leg at N > 1
(bench.py reports the all-ranks figure as e2e.ceiling_gbs) and which GPUs share a host link.  One process, one stream per device."""
import itertools
import sys
import time

import torch

n_dev = torch.cuda.device_count()
size = 1 << 30
host = [torch.empty(size, dtype=torch.uint8).pin_memory() for _ in range(n_dev)]
dev = [torch.empty(size, dtype=torch.uint8, device=f"cuda:{d}") for d in range(n_dev)]
streams = [torch.cuda.Stream(device=d) for d in range(n_dev)]


def run(subset, reps=3):
    best = 0.0
    for _ in range(reps):
        for d in subset:
            torch.cuda.synchronize(d)
        t0 = time.perf_counter()
        for d in subset:
            with torch.cuda.stream(streams[d]):
                dev[d].copy_(host[d], non_blocking=True)
        for d in subset:
            streams[d].synchronize()
        best = max(best, len(subset) * size / (time.perf_counter() - t0) / 1e9)
    return best


subsets = [(d,) for d in range(n_dev)]
subsets += [p for p in itertools.combinations(range(n_dev), 2) if p[0] == 0 or p[1] == p[0] + 1]
if n_dev >= 4:
    subsets += [(0, 1, 2, 3)]
if n_dev >= 8:
    subsets += [(4, 5, 6, 7), (0, 1, 4, 5), (0, 2, 4, 6), (0, 1, 2, 3, 4, 5), tuple(range(8))]
print(f"{n_dev} GPUs, 1 GiB pinned buffers, best of 3")
for s in subsets:
    g = run(s)
    print(f"GPUs {','.join(map(str, s)):16s} {g:7.1f} GB/s aggregate  {g / len(s):6.1f} GB/s per GPU", flush=True)
