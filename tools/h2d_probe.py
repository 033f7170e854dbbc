#!/usr/bin/env python
"""How fast does a pinned 5 MB buffer cross PCIe: one copy vs chunks, with and without kernels
running on another stream (decides whether chunked upload + overlapped scan can pay off)."""
import json
import sys

import torch


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
    dev = torch.device("cuda", 0)
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    host.random_(0, 255)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream()
    busy = torch.empty(64 << 20, dtype=torch.float32, device=dev)

    def run(chunks, with_kernels, reps=50):
        cuts = [0]
        for f in chunks:
            cuts.append(min(n, cuts[-1] + int(f * n)))
        cuts[-1] = n
        ts = []
        for _ in range(reps):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if with_kernels:
                for _ in range(6):
                    busy.mul_(1.0001)          # ~40 us each on the default stream
            with torch.cuda.stream(side):
                a.record()
                for lo, hi in zip(cuts[:-1], cuts[1:]):
                    d[lo:hi].copy_(host[lo:hi], non_blocking=True)
                b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    out = {}
    for name, ch in (("1", [1.0]), ("3", [0.5, 0.3, 0.2]), ("4", [0.35, 0.3, 0.2, 0.15]), ("8", [0.125] * 8)):
        out["chunks_%s_idle_us" % name] = run(ch, False)
        out["chunks_%s_busy_us" % name] = run(ch, True)
    out["gb_per_s_single_idle"] = n / out["chunks_1_idle_us"] * 1e-3
    print(json.dumps(out))


if __name__ == "__main__":
    main()
