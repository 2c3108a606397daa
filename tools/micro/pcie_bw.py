#!/usr/bin/env python
"""Raw host<->device copy rates of the box, pinned memory, the sizes one bench step moves (316 MB in, 80 MB out):
H2D alone, D2H alone, both at once on two streams. Prints one JSON line. usage: python tools/micro/pcie_bw.py"""
import json
import torch


def rate(fn, nbytes, reps=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record(); torch.cuda.synchronize()
    return nbytes * reps / (a.elapsed_time(b) * 1e-3) / 1e9


def main():
    n_in, n_out = 316_000_016, 80_000_000
    hi = torch.empty(n_in, dtype=torch.uint8).pin_memory(); di = torch.empty(n_in, dtype=torch.uint8, device="cuda")
    ho = torch.empty(n_out, dtype=torch.uint8).pin_memory(); do = torch.empty(n_out, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    out = {"h2d_gb_s": rate(lambda: di.copy_(hi, non_blocking=True), n_in), "d2h_gb_s": rate(lambda: ho.copy_(do, non_blocking=True), n_out)}

    def both():
        with torch.cuda.stream(s1):
            di.copy_(hi, non_blocking=True)
        with torch.cuda.stream(s2):
            ho.copy_(do, non_blocking=True)
    both(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(10):
        both()
    s1.synchronize(); s2.synchronize(); b.record(); torch.cuda.synchronize()
    t = a.elapsed_time(b) * 1e-3
    out["both_h2d_gb_s"] = n_in * 10 / t / 1e9; out["both_d2h_gb_s"] = n_out * 10 / t / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
