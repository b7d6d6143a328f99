"""H2D, D2H and simultaneous H2D + D2H bandwidth from pinned memory (what bounds dftd3_host)."""
import time

import torch

dev = torch.device("cuda:0")
n = 16 * 1024 * 1024
h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
d_in, d_out = torch.empty(n, dtype=torch.uint8, device=dev), torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def run(h2d, d2h, reps=20, pieces=1):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step = n // pieces
    for _ in range(reps):
        for p in range(pieces):
            sl = slice(p * step, (p + 1) * step)
            if h2d:
                with torch.cuda.stream(s1):
                    d_in[sl].copy_(h_in[sl], non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out[sl].copy_(d_out[sl], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


for pieces in (1, 8):
    for name, a, b in (("H2D", True, False), ("D2H", False, True), ("both", True, True)):
        run(a, b, 3, pieces)
        t = run(a, b, 20, pieces)
        gb = n * (int(a) + int(b)) / 1e9
        print(f"pieces {pieces} {name:5s}: {t * 1e3:.3f} ms per 16 MiB each way -> {gb / t:.1f} GB/s total")
