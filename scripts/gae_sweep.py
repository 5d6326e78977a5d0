"""BASELINE.json configs[2]: GAE-only bandwidth sweep, T in {16..1024} x N in {1K..1M}, fp32, against the HBM roofline.
Algorithmic bytes = 17 B / transition (+ 4 B / env for last_val).  CUDA events, warm, inputs rotate through enough
distinct buffers to exceed L2 when one problem fits in it.  Prints one JSON line per shape + a table."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minppo_b200.learner import Memory, calculate_gae

dev = torch.device("cuda:0")
peak = 6541.8
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
L2 = 126e6
rows = []
# full configs[2] grid: every power of two T = 16 .. 1024, N = 2^10 .. 2^20 (largest case 2^30 transitions = 18.3 GB)
for T in (16, 32, 64, 128, 256, 512, 1024):
    for N in (1 << 10, 1 << 12, 1 << 14, 1 << 16, 1 << 18, 1 << 20):
        nbytes = T * N * 17 + 4 * N
        copies = max(1, min(16, int(2 * L2 / nbytes) + 1))   # rotate buffers so that small problems do not live in L2
        g = torch.Generator(device=dev).manual_seed(0)
        bufs = []
        for _ in range(copies):
            bufs.append((Memory(torch.rand(T, N, device=dev, generator=g) < 0.01, None, torch.randn(T, N, device=dev, generator=g),
                                torch.randn(T, N, device=dev, generator=g), None, None), torch.randn(N, device=dev, generator=g)))
        for m, lv in bufs[:2]:
            calculate_gae(m, lv, 0.99, 0.95)
        torch.cuda.synchronize()
        reps = max(3, min(200, int(2e9 / nbytes)))
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(reps):
            m, lv = bufs[i % copies]
            calculate_gae(m, lv, 0.99, 0.95)
        b.record()
        torch.cuda.synchronize()
        us = a.elapsed_time(b) / reps * 1e3
        gbs = nbytes / (us * 1e-6) / 1e9
        rows.append({"T": T, "N": N, "us": round(us, 2), "GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 3),
                     "frac_of_8TBs": round(gbs / 8000, 3), "MB": round(nbytes / 1e6, 2), "rotating_buffers": copies})
        print(json.dumps(rows[-1]), flush=True)
        del bufs, m, lv
        torch.cuda.empty_cache()
print("note: time includes the host-side launch path of calculate_gae (one allocation of the outputs + one launch); shapes below ~10 MB are launch-latency numbers, not bandwidth")
