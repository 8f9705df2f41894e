"""One-process sweep of the row kernel's residency knobs (run under gpurun): CTAs/SM x stages x consumer threads x tile
bytes on chosen workloads.  Prints the fraction of the measured copy peak for each setting (0 = automatic)."""
import itertools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import csic_b200 as csic
from bench import WORKLOADS, ORD, algorithmic_bytes_per_frame, load_peak

peak, _ = load_peak()
ctx = csic.Context(0)
names = sys.argv[1].split(",")
ctas = [int(v) for v in sys.argv[2].split(",")]
stages = [int(v) for v in sys.argv[3].split(",")]
threads = [int(v) for v in sys.argv[4].split(",")]
tiles = [int(v) for v in (sys.argv[5].split(",") if len(sys.argv) > 5 else ["0"])]
for name in names:
    W, H, frames, a, b, q, f, order, fmt, _ = WORKLOADS[name]
    frames = max(8, min(frames, int(3e9 // (W * H * 3))))
    pool = 1 if name.endswith("avg") else 0
    p = csic.make_params(W, H, a, b, q[0], q[1], q[2], f, tuple(ORD[c] for c in order), pool_mode=pool, out_format=fmt)
    fb = csic.out_shape(p)[3]
    rgb = torch.randint(0, 256, (frames, H, W, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty((frames, fb), dtype=torch.uint8, device="cuda")
    alg = algorithmic_bytes_per_frame(W, H, f, fb, average=bool(pool)) * frames
    res = []
    for c, st, th, tb in itertools.product(ctas, stages, threads, tiles):
        ctx.set_option(2, c); ctx.set_option(3, st); ctx.set_option(5, th); ctx.set_option(4, tb)
        try:
            for _ in range(2):
                ctx.process_torch(p, rgb, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(8):
                ctx.process_torch(p, rgb, out=out)
            e1.record(); torch.cuda.synchronize()
            frac = alg / (e0.elapsed_time(e1) / 8 / 1e3) / 1e9 / peak
            fam = ctx.last_kernel()[0]
        except Exception:
            frac, fam = 0.0, -1
        res.append((frac, c, st, th, tb, fam))
        print(f"{name:9s} ctas={c} stages={st} threads={th:3d} tile={tb:5d} fam={fam} frac={frac:.3f}", flush=True)
    print("== best", name, sorted(res, reverse=True)[:3], flush=True)
    del rgb, out
    torch.cuda.empty_cache()
