"""One-process parameter sweep of the flex kernel (run under gpurun): consumer threads x tile bytes x stages.
Prints GB/s of algorithmic traffic as a fraction of the measured copy peak for each setting."""
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import csic_b200 as csic
from bench import WORKLOADS, ORD, algorithmic_bytes_per_frame, load_peak

peak, _ = load_peak()
ctx = csic.Context(0)
ctx.set_option(0, 2)
names = sys.argv[1].split(",") if len(sys.argv) > 1 else ["cfg4odd", "cfg3odd", "cfg2", "cfg5", "cfg3b", "cfg3p"]
threads = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["160", "192", "224", "256"])]
tiles = [int(v) for v in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["8192", "12288", "16384", "24576"])]
stages = [int(v) for v in (sys.argv[4].split(",") if len(sys.argv) > 4 else ["2", "3"])]
for name in names:
    W, H, frames, a, b, q, f, order, fmt, _ = WORKLOADS[name]
    frames = max(8, min(frames, int(1.5e9 // (W * H * 3))))
    p = csic.make_params(W, H, a, b, q[0], q[1], q[2], f, tuple(ORD[c] for c in order), out_format=fmt)
    fb = csic.out_shape(p)[3]
    rgb = torch.randint(0, 256, (frames, H, W, 3), dtype=torch.uint8, device="cuda")
    out = torch.empty((frames, fb), dtype=torch.uint8, device="cuda")
    alg = algorithmic_bytes_per_frame(W, H, f, fb) * frames
    best = (0, None)
    for th, tb, st in itertools.product(threads, tiles, stages):
        ctx.set_option(5, th); ctx.set_option(4, tb); ctx.set_option(3, st)
        try:
            for _ in range(2):
                ctx.process_torch(p, rgb, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ctx.process_torch(p, rgb, out=out)
            e1.record(); torch.cuda.synchronize()
            fam = ctx.last_kernel()[0]
            ms = e0.elapsed_time(e1) / 5
            frac = alg / (ms / 1e3) / 1e9 / peak
        except Exception as ex:   # noqa
            fam, frac = -1, 0.0
        print(f"{name:8s} threads={th:3d} tile={tb:5d} stages={st} fam={fam} frac={frac:.3f}", flush=True)
        if frac > best[0]:
            best = (frac, (th, tb, st))
    print(f"== {name}: best {best}", flush=True)
    del rgb, out
    torch.cuda.empty_cache()
