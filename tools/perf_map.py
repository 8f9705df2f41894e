"""Performance map of the automatic kernel choice (run under gpurun): common and awkward resolutions x factor x
format.  Batches are sized by the ALGORITHMIC bytes of a launch (~2 GB, i.e. >= 0.3 ms at the copy peak; input capped at
24 GB): round 1 sized them by input bytes (1.5 GB), which at f = 8 left 35 us launches whose ramp-up and tail cost 10 %
(the same kernel measured 0.97 on a 1024-frame 4K batch and 0.84 here).  Prints kernel family and the fraction of the
measured copy peak."""
import itertools
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import csic_b200 as csic
from bench import ORD, algorithmic_bytes_per_frame, load_peak

peak, _ = load_peak()
ctx = csic.Context(0)
sizes = [(640, 480), (1280, 720), (1366, 768), (1920, 1080), (2560, 1440), (3840, 2160), (4096, 2160), (7680, 4320),
         (1000, 1000), (500, 333), (333, 500), (1080, 1920), (96, 96), (200, 200)]
orders = sys.argv[1].split(",") if len(sys.argv) > 1 else ["CSQ"]
pools = [int(v) for v in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"])]
only_fmts = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else None      # e.g. "1" = RGB888 only
FAM = {1: "generic", 2: "rows", 3: "pool", 4: "flex"}
worst = []
for (W, H), f, (fmt, q), order, pool in itertools.product(sizes, (1, 2, 4, 8), ((0, (8, 8, 8)), (3, (8, 8, 8)), (1, (6, 5, 5))), orders, pools):
    if pool and (f == 1 or W % f or H % f):
        continue
    if only_fmts is not None and fmt not in only_fmts:
        continue
    p = csic.make_params(W, H, 2, 0, q[0], q[1], q[2], f, tuple(ORD[c] for c in order), pool_mode=pool, out_format=fmt)
    fb = csic.out_shape(p)[3]
    per_frame = algorithmic_bytes_per_frame(W, H, f, fb, average=bool(pool))
    frames = max(2, min(int(2.0e9 // per_frame), int(24e9 // (W * H * 3))))
    rgb = torch.empty((frames, H, W, 3), dtype=torch.uint8, device="cuda")
    rgb.random_(0, 256)
    out = torch.empty((frames, fb), dtype=torch.uint8, device="cuda")
    alg = algorithmic_bytes_per_frame(W, H, f, fb, average=bool(pool)) * frames
    for _ in range(2):
        ctx.process_torch(p, rgb, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ctx.process_torch(p, rgb, out=out)
    e1.record(); torch.cuda.synchronize()
    frac = alg / (e0.elapsed_time(e1) / 5 / 1e3) / 1e9 / peak
    fam = ctx.last_kernel()[0]
    tag = f"{W}x{H} f={f} fmt={fmt} {order} pool={pool}"
    print(f"{tag:34s} {FAM.get(fam, fam):8s} {frac:.3f}", flush=True)
    worst.append((frac, tag, FAM.get(fam, fam)))
    del rgb, out
print("== ten slowest:")
for w in sorted(worst)[:10]:
    print("  ", w)
