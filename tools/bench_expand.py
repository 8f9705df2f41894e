"""Times the PLANAR decoder (csic_expand_planar_device) on the cfg3p geometry; prints GB/s of algorithmic traffic
(planar bytes read + 3 B/px written) as a fraction of the measured copy peak.  Run under gpurun."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import csic_b200 as csic
from bench import load_peak

peak, _ = load_peak()
ctx = csic.Context(0)
for (W, H, n, a, b, f) in ((1920, 1080, 256, 2, 0, 1), (3840, 2160, 64, 2, 2, 2), (1918, 1078, 256, 2, 0, 1), (1366, 768, 512, 2, 0, 1),
                           (333, 500, 2048, 2, 2, 1), (1000, 1000, 256, 4, 4, 1), (2732, 1536, 128, 2, 0, 2), (96, 96, 32768, 2, 0, 1),
                           (32, 32, 262144, 2, 0, 1), (1001, 999, 256, 2, 0, 1)):
    p = csic.make_params(W, H, a, b, 8, 8, 8, f, out_format=4)
    ow, oh, _, fb = csic.out_shape(p)
    rgb = torch.randint(0, 256, (n, H, W, 3), dtype=torch.uint8, device="cuda")
    planar = ctx.process_torch(p, rgb)
    torch.cuda.synchronize(); ctx.synchronize()
    for to_rgb in (False, True):
        for _ in range(2):
            out = ctx.expand_planar_torch(p, planar, to_rgb)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out = ctx.expand_planar_torch(p, planar, to_rgb)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        gbs = n * (fb + ow * oh * 3) / (ms / 1e3) / 1e9
        print(f"expand {W}x{H} f={f} 4:{a}:{b} to_rgb={to_rgb}: {ms:.3f} ms, {gbs:.0f} GB/s = {gbs / peak:.3f} of measured peak", flush=True)
    del rgb, planar, out
