"""csic_process_host with pageable vs pinned host buffers (one GPU, cfg4 geometry)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import csic_b200 as csic
S = csic.ProcessingStep
W, H, n = 3840, 2160, 96
p = csic.make_params(W, H, 2, 0, 8, 8, 8, 2, (S.ChromaSubsampling, S.SpatialSampling, S.ColorQuantization), out_format=csic.OutFormat.BUNDLE128)
fb = csic.out_shape(p)[3]
base = np.random.default_rng(1).integers(0, 256, size=(4, H, W, 3), dtype=np.uint8)
with csic.Context(0) as ctx:
    for kind in ("pinned", "pageable"):
        if kind == "pinned":
            pi, po = csic.PinnedBuffer(n * H * W * 3), csic.PinnedBuffer(n * fb)
            rgb, out = pi.array.reshape(n, H, W, 3), po.array.reshape(n, fb)
        else:
            rgb, out = np.empty((n, H, W, 3), np.uint8), np.empty((n, fb), np.uint8)
        for i in range(n):
            rgb[i] = base[i % 4]
        ctx.process_host(p, rgb, out=out)
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.process_host(p, rgb, out=out)
        dt = (time.perf_counter() - t0) / 3
        print(f"{kind}: {n * W * H / 1e6 / dt:,.0f} MP/s e2e ({dt * 1e3:.1f} ms for {n} frames)")
