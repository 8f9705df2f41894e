"""Single-process multi-GPU end-to-end (csic_multi_*): one host process, one thread + context per GPU.
  python tools/multi_e2e.py [--frames 256]      # frames sharded across all visible GPUs (cfg4 geometry)
Also times ONE 8K frame cut into row bands across the GPUs (cfg5 geometry)."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import csic_b200 as csic
import oracle

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=128)
args = ap.parse_args()
S = csic.ProcessingStep
ops = (S.ChromaSubsampling, S.SpatialSampling, S.ColorQuantization)
ndev = csic.device_count()
for devs in ([0], list(range(ndev))):
    with csic.MultiContext(devs) as m:
        # cfg4 geometry, frames sharded
        W, H, n = 3840, 2160, args.frames
        p = csic.make_params(W, H, 2, 0, 8, 8, 8, 2, ops, out_format=csic.OutFormat.BUNDLE128)
        fb = csic.out_shape(p)[3]
        pin_in, pin_out = csic.PinnedBuffer(n * H * W * 3), csic.PinnedBuffer(n * fb)
        base = np.random.default_rng(1).integers(0, 256, size=(4, H, W, 3), dtype=np.uint8)
        rgb = pin_in.array.reshape(n, H, W, 3)
        for i in range(n):
            rgb[i] = base[i % 4]
        out = pin_out.array.reshape(n, fb)
        m.process_host(p, rgb, out=out)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            m.process_host(p, rgb, out=out)
        dt = (time.perf_counter() - t0) / reps
        want = oracle.process(oracle.make_params(W, H, 2, 0, (8, 8, 8), 2, "CSQ", out_format=3), base[1])
        print(f"gpus={len(devs)} cfg4 x{n} frames: {n * W * H / 1e6 / dt:,.0f} MP/s e2e ({dt * 1e3:.1f} ms)  parity={np.array_equal(out[1], want[0])}")
        pin_in.free(); pin_out.free()
        # cfg5 geometry, ONE frame cut into row bands
        W, H = 7680, 4320
        p = csic.make_params(W, H, 2, 0, 6, 5, 5, 4, ops, out_format=csic.OutFormat.RGB888)
        fb = csic.out_shape(p)[3]
        pin_in, pin_out = csic.PinnedBuffer(H * W * 3), csic.PinnedBuffer(fb)
        rgb = pin_in.array.reshape(1, H, W, 3)
        rgb[0] = np.tile(base[0], (2, 2, 1))
        out = pin_out.array.reshape(1, fb)
        m.process_host(p, rgb, out=out)
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            m.process_host(p, rgb, out=out)
        dt = (time.perf_counter() - t0) / reps
        want = oracle.process(oracle.make_params(W, H, 2, 0, (6, 5, 5), 4, "CSQ", out_format=1), rgb[0])
        print(f"gpus={len(devs)} cfg5 one 8K frame in {len(devs)} row bands: {dt * 1e3:.2f} ms per frame e2e ({W * H / 1e6 / dt:,.0f} MP/s)  parity={np.array_equal(out[0], want[0])}")
        pin_in.free(); pin_out.free()
