"""Single-process multi-GPU end-to-end (csic_multi_*): one host process, one thread + context per GPU -- the shape of the
reference's host (one JVM).  Run under gpurun --gpus N:
  python tools/multi_e2e.py [--frames 512] [--json gpurun_out/multi_e2e.json]
cfg4 geometry, pinned host buffers.  Compares, on the same batch:
  * one GPU;
  * all GPUs with the even frame split of round 1 (CSIC_OPT_MULTI_STATIC_SPLIT);
  * all GPUs pulling chunks from the shared cursor (default): GPUs behind faster host links take more of the batch;
  * the second half of the GPUs only (on this pool's 8-GPU boxes GPUs 4-7 have the better path to host memory and
    are SLOWED DOWN when GPUs 0-3 copy at the same time: profiles/r2/pcie_ceiling.json).
Also times ONE 8K frame cut into row bands across the GPUs (cfg5 geometry)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import csic_b200 as csic
import oracle

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=512)
ap.add_argument("--json", default="")
args = ap.parse_args()
S = csic.ProcessingStep
ops = (S.ChromaSubsampling, S.SpatialSampling, S.ColorQuantization)
ndev = csic.device_count()
W, H, n = 3840, 2160, args.frames
p = csic.make_params(W, H, 2, 0, 8, 8, 8, 2, ops, out_format=csic.OutFormat.BUNDLE128)
fb = csic.out_shape(p)[3]
pin_in, pin_out = csic.PinnedBuffer(n * H * W * 3), csic.PinnedBuffer(n * fb)
base = np.random.default_rng(1).integers(0, 256, size=(4, H, W, 3), dtype=np.uint8)
rgb = pin_in.array.reshape(n, H, W, 3)
for i in range(n):
    rgb[i] = base[i % 4]
out = pin_out.array.reshape(n, fb)
want = oracle.process(oracle.make_params(W, H, 2, 0, (8, 8, 8), 2, "CSQ", out_format=3), base[1])
results = []
cases = [("one GPU", [0], 0)]
if ndev > 1:
    cases += [(f"{ndev} GPUs, even split", list(range(ndev)), 1), (f"{ndev} GPUs, shared cursor", list(range(ndev)), 0)]
if ndev >= 4:
    half = list(range(ndev // 2, ndev))
    cases += [(f"GPUs {half}, shared cursor", half, 0), (f"GPUs {list(range(ndev // 2))}, shared cursor", list(range(ndev // 2)), 0)]
for name, devs, static in cases:
    with csic.MultiContext(devs) as m:
        m.set_option(csic.MultiContext.STATIC_SPLIT, static)
        out[:] = 0
        m.process_host(p, rgb, out=out)
        b0 = m.host_bytes()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            m.process_host(p, rgb, out=out)
        dt = (time.perf_counter() - t0) / reps
        share = [round((b - a) / max(1, sum(m.host_bytes()) - sum(b0)), 3) for a, b in zip(b0, m.host_bytes())]
        ok = bool(np.array_equal(out[1], want[0]) and np.array_equal(out[n - 3], oracle.process(oracle.make_params(W, H, 2, 0, (8, 8, 8), 2, "CSQ", out_format=3), base[(n - 3) % 4])[0]))
        r = {"case": name, "devices": devs, "static_split": bool(static), "frames": n, "mp_per_s": round(n * W * H / 1e6 / dt, 1),
             "ms": round(dt * 1e3, 2), "share_of_bytes_per_gpu": share, "parity": ok}
        results.append(r)
        print(json.dumps(r), flush=True)
pin_in.free(); pin_out.free()
# cfg5 geometry, ONE frame cut into row bands
W, H = 7680, 4320
p = csic.make_params(W, H, 2, 0, 6, 5, 5, 4, ops, out_format=csic.OutFormat.RGB888)
fb = csic.out_shape(p)[3]
pin_in, pin_out = csic.PinnedBuffer(H * W * 3), csic.PinnedBuffer(fb)
rgb = pin_in.array.reshape(1, H, W, 3)
rgb[0] = np.tile(base[0], (2, 2, 1))
out = pin_out.array.reshape(1, fb)
want = oracle.process(oracle.make_params(W, H, 2, 0, (6, 5, 5), 4, "CSQ", out_format=1), rgb[0])
for devs in ([0], list(range(ndev))):
    with csic.MultiContext(devs) as m:
        m.process_host(p, rgb, out=out)
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            m.process_host(p, rgb, out=out)
        dt = (time.perf_counter() - t0) / reps
        r = {"case": f"one 8K frame in {len(devs)} row band(s)", "devices": devs, "ms_per_frame": round(dt * 1e3, 3),
             "mp_per_s": round(W * H / 1e6 / dt, 1), "parity": bool(np.array_equal(out[0], want[0]))}
        results.append(r)
        print(json.dumps(r), flush=True)
    if ndev == 1:
        break
pin_in.free(); pin_out.free()
if args.json:
    json.dump(results, open(args.json, "w"), indent=1)
