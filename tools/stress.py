"""One-off randomised stress on the GPU (run under gpurun; not part of the test suite): for ~N seconds each,
(1) forward: random geometry / mode / order / factor / format / input format / frame count (a third of the cases: a random
    row band into a canary-filled buffer), the automatic kernel choice against the gather kernel (family option 1) on the
    same device buffers, byte for byte;
(2) decoder: random planes through csic_expand_planar_device against a torch gather.
(3) host path: csic_process_host (pageable NumPy buffers, random chunk size -> the 3-stream chunk pipeline, re-pitching of
    odd widths, every-f-th-row shipping) and csic_process_host_band on a random band, against the device path.
Prints the kernel families seen and the first mismatch, exits non-zero on one."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import csic_b200 as csic

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1234)
ctx = csic.Context(0)
ORD = {"S": 1, "Q": 2, "C": 3}
AB = [(4, 4), (4, 0), (2, 2), (2, 0), (1, 1), (1, 0)]
ORDERS = ["SQC", "SCQ", "QSC", "QCS", "CSQ", "CQS"]
seen, n_fwd, n_dec = {}, 0, 0
t0 = time.time()
while time.time() - t0 < secs:
    f = int(rng.choice([1, 1, 2, 2, 4, 8]))
    kind = rng.integers(0, 4)
    if kind == 0:      # aligned, row-kernel shapes
        W, H = 16 * f * int(rng.integers(1, 40)), f * 2 * int(rng.integers(1, 60))
    elif kind == 1:    # anything
        W, H = int(rng.integers(1, 700)), int(rng.integers(1, 300))
    elif kind == 2:    # tiny frames, many of them
        W, H = int(rng.integers(1, 9)) * 16, int(rng.integers(1, 70))
    else:              # wide
        W, H = int(rng.integers(2000, 9000)), int(rng.integers(1, 12))
    a, b = AB[rng.integers(0, len(AB))]
    order = ORDERS[rng.integers(0, 6)]
    fmt = int(rng.choice([0, 1, 2, 3, 4]))
    q = [(8, 8, 8), (6, 5, 5), (3, 3, 2), (4, 4, 4), (8, 7, 8)][rng.integers(0, 5)]
    inf = int(rng.choice([0, 0, 1, 2]))
    pool = int(rng.choice([0, 0, 1]))
    rm = int(rng.integers(0, 2))
    n = int(rng.choice([1, 2, 3, 7, 33, 150])) if W * H < 40000 else int(rng.choice([1, 2, 5]))
    try:
        p = csic.make_params(W, H, a, b, q[0], q[1], q[2], f, tuple(ORD[c] for c in order), rm, pool, fmt, inf)
    except csic.IllegalArgumentException:
        continue
    ch = 3 if inf == 0 else 4
    rgb = torch.randint(0, 256, (n, H, W, ch), dtype=torch.uint8, device="cuda")
    oh, fb = csic.out_shape(p)[1], csic.out_shape(p)[3]
    band = (None, None)
    if fmt != 4 and oh >= 2 and rng.integers(0, 3) == 0:        # a third of the cases: a random row band into a canary-filled buffer
        r0 = int(rng.integers(0, oh - 1))
        band = (r0, int(rng.integers(1, oh - r0 + 1)))
    ctx.set_option(0, 0)
    out = ctx.process_torch(p, rgb, torch.full((n, fb), 0xA5, dtype=torch.uint8, device="cuda"), *band)
    ctx.synchronize()
    fam = ctx.last_kernel()[0]
    seen[fam] = seen.get(fam, 0) + 1
    ctx.set_option(0, 1)
    ref = ctx.process_torch(p, rgb, torch.full((n, fb), 0xA5, dtype=torch.uint8, device="cuda"), *band)
    ctx.synchronize()
    ctx.set_option(0, 0)
    n_fwd += 1
    if not torch.equal(out, ref):
        bad = int((out != ref).flatten().nonzero()[0])
        print("FORWARD MISMATCH", dict(W=W, H=H, a=a, b=b, order=order, fmt=fmt, q=q, inf=inf, pool=pool, rm=rm, n=n, f=f, fam=fam, band=band, first=bad))
        sys.exit(1)
print(f"forward: {n_fwd} cases, kernel families {seen}: all equal to the gather kernel", flush=True)

t0 = time.time()
while time.time() - t0 < secs:
    f = int(rng.choice([1, 1, 2]))
    kind = rng.integers(0, 3)
    if kind == 0:
        W, H = 16 * f * int(rng.integers(1, 130)), f * int(rng.integers(1, 80))
    elif kind == 1:
        W, H = f * int(rng.integers(1, 1200)), f * int(rng.integers(1, 200))
    else:
        W, H = f * int(rng.integers(1, 9)) * 4, f * int(rng.integers(1, 70))
    a, b = AB[rng.integers(0, len(AB))]
    n = int(rng.choice([1, 2, 9, 64, 300])) if W * H < 30000 else int(rng.choice([1, 3, 6]))
    try:
        p = csic.make_params(W, H, a, b, 8, 8, 8, f, (3, 1, 2), 0, 0, 4, 0)
    except csic.IllegalArgumentException:
        continue
    w, h, _, fb = csic.out_shape(p)
    cw, chh, ob, orr = csic.planar_shape(p)
    hf, vf = 4 // a, (2 if b == 0 else 1)
    hs, vs = max(1, hf // f), max(1, vf // f)
    last_c = (((W - 1) // hf) * hf // f) // hs
    planar = torch.randint(0, 256, (n, fb), dtype=torch.uint8, device="cuda")
    rows = torch.arange(h, device="cuda")
    held = (rows & 1).bool() if vs == 2 else torch.zeros(h, dtype=torch.bool, device="cuda")
    crow = (rows - held.long()) // vs
    ccol = (torch.arange(w, device="cuda") // hs)[None, :].expand(h, w).clone()
    ccol[held] = last_c
    idx = (crow[:, None] * cw + ccol).reshape(-1)
    want = torch.stack([planar[:, :w * h], planar[:, ob:ob + cw * chh][:, idx], planar[:, orr:orr + cw * chh][:, idx]], -1)
    got = ctx.expand_planar_torch(p, planar, to_rgb=False)
    ctx.synchronize()
    n_dec += 1
    if not torch.equal(got.reshape(n, -1, 3), want):
        print("DECODER MISMATCH", dict(W=W, H=H, a=a, b=b, f=f, n=n))
        sys.exit(1)
print(f"decoder: {n_dec} cases: all equal to the torch gather")

t0 = time.time()
n_host = 0
while time.time() - t0 < secs / 2:
    f = int(rng.choice([1, 2, 2, 4, 8]))
    W, H = (16 * f * int(rng.integers(1, 60)), f * int(rng.integers(1, 120))) if rng.integers(0, 2) else (int(rng.integers(1, 900)), int(rng.integers(1, 200)))
    a, b = AB[rng.integers(0, len(AB))]
    order = ORDERS[rng.integers(0, 6)]
    fmt = int(rng.choice([0, 1, 2, 3]))
    q = [(8, 8, 8), (6, 5, 5), (3, 3, 2)][rng.integers(0, 3)]
    inf = int(rng.choice([0, 0, 2]))
    pool = int(rng.choice([0, 0, 1]))
    n = int(rng.choice([1, 3, 17, 64]))
    if n * W * H > 3e7:
        n = 2
    try:
        p = csic.make_params(W, H, a, b, q[0], q[1], q[2], f, tuple(ORD[c] for c in order), 0, pool, fmt, inf)
    except csic.IllegalArgumentException:
        continue
    ch = 3 if inf == 0 else 4
    rgb = rng.integers(0, 256, size=(n, H, W, ch), dtype=np.uint8)
    ctx.set_option(0, 0)
    ref = ctx.process_torch(p, torch.from_numpy(rgb).cuda()).cpu().numpy()
    ctx.synchronize()
    ctx.set_option(1, int(rng.choice([0, 1 << 16, 1 << 20, 3 << 20, 16 << 20])))     # host chunk bytes
    got = ctx.process_host(p, rgb)
    ok = np.array_equal(got.reshape(ref.shape), ref)
    oh = csic.out_shape(p)[1]
    if ok and oh >= 2:
        r0 = int(rng.integers(0, oh - 1)); nr = int(rng.integers(1, oh - r0 + 1))
        band = np.full_like(got, 0xA5)
        ctx.process_host_band(p, rgb, band, r0, nr)
        rb = csic.out_shape(p)[2]
        bv, rv = band.reshape(n, oh, rb), ref.reshape(n, oh, rb)
        ok = np.array_equal(bv[:, r0:r0 + nr], rv[:, r0:r0 + nr]) and bool((bv[:, :r0] == 0xA5).all()) and bool((bv[:, r0 + nr:] == 0xA5).all())
    ctx.set_option(1, 0)
    n_host += 1
    if not ok:
        print("HOST PATH MISMATCH", dict(W=W, H=H, a=a, b=b, order=order, fmt=fmt, q=q, inf=inf, pool=pool, n=n, f=f))
        sys.exit(1)
print(f"host path: {n_host} cases (chunked pipeline + a random band each): all equal to the device path")
