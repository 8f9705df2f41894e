"""Small driver for compute-sanitizer: every kernel family / format / factor once on tiny frames, checked
against the oracle.  Usage (on the GPU box): compute-sanitizer --tool memcheck python tools/sanitize.py"""
import itertools, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import csic_b200 as csic
import oracle

rng = np.random.default_rng(1)
ORD = {"S": 1, "Q": 2, "C": 3}
n_ok = 0
with csic.Context(0) as ctx:
    for (W, H), f, (a, b), order, (fmt, q) in itertools.product(
            [(128, 24), (64, 16), (40, 9)], (1, 2, 4, 8), [(4, 4), (2, 0), (1, 0)], ("CSQ", "SQC"),
            [(0, (8, 8, 8)), (1, (6, 5, 5)), (2, (3, 3, 2)), (3, (6, 5, 5)), (3, (8, 8, 8))]):
        rgb = rng.integers(0, 256, size=(5, H, W, 3), dtype=np.uint8)
        p = csic.make_params(W, H, a, b, *q, f, tuple(ORD[c] for c in order), out_format=fmt)
        want = oracle.process(oracle.make_params(W, H, a, b, q, f, order, out_format=fmt), rgb)
        for fam in (0, 1):
            ctx.set_option(0, fam)
            got = ctx.process_host(p, rgb)
            assert np.array_equal(got, want), (W, H, f, a, b, order, fmt, q, fam)
            n_ok += 1
print("sanitize driver ok:", n_ok, "launch configurations")
