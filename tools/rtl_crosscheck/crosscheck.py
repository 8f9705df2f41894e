#!/usr/bin/env python3
"""Diff the reference's real RTL simulation against the oracle (and the GPU path) on random small images.
Needs a JDK + sbt and a checkout of the reference; see README.md in this directory.  Not runnable in the
authoring container (no JVM) -- which is exactly why it exists."""
import argparse, os, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
STEP = {1: "spatial", 2: "color", 3: "chroma"}
TAG = {1: "Sp", 2: "Co", 3: "Ch"}
ORDERS = {"SQC": (1, 2, 3), "SCQ": (1, 3, 2), "QSC": (2, 1, 3), "QCS": (2, 3, 1), "CSQ": (3, 1, 2), "CQS": (3, 2, 1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", required=True)
    ap.add_argument("--cases", type=int, default=20)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--gpu", action="store_true", help="also compare csic_b200 (needs a B200)")
    ap.add_argument("--runner", default="sbt",
                    help="program that runs the reference CLI in --reference: called as <runner> '<sbt command line>'.  "
                         "Default sbt; the self-test passes tools/rtl_crosscheck/fake_reference.py (no JVM needed)")
    args = ap.parse_args()
    from PIL import Image
    import oracle
    rng = np.random.default_rng(args.seed)
    bad = 0
    for case in range(args.cases):
        f = int(rng.choice([1, 2, 4, 8]))
        W, H = f * int(rng.integers(1, 9)), f * int(rng.integers(1, 9))
        a, b = [(4, 4), (4, 0), (2, 2), (2, 0), (1, 1), (1, 0)][int(rng.integers(0, 6))]
        q = [int(v) for v in rng.integers(1, 9, size=3)]
        name, ops = list(ORDERS.items())[int(rng.integers(0, 6))]
        rgb = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        with tempfile.TemporaryDirectory() as td:
            inp = os.path.join(td, f"case{case}.png")
            Image.fromarray(rgb, "RGB").save(inp)
            cli = (f"Test / runMain jpeg.ImageCompressionApp --input {inp} --a {a} --b {b} --yq {q[0]} --cbq {q[1]} "
                   f"--crq {q[2]} --sf {f} --op1 {STEP[ops[0]]} --op2 {STEP[ops[1]]} --op3 {STEP[ops[2]]}")
            subprocess.run(args.runner.split() + [cli], cwd=args.reference, check=True, capture_output=True)
            # the order tag is `opN.toString.split('.').last.take(2)` (ImageCompressorTopApp.scala:188): "Sp-Co-Ch" as
            # intended, but "Pr-Pr-Pr" with Chisel versions whose ChiselEnum prints "ProcessingStep(1=SpatialSampling)"
            # (the reference's own committed APP_OUTPUT file is named that way) -- accept either
            import glob
            pat = os.path.join(args.reference, "APP_OUTPUT",
                               f"case{case}_processed_chroma4-{a}-{b}_Y{q[0]}Cb{q[1]}Cr{q[2]}_sf{f}_order-*.png")
            hits = sorted(glob.glob(pat), key=os.path.getmtime)
            if not hits:
                sys.exit(f"the reference wrote no {pat}")
            rtl = np.asarray(Image.open(hits[-1]).convert("RGB"))
            for h in hits:
                os.remove(h)
        po = oracle.make_params(W, H, a, b, tuple(q), f, name, out_format=1)
        want = oracle.process(po, rgb).reshape(H // f, W // f, 3)
        ok = np.array_equal(rtl, want)
        if args.gpu:
            import csic_b200 as csic
            top = csic.ImageCompressorTop(W, H, a, b, *q, f, *ops, out_format=csic.OutFormat.RGB888)
            ok = ok and np.array_equal(top.process(rgb)[0], rtl)
        print(f"case {case}: {W}x{H} 4:{a}:{b} q={q} f={f} {name} -> {'ok' if ok else 'MISMATCH'}")
        bad += not ok
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
