#!/usr/bin/env python3
"""Stand-in for `sbt "Test / runMain jpeg.ImageCompressionApp ..."` so that crosscheck.py's own plumbing (CLI flags,
PNG in, APP_OUTPUT naming, PNG out, comparison) can be exercised where no JVM exists (tests/test_host_logic.py).
TEST INFRASTRUCTURE: it answers with the CPU oracle, so the comparison it feeds is trivially equal -- except with
--corrupt, which flips one output byte to prove that crosscheck.py reports a mismatch.  The real cross-check needs the
real reference: `crosscheck.py --reference <checkout> --runner sbt`."""
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle

argv = sys.argv[1:]
corrupt = "--corrupt" in argv
line = [a for a in argv if a != "--corrupt"][-1]                      # the sbt command line, one string
tok = line.split()
assert tok[:4] == ["Test", "/", "runMain", "jpeg.ImageCompressionApp"], tok[:4]
kv = dict(zip(tok[4::2], tok[5::2]))
step = {"spatial": "S", "color": "Q", "chroma": "C"}
order = "".join(step[kv[k]] for k in ("--op1", "--op2", "--op3"))
a, b, f = int(kv["--a"]), int(kv["--b"]), int(kv["--sf"])
q = tuple(int(kv[k]) for k in ("--yq", "--cbq", "--crq"))
rgb = np.asarray(Image.open(kv["--input"]).convert("RGB"))
H, W = rgb.shape[:2]
out = oracle.process(oracle.make_params(W, H, a, b, q, f, order, out_format=1), rgb).reshape(H // f, W // f, 3).copy()
if corrupt:
    out[0, 0, 0] ^= 1
name = os.path.basename(kv["--input"]).split(".")[0]
os.makedirs("APP_OUTPUT", exist_ok=True)
Image.fromarray(out, "RGB").save(f"APP_OUTPUT/{name}_processed_chroma4-{a}-{b}_Y{q[0]}Cb{q[1]}Cr{q[2]}_sf{f}_order-Pr-Pr-Pr.png")
