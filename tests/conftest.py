import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_png_rgb(name):
    """H x W x 3 uint8; alpha dropped, exactly as pixel.red()/green()/blue() does
    (reference: src/test/scala/jpeg/ImageCompressorTopApp.scala:86-89)."""
    from PIL import Image
    im = Image.open(os.path.join(GOLDEN, name))
    if im.mode not in ("RGB", "RGBA"):
        im = im.convert("RGBA")
    return np.ascontiguousarray(np.asarray(im)[..., :3], dtype=np.uint8)


def manifest():
    return json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))


@pytest.fixture(scope="session")
def golden_manifest():
    return manifest()


def synth_frames(n, h, w, seed):
    """Deterministic uniform bytes: exercises every shift / clamp path (inverse clamps fire ~40 %)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)


ALL_AB = [(4, 4), (4, 0), (2, 2), (2, 0), (1, 1), (1, 0)]       # ChromaSubsampler.scala:17-18
ALL_ORDERS = ["SQC", "SCQ", "QSC", "QCS", "CSQ", "CQS"]         # ImageCompressorTop.scala:27-31
