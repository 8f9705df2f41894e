"""csic_app: the native (C++) host program with the reference's ImageCompressionApp command line
(src/test/scala/jpeg/ImageCompressorTopApp.scala:147-216) on top of the C ABI and a zlib PNG codec."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_png_rgb

APP = os.path.join(ROOT, "chroma-subsampling-image-compressor_b200", "csic_app")


def test_png_codec_roundtrip(tmp_path):
    """read_png handles the reference's RGB and RGBA inputs; write_png_rgb output is readable by Pillow."""
    for name, ch in (("in16x16.png", 3), ("in128x128.png", 4), ("in512x512.png", 3)):
        out = tmp_path / name
        r = subprocess.run([APP, "--selftest-png", os.path.join(GOLDEN, name), str(out)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert f"channels={ch}" in r.stdout
        assert np.array_equal(load_png_rgb(name), np.asarray(__import__("PIL.Image").Image.open(out)))


def test_cli_errors_without_gpu(tmp_path):
    r = subprocess.run([APP, "--input", "nope.png"], capture_output=True, text=True)
    assert r.returncode == 1 and "[ERROR] Input image not found: nope.png" in r.stdout
    r = subprocess.run([APP, "--input", os.path.join(GOLDEN, "in16x16.png"), "--op1", "blur"], capture_output=True, text=True)
    assert r.returncode == 3 and "Unknown processing step: blur" in r.stderr
    r = subprocess.run([APP, "--input", os.path.join(GOLDEN, "in16x16.png"), "--sf", "3", "--outdir", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 3 and "requirement failed: Factor must be 1, 2, 4, or 8" in r.stderr
    # --sf 0: the reference divides by the factor before it constructs anything (ImageCompressorTopApp.scala:44)
    r = subprocess.run([APP, "--input", os.path.join(GOLDEN, "in16x16.png"), "--sf", "0", "--outdir", str(tmp_path)],
                       capture_output=True, text=True)
    assert r.returncode == 3 and "ArithmeticException: / by zero" in r.stderr


def test_png_reader_rejects_malformed_headers(tmp_path):
    """A short IHDR or absurd dimensions end in an error message, not in a crash."""
    import struct, zlib
    def chunk(t, body):
        return struct.pack(">I", len(body)) + t + body + struct.pack(">I", zlib.crc32(t + body))
    sig = b"\x89PNG\r\n\x1a\n"
    short = tmp_path / "short.png"
    short.write_bytes(sig + chunk(b"IHDR", b"\0" * 8) + chunk(b"IEND", b""))
    huge = tmp_path / "huge.png"
    huge.write_bytes(sig + chunk(b"IHDR", struct.pack(">IIBBBBB", 0x7FFFFFFF, 0x7FFFFFFF, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(b"\0")) + chunk(b"IEND", b""))
    for f, text in ((short, "IHDR shorter than 13 bytes"), (huge, "PNG too large")):
        r = subprocess.run([APP, "--input", str(f), "--outdir", str(tmp_path)], capture_output=True, text=True)
        assert r.returncode not in (0, -8, -11, -6) and text in (r.stdout + r.stderr), (r.returncode, r.stdout, r.stderr)


@pytest.mark.gpu
def test_native_app_reproduces_goldens(tmp_path):
    """BASELINE configs[0] (G27) and the 4:2:2 + sf2 top output (G26) through the native CLI."""
    cases = [
        (["--a", "2", "--b", "0", "--yq", "3", "--cbq", "3", "--crq", "2", "--sf", "1", "--op1", "chroma", "--op2", "color",
          "--op3", "spatial"], "in128x128_processed_chroma4-2-0_Y3Cb3Cr2_sf1_order-Ch-Co-Sp.png",
         "G_top_legacy_CHROMA_420_Q_8BIT_sf1_128x128.png"),
        (["--a", "2", "--b", "2", "--yq", "8", "--cbq", "8", "--crq", "8", "--sf", "2", "--op1", "chroma", "--op2", "spatial",
          "--op3", "color"], "in128x128_processed_chroma4-2-2_Y8Cb8Cr8_sf2_order-Ch-Sp-Co.png",
         "G_top_422_Y8Cb8Cr8_sf2_128x128.png"),
    ]
    for flags, out_name, golden in cases:
        r = subprocess.run([APP, "--input", os.path.join(GOLDEN, "in128x128.png"), "--outdir", str(tmp_path)] + flags,
                           capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout + r.stderr
        assert "Image processing complete. Output saved to:" in r.stdout
        got = np.asarray(__import__("PIL.Image").Image.open(tmp_path / out_name))
        assert np.array_equal(got, load_png_rgb(golden))
