"""Pins the oracle: every known-answer vector in the reference's tests and every PNG the reference
committed as an output must be reproduced bit-exactly (SURVEY.md sections 4.3, 8(c) C4).  CPU only."""
import numpy as np
import pytest

import oracle
from oracle import csic_oracle_np as onp
from conftest import ALL_AB, ALL_ORDERS, load_png_rgb, manifest, synth_frames


# ---- RGB2YCbCrTester.scala:12-31 ------------------------------------------------------------------
FWD_KAT = [((0, 0, 0), (0, 128, 128)), ((255, 255, 255), (255, 128, 128)), ((255, 0, 0), (77, 85, 255)),
           ((0, 255, 0), (149, 43, 21)), ((0, 0, 255), (29, 255, 107))]
FWD_KAT_TRUNC = [((255, 0, 0), (77, 86, 255)), ((0, 255, 0), (149, 44, 22)), ((0, 0, 255), (29, 255, 108))]


@pytest.mark.parametrize("rgb,ycc", FWD_KAT)
def test_forward_kat_floor(rgb, ycc):
    assert oracle.rgb2ycbcr(*rgb, 0) == ycc
    assert tuple(onp.forward(np.array(rgb, np.uint8), 0)) == ycc


@pytest.mark.parametrize("rgb,ycc", FWD_KAT_TRUNC)
def test_forward_kat_trunc(rgb, ycc):
    assert oracle.rgb2ycbcr(*rgb, 1) == ycc
    assert tuple(onp.forward(np.array(rgb, np.uint8), 1)) == ycc


def test_forward_exhaustive_c_vs_numpy_and_ranges():
    """All 2^24 colours, both roundings: C streaming form == NumPy form; the only clamp that fires is
    256 -> 255, once per chroma channel; floor and trunc differ on 12,472,897 colours (F3)."""
    g, b = np.mgrid[0:256, 0:256].astype(np.uint8)
    ndiff = 0
    for r in range(256):
        rgb = np.stack([np.full_like(g, r), g, b], -1)                    # 256 x 256 x 3
        p = oracle.make_params(256, 256)
        outs = []
        for mode in (0, 1):
            p.round_mode = mode
            c = oracle.process(p, rgb).reshape(256, 256, 3)
            n = onp.forward(rgb, mode)
            assert np.array_equal(c, n)
            outs.append(c)
        ndiff += int((outs[0] != outs[1]).any(-1).sum())
    assert ndiff == 12472897
    assert oracle.rgb2ycbcr(0, 0, 255, 0)[1] == 255 and oracle.rgb2ycbcr(255, 0, 0, 0)[2] == 255


# ---- derived inverse KATs ------------------------------------------------------------
INV_KAT = [((0, 128, 128), (0, 0, 0)), ((255, 128, 128), (255, 255, 255)), ((77, 85, 255), (255, 3, 3)),
           ((149, 43, 21), (2, 255, 2)), ((29, 255, 107), (0, 1, 255)), ((16, 128, 128), (19, 19, 19))]


@pytest.mark.parametrize("ycc,rgb", INV_KAT)
def test_inverse_kat(ycc, rgb):
    assert oracle.ycbcr2rgb(*ycc) == rgb
    assert tuple(onp.inverse(np.array(ycc, np.uint8))) == rgb


# ---- ColorQuantizerSpec.scala:43-61 ---------------------------------------------------------------
Q_PIX = [(0, 0, 0), (255, 255, 255), (128, 128, 128), (77, 150, 29), (200, 50, 220), (16, 16, 16), (235, 240, 240)]
Q_KAT = {  # (yBits, cbBits, crBits) -> expected for pixels 3, 4, 6
    (6, 5, 5): [(76, 144, 24), (200, 48, 216), (232, 240, 240)],
    (3, 3, 2): [(64, 128, 0), (192, 32, 192), (224, 224, 192)],
    (8, 1, 1): [(77, 128, 0), (200, 0, 128), (235, 128, 128)],
    (1, 8, 8): [(0, 150, 29), (128, 50, 220), (128, 240, 240)],
    (4, 4, 4): [(64, 144, 16), (192, 48, 208), (224, 240, 240)],
    (8, 8, 8): [(77, 150, 29), (200, 50, 220), (235, 240, 240)],
}


@pytest.mark.parametrize("bits", list(Q_KAT))
def test_quant_kat(bits):
    got = oracle.quant_stream(np.array(Q_PIX, np.uint8), *bits)
    assert [tuple(int(v) for v in got[i]) for i in (3, 4, 6)] == Q_KAT[bits]
    # quantizePixelSW (ColorQuantizerSpec.scala:19-40) for every pixel
    for px, g in zip(Q_PIX, got):
        assert tuple(g) == tuple((v >> (8 - t)) << (8 - t) for v, t in zip(px, bits))


# ---- SpatialDownsamplerSpec.scala:20-46, 60-88, 90-118, 120-145 -----------------------------------
def _idx_stream(n):
    s = np.zeros((n, 3), np.uint8)
    s[:, 0] = np.arange(n) % 256
    return s


@pytest.mark.parametrize("W,H,f,expected", [
    (4, 4, 2, [0, 2, 8, 10]),
    (8, 8, 4, [r * 8 + c for r in range(0, 8, 4) for c in range(0, 8, 4)]),
    (16, 16, 8, [r * 16 + c for r in range(0, 16, 8) for c in range(0, 16, 8)]),
    (5, 3, 2, [0, 2, 4, 10, 12, 14]),                      # non-power-of-two: ceil dims
])
def test_spatial_kat(W, H, f, expected):
    out = oracle.spatial_stream(_idx_stream(W * H), W, H, f)
    assert list(out[:, 0]) == expected


# ---- ChromaSubsampler 4x4 KATs -----------------------------------------------------
def _chroma_idx(W, H, a, b, n=None):
    n = W * H if n is None else n
    s = np.zeros((n, 3), np.uint8)
    s[:, 1] = np.arange(n)
    return oracle.chroma_stream(s, W, H, a, b)[:, 1].reshape(-1)


def test_chroma_kat_4x4():
    assert list(_chroma_idx(4, 4, 2, 0)) == [0, 0, 2, 2, 2, 2, 2, 2, 8, 8, 10, 10, 10, 10, 10, 10]
    assert list(_chroma_idx(4, 4, 1, 1)) == [0] * 4 + [4] * 4 + [8] * 4 + [12] * 4
    assert list(_chroma_idx(4, 4, 1, 0)) == [0] * 8 + [8] * 8
    assert list(_chroma_idx(4, 4, 2, 2)) == [0, 0, 2, 2, 4, 4, 6, 6, 8, 8, 10, 10, 12, 12, 14, 14]
    assert list(_chroma_idx(4, 4, 4, 4)) == list(range(16))


def test_chroma_after_spatial_kat_8x8():
    """Case B: 8x8, f=2, 4:2:0, spatial before chroma; indices of the original pixels."""
    s = np.zeros((64, 3), np.uint8)
    s[:, 0] = np.arange(64)
    s[:, 1] = np.arange(64)
    d = oracle.spatial_stream(s, 8, 8, 2)
    o = oracle.chroma_stream(d, 8, 8, 2, 0)
    assert o[:, 0].reshape(4, 4).tolist() == [[0, 2, 4, 6], [16, 18, 20, 22], [32, 34, 36, 38], [48, 50, 52, 54]]
    assert o[:, 1].reshape(4, 4).tolist() == [[0, 0, 4, 4], [16, 16, 20, 20], [20] * 4, [20] * 4]


# ---- the 29 committed PNGs ------------------------------------------------------------------------
@pytest.mark.parametrize("entry", manifest()["goldens"], ids=lambda e: e["file"])
def test_golden_png(entry):
    rgb = load_png_rgb(entry["input"])
    want = load_png_rgb(entry["file"])
    if entry["forward"] == "identity":
        assert np.array_equal(rgb, want)
        return
    H, W = rgb.shape[:2]
    p = oracle.make_params(W, H, entry["a"], entry["b"], tuple(entry["q"]), entry["factor"], entry["order"],
                           round_mode=0 if entry["forward"] == "floor" else 1, out_format=1)
    f = entry["factor"]
    got = oracle.process(p, rgb).reshape(H // f, W // f, 3)
    assert got.shape == want.shape
    assert np.array_equal(got, want), f"{int((got != want).any(-1).sum())} pixels differ"
    assert np.array_equal(onp.process(p, rgb).reshape(want.shape), want)


def test_golden_g26_order_sensitivity():
    """G26 is reproduced by the three chroma-before-spatial orders only (595 pixels differ otherwise)."""
    e = [g for g in manifest()["goldens"] if g["file"].startswith("G_top_422")][0]
    rgb, want = load_png_rgb(e["input"]), load_png_rgb(e["file"])
    for order in ALL_ORDERS:
        p = oracle.make_params(128, 128, 2, 2, (8, 8, 8), 2, order, out_format=1)
        diff = int((oracle.process(p, rgb).reshape(64, 64, 3) != want).any(-1).sum())
        assert diff == (0 if order in ("CSQ", "CQS", "QCS") else 595), (order, diff)


# ---- closed form == streaming state machines ------------------------------------------------------
SIZES = [(16, 16), (32, 8), (8, 32), (24, 10), (5, 3), (7, 9), (33, 17), (64, 2), (1, 1), (3, 1), (1, 5)]


@pytest.mark.parametrize("ab", ALL_AB)
@pytest.mark.parametrize("order", ALL_ORDERS)
def test_closed_form_matches_streaming(ab, order):
    for (W, H) in SIZES:
        rgb = synth_frames(1, H, W, seed=W * 131 + H)
        for f in (1, 2, 4, 8):
            for fmt, q in ((0, (8, 8, 8)), (1, (6, 5, 5)), (2, (3, 3, 2)), (3, (8, 8, 8)), (3, (5, 4, 4))):
                p = oracle.make_params(W, H, ab[0], ab[1], q, f, order, out_format=fmt)
                a, b = oracle.process(p, rgb), onp.process(p, rgb)
                assert np.array_equal(a, b), (W, H, f, fmt, q)


@pytest.mark.parametrize("order", ALL_ORDERS)
def test_closed_form_matches_streaming_average_ext(order):
    for (W, H) in [(16, 16), (32, 8), (8, 32), (24, 16)]:
        rgb = synth_frames(1, H, W, seed=7)
        for ab in ALL_AB:
            for f in (2, 4, 8):
                p = oracle.make_params(W, H, ab[0], ab[1], (5, 4, 3), f, order, pool_mode=1)
                assert np.array_equal(oracle.process(p, rgb), onp.process(p, rgb)), (W, H, ab, f)


def test_quant_position_is_irrelevant_decimate():
    """ColorQuantizer is pointwise: only chroma-vs-spatial order matters."""
    rgb = synth_frames(2, 16, 24, seed=3)
    for ab in ALL_AB:
        for f in (1, 2, 4):
            outs = {o: oracle.process(oracle.make_params(24, 16, *ab, (4, 3, 5), f, o), rgb) for o in ALL_ORDERS}
            assert np.array_equal(outs["CSQ"], outs["CQS"]) and np.array_equal(outs["CSQ"], outs["QCS"])
            assert np.array_equal(outs["SQC"], outs["SCQ"]) and np.array_equal(outs["SQC"], outs["QSC"])
            if f == 1:
                assert np.array_equal(outs["CSQ"], outs["SQC"])


def test_frames_are_independent_and_threads_agree():
    rgb = synth_frames(5, 16, 16, seed=11)
    p = oracle.make_params(16, 16, 2, 0, (6, 5, 5), 2, "SQC")
    whole = oracle.process(p, rgb, threads=3)
    for k in range(5):
        assert np.array_equal(whole[k], oracle.process(p, rgb[k])[0])
