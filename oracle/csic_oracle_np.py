"""Closed-form NumPy twin of the oracle -- TEST INFRASTRUCTURE ONLY (same import rules as oracle/).

csic_oracle.c restates the reference as sequential state machines.  This module restates it as
*closed-form gathers* (SURVEY.md section 8(a) row A3), i.e. the same index algebra the CUDA kernels use,
so `tests/test_oracle_golden.py::test_closed_form_matches_streaming` proves on the CPU that the
algebra is equivalent to the RTL's counters before any kernel is trusted with it.

Reference citations are relative to the reference root, src/main/scala/jpeg/.
"""
import numpy as np


def forward(rgb, round_mode=0):
    """RGB2YCbCr.scala:33-35,55-65,74-76 (FLOOR) / :95-121 (TRUNC).  rgb [...,3] uint8 -> ycc [...,3] uint8."""
    v = rgb.astype(np.int32)
    r, g, b = v[..., 0], v[..., 1], v[..., 2]
    yi = 77 * r + 150 * g + 29 * b + 128
    cbi = -43 * r - 85 * g + 128 * b + 128
    cri = 128 * r - 107 * g - 21 * b + 128
    if round_mode == 0:
        y, cb, cr = yi >> 8, (cbi >> 8) + 128, (cri >> 8) + 128          # numpy >> on int32 is arithmetic
    else:
        tr = lambda t: np.sign(t) * (np.abs(t) // 256)                  # Scala Int `/`: toward zero
        y, cb, cr = tr(yi), tr(cbi) + 128, tr(cri) + 128
    return np.clip(np.stack([y, cb, cr], -1), 0, 255).astype(np.uint8)


def inverse(ycc):
    """YCbCrUtils.ycbcr2rgb, RGB2YCbCr.scala:123-132."""
    v = ycc.astype(np.int32)
    c, d, e = v[..., 0], v[..., 1] - 128, v[..., 2] - 128
    r = (298 * c + 409 * e + 128) >> 8
    g = (298 * c - 100 * d - 208 * e + 128) >> 8
    b = (298 * c + 516 * d + 128) >> 8
    return np.clip(np.stack([r, g, b], -1), 0, 255).astype(np.uint8)


def quant(ycc, q):
    """ColorQuantizer.scala:29-31,42-44."""
    s = 8 - np.asarray(q, dtype=np.int32)
    return ((ycc.astype(np.int32) >> s) << s).astype(np.uint8)


def chroma_source_full(W, H, a, b):
    """Chroma source (row, col) for every pixel of a full-resolution W x H stream.
    ChromaSubsampler.scala:26-27,37-38,52-65 in closed form."""
    hf, vf = 4 // a, (2 if b == 0 else 1)
    r, c = np.mgrid[0:H, 0:W]
    held = (r % vf) != 0                       # vf == 2 and odd line: nothing sampled on this line
    last = ((W - 1) // hf) * hf                # last sample column of the previous line
    return np.where(held, r - 1, r), np.where(held, last, c - c % hf)


def source_maps(W, H, a, b, f, chroma_first):
    """For each output pixel: (row, col) of the input pixel supplying Y, and of the one supplying Cb/Cr
    (DECIMATE).  chroma_first = ChromaSubsampling precedes SpatialSampling in op1..op3."""
    hf, vf = 4 // a, (2 if b == 0 else 1)
    Wo, Ho = -(-W // f), -(-H // f)
    ro, co = np.mgrid[0:Ho, 0:Wo]
    yr, yc = ro * f, co * f
    if chroma_first or f == 1:
        fr, fc = chroma_source_full(W, H, a, b)
        return yr, yc, fr[yr, yc], fc[yr, yc]
    # spatial first: the chroma stage sees the short stream but still counts with the full W, H
    # (ImageCompressorTop.scala:52-58)
    m = ro * Wo + co
    col, line = m % W, (m // W) % H
    held = (line % vf) != 0
    src = np.where(held, (line - 1) * W + ((W - 1) // hf) * hf, m - col % hf)
    return yr, yc, (src // Wo) * f, (src % Wo) * f


def _avg(x, f):
    H, W = x.shape[:2]
    s = x.astype(np.int32).reshape(H // f, f, W // f, f, -1).sum(axis=(1, 3))
    return ((s + (f * f) // 2) >> (2 * (f.bit_length() - 1))).astype(np.uint8)


def planar_factors(a, b, f):
    hf, vf = 4 // a, (2 if b == 0 else 1)
    return max(1, hf // f), max(1, vf // f)


def expand_planar(planar, W, H, a, b, f, to_rgb=False):
    """Decoder of the PLANAR format: re-applies the replay rule of ChromaSubsampler.scala:52-65 in output
    coordinates (chroma before spatial, or f == 1).  planar: flat uint8 of one frame.  Returns [Ho, Wo, 3]."""
    hf, vf = 4 // a, (2 if b == 0 else 1)
    hs, vs = planar_factors(a, b, f)
    Wo, Ho = -(-W // f), -(-H // f)
    cw, ch = -(-Wo // hs), -(-Ho // vs)
    y = planar[:Wo * Ho].reshape(Ho, Wo)
    cb = planar[Wo * Ho:Wo * Ho + cw * ch].reshape(ch, cw)
    cr = planar[Wo * Ho + cw * ch:].reshape(ch, cw)
    ro, co = np.mgrid[0:Ho, 0:Wo]
    held = ((ro * f) % vf) != 0                              # only possible for f == 1
    last = (((W - 1) // hf) * hf) // f                       # last sample column, in output pixels
    src_r = np.where(held, ro - 1, ro) // vs
    src_c = np.where(held, last, co - co % hs) // hs
    out = np.stack([y, cb[src_r, src_c], cr[src_r, src_c]], -1)
    return inverse(out) if to_rgb else out


def slot_bits(q):
    t = sum(q)
    return 8 if t <= 8 else (16 if t <= 16 else 32)


def process_frame(rgb, a=4, b=4, q=(8, 8, 8), factor=1, ops=(3, 1, 2), round_mode=0, pool_mode=0, out_format=0,
                  in_format=0):
    """One frame [H,W,3|4] uint8 -> flat uint8 output in the csic_out_format layout."""
    H, W = rgb.shape[:2]
    rgb = rgb[..., [2, 1, 0]] if in_format == 2 else rgb[..., :3]      # alpha is never read
    ops = tuple(ops)
    chroma_first = ops.index(3) < ops.index(1)
    quant_first = ops.index(2) < ops.index(1)
    ycc = forward(rgb, round_mode)
    f = factor
    if pool_mode == 1 and f > 1:
        x = ycc
        if quant_first:
            x = quant(x, q)
        if chroma_first:
            fr, fc = chroma_source_full(W, H, a, b)
            x = np.stack([x[..., 0], x[fr, fc, 1], x[fr, fc, 2]], -1)
        x = _avg(x, f)
        if not quant_first:
            x = quant(x, q)
        if not chroma_first:
            Wo, Ho = W // f, H // f
            _, _, cr_, cc_ = source_maps(W, H, a, b, f, False)
            x = np.stack([x[..., 0], x[cr_ // f, cc_ // f, 1], x[cr_ // f, cc_ // f, 2]], -1)
        o = x
    else:
        yr, yc, cr_, cc_ = source_maps(W, H, a, b, f, chroma_first)
        o = quant(np.stack([ycc[yr, yc, 0], ycc[cr_, cc_, 1], ycc[cr_, cc_, 2]], -1), q)
    Ho, Wo = o.shape[:2]
    if out_format == 4:                                   # PLANAR: Y plane + surviving chroma sample points
        hs, vs = planar_factors(a, b, f)
        return np.concatenate([o[..., 0].reshape(-1), o[::vs, ::hs, 1].reshape(-1), o[::vs, ::hs, 2].reshape(-1)])
    if out_format == 0:
        return o.reshape(-1)
    if out_format == 1:
        return inverse(o).reshape(-1)
    sb = slot_bits(q) // 8
    word = 8 if out_format == 2 else 16
    row_bytes = -(-(Wo * sb) // word) * word
    v = ((o[..., 0].astype(np.uint32) >> (8 - q[0])) << (q[1] + q[2])) | \
        ((o[..., 1].astype(np.uint32) >> (8 - q[1])) << q[2]) | (o[..., 2].astype(np.uint32) >> (8 - q[2]))
    out = np.zeros((Ho, row_bytes), dtype=np.uint8)
    for k in range(sb):
        out[:, k:Wo * sb:sb] = ((v >> (8 * k)) & 0xFF).astype(np.uint8)
    return out.reshape(-1)


def process(p, rgb):
    """p: oracle.Params (or anything with the csic_params fields); rgb [n,H,W,3]."""
    rgb = np.asarray(rgb, dtype=np.uint8)
    if rgb.ndim == 3:
        rgb = rgb[None]
    return np.stack([process_frame(fr, p.chroma_a, p.chroma_b, (p.y_bits, p.cb_bits, p.cr_bits), p.factor,
                                   tuple(p.op), p.round_mode, p.pool_mode, p.out_format, getattr(p, 'in_format', 0)) for fr in rgb])
