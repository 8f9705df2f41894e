"""CPU oracle for the pixel pipeline -- TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py (cpu_baseline / --impl reference).
The product package never imports this module (tests/test_host_logic.py::test_product_never_touches_the_oracle).

`lib()` loads oracle/_build/libcsic_oracle.so (built from csic_oracle.c by `make -C oracle`), the
literal streaming restatement of the reference.  `csic_oracle_np` is an independent closed-form
NumPy twin used to cross-check it.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libcsic_oracle.so")
_lib = None


class Params(ctypes.Structure):
    """Mirror of `csic_params` (include/csic.h): 16 x int32."""
    _fields_ = [("width", ctypes.c_int32), ("height", ctypes.c_int32),
                ("chroma_a", ctypes.c_int32), ("chroma_b", ctypes.c_int32),
                ("y_bits", ctypes.c_int32), ("cb_bits", ctypes.c_int32), ("cr_bits", ctypes.c_int32),
                ("factor", ctypes.c_int32), ("op", ctypes.c_int32 * 3),
                ("round_mode", ctypes.c_int32), ("pool_mode", ctypes.c_int32), ("out_format", ctypes.c_int32),
                ("in_format", ctypes.c_int32), ("reserved", ctypes.c_int32)]


def build(force=False):
    src = os.path.join(_HERE, "csic_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        P = ctypes.POINTER(Params)
        u8p = ctypes.c_void_p
        L.csic_oracle_process.argtypes = [P, u8p, ctypes.c_size_t, u8p, ctypes.c_int]
        L.csic_oracle_process.restype = ctypes.c_int
        L.csic_oracle_out_bytes_per_frame.argtypes = [P]
        L.csic_oracle_out_bytes_per_frame.restype = ctypes.c_size_t
        ip = ctypes.POINTER(ctypes.c_int)
        L.csic_oracle_rgb2ycbcr.argtypes = [ctypes.c_int] * 4 + [ip] * 3
        L.csic_oracle_ycbcr2rgb.argtypes = [ctypes.c_int] * 3 + [ip] * 3
        L.csic_oracle_ycbcr2rgb_array.argtypes = [u8p, ctypes.c_size_t, u8p]
        L.csic_oracle_ycbcr2rgb_array.restype = None
        L.csic_oracle_quant.argtypes = [ctypes.c_int, ctypes.c_int]
        L.csic_oracle_quant.restype = ctypes.c_int
        for fn in (L.csic_oracle_chroma_stream, L.csic_oracle_spatial_stream, L.csic_oracle_quant_stream):
            fn.restype = ctypes.c_size_t
        L.csic_oracle_chroma_stream.argtypes = [u8p, ctypes.c_size_t] + [ctypes.c_int] * 4 + [u8p]
        L.csic_oracle_spatial_stream.argtypes = [u8p, ctypes.c_size_t] + [ctypes.c_int] * 3 + [u8p]
        L.csic_oracle_quant_stream.argtypes = [u8p, ctypes.c_size_t] + [ctypes.c_int] * 3 + [u8p]
        _lib = L
    return _lib


STEP = {"spatial": 1, "color": 2, "chroma": 3}
ORDERS = {"SQC": (1, 2, 3), "SCQ": (1, 3, 2), "QSC": (2, 1, 3), "QCS": (2, 3, 1), "CSQ": (3, 1, 2), "CQS": (3, 2, 1)}


def make_params(width, height, a=4, b=4, q=(8, 8, 8), factor=1, order="CSQ", round_mode=0, pool_mode=0,
                out_format=0, in_format=0):
    p = Params()
    p.width, p.height, p.chroma_a, p.chroma_b = width, height, a, b
    p.y_bits, p.cb_bits, p.cr_bits = q
    p.factor = factor
    ops = ORDERS[order] if isinstance(order, str) else tuple(order)
    p.op[0], p.op[1], p.op[2] = ops
    p.round_mode, p.pool_mode, p.out_format = round_mode, pool_mode, out_format
    p.in_format = in_format
    return p


def process(p, rgb, threads=1):
    """rgb: uint8 array [n, H, W, 3] (or [H, W, 3]; 4 channels for in_format 1/2).  Returns uint8 [n, bytes_per_frame]."""
    L = lib()
    rgb = np.ascontiguousarray(rgb, dtype=np.uint8)
    if rgb.ndim == 3:
        rgb = rgb[None]
    n = rgb.shape[0]
    assert rgb.shape[1:] == (p.height, p.width, 3 if p.in_format == 0 else 4), (rgb.shape, p.height, p.width)
    out = np.empty((n, L.csic_oracle_out_bytes_per_frame(ctypes.byref(p))), dtype=np.uint8)
    rc = L.csic_oracle_process(ctypes.byref(p), rgb.ctypes.data, n, out.ctypes.data, threads)
    if rc != 0:
        raise ValueError(f"csic_oracle_process failed: {rc}")
    return out


def rgb2ycbcr(r, g, b, round_mode=0):
    y, cb, cr = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib().csic_oracle_rgb2ycbcr(r, g, b, round_mode, ctypes.byref(y), ctypes.byref(cb), ctypes.byref(cr))
    return y.value, cb.value, cr.value


def ycbcr2rgb(y, cb, cr):
    r, g, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib().csic_oracle_ycbcr2rgb(y, cb, cr, ctypes.byref(r), ctypes.byref(g), ctypes.byref(b))
    return r.value, g.value, b.value


def ycbcr2rgb_array(ycc):
    """uint8 [..., 3] (Y,Cb,Cr) -> uint8 [..., 3] (R,G,B): csic_oracle_ycbcr2rgb applied to every triple."""
    ycc = np.ascontiguousarray(ycc, dtype=np.uint8)
    assert ycc.shape[-1] == 3
    out = np.empty_like(ycc)
    lib().csic_oracle_ycbcr2rgb_array(ycc.ctypes.data, ycc.size // 3, out.ctypes.data)
    return out


def quant(v, bits):
    return lib().csic_oracle_quant(v, bits)


def _stream(fn, ycc, *args, out_len=None):
    ycc = np.ascontiguousarray(ycc, dtype=np.uint8).reshape(-1, 3)
    out = np.empty((out_len if out_len is not None else len(ycc), 3), dtype=np.uint8)
    m = fn(ycc.ctypes.data, len(ycc), *args, out.ctypes.data)
    return out[:m]


def chroma_stream(ycc, W, H, a, b):
    return _stream(lib().csic_oracle_chroma_stream, ycc, W, H, a, b)


def spatial_stream(ycc, W, H, f):
    return _stream(lib().csic_oracle_spatial_stream, ycc, W, H, f)


def quant_stream(ycc, yb, cbb, crb):
    return _stream(lib().csic_oracle_quant_stream, ycc, yb, cbb, crb)
