"""Host-side mirror of the reference's interface for the pixel path.

Same names, argument meaning and error behaviour as the Scala sources (paths relative to the
reference root) so that callers -- and the parity tests -- read like the reference's own:

  ImageCompressorTop(width, height, a, b, yBits, cbBits, crBits, downFactor, op1, op2, op3)
                                              src/main/scala/jpeg/ImageCompressorTop.scala:11-25
  ImageProcessorParams / ImageProcessor       src/main/scala/jpeg/ImageProcessor.scala:15-63
  ImageProcessorModel.{readImage, writeImage, getImageParams, getImagePixels}
                                              src/test/scala/jpeg/ImageProcessorModel.scala:14-52
  YCbCrUtils.ycbcr2rgb is fused into the kernel (out_format RGB888) instead of running on the host.

Where the reference pushes one pixel per simulated clock through a Chisel DUT, these classes make
one C-ABI call (one fused kernel launch) per batch of frames.
"""
import dataclasses
import os

import numpy as np

from . import api
from .api import (IllegalArgumentException, InFormat, OutFormat, PoolMode, ProcessingStep, RoundMode)

_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = api.Context(device)
    return _default_ctx[device]


class ImageCompressorTop:
    """Reorderable top: toYC -> op1 -> op2 -> op3 (ImageCompressorTop.scala:80-114).

    Constructing it validates exactly what the Scala constructor `require`s (:27-31 plus the
    sub-module requires); `process` replaces driving the DUT pixel by pixel
    (ImageCompressorTopApp.scala:53-131)."""

    def __init__(self, width, height, chroma_param_a_config, chroma_param_b_config, yTargetQuantBitsConfig,
                 cbTargetQuantBitsConfig, crTargetQuantBitsConfig, downFactorConfig, op1Type, op2Type, op3Type,
                 round_mode=RoundMode.FLOOR, pool_mode=PoolMode.DECIMATE, out_format=OutFormat.YCC888,
                 in_format=InFormat.RGB24, ctx=None):
        self.params = api.make_params(width, height, chroma_param_a_config, chroma_param_b_config,
                                      yTargetQuantBitsConfig, cbTargetQuantBitsConfig, crTargetQuantBitsConfig,
                                      downFactorConfig, (op1Type, op2Type, op3Type), round_mode, pool_mode, out_format,
                                      in_format)
        self._ctx = ctx

    @classmethod
    def legacy(cls, width, height, chroma_mode, quant_mode, factor=1, **kw):
        """`ImageCompressorTop(w, h, ChromaSubsamplingMode, QuantizationMode, factor)` of the enum era."""
        p = api.params_from_legacy(width, height, chroma_mode, quant_mode, factor)
        return cls(width, height, p.chroma_a, p.chroma_b, p.y_bits, p.cb_bits, p.cr_bits, factor,
                   p.op[0], p.op[1], p.op[2], **kw)

    @property
    def ctx(self):
        return self._ctx or default_context()

    @property
    def out_shape(self):
        """(out_h, out_w) of the emitted stream: ceil(H/f) x ceil(W/f)."""
        w, h, _, _ = api.out_shape(self.params)
        return h, w

    def process(self, rgb):
        """rgb: uint8 [H,W,3] or [n,H,W,3] (4 channels for in_format RGBA32/BGRA32) -> uint8 [n, out_h, out_w, 3]
        (YCC888/RGB888) or [n, bytes] (bundles)."""
        out = self.ctx.process_host(self.params, rgb)
        if self.params.out_format in (OutFormat.YCC888, OutFormat.RGB888):
            h, w = self.out_shape
            out = out.reshape(-1, h, w, 3)
        return out


@dataclasses.dataclass(frozen=True)
class ImageProcessorParams:
    """case class ImageProcessorParams -- ImageProcessor.scala:15-29 (same requires, same messages)."""
    width: int
    height: int
    factor: int
    chromaParamA: int
    chromaParamB: int

    def __post_init__(self):
        import ctypes
        from . import _ffi
        p = _ffi.CsicParams()
        if self.width <= 0:
            raise IllegalArgumentException(-1, "width must be positive")
        if self.height <= 0:
            raise IllegalArgumentException(-1, "height must be positive")
        rc = _ffi.lib().csic_params_from_image_processor(self.width, self.height, self.factor, self.chromaParamA,
                                                         self.chromaParamB, ctypes.byref(p))
        if rc == -2:
            raise IllegalArgumentException(rc, "factor must be 1, 2, 4, or 8")
        if rc == -4:
            raise IllegalArgumentException(rc, f"chromaParamA must be 4, 2, or 1. Got {self.chromaParamA}")
        if rc == -5:
            raise IllegalArgumentException(
                rc, f"chromaParamB must be equal to chromaParamA ({self.chromaParamA}) or 0. Got {self.chromaParamB}")
        api.check(rc)
        object.__setattr__(self, "_csic", p)


class ImageProcessor:
    """class ImageProcessor(p) -- ImageProcessor.scala:31-63: toYC -> chroma -> spatial, no quantiser."""

    def __init__(self, p: ImageProcessorParams, out_format=OutFormat.YCC888, ctx=None):
        from . import _ffi
        self.p = p
        self.params = _ffi.CsicParams.from_buffer_copy(p._csic)    # own copy: p is frozen and may be shared
        self.params.out_format = int(out_format)
        api.validate(self.params)
        self._ctx = ctx

    def process(self, rgb):
        ctx = self._ctx or default_context()
        out = ctx.process_host(self.params, rgb)
        w, h, _, _ = api.out_shape(self.params)
        return out.reshape(-1, h, w, 3) if self.params.out_format <= 1 else out


class ImageProcessorModel:
    """object ImageProcessorModel -- src/test/scala/jpeg/ImageProcessorModel.scala:9-53.
    scrimage is replaced by Pillow; an image is a uint8 [H,W,3] array (alpha dropped, as
    pixel.red()/green()/blue() does)."""

    @staticmethod
    def readImage(file):                                    # :14-16
        from PIL import Image
        im = Image.open(file)
        if im.mode not in ("RGB", "RGBA"):
            im = im.convert("RGBA")
        return np.ascontiguousarray(np.asarray(im)[..., :3], dtype=np.uint8)

    @staticmethod
    def writeImage(image, file, p=None):                    # :18-22 and :24-28 (Array[Pixel] + params)
        from PIL import Image
        arr = np.asarray(image, dtype=np.uint8)
        if p is not None:
            arr = arr.reshape(p.height, p.width, 3)
        parent = os.path.dirname(os.path.abspath(file))
        os.makedirs(parent, exist_ok=True)                  # getParentFile().mkdirs()
        Image.fromarray(arr, "RGB").save(file, format="PNG")

    @staticmethod
    def getImageParams(image, numPixelsPerCycle):           # :33-41 -- defaults to 4:4:4
        h, w = image.shape[:2]
        return ImageProcessorParams(width=w, height=h, factor=numPixelsPerCycle, chromaParamA=4, chromaParamB=4)

    @staticmethod
    def getImagePixels(image):                              # :43-52 -- H x W x [r,g,b]
        return np.asarray(image, dtype=np.uint8)[..., :3].astype(int).tolist()
