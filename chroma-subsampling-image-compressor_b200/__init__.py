"""csic_b200 -- B200-native pixel pipeline behind the reference's parameter surface.

Product code.  It never imports `oracle/` (CPU restatement = test infrastructure) and has no CPU
compute path: importing works anywhere (parameter validation is host-only), processing needs a B200.
"""
from . import _ffi
from .api import (ChromaSubsamplingMode, Context, CsicError, IllegalArgumentException, InFormat, MultiContext, OutFormat, PinnedBuffer,
                  PoolMode, ProcessingStep, QuantizationMode, RoundMode, band_input_rows, device_count, make_params,
                  out_shape, planar_shape, params_from_legacy, parse_processing_step, validate)
from .model import (ImageCompressorTop, ImageProcessor, ImageProcessorModel, ImageProcessorParams, default_context)
from .sharding import band_plan, frame_shard

LIB_PATH = _ffi.LIB_PATH
__all__ = [n for n in dir() if not n.startswith("_")]
