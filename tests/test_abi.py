"""The C-ABI library loads, exports every symbol include/csic.h declares, and its host-only half
(validation, geometry, names) behaves like the reference's constructors.  No compute, no GPU."""
import ctypes
import os
import re

import pytest

import csic_b200 as csic
from csic_b200 import _ffi
from conftest import ALL_AB, ROOT

HEADER = open(os.path.join(ROOT, "include", "csic.h")).read()


def header_symbols():
    return sorted(set(re.findall(r"^CSIC_API\s+[\w\s\*]+?\b(csic_\w+)\s*\(", HEADER, flags=re.M)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 20
    L = ctypes.CDLL(_ffi.LIB_PATH)
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/csic.h but not exported by libcsic.so"
    assert sorted(_ffi.PROTOTYPES) == syms, "python binding table out of sync with the header"
    assert _ffi.lib().csic_abi_version() == 1


def test_params_struct_layout_matches_header():
    assert ctypes.sizeof(_ffi.CsicParams) == 16 * 4
    m = re.search(r"typedef struct csic_params \{(.*?)\} csic_params;", HEADER, flags=re.S).group(1)
    fields = re.findall(r"int32_t ([^;]+);", m)
    names = [n.strip().split("[")[0] for f in fields for n in f.split(",")]
    assert names == [f[0] for f in _ffi.CsicParams._fields_]


def test_header_cites_reference_for_every_entry_point():
    # every prototype is preceded by a comment naming the reference file it replaces or mirrors
    assert HEADER.count(".scala") >= 25


# ---- require(...) predicates and their messages -----------------------------------------------------
def msg_of(**kw):
    base = dict(width=16, height=16, a=4, b=4, y_bits=8, cb_bits=8, cr_bits=8, factor=1)
    base.update(kw)
    with pytest.raises(csic.IllegalArgumentException) as e:
        csic.make_params(**base)
    return str(e.value), e.value.status


def test_validation_messages_match_reference():
    assert msg_of(factor=3) == ("requirement failed: Factor must be 1, 2, 4, or 8", -2)        # SpatialDownsampler.scala:8
    assert msg_of(width=0)[1] == -1 and msg_of(height=-4)[1] == -1                               # :7
    assert msg_of(a=3) == ("requirement failed: param_a must be 4, 2, or 1. Got 3", -4)         # ChromaSubsampler.scala:17
    assert msg_of(a=2, b=1) == ("requirement failed: param_b must be equal to param_a (2) or 0. Got 1", -5)   # :18
    assert msg_of(y_bits=0) == ("requirement failed: Y target bits must be between 1 and 8. Got 0", -6)       # ColorQuantizer.scala:13
    assert msg_of(cb_bits=9)[0].startswith("requirement failed: Cb target bits must be between 1 and 8")
    assert msg_of(cr_bits=-1)[0].startswith("requirement failed: Cr target bits")
    S = csic.ProcessingStep
    assert msg_of(ops=(S.SpatialSampling, S.SpatialSampling, S.ChromaSubsampling))[1] == -7     # ImageCompressorTop.scala:31
    assert "op2Type must be a valid reorderable operation" in msg_of(ops=(S.SpatialSampling, S.NoOp, S.ChromaSubsampling))[0]
    assert msg_of(factor=2, width=10, height=7, pool_mode=csic.PoolMode.AVERAGE)[1] == -3


def test_doubly_invalid_reports_the_constructors_first_failure():
    """`new ImageCompressorTop(...)` evaluates its requires in the order ops (:28-31) -> SpatialDownsampler (:45;
    dims, factor) -> ColorQuantizer (:46-51; Y, Cb, Cr) -> ChromaSubsampler (:53-59; a, b): a parameter set that breaks
    several of them must raise the message of the FIRST one in that order (missing #6)."""
    S = csic.ProcessingStep
    bad_ops = (S.SpatialSampling, S.SpatialSampling, S.ChromaSubsampling)
    assert msg_of(ops=bad_ops, width=0)[1] == -7                         # ops before dims
    assert msg_of(ops=bad_ops, factor=3, a=3, y_bits=0)[1] == -7
    assert "op1Type" in msg_of(ops=(S.NoOp, S.NoOp, S.NoOp), factor=5)[0]
    assert msg_of(width=0, factor=3)[1] == -1                            # dims before factor
    assert msg_of(factor=3, y_bits=0)[1] == -2                           # factor (spatial) before quantiser
    assert msg_of(factor=3, a=3)[1] == -2                                # factor before chroma
    assert msg_of(y_bits=0, a=3)[1] == -6                                # quantiser before chroma (was the other way round)
    assert msg_of(cb_bits=9, a=2, b=1)[1] == -6
    assert msg_of(y_bits=0, cb_bits=0)[0].endswith("Y target bits must be between 1 and 8. Got 0")
    assert msg_of(cb_bits=0, cr_bits=0)[0].startswith("requirement failed: Cb target bits")
    assert msg_of(a=3, b=1)[1] == -4                                     # param_a before param_b
    assert msg_of(a=3, out_format=9)[1] == -4                            # extension fields last
    assert msg_of(out_format=9)[1] == -8


def test_image_processor_params_doubly_invalid_order():
    """case class ImageProcessorParams (ImageProcessor.scala:22-28): width, height, factor, divisibility, chromaParamA,
    chromaParamB -- divisibility is checked BEFORE the chroma parameters."""
    P = csic.ImageProcessorParams
    for kw, text in (((0, 0, 3, 3, 1), "width must be positive"), ((16, 0, 3, 3, 1), "height must be positive"),
                     ((10, 16, 3, 3, 1), "factor must be 1, 2, 4, or 8"),
                     ((10, 16, 4, 3, 1), "Image dimensions must be divisible by spatial downsampling factor."),
                     ((16, 16, 4, 3, 1), "chromaParamA must be 4, 2, or 1. Got 3")):
        with pytest.raises(csic.IllegalArgumentException) as e:
            P(*kw)
        assert text in str(e.value), (kw, str(e.value))


def test_image_processors_do_not_share_a_params_struct():
    """Two ImageProcessors built from one frozen ImageProcessorParams keep their own out_format."""
    pp = csic.ImageProcessorParams(16, 16, 2, 2, 0)
    a = csic.ImageProcessor(pp, out_format=csic.OutFormat.YCC888)
    b = csic.ImageProcessor(pp, out_format=csic.OutFormat.RGB888)
    assert a.params.out_format == csic.OutFormat.YCC888 and b.params.out_format == csic.OutFormat.RGB888
    assert pp._csic.out_format == csic.OutFormat.YCC888
    with pytest.raises(csic.IllegalArgumentException):
        csic.ImageProcessor(pp, out_format=9)


def test_all_legal_surface_accepted():
    for a, b in ALL_AB:
        for f in (1, 2, 4, 8):
            for bits in (1, 5, 8):
                csic.make_params(5, 3, a, b, bits, bits, bits, f)       # the top has no divisibility require
    for reject in ((4, 2), (2, 4), (1, 2), (0, 0), (8, 8)):
        with pytest.raises(csic.IllegalArgumentException):
            csic.make_params(8, 8, *reject)


def test_image_processor_params_requires():
    """ImageProcessor.scala:22-28, incl. the divisibility predicate only ImageProcessorParams has."""
    P = csic.ImageProcessorParams
    P(16, 16, 2, 2, 0)
    for kw, text in (((0, 16, 1, 4, 4), "width must be positive"), ((16, 0, 1, 4, 4), "height must be positive"),
                     ((16, 16, 3, 4, 4), "factor must be 1, 2, 4, or 8"),
                     ((10, 16, 4, 4, 4), "Image dimensions must be divisible by spatial downsampling factor."),
                     ((16, 16, 2, 3, 3), "chromaParamA must be 4, 2, or 1. Got 3"),
                     ((16, 16, 2, 2, 1), "chromaParamB must be equal to chromaParamA (2) or 0. Got 1")):
        with pytest.raises(csic.IllegalArgumentException) as e:
            P(*kw)
        assert text in str(e.value)
    import numpy as np
    p = csic.ImageProcessorModel.getImageParams(np.zeros((8, 12, 3), np.uint8), 4)   # ImageProcessorModel.scala:33-41
    assert (p.width, p.height, p.factor, p.chromaParamA, p.chromaParamB) == (12, 8, 4, 4, 4)


def test_legacy_enum_mapping():
    """CHROMA_444/422/420 and Q_24BIT/16BIT/8BIT (pinned by goldens G11-G22, G27)."""
    C, Q = csic.ChromaSubsamplingMode, csic.QuantizationMode
    want_c = {C.CHROMA_444: (4, 4), C.CHROMA_422: (2, 2), C.CHROMA_420: (2, 0)}
    want_q = {Q.Q_24BIT: (8, 8, 8), Q.Q_16BIT: (6, 5, 5), Q.Q_8BIT: (3, 3, 2)}
    for c, ab in want_c.items():
        for q, bits in want_q.items():
            p = csic.params_from_legacy(64, 32, c, q, 2)
            assert (p.chroma_a, p.chroma_b) == ab and (p.y_bits, p.cb_bits, p.cr_bits) == bits and p.factor == 2
    with pytest.raises(csic.IllegalArgumentException):
        csic.params_from_legacy(64, 32, 3, 0, 1)


def test_out_shape_and_bundle_geometry():
    F = csic.OutFormat
    assert csic.out_shape(csic.make_params(5, 3, factor=2)) == (3, 2, 9, 18)            # ceil dims, SpatialDownsamplerSpec.scala:120-123
    assert csic.out_shape(csic.make_params(3840, 2160, 2, 0, factor=2, out_format=F.BUNDLE128)) == (1920, 1080, 7680, 8294400)
    assert csic.out_shape(csic.make_params(10, 4, y_bits=3, cb_bits=3, cr_bits=2, out_format=F.BUNDLE64))[2] == 16   # 10 slots of 1 B -> 2 words
    assert csic.out_shape(csic.make_params(10, 4, y_bits=6, cb_bits=5, cr_bits=5, out_format=F.BUNDLE128))[2] == 32  # 20 B -> 2 words
    assert csic.out_shape(csic.make_params(10, 4, out_format=F.BUNDLE64))[2] == 40


def test_parse_processing_step():
    """ImageCompressorTopApp.scala:154-161."""
    S = csic.ProcessingStep
    for name, want in (("spatial", S.SpatialSampling), ("SpatialSampling", S.SpatialSampling), ("COLOR", S.ColorQuantization),
                       ("colorquantization", S.ColorQuantization), ("Chroma", S.ChromaSubsampling),
                       ("chromasubsampling", S.ChromaSubsampling)):
        assert csic.parse_processing_step(name) == want
    with pytest.raises(csic.IllegalArgumentException) as e:
        csic.parse_processing_step("blur")
    assert "Unknown processing step: blur. Use 'spatial', 'color', or 'chroma'." in str(e.value)


def test_band_input_rows():
    p = csic.make_params(64, 32, 2, 0, factor=1)                               # 4:2:0 full res, chroma first
    assert csic.band_input_rows(p, 0, 8) == (0, 8)
    assert csic.band_input_rows(p, 3, 2) == (2, 3)                             # odd start: needs the row above
    p = csic.make_params(64, 32, 2, 0, factor=2)
    assert csic.band_input_rows(p, 4, 4) == (8, 7)                             # rows 8,10,12,14
    S = csic.ProcessingStep
    p = csic.make_params(64, 32, 2, 0, factor=2, ops=(S.SpatialSampling, S.ColorQuantization, S.ChromaSubsampling))
    assert csic.band_input_rows(p, 2, 2) == (2, 5)                             # counter line 1 is odd: held from out row 1
    with pytest.raises(csic.IllegalArgumentException):
        csic.band_input_rows(p, 15, 4)


def test_no_device_means_error_not_fallback():
    """Without a usable GPU the process path must fail loudly (CSIC_ENODEVICE), never compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is covered on the CPU box")
    assert csic.device_count() == 0
    with pytest.raises(csic.CsicError) as e:
        csic.Context(0)
    assert e.value.status == -10 and "no CPU fallback" in str(e.value)


def test_library_exports_only_the_c_abi():
    """nm -D: every defined dynamic symbol is a csic_* function of include/csic.h (linker version script csrc/csic.map);
    no C++ template specialisation or libstdc++ weak symbol leaks (VERDICT r1 hygiene)."""
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", csic.LIB_PATH], capture_output=True, text=True, check=True).stdout
    names = [l.split()[-1] for l in out.splitlines() if l.strip()]
    assert names and all(n.startswith("csic_") for n in names), [n for n in names if not n.startswith("csic_")]
    assert set(names) == set(_ffi.PROTOTYPES), set(names) ^ set(_ffi.PROTOTYPES)
