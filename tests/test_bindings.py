"""The JVM-facing bindings, as far as an image without a JDK can take them.

bindings/jni/csic_jni.c is compiled against a stand-in <jni.h> (tests/c/jni_stub) and driven through a fake JNIEnv
(tests/c/jni_harness.c): error paths on the CPU, the data path on the GPU against the oracle.  The Scala sources
(bindings/scala/*.scala, bindings/jni/CsicJni.scala) cannot be compiled here; the checks below keep them in step with
the reference's surfaces and with include/csic.h, and bindings/ci/scala-bindings.yml is the job that compiles them."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

PKG = os.path.join(ROOT, "chroma-subsampling-image-compressor_b200")
B = os.path.join(ROOT, "bindings")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    d = tmp_path_factory.mktemp("jni")
    exe = str(d / "jni_harness")
    flags = ["-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", f"-I{ROOT}/tests/c/jni_stub", f"-I{ROOT}/include"]
    subprocess.check_call(["gcc", *flags, os.path.join(B, "jni", "csic_jni.c"), os.path.join(ROOT, "tests", "c", "jni_harness.c"),
                           f"-L{PKG}", "-lcsic", f"-Wl,-rpath,{PKG}", "-o", exe])
    return exe


def test_jni_shim_error_paths(harness):
    """Argument and `require` failures become the exceptions the Scala side expects (no GPU involved)."""
    r = subprocess.run([harness, "errors"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok  ") == 9 and "FAIL" not in r.stdout, r.stdout


def test_jni_shim_has_no_critical_sections():
    """csic_process_host allocates, spawns threads and blocks on CUDA: forbidden inside Get/ReleasePrimitiveArrayCritical."""
    src = open(os.path.join(B, "jni", "csic_jni.c")).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "PrimitiveArrayCritical" not in code
    native = set(re.findall(r"Java_jpeg_CsicJni_(\w+)\(", code))
    declared = set(re.findall(r"@native def (\w+)\(", open(os.path.join(B, "jni", "CsicJni.scala")).read()))
    assert native == declared, (native, declared)


def test_scala_twins_mirror_the_reference_surfaces():
    """ImageProcessorModelGpu has the method set of ImageProcessorModel.scala:14-52; the app twin has the flags and
    defaults of ImageCompressorTopApp.scala:164-173; every C symbol the Panama binding looks up is exported."""
    model = open(os.path.join(B, "scala", "ImageProcessorModelGpu.scala")).read()
    for sig in ("def readImage(file: String): ImmutableImage", "def writeImage(image: MutableImage, file: String): Unit",
                "def writeImage(image: Array[Pixel], p: ImageProcessorParams, file: String): Unit",
                "def getImageParams(image: ImmutableImage, numPixelsPerCycle: Int): ImageProcessorParams",
                "def getImagePixels(image: ImmutableImage): ImageType"):
        assert sig in model, sig
    app = open(os.path.join(B, "scala", "ImageCompressionAppGpu.scala")).read()
    for flag, default in (("--input", "test_images/in128x128.png"), ("--a", "4"), ("--b", "4"), ("--yq", "8"), ("--cbq", "8"),
                          ("--crq", "8"), ("--sf", "8"), ("--op1", "spatial"), ("--op2", "color"), ("--op3", "chroma")):
        assert f'argsMap.getOrElse("{flag}", "{default}")' in app, flag
    assert "def processImage(inputImagePath: String, outputImagePath: String, chromaParamA: Int, chromaParamB: Int," in app
    assert '"APP_OUTPUT"' in app and "_processed_" in app
    gpu = open(os.path.join(B, "scala", "CsicGpu.scala")).read()
    header = open(os.path.join(ROOT, "include", "csic.h")).read()
    for sym in re.findall(r'fn\("(csic_\w+)"', gpu):
        assert re.search(rf"\b{sym}\(", header), sym
    assert gpu.count("Arena.ofConfined()") == 2 and "finally call.close()" in gpu      # per-call arena
    ci = open(os.path.join(B, "ci", "scala-bindings.yml")).read()
    assert "sbt Test/compile" in ci and "csic_jni.c" in ci


@pytest.mark.gpu
def test_jni_shim_data_path_matches_oracle(harness, tmp_path):
    import oracle
    W, H = 96, 40
    out = tmp_path / "o.bin"
    r = subprocess.run([harness, "run", str(W), str(H), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    blob = np.fromfile(out, dtype=np.uint8)
    rgb, got = blob[:W * H * 3].reshape(1, H, W, 3), blob[W * H * 3:]
    want = oracle.process(oracle.make_params(W, H, 2, 0, (6, 5, 5), 2, "CSQ", out_format=1), rgb)
    assert np.array_equal(got, want[0])
