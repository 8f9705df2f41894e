#!/usr/bin/env python3
"""Copy the reference's committed input/output PNGs into tests/golden/ and write MANIFEST.json.

The reference (/root/reference, Scala/Chisel) cannot run in this container (no JVM), so its
committed PNGs are the only outputs of the real reference we have.  They are *test vectors*,
not source.  This script is the provenance record: it must be run in the authoring container,
where /root/reference exists; the GPU box only ever sees the copies.

Each golden is tagged with the recipe that the oracle must
reproduce bit-exactly; tests/test_oracle_golden.py walks the manifest.
"""
import hashlib, json, os, shutil, sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))

def R(src, dst, inp, fwd, a, b, q, f, order, note):
    return dict(src=src, file=dst, input=inp, forward=fwd, a=a, b=b, q=list(q), factor=f,
                order=order, note=note)

CH = "APP_OUTPUT/chroma_subsampler_parameterized_tests"
QZ = "APP_OUTPUT/quantizer_parameterized_tests"
ENTRIES = []
# G1-G4: ChromaSubsamplerImageSpec.scala:106-111 (stage bench, trunc forward model :28-42)
for tag, a, b in (("4-4_444", 4, 4), ("2-2_422", 2, 2), ("2-0_420", 2, 0), ("1-1_411", 1, 1)):
    n = f"output_chroma_4-{tag}_16x16.png"
    ENTRIES.append(R(f"{CH}/{n}", f"G_chroma_param_{n}", "in16x16.png", "trunc", a, b, (8, 8, 8), 1, "CSQ",
                     "ChromaSubsamplerImageSpec"))
# G5-G10: ColorQuantizerImageSpec.scala:84-91 (stage bench, trunc forward model :28-48)
for y, cb, cr in ((8, 8, 8), (6, 5, 5), (3, 3, 2), (8, 4, 4), (4, 4, 4), (1, 1, 1)):
    n = f"output_quantized_Y{y}Cb{cb}Cr{cr}_128x128.png"
    ENTRIES.append(R(f"{QZ}/{n}", f"G_quant_param_{n}", "in128x128.png", "trunc", 4, 4, (y, cb, cr), 1, "CSQ",
                     "ColorQuantizerImageSpec"))
# G11-G13: legacy QuantizationMode enum outputs (pins Q_24BIT/Q_16BIT/Q_8BIT)
for tag, q in (("Q24bit", (8, 8, 8)), ("Q16bit", (6, 5, 5)), ("Q8bit", (3, 3, 2))):
    n = f"output_quantized_{tag}_128x128.png"
    ENTRIES.append(R(f"output_images_quantizer/{n}", f"G_quant_legacy_{n}", "in128x128.png", "trunc", 4, 4, q, 1,
                     "CSQ", "legacy QuantizationMode enum"))
# G14-G22: legacy ChromaSubsamplingMode enum outputs
for tag, (a, b) in (("444", (4, 4)), ("422", (2, 2)), ("420", (2, 0))):
    for s in (16, 128, 512):
        n = f"output_chroma_{tag}_{s}x{s}.png"
        ENTRIES.append(R(f"output_images_chroma/{n}", f"G_chroma_legacy_{n}", f"in{s}x{s}.png", "trunc", a, b,
                         (8, 8, 8), 1, "CSQ", "legacy ChromaSubsamplingMode enum"))
# G23: ImageProcessor integration (RTL simulation output), SpatialDownsamplerSpec.scala:155-230
ENTRIES.append(R("APP_OUTPUT/spatial_downsampler_integration_420_sf2.png", "G_imageprocessor_420_sf2_16x16.png",
                 "in16x16.png", "floor", 2, 0, (8, 8, 8), 2, "CSQ", "ImageProcessor RTL sim"))
ENTRIES.append(R("output_images/out16x16_processed.png", "G_out16x16_processed.png",
                 "in16x16.png", "floor", 2, 0, (8, 8, 8), 2, "CSQ", "ImageProcessor RTL sim (copy)"))
# G24: spatial only
for n in ("out16x16.png", "out8x8.png"):
    ENTRIES.append(R(f"output_images/{n}", f"G_{n}", "in16x16.png", "floor", 4, 4, (8, 8, 8), 2, "CSQ",
                     "spatial f=2 only, RTL sim"))
# G25: identity copy of the input (no pipeline): recipe 'identity'
ENTRIES.append(R("output_images/out16x16_model_copy.png", "G_out16x16_model_copy.png", "in16x16.png", "identity",
                 4, 4, (8, 8, 8), 1, "CSQ", "identity copy"))
# G26: ImageCompressorTop via ImageCompressionApp (RTL sim), chroma before spatial
ENTRIES.append(R("APP_OUTPUT/in128x128_processed_chroma4-2-2_Y8Cb8Cr8_sf2_order-Pr-Pr-Pr.png",
                 "G_top_422_Y8Cb8Cr8_sf2_128x128.png", "in128x128.png", "floor", 2, 2, (8, 8, 8), 2, "CSQ",
                 "ImageCompressorTop RTL sim; order string lost (Pr-Pr-Pr); chroma-before-spatial matches"))
# G27: BASELINE config 1 (legacy enum top): 4:2:0 + Q_8BIT + sf1
ENTRIES.append(R("APP_OUTPUT/in128x128_processed_chromaChromaSubsamplingMode(2=CHROMA_420)_quantQuantizationMode(2=Q_8BIT)_sf1.png",
                 "G_top_legacy_CHROMA_420_Q_8BIT_sf1_128x128.png", "in128x128.png", "floor", 2, 0, (3, 3, 2), 1,
                 "CSQ", "BASELINE.json configs[0]; legacy enum top, RTL sim"))

INPUTS = ["in16x16.png", "in128x128.png", "in512x512.png"]

def sha1(p):
    return hashlib.sha1(open(p, "rb").read()).hexdigest()

def main():
    if not os.path.isdir(REF):
        sys.exit("run this where /root/reference exists (authoring container)")
    man = dict(reference="Andurdur/Chroma-Subsampling-Image-Compressor", inputs={}, goldens=[])
    for n in INPUTS:
        shutil.copyfile(f"{REF}/test_images/{n}", f"{HERE}/{n}")
        os.chmod(f"{HERE}/{n}", 0o644)
        man["inputs"][n] = dict(src=f"test_images/{n}", sha1=sha1(f"{HERE}/{n}"))
    for e in ENTRIES:
        shutil.copyfile(f"{REF}/{e['src']}", f"{HERE}/{e['file']}")
        os.chmod(f"{HERE}/{e['file']}", 0o644)
        e = dict(e, sha1=sha1(f"{HERE}/{e['file']}"))
        man["goldens"].append(e)
    json.dump(man, open(f"{HERE}/MANIFEST.json", "w"), indent=1)
    print(f"{len(INPUTS)} inputs, {len(ENTRIES)} goldens -> {HERE}")

if __name__ == "__main__":
    main()
