"""`ImageCompressionApp` -- mirror of src/test/scala/jpeg/ImageCompressorTopApp.scala.

    python -m csic_b200.app --input test.png --a 2 --b 0 --yq 3 --cbq 3 --crq 2 --sf 1 \
                            --op1 chroma --op2 color --op3 spatial

Same flags, defaults (:164-173), console banner (:177-185) and output file naming (:187-190).
The Chisel DUT + per-pixel collector loop (:53-131) and the host ycbcr2rgb (:118) are one fused
kernel launch.
"""
import os
import sys

import numpy as np

from .api import OutFormat, ProcessingStep, parse_processing_step
from .model import ImageCompressorTop, ImageProcessorModel


def processImage(inputImagePath, outputImagePath, chromaParamA, chromaParamB, yTargetBits, cbTargetBits,
                 crTargetBits, spatialFactorToUse, op1, op2, op3, ctx=None):
    """ImageCompressionApp.processImage -- ImageCompressorTopApp.scala:23-145."""
    inputImage = ImageProcessorModel.readImage(inputImagePath)                  # :39
    imageHeight, imageWidth = inputImage.shape[:2]
    spatialInPipeline = ProcessingStep.SpatialSampling in (op1, op2, op3)        # :43
    finalW = imageWidth // spatialFactorToUse if spatialInPipeline else imageWidth      # :44
    finalH = imageHeight // spatialFactorToUse if spatialInPipeline else imageHeight    # :45
    if spatialInPipeline and (imageWidth % spatialFactorToUse or imageHeight % spatialFactorToUse):
        print(f"[WARN] Image dimensions ({imageWidth}x{imageHeight}) are not perfectly divisible by spatialFactor "
              f"({spatialFactorToUse}). SpatialDownsampler might truncate.")      # :47-49
    top = ImageCompressorTop(imageWidth, imageHeight, chromaParamA, chromaParamB, yTargetBits, cbTargetBits,
                             crTargetBits, spatialFactorToUse, op1, op2, op3, out_format=OutFormat.RGB888, ctx=ctx)
    stream = top.process(inputImage).reshape(-1, 3)       # what the collector loop gathers, in emission order
    # :108-142 -- the first finalW*finalH emitted pixels are laid out row-major in a finalW-wide image
    # (a magenta canvas shows through if the DUT emitted fewer).
    expected = finalW * finalH
    canvas = np.empty((expected, 3), np.uint8)
    canvas[:] = (255, 0, 255)
    n = min(expected, len(stream))
    canvas[:n] = stream[:n]
    ImageProcessorModel.writeImage(canvas.reshape(finalH, finalW, 3), outputImagePath)  # :144
    return canvas.reshape(finalH, finalW, 3)


def _step_tag(step):
    return ProcessingStep(step).name[:2]                                            # :188 `.take(2)`


def main(argv=None):
    args = list(sys.argv[1:] if argv is None else argv)
    argsMap = {}
    for i in range(0, len(args) - 1, 2):                                            # args.sliding(2, 2), :149-151
        if args[i].startswith("--"):
            argsMap[args[i]] = args[i + 1]
    inputPath = argsMap.get("--input", "test_images/in128x128.png")
    a = int(argsMap.get("--a", "4"))
    b = int(argsMap.get("--b", "4"))
    yq = int(argsMap.get("--yq", "8"))
    cbq = int(argsMap.get("--cbq", "8"))
    crq = int(argsMap.get("--crq", "8"))
    sf = int(argsMap.get("--sf", "8"))
    op1 = parse_processing_step(argsMap.get("--op1", "spatial"))
    op2 = parse_processing_step(argsMap.get("--op2", "color"))
    op3 = parse_processing_step(argsMap.get("--op3", "chroma"))
    imageName = os.path.basename(inputPath).split(".")[0]
    bar = "-" * 52
    print(bar); print("Image Compressor Application Parameters:"); print(bar)
    print(f"Input Image: {inputPath}")
    print(f"Selected Chroma Subsampling (J:a:b): 4:{a}:{b}")
    print(f"Selected Quantization Bits (Y/Cb/Cr): {yq}/{cbq}/{crq}")
    print(f"Selected Spatial Downsampling Factor: {sf}")
    print(f"Selected Pipeline Order: {op1.name} -> {op2.name} -> {op3.name}")
    print(bar)
    outDir = argsMap.get("--outdir", "APP_OUTPUT")
    order = f"order-{_step_tag(op1)}-{_step_tag(op2)}-{_step_tag(op3)}"
    suffix = f"chroma4-{a}-{b}_Y{yq}Cb{cbq}Cr{crq}_sf{sf}_{order}"
    outputPath = f"{outDir}/{imageName}_processed_{suffix}.png"
    if not os.path.exists(inputPath):
        print(f"[ERROR] Input image not found: {inputPath}")
        return 1
    processImage(inputPath, outputPath, a, b, yq, cbq, crq, sf, op1, op2, op3)
    print(f"Image processing complete. Output saved to: {outputPath}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
