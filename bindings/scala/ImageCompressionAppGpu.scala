// Drop-in twin of `object ImageCompressionApp` (src/test/scala/jpeg/ImageCompressorTopApp.scala) whose
// processImage calls libcsic.so instead of simulating the Chisel DUT: same signature as :23-37, same PNG in / PNG
// out via ImageProcessorModel, and a `main` with the flags, defaults, banner and output naming of :149-190.
// Compiled by bindings/ci/scala-bindings.yml; never compiled in the authoring image (no JVM there).
//   sbt "Test/runMain jpeg.ImageCompressionAppGpu --input test_images/in128x128.png --a 2 --b 0 --sf 2"
package jpeg

import com.sksamuel.scrimage.ImmutableImage
import java.awt.{Color => AwtColor}

object ImageCompressionAppGpu extends App {
  def processImage(inputImagePath: String, outputImagePath: String, chromaParamA: Int, chromaParamB: Int,
                   yTargetBits: Int, cbTargetBits: Int, crTargetBits: Int, spatialFactorToUse: Int,
                   op1: ProcessingStep.Type, op2: ProcessingStep.Type, op3: ProcessingStep.Type): Unit = {
    val inputImage = ImageProcessorModel.readImage(inputImagePath)                       // :39
    val (w, h) = (inputImage.width, inputImage.height)
    val spatial = Seq(op1, op2, op3).contains(ProcessingStep.SpatialSampling)              // :43
    val outW = if (spatial) w / spatialFactorToUse else w                                  // :44
    val outH = if (spatial) h / spatialFactorToUse else h                                  // :45
    val rgb = new Array[Byte](w * h * 3)
    for (y <- 0 until h; x <- 0 until w) {                                                 // raster order of :77-89
      val px = inputImage.pixel(x, y); val i = 3 * (y * w + x)
      rgb(i) = px.red().toByte; rgb(i + 1) = px.green().toByte; rgb(i + 2) = px.blue().toByte
    }
    val gpu = new CsicGpu(0)
    val out = try gpu.process(w, h, chromaParamA, chromaParamB, yTargetBits, cbTargetBits, crTargetBits,
                              spatialFactorToUse, op1.litValue.toInt, op2.litValue.toInt, op3.litValue.toInt, rgb, 1)
              finally gpu.close()
    val img = ImmutableImage.filled(outW, outH, AwtColor.MAGENTA).copy()                   // :133
    for (i <- 0 until math.min(outW * outH, out.length / 3))                               // :134-142
      img.setColor(i % outW, i / outW, new com.sksamuel.scrimage.color.RGBColor(
        out(3 * i) & 0xFF, out(3 * i + 1) & 0xFF, out(3 * i + 2) & 0xFF, 255))
    ImageProcessorModel.writeImage(img, outputImagePath)                                   // :144
  }

  // --- Main execution with command-line argument parsing: flag for flag ImageCompressorTopApp.scala:149-190 ---
  val argsMap = args.sliding(2, 2).collect {                                               // :149-151
    case Array(key, value) if key.startsWith("--") => key -> value
  }.toMap

  def parseProcessingStep(name: String): ProcessingStep.Type = name.toLowerCase match {    // :154-161
    case "spatial" | "spatialsampling"  => ProcessingStep.SpatialSampling
    case "color" | "colorquantization"  => ProcessingStep.ColorQuantization
    case "chroma" | "chromasubsampling" => ProcessingStep.ChromaSubsampling
    case _ => throw new IllegalArgumentException(s"Unknown processing step: $name. Use 'spatial', 'color', or 'chroma'.")
  }

  val inputPath = argsMap.getOrElse("--input", "test_images/in128x128.png")                // :164-173: same defaults
  val selectedChromaParamA = argsMap.getOrElse("--a", "4").toInt
  val selectedChromaParamB = argsMap.getOrElse("--b", "4").toInt
  val yTargetQuantBits = argsMap.getOrElse("--yq", "8").toInt
  val cbTargetQuantBits = argsMap.getOrElse("--cbq", "8").toInt
  val crTargetQuantBits = argsMap.getOrElse("--crq", "8").toInt
  val selectedSpatialFactor = argsMap.getOrElse("--sf", "8").toInt
  val op1_choice = parseProcessingStep(argsMap.getOrElse("--op1", "spatial"))
  val op2_choice = parseProcessingStep(argsMap.getOrElse("--op2", "color"))
  val op3_choice = parseProcessingStep(argsMap.getOrElse("--op3", "chroma"))

  val imageName = new java.io.File(inputPath).getName.takeWhile(_ != '.')                  // :175

  println("----------------------------------------------------")                          // :177-185
  println("Image Compressor Application Parameters:")
  println("----------------------------------------------------")
  println(s"Input Image: $inputPath")
  println(s"Selected Chroma Subsampling (J:a:b): 4:$selectedChromaParamA:$selectedChromaParamB")
  println(s"Selected Quantization Bits (Y/Cb/Cr): $yTargetQuantBits/$cbTargetQuantBits/$crTargetQuantBits")
  println(s"Selected Spatial Downsampling Factor: $selectedSpatialFactor")
  println(s"Selected Pipeline Order: $op1_choice -> $op2_choice -> $op3_choice")
  println("----------------------------------------------------")

  val outputBaseDirName = "APP_OUTPUT"                                                      // :187-195
  val pipelineOrderString = s"order-${op1_choice.toString.split('.').last.take(2)}-${op2_choice.toString.split('.').last.take(2)}-${op3_choice.toString.split('.').last.take(2)}"
  val outputFileNameSuffix = s"chroma4-${selectedChromaParamA}-${selectedChromaParamB}_Y${yTargetQuantBits}Cb${cbTargetQuantBits}Cr${crTargetQuantBits}_sf${selectedSpatialFactor}_${pipelineOrderString}"
  val outputPath = s"$outputBaseDirName/${imageName}_processed_${outputFileNameSuffix}.png"
  val outputDir = new java.io.File(outputBaseDirName)
  if (!outputDir.exists()) outputDir.mkdirs()

  val inputFile = new java.io.File(inputPath)                                               // :197-215
  if (!inputFile.exists()) {
    println(s"[ERROR] Input image not found: $inputPath")
  } else {
    processImage(inputPath, outputPath, selectedChromaParamA, selectedChromaParamB, yTargetQuantBits, cbTargetQuantBits,
                 crTargetQuantBits, selectedSpatialFactor, op1_choice, op2_choice, op3_choice)
    println(s"Image processing complete. Output saved to: $outputPath")
  }
}
