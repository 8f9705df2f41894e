// Drop-in twin of `object ImageCompressionApp` (src/test/scala/jpeg/ImageCompressorTopApp.scala) whose
// processImage calls libcsic.so instead of simulating the Chisel DUT.  SOURCE ONLY (no JVM in the authoring
// image).  Same signature as ImageCompressorTopApp.scala:23-37, same PNG in / PNG out via ImageProcessorModel.
package jpeg

import com.sksamuel.scrimage.ImmutableImage
import java.awt.{Color => AwtColor}

object ImageCompressionAppGpu {
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
}
