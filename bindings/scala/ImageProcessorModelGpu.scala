// GPU-backed twin of `object ImageProcessorModel` (src/test/scala/jpeg/ImageProcessorModel.scala:9-53): the same five
// entry points with the same signatures -- readImage, writeImage (both overloads), getImageParams, getImagePixels --
// plus `process`, which is what the reference's specs obtain by pushing the pixels of getImagePixels through a
// simulated `ImageProcessor` DUT one handshake at a time (SpatialDownsamplerSpec.scala:155-230).  Compiled by
// bindings/ci/scala-bindings.yml; never compiled in the authoring image (no JVM there).
package jpeg

import java.io.File
import com.sksamuel.scrimage.{ImmutableImage, MutableImage}
import com.sksamuel.scrimage.nio.PngWriter
import com.sksamuel.scrimage.pixels.Pixel

object ImageProcessorModelGpu {

  type PixelType = ImageProcessorModel.PixelType      // Seq[Int]
  type ImageType = ImageProcessorModel.ImageType      // Seq[Seq[PixelType]]

  def readImage(file: String): ImmutableImage = ImageProcessorModel.readImage(file)                               // :14-16

  def writeImage(image: MutableImage, file: String): Unit = ImageProcessorModel.writeImage(image, file)           // :18-22

  def writeImage(image: Array[Pixel], p: ImageProcessorParams, file: String): Unit =                              // :24-28
    ImageProcessorModel.writeImage(image, p, file)

  def getImageParams(image: ImmutableImage, numPixelsPerCycle: Int): ImageProcessorParams =                       // :33-41
    ImageProcessorModel.getImageParams(image, numPixelsPerCycle)

  def getImagePixels(image: ImmutableImage): ImageType = ImageProcessorModel.getImagePixels(image)                // :43-52

  /** Packed RGB24 in raster order: what `pixel.red()/green()/blue()` yields (ImageCompressorTopApp.scala:86-89). */
  def packRgb(image: ImmutableImage): Array[Byte] = {
    val (w, h) = (image.width, image.height)
    val rgb = new Array[Byte](w * h * 3)
    var i = 0
    for (r <- 0 until h; c <- 0 until w) {
      val px = image.pixel(c, r)
      rgb(i) = px.red().toByte; rgb(i + 1) = px.green().toByte; rgb(i + 2) = px.blue().toByte
      i += 3
    }
    rgb
  }

  /** `ImageProcessor(p)` (ImageProcessor.scala:31-63: toYC -> chroma -> spatial, no quantiser) on the GPU: returns the
    * output pixels in raster order, already converted back to RGB (YCbCrUtils.ycbcr2rgb, RGB2YCbCr.scala:123-132), ready
    * for writeImage(pixels, outParams, file).  outParams carries the output size. */
  def process(image: ImmutableImage, p: ImageProcessorParams, gpu: CsicGpu): (Array[Pixel], ImageProcessorParams) = {
    require(image.width == p.width && image.height == p.height, "image size must match ImageProcessorParams")
    // ProcessingStep ids (ImageCompressorTop.scala:7-9): SpatialSampling 1, ColorQuantization 2, ChromaSubsampling 3
    val out = gpu.process(p.width, p.height, p.chromaParamA, p.chromaParamB, 8, 8, 8, p.factor, 3, 1, 2, packRgb(image), 1, outFormat = 1)
    val (ow, oh) = gpu.outSize(p.width, p.height, p.factor)
    val pixels = Array.tabulate(ow * oh) { i =>
      new Pixel(i % ow, i / ow, out(3 * i) & 0xFF, out(3 * i + 1) & 0xFF, out(3 * i + 2) & 0xFF, 255)
    }
    (pixels, p.copy(width = ow, height = oh, factor = 1))
  }
}
