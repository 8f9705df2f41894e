package jpeg

/** JNI flavour of bindings/scala/CsicGpu.scala for JDK < 21 (no java.lang.foreign); native side: csic_jni.c.
  * Source only -- this image has no JDK.  `params` = the 16 ints of `csic_params` (include/csic.h) in order:
  * width, height, chroma_a, chroma_b, y_bits, cb_bits, cr_bits, factor, op1, op2, op3 (ProcessingStep ids,
  * ImageCompressorTop.scala:7-9), round_mode, pool_mode, out_format, in_format, 0. */
object CsicJni {
  System.loadLibrary("csic_jni")
  @native def create(device: Int): Long
  @native def destroy(ctx: Long): Unit
  @native def outBytesPerFrame(params: Array[Int]): Long           // IllegalArgumentException as the Scala `require`s
  @native def processHost(ctx: Long, params: Array[Int], rgb: Array[Byte], nFrames: Long, out: Array[Byte]): Unit
}

/** Drop-in for the body of ImageCompressionApp.processImage (ImageCompressorTopApp.scala:53-131). */
final class CsicGpuJni(device: Int = 0) extends AutoCloseable {
  private val ctx = CsicJni.create(device)
  def process(width: Int, height: Int, a: Int, b: Int, yBits: Int, cbBits: Int, crBits: Int, factor: Int,
              op1: Int, op2: Int, op3: Int, rgb: Array[Byte], nFrames: Int, outFormat: Int = 1): Array[Byte] = {
    val params = Array(width, height, a, b, yBits, cbBits, crBits, factor, op1, op2, op3, 0, 0, outFormat, 0, 0)
    val out = new Array[Byte]((CsicJni.outBytesPerFrame(params) * nFrames).toInt)
    CsicJni.processHost(ctx, params, rgb, nFrames.toLong, out)
    out
  }
  override def close(): Unit = CsicJni.destroy(ctx)
}
