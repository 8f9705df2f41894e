package jpeg

import java.nio.ByteBuffer

/** JNI flavour of bindings/scala/CsicGpu.scala for JDK < 21 (no java.lang.foreign); native side: csic_jni.c.
  * The authoring image has no JDK; the C side is compiled and exercised there through a fake JNIEnv
  * (tests/test_bindings.py), this file is compiled by bindings/ci/scala-bindings.yml.
  * `params` = the 16 ints of `csic_params` (include/csic.h) in order: width, height, chroma_a, chroma_b, y_bits,
  * cb_bits, cr_bits, factor, op1, op2, op3 (ProcessingStep ids, ImageCompressorTop.scala:7-9), round_mode, pool_mode,
  * out_format, in_format, 0. */
object CsicJni {
  System.loadLibrary("csic_jni")
  @native def create(device: Int): Long
  @native def destroy(ctx: Long): Unit
  @native def outBytesPerFrame(params: Array[Int]): Long           // IllegalArgumentException as the Scala `require`s
  @native def processHost(ctx: Long, params: Array[Int], rgb: Array[Byte], nFrames: Long, out: Array[Byte]): Unit
  @native def processHostDirect(ctx: Long, params: Array[Int], rgb: ByteBuffer, nFrames: Long, out: ByteBuffer): Unit
  @native def hostAlloc(bytes: Long): ByteBuffer                    // pinned memory as a direct buffer
  @native def hostFree(buf: ByteBuffer): Unit
}

/** Drop-in for the body of ImageCompressionApp.processImage (ImageCompressorTopApp.scala:53-131). */
final class CsicGpuJni(device: Int = 0) extends AutoCloseable {
  private val ctx = CsicJni.create(device)
  def process(width: Int, height: Int, a: Int, b: Int, yBits: Int, cbBits: Int, crBits: Int, factor: Int,
              op1: Int, op2: Int, op3: Int, rgb: Array[Byte], nFrames: Int, outFormat: Int = 1): Array[Byte] = {
    val params = Array(width, height, a, b, yBits, cbBits, crBits, factor, op1, op2, op3, 0, 0, outFormat, 0, 0)
    val total = Math.multiplyExact(CsicJni.outBytesPerFrame(params), nFrames.toLong)
    require(total <= Int.MaxValue, s"output of $total bytes does not fit a JVM array: use processHostDirect")
    val out = new Array[Byte](total.toInt)
    CsicJni.processHost(ctx, params, rgb, nFrames.toLong, out)
    out
  }
  override def close(): Unit = CsicJni.destroy(ctx)
}
