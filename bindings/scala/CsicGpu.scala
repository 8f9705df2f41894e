// Panama (java.lang.foreign, JDK 21+) binding of include/csic.h for the reference's Scala code base.
// SOURCE ONLY: the authoring image has no JVM, so this file has never been compiled here; it is the file a
// maintainer drops into src/test/scala/jpeg/ next to ImageCompressorTopApp.scala (see INTEGRATION.md).
package jpeg

import java.lang.foreign._
import java.lang.foreign.ValueLayout._
import java.lang.invoke.MethodHandle

/** One instance per (thread, GPU).  Replaces `chiseltest.RawTester.test(new ImageCompressorTop(...)){...}`
  * in ImageCompressionApp.processImage (ImageCompressorTopApp.scala:53-131). */
final class CsicGpu(device: Int = 0) extends AutoCloseable {
  import CsicGpu._
  private val arena = Arena.ofConfined()
  private val ctx: MemorySegment = {
    val out = arena.allocate(ADDRESS)
    check(create.invoke(device, out).asInstanceOf[Int], null)
    out.get(ADDRESS, 0)
  }

  /** rgb: nFrames * H * W * 3 bytes in raster order (pixel.red/green/blue, :86-89).
    * outFormat 0 = Y,Cb,Cr bytes (what the DUT emits), 1 = R,G,B after the fused YCbCrUtils.ycbcr2rgb (:118). */
  def process(width: Int, height: Int, a: Int, b: Int, yBits: Int, cbBits: Int, crBits: Int, factor: Int,
              op1: Int, op2: Int, op3: Int, rgb: Array[Byte], nFrames: Int, outFormat: Int = 1): Array[Byte] = {
    val p = arena.allocate(PARAMS)
    val ints = Array(width, height, a, b, yBits, cbBits, crBits, factor, op1, op2, op3, 0, 0, outFormat, 0, 0)
    ints.zipWithIndex.foreach { case (v, i) => p.setAtIndex(JAVA_INT, i.toLong, v) }
    val msg = arena.allocate(256)
    check(validate.invoke(p, msg, 256L).asInstanceOf[Int], msg)
    val fb = arena.allocate(JAVA_LONG)
    check(outShape.invoke(p, MemorySegment.NULL, MemorySegment.NULL, MemorySegment.NULL, fb).asInstanceOf[Int], null)
    val in  = arena.allocate(rgb.length.toLong); in.copyFrom(MemorySegment.ofArray(rgb))
    val out = arena.allocate(fb.get(JAVA_LONG, 0) * nFrames)
    check(processHost.invoke(ctx, p, in, nFrames.toLong, out).asInstanceOf[Int], null)
    out.toArray(JAVA_BYTE)
  }

  override def close(): Unit = { destroy.invoke(ctx); arena.close() }
}

object CsicGpu {
  private val linker = Linker.nativeLinker()
  private val lib = SymbolLookup.libraryLookup(sys.props.getOrElse("csic.lib", "libcsic.so"), Arena.global())
  private def fn(name: String, d: FunctionDescriptor): MethodHandle = linker.downcallHandle(lib.find(name).get, d)
  val PARAMS: MemoryLayout = MemoryLayout.sequenceLayout(16, JAVA_INT)   // struct csic_params
  private val create      = fn("csic_create",       FunctionDescriptor.of(JAVA_INT, JAVA_INT, ADDRESS))
  private val destroy     = fn("csic_destroy",      FunctionDescriptor.of(JAVA_INT, ADDRESS))
  private val validate    = fn("csic_validate",     FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG))
  private val outShape    = fn("csic_out_shape",    FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS))
  private val processHost = fn("csic_process_host", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS))
  private val strerror    = fn("csic_strerror",     FunctionDescriptor.of(ADDRESS, JAVA_INT))

  /** CSIC_EINVAL_* (-1..-9) are the reference's `require` failures: rethrow what Scala's require throws. */
  private def check(rc: Int, msg: MemorySegment): Unit = if (rc != 0) {
    val text = if (msg != null && msg.get(JAVA_BYTE, 0) != 0) msg.getUtf8String(0)
               else strerror.invoke(rc).asInstanceOf[MemorySegment].reinterpret(256).getUtf8String(0)
    if (rc >= -9) throw new IllegalArgumentException(s"requirement failed: $text")
    else throw new RuntimeException(s"csic status $rc: $text")
  }
}
