// Panama (java.lang.foreign, JDK 21+) binding of include/csic.h for the reference's Scala code base.
// The authoring image has no JVM, so this file has never been compiled there; bindings/ci/scala-bindings.yml is the
// job that compiles it (against the reference's own build.sbt) wherever a JDK exists.  It is the file a maintainer
// drops into src/test/scala/jpeg/ next to ImageCompressorTopApp.scala (see INTEGRATION.md).
package jpeg

import java.lang.foreign._
import java.lang.foreign.ValueLayout._
import java.lang.invoke.MethodHandle

/** One instance per (thread, GPU).  Replaces `chiseltest.RawTester.test(new ImageCompressorTop(...)){...}`
  * in ImageCompressionApp.processImage (ImageCompressorTopApp.scala:53-131). */
final class CsicGpu(device: Int = 0) extends AutoCloseable {
  import CsicGpu._
  // lives as long as the instance: only the context handle.  Everything a call needs is allocated from an arena that
  // is closed when the call returns, so a long-running host does not grow native memory call by call.
  private val arena = Arena.ofConfined()
  private val ctx: MemorySegment = {
    val out = arena.allocate(ADDRESS)
    check(create.invoke(device, out).asInstanceOf[Int], null)
    out.get(ADDRESS, 0)
  }

  /** The 16 ints of `csic_params` for the reference's constructor arguments (ImageCompressorTop.scala:11-25). */
  private def params(a: Arena, width: Int, height: Int, ca: Int, cb: Int, yBits: Int, cbBits: Int, crBits: Int,
                     factor: Int, op1: Int, op2: Int, op3: Int, outFormat: Int): MemorySegment = {
    val p = a.allocate(PARAMS)
    val ints = Array(width, height, ca, cb, yBits, cbBits, crBits, factor, op1, op2, op3, 0, 0, outFormat, 0, 0)
    ints.zipWithIndex.foreach { case (v, i) => p.setAtIndex(JAVA_INT, i.toLong, v) }
    val msg = a.allocate(256)
    check(validate.invoke(p, msg, 256L).asInstanceOf[Int], msg)       // IllegalArgumentException like `require`
    p
  }

  /** rgb: nFrames * H * W * 3 bytes in raster order (pixel.red/green/blue, :86-89).
    * outFormat 0 = Y,Cb,Cr bytes (what the DUT emits), 1 = R,G,B after the fused YCbCrUtils.ycbcr2rgb (:118). */
  def process(width: Int, height: Int, a: Int, b: Int, yBits: Int, cbBits: Int, crBits: Int, factor: Int,
              op1: Int, op2: Int, op3: Int, rgb: Array[Byte], nFrames: Int, outFormat: Int = 1): Array[Byte] = {
    require(nFrames > 0, "nFrames must be positive")
    require(rgb.length.toLong == nFrames.toLong * width * height * 3, s"rgb must hold $nFrames frames of ${width}x$height RGB24")
    val call = Arena.ofConfined()
    try {
      val p = params(call, width, height, a, b, yBits, cbBits, crBits, factor, op1, op2, op3, outFormat)
      val fb = call.allocate(JAVA_LONG)
      check(outShape.invoke(p, MemorySegment.NULL, MemorySegment.NULL, MemorySegment.NULL, fb).asInstanceOf[Int], null)
      val total = Math.multiplyExact(fb.get(JAVA_LONG, 0), nFrames.toLong)
      require(total <= Int.MaxValue, s"output of $total bytes does not fit a JVM array")
      val in = call.allocate(rgb.length.toLong); in.copyFrom(MemorySegment.ofArray(rgb))
      val out = call.allocate(total)
      check(processHost.invoke(ctx, p, in, nFrames.toLong, out).asInstanceOf[Int], null)
      out.toArray(JAVA_BYTE)
    } finally call.close()
  }

  /** Output geometry: (outWidth, outHeight) = ceil(W/f) x ceil(H/f), what the DUT emits (SpatialDownsamplerSpec.scala:120-123). */
  def outSize(width: Int, height: Int, factor: Int): (Int, Int) = ((width + factor - 1) / factor, (height + factor - 1) / factor)

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
  private val lastError   = fn("csic_last_error",   FunctionDescriptor.of(ADDRESS))

  /** CSIC_EINVAL_* (-1..-9) are the reference's `require` failures: rethrow what Scala's require throws. */
  private def check(rc: Int, msg: MemorySegment): Unit = if (rc != 0) {
    def cstr(h: MethodHandle, args: Int*): String =
      (if (args.isEmpty) h.invoke() else h.invoke(args.head)).asInstanceOf[MemorySegment].reinterpret(512).getUtf8String(0)
    val text = if (msg != null && msg.get(JAVA_BYTE, 0) != 0) msg.getUtf8String(0)
               else if (rc == -11) cstr(lastError) else cstr(strerror, rc)
    if (rc >= -9) throw new IllegalArgumentException(s"requirement failed: $text")
    else throw new RuntimeException(s"csic status $rc: $text")
  }
}
