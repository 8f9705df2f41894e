/*
 * csic_jni.c -- JNI shim over include/csic.h for JVMs without java.lang.foreign (JDK < 21).
 *
 * Build on a machine with a JDK:
 *   gcc -shared -fPIC -I"$JAVA_HOME/include" -I"$JAVA_HOME/include/linux" -I../../include \
 *       csic_jni.c -L../../chroma-subsampling-image-compressor_b200 -lcsic -o libcsic_jni.so
 * This image has no JDK: tests/test_bindings.py compiles the file against a stand-in <jni.h> (tests/c/jni_stub) and
 * drives every entry point through a fake JNIEnv (tests/c/jni_harness.c), on the CPU for the error paths and on the
 * GPU for the data path.  The Scala side is bindings/jni/CsicJni.scala; it mirrors bindings/scala/CsicGpu.scala
 * (Panama) method for method.
 *
 * What it replaces in the reference: the body of ImageCompressionApp.processImage between reading the pixels and
 * writing the PNG (src/test/scala/jpeg/ImageCompressorTopApp.scala:53-131).  Errors follow the reference: a failed
 * `require` is an IllegalArgumentException (csic_validate carries the reference's message text).
 *
 * No GetPrimitiveArrayCritical: csic_process_host allocates, spawns threads and blocks on CUDA for milliseconds, which
 * JNI forbids inside a critical region (GCLocker would stall every thread that needs a GC).  Heap arrays are copied
 * with Get/SetByteArrayRegion through a native staging buffer; callers that want zero copies allocate direct buffers
 * over pinned memory with hostAlloc and use processHostDirect.
 */
#include <jni.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "csic.h"

static void throw_for(JNIEnv* env, int rc, const char* msg) {
  const char* cls = (rc >= CSIC_EINVAL_ARG) ? "java/lang/IllegalArgumentException" : "java/lang/RuntimeException";
  char text[320];
  if (rc >= CSIC_EINVAL_ARG) snprintf(text, sizeof text, "requirement failed: %s", (msg && msg[0]) ? msg : csic_strerror(rc));
  else snprintf(text, sizeof text, "csic status %d: %s", rc, rc == CSIC_ECUDA ? csic_last_error() : csic_strerror(rc));
  (*env)->ThrowNew(env, (*env)->FindClass(env, cls), text);
}

/* params: the 16 ints of csic_params in declaration order (ImageCompressorTop.scala:11-25 + build-side selectors) */
static int load_params(JNIEnv* env, jintArray jparams, csic_params* p) {
  if (!jparams || (*env)->GetArrayLength(env, jparams) != 16) return CSIC_EINVAL_ARG;
  (*env)->GetIntArrayRegion(env, jparams, 0, 16, (jint*)p);
  return CSIC_OK;
}

/* Validates the parameters and the frame count and returns the byte sizes of the batch; 0 = ok, otherwise the
 * exception has been thrown.  Every product is checked for overflow (a negative nFrames must not wrap). */
static int batch_sizes(JNIEnv* env, jintArray jparams, jlong nFrames, csic_params* p, size_t* in_bytes, size_t* out_bytes) {
  char msg[256] = {0};
  int rc = load_params(env, jparams, p);
  if (rc == CSIC_OK) rc = csic_validate(p, msg, sizeof msg);
  if (rc != CSIC_OK) { throw_for(env, rc, msg); return rc; }
  if (nFrames <= 0) { throw_for(env, CSIC_EINVAL_ARG, "nFrames must be positive"); return CSIC_EINVAL_ARG; }
  size_t fb = 0;
  rc = csic_out_shape(p, NULL, NULL, NULL, &fb);
  if (rc != CSIC_OK) { throw_for(env, rc, NULL); return rc; }
  const uint64_t frame_in = (uint64_t)p->width * (uint64_t)p->height * (p->in_format == CSIC_IN_RGB24 ? 3u : 4u);
  const uint64_t n = (uint64_t)nFrames;
  if ((frame_in && n > UINT64_MAX / frame_in) || (fb && n > UINT64_MAX / fb) || n * frame_in > (uint64_t)SIZE_MAX ||
      n * fb > (uint64_t)SIZE_MAX) {
    throw_for(env, CSIC_EINVAL_ARG, "nFrames * frame size overflows");
    return CSIC_EINVAL_ARG;
  }
  *in_bytes = (size_t)(n * frame_in);
  *out_bytes = (size_t)(n * fb);
  return CSIC_OK;
}

JNIEXPORT jlong JNICALL Java_jpeg_CsicJni_create(JNIEnv* env, jclass cls, jint device) {
  (void)cls;
  csic_ctx* ctx = NULL;
  int rc = csic_create((int)device, &ctx);
  if (rc != CSIC_OK) { throw_for(env, rc, NULL); return 0; }
  return (jlong)(uintptr_t)ctx;
}

JNIEXPORT void JNICALL Java_jpeg_CsicJni_destroy(JNIEnv* env, jclass cls, jlong ctx) {
  (void)env; (void)cls;
  csic_destroy((csic_ctx*)(uintptr_t)ctx);
}

/* bytes per output frame for these parameters; throws what the reference's constructors throw */
JNIEXPORT jlong JNICALL Java_jpeg_CsicJni_outBytesPerFrame(JNIEnv* env, jclass cls, jintArray jparams) {
  (void)cls;
  csic_params p;
  char msg[256] = {0};
  int rc = load_params(env, jparams, &p);
  if (rc == CSIC_OK) rc = csic_validate(&p, msg, sizeof msg);
  size_t fb = 0;
  if (rc == CSIC_OK) rc = csic_out_shape(&p, NULL, NULL, NULL, &fb);
  if (rc != CSIC_OK) { throw_for(env, rc, msg); return 0; }
  return (jlong)fb;
}

/* rgb: nFrames * H * W * (3|4) bytes (pixel.red/green/blue, ImageCompressorTopApp.scala:86-89); out: nFrames * bytes
 * per frame.  The arrays are copied through a native staging buffer (pageable: csic_process_host moves it through its
 * own pinned bounce buffers on several threads). */
JNIEXPORT void JNICALL Java_jpeg_CsicJni_processHost(JNIEnv* env, jclass cls, jlong ctx, jintArray jparams,
                                                     jbyteArray rgb, jlong nFrames, jbyteArray out) {
  (void)cls;
  csic_params p;
  size_t in_bytes = 0, out_bytes = 0;
  if (batch_sizes(env, jparams, nFrames, &p, &in_bytes, &out_bytes) != CSIC_OK) return;
  if (!rgb || !out || (uint64_t)(*env)->GetArrayLength(env, rgb) < in_bytes || (uint64_t)(*env)->GetArrayLength(env, out) < out_bytes) {
    throw_for(env, CSIC_EINVAL_ARG, "rgb / out array shorter than nFrames frames");
    return;
  }
  uint8_t* stage = (uint8_t*)malloc(in_bytes + out_bytes + 1);
  if (!stage) { throw_for(env, CSIC_ENOMEM, NULL); return; }
  (*env)->GetByteArrayRegion(env, rgb, 0, (jsize)in_bytes, (jbyte*)stage);
  const int rc = csic_process_host((csic_ctx*)(uintptr_t)ctx, &p, stage, (size_t)nFrames, stage + in_bytes);
  if (rc == CSIC_OK) (*env)->SetByteArrayRegion(env, out, 0, (jsize)out_bytes, (const jbyte*)(stage + in_bytes));
  free(stage);
  if (rc != CSIC_OK) throw_for(env, rc, NULL);
}

/* The same call on direct ByteBuffers: no copy on the JVM side.  Buffers from hostAlloc are pinned, so the library
 * DMAs straight from / into them. */
JNIEXPORT void JNICALL Java_jpeg_CsicJni_processHostDirect(JNIEnv* env, jclass cls, jlong ctx, jintArray jparams,
                                                           jobject rgb, jlong nFrames, jobject out) {
  (void)cls;
  csic_params p;
  size_t in_bytes = 0, out_bytes = 0;
  if (batch_sizes(env, jparams, nFrames, &p, &in_bytes, &out_bytes) != CSIC_OK) return;
  void* in = rgb ? (*env)->GetDirectBufferAddress(env, rgb) : NULL;
  void* o = out ? (*env)->GetDirectBufferAddress(env, out) : NULL;
  if (!in || !o || (uint64_t)(*env)->GetDirectBufferCapacity(env, rgb) < in_bytes ||
      (uint64_t)(*env)->GetDirectBufferCapacity(env, out) < out_bytes) {
    throw_for(env, CSIC_EINVAL_ARG, "rgb / out must be direct ByteBuffers of at least nFrames frames");
    return;
  }
  const int rc = csic_process_host((csic_ctx*)(uintptr_t)ctx, &p, (const uint8_t*)in, (size_t)nFrames, (uint8_t*)o);
  if (rc != CSIC_OK) throw_for(env, rc, NULL);
}

/* Pinned host memory (csic_host_alloc) as a direct ByteBuffer; release it with hostFree. */
JNIEXPORT jobject JNICALL Java_jpeg_CsicJni_hostAlloc(JNIEnv* env, jclass cls, jlong bytes) {
  (void)cls;
  void* ptr = NULL;
  if (bytes <= 0) { throw_for(env, CSIC_EINVAL_ARG, "bytes must be positive"); return NULL; }
  const int rc = csic_host_alloc((size_t)bytes, &ptr);
  if (rc != CSIC_OK) { throw_for(env, rc, NULL); return NULL; }
  jobject buf = (*env)->NewDirectByteBuffer(env, ptr, bytes);
  if (!buf) csic_host_free(ptr);
  return buf;
}

JNIEXPORT void JNICALL Java_jpeg_CsicJni_hostFree(JNIEnv* env, jclass cls, jobject buf) {
  (void)cls;
  void* ptr = buf ? (*env)->GetDirectBufferAddress(env, buf) : NULL;
  if (ptr) csic_host_free(ptr);
}
