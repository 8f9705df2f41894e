/*
 * csic_jni.c -- JNI shim over include/csic.h for JVMs without java.lang.foreign (JDK < 21).
 *
 * Source only: this image has no JDK (no jni.h), so the file is not built here; on a machine with one:
 *   gcc -shared -fPIC -I"$JAVA_HOME/include" -I"$JAVA_HOME/include/linux" -I../../include \
 *       csic_jni.c -L../../chroma-subsampling-image-compressor_b200 -lcsic -o libcsic_jni.so
 * The Scala side is bindings/jni/CsicJni.scala; it mirrors bindings/scala/CsicGpu.scala (Panama) method for method.
 *
 * What it replaces in the reference: the body of ImageCompressionApp.processImage between reading the pixels and
 * writing the PNG (src/test/scala/jpeg/ImageCompressorTopApp.scala:53-131).  Errors follow the reference: a failed
 * `require` is an IllegalArgumentException (csic_validate carries the reference's message text).
 */
#include <jni.h>
#include <stdint.h>
#include <stdio.h>
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
  if ((*env)->GetArrayLength(env, jparams) != 16) return CSIC_EINVAL_ARG;
  (*env)->GetIntArrayRegion(env, jparams, 0, 16, (jint*)p);
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
 * per frame.  Heap arrays are pageable memory: csic_process_host stages them through its own pinned bounce buffers. */
JNIEXPORT void JNICALL Java_jpeg_CsicJni_processHost(JNIEnv* env, jclass cls, jlong ctx, jintArray jparams,
                                                     jbyteArray rgb, jlong nFrames, jbyteArray out) {
  (void)cls;
  csic_params p;
  char msg[256] = {0};
  int rc = load_params(env, jparams, &p);
  if (rc == CSIC_OK) rc = csic_validate(&p, msg, sizeof msg);
  if (rc != CSIC_OK) { throw_for(env, rc, msg); return; }
  size_t fb = 0;
  csic_out_shape(&p, NULL, NULL, NULL, &fb);
  const size_t in_bytes = (size_t)nFrames * (size_t)p.width * (size_t)p.height * (p.in_format == CSIC_IN_RGB24 ? 3u : 4u);
  if ((size_t)(*env)->GetArrayLength(env, rgb) < in_bytes || (size_t)(*env)->GetArrayLength(env, out) < (size_t)nFrames * fb) {
    throw_for(env, CSIC_EINVAL_ARG, "rgb / out array shorter than nFrames frames");
    return;
  }
  jbyte* in = (*env)->GetPrimitiveArrayCritical(env, rgb, NULL);
  jbyte* o = in ? (*env)->GetPrimitiveArrayCritical(env, out, NULL) : NULL;
  rc = (in && o) ? csic_process_host((csic_ctx*)(uintptr_t)ctx, &p, (const uint8_t*)in, (size_t)nFrames, (uint8_t*)o)
                 : CSIC_ENOMEM;
  if (o) (*env)->ReleasePrimitiveArrayCritical(env, out, o, 0);
  if (in) (*env)->ReleasePrimitiveArrayCritical(env, rgb, in, JNI_ABORT);
  if (rc != CSIC_OK) throw_for(env, rc, NULL);
}
