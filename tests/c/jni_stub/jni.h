/* Stand-in <jni.h> for images without a JDK (TEST INFRASTRUCTURE): exactly the JNI types and JNIEnv functions
 * bindings/jni/csic_jni.c uses, with the signatures of the JNI specification (Java SE 11, chapter 4).  The member
 * ORDER of the function table is not the JDK's -- shim and harness are both compiled against this header, so they
 * agree -- which is why this header is only good for tests/c/jni_harness.c, never for a real JVM. */
#ifndef CSIC_TEST_JNI_STUB_H_
#define CSIC_TEST_JNI_STUB_H_
#include <stdint.h>

#define JNIEXPORT __attribute__((visibility("default")))
#define JNICALL

typedef int32_t jint;
typedef int64_t jlong;
typedef int8_t jbyte;
typedef jint jsize;

struct fake_object;
typedef struct fake_object* jobject;
typedef jobject jclass;
typedef jobject jarray;
typedef jarray jintArray;
typedef jarray jbyteArray;

struct JNINativeInterface_;
typedef const struct JNINativeInterface_* JNIEnv;

struct JNINativeInterface_ {
  jclass (*FindClass)(JNIEnv* env, const char* name);
  jint (*ThrowNew)(JNIEnv* env, jclass clazz, const char* msg);
  jsize (*GetArrayLength)(JNIEnv* env, jarray array);
  void (*GetIntArrayRegion)(JNIEnv* env, jintArray array, jsize start, jsize len, jint* buf);
  void (*GetByteArrayRegion)(JNIEnv* env, jbyteArray array, jsize start, jsize len, jbyte* buf);
  void (*SetByteArrayRegion)(JNIEnv* env, jbyteArray array, jsize start, jsize len, const jbyte* buf);
  jobject (*NewDirectByteBuffer)(JNIEnv* env, void* address, jlong capacity);
  void* (*GetDirectBufferAddress)(JNIEnv* env, jobject buf);
  jlong (*GetDirectBufferCapacity)(JNIEnv* env, jobject buf);
};
#endif
