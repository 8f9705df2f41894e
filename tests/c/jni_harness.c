/* Drives bindings/jni/csic_jni.c through a fake JNIEnv (TEST INFRASTRUCTURE, see tests/c/jni_stub/jni.h).
 *   jni_harness errors          CPU only: every argument / `require` failure path throws the right exception
 *   jni_harness run W H out.bin GPU: one synthetic frame through processHost (arrays) and processHostDirect (pinned
 *                               direct buffers); both outputs must agree and are written to out.bin for the Python
 *                               side to compare with the oracle
 */
#include <jni.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct fake_object { int kind; size_t len; void* data; const char* name; };   /* kind: 1 class, 2 int[], 3 byte[], 4 direct buffer */

static char g_thrown_class[128], g_thrown_msg[512];
static int g_thrown;

static jclass f_FindClass(JNIEnv* env, const char* name) {
  (void)env;
  struct fake_object* o = calloc(1, sizeof *o);
  o->kind = 1; o->name = name;
  return o;
}
static jint f_ThrowNew(JNIEnv* env, jclass c, const char* msg) {
  (void)env;
  g_thrown = 1;
  snprintf(g_thrown_class, sizeof g_thrown_class, "%s", c->name);
  snprintf(g_thrown_msg, sizeof g_thrown_msg, "%s", msg);
  return 0;
}
static jsize f_GetArrayLength(JNIEnv* env, jarray a) { (void)env; return (jsize)a->len; }
static void f_GetIntArrayRegion(JNIEnv* env, jintArray a, jsize s, jsize n, jint* buf) { (void)env; memcpy(buf, (jint*)a->data + s, (size_t)n * 4); }
static void f_GetByteArrayRegion(JNIEnv* env, jbyteArray a, jsize s, jsize n, jbyte* buf) { (void)env; memcpy(buf, (jbyte*)a->data + s, (size_t)n); }
static void f_SetByteArrayRegion(JNIEnv* env, jbyteArray a, jsize s, jsize n, const jbyte* buf) { (void)env; memcpy((jbyte*)a->data + s, buf, (size_t)n); }
static jobject f_NewDirectByteBuffer(JNIEnv* env, void* p, jlong cap) {
  (void)env;
  struct fake_object* o = calloc(1, sizeof *o);
  o->kind = 4; o->len = (size_t)cap; o->data = p;
  return o;
}
static void* f_GetDirectBufferAddress(JNIEnv* env, jobject b) { (void)env; return b->kind == 4 ? b->data : NULL; }
static jlong f_GetDirectBufferCapacity(JNIEnv* env, jobject b) { (void)env; return b->kind == 4 ? (jlong)b->len : -1; }

static const struct JNINativeInterface_ g_table = {f_FindClass, f_ThrowNew, f_GetArrayLength, f_GetIntArrayRegion, f_GetByteArrayRegion,
                                                   f_SetByteArrayRegion, f_NewDirectByteBuffer, f_GetDirectBufferAddress,
                                                   f_GetDirectBufferCapacity};
static JNIEnv g_env = &g_table;

jlong Java_jpeg_CsicJni_create(JNIEnv*, jclass, jint);
void Java_jpeg_CsicJni_destroy(JNIEnv*, jclass, jlong);
jlong Java_jpeg_CsicJni_outBytesPerFrame(JNIEnv*, jclass, jintArray);
void Java_jpeg_CsicJni_processHost(JNIEnv*, jclass, jlong, jintArray, jbyteArray, jlong, jbyteArray);
void Java_jpeg_CsicJni_processHostDirect(JNIEnv*, jclass, jlong, jintArray, jobject, jlong, jobject);
jobject Java_jpeg_CsicJni_hostAlloc(JNIEnv*, jclass, jlong);
void Java_jpeg_CsicJni_hostFree(JNIEnv*, jclass, jobject);

static struct fake_object* int_array(const jint* v, size_t n) {
  struct fake_object* o = calloc(1, sizeof *o);
  o->kind = 2; o->len = n; o->data = malloc(n * 4 + 4);
  memcpy(o->data, v, n * 4);
  return o;
}
static struct fake_object* byte_array(size_t n) {
  struct fake_object* o = calloc(1, sizeof *o);
  o->kind = 3; o->len = n; o->data = calloc(1, n + 1);
  return o;
}
static int expect(const char* what, const char* cls, const char* needle) {
  const int ok = g_thrown && strstr(g_thrown_class, cls) && strstr(g_thrown_msg, needle);
  printf("%s %s: %s | %s\n", ok ? "ok  " : "FAIL", what, g_thrown ? g_thrown_class : "(nothing thrown)", g_thrown_msg);
  g_thrown = 0; g_thrown_msg[0] = 0; g_thrown_class[0] = 0;
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "errors")) {
    int bad = 0;
    /* width,height,a,b,y,cb,cr,f,op1,op2,op3,round,pool,outfmt,infmt,0 */
    jint ok16[16] = {16, 8, 2, 0, 8, 8, 8, 2, 3, 1, 2, 0, 0, 1, 0, 0};
    struct fake_object* good = int_array(ok16, 16);
    jlong fb = Java_jpeg_CsicJni_outBytesPerFrame(&g_env, NULL, good);
    printf("%s outBytesPerFrame(16x8 f=2 RGB888) = %lld\n", (fb == 8 * 4 * 3 && !g_thrown) ? "ok  " : "FAIL", (long long)fb);
    bad += !(fb == 96 && !g_thrown);
    jint f3[16]; memcpy(f3, ok16, sizeof f3); f3[7] = 3;
    Java_jpeg_CsicJni_outBytesPerFrame(&g_env, NULL, int_array(f3, 16));
    bad += expect("factor 3", "IllegalArgumentException", "requirement failed: Factor must be 1, 2, 4, or 8");
    jint a3[16]; memcpy(a3, ok16, sizeof a3); a3[2] = 3;
    Java_jpeg_CsicJni_outBytesPerFrame(&g_env, NULL, int_array(a3, 16));
    bad += expect("param_a 3", "IllegalArgumentException", "param_a must be 4, 2, or 1. Got 3");
    Java_jpeg_CsicJni_outBytesPerFrame(&g_env, NULL, int_array(ok16, 15));
    bad += expect("15 ints", "IllegalArgumentException", "requirement failed");
    struct fake_object *in = byte_array(16 * 8 * 3), *out = byte_array(96);
    Java_jpeg_CsicJni_processHost(&g_env, NULL, 0, good, in, -1, out);
    bad += expect("nFrames -1", "IllegalArgumentException", "nFrames must be positive");
    Java_jpeg_CsicJni_processHost(&g_env, NULL, 0, good, in, 0x4000000000000000LL, out);
    bad += expect("nFrames 2^62", "IllegalArgumentException", "overflows");
    Java_jpeg_CsicJni_processHost(&g_env, NULL, 0, good, in, 2, out);
    bad += expect("short arrays", "IllegalArgumentException", "shorter than nFrames frames");
    Java_jpeg_CsicJni_processHostDirect(&g_env, NULL, 0, good, in, 1, out);     /* heap arrays are not direct buffers */
    bad += expect("not direct", "IllegalArgumentException", "direct ByteBuffers");
    Java_jpeg_CsicJni_hostAlloc(&g_env, NULL, 0);
    bad += expect("hostAlloc(0)", "IllegalArgumentException", "bytes must be positive");
    return bad ? 1 : 0;
  }
  if (argc >= 5 && !strcmp(argv[1], "run")) {
    const int W = atoi(argv[2]), H = atoi(argv[3]);
    jint p16[16] = {W, H, 2, 0, 6, 5, 5, 2, 3, 1, 2, 0, 0, 1, 0, 0};     /* 4:2:0, 6/5/5, f=2, chroma-spatial-color, RGB888 */
    struct fake_object* params = int_array(p16, 16);
    const jlong fb = Java_jpeg_CsicJni_outBytesPerFrame(&g_env, NULL, params);
    const size_t nin = (size_t)W * H * 3;
    struct fake_object *in = byte_array(nin), *out = byte_array((size_t)fb);
    unsigned s = 12345u;
    for (size_t i = 0; i < nin; ++i) { s = s * 1664525u + 1013904223u; ((unsigned char*)in->data)[i] = (unsigned char)(s >> 24); }
    const jlong ctx = Java_jpeg_CsicJni_create(&g_env, NULL, 0);
    if (g_thrown) { printf("create threw: %s\n", g_thrown_msg); return 2; }
    Java_jpeg_CsicJni_processHost(&g_env, NULL, ctx, params, in, 1, out);
    if (g_thrown) { printf("processHost threw: %s\n", g_thrown_msg); return 3; }
    jobject din = Java_jpeg_CsicJni_hostAlloc(&g_env, NULL, (jlong)nin), dout = Java_jpeg_CsicJni_hostAlloc(&g_env, NULL, fb);
    if (g_thrown || !din || !dout) { printf("hostAlloc threw: %s\n", g_thrown_msg); return 4; }
    memcpy(din->data, in->data, nin);
    Java_jpeg_CsicJni_processHostDirect(&g_env, NULL, ctx, params, din, 1, dout);
    if (g_thrown) { printf("processHostDirect threw: %s\n", g_thrown_msg); return 5; }
    if (memcmp(dout->data, out->data, (size_t)fb)) { printf("array path and direct path disagree\n"); return 6; }
    FILE* f = fopen(argv[4], "wb");
    fwrite(in->data, 1, nin, f); fwrite(out->data, 1, (size_t)fb, f); fclose(f);
    Java_jpeg_CsicJni_hostFree(&g_env, NULL, din); Java_jpeg_CsicJni_hostFree(&g_env, NULL, dout);
    Java_jpeg_CsicJni_destroy(&g_env, NULL, ctx);
    printf("ok run %dx%d: %lld output bytes\n", W, H, (long long)fb);
    return 0;
  }
  fprintf(stderr, "usage: jni_harness errors | run W H out.bin\n");
  return 64;
}
