/* Pure-C caller of include/csic.h: proves the boundary is a plain C ABI (no Python, no C++ types).
 * Built by tests/test_c_driver.py with:  gcc -std=c11 -Iinclude abi_driver.c -L<pkg> -lcsic
 *
 *   abi_driver host                 parameter / validation checks only; exits 0 (no GPU needed)
 *   abi_driver gpu W H a b f fmt    fills n=3 frames with a fixed pattern, runs csic_process_host and
 *                                   csic_process_band, prints "fnv <hex>" of the output (compared with the
 *                                   oracle by the test) */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "csic.h"

static uint64_t fnv1a(const uint8_t* p, size_t n) {
  uint64_t h = 1469598103934665603ull;
  for (size_t i = 0; i < n; ++i) { h ^= p[i]; h *= 1099511628211ull; }
  return h;
}

static int expect(int cond, const char* what) {
  if (!cond) { fprintf(stderr, "FAIL: %s\n", what); return 1; }
  return 0;
}

static int host_checks(void) {
  int bad = 0;
  char msg[128];
  csic_params p;
  bad |= expect(csic_abi_version() == CSIC_ABI_VERSION, "abi version");
  bad |= expect(sizeof(csic_params) == 64, "sizeof(csic_params) == 64");
  bad |= expect(csic_params_default(128, 128, &p) == CSIC_OK && p.factor == 8 && p.op[0] == CSIC_STEP_SPATIAL, "defaults");
  bad |= expect(csic_params_from_legacy(128, 128, CSIC_CHROMA_420, CSIC_Q_8BIT, 1, &p) == CSIC_OK && p.chroma_a == 2 &&
                    p.chroma_b == 0 && p.y_bits == 3 && p.cb_bits == 3 && p.cr_bits == 2, "legacy enums");
  p.factor = 3;
  bad |= expect(csic_validate(&p, msg, sizeof msg) == CSIC_EINVAL_FACTOR && strcmp(msg, "Factor must be 1, 2, 4, or 8") == 0,
                "factor message");
  p.factor = 2; p.chroma_b = 1;
  bad |= expect(csic_validate(&p, msg, sizeof msg) == CSIC_EINVAL_CHROMA_B &&
                    strcmp(msg, "param_b must be equal to param_a (2) or 0. Got 1") == 0, "param_b message");
  bad |= expect(csic_params_from_image_processor(10, 16, 4, 4, 4, &p) == CSIC_EINVAL_DIVISIBLE, "ImageProcessorParams divisibility");
  bad |= expect(csic_parse_step("ChromaSubsampling") == CSIC_STEP_CHROMA && csic_parse_step("blur") == CSIC_EINVAL_OPS, "parse_step");
  int32_t w, h; size_t rb, fb;
  csic_params_from_legacy(3840, 2160, CSIC_CHROMA_420, CSIC_Q_24BIT, 2, &p);
  p.out_format = CSIC_OUT_BUNDLE128;
  bad |= expect(csic_out_shape(&p, &w, &h, &rb, &fb) == CSIC_OK && w == 1920 && h == 1080 && rb == 7680 && fb == 8294400, "out_shape");
  if (csic_device_count() <= 0) {
    csic_ctx* c = NULL;
    bad |= expect(csic_create(0, &c) == CSIC_ENODEVICE && c == NULL, "no device -> CSIC_ENODEVICE, never a CPU path");
  }
  printf(bad ? "host checks FAILED\n" : "host checks ok\n");
  return bad;
}

int main(int argc, char** argv) {
  if (argc >= 2 && strcmp(argv[1], "host") == 0) return host_checks();
  if (argc < 8 || strcmp(argv[1], "gpu") != 0) { fprintf(stderr, "usage: abi_driver host | gpu W H a b f fmt\n"); return 2; }
  const int W = atoi(argv[2]), H = atoi(argv[3]), a = atoi(argv[4]), b = atoi(argv[5]), f = atoi(argv[6]), fmt = atoi(argv[7]);
  const size_t n = 3;
  csic_params p;
  int rc = csic_params_default(W, H, &p);
  p.chroma_a = a; p.chroma_b = b; p.factor = f; p.out_format = fmt;
  p.y_bits = 6; p.cb_bits = 5; p.cr_bits = 5;
  p.op[0] = CSIC_STEP_CHROMA; p.op[1] = CSIC_STEP_SPATIAL; p.op[2] = CSIC_STEP_COLOR;
  char msg[128];
  if ((rc = csic_validate(&p, msg, sizeof msg)) != CSIC_OK) { fprintf(stderr, "invalid: %s\n", msg); return 3; }
  int32_t ow, oh; size_t rb, fb;
  csic_out_shape(&p, &ow, &oh, &rb, &fb);
  const size_t in_bytes = n * (size_t)W * H * 3;
  uint8_t *in = NULL, *out = NULL, *out2 = NULL;
  if (csic_host_alloc(in_bytes, (void**)&in) || csic_host_alloc(n * fb, (void**)&out)) { fprintf(stderr, "host_alloc: %s\n", csic_last_error()); return 4; }
  out2 = (uint8_t*)calloc(n, fb);
  uint32_t s = 12345u;                                   /* LCG pattern the test replays in NumPy */
  for (size_t i = 0; i < in_bytes; ++i) { s = s * 1664525u + 1013904223u; in[i] = (uint8_t)(s >> 24); }
  csic_ctx* ctx = NULL;
  if ((rc = csic_create(0, &ctx)) != CSIC_OK) { fprintf(stderr, "csic_create: %s / %s\n", csic_strerror(rc), csic_last_error()); return 5; }
  if ((rc = csic_process_host(ctx, &p, in, n, out)) != CSIC_OK) { fprintf(stderr, "process_host: %s / %s\n", csic_strerror(rc), csic_last_error()); return 6; }
  int32_t fam = 0; int64_t launches = 0;
  csic_last_kernel(ctx, &fam, &launches);
  printf("family %d launches %lld\n", fam, (long long)launches);
  printf("fnv %016llx\n", (unsigned long long)fnv1a(out, n * fb));
  /* force the other kernel family through the same ABI: must give the same bytes */
  csic_set_option(ctx, CSIC_OPT_KERNEL_FAMILY, 1);
  if ((rc = csic_process_host(ctx, &p, in, n, out2)) != CSIC_OK) return 7;
  printf("generic %s\n", memcmp(out, out2, n * fb) == 0 ? "same" : "DIFFERENT");
  csic_destroy(ctx);
  csic_host_free(in); csic_host_free(out); free(out2);
  return 0;
}
