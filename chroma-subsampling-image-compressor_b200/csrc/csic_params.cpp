// Host-only half of the C ABI: parameter construction, validation, geometry, names.
// Mirrors the reference's constructor `require(...)` predicates and their message texts
// (citations relative to the reference root, src/main/scala/jpeg/).  No CUDA in this file.
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstring>
#include <string>

#include "csic_internal.h"

namespace {

int fail(int code, char* msg, size_t n, const std::string& text) {
  if (msg && n) {
    std::snprintf(msg, n, "%s", text.c_str());
  }
  return code;
}

bool is_step(int v) { return v == CSIC_STEP_SPATIAL || v == CSIC_STEP_COLOR || v == CSIC_STEP_CHROMA; }

}  // namespace

extern "C" {

int csic_abi_version(void) { return CSIC_ABI_VERSION; }

// ImageCompressorTopApp.scala:164-173 defaults.
int csic_params_default(int32_t width, int32_t height, csic_params* out) {
  if (!out) return CSIC_EINVAL_ARG;
  std::memset(out, 0, sizeof(*out));
  out->width = width;
  out->height = height;
  out->chroma_a = 4;
  out->chroma_b = 4;
  out->y_bits = out->cb_bits = out->cr_bits = 8;
  out->factor = 8;
  out->op[0] = CSIC_STEP_SPATIAL;
  out->op[1] = CSIC_STEP_COLOR;
  out->op[2] = CSIC_STEP_CHROMA;
  out->round_mode = CSIC_ROUND_FLOOR;
  out->pool_mode = CSIC_POOL_DECIMATE;
  out->out_format = CSIC_OUT_YCC888;
  return csic_validate(out, nullptr, 0);
}

// ImageProcessor.scala:15-29 (params + requires, evaluated in the case class's order: width, height, factor,
// divisibility, chromaParamA, chromaParamB), :42-62 (toYC -> chroma -> spatial, no quantiser).
int csic_params_from_image_processor(int32_t width, int32_t height, int32_t factor, int32_t chroma_a,
                                     int32_t chroma_b, csic_params* out) {
  if (!out) return CSIC_EINVAL_ARG;
  std::memset(out, 0, sizeof(*out));
  out->width = width;
  out->height = height;
  out->chroma_a = chroma_a;
  out->chroma_b = chroma_b;
  out->y_bits = out->cb_bits = out->cr_bits = 8;
  out->factor = factor;
  out->op[0] = CSIC_STEP_CHROMA;
  out->op[1] = CSIC_STEP_SPATIAL;
  out->op[2] = CSIC_STEP_COLOR;   // 8/8/8 quantiser == identity; keeps op[] a permutation
  if (width <= 0 || height <= 0) return CSIC_EINVAL_DIMS;                                     // :22-23
  if (!(factor == 1 || factor == 2 || factor == 4 || factor == 8)) return CSIC_EINVAL_FACTOR;  // :24
  if (width % factor != 0 || height % factor != 0) return CSIC_EINVAL_DIVISIBLE;               // :25 (only here)
  if (!(chroma_a == 4 || chroma_a == 2 || chroma_a == 1)) return CSIC_EINVAL_CHROMA_A;         // :27
  if (!(chroma_b == chroma_a || chroma_b == 0)) return CSIC_EINVAL_CHROMA_B;                   // :28
  return csic_validate(out, nullptr, 0);
}

// Legacy enum surface (pinned by goldens G11-G13, G14-G22, G27).
int csic_params_from_legacy(int32_t width, int32_t height, int32_t chroma_mode, int32_t quant_mode,
                            int32_t factor, csic_params* out) {
  if (!out) return CSIC_EINVAL_ARG;
  static const int AB[3][2] = {{4, 4}, {2, 2}, {2, 0}};
  static const int Q[3][3] = {{8, 8, 8}, {6, 5, 5}, {3, 3, 2}};
  if (chroma_mode < 0 || chroma_mode > 2 || quant_mode < 0 || quant_mode > 2) return CSIC_EINVAL_MODE;
  std::memset(out, 0, sizeof(*out));
  out->width = width;
  out->height = height;
  out->chroma_a = AB[chroma_mode][0];
  out->chroma_b = AB[chroma_mode][1];
  out->y_bits = Q[quant_mode][0];
  out->cb_bits = Q[quant_mode][1];
  out->cr_bits = Q[quant_mode][2];
  out->factor = factor;
  out->op[0] = CSIC_STEP_CHROMA;
  out->op[1] = CSIC_STEP_COLOR;
  out->op[2] = CSIC_STEP_SPATIAL;
  return csic_validate(out, nullptr, 0);
}

int csic_validate(const csic_params* p, char* msg, size_t n) {
  if (!p) return fail(CSIC_EINVAL_ARG, msg, n, "params is NULL");
  // The checks run in the order the reference's constructor evaluates its `require`s, so that a parameter set that
  // is invalid in several ways reports the SAME first failure as `new ImageCompressorTop(...)`:
  //   ImageCompressorTop.scala:27-31 (ops) -> :45 new SpatialDownsampler (SpatialDownsampler.scala:7-8: dims, factor)
  //   -> :46-51 new ColorQuantizer (ColorQuantizer.scala:13-15: Y, Cb, Cr bits) -> :53-59 new ChromaSubsampler
  //   (ChromaSubsampler.scala:13-18: dims again, param_a, param_b).
  // ImageCompressorTop.scala:28-31
  for (int i = 0; i < 3; ++i)
    if (!is_step(p->op[i]))
      return fail(CSIC_EINVAL_OPS, msg, n,
                  "op" + std::to_string(i + 1) + "Type must be a valid reorderable operation.");
  if (p->op[0] == p->op[1] || p->op[0] == p->op[2] || p->op[1] == p->op[2])
    return fail(CSIC_EINVAL_OPS, msg, n, "op1, op2, and op3 types must be distinct and form a permutation.");
  // SpatialDownsampler.scala:7 (ChromaSubsampler.scala:13-14 can no longer fire after it)
  if (p->width <= 0 || p->height <= 0)
    return fail(CSIC_EINVAL_DIMS, msg, n, "Width and height must be positive");
  // SpatialDownsampler.scala:8
  if (!(p->factor == 1 || p->factor == 2 || p->factor == 4 || p->factor == 8))
    return fail(CSIC_EINVAL_FACTOR, msg, n, "Factor must be 1, 2, 4, or 8");
  // ColorQuantizer.scala:13-15 (originalBitWidth is fixed to 8 by the tops)
  const int bits[3] = {p->y_bits, p->cb_bits, p->cr_bits};
  const char* names[3] = {"Y", "Cb", "Cr"};
  for (int i = 0; i < 3; ++i)
    if (bits[i] < 1 || bits[i] > 8)
      return fail(CSIC_EINVAL_QBITS, msg, n,
                  std::string(names[i]) + " target bits must be between 1 and 8. Got " + std::to_string(bits[i]));
  // ChromaSubsampler.scala:17
  if (!(p->chroma_a == 4 || p->chroma_a == 2 || p->chroma_a == 1))
    return fail(CSIC_EINVAL_CHROMA_A, msg, n,
                "param_a must be 4, 2, or 1. Got " + std::to_string(p->chroma_a));
  // ChromaSubsampler.scala:18
  if (!(p->chroma_b == p->chroma_a || p->chroma_b == 0))
    return fail(CSIC_EINVAL_CHROMA_B, msg, n,
                "param_b must be equal to param_a (" + std::to_string(p->chroma_a) + ") or 0. Got " +
                    std::to_string(p->chroma_b));
  // ---- fields the reference does not have (extensions): checked last ----
  if (p->round_mode != CSIC_ROUND_FLOOR && p->round_mode != CSIC_ROUND_TRUNC)
    return fail(CSIC_EINVAL_MODE, msg, n, "round_mode must be 0 (FLOOR) or 1 (TRUNC)");
  if (p->pool_mode != CSIC_POOL_DECIMATE && p->pool_mode != CSIC_POOL_AVERAGE)
    return fail(CSIC_EINVAL_MODE, msg, n, "pool_mode must be 0 (DECIMATE) or 1 (AVERAGE)");
  if (p->out_format < CSIC_OUT_YCC888 || p->out_format > CSIC_OUT_PLANAR)
    return fail(CSIC_EINVAL_MODE, msg, n, "out_format must be 0..4");
  if (p->in_format < CSIC_IN_RGB24 || p->in_format > CSIC_IN_BGRA32)
    return fail(CSIC_EINVAL_MODE, msg, n, "in_format must be 0 (RGB24), 1 (RGBA32) or 2 (BGRA32)");
  if (p->reserved != 0) return fail(CSIC_EINVAL_MODE, msg, n, "reserved field must be 0");
  // The AVERAGE extension is defined on whole f x f blocks only (same predicate as ImageProcessor.scala:25).
  if (p->pool_mode == CSIC_POOL_AVERAGE && (p->width % p->factor != 0 || p->height % p->factor != 0))
    return fail(CSIC_EINVAL_DIVISIBLE, msg, n,
                "Image dimensions must be divisible by spatial downsampling factor.");
  if (p->out_format == CSIC_OUT_PLANAR) {
    int ic = 0, is = 0;
    for (int i = 0; i < 3; ++i) {
      if (p->op[i] == CSIC_STEP_CHROMA) ic = i;
      if (p->op[i] == CSIC_STEP_SPATIAL) is = i;
    }
    if (p->pool_mode != CSIC_POOL_DECIMATE || (p->factor > 1 && ic > is))
      return fail(CSIC_EINVAL_MODE, msg, n,
                  "PLANAR output needs ChromaSubsampling before SpatialSampling (or factor 1) and DECIMATE");
  }
  if (msg && n) msg[0] = '\0';
  return CSIC_OK;
}

int csic_planar_shape(const csic_params* p, int32_t* chroma_w, int32_t* chroma_h, size_t* cb_offset, size_t* cr_offset) {
  int rc = csic_validate(p, nullptr, 0);
  if (rc != CSIC_OK) return rc;
  if (p->out_format != CSIC_OUT_PLANAR) return CSIC_EINVAL_MODE;
  const csic::Geometry g = csic::geometry(*p);
  if (chroma_w) *chroma_w = g.planar_cw;
  if (chroma_h) *chroma_h = g.planar_ch;
  if (cb_offset) *cb_offset = (size_t)g.out_w * (size_t)g.out_h;
  if (cr_offset) *cr_offset = (size_t)g.out_w * (size_t)g.out_h + (size_t)g.planar_cw * (size_t)g.planar_ch;
  return CSIC_OK;
}

int csic_out_shape(const csic_params* p, int32_t* out_w, int32_t* out_h, size_t* out_row_bytes,
                   size_t* out_bytes_per_frame) {
  int rc = csic_validate(p, nullptr, 0);
  if (rc != CSIC_OK) return rc;
  csic::Geometry g = csic::geometry(*p);
  if (out_w) *out_w = g.out_w;
  if (out_h) *out_h = g.out_h;
  if (out_row_bytes) *out_row_bytes = g.out_row_bytes;
  if (out_bytes_per_frame) *out_bytes_per_frame = g.out_frame_bytes;
  return CSIC_OK;
}

// ImageCompressorTopApp.scala:154-161
int csic_parse_step(const char* name) {
  if (!name) return CSIC_EINVAL_OPS;
  std::string s;
  for (const char* c = name; *c; ++c) s.push_back((char)std::tolower((unsigned char)*c));
  if (s == "spatial" || s == "spatialsampling") return CSIC_STEP_SPATIAL;
  if (s == "color" || s == "colorquantization") return CSIC_STEP_COLOR;
  if (s == "chroma" || s == "chromasubsampling") return CSIC_STEP_CHROMA;
  return CSIC_EINVAL_OPS;
}

const char* csic_strerror(int status) {
  switch (status) {
    case CSIC_OK: return "ok";
    case CSIC_EINVAL_DIMS: return "Width and height must be positive";
    case CSIC_EINVAL_FACTOR: return "Factor must be 1, 2, 4, or 8";
    case CSIC_EINVAL_DIVISIBLE: return "Image dimensions must be divisible by spatial downsampling factor.";
    case CSIC_EINVAL_CHROMA_A: return "param_a must be 4, 2, or 1";
    case CSIC_EINVAL_CHROMA_B: return "param_b must be equal to param_a or 0";
    case CSIC_EINVAL_QBITS: return "target bits must be between 1 and 8";
    case CSIC_EINVAL_OPS: return "op1, op2, and op3 types must be distinct and form a permutation.";
    case CSIC_EINVAL_MODE: return "mode selector out of range";
    case CSIC_EINVAL_ARG: return "invalid argument";
    case CSIC_ENODEVICE: return "no usable CUDA device (there is no CPU fallback)";
    case CSIC_ECUDA: return "CUDA error (see csic_last_error)";
    case CSIC_ENOMEM: return "out of memory";
    default: return "unknown csic status";
  }
}

}  // extern "C"

namespace csic {

int slot_bits(const csic_params& p) {
  int t = p.y_bits + p.cb_bits + p.cr_bits;
  return t <= 8 ? 8 : (t <= 16 ? 16 : 32);
}

Geometry geometry(const csic_params& p) {
  Geometry g{};
  const int f = p.factor;
  g.out_w = (p.width + f - 1) / f;   // what the DUT emits: SpatialDownsamplerSpec.scala:120-123
  g.out_h = (p.height + f - 1) / f;
  g.in_px_bytes = p.in_format == CSIC_IN_RGB24 ? 3 : 4;
  g.in_row_bytes = (size_t)p.width * (size_t)g.in_px_bytes;
  g.in_frame_bytes = g.in_row_bytes * (size_t)p.height;
  if (p.out_format == CSIC_OUT_BUNDLE64 || p.out_format == CSIC_OUT_BUNDLE128) {
    const size_t word = p.out_format == CSIC_OUT_BUNDLE64 ? 8 : 16;
    g.out_px_bytes = slot_bits(p) / 8;
    const size_t bytes = (size_t)g.out_w * g.out_px_bytes;
    g.out_row_bytes = (bytes + word - 1) / word * word;
  } else if (p.out_format == CSIC_OUT_PLANAR) {
    g.out_px_bytes = 1;
    g.out_row_bytes = (size_t)g.out_w;            // a row of the Y plane
  } else {
    g.out_px_bytes = 3;
    g.out_row_bytes = (size_t)g.out_w * 3;
  }
  g.out_frame_bytes = g.out_row_bytes * (size_t)g.out_h;
  {
    const int hf = 4 / p.chroma_a, vf = (p.chroma_b == 0) ? 2 : 1;
    g.planar_hs = std::max(1, hf / f);
    g.planar_vs = std::max(1, vf / f);
    g.planar_cw = (g.out_w + g.planar_hs - 1) / g.planar_hs;
    g.planar_ch = (g.out_h + g.planar_vs - 1) / g.planar_vs;
    if (p.out_format == CSIC_OUT_PLANAR) g.out_frame_bytes += 2 * (size_t)g.planar_cw * (size_t)g.planar_ch;
  }
  int ic = -1, is = -1, iq = -1;
  for (int i = 0; i < 3; ++i) {
    if (p.op[i] == CSIC_STEP_CHROMA) ic = i;
    if (p.op[i] == CSIC_STEP_SPATIAL) is = i;
    if (p.op[i] == CSIC_STEP_COLOR) iq = i;
  }
  g.chroma_first = ic < is;
  g.quant_first = iq < is;
  g.hf = 4 / p.chroma_a;
  g.vf = (p.chroma_b == 0) ? 2 : 1;
  return g;
}

}  // namespace csic
