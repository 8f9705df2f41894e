/*
 * csic_oracle.c -- CPU restatement of the reference's pixel pipeline.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library -- as the checker or the timed CPU baseline, never as the product path.
 * libcsic.so (the product) does not link, include or call anything in this directory.
 *
 * Form: deliberately the slow, literal one.  Every stage is a sequential state machine over the
 * raster stream -- counters and latched chroma exactly as the Chisel RTL holds them -- and the
 * pipeline is "run stage op1 over the whole stream, then op2, then op3".  The CUDA kernels use
 * closed-form gathers instead, so this is an independent derivation of the same semantics.
 *
 * Parity pinning: the Scala/Chisel reference cannot run here (no JVM, no sbt, no RTL simulator),
 * so this oracle is pinned against (a) every known-answer vector in the reference's own tests and
 * (b) all 29 PNGs the reference committed as outputs of its RTL simulations and stage benches
 * (tests/golden/MANIFEST.json; tests/test_oracle_golden.py).  Cases no reference artefact pins are
 * listed in DESIGN.md ("parity unpinned": spatial-before-chroma with f>1, (a,b) in {(4,0),(1,0)},
 * non-divisible sizes through the top, BUNDLE word layout, AVERAGE pooling).
 *
 * All paths below are relative to the reference root, src/main/scala/jpeg/ unless stated.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

#include "../include/csic.h"

typedef struct { uint8_t y, cb, cr; } ycc_t;   /* PixelYCbCrBundle, PixelBundle.scala:11-15 */

static inline int clamp255(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }

/* Arithmetic shift right by 8 on a possibly negative int == floor(v / 256).
 * RGB2YCbCr.scala:50-52 `numerator >> 8` on SInt; ReferenceModel.scala:15-17 `>> 8` on Int. */
static inline int asr8(int v) { return (v >= 0) ? (v >> 8) : -(((-v) + 255) >> 8); }

/* Scala `/ 256` on Int truncates toward zero.  RGB2YCbCr.scala:111-113 (YCbCrUtils.rgbToYCbCr). */
static inline int div256_trunc(int v) { return v / 256; }

/* RGB2YCbCr.scala:33-35 (MAC), :55-65 (bias, shift, +128 AFTER the shift), :74-76 (clamp last).
 * FLOOR: class RGB2YCbCr / ReferenceModel.rgb2ycbcr (ReferenceModel.scala:8-19).
 * TRUNC: object YCbCrUtils.rgbToYCbCr (RGB2YCbCr.scala:95-121). */
void csic_oracle_rgb2ycbcr(int r, int g, int b, int round_mode, int* y, int* cb, int* cr) {
  int yi = 77 * r + 150 * g + 29 * b;
  int cbi = -43 * r - 85 * g + 128 * b;
  int cri = 128 * r - 107 * g - 21 * b;
  if (round_mode == CSIC_ROUND_TRUNC) {
    *y = clamp255(div256_trunc(yi + 128));
    *cb = clamp255(div256_trunc(cbi + 128) + 128);
    *cr = clamp255(div256_trunc(cri + 128) + 128);
  } else {
    *y = clamp255(asr8(yi + 128));
    *cb = clamp255(asr8(cbi + 128) + 128);
    *cr = clamp255(asr8(cri + 128) + 128);
  }
}

/* YCbCrUtils.ycbcr2rgb -- RGB2YCbCr.scala:123-132 == YCbCr2RGB.scala:17-26.
 * c = y (no -16), d = cb-128, e = cr-128; `>> 8` is arithmetic on negative Ints; clamp after. */
void csic_oracle_ycbcr2rgb(int y, int cb, int cr, int* r, int* g, int* b) {
  int c = y, d = cb - 128, e = cr - 128;
  *r = clamp255(asr8(298 * c + 409 * e + 128));
  *g = clamp255(asr8(298 * c - 100 * d - 208 * e + 128));
  *b = clamp255(asr8(298 * c + 516 * d + 128));
}

/* The same function over an array of n interleaved (Y,Cb,Cr) byte triples -> (R,G,B) byte triples: lets a test walk
 * the whole 2^24 YCbCr cube (the clamps of YCbCr2RGB.scala:17-26 fire on a large part of it, most of which no RGB
 * input ever reaches through the forward transform). */
void csic_oracle_ycbcr2rgb_array(const uint8_t* ycc, size_t n, uint8_t* rgb) {
  for (size_t i = 0; i < n; ++i) {
    int r, g, b;
    csic_oracle_ycbcr2rgb(ycc[3 * i], ycc[3 * i + 1], ycc[3 * i + 2], &r, &g, &b);
    rgb[3 * i] = (uint8_t)r; rgb[3 * i + 1] = (uint8_t)g; rgb[3 * i + 2] = (uint8_t)b;
  }
}

/* ColorQuantizer.scala:29-31 (shift = 8 - targetBits), :42-44 ((v >> s) << s). */
int csic_oracle_quant(int v, int target_bits) {
  int s = 8 - target_bits;
  return (v >> s) << s;
}

/* ---- stream stages ----------------------------------------------------------------------------
 * Each consumes a stream of n triples and writes its output stream; returns the output count.
 * W,H are the *configured* sizes handed to the module constructors, which are always the input
 * frame's (ImageCompressorTop.scala:44,52-58) even when an earlier stage already shortened the
 * stream -- that mismatch is part of the reference's behaviour. */

/* ChromaSubsampler.scala:26-27 (factors), :34-35 (latches, reset 0), :37-38 (counters advance on every
 * accepted pixel, wrap at W and H), :52-65 (sample / hold).  Identical to the benches' software model
 * subsampleChromaSw, src/test/scala/jpeg/ChromaSubsamplerImageSpec.scala:45-78. */
size_t csic_oracle_chroma_stream(const ycc_t* in, size_t n, int W, int H, int a, int b, ycc_t* out) {
  int hf = 4 / a;
  int vf = (b == 0 && a != 0) ? 2 : 1;
  int pixel_counter = 0, line_counter = 0;
  uint8_t last_cb = 0, last_cr = 0;
  for (size_t i = 0; i < n; ++i) {
    ycc_t px = in[i];
    int sample = (pixel_counter % hf == 0) && (line_counter % vf == 0);
    if (sample) {
      last_cb = px.cb;
      last_cr = px.cr;
    } else {
      px.cb = last_cb;
      px.cr = last_cr;
    }
    out[i] = px;
    if (++pixel_counter == W) {            /* Counter(io.dataIn.fire, imageWidth) */
      pixel_counter = 0;
      if (++line_counter == H) line_counter = 0;
    }
  }
  return n;
}

/* SpatialDownsampler.scala:17-31 (counters), :33-45 (sampleH && sampleV on the low bits),
 * :55 (payload forwarded unchanged).  sof/eol (:11-12) are never read. */
size_t csic_oracle_spatial_stream(const ycc_t* in, size_t n, int W, int H, int f, ycc_t* out) {
  int col = 0, row = 0;
  size_t m = 0;
  for (size_t i = 0; i < n; ++i) {
    if ((col & (f - 1)) == 0 && (row & (f - 1)) == 0) out[m++] = in[i];
    if (col == W - 1) {
      col = 0;
      row = (row == H - 1) ? 0 : row + 1;
    } else {
      ++col;
    }
  }
  return m;
}

/* AVERAGE extension (not in the reference; README.md:44 mentions "average pooling" in prose only).
 * Mean of each f x f block of the stage's input stream, per channel, round half up:
 * (sum + f*f/2) >> (2*log2 f); emitted in raster order of the blocks.  Needs n == W*H, W%f==H%f==0. */
static size_t spatial_average_stream(const ycc_t* in, size_t n, int W, int H, int f, ycc_t* out) {
  if (n != (size_t)W * H) return 0;
  int sh = 0;
  while ((1 << sh) < f) ++sh;
  sh *= 2;
  int half = (f * f) / 2;
  size_t m = 0;
  for (int r = 0; r < H; r += f)
    for (int c = 0; c < W; c += f) {
      int sy = 0, scb = 0, scr = 0;
      for (int dr = 0; dr < f; ++dr)
        for (int dc = 0; dc < f; ++dc) {
          const ycc_t* p = &in[(size_t)(r + dr) * W + (c + dc)];
          sy += p->y; scb += p->cb; scr += p->cr;
        }
      out[m].y = (uint8_t)((sy + half) >> sh);
      out[m].cb = (uint8_t)((scb + half) >> sh);
      out[m].cr = (uint8_t)((scr + half) >> sh);
      ++m;
    }
  return m;
}

/* ColorQuantizer.scala:35-47: pointwise, one register stage. */
size_t csic_oracle_quant_stream(const ycc_t* in, size_t n, int yb, int cbb, int crb, ycc_t* out) {
  for (size_t i = 0; i < n; ++i) {
    out[i].y = (uint8_t)csic_oracle_quant(in[i].y, yb);
    out[i].cb = (uint8_t)csic_oracle_quant(in[i].cb, cbb);
    out[i].cr = (uint8_t)csic_oracle_quant(in[i].cr, crb);
  }
  return n;
}

/* ---- geometry of the build-defined output formats (include/csic.h, csic_out_format) ---------- */
static int slot_bits(const csic_params* p) {
  int t = p->y_bits + p->cb_bits + p->cr_bits;
  return t <= 8 ? 8 : (t <= 16 ? 16 : 32);
}

/* CSIC_OUT_PLANAR chroma decimation in output pixels (include/csic.h). */
static void planar_factors(const csic_params* p, int* hs, int* vs) {
  int hf = 4 / p->chroma_a, vf = (p->chroma_b == 0) ? 2 : 1, f = p->factor;
  *hs = hf / f > 1 ? hf / f : 1;
  *vs = vf / f > 1 ? vf / f : 1;
}

static void out_geometry(const csic_params* p, int* ow, int* oh, size_t* row_bytes) {
  int f = p->factor;
  *ow = (p->width + f - 1) / f;
  *oh = (p->height + f - 1) / f;
  if (p->out_format == CSIC_OUT_PLANAR) { *row_bytes = (size_t)*ow; return; }
  if (p->out_format == CSIC_OUT_BUNDLE64 || p->out_format == CSIC_OUT_BUNDLE128) {
    size_t word = p->out_format == CSIC_OUT_BUNDLE64 ? 8 : 16;
    size_t bytes = (size_t)*ow * slot_bits(p) / 8;
    *row_bytes = (bytes + word - 1) / word * word;
  } else {
    *row_bytes = (size_t)*ow * 3;
  }
}

size_t csic_oracle_out_bytes_per_frame(const csic_params* p) {
  int ow, oh; size_t rb;
  out_geometry(p, &ow, &oh, &rb);
  if (p->out_format == CSIC_OUT_PLANAR) {
    int hs, vs;
    planar_factors(p, &hs, &vs);
    return (size_t)ow * oh + 2 * (size_t)((ow + hs - 1) / hs) * ((oh + vs - 1) / vs);
  }
  return rb * oh;
}

/* One frame through the top.  ImageCompressorTop.scala:80-81 (toYC first), :83-114 (op1, op2, op3 as
 * configured); ImageProcessor.scala:42-62 is the special case chroma -> spatial with 8/8/8 bits. */
static int process_frame(const csic_params* p, const uint8_t* rgb, uint8_t* out, ycc_t* s0, ycc_t* s1) {
  const int W = p->width, H = p->height;
  size_t n = (size_t)W * H;
  /* pixel.red()/green()/blue() -- alpha is never read (ImageCompressorTopApp.scala:86-89) */
  const size_t ipb = p->in_format == CSIC_IN_RGB24 ? 3 : 4;
  const int ir = p->in_format == CSIC_IN_BGRA32 ? 2 : 0, ib = p->in_format == CSIC_IN_BGRA32 ? 0 : 2;
  for (size_t i = 0; i < n; ++i) {           /* raster order, ImageCompressorTopApp.scala:77-89 */
    int y, cb, cr;
    csic_oracle_rgb2ycbcr(rgb[ipb * i + ir], rgb[ipb * i + 1], rgb[ipb * i + ib], p->round_mode, &y, &cb, &cr);
    s0[i].y = (uint8_t)y; s0[i].cb = (uint8_t)cb; s0[i].cr = (uint8_t)cr;
  }
  ycc_t *cur = s0, *nxt = s1;
  for (int k = 0; k < 3; ++k) {
    switch (p->op[k]) {
      case CSIC_STEP_SPATIAL:
        if (p->pool_mode == CSIC_POOL_AVERAGE && p->factor > 1)
          n = spatial_average_stream(cur, n, W, H, p->factor, nxt);
        else
          n = csic_oracle_spatial_stream(cur, n, W, H, p->factor, nxt);
        break;
      case CSIC_STEP_COLOR:
        n = csic_oracle_quant_stream(cur, n, p->y_bits, p->cb_bits, p->cr_bits, nxt);
        break;
      case CSIC_STEP_CHROMA:
        n = csic_oracle_chroma_stream(cur, n, W, H, p->chroma_a, p->chroma_b, nxt);
        break;
      default:
        return CSIC_EINVAL_OPS;
    }
    ycc_t* t = cur; cur = nxt; nxt = t;
  }
  int ow, oh; size_t row_bytes;
  out_geometry(p, &ow, &oh, &row_bytes);
  if (n != (size_t)ow * oh) return CSIC_EINVAL_DIMS;

  if (p->out_format == CSIC_OUT_PLANAR) {
    /* Y plane, then the chroma of the surviving sample points: output pixel (rc*vs, cc*hs) carries its own
     * (just sampled) chroma in the stream, so the planes can be read straight off it. */
    int hs, vs;
    planar_factors(p, &hs, &vs);
    int cw = (ow + hs - 1) / hs, ch = (oh + vs - 1) / vs;
    uint8_t *yp = out, *cbp = out + (size_t)ow * oh, *crp = cbp + (size_t)cw * ch;
    for (size_t i = 0; i < n; ++i) yp[i] = cur[i].y;
    for (int rc = 0; rc < ch; ++rc)
      for (int cc = 0; cc < cw; ++cc) {
        const ycc_t* q = &cur[(size_t)(rc * vs) * ow + (size_t)cc * hs];
        cbp[(size_t)rc * cw + cc] = q->cb;
        crp[(size_t)rc * cw + cc] = q->cr;
      }
    return CSIC_OK;
  }
  if (p->out_format == CSIC_OUT_YCC888) {
    for (size_t i = 0; i < n; ++i) { out[3 * i] = cur[i].y; out[3 * i + 1] = cur[i].cb; out[3 * i + 2] = cur[i].cr; }
  } else if (p->out_format == CSIC_OUT_RGB888) {     /* ImageCompressorTopApp.scala:118 */
    for (size_t i = 0; i < n; ++i) {
      int r, g, b;
      csic_oracle_ycbcr2rgb(cur[i].y, cur[i].cb, cur[i].cr, &r, &g, &b);
      out[3 * i] = (uint8_t)r; out[3 * i + 1] = (uint8_t)g; out[3 * i + 2] = (uint8_t)b;
    }
  } else {
    int sb = slot_bits(p) / 8;
    memset(out, 0, row_bytes * oh);
    for (int r = 0; r < oh; ++r)
      for (int c = 0; c < ow; ++c) {
        const ycc_t* q = &cur[(size_t)r * ow + c];
        uint32_t v = ((uint32_t)(q->y >> (8 - p->y_bits)) << (p->cb_bits + p->cr_bits)) |
                     ((uint32_t)(q->cb >> (8 - p->cb_bits)) << p->cr_bits) |
                     (uint32_t)(q->cr >> (8 - p->cr_bits));
        uint8_t* d = out + (size_t)r * row_bytes + (size_t)c * sb;
        for (int k = 0; k < sb; ++k) d[k] = (uint8_t)(v >> (8 * k));   /* little endian */
      }
  }
  return CSIC_OK;
}

typedef struct {
  const csic_params* p; const uint8_t* rgb; uint8_t* out;
  size_t f0, f1, in_stride, out_stride; int rc;
} job_t;

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  size_t n = (size_t)j->p->width * j->p->height;
  ycc_t* s0 = (ycc_t*)malloc(n * sizeof(ycc_t));
  ycc_t* s1 = (ycc_t*)malloc(n * sizeof(ycc_t));
  j->rc = (s0 && s1) ? CSIC_OK : CSIC_ENOMEM;
  for (size_t k = j->f0; k < j->f1 && j->rc == CSIC_OK; ++k)
    j->rc = process_frame(j->p, j->rgb + k * j->in_stride, j->out + k * j->out_stride, s0, s1);
  free(s0); free(s1);
  return NULL;
}

/* n_frames independent frames (the reference elaborates a fresh DUT per image,
 * ImageCompressorTopApp.scala:53), split across `threads` host threads.  Parameters are assumed
 * valid (tests validate through libcsic's csic_validate, which mirrors the reference's requires). */
int csic_oracle_process(const csic_params* p, const uint8_t* rgb, size_t n_frames, uint8_t* out, int threads) {
  if (!p || !rgb || !out) return CSIC_EINVAL_ARG;
  if (threads < 1) threads = 1;
  if ((size_t)threads > n_frames) threads = n_frames ? (int)n_frames : 1;
  size_t in_stride = (size_t)p->width * p->height * (p->in_format == CSIC_IN_RGB24 ? 3 : 4), out_stride = csic_oracle_out_bytes_per_frame(p);
  job_t* jobs = (job_t*)calloc((size_t)threads, sizeof(job_t));
  pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
  if (!jobs || !th) { free(jobs); free(th); return CSIC_ENOMEM; }
  for (int t = 0; t < threads; ++t) {
    jobs[t] = (job_t){p, rgb, out, n_frames * t / threads, n_frames * (t + 1) / threads, in_stride, out_stride, 0};
    if (t > 0) pthread_create(&th[t], NULL, worker, &jobs[t]);
  }
  worker(&jobs[0]);
  int rc = jobs[0].rc;
  for (int t = 1; t < threads; ++t) { pthread_join(th[t], NULL); if (jobs[t].rc) rc = jobs[t].rc; }
  free(jobs); free(th);
  return rc;
}
