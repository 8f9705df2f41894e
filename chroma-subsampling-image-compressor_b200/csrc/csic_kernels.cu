// CUDA kernels of the pixel pipeline, written for sm_100a (B200).
//
//   csic_rows_kernel<F, FMT>   the hot path.  Persistent CTAs; packed RGB24 row segments are staged into a
//                              multi-stage shared-memory ring by the TMA engine (cp.async.bulk + mbarrier
//                              complete_tx), each thread converts granules of 4 output pixels with dp4a,
//                              resolves chroma sample-and-hold inside the granule (or from one TMA-fetched
//                              held pixel per row), quantises, packs, and the tile leaves through a
//                              double-buffered shared-memory staging area with a TMA bulk store.
//   csic_generic_kernel        one thread per output slot, closed-form gather; any legal parameter set
//                              (odd sizes, unaligned pointers, AVERAGE extension).  Not a fallback to the CPU:
//                              it is the same GPU path for shapes the TMA kernel's alignment rules exclude.
//
// Semantics (bit-exact with the reference; citations relative to its root, src/main/scala/jpeg/):
//   forward   RGB2YCbCr.scala:33-35,50-65,74-76 (FLOOR)   RGB2YCbCr.scala:95-121 (TRUNC)
//   chroma    ChromaSubsampler.scala:26-27,34-38,52-65      spatial  SpatialDownsampler.scala:17-55
//   quant     ColorQuantizer.scala:29-31,42-44              inverse  RGB2YCbCr.scala:123-132
//   order / misaligned counters   ImageCompressorTop.scala:43-58,83-114
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "csic_internal.h"

namespace csic {

// ------------------------------------------------------------------------------------------------
// Forward transform on one packed pixel word p = R | G<<8 | B<<16 | (don't care)<<24.
//
//   Y  = (77R + 150G + 29B + 128) >> 8                      never clamps (max 255)
//   Cb = clamp(((-43R - 85G + 128B + 128) >> 8) + 128)      only 256 -> 255 ever clamps
// The chroma rows are evaluated *negated* so that every coefficient fits a signed byte for dp4a:
//   x  = max(43R + 85G - 128B + 32639, 0)   ==  65535 - (cbi + 128 + 32768)   (clamped)
//   Cb = 255 - (x >> 8)                     ==  ~byte1(x)
// which equals the reference for all 2^24 colours (tests/test_device_math.py replays this identity
// exhaustively; the kernel itself is checked against the oracle on the full colour cube).
// TRUNC (Scala `/ 256`, toward zero) differs from floor only for negative numerators, i.e. x >= 32768:
//   x -= 255 there.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int32_t dp4a_us(uint32_t a, uint32_t b_s8x4, int32_t c) {
  int32_t d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b_s8x4), "r"(c));
  return d;
}

constexpr uint32_t kCoefY = 0x001D964Du;     //  77, 150,  29, 0   (u8)
constexpr uint32_t kCoefNCb = 0x0080552Bu;   //  43,  85,-128, 0   (s8)  == -cb row
constexpr uint32_t kCoefNCr = 0x00156B80u;   //-128, 107,  21, 0   (s8)  == -cr row

// byte 1 of the result is Y
__device__ __forceinline__ uint32_t fwd_y16(uint32_t p) { return dp4a_uu(p, kCoefY, 128u); }
// byte 1 of the result is ~Cb / ~Cr; result < 65536
template <bool TRUNC>
__device__ __forceinline__ uint32_t fwd_nc16(uint32_t p, uint32_t coef) {
  int32_t x = max(dp4a_us(p, coef, 32639), 0);
  if (TRUNC) x -= (x >> 15) * 255;
  return (uint32_t)x;
}

__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }

// YCbCrUtils.ycbcr2rgb with the -128 offsets folded into the constants.  Returns R | G<<8 | B<<16.
__device__ __forceinline__ uint32_t inverse_rgb(int y, int cb, int cr) {
  const int c = 298 * y;
  const int r = clamp255((c + 409 * cr - 52224) >> 8);
  const int g = clamp255((c - 100 * cb - 208 * cr + 39552) >> 8);
  const int b = clamp255((c + 516 * cb - 65920) >> 8);
  return (uint32_t)r | ((uint32_t)g << 8) | ((uint32_t)b << 16);
}

// ================================================================================================
// Generic gather kernel
// ================================================================================================
__device__ __forceinline__ uint32_t load_px(const uint8_t* __restrict__ frame, uint32_t row_bytes, int r, int c) {
  const uint8_t* q = frame + (size_t)r * row_bytes + (size_t)c * 3;
  return (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16);
}

// Chroma source of full-resolution pixel (r, c): ChromaSubsampler.scala:52-65 in closed form.
__device__ __forceinline__ void chroma_src_full(const KPlan& P, int r, int c, int& sr, int& sc) {
  if (P.vf == 2 && (r & 1)) {
    sr = r - 1;                 // nothing is sampled on an odd line: the latch still holds the last
    sc = P.last_sample_col;     // sample point of the line above
  } else {
    sr = r;
    sc = c - (c % P.hf);
  }
}

// Chroma source, in *output grid* coordinates, when the chroma stage runs on the downsampled stream
// but counts with the full W x H (ImageCompressorTop.scala:52-58).
__device__ __forceinline__ void chroma_src_case_b(const KPlan& P, int ro, int co, int& sro, int& sco) {
  const uint32_t m = (uint32_t)ro * (uint32_t)P.Wo + (uint32_t)co;
  const uint32_t col = m % (uint32_t)P.W;
  const uint32_t line = (m / (uint32_t)P.W) % (uint32_t)P.H;
  uint32_t src;
  if (P.vf == 2 && (line & 1)) src = (line - 1) * (uint32_t)P.W + (uint32_t)P.last_sample_col;
  else src = m - (col % (uint32_t)P.hf);
  sro = (int)(src / (uint32_t)P.Wo);
  sco = (int)(src % (uint32_t)P.Wo);
}

template <bool TRUNC>
__device__ __forceinline__ void ycc_of(uint32_t p, int& y, int& cb, int& cr) {
  y = (int)(fwd_y16(p) >> 8);
  cb = 255 - (int)(fwd_nc16<TRUNC>(p, kCoefNCb) >> 8);
  cr = 255 - (int)(fwd_nc16<TRUNC>(p, kCoefNCr) >> 8);
}

template <bool TRUNC>
__global__ void __launch_bounds__(256) csic_generic_kernel(const __grid_constant__ KPlan P) {
  const uint64_t total = (uint64_t)P.n_frames * (uint64_t)P.band_rows * (uint64_t)P.slots_per_row;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (uint64_t)gridDim.x * blockDim.x) {
    const int co = (int)(idx % (uint32_t)P.slots_per_row);
    const uint64_t t = idx / (uint32_t)P.slots_per_row;
    const int ro = P.row0 + (int)(t % (uint32_t)P.band_rows);
    const uint64_t k = t / (uint32_t)P.band_rows;
    uint8_t* orow = P.out + k * P.out_frame_bytes + (size_t)ro * P.out_row_bytes;
    if (co >= P.Wo) {   // BUNDLE row padding: zero slots
      if (P.slot_bytes == 1) orow[co] = 0;
      else if (P.slot_bytes == 2) reinterpret_cast<uint16_t*>(orow)[co] = 0;
      else reinterpret_cast<uint32_t*>(orow)[co] = 0;
      continue;
    }
    const uint8_t* frame = P.in + k * P.in_frame_bytes;
    const int f = P.f;
    int y, cb, cr;
    if (!P.average || f == 1) {
      const int yr = ro * f, yc = co * f;
      int sr, sc;
      if (!P.case_b) {
        chroma_src_full(P, yr, yc, sr, sc);
      } else {
        int sro, sco;
        chroma_src_case_b(P, ro, co, sro, sco);
        sr = sro * f;
        sc = sco * f;
      }
      y = (int)(fwd_y16(load_px(frame, P.in_row_bytes, yr, yc)) >> 8);
      const uint32_t pc = load_px(frame, P.in_row_bytes, sr, sc);
      cb = 255 - (int)(fwd_nc16<TRUNC>(pc, kCoefNCb) >> 8);
      cr = 255 - (int)(fwd_nc16<TRUNC>(pc, kCoefNCr) >> 8);
      y = (y >> P.sy) << P.sy;
      cb = (cb >> P.scb) << P.scb;
      cr = (cr >> P.scr) << P.scr;
    } else {
      // AVERAGE extension: mean over the f x f block of the stream entering the spatial stage.
      const int qy = P.quant_first ? P.sy : 0, qcb = P.quant_first ? P.scb : 0, qcr = P.quant_first ? P.scr : 0;
      int bro = ro, bco = co;   // block supplying chroma
      if (P.case_b) chroma_src_case_b(P, ro, co, bro, bco);
      int sy_ = 0, scb_ = 0, scr_ = 0;
      for (int dr = 0; dr < f; ++dr)
        for (int dc = 0; dc < f; ++dc) {
          const int r = ro * f + dr, c = co * f + dc;
          int yy = (int)(fwd_y16(load_px(frame, P.in_row_bytes, r, c)) >> 8);
          sy_ += (yy >> qy) << qy;
          int sr, sc;
          if (!P.case_b) {
            chroma_src_full(P, r, c, sr, sc);   // chroma stage ran at full resolution, before pooling
          } else {
            sr = bro * f + dr;                  // pooling first: own chroma of the source block
            sc = bco * f + dc;
          }
          const uint32_t pc = load_px(frame, P.in_row_bytes, sr, sc);
          const int b0 = 255 - (int)(fwd_nc16<TRUNC>(pc, kCoefNCb) >> 8);
          const int r0 = 255 - (int)(fwd_nc16<TRUNC>(pc, kCoefNCr) >> 8);
          scb_ += (b0 >> qcb) << qcb;
          scr_ += (r0 >> qcr) << qcr;
        }
      const int sh = 2 * (31 - __clz(f)), half = (f * f) >> 1;
      y = (sy_ + half) >> sh;
      cb = (scb_ + half) >> sh;
      cr = (scr_ + half) >> sh;
      if (!P.quant_first) {
        y = (y >> P.sy) << P.sy;
        cb = (cb >> P.scb) << P.scb;
        cr = (cr >> P.scr) << P.scr;
      }
    }
    if (P.kformat == KF_YCC888) {
      uint8_t* o = orow + (size_t)co * 3;
      o[0] = (uint8_t)y; o[1] = (uint8_t)cb; o[2] = (uint8_t)cr;
    } else if (P.kformat == KF_RGB888) {
      const uint32_t v = inverse_rgb(y, cb, cr);
      uint8_t* o = orow + (size_t)co * 3;
      o[0] = (uint8_t)v; o[1] = (uint8_t)(v >> 8); o[2] = (uint8_t)(v >> 16);
    } else {
      const uint32_t v = ((uint32_t)(y >> P.sy) << (P.cb_bits + P.cr_bits)) |
                         ((uint32_t)(cb >> P.scb) << P.cr_bits) | (uint32_t)(cr >> P.scr);
      if (P.slot_bytes == 1) orow[co] = (uint8_t)v;
      else if (P.slot_bytes == 2) reinterpret_cast<uint16_t*>(orow)[co] = (uint16_t)v;
      else reinterpret_cast<uint32_t*>(orow)[co] = v;
    }
  }
}

int launch_generic(const KPlan& k, void* stream) {
  const uint64_t total = (uint64_t)k.n_frames * (uint64_t)k.band_rows * (uint64_t)k.slots_per_row;
  if (total == 0) return (int)cudaSuccess;
  const uint64_t blocks = std::min<uint64_t>((total + 255) / 256, (uint64_t)148 * 64);
  if (k.trunc) csic_generic_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(k);
  else csic_generic_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(k);
  return (int)cudaGetLastError();
}

// ================================================================================================
// TMA-staged row kernel
// ================================================================================================
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
// global -> shared bulk copy performed by the TMA engine; completion counted in bytes on `bar`.
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar, uint64_t pol) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(pol)
      : "memory");
}
// shared -> global bulk copy (bulk async-group completion).
__device__ __forceinline__ void tma_store_1d(void* dst, uint32_t src, uint32_t bytes, uint64_t pol) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst), "r"(src),
               "r"(bytes), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}

__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}

// The four sampled pixels of a granule, each as a word whose low three bytes are R,G,B.
// A granule is 4 consecutive output pixels = 4 input pixels at a stride of F pixels (3F bytes);
// `a` is the shared-memory address of its first byte.
template <int F>
__device__ __forceinline__ void load_granule(uint32_t a, uint32_t (&p)[4]) {
  if (F == 1) {                                  // 12 bytes; word stride 3 across lanes: conflict free
    const uint32_t w0 = lds32(a), w1 = lds32(a + 4), w2 = lds32(a + 8);
    p[0] = w0;
    p[1] = __funnelshift_r(w0, w1, 24);
    p[2] = __funnelshift_r(w1, w2, 16);
    p[3] = w2 >> 8;
  } else if (F == 2) {                           // 24 bytes, 8-byte aligned; conflict free per half warp
    const uint2 u0 = lds64(a), u1 = lds64(a + 8), u2 = lds64(a + 16);
    p[0] = u0.x;                                 // bytes 0..2
    p[1] = __funnelshift_r(u0.y, u1.x, 16);      // bytes 6..8
    p[2] = u1.y;                                 // bytes 12..14
    p[3] = __funnelshift_r(u2.x, u2.y, 16);      // bytes 18..20
  } else {                                       // pixels sit on word boundaries
#pragma unroll
    for (int j = 0; j < 4; ++j) p[j] = lds32(a + j * 3u * F);
  }
}

// Per-stage tile descriptor: written by the producer thread before it arms the stage's mbarrier
// (release), read by every thread after the barrier's phase flips (acquire).
constexpr int kMaxTileRows = 16;
struct TileMeta {
  uint64_t out_base;                 // global address of the tile's first output byte
  uint32_t n_granules;               // rows * granules per row segment
  uint32_t any_held;                 // some row of the tile replays a held chroma pair
  uint32_t held_addr[kMaxTileRows];  // per row: 0, or shared address of the RGB pixel whose chroma the row replays
};

// Runtime constants of the inner loop, hoisted into registers once per kernel.
struct LoopConst {
  uint32_t qm0, qm1, qm2;            // YCC888: quantiser keep-masks over the three packed words
  uint32_t my, mcb, mcr;             // RGB888
  int shy, shb, shr, ly, lb;         // bundles
  uint32_t gran_per_row;
  uint32_t row0_of_thread, rem0_of_thread;   // threadIdx.x / gran_per_row, threadIdx.x % gran_per_row
  uint32_t drow, drem;                       // blockDim.x / gran_per_row, blockDim.x % gran_per_row
  bool trunc;
};

__device__ __forceinline__ uint32_t fwd_nc16_rt(uint32_t p, uint32_t coef, bool trunc) {
  int32_t x = max(dp4a_us(p, coef, 32639), 0);
  if (trunc) x -= (x >> 15) * 255;
  return (uint32_t)x;
}

// One tile: a flat loop over its granules.  Rows are packed back to back in the stage (and, for the
// staged formats, in the output buffer), so granule q lives at  base + q * granule_bytes  and the row
// index is only needed to look up a held chroma pair (HELD tiles).
//   HFE   chroma hold width inside a granule, in output pixels (1, 2 or 4)
//   HELD  the tile may contain rows that replay a held pair (odd 4:2:0 / 4:1:0 lines)
//   Q8    8/8/8 bits in a 32-bit slot: pure byte permutes
template <int F, int FMT, int HFE, bool HELD, bool Q8>
__device__ __forceinline__ void tile_loop(uint32_t in_s, uint32_t out_s, uint8_t* __restrict__ out_g,
                                          const TileMeta* __restrict__ meta, const LoopConst& C) {
  const uint32_t n = meta->n_granules;
  uint32_t row = C.row0_of_thread, rem = C.rem0_of_thread;   // row of granule q, tracked without a division
  for (uint32_t q = threadIdx.x; q < n; q += blockDim.x) {
    uint32_t p[4];
    load_granule<F>(in_s + q * (12u * F), p);
    uint32_t dy[4], xb[4], xr[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) dy[j] = fwd_y16(p[j]);
    uint32_t haddr = 0;
    if (HELD) {
      haddr = meta->held_addr[row];
      row += C.drow;
      rem += C.drem;
      if (rem >= C.gran_per_row) { rem -= C.gran_per_row; ++row; }
    }
    if (HELD && haddr != 0) {
      const uint32_t hp = lds8(haddr) | (lds8(haddr + 1) << 8) | (lds8(haddr + 2) << 16);
      const uint32_t hb = fwd_nc16_rt(hp, kCoefNCb, C.trunc), hr = fwd_nc16_rt(hp, kCoefNCr, C.trunc);
#pragma unroll
      for (int j = 0; j < 4; ++j) { xb[j] = hb; xr[j] = hr; }
    } else {
      // sample where j % HFE == 0, hold in between (ChromaSubsampler.scala:57-65)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j % HFE == 0) {
          xb[j] = fwd_nc16_rt(p[j], kCoefNCb, C.trunc);
          xr[j] = fwd_nc16_rt(p[j], kCoefNCr, C.trunc);
        } else {
          xb[j] = xb[j - 1];
          xr[j] = xr[j - 1];
        }
      }
    }

    if (FMT == KF_YCC888) {
      // byte 1 of dy is Y, byte 1 of xb/xr is ~Cb/~Cr: gather with PRMT, flip and quantise per word.
      uint32_t t, u;
      t = __byte_perm(dy[0], xb[0], 0x0051); u = __byte_perm(xr[0], dy[1], 0x0051);
      const uint32_t w0 = (__byte_perm(t, u, 0x5410) ^ 0x00FFFF00u) & C.qm0;
      t = __byte_perm(xb[1], xr[1], 0x0051); u = __byte_perm(dy[2], xb[2], 0x0051);
      const uint32_t w1 = (__byte_perm(t, u, 0x5410) ^ 0xFF00FFFFu) & C.qm1;
      t = __byte_perm(xr[2], dy[3], 0x0051); u = __byte_perm(xb[3], xr[3], 0x0051);
      const uint32_t w2 = (__byte_perm(t, u, 0x5410) ^ 0xFFFF00FFu) & C.qm2;
      const uint32_t a = out_s + q * 12u;
      sts32(a, w0); sts32(a + 4, w1); sts32(a + 8, w2);
    } else if (FMT == KF_RGB888) {
      uint32_t v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int y = (int)((dy[j] >> 8) & C.my);
        const int cb = (int)((255u - (xb[j] >> 8)) & C.mcb);
        const int cr = (int)((255u - (xr[j] >> 8)) & C.mcr);
        v[j] = inverse_rgb(y, cb, cr);
      }
      const uint32_t a = out_s + q * 12u;
      sts32(a, v[0] | (v[1] << 24));
      sts32(a + 4, (v[1] >> 8) | (v[2] << 16));
      sts32(a + 8, (v[2] >> 16) | (v[3] << 8));
    } else {
      // bundle slots go straight to global memory: one coalesced 4/8/16-byte store per granule
      uint32_t v[4];
      if (Q8) {
#pragma unroll
        for (int j = 0; j < 4; ++j)   // (Cr, Cb, Y, 0): dy < 65536 so its byte 3 is the zero pad
          v[j] = __byte_perm(__byte_perm(xr[j], xb[j], 0x0051), dy[j], 0x7510) ^ 0x0000FFFFu;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          v[j] = ((dy[j] >> C.shy) << C.ly) | (((xb[j] ^ 0xFFFFu) >> C.shb) << C.lb) | ((xr[j] ^ 0xFFFFu) >> C.shr);
      }
      if (FMT == KF_SLOT32) __stcs(reinterpret_cast<uint4*>(out_g) + q, make_uint4(v[0], v[1], v[2], v[3]));
      else if (FMT == KF_SLOT16) __stcs(reinterpret_cast<uint2*>(out_g) + q, make_uint2(v[0] | (v[1] << 16), v[2] | (v[3] << 16)));
      else __stcs(reinterpret_cast<uint32_t*>(out_g) + q, v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24));
    }
  }
}

template <int F, int FMT, bool Q8>
__global__ void __launch_bounds__(256) csic_rows_kernel(const __grid_constant__ KPlan P) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr bool kStaged = (FMT == KF_YCC888 || FMT == KF_RGB888);   // output leaves through smem + TMA store
  const uint32_t tid = threadIdx.x;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t S = (uint32_t)P.stages;
  const uint32_t n_my = (P.n_tiles > blockIdx.x) ? (P.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint64_t pol = policy_evict_first();   // every byte is touched exactly once: do not keep it in L2

  // -- producer (thread 0): tile i of this CTA -> stage i % S ------------------------------------
  auto issue_load = [&](uint32_t i) {
    const uint32_t tile = blockIdx.x + i * gridDim.x;
    const uint32_t t2 = tile / (uint32_t)P.nsplit;
    const uint32_t seg = tile - t2 * (uint32_t)P.nsplit;
    const uint32_t k = t2 / P.tiles_per_band;
    const uint32_t tb = t2 - k * P.tiles_per_band;
    const uint32_t ro0 = (uint32_t)P.row0 + tb * (uint32_t)P.tile_rows;
    const uint32_t nrows = min((uint32_t)P.tile_rows, (uint32_t)(P.row0 + P.band_rows) - ro0);
    const uint32_t s = i % S;
    const uint8_t* frame = P.in + (uint64_t)k * P.in_frame_bytes;
    const uint32_t bar = sbase + P.bar_off + s * 8u;
    const uint32_t dst = sbase + s * P.stage_stride;
    const uint32_t aux = dst + (uint32_t)P.tile_rows * P.tile_in_bytes;    // 32-byte window per row
    TileMeta* m = reinterpret_cast<TileMeta*>(smem + P.meta_off) + s;

    // Which rows replay a held chroma pair, and from where?  (KPlan::hfe covers the in-row hold.)
    uint32_t n_aux = 0, any = 0;
    const uint8_t* aux_src[kMaxTileRows];
    if (P.vf == 2) {
      for (uint32_t j = 0; j < nrows; ++j) {
        const uint32_t ro = ro0 + j;
        uint32_t h = 0;
        const uint8_t* hp = nullptr;
        if (!P.case_b) {
          if (F == 1 && (ro & 1)) {          // odd line at full resolution: last sample point of the line above
            if (j > 0 && P.nsplit == 1) h = dst + (j - 1) * P.tile_in_bytes + (uint32_t)P.last_sample_col * 3u;   // in this tile
            else hp = frame + (uint64_t)(ro - 1) * P.in_row_bytes + (uint32_t)P.last_sample_col * 3u;
          }
        } else {
          const uint32_t line = ro / F;      // W == F * Wo: one counter line spans F output rows
          if (line & 1) {
            const uint32_t srow = (line - 1) * F + (uint32_t)P.last_sample_col / (uint32_t)P.Wo;
            const uint32_t scol = (uint32_t)P.last_sample_col % (uint32_t)P.Wo;
            hp = frame + (uint64_t)(srow * F) * P.in_row_bytes + (uint64_t)scol * (3u * F);
          }
        }
        if (hp) {
          const uint64_t a = reinterpret_cast<uint64_t>(hp);
          h = aux + j * 32u + (uint32_t)(a & 15u);
          aux_src[j] = reinterpret_cast<const uint8_t*>(a & ~(uint64_t)15);
          ++n_aux;
        } else {
          aux_src[j] = nullptr;
        }
        m->held_addr[j] = h;
        any |= h;
      }
    }
    m->any_held = any;
    m->n_granules = nrows * ((uint32_t)P.tile_px >> 2);
    m->out_base = reinterpret_cast<uint64_t>(P.out) + (uint64_t)k * P.out_frame_bytes + (uint64_t)ro0 * P.out_row_bytes +
                  (uint64_t)seg * P.tile_out_bytes;
    mbar_expect_tx(bar, nrows * P.tile_in_bytes + n_aux * 32u);
    const uint8_t* src = frame + (uint64_t)(ro0 * F) * P.in_row_bytes + (uint64_t)seg * P.tile_in_bytes;
    if (F == 1 && P.nsplit == 1) {           // consecutive rows are contiguous in memory: one bulk copy
      tma_load_1d(dst, src, nrows * P.tile_in_bytes, bar, pol);
    } else {
      for (uint32_t j = 0; j < nrows; ++j)
        tma_load_1d(dst + j * P.tile_in_bytes, src + (uint64_t)j * F * P.in_row_bytes, P.tile_in_bytes, bar, pol);
    }
    if (n_aux) {
      for (uint32_t j = 0; j < nrows; ++j)
        if (aux_src[j]) tma_load_1d(aux + j * 32u, aux_src[j], 32u, bar, pol);
    }
  };

  if (tid == 0) {
    for (uint32_t s = 0; s < S; ++s) mbar_init(sbase + P.bar_off + s * 8u, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    for (uint32_t i = 0; i + 1 < S && i < n_my; ++i) issue_load(i);
  }

  LoopConst C;
  {
    const uint32_t my = P.qmask & 0xFFu, mcb = (P.qmask >> 8) & 0xFFu, mcr = (P.qmask >> 16) & 0xFFu;
    C.qm0 = my | (mcb << 8) | (mcr << 16) | (my << 24);
    C.qm1 = mcb | (mcr << 8) | (my << 16) | (mcb << 24);
    C.qm2 = mcr | (my << 8) | (mcb << 16) | (mcr << 24);
    C.my = my; C.mcb = mcb; C.mcr = mcr;
    C.shy = 8 + P.sy; C.shb = 8 + P.scb; C.shr = 8 + P.scr;
    C.ly = P.cb_bits + P.cr_bits; C.lb = P.cr_bits;
    C.gran_per_row = (uint32_t)P.tile_px >> 2;
    C.row0_of_thread = tid / C.gran_per_row;
    C.rem0_of_thread = tid % C.gran_per_row;
    C.drow = blockDim.x / C.gran_per_row;
    C.drem = blockDim.x % C.gran_per_row;
    C.trunc = P.trunc != 0;
  }
  const int hfe = P.hfe;

  for (uint32_t i = 0; i < n_my; ++i) {
    const uint32_t s = i % S;
    // Refill the stage that was consumed in iteration i-1 (everyone passed that iteration's barrier).
    if (tid == 0 && i + S - 1 < n_my) issue_load(i + S - 1);
    mbar_wait(sbase + P.bar_off + s * 8u, (i / S) & 1u);

    const uint32_t in_s = sbase + s * P.stage_stride;
    const uint32_t out_s = sbase + P.out_buf_off + (i & 1u) * P.out_buf_stride;
    const TileMeta* m = reinterpret_cast<const TileMeta*>(smem + P.meta_off) + s;
    uint8_t* out_g = reinterpret_cast<uint8_t*>(m->out_base);
    if (m->any_held) {
      if (hfe == 1) tile_loop<F, FMT, 1, true, Q8>(in_s, out_s, out_g, m, C);
      else if (hfe == 2) tile_loop<F, FMT, 2, true, Q8>(in_s, out_s, out_g, m, C);
      else tile_loop<F, FMT, 4, true, Q8>(in_s, out_s, out_g, m, C);
    } else {
      if (hfe == 1) tile_loop<F, FMT, 1, false, Q8>(in_s, out_s, out_g, m, C);
      else if (hfe == 2) tile_loop<F, FMT, 2, false, Q8>(in_s, out_s, out_g, m, C);
      else tile_loop<F, FMT, 4, false, Q8>(in_s, out_s, out_g, m, C);
    }

    if (kStaged) {
      // Hand the tile to the TMA engine.  Generic-proxy writes must be fenced before the async proxy
      // reads them; the staging buffer used two tiles ago must have been read out before it is reused.
      const uint32_t bytes = m->n_granules * 12u;
      fence_proxy_async_smem();
      if (tid == 0) tma_store_wait_read0();
      __syncthreads();
      if (tid == 0) {
        tma_store_1d(out_g, out_s, bytes, pol);
        tma_store_commit();
      }
    } else {
      __syncthreads();    // everyone is done reading stage s (and its meta) before it is refilled
    }
  }
  if (kStaged && tid == 0) tma_store_wait_all();
}

bool plan_rows_kernel(KPlan& k, int sm_count, size_t max_smem_optin, int force_stages, uint32_t force_tile_bytes) {
  if (k.average && k.f > 1) return false;                       // AVERAGE extension: generic kernel
  if (k.W % k.f != 0) return false;                             // a counter line must be whole output rows
  if (k.Wo % 16 != 0) return false;                             // 16-byte TMA granularity on the output rows
  if (k.in_row_bytes % 16 != 0) return false;                   // ... and on the input rows (W % 16 == 0)
  if ((reinterpret_cast<uintptr_t>(k.in) | reinterpret_cast<uintptr_t>(k.out)) & 15u) return false;
  if ((k.in_frame_bytes | k.out_frame_bytes | k.out_row_bytes) & 15u) return false;
  if (k.slots_per_row != k.Wo) return false;                    // BUNDLE rows with padding
  if (k.case_b && k.Wo < 4) return false;
  if (k.band_rows <= 0 || k.n_frames == 0) return false;

  // hold width inside a granule, in output pixels
  if (!k.case_b) k.hfe = std::max(1, k.hf / k.f);
  else k.hfe = k.hf;

  const bool staged = k.kformat <= KF_RGB888;
  const uint32_t opx = staged ? 3u : (uint32_t)k.slot_bytes;
  // Tile budget: input bytes of one tile.  Staged formats also hold two output buffers per CTA.
  const uint32_t tile_budget = force_tile_bytes ? force_tile_bytes : (staged ? 12u * 1024u : 24u * 1024u);
  const uint32_t row_in = (uint32_t)k.Wo * 3u * (uint32_t)k.f;  // == in_row_bytes
  int nsplit = 0;
  for (int n = (int)((row_in + tile_budget - 1) / tile_budget); n <= 64; ++n) {
    if (k.Wo % (16 * n) == 0) { nsplit = n; break; }
  }
  if (nsplit == 0) return false;
  k.nsplit = nsplit;
  k.tile_px = k.Wo / nsplit;
  k.tile_in_bytes = (uint32_t)k.tile_px * 3u * (uint32_t)k.f;    // one row segment
  k.tile_out_bytes = (uint32_t)k.tile_px * opx;
  // Rows per tile: whole rows only (so the tile's output is contiguous), as many as fit the budget,
  // but keep at least ~4 tiles per SM so small batches still spread over the chip.
  int rows = 1;
  if (nsplit == 1) {
    rows = (int)std::min<uint32_t>((uint32_t)kMaxTileRows, std::max<uint32_t>(1u, tile_budget / k.tile_in_bytes));
    rows = std::min(rows, k.band_rows);
    auto tiles_for = [&](int r) { return (uint64_t)k.n_frames * (uint64_t)((k.band_rows + r - 1) / r); };
    while (rows > 1 && tiles_for(rows) < (uint64_t)sm_count * 4u) rows = (rows + 1) / 2;
  }
  k.tile_rows = rows;
  k.tiles_per_band = (uint32_t)((k.band_rows + rows - 1) / rows);
  const uint64_t n_tiles = (uint64_t)k.n_frames * (uint64_t)k.tiles_per_band * (uint64_t)nsplit;
  if (n_tiles >= (1ull << 31)) return false;
  k.n_tiles = (uint32_t)n_tiles;

  auto up128 = [](uint32_t v) { return (v + 127u) & ~127u; };
  k.stage_stride = up128((uint32_t)rows * (k.tile_in_bytes + 32u));
  k.out_buf_stride = staged ? up128((uint32_t)rows * k.tile_out_bytes) : 0u;
  int stages = force_stages >= 2 ? force_stages : 3;
  auto need = [&](int s) {
    return (uint32_t)s * k.stage_stride + 2u * k.out_buf_stride + (uint32_t)s * (uint32_t)sizeof(TileMeta) + (uint32_t)s * 8u + 256u;
  };
  while (force_stages < 2 && !force_tile_bytes && stages > 2 && need(stages) > 76u * 1024u) --stages;
  if (need(stages) > max_smem_optin) return false;
  k.stages = stages;
  k.out_buf_off = (uint32_t)stages * k.stage_stride;
  k.meta_off = up128(k.out_buf_off + 2u * k.out_buf_stride);
  k.bar_off = up128(k.meta_off + (uint32_t)stages * (uint32_t)sizeof(TileMeta));
  k.smem_bytes = k.bar_off + (uint32_t)stages * 8u;
  return true;
}

template <int F, int FMT>
static int launch_rows_t(const KPlan& k, unsigned grid, cudaStream_t st) {
  const bool q8 = FMT == KF_SLOT32 && k.sy == 0 && k.scb == 0 && k.scr == 0;
  if (FMT == KF_SLOT32 && q8) csic_rows_kernel<F, FMT, (FMT == KF_SLOT32)><<<grid, 256, k.smem_bytes, st>>>(k);
  else csic_rows_kernel<F, FMT, false><<<grid, 256, k.smem_bytes, st>>>(k);
  return (int)cudaGetLastError();
}

template <int F>
static int launch_rows_f(const KPlan& k, unsigned grid, cudaStream_t st) {
  switch (k.kformat) {
    case KF_YCC888: return launch_rows_t<F, KF_YCC888>(k, grid, st);
    case KF_RGB888: return launch_rows_t<F, KF_RGB888>(k, grid, st);
    case KF_SLOT8: return launch_rows_t<F, KF_SLOT8>(k, grid, st);
    case KF_SLOT16: return launch_rows_t<F, KF_SLOT16>(k, grid, st);
    default: return launch_rows_t<F, KF_SLOT32>(k, grid, st);
  }
}

template <int F, int FMT, bool Q8>
static cudaError_t set_attr_one(size_t bytes) {
  return cudaFuncSetAttribute(csic_rows_kernel<F, FMT, Q8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}
template <int F>
static cudaError_t set_attr_f(size_t b) {
  cudaError_t e;
  if ((e = set_attr_one<F, KF_YCC888, false>(b)) != cudaSuccess) return e;
  if ((e = set_attr_one<F, KF_RGB888, false>(b)) != cudaSuccess) return e;
  if ((e = set_attr_one<F, KF_SLOT8, false>(b)) != cudaSuccess) return e;
  if ((e = set_attr_one<F, KF_SLOT16, false>(b)) != cudaSuccess) return e;
  if ((e = set_attr_one<F, KF_SLOT32, false>(b)) != cudaSuccess) return e;
  if ((e = set_attr_one<F, KF_SLOT32, true>(b)) != cudaSuccess) return e;
  return cudaSuccess;
}

int rows_kernel_set_attributes(size_t max_smem_optin) {
  cudaError_t e;
  if ((e = set_attr_f<1>(max_smem_optin)) != cudaSuccess) return (int)e;
  if ((e = set_attr_f<2>(max_smem_optin)) != cudaSuccess) return (int)e;
  if ((e = set_attr_f<4>(max_smem_optin)) != cudaSuccess) return (int)e;
  if ((e = set_attr_f<8>(max_smem_optin)) != cudaSuccess) return (int)e;
  return (int)cudaSuccess;
}

int launch_rows(const KPlan& k, int sm_count, int force_ctas_per_sm, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  // Persistent grid: a whole number of CTAs per SM, as many as the shared-memory footprint allows.
  uint32_t per_sm = std::max<uint32_t>(1u, std::min<uint32_t>(8u, (uint32_t)(224u * 1024u / (k.smem_bytes + 1024u))));
  if (force_ctas_per_sm > 0) per_sm = std::min<uint32_t>(per_sm, (uint32_t)force_ctas_per_sm);
  const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)k.n_tiles, (uint64_t)sm_count * per_sm);
  switch (k.f) {
    case 1: return launch_rows_f<1>(k, grid, st);
    case 2: return launch_rows_f<2>(k, grid, st);
    case 4: return launch_rows_f<4>(k, grid, st);
    default: return launch_rows_f<8>(k, grid, st);
  }
}

}  // namespace csic
