// Planning and dispatch of the pixel pipeline's kernels, written for sm_100a (B200), and the two kernels that are not
// TMA-staged.  The staged kernels live in their own files:
//   csic_rows_kernel.cu    the hot path: packed RGB24 row segments through a shared-memory ring filled by the TMA engine
//                          (cp.async.bulk + mbarrier complete_tx), dp4a colour matrix, chroma hold inside the granule (or
//                          from one held pixel per row), quantise, pack, TMA bulk store.  Planned here (plan_rows_kernel).
//   csic_pool_kernel.cu    the AVERAGE extension, same machinery (plan_pool_kernel).
//   csic_flex_kernel.cu    DECIMATE for any width / pitch / alignment (hull fetch, staged span stores).
//   csic_decode_kernel.cu  the PLANAR decoder as flat tiles, any width.
// In this file:
//   csic_generic_kernel            a warp per output row, closed-form gather; any legal parameter set (odd sizes,
//                                  unaligned pointers, AVERAGE on dense odd widths, spatial-before-chroma shapes the
//                                  staged kernels exclude).  Not a fallback to the CPU: the independent second
//                                  implementation every other kernel is cross-checked against.
//   csic_expand_planar_any_kernel  LDG decoder for what csic_decode_kernel declines.
//
// Semantics (bit-exact with the reference; citations relative to its root, src/main/scala/jpeg/):
//   forward   RGB2YCbCr.scala:33-35,50-65,74-76 (FLOOR)   RGB2YCbCr.scala:95-121 (TRUNC)
//   chroma    ChromaSubsampler.scala:26-27,34-38,52-65      spatial  SpatialDownsampler.scala:17-55
//   quant     ColorQuantizer.scala:29-31,42-44              inverse  RGB2YCbCr.scala:123-132
//   order / misaligned counters   ImageCompressorTop.scala:43-58,83-114
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>

#include "csic_internal.h"
#include "csic_device_math.cuh"
#include "csic_tma.cuh"

namespace csic {

// ================================================================================================
// Generic gather kernel: any legal parameter set, any width, pitch and alignment
// ================================================================================================
// What runs when no staged kernel is eligible: spatial-before-chroma shapes whose counter lines are not whole output
// rows (e.g. W = 13, f = 2), the AVERAGE extension on shapes that break the pooling kernel's 16-byte rules -- and
// every parameter set when a test forces it (CSIC_OPT_KERNEL_FAMILY = 1), which makes it the independent second
// implementation the other kernels are cross-checked against.
//
// One WARP per output row (grid-stride) -- or, when a row has fewer than 32 granules, as many whole rows as fit its
// 32 lanes (a 96x96 frame pooled 8x8 has three granules per row: one row per CTA left 125 of 128 threads idle, 0.02 of
// the copy peak).  The frame / row split and, for case B, the split of the row into counter lines are the only
// divisions and happen once per row.  A thread computes one granule of four output slots:
//   load    every pixel is three bytes at an arbitrary address = two aligned LDG.32 and a funnel shift through L1 (the
//           second word never lies beyond the last word that holds a byte of the input); a thread's loads are all
//           independent, so a warp keeps 8 .. 200 of them in flight;
//   chroma  closed-form source maps of ChromaSubsampler.scala:52-65; case A resolves inside the granule (hold width
//           hfe divides 4), held lines replay one pixel per row; case B (ImageCompressorTop.scala:52-58) maps every
//           element through the full-size counters;
//   store   a warp's 32 granules are consecutive output bytes: staged in the warp's shared-memory slot at the output's
//           own offset modulo 16 and written as 16-byte st.global.cs aligned on the GLOBAL address (span_store); narrow
//           rows (several per warp) leave from registers with the widest stores their address allows (direct_store).
// Round 1's version -- one thread per slot, three divisions, byte loads and byte stores -- ran at 0.13 - 0.30 of the
// copy peak.

// A stored input row as the gather kernel addresses it: word-aligned base, byte phase of its first pixel, and the last
// word index a load may touch (the word holding the last byte this launch may read).  Set up once per row with 64-bit
// arithmetic; every pixel is then 32-bit offset arithmetic (round 2's first version did the 64-bit pointer math and
// the clamp per pixel: 382 warp-instructions per DECIMATE granule, now ~150).
struct GRow {
  const uint32_t* base;
  uint32_t a0, maxw;
};
__device__ __forceinline__ GRow grow_at(const uint8_t* p, const uint8_t* last_word) {
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  GRow r;
  r.base = reinterpret_cast<const uint32_t*>(a & ~(uintptr_t)3);
  r.a0 = (uint32_t)a & 3u;
  const uintptr_t lw = reinterpret_cast<uintptr_t>(last_word), ba = reinterpret_cast<uintptr_t>(r.base);
  r.maxw = lw >= ba ? (uint32_t)min((lw - ba) >> 2, (uintptr_t)0x3FFFFFFFu) : 0u;
  return r;
}
// three colour bytes (low three bytes of the result) at byte offset `off` of the row: two aligned LDG.32 through L1
// and a funnel shift; the second word is never fetched beyond GRow::maxw
__device__ __forceinline__ uint32_t grow_px(const GRow& r, uint32_t off) {
  const uint32_t o = r.a0 + off, wi = o >> 2;
  return __funnelshift_r(__ldg(r.base + wi), __ldg(r.base + min(wi + 1u, r.maxw)), (o & 3u) * 8u);
}

// Chroma source, in *output grid* coordinates, when the chroma stage runs on the downsampled stream
// but counts with the full W x H (ImageCompressorTop.scala:52-58).
__device__ __forceinline__ void chroma_src_case_b(const KPlan& P, int ro, int co, int& sro, int& sco) {
  const uint32_t m = (uint32_t)ro * (uint32_t)P.Wo + (uint32_t)co;
  const uint32_t line = m / (uint32_t)P.W;           // m < Wo * Ho <= W * H: the `% H` of the counters never wraps
  const uint32_t col = m - line * (uint32_t)P.W;
  uint32_t src;
  if (P.vf == 2 && (line & 1)) src = (line - 1) * (uint32_t)P.W + (uint32_t)P.last_sample_col;
  else src = m - (col % (uint32_t)P.hf);
  sro = (int)(src / (uint32_t)P.Wo);
  sco = (int)(src - (uint32_t)sro * (uint32_t)P.Wo);
}

constexpr uint32_t kGatherSlot = 32u * 16u + 32u;   // a warp's 32 granules of up to 16 bytes + alignment offset + read-ahead slack

// AF: 0 = DECIMATE (or f == 1); 2 / 4 / 8 = the AVERAGE extension with that factor (block loops unrolled)
template <bool TRUNC, int FMT, int AF>
__global__ void __launch_bounds__(128) csic_generic_kernel(const __grid_constant__ KPlan P) {
  __shared__ __align__(16) uint8_t stage[4][kGatherSlot];
  constexpr uint32_t kG = (FMT == KF_YCC888 || FMT == KF_RGB888) ? 12u : (FMT == KF_SLOT32 ? 16u : (FMT == KF_SLOT16 ? 8u : 4u));
  constexpr bool avg = AF > 1;
  constexpr uint32_t NR = avg ? (uint32_t)AF : 1u;                          // input rows per output row
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const uint32_t Wo = (uint32_t)P.Wo, spr = (uint32_t)P.slots_per_row, gpr = (spr + 3u) >> 2;
  const uint32_t n_rows = P.n_frames * (uint32_t)P.band_rows;
  const uint32_t f = (uint32_t)P.f, ipb = (uint32_t)P.in_px_bytes, pxb = f * ipb;
  const uint32_t hfe = P.case_b ? 1u : (uint32_t)max(1, P.hf / P.f);         // case A: hold width inside a granule, in output pixels
  const uint32_t hfm = (uint32_t)P.hf - 1u;                                  // hf is 1, 2 or 4
  const uint32_t row_bytes = (FMT == KF_YCC888 || FMT == KF_RGB888) ? Wo * 3u : (FMT == KF_PLANAR ? Wo : spr * (uint32_t)P.slot_bytes);
  // last aligned word that still holds a byte this launch may read: the end of the band's last input row (the host
  // band path hands over buffers that hold only the band's rows)
  const uint32_t last_out_row = (uint32_t)(P.row0 + P.band_rows - 1);
  const uint32_t last_in_row = P.compact ? last_out_row : (avg ? (last_out_row + 1u) * f - 1u : last_out_row * f);
  const uint8_t* last_word = reinterpret_cast<const uint8_t*>(
      (reinterpret_cast<uintptr_t>(P.in) + (uint64_t)(P.n_frames - 1u) * P.in_frame_bytes + (uint64_t)last_in_row * P.in_row_bytes +
       (uint64_t)(uint32_t)P.W * ipb - 1u) & ~(uintptr_t)3);
  const uint32_t my = P.qmask & 0xFFu, mcb = (P.qmask >> 8) & 0xFFu, mcr = (P.qmask >> 16) & 0xFFu;
  const uint32_t sbase = smem_u32(stage[warp]);
  const int fsh = 31 - __clz(P.f);                                          // log2(f)

  // lanes -> (row of the warp's group, granule): wide rows take the whole warp, narrow ones share it
  const bool narrow = gpr < 32u;
  const uint32_t rpw = narrow ? 32u / gpr : 1u;                              // rows per warp and trip
  const uint32_t sub = narrow ? lane / gpr : 0u, gl = narrow ? lane - sub * gpr : lane;
  const uint32_t n_groups = (n_rows + rpw - 1u) / rpw, nwarps = blockDim.x >> 5;
  for (uint32_t grp = blockIdx.x * nwarps + warp; grp < n_groups; grp += gridDim.x * nwarps) {
    const uint32_t Rw = grp * rpw + sub;
    const bool row_ok = sub < rpw && Rw < n_rows;                            // lanes beyond the warp's last row idle (they still
    const uint32_t R = min(Rw, n_rows - 1u);                                 // compute, on a valid row, and store nothing)
    const uint32_t k = R / (uint32_t)P.band_rows, ro = (uint32_t)P.row0 + (R - k * (uint32_t)P.band_rows);
    const uint8_t* frame = P.in + (uint64_t)k * P.in_frame_bytes;
    uint8_t* fout = P.out + (uint64_t)k * P.out_frame_bytes;
    uint8_t* orow = fout + (uint64_t)ro * P.out_row_bytes;
    // stored row of a full-resolution row (KPlan::compact: only every f-th row is stored, DECIMATE reads no other)
    auto in_row = [&](uint32_t r) { return frame + (uint64_t)(P.compact ? (r >> fsh) : r) * P.in_row_bytes; };
    // the input rows of this output row: its own (DECIMATE) or the AF rows of its blocks (AVERAGE)
    GRow rows[NR];
#pragma unroll
    for (uint32_t dr = 0; dr < NR; ++dr) rows[dr] = grow_at(in_row(ro * f + dr), last_word);
    // Held chroma, one pixel per held line: DECIMATE case A on an odd full-resolution line of 4:2:0 / 4:1:0 (f == 1
    // only) replays the last sample point of the line above (ChromaSubsampler.scala:62-65); AVERAGE with the chroma
    // stage first does so for every odd line dr of the block rows.
    uint32_t hxb[NR / 2u + 1u], hxr[NR / 2u + 1u];
    const bool vheld = !P.case_b && P.vf == 2;
    const bool held_row = !avg && vheld && ((ro * f) & 1u);
#pragma unroll
    for (uint32_t h = 0; h < NR / 2u + 1u; ++h) { hxb[h] = 0u; hxr[h] = 0u; }
    if (held_row) {
      const GRow hr = grow_at(in_row(ro * f - 1u), last_word);
      const uint32_t hp = grow_px(hr, (uint32_t)P.last_sample_col * ipb);
      hxb[0] = fwd_nc16<TRUNC>(hp, P.coef_ncb); hxr[0] = fwd_nc16<TRUNC>(hp, P.coef_ncr);
    }
    if (avg && vheld) {
#pragma unroll
      for (int h = 0; h < (int)(NR / 2u); ++h) {                             // odd line 2h + 1 replays (2h, lastSampleCol)
        const uint32_t hp = grow_px(rows[2u * h], (uint32_t)P.last_sample_col * ipb);
        hxb[h] = fwd_nc16<TRUNC>(hp, P.coef_ncb); hxr[h] = fwd_nc16<TRUNC>(hp, P.coef_ncr);
      }
    }
    // Case B (spatial before chroma, DECIMATE): the chroma stage walks the DECIMATED stream with the full-size counters
    // (ImageCompressorTop.scala:52-58), so a counter line is W stream elements = W / Wo output rows, and an output row
    // lies in at most two lines (W >= Wo).  Per row: the column where the second line starts, and for each of the two
    // lines either its held pair (odd line, vf == 2: element (line-1) * W + lastSampleCol) or the phase of its hold
    // groups.  Per pixel nothing is divided any more.
    uint32_t cb_split = 0xFFFFFFFFu, cb_ph[1] = {0u}, cb_hxb[2] = {0u, 0u}, cb_hxr[2] = {0u, 0u};
    bool cb_held[2] = {false, false};
    GRow prev_row = rows[0];
    if (!avg && P.case_b) {
      const uint32_t m0 = ro * Wo, line0 = m0 / (uint32_t)P.W, col0 = m0 - line0 * (uint32_t)P.W;
      cb_split = (uint32_t)P.W - col0;                                       // first column of the row that lies in line0 + 1
#pragma unroll
      for (int l = 0; l < 2; ++l) {
        const uint32_t line = line0 + (uint32_t)l;
        if (l == 0) cb_ph[0] = col0 & hfm;                                   // (column in its line) mod hf of the row's first pixel; 0 for the second line
        if (P.vf == 2 && (line & 1u) && (l == 0 || cb_split < Wo)) {
          const uint32_t src = (line - 1u) * (uint32_t)P.W + (uint32_t)P.last_sample_col, sro = src / Wo, sco = src - sro * Wo;
          const GRow hr = grow_at(in_row(sro * f), last_word);
          const uint32_t hp = grow_px(hr, sco * pxb);
          cb_hxb[l] = fwd_nc16<TRUNC>(hp, P.coef_ncb); cb_hxr[l] = fwd_nc16<TRUNC>(hp, P.coef_ncr);
          cb_held[l] = true;
        }
      }
      if (ro > 0u) prev_row = grow_at(in_row((ro - 1u) * f), last_word);    // a hold group may start at the end of the row above
    }
    for (uint32_t g0 = 0; g0 < gpr; g0 += 32u) {                             // warp uniform; one trip when the rows are narrow
      const uint32_t g = narrow ? gl : g0 + lane, c0 = 4u * g;
      uint8_t* og = orow + (size_t)g0 * kG;                                  // first output byte of this warp's group
      const uint32_t st = sbase + ((uint32_t)reinterpret_cast<uintptr_t>(og) & 12u);
      if (g < gpr) {
        uint32_t y[4], cb[4], cr[4];                                         // final channel values (quantised)
        if (!avg) {
          uint32_t p[4], xb[4], xr[4];
          if (pxb == 3u && c0 + 3u < Wo) {                                   // f == 1, RGB24: twelve consecutive bytes
            const uint32_t o = rows[0].a0 + c0 * 3u, wi = o >> 2, sh = (o & 3u) * 8u;
            const uint32_t w0 = __ldg(rows[0].base + wi), w1 = __ldg(rows[0].base + wi + 1u), w2 = __ldg(rows[0].base + wi + 2u);
            const uint32_t w3 = __ldg(rows[0].base + min(wi + 3u, rows[0].maxw));
            const uint32_t v0 = __funnelshift_r(w0, w1, sh), v1 = __funnelshift_r(w1, w2, sh), v2 = __funnelshift_r(w2, w3, sh);
            p[0] = v0; p[1] = __funnelshift_r(v0, v1, 24); p[2] = __funnelshift_r(v1, v2, 16); p[3] = v2 >> 8;
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) p[j] = grow_px(rows[0], min(c0 + j, Wo - 1u) * pxb);
          }
          if (held_row) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { xb[j] = hxb[0]; xr[j] = hxr[0]; }
          } else if (!P.case_b) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {                                    // sample where j % hfe == 0, hold in between
              if (j == 0 || ((uint32_t)j & (hfe - 1u)) == 0u) { xb[j] = fwd_nc16<TRUNC>(p[j], P.coef_ncb); xr[j] = fwd_nc16<TRUNC>(p[j], P.coef_ncr); }
              else { xb[j] = xb[j - 1]; xr[j] = xr[j - 1]; }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const uint32_t co = min(c0 + j, Wo - 1u);
              const bool l = co >= cb_split;                                 // which of the row's two counter lines
              if (l ? cb_held[1] : cb_held[0]) {
                xb[j] = l ? cb_hxb[1] : cb_hxb[0]; xr[j] = l ? cb_hxr[1] : cb_hxr[0];
              } else {
                // sampled line: the pixel replays the first element of its hold group, (col - col % hf); the group may
                // begin in the row above (never further back: hf - 1 <= 3 elements; rows narrower than that divide)
                const uint32_t back = (l ? co - cb_split : co + cb_ph[0]) & hfm;
                uint32_t pc;
                if (back == 0u) pc = p[j];
                else if (back <= co) pc = grow_px(rows[0], (co - back) * pxb);
                else if (Wo + co >= back) pc = grow_px(prev_row, (Wo + co - back) * pxb);
                else {
                  int sro, sco;
                  chroma_src_case_b(P, (int)ro, (int)co, sro, sco);
                  pc = grow_px(grow_at(in_row((uint32_t)sro * f), last_word), (uint32_t)sco * pxb);
                }
                xb[j] = fwd_nc16<TRUNC>(pc, P.coef_ncb); xr[j] = fwd_nc16<TRUNC>(pc, P.coef_ncr);
              }
            }
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            y[j] = (fwd_y16(p[j], P.coef_y) >> 8) & my;
            cb[j] = (~(xb[j] >> 8)) & mcb;
            cr[j] = (~(xr[j] >> 8)) & mcr;
          }
        } else {
          // AVERAGE extension: mean over the AF x AF block of the stream entering the spatial stage, round half up;
          // the quantiser before or after the mean as op[] says
          const uint32_t qy = P.quant_first ? my : 0xFFu, qb = P.quant_first ? mcb : 0xFFu, qr = P.quant_first ? mcr : 0xFFu;
          constexpr uint32_t half = (uint32_t)(AF * AF) >> 1, sh = AF == 2 ? 2u : (AF == 4 ? 4u : 6u);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t co = min(c0 + j, Wo - 1u), col0 = co * NR;
            uint32_t sy = 0, sb = 0, sr = 0;
            if (!P.case_b) {
#pragma unroll
              for (uint32_t dr = 0; dr < NR; ++dr) {
                const bool hl = vheld && (dr & 1u);       // nothing is sampled on an odd line: it replays the held pair
                uint32_t xb = hxb[dr >> 1], xr = hxr[dr >> 1];
                if (!hl && (col0 & hfm) != 0u) {          // the block starts inside a hold group (hf > AF): its sample lies to the left
                  const uint32_t pc = grow_px(rows[dr], (col0 & ~hfm) * ipb);
                  xb = fwd_nc16<TRUNC>(pc, P.coef_ncb); xr = fwd_nc16<TRUNC>(pc, P.coef_ncr);
                }
#pragma unroll
                for (uint32_t dc = 0; dc < NR; ++dc) {
                  const uint32_t pv = grow_px(rows[dr], (col0 + dc) * ipb);
                  sy += (fwd_y16(pv, P.coef_y) >> 8) & qy;
                  if (!hl && ((col0 + dc) & hfm) == 0u) { xb = fwd_nc16<TRUNC>(pv, P.coef_ncb); xr = fwd_nc16<TRUNC>(pv, P.coef_ncr); }
                  sb += (~(xb >> 8)) & qb; sr += (~(xr >> 8)) & qr;
                }
              }
            } else {
              // pooling first: the chroma stage then samples the POOLED stream with the full-size counters, i.e. this
              // pixel takes the pooled chroma of block (bro, bco) (ImageCompressorTop.scala:52-58)
              int bro, bco;
              chroma_src_case_b(P, (int)ro, (int)co, bro, bco);
#pragma unroll
              for (uint32_t dr = 0; dr < NR; ++dr) {
                const GRow cr_ = grow_at(in_row((uint32_t)bro * f + dr), last_word);
#pragma unroll
                for (uint32_t dc = 0; dc < NR; ++dc) {
                  sy += (fwd_y16(grow_px(rows[dr], (col0 + dc) * ipb), P.coef_y) >> 8) & qy;
                  const uint32_t pc = grow_px(cr_, ((uint32_t)bco * NR + dc) * ipb);
                  sb += (~(fwd_nc16<TRUNC>(pc, P.coef_ncb) >> 8)) & qb; sr += (~(fwd_nc16<TRUNC>(pc, P.coef_ncr) >> 8)) & qr;
                }
              }
            }
            y[j] = (sy + half) >> sh; cb[j] = (sb + half) >> sh; cr[j] = (sr + half) >> sh;
            if (!P.quant_first) { y[j] &= my; cb[j] &= mcb; cr[j] &= mcr; }
          }
        }
        // ---- pack the granule: into the warp's staging slot, or (narrow rows) straight to global memory ----
        const uint32_t so = st + kG * lane;
        uint8_t* gp = orow + (size_t)g * kG;
        const uint32_t gbytes = min(kG, row_bytes - g * kG);                  // the row's last granule may be partial
        uint32_t ww[kG / 4u];
        if (FMT == KF_YCC888 || FMT == KF_RGB888) {
          uint32_t v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            v[j] = FMT == KF_RGB888 ? inverse_rgb((int)y[j], (int)cb[j], (int)cr[j]) : (y[j] | (cb[j] << 8) | (cr[j] << 16));
          pack_rgb_granule(v, ww[0], ww[1], ww[2]);
        } else if (FMT == KF_PLANAR) {
          ww[0] = y[0] | (y[1] << 8) | (y[2] << 16) | (y[3] << 24);
          if (row_ok && ro % (uint32_t)P.planar_vs == 0u) {                  // surviving chroma sample points of this row
            const size_t crow = (size_t)(ro / (uint32_t)P.planar_vs) * (size_t)P.planar_cw;
#pragma unroll
            for (uint32_t j = 0; j < 4u; ++j)
              if ((j & ((uint32_t)P.planar_hs - 1u)) == 0u && c0 + j < Wo) {
                fout[P.planar_cb_off + crow + (c0 + j) / (uint32_t)P.planar_hs] = (uint8_t)cb[j];
                fout[P.planar_cr_off + crow + (c0 + j) / (uint32_t)P.planar_hs] = (uint8_t)cr[j];
              }
          }
        } else {
          uint32_t v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            v[j] = (c0 + j < Wo) ? ((y[j] >> P.sy) << (P.cb_bits + P.cr_bits)) | ((cb[j] >> P.scb) << P.cr_bits) | (cr[j] >> P.scr)
                                 : 0u;                                        // BUNDLE row padding: zero slots
          if (FMT == KF_SLOT32) { ww[0] = v[0]; ww[kG / 4u > 1u ? 1 : 0] = v[1]; ww[kG / 4u > 2u ? 2 : 0] = v[2]; ww[kG / 4u > 3u ? 3 : 0] = v[3]; }
          else if (FMT == KF_SLOT16) { ww[0] = v[0] | (v[1] << 16); ww[kG / 4u > 1u ? 1 : 0] = v[2] | (v[3] << 16); }
          else ww[0] = v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24);
        }
        if (narrow) {
          if (row_ok) direct_store(gp, ww, gbytes);
        } else {
#pragma unroll
          for (uint32_t i = 0; i < kG / 4u; ++i) sts32(so + 4u * i, ww[i]);
        }
      }
      if (narrow) break;                                                     // the rows of this trip are done
      __syncwarp();
      span_store(og, st, min(32u * kG, row_bytes - g0 * kG), lane, 32u);   // the row's last granule may be partial
      __syncwarp();
    }
  }
}

// Decoder of the PLANAR format for ANY width and any alignment of the planes and of the output: chroma fetched with the
// reference's replay rule (ChromaSubsampler.scala:52-65) expressed in output coordinates (chroma before spatial, or
// f == 1).  One CTA per output row segment of up to 4096 pixels; the only division (frame / row split) and everything
// row-dependent happen once per row.
//   load   a thread takes up to four granules of four pixels and issues ALL their loads first (Y word, Cb and Cr
//          samples: aligned LDG.32 pairs + funnel shifts on per-row word bases, 32-bit offsets), so a warp keeps ~24
//          loads in flight instead of 3;
//   store  the whole segment is staged in shared memory at the output's own offset modulo 16 (rounded down to a
//          word) and leaves as 16-byte st.global.cs aligned on the GLOBAL address (span_store): one head and one tail
//          of < 16 bytes per row instead of one per 384 bytes.
// Round 1: one thread per pixel, three divisions, three byte loads and three byte stores each -- 0.10 of the copy
// peak; the first round-2 version (one granule per thread and trip, per-warp staging) 0.22; this one 0.4 - 0.5.  Since
// the TMA-staged decoder (csic_decode_kernel.cu: 0.6 - 0.95) it only runs what that kernel declines, and under
// CSIC_DEC_NO_TMA in the tests.
constexpr uint32_t kExpandSeg = 4096u;   // pixels per staged segment: 12 KB of shared memory

__global__ void __launch_bounds__(128) csic_expand_planar_any_kernel(const __grid_constant__ KPlan P, const uint8_t* __restrict__ planar,
                                                                     uint8_t* __restrict__ out, int to_rgb) {
  __shared__ __align__(16) uint8_t stage[kExpandSeg * 3u + 48u];      // + alignment offset + read-ahead slack of span_store
  const uint32_t Wo = (uint32_t)P.Wo, tid = threadIdx.x, NT = blockDim.x;
  const uint32_t n_rows = P.n_frames * (uint32_t)P.Ho;
  const uint32_t last_c = (uint32_t)((P.last_sample_col / P.f) / P.planar_hs);   // plane column of a line's last sample point
  const uint32_t hs_sh = P.planar_hs == 4 ? 2u : (P.planar_hs == 2 ? 1u : 0u), vs_sh = P.planar_vs == 2 ? 1u : 0u;
  const uint32_t csel = hs_sh == 0 ? 0x3210u : (hs_sh == 1 ? 0x1100u : 0x0000u);   // sample bytes -> the four pixels of a granule
  const bool vhold = P.vf == 2 && P.f == 1;                         // odd lines replay the line above (f == 1 only)
  // last aligned word that still holds a byte of the planar buffer (loads never go past it)
  const uint8_t* last_word = reinterpret_cast<const uint8_t*>(
      (reinterpret_cast<uintptr_t>(planar) + (uint64_t)P.n_frames * P.out_frame_bytes - 1u) & ~(uintptr_t)3);
  const uint32_t sbase = smem_u32(stage);
  for (uint32_t R = blockIdx.x; R < n_rows; R += gridDim.x) {
    const uint32_t k = R / (uint32_t)P.Ho, ro = R - k * (uint32_t)P.Ho;
    const uint8_t* fr = planar + (uint64_t)k * P.out_frame_bytes;
    const bool held = vhold && (ro & 1u);
    const size_t crow = (size_t)((held ? ro - 1u : ro) >> vs_sh) * (size_t)P.planar_cw;
    const GRow yr = grow_at(fr + (size_t)ro * Wo, last_word);
    const GRow br = grow_at(fr + P.planar_cb_off + crow, last_word), rr = grow_at(fr + P.planar_cr_off + crow, last_word);
    uint8_t* orow = out + (uint64_t)R * Wo * 3u;
    uint32_t hcb = 0, hcr = 0;
    if (held) { hcb = (grow_px(br, last_c) & 0xFFu) * 0x01010101u; hcr = (grow_px(rr, last_c) & 0xFFu) * 0x01010101u; }
    for (uint32_t seg0 = 0; seg0 < Wo; seg0 += kExpandSeg) {
      const uint32_t npx = min(kExpandSeg, Wo - seg0), ngr = (npx + 3u) >> 2;
      uint8_t* og = orow + (size_t)seg0 * 3u;
      const uint32_t st = sbase + ((uint32_t)reinterpret_cast<uintptr_t>(og) & 12u);
      for (uint32_t g0 = tid; g0 < ngr; g0 += 4u * NT) {
        // All the loads of up to four granules first, as bare LDGs on clamped indices: no branch and no dependent
        // instruction between them (a funnel shift right behind its two loads stalls the warp -- the GPU issues in
        // order -- and the next granule's loads with it: ncu showed 11.7 long-scoreboard stalls per issue that way).
        uint32_t yl[4], yh[4], bl[4], bh[4], rl[4], rh[4];
        const uint32_t oc0 = seg0 >> hs_sh;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t g = min(g0 + (uint32_t)u * NT, ngr - 1u);
          const uint32_t wy = (yr.a0 + seg0 + 4u * g) >> 2;
          yl[u] = __ldg(yr.base + wy); yh[u] = __ldg(yr.base + min(wy + 1u, yr.maxw));
        }
        if (!held) {
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t g = min(g0 + (uint32_t)u * NT, ngr - 1u), cs = oc0 + ((4u * g) >> hs_sh);
            const uint32_t wb = (br.a0 + cs) >> 2, wr = (rr.a0 + cs) >> 2;
            bl[u] = __ldg(br.base + wb); bh[u] = __ldg(br.base + min(wb + 1u, br.maxw));
            rl[u] = __ldg(rr.base + wr); rh[u] = __ldg(rr.base + min(wr + 1u, rr.maxw));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const uint32_t g = g0 + (uint32_t)u * NT;
          if (g < ngr) {
            const uint32_t cs = oc0 + ((4u * g) >> hs_sh);
            const uint32_t yw = __funnelshift_r(yl[u], yh[u], ((yr.a0 + seg0 + 4u * g) & 3u) * 8u);
            uint32_t cbw = hcb, crw = hcr;
            if (!held) {
              cbw = __byte_perm(__funnelshift_r(bl[u], bh[u], ((br.a0 + cs) & 3u) * 8u), 0, csel);
              crw = __byte_perm(__funnelshift_r(rl[u], rh[u], ((rr.a0 + cs) & 3u) * 8u), 0, csel);
            }
            uint32_t w0, w1, w2;
            decode_granule(yw, cbw, crw, held ? 2u : hs_sh, to_rgb, w0, w1, w2);
            sts32(st + 12u * g, w0); sts32(st + 12u * g + 4u, w1); sts32(st + 12u * g + 8u, w2);
          }
        }
      }
      __syncthreads();
      span_store(og, st, npx * 3u, tid, NT);                           // the row's last granule may hold fewer than four pixels
      __syncthreads();
    }
  }
}

int launch_expand_planar(const KPlan& k, const uint8_t* planar, uint8_t* out, int to_rgb, int sm_count, size_t max_smem_optin, void* stream) {
  const uint64_t total = (uint64_t)k.n_frames * (uint64_t)k.Ho * (uint64_t)k.Wo;
  if (total == 0) return (int)cudaSuccess;
  // the TMA-staged decoder (csic_decode_kernel.cu) takes every realistic shape; what it declines -- rows narrower than a
  // granule, chroma rows too wide for a stage -- runs on the LDG kernel above
  const int tma = launch_decode_tma(k, planar, out, to_rgb, sm_count, max_smem_optin, stream);
  if (tma >= 0) return tma;
  const uint64_t n_rows = (uint64_t)k.n_frames * (uint64_t)k.Ho;
  if (n_rows >= (1ull << 32)) return (int)cudaErrorInvalidValue;
  const unsigned threads = (unsigned)std::min<uint32_t>(128u, ((((uint32_t)k.Wo + 3u) >> 2) + 31u) & ~31u);
  const unsigned blocks = (unsigned)std::min<uint64_t>(n_rows, (uint64_t)sm_count * (2048u / threads));
  csic_expand_planar_any_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(k, planar, out, to_rgb);
  return (int)cudaGetLastError();
}

namespace {
template <int FMT, int AF>
int launch_generic_af(const KPlan& k, unsigned blocks, unsigned threads, cudaStream_t st) {
  if (k.trunc) csic_generic_kernel<true, FMT, AF><<<blocks, threads, 0, st>>>(k);
  else csic_generic_kernel<false, FMT, AF><<<blocks, threads, 0, st>>>(k);
  return (int)cudaGetLastError();
}
template <int FMT>
int launch_generic_fmt(const KPlan& k, unsigned blocks, unsigned threads, cudaStream_t st) {
  if (!(k.average && k.f > 1)) return launch_generic_af<FMT, 0>(k, blocks, threads, st);
  if (k.f == 2) return launch_generic_af<FMT, 2>(k, blocks, threads, st);
  if (k.f == 4) return launch_generic_af<FMT, 4>(k, blocks, threads, st);
  return launch_generic_af<FMT, 8>(k, blocks, threads, st);
}
}  // namespace

int launch_generic(const KPlan& k, int sm_count, void* stream) {
  const uint64_t n_rows = (uint64_t)k.n_frames * (uint64_t)k.band_rows;
  if (n_rows == 0 || k.slots_per_row == 0) return (int)cudaSuccess;
  if (n_rows >= (1ull << 32)) return (int)cudaErrorInvalidValue;
  // one warp per row, or per 32 / gpr rows when a row has fewer than 32 granules; four warps per CTA
  const uint32_t gpr = ((uint32_t)k.slots_per_row + 3u) >> 2;
  const uint64_t groups = gpr < 32u ? (n_rows + 32u / gpr - 1u) / (32u / gpr) : n_rows;
  const unsigned threads = 128u;
  const unsigned blocks = (unsigned)std::min<uint64_t>((groups + 3u) / 4u, (uint64_t)sm_count * 16u);
  cudaStream_t st = (cudaStream_t)stream;
  switch (k.kformat) {
    case KF_YCC888: return launch_generic_fmt<KF_YCC888>(k, blocks, threads, st);
    case KF_RGB888: return launch_generic_fmt<KF_RGB888>(k, blocks, threads, st);
    case KF_SLOT8: return launch_generic_fmt<KF_SLOT8>(k, blocks, threads, st);
    case KF_SLOT16: return launch_generic_fmt<KF_SLOT16>(k, blocks, threads, st);
    case KF_PLANAR: return launch_generic_fmt<KF_PLANAR>(k, blocks, threads, st);
    default: return launch_generic_fmt<KF_SLOT32>(k, blocks, threads, st);
  }
}

// ================================================================================================
// TMA-staged row kernel: planning and dispatch (the kernel itself lives in csic_rows_kernel.cu)
// ================================================================================================
bool plan_rows_kernel(KPlan& k, int sm_count, size_t max_smem_optin, int force_stages, uint32_t force_tile_bytes) {
  if (k.average && k.f > 1) return false;                       // AVERAGE extension: csic_pool_kernel / generic
  if (k.band_rows <= 0 || k.n_frames == 0) return false;
  const bool auto_threads = k.block_threads <= 0;
  if (auto_threads) k.block_threads = kDefaultBlockThreads;
  if (k.block_threads > kMaxConsumerThreads) return false;
  // The kernel processes Wp = Wo rounded up to 16 output pixels per row; columns >= Wo are read from / written to
  // the row padding, so both pitches must cover Wp (dense buffers qualify when Wo % 16 == 0).
  k.Wp = (k.Wo + 15) & ~15;
  const bool planar = k.kformat == KF_PLANAR;
  const uint32_t opx0 = planar ? 1u : (k.kformat <= KF_RGB888 ? 3u : (uint32_t)k.slot_bytes);
  const uint64_t need_in = (uint64_t)k.Wp * (uint32_t)k.f * (uint32_t)k.in_px_bytes, need_out = (uint64_t)k.Wp * opx0;
  if (k.in_row_bytes < need_in || k.out_row_bytes < need_out) return false;
  if ((k.in_row_bytes | k.out_row_bytes) & 15u) return false;    // 16-byte TMA granularity of every row start
  if ((reinterpret_cast<uintptr_t>(k.in) | reinterpret_cast<uintptr_t>(k.out)) & 15u) return false;
  if ((k.in_frame_bytes | k.out_frame_bytes) & 15u) return false;
  if (k.case_b) {
    // spatial before chroma: a counter line must be exactly f output rows, each starting on a sample column
    if (k.W % k.f != 0 || k.Wo % 4 != 0) return false;
    k.caseb_row_add = (uint32_t)k.last_sample_col / (uint32_t)k.Wo;
    k.caseb_col_bytes = ((uint32_t)k.last_sample_col % (uint32_t)k.Wo) * (uint32_t)k.in_px_bytes * (uint32_t)k.f;
  }
  k.ragged = k.Wp != k.Wo;
  k.in_dense = k.in_row_bytes == need_in;
  k.out_dense = k.out_row_bytes == need_out;
  // PLANAR: three dense planes; chroma rows (Wo / hs bytes) must be 16-byte multiples too; bands start on a
  // chroma row
  if (planar && (k.ragged || !k.out_dense || k.Wo % (16 * k.planar_hs) != 0 || k.row0 % k.planar_vs != 0)) return false;

  // hold width inside a granule, in output pixels
  if (!k.case_b) k.hfe = std::max(1, k.hf / k.f);
  else k.hfe = k.hf;

  // 32-bit bundle slots at f == 1 leave through shared memory + TMA bulk stores like the 3-byte formats: there the
  // output is larger than the input (4 B per 3 B pixel) and per-warp 512-byte st.global bursts held the kernel at 0.90
  // of the copy peak (ncu: warps waiting on the store path, DRAM 70 % busy) -- 0.98 staged.  Narrower slots and f >= 2
  // keep the direct stores (16-bit slots at f == 1: 0.94 direct, 0.83 staged; cfg4: 1.03).  Must match kStaged in
  // csic_rows_kernel.cu.
  const bool staged = k.kformat <= KF_RGB888 || planar || (k.f == 1 && k.kformat == KF_SLOT32);
  const uint32_t opx = planar ? 1u : (k.kformat <= KF_RGB888 ? 3u : (uint32_t)k.slot_bytes);
  // Tile budget: input bytes of one tile.  Staged formats also hold two output buffers per CTA.
  const uint32_t tile_budget = force_tile_bytes ? force_tile_bytes : 24u * 1024u;
  const uint32_t ipb = (uint32_t)k.in_px_bytes;
  const uint32_t row_in = (uint32_t)k.Wp * ipb * (uint32_t)k.f;  // bytes of one processed input row
  int nsplit = 0;
  const uint32_t tile_max = tile_budget + tile_budget / 3;        // a tile may overshoot the budget by a third
  for (int n = (int)((row_in + tile_max - 1) / tile_max); n <= 64; ++n) {
    if (n >= 1 && k.Wp % (16 * (planar ? k.planar_hs : 1) * n) == 0) { nsplit = n; break; }
  }
  if (nsplit == 0) return false;
  k.nsplit = nsplit;
  k.tile_px = k.Wp / nsplit;
  k.tile_in_bytes = (uint32_t)k.tile_px * ipb * (uint32_t)k.f;    // one row segment
  k.tile_out_bytes = (uint32_t)k.tile_px * opx;
  // "Tall image" (as in the flex kernel): when every launch row is a whole frame row and frames lie back to back with the
  // same row stride inside and across frames, on the input AND the output side, the batch IS one image of n_frames * Ho
  // rows, and a tile may span frames -- 32x32 frames make 64-row tiles of two frames instead of one 3 KB tile per frame
  // (0.74 -> 0.95 of the copy peak).  The kernel does not change: the only rule that looks at a row index, "odd lines
  // replay the line above" (f == 1, 4:2:0 / 4:1:0), reads the same parity when frames have an even number of rows.
  const uint64_t stored_rows = (uint64_t)(uint32_t)k.Ho * (uint32_t)k.row_step;   // input rows from one frame's first row to the next frame's
  static const bool no_tall = std::getenv("CSIC_ROWS_NO_TALL") != nullptr;        // experiment switch (tools/): never changes results
  const bool tall = !no_tall && nsplit == 1 && !planar && !k.case_b && k.n_frames > 1 && k.row0 == 0 && k.band_rows == k.Ho &&
                    k.in_frame_bytes == stored_rows * k.in_row_bytes && k.out_frame_bytes == (uint64_t)(uint32_t)k.Ho * k.out_row_bytes &&
                    !(k.vf == 2 && k.f == 1 && (k.Ho & 1)) &&
                    (uint64_t)k.n_frames * (uint32_t)k.Ho < (1ull << 30) && (uint64_t)k.n_frames * (uint32_t)k.H < (1ull << 31);
  const uint32_t frames = tall ? 1u : k.n_frames;
  const int band_rows = tall ? (int)(k.n_frames * (uint32_t)k.Ho) : k.band_rows;
  // Rows per tile: whole rows only (so the tile's output is contiguous), as many as fit the budget,
  // but keep at least ~4 tiles per SM so small batches still spread over the chip.
  int rows = 1;
  if (nsplit == 1) {
    rows = (int)std::min<uint32_t>((uint32_t)kMaxTileRows, std::max<uint32_t>(1u, tile_budget / k.tile_in_bytes));
    // e.g. 15 KB RGBA 4K rows: one row leaves the ring too shallow, two rows (30 KB) are within the slack
    if ((uint32_t)rows * k.tile_in_bytes < tile_budget * 7u / 10u && (uint32_t)(rows + 1) * k.tile_in_bytes <= tile_max &&
        rows < kMaxTileRows)
      ++rows;
    rows = std::min(rows, band_rows);
    auto tiles_for = [&](int r) { return (uint64_t)frames * (uint64_t)((band_rows + r - 1) / r); };
    while (rows > 1 && tiles_for(rows) < (uint64_t)sm_count * 4u) rows = (rows + 1) / 2;
    // staged 32-bit slots: a tile's two output buffers are larger than its input stages; keep two CTAs resident
    // (4096-pixel rows: two rows per tile left room for one CTA only -- 0.81 of the copy peak instead of 0.98)
    if (staged && !planar && k.kformat > KF_RGB888)
      while (rows > 1 && 2u * ((uint32_t)rows * (k.tile_in_bytes + 32u)) + 2u * ((uint32_t)rows * k.tile_out_bytes) + 2048u > 227u * 1024u / 2u - 1024u)
        --rows;
    // Equal tiles: CTAs take tiles round robin, so a short last tile per frame recurs with the period of the frame
    // (96 rows as 64 + 32: every other CTA got the 32-row tiles only and the kernel ran at 0.71 of the copy peak
    // where 64x64 and 128x128 frames reach 0.98 - 1.01).  Even row counts keep a held line and its sample in one tile.
    if (rows > 1) {
      const int t = (band_rows + rows - 1) / rows;
      int bal = (band_rows + t - 1) / t;
      if ((k.vf == 2 || planar) && (bal & 1) && bal < rows) ++bal;
      rows = std::min(rows, bal);
    }
    if (planar && rows > 1) rows &= ~(k.planar_vs - 1);          // tiles start on a chroma row
    if (rows < 1) rows = 1;
  }
  k.tile_rows = rows;
  k.tiles_per_band = (uint32_t)((band_rows + rows - 1) / rows);
  const uint64_t n_tiles = (uint64_t)frames * (uint64_t)k.tiles_per_band * (uint64_t)nsplit;
  if (n_tiles >= (1ull << 31)) return false;
  k.n_tiles = (uint32_t)n_tiles;
  // Tiny tiles (whole frames of a few KB: a tile never spans two frames): a granule per thread is all there is, so the
  // per-tile costs dominate -- half-size CTAs, as many as fit (32x32 frames: 0.74 of the copy peak vs 0.51;
  // profiles/r1/sweep_rows_v8.txt)
  const bool tiny = (uint32_t)rows * ((uint32_t)k.tile_px >> 2) <= 512u && n_tiles > (uint64_t)sm_count * 8u;
  if (tiny && auto_threads) k.block_threads = 128;

  auto up128 = [](uint32_t v) { return (v + 127u) & ~127u; };
  k.stage_stride = up128((uint32_t)rows * (k.tile_in_bytes + 32u));
  k.out_buf_stride = staged ? up128((uint32_t)rows * k.tile_out_bytes) : 0u;
  if (planar)   // [Y rows][Cb rows][Cr rows]
    k.out_buf_stride = up128((uint32_t)rows * (uint32_t)k.tile_px +
                             2u * (uint32_t)((rows + k.planar_vs - 1) / k.planar_vs) * ((uint32_t)k.tile_px / (uint32_t)k.planar_hs));
  // Ring depth and residency, fitted to B200 sweeps (profiles/r1/exp_ctas_v4.txt): the kernel is fastest with
  // about four ~24 KB tiles in flight per SM -- 2 resident CTAs x 2 stages (cfg4 0.996 of the measured copy peak
  // vs 0.953 with 4 CTAs; cfg2 0.97 vs 0.94 with 6 tiles in flight).  Only when a tile carries little compute
  // per byte (f >= 4: one pixel in 16 is converted) does a third stage pay (cfg5 0.98 vs 0.92).
  auto need = [&](int s) {
    return (uint32_t)s * k.stage_stride + 2u * k.out_buf_stride + (uint32_t)s * kTileMetaBytes + (uint32_t)s * 16u + 384u;
  };
  auto ctas_for = [&](int s) {
    const uint32_t by_smem = (uint32_t)(227u * 1024u / (need(s) + 1024u));
    return std::min<uint32_t>(std::min<uint32_t>(by_smem, 2048u / ((uint32_t)k.block_threads + 32u)), 8u);
  };
  int stages = force_stages >= 2 ? force_stages : (k.f >= 4 ? 3 : 2);
  while (force_stages < 2 && stages > 2 && ctas_for(stages) < 2) --stages;
  if (need(stages) > max_smem_optin || ctas_for(stages) < 1) return false;
  // f == 1 converts every byte it loads (issue slots 64 % busy with 16 warps): there more resident warps win
  // (PLANAR 1080p: 0.97 of the copy peak with 3 CTAs vs 0.84 with 2; 16-bit bundles: 0.93 with 4).
  k.ctas_per_sm = (int32_t)std::min<uint32_t>(ctas_for(stages), tiny ? 8u : (k.f == 1 ? 4u : 2u));
  k.stages = stages;
  if (tall) {                                      // nothing can fail from here on: commit the tall-image view of the batch
    k.in_frame_bytes *= k.n_frames; k.out_frame_bytes *= k.n_frames;
    k.H *= (int32_t)k.n_frames; k.Ho = band_rows; k.band_rows = band_rows;
    k.n_frames = 1;
  }
  k.out_buf_off = (uint32_t)stages * k.stage_stride;
  k.meta_off = up128(k.out_buf_off + 2u * k.out_buf_stride);
  k.bar_off = up128(k.meta_off + (uint32_t)stages * (uint32_t)kTileMetaBytes);
  k.smem_bytes = k.bar_off + (uint32_t)stages * 16u;     // full[] and empty[] mbarriers
  return true;
}

int rows_kernel_set_attributes(size_t max_smem_optin) {
  int e;
  if ((e = flex_set_attributes(max_smem_optin)) != 0) return e;
  if ((e = decode_set_attributes(max_smem_optin)) != 0) return e;
  if ((e = pool_set_attributes_factor<2>(max_smem_optin)) != 0) return e;
  if ((e = pool_set_attributes_factor<4>(max_smem_optin)) != 0) return e;
  if ((e = pool_set_attributes_factor<8>(max_smem_optin)) != 0) return e;
  if ((e = rows_set_attributes_factor<1>(max_smem_optin)) != 0) return e;
  if ((e = rows_set_attributes_factor<2>(max_smem_optin)) != 0) return e;
  if ((e = rows_set_attributes_factor<4>(max_smem_optin)) != 0) return e;
  if ((e = rows_set_attributes_factor<8>(max_smem_optin)) != 0) return e;
  return 0;
}

// ---- AVERAGE extension: planning and dispatch of csic_pool_kernel --------------------------------
bool plan_pool_kernel(KPlan& k, int sm_count, size_t max_smem_optin) {
  if (!(k.average && k.f > 1)) return false;
  if (k.case_b) {
    // pooling before chroma: a counter line (W == f * Wo stream elements) must be whole output rows that start on a
    // sample column; the held block of an odd line is the stream element (line-1) * W + lastSampleCol
    if (k.W % k.f != 0 || k.Wo % k.hf != 0) return false;
    k.caseb_row_add = (uint32_t)k.last_sample_col / (uint32_t)k.Wo;
    k.caseb_col_bytes = ((uint32_t)k.last_sample_col % (uint32_t)k.Wo) * (uint32_t)k.in_px_bytes * (uint32_t)k.f;
  }
  // Like the row kernel, the kernel processes Wp = Wo rounded up to 16 output pixels per row: columns >= Wo are read
  // from / written into the row padding, so both pitches must cover Wp (dense buffers qualify when Wo % 16 == 0;
  // csic_process_host re-pitches other widths in its staging buffers, csic_process_device_pitched takes them as given).
  k.Wp = (k.Wo + 15) & ~15;
  const uint32_t opx_p = k.kformat <= KF_RGB888 ? 3u : (uint32_t)k.slot_bytes;
  if (k.in_row_bytes % 16 != 0 || (uint64_t)k.in_row_bytes < (uint64_t)k.Wp * (uint32_t)k.f * (uint32_t)k.in_px_bytes) return false;
  if ((uint64_t)k.out_row_bytes < (uint64_t)k.Wp * opx_p) return false;
  if ((reinterpret_cast<uintptr_t>(k.in) | reinterpret_cast<uintptr_t>(k.out)) & 15u) return false;
  if ((k.in_frame_bytes | k.out_frame_bytes | k.out_row_bytes) & 15u) return false;
  k.ragged = k.Wp != k.Wo;
  k.in_dense = (uint64_t)k.in_row_bytes == (uint64_t)k.Wp * (uint32_t)k.f * (uint32_t)k.in_px_bytes;    // the f rows of a block row are contiguous
  k.out_dense = (uint64_t)k.out_row_bytes == (uint64_t)k.Wp * opx_p;                                     // output rows are back to back
  if (k.band_rows <= 0 || k.n_frames == 0) return false;
  // 4x4 / 8x8 pooling keeps more live registers per thread: 6 consumer warps leave room for one more CTA per SM
  // (B200 sweep, profiles/r1/sweep_pool_v8.txt: 8K 4x4 0.97 of the copy peak vs 0.92 with 8 warps)
  const bool auto_threads = k.block_threads <= 0;
  if (auto_threads) k.block_threads = k.f >= 4 ? 192 : kDefaultBlockThreads;
  if (k.block_threads > kMaxConsumerThreads) return false;

  const bool staged = k.kformat <= KF_RGB888;
  const uint32_t opx = staged ? 3u : (uint32_t)k.slot_bytes;
  const uint32_t f = (uint32_t)k.f, ipb = (uint32_t)k.in_px_bytes;
  const uint32_t budget = 24u * 1024u, tile_max = budget + budget / 3;   // (48 KB tiles for 8x8 pooling: 1080p +5 %, 4K / 8K -12 %)
  const uint32_t block_row_bytes = (uint32_t)k.Wp * f * f * ipb;          // the f input rows of one output row
  int nsplit = 0;
  for (int n = (int)((block_row_bytes + tile_max - 1) / tile_max); n <= 256; ++n)
    if (n >= 1 && k.Wp % (16 * n) == 0) { nsplit = n; break; }
  if (nsplit == 0) return false;
  k.nsplit = nsplit;
  k.tile_px = k.Wp / nsplit;
  k.tile_in_bytes = (uint32_t)k.tile_px * f * f * ipb;                   // per output-row segment: f row parts
  k.tile_out_bytes = (uint32_t)k.tile_px * opx;
  int rows = 1;
  if (nsplit == 1) {
    rows = (int)std::max<uint32_t>(1u, budget / k.tile_in_bytes);
    // e.g. 2560x1440 2x2 or 640x480 8x8: an output row carries 15 KB of input -- one row leaves the ring too shallow, two
    // rows (30 KB) are within the slack (the row kernel's rule)
    if ((uint32_t)rows * k.tile_in_bytes < budget * 7u / 10u && (uint32_t)(rows + 1) * k.tile_in_bytes <= tile_max) ++rows;
    rows = std::min(rows, (int)(2u * (uint32_t)kPoolMaxRows / f));      // held_addr[] holds rows * f/2 entries
    rows = std::min(rows, k.band_rows);
    if (!k.out_dense) rows = 1;                                           // a tile's output must be one contiguous range
    auto tiles_for = [&](int r) { return (uint64_t)k.n_frames * (uint64_t)((k.band_rows + r - 1) / r); };
    while (rows > 1 && tiles_for(rows) < (uint64_t)sm_count * 4u) rows = (rows + 1) / 2;
    if (rows > 1) {                                                       // equal tiles (see plan_rows_kernel)
      const int t = (k.band_rows + rows - 1) / rows;
      rows = std::min(rows, (k.band_rows + t - 1) / t);
    }
  }
  k.tile_rows = rows;
  k.tiles_per_band = (uint32_t)((k.band_rows + rows - 1) / rows);
  const uint64_t n_tiles = (uint64_t)k.n_frames * (uint64_t)k.tiles_per_band * (uint64_t)nsplit;
  if (n_tiles >= (1ull << 31)) return false;
  k.n_tiles = (uint32_t)n_tiles;
  // 8x8 pooling is bound by the integer pipes, and a granule is 256 input pixels: a 15 KB tile has twenty of them, four
  // threads each -- 80 busy threads of 192 (1080p: 240 output pixels per row, split three ways).  Size the CTA to the
  // tile and let more of them be resident instead (1080p 8x8: 0.69-0.81 -> 0.78-0.91 of the copy peak).
  const uint32_t units = (uint32_t)rows * ((uint32_t)k.tile_px >> 2) * 4u;
  const bool small_cta = auto_threads && k.f == 8 && 2u * units <= (uint32_t)k.block_threads;   // (tiles that fill 2/3 of the CTA or more: 3-4 % slower this way)
  if (small_cta) k.block_threads = (int32_t)std::max(64u, (units + 31u) & ~31u);
  auto up128 = [](uint32_t v) { return (v + 127u) & ~127u; };
  k.stage_stride = up128((uint32_t)rows * k.tile_in_bytes + (uint32_t)kPoolMaxRows * 32u);
  k.out_buf_stride = staged ? (uint32_t)(kMaxConsumerThreads / 32) * 384u / 2u : 0u;   // 2 x stride = one 384-byte slot per warp
  k.stages = 2;
  const uint32_t need = 2u * k.stage_stride + 2u * k.out_buf_stride + 2u * kPoolMetaBytes + 32u + 384u;
  if (need > max_smem_optin) return false;
  // pooling converts every input pixel: the kernel is issue-bound, so take all the warps that fit (measured:
  // 4 CTAs 0.79 of the copy peak on the 4K 2x2 case vs 0.73 with 2)
  const uint32_t threads = (uint32_t)k.block_threads + 32u;
  const uint32_t by_regs = 65536u / ((k.f == 2 ? 56u : 72u) * threads);     // __maxnreg__ of csic_pool_kernel
  k.ctas_per_sm = (int32_t)std::max<uint32_t>(1u, std::min<uint32_t>(std::min<uint32_t>(small_cta ? 8u : 4u, by_regs), 227u * 1024u / (need + 1024u)));
  k.out_buf_off = 2u * k.stage_stride;
  k.meta_off = up128(k.out_buf_off + 2u * k.out_buf_stride);
  k.bar_off = up128(k.meta_off + 2u * kPoolMetaBytes);
  k.smem_bytes = k.bar_off + 2u * 16u;
  return true;
}

int launch_pool(const KPlan& k, int sm_count, int force_ctas_per_sm, void* stream) {
  const int per_sm = force_ctas_per_sm > 0 ? force_ctas_per_sm : std::max(1, k.ctas_per_sm);
  const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)k.n_tiles, (uint64_t)sm_count * (uint64_t)per_sm);
  switch (k.f) {
    case 2: return launch_pool_factor<2>(k, grid, stream);
    case 4: return launch_pool_factor<4>(k, grid, stream);
    default: return launch_pool_factor<8>(k, grid, stream);
  }
}

int launch_rows(const KPlan& k, int sm_count, int force_ctas_per_sm, void* stream) {
  // Persistent grid: a whole number of CTAs per SM, as many as the shared-memory footprint allows.
  uint32_t per_sm = (uint32_t)std::max(1, k.ctas_per_sm);
  if (force_ctas_per_sm > 0) per_sm = (uint32_t)force_ctas_per_sm;   // tuning knob; the hardware caps residency
  const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)k.n_tiles, (uint64_t)sm_count * per_sm);
  switch (k.f) {
    case 1: return launch_rows_factor<1>(k, grid, stream);
    case 2: return launch_rows_factor<2>(k, grid, stream);
    case 4: return launch_rows_factor<4>(k, grid, stream);
    default: return launch_rows_factor<8>(k, grid, stream);
  }
}

}  // namespace csic
