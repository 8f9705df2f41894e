// csic_flex_kernel<FMT, TRUNC> -- the DECIMATE pipeline for buffers the TMA row kernel's 16-byte rules exclude:
// any frame width, any row pitch, any base-pointer alignment (dense odd-width frames on the device, sub-views of a
// larger buffer, bundle rows with pad slots).  Same arithmetic and the same closed-form source maps as
// csic_rows_kernel; what differs is how bytes move:
//
//   load     a tile = up to 16 whole output rows (or one row segment).  Every input row span is copied into shared
//            memory at the SAME offset modulo 16 it has in global memory, as whole 16-byte chunks through cp.async
//            (LDGSTS.128, L2 evict-first).  The chunks that straddle a span's ends bring along a few bytes of the
//            neighbouring pixels / rows of the same buffer; only at the two ends of the byte range the launch may touch
//            are the < 16 edge bytes copied one by one, so the kernel is exact on sub-buffers.
//            Two input stages per CTA: the next tile's copies are in flight while the current one is converted.
//   compute  one thread per granule of 4 output pixels; a pixel at an arbitrary byte address is two aligned LDS.32
//            and one funnel shift.  dp4a colour matrix, in-granule chroma hold, held rows from one pixel per row
//            fetched with the tile (ChromaSubsampler.scala:52-65), quantise, pack -- into an aligned staging area.
//   store    the staging area leaves as 16-byte st.global.cs words aligned on the GLOBAL address (shared-memory side
//            re-aligned with funnel shifts), head / tail bytes with byte stores: coalesced whatever the row size.
//
// Reference semantics as in csic_kernels.cu's header (RGB2YCbCr.scala:33-76, ChromaSubsampler.scala:26-65,
// SpatialDownsampler.scala:17-55, ColorQuantizer.scala:29-44, RGB2YCbCr.scala:123-132, ImageCompressorTop.scala:43-58).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "csic_internal.h"
#include "csic_device_math.cuh"
#include "csic_tma.cuh"

namespace csic {

namespace {

constexpr int kFlexThreads = 256;
constexpr uint32_t kFlexTileBytes = 12u * 1024u;     // input bytes of one tile
constexpr uint32_t kCtaWideSpan = 2048u;             // row spans at least this long are copied by the whole CTA

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, uint64_t pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void sts8(uint32_t a, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(a), "r"(v) : "memory");
}
// issued where it is written (never sunk towards its use): the value is consumed a whole tile later
__device__ __forceinline__ uint32_t ldg8_now(const uint8_t* p) {
  uint32_t v;
  asm volatile("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

// global [g, g+len) -> shared, byte i at  slot + (g & 15) + i  (slot is 16-byte aligned).  Whole 16-byte chunks go
// through cp.async; the chunks that straddle the ends of the span are fetched whole too -- the extra bytes belong to
// the neighbouring pixels / rows / pitch padding of the same buffer -- unless that would leave [lo, hi), the byte range
// this launch may touch: only there (first and last span of a launch) the < 16 edge bytes are copied one by one.
__device__ __forceinline__ void span_load(uint32_t slot, const uint8_t* __restrict__ g, uint32_t len, uint32_t t, uint32_t nthr,
                                          uint64_t pol, uintptr_t lo, uintptr_t hi) {
  const uintptr_t A = reinterpret_cast<uintptr_t>(g), B = A + len, base = A & ~(uintptr_t)15;
  uintptr_t start = base, end = (B + 15) & ~(uintptr_t)15;
  if (start < lo) start += 16;                   // start > A: head bytes [A, min(start, B)) by hand
  if (end > hi) end -= 16;                       // end < B: tail bytes by hand
  const uint8_t* gb = reinterpret_cast<const uint8_t*>(base);
  if (end > start) {
    const uint32_t o = (uint32_t)(start - base), nchunk = (uint32_t)(end - start) >> 4;
    for (uint32_t c = t; c < nchunk; c += nthr) cp_async16(slot + o + (c << 4), gb + o + (c << 4), pol);
  }
  if (start > A || end < B) {
    const uintptr_t h1 = start > A ? (start < B ? start : B) : A;          // head is [A, h1)
    const uintptr_t t0 = end < B ? (end > h1 ? end : h1) : B;              // tail is [t0, B)
    const uint32_t nh = (uint32_t)(h1 - A), nt = (uint32_t)(B - t0);
    for (uint32_t i = t; i < nh + nt; i += nthr) {
      const uint32_t off = (uint32_t)(A - base) + (i < nh ? i : (uint32_t)(t0 - A) + (i - nh));
      sts8(slot + off, __ldg(gb + off));
    }
  }
}

// shared [ssrc, ssrc+len) -> global [g, g+len): 16-byte stores aligned on the global address.
__device__ __forceinline__ void span_store(uint8_t* __restrict__ g, uint32_t ssrc, uint32_t len, uint32_t t,
                                           uint32_t nthr) {
  const uint32_t head = min(len, (16u - ((uint32_t)reinterpret_cast<uintptr_t>(g) & 15u)) & 15u);
  const uint32_t nchunk = (len - head) >> 4;
  const uint32_t tail = len - head - (nchunk << 4);
  const uint32_t s0 = ssrc + head, sa = s0 & ~3u, sh = (s0 & 3u) * 8u;
  if (sh == 0 && (sa & 15u) == 0) {
    for (uint32_t c = t; c < nchunk; c += nthr) __stcs(reinterpret_cast<uint4*>(g + head + (c << 4)), lds128(sa + (c << 4)));
  } else {
    for (uint32_t c = t; c < nchunk; c += nthr) {
      const uint32_t a = sa + (c << 4);
      const uint32_t w0 = lds32(a), w1 = lds32(a + 4), w2 = lds32(a + 8), w3 = lds32(a + 12), w4 = lds32(a + 16);
      __stcs(reinterpret_cast<uint4*>(g + head + (c << 4)),
             make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                        __funnelshift_r(w3, w4, sh)));
    }
  }
  for (uint32_t i = nthr - 1u - t; i < head + tail; i += nthr) {
    const uint32_t off = i < head ? i : len - tail + (i - head);
    g[off] = (uint8_t)lds8(ssrc + off);
  }
}

// `n` row spans of `len` bytes: one after the other with the whole CTA when they are long, one warp per span when short.
template <typename F>
__device__ __forceinline__ void for_each_span(uint32_t n, uint32_t len, F&& fn) {
  const uint32_t tid = threadIdx.x, NT = blockDim.x;
  if (len >= kCtaWideSpan || n == 1) {
    for (uint32_t j = 0; j < n; ++j) fn(j, tid, NT);
  } else {
    for (uint32_t j = tid >> 5; j < n; j += NT >> 5) fn(j, tid & 31u, 32u);
  }
}

// three colour bytes at an arbitrary shared-memory address (low three bytes of the result)
__device__ __forceinline__ uint32_t lds_px(uint32_t a) {
  const uint32_t b = a & ~3u;
  return __funnelshift_r(lds32(b), lds32(b + 4), (a & 3u) * 8u);
}

template <int FMT> struct FlexFmt {
  // staging bytes per granule of four slots
  static constexpr uint32_t kUnit = (FMT == KF_YCC888 || FMT == KF_RGB888) ? 12u : (FMT == KF_SLOT32 ? 16u : (FMT == KF_SLOT16 ? 8u : 4u));
};

// Where a tile sits: everything the three phases need, derived from the tile index.
struct FlexTile {
  const uint8_t* frame;      // input frame
  const uint8_t* src0;       // first input byte of the tile's first row span
  uint32_t k, ro0, nrows;    // frame, first output row, rows
  uint32_t col0, ncols, npx; // first slot, slots (pad slots included), pixels (>= 1)
  uint32_t len_in;           // bytes from the first sampled pixel's first byte to the last one's last byte
};
__device__ __forceinline__ FlexTile flex_tile(const KPlan& P, uint32_t tile) {
  FlexTile T;
  const uint32_t nsplit = (uint32_t)P.nsplit, tile_px = (uint32_t)P.tile_px;
  const uint32_t t2 = tile / nsplit, seg = tile - t2 * nsplit;
  T.k = t2 / P.tiles_per_band;
  const uint32_t tb = t2 - T.k * P.tiles_per_band;
  T.ro0 = (uint32_t)P.row0 + tb * (uint32_t)P.tile_rows;
  T.nrows = min((uint32_t)P.tile_rows, (uint32_t)(P.row0 + P.band_rows) - T.ro0);
  T.col0 = seg * tile_px;
  T.ncols = min(tile_px, (uint32_t)P.slots_per_row - T.col0);
  T.npx = min((uint32_t)P.Wo, T.col0 + T.ncols) - T.col0;
  const uint32_t pxb = (uint32_t)P.f * (uint32_t)P.in_px_bytes;
  T.len_in = (T.npx - 1u) * pxb + (uint32_t)P.in_px_bytes;
  T.frame = P.in + (uint64_t)T.k * P.in_frame_bytes;
  T.src0 = T.frame + (uint64_t)(T.ro0 * (uint32_t)P.row_step) * P.in_row_bytes + (uint64_t)T.col0 * pxb;
  return T;
}

}  // namespace

template <int FMT, bool TRUNC>
__global__ void __launch_bounds__(kFlexThreads) csic_flex_kernel(const __grid_constant__ KPlan P) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr uint32_t kUnit = FlexFmt<FMT>::kUnit, kOpx = kUnit / 4u;
  const uint32_t tid = threadIdx.x, NT = blockDim.x;
  const uint32_t sbase = smem_u32(smem), out_s = sbase + P.out_buf_off, held_base = sbase + P.meta_off;
  const uint32_t in_stage = P.stage_stride * (uint32_t)P.tile_rows + 32u;     // bytes of one input stage
  const uint64_t pol = policy_evict_first();
  const uint32_t ipb = (uint32_t)P.in_px_bytes, f = (uint32_t)P.f, pxb = f * ipb;
  const uint32_t nsplit = (uint32_t)P.nsplit;
  const uint32_t hfe = (uint32_t)P.hfe;
  const bool vhold = P.vf == 2;
  const uint64_t rstep = (uint64_t)(uint32_t)P.row_step * P.in_row_bytes;
  // the byte range of the input this launch may touch (see span_load)
  const uintptr_t lim_lo = reinterpret_cast<uintptr_t>(P.in) + (uint64_t)((uint32_t)P.row0 * (uint32_t)P.row_step) * P.in_row_bytes;
  const uintptr_t lim_hi = reinterpret_cast<uintptr_t>(P.in) + (uint64_t)(P.n_frames - 1u) * P.in_frame_bytes +
                           (uint64_t)((uint32_t)(P.row0 + P.band_rows - 1) * (uint32_t)P.row_step) * P.in_row_bytes +
                           ((uint32_t)P.Wo - 1u) * pxb + ipb;
  // quantiser masks / shifts
  const uint32_t my = P.qmask & 0xFFu, mcb = (P.qmask >> 8) & 0xFFu, mcr = (P.qmask >> 16) & 0xFFu;
  const uint32_t qm0 = my | (mcb << 8) | (mcr << 16) | (my << 24);
  const uint32_t qm1 = mcb | (mcr << 8) | (my << 16) | (mcb << 24);
  const uint32_t qm2 = mcr | (my << 8) | (mcb << 16) | (mcr << 24);
  const int shy = 8 + P.sy, shb = 8 + P.scb, shr = 8 + P.scr, ly = P.cb_bits + P.cr_bits, lb = P.cr_bits;
  const uint32_t vs_sh = P.planar_vs == 2 ? 1u : 0u, hs_sh = P.planar_hs == 4 ? 2u : (P.planar_hs == 2 ? 1u : 0u);

  // Input rows of a tile land in stage s at  in(s) + j * rs_mul + ((a0 + j * rs_add) & 15),  a0 = src0 & 15.
  const uint32_t rs_mul = P.in_dense ? P.in_row_bytes : P.stage_stride;
  const uint32_t rs_add = P.in_dense ? 0u : ((uint32_t)rstep & 15u);

  // Hands the tile's input to the copy engine (cp.async) and fetches the pixel a held row replays into registers;
  // nothing here waits for memory.
  uint32_t h0 = 0, h1 = 0, h2 = 0, hvalid = 0;
  auto issue = [&](const FlexTile& T, uint32_t s) {
    const uint32_t in_s = sbase + s * in_stage;
    if (P.in_dense) {          // consecutive rows are contiguous in memory: one span
      span_load(in_s, T.src0, (T.nrows - 1u) * P.in_row_bytes + T.len_in, tid, NT, pol, lim_lo, lim_hi);
    } else {
      for_each_span(T.nrows, T.len_in, [&](uint32_t j, uint32_t t, uint32_t n) {
        span_load(in_s + j * rs_mul, T.src0 + (uint64_t)j * rstep, T.len_in, t, n, pol, lim_lo, lim_hi);
      });
    }
    hvalid = 0;
    if (vhold && tid < T.nrows) {   // the pixel whose chroma a held row replays (ChromaSubsampler.scala:62-65)
      const uint32_t ro = T.ro0 + tid;
      const uint8_t* hp = nullptr;
      if (!P.case_b) {
        if (f == 1 && (ro & 1u)) hp = T.frame + (uint64_t)(ro - 1u) * P.in_row_bytes + (uint32_t)P.last_sample_col * ipb;
      } else {
        const uint32_t line = ro / f;        // W == f * Wo: one counter line spans f output rows
        if (line & 1u) {
          const uint32_t srow = (line - 1u) * f + P.caseb_row_add;
          hp = T.frame + (uint64_t)(srow * (uint32_t)P.row_step) * P.in_row_bytes + P.caseb_col_bytes;
        }
      }
      if (hp) { h0 = ldg8_now(hp); h1 = ldg8_now(hp + 1); h2 = ldg8_now(hp + 2); hvalid = 0x80000000u; }
    }
    cp_async_commit();
  };
  auto publish_held = [&](uint32_t s) {      // first use of the registers `issue` filled
    if (vhold && tid < (uint32_t)P.tile_rows) sts32(held_base + (s * (uint32_t)kMaxTileRows + tid) * 4u, hvalid ? (hvalid | h0 | (h1 << 8) | (h2 << 16)) : 0u);
  };

  uint32_t tile = blockIdx.x;
  if (tile >= P.n_tiles) return;
  FlexTile T = flex_tile(P, tile);
  issue(T, 0);
  publish_held(0);
  for (uint32_t it = 0;; ++it) {
    const uint32_t s = it & 1u;
    const uint32_t next = tile + gridDim.x;
    const bool has_next = next < P.n_tiles;
    FlexTile Tn;
    if (has_next) {
      Tn = flex_tile(P, next);
      issue(Tn, s ^ 1u);       // stage s^1 was last read two barriers ago
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    // ---- compute -------------------------------------------------------------------------------------------
    const uint32_t in_s = sbase + s * in_stage, held_s = held_base + s * (uint32_t)kMaxTileRows * 4u;
    const uint32_t a0 = (uint32_t)reinterpret_cast<uintptr_t>(T.src0) & 15u;
    const uint32_t gpr = (T.ncols + 3u) >> 2;                  // granules per row
    const uint32_t n_gran = T.nrows * gpr, srow = gpr * kUnit;
    const uint32_t gpr_magic = gpr > 1u ? 0xFFFFFFFFu / gpr + 1u : 0u;   // q / gpr == umulhi(q, magic) for q < 65536
    // PLANAR: chroma rows of the tile are the output rows with ro % vs == 0
    const uint32_t c_first = (T.ro0 + (1u << vs_sh) - 1u) >> vs_sh;
    const uint32_t c_last1 = ((T.ro0 + T.nrows - 1u) >> vs_sh) + 1u;
    const uint32_t nrc = (FMT == KF_PLANAR && c_last1 > c_first) ? c_last1 - c_first : 0u;
    const uint32_t ccols = (T.npx + (1u << hs_sh) - 1u) >> hs_sh;
    const uint32_t cb_s = out_s + T.nrows * srow, cr_s = cb_s + nrc * ccols;
    const uint32_t last_px = T.npx - 1u;
    for (uint32_t q = tid; q < n_gran; q += NT) {
      const uint32_t row = gpr > 1u ? __umulhi(q, gpr_magic) : q, g = q - row * gpr;
      const uint32_t rs = in_s + row * rs_mul + ((a0 + row * rs_add) & 15u);
      const uint32_t c = g * 4u;
      uint32_t p[4], dy[4], xb[4], xr[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) p[j] = lds_px(rs + min(c + j, last_px) * pxb);
#pragma unroll
      for (int j = 0; j < 4; ++j) dy[j] = fwd_y16(p[j], P.coef_y);
      const uint32_t hv = vhold ? lds32(held_s + row * 4u) : 0u;
      if (hv) {
        const uint32_t hb = fwd_nc16<TRUNC>(hv & 0x00FFFFFFu, P.coef_ncb), hr = fwd_nc16<TRUNC>(hv & 0x00FFFFFFu, P.coef_ncr);
#pragma unroll
        for (int j = 0; j < 4; ++j) { xb[j] = hb; xr[j] = hr; }
      } else {
        // sample where j % hfe == 0, hold in between (ChromaSubsampler.scala:57-65)
        xb[0] = fwd_nc16<TRUNC>(p[0], P.coef_ncb); xr[0] = fwd_nc16<TRUNC>(p[0], P.coef_ncr);
        if (hfe == 1) { xb[1] = fwd_nc16<TRUNC>(p[1], P.coef_ncb); xr[1] = fwd_nc16<TRUNC>(p[1], P.coef_ncr); }
        else { xb[1] = xb[0]; xr[1] = xr[0]; }
        if (hfe <= 2) { xb[2] = fwd_nc16<TRUNC>(p[2], P.coef_ncb); xr[2] = fwd_nc16<TRUNC>(p[2], P.coef_ncr); }
        else { xb[2] = xb[0]; xr[2] = xr[0]; }
        if (hfe == 1) { xb[3] = fwd_nc16<TRUNC>(p[3], P.coef_ncb); xr[3] = fwd_nc16<TRUNC>(p[3], P.coef_ncr); }
        else { xb[3] = xb[2]; xr[3] = xr[2]; }
      }
      const uint32_t so = out_s + row * srow + g * kUnit;
      if (FMT == KF_YCC888) {
        // byte 1 of dy is Y, byte 1 of xb/xr is ~Cb/~Cr: gather with PRMT, flip and quantise per word
        uint32_t t, u;
        t = __byte_perm(dy[0], xb[0], 0x0051); u = __byte_perm(xr[0], dy[1], 0x0051);
        sts32(so, (__byte_perm(t, u, 0x5410) ^ 0x00FFFF00u) & qm0);
        t = __byte_perm(xb[1], xr[1], 0x0051); u = __byte_perm(dy[2], xb[2], 0x0051);
        sts32(so + 4, (__byte_perm(t, u, 0x5410) ^ 0xFF00FFFFu) & qm1);
        t = __byte_perm(xr[2], dy[3], 0x0051); u = __byte_perm(xb[3], xr[3], 0x0051);
        sts32(so + 8, (__byte_perm(t, u, 0x5410) ^ 0xFFFF00FFu) & qm2);
      } else if (FMT == KF_RGB888) {
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
          v[j] = inverse_rgb((int)((dy[j] >> 8) & my), (int)((255u - (xb[j] >> 8)) & mcb), (int)((255u - (xr[j] >> 8)) & mcr));
        sts32(so, v[0] | (v[1] << 24));
        sts32(so + 4, (v[1] >> 8) | (v[2] << 16));
        sts32(so + 8, (v[2] >> 16) | (v[3] << 8));
      } else if (FMT == KF_PLANAR) {
        const uint32_t my4 = my * 0x01010101u;
        sts32(so, __byte_perm(__byte_perm(dy[0], dy[1], 0x0051), __byte_perm(dy[2], dy[3], 0x0051), 0x5410) & my4);
        if (!hv) {               // a sampled line: its sample points go to the chroma planes (hs == hfe here)
          const uint32_t crow = (((T.ro0 + row) >> vs_sh) - c_first) * ccols + (c >> hs_sh);
          if (hfe == 1) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (c + j <= last_px) { sts8(cb_s + crow + j, (~(xb[j] >> 8)) & mcb); sts8(cr_s + crow + j, (~(xr[j] >> 8)) & mcr); }
          } else if (hfe == 2) {
            sts8(cb_s + crow, (~(xb[0] >> 8)) & mcb); sts8(cr_s + crow, (~(xr[0] >> 8)) & mcr);
            if (c + 2 <= last_px) { sts8(cb_s + crow + 1, (~(xb[2] >> 8)) & mcb); sts8(cr_s + crow + 1, (~(xr[2] >> 8)) & mcr); }
          } else {
            sts8(cb_s + crow, (~(xb[0] >> 8)) & mcb); sts8(cr_s + crow, (~(xr[0] >> 8)) & mcr);
          }
        }
      } else {
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = ((dy[j] >> shy) << ly) | (((xb[j] ^ 0xFFFFu) >> shb) << lb) | ((xr[j] ^ 0xFFFFu) >> shr);
          if (c + j > last_px) v[j] = 0u;       // the row's zero pad slots
        }
        if (FMT == KF_SLOT32) { sts32(so, v[0]); sts32(so + 4, v[1]); sts32(so + 8, v[2]); sts32(so + 12, v[3]); }
        else if (FMT == KF_SLOT16) { sts32(so, v[0] | (v[1] << 16)); sts32(so + 4, v[2] | (v[3] << 16)); }
        else sts32(so, v[0] | (v[1] << 8) | (v[2] << 16) | (v[3] << 24));
      }
    }
    __syncthreads();

    // ---- store ---------------------------------------------------------------------------------------------
    uint8_t* fout = P.out + (uint64_t)T.k * P.out_frame_bytes;
    uint8_t* obase = fout + (uint64_t)T.ro0 * P.out_row_bytes + (uint64_t)T.col0 * kOpx;
    const uint32_t row_out = T.ncols * kOpx;
    if (P.out_dense && nsplit == 1 && row_out == srow) {      // whole dense rows, packed in the staging area: one span
      span_store(obase, out_s, T.nrows * row_out, tid, NT);
    } else {
      for_each_span(T.nrows, row_out, [&](uint32_t j, uint32_t t, uint32_t n) {
        span_store(obase + (uint64_t)j * P.out_row_bytes, out_s + j * srow, row_out, t, n);
      });
    }
    if (FMT == KF_PLANAR && nrc) {
      const uint64_t coff = (uint64_t)c_first * (uint32_t)P.planar_cw + (T.col0 >> hs_sh);
      uint8_t* cb_g = fout + P.planar_cb_off + coff;
      uint8_t* cr_g = fout + P.planar_cr_off + coff;
      if (nsplit == 1) {                                       // ccols == planar_cw: chroma rows are contiguous
        span_store(cb_g, cb_s, nrc * ccols, tid, NT);
        span_store(cr_g, cr_s, nrc * ccols, tid, NT);
      } else {
        for_each_span(nrc, ccols, [&](uint32_t j, uint32_t t, uint32_t n) {
          span_store(cb_g + (uint64_t)j * (uint32_t)P.planar_cw, cb_s + j * ccols, ccols, t, n);
          span_store(cr_g + (uint64_t)j * (uint32_t)P.planar_cw, cr_s + j * ccols, ccols, t, n);
        });
      }
    }
    if (!has_next) break;
    // The held words of the next tile: stage s^1's were last read in the compute phase two barriers back; the next
    // compute phase starts behind the next barrier.  The staging area is rewritten only behind that barrier too.
    publish_held(s ^ 1u);
    T = Tn;
    tile = next;
  }
}

// ---- planning and dispatch ---------------------------------------------------------------------------------
bool plan_flex_kernel(KPlan& k, int sm_count, size_t max_smem_optin) {
  if (k.average && k.f > 1) return false;                       // AVERAGE extension: csic_pool_kernel / generic
  if (k.band_rows <= 0 || k.n_frames == 0) return false;
  const uint32_t ipb = (uint32_t)k.in_px_bytes, f = (uint32_t)k.f, pxb = f * ipb;
  if (k.case_b) {
    // spatial before chroma: a counter line must be exactly f output rows, each starting on a sample column
    if (k.W % k.f != 0 || k.Wo % k.hf != 0) return false;
    k.caseb_row_add = (uint32_t)k.last_sample_col / (uint32_t)k.Wo;
    k.caseb_col_bytes = ((uint32_t)k.last_sample_col % (uint32_t)k.Wo) * pxb;
    k.hfe = k.hf;
  } else {
    k.hfe = std::max(1, k.hf / k.f);
  }
  const bool planar = k.kformat == KF_PLANAR;
  const uint32_t unit = (k.kformat <= KF_RGB888) ? 12u : (planar ? 4u : 4u * (uint32_t)k.slot_bytes);
  const uint32_t S = (uint32_t)k.slots_per_row, Wo = (uint32_t)k.Wo;
  const uint32_t tile_px_max = std::max(16u, (kFlexTileBytes / pxb) & ~15u);
  k.tile_px = (int32_t)std::min(tile_px_max, (S + 15u) & ~15u);
  k.nsplit = (int32_t)((S + (uint32_t)k.tile_px - 1u) / (uint32_t)k.tile_px);
  const uint32_t len_in_max = ((std::min((uint32_t)k.tile_px, Wo) - 1u) * f + 1u) * ipb;
  // consecutive processed rows contiguous in memory?  (dense rows, every stored row is read)
  const bool contiguous = k.nsplit == 1 && k.row_step == 1 && k.in_row_bytes == (uint32_t)k.W * ipb;
  int rows = 1;
  if (k.nsplit == 1) {
    const uint32_t per_row = contiguous ? k.in_row_bytes : len_in_max;
    rows = (int)std::min<uint32_t>((uint32_t)kMaxTileRows, std::max<uint32_t>(1u, kFlexTileBytes / std::max(1u, per_row)));
    rows = std::min(rows, k.band_rows);
    auto tiles_for = [&](int r) { return (uint64_t)k.n_frames * (uint64_t)((k.band_rows + r - 1) / r); };
    while (rows > 1 && tiles_for(rows) < (uint64_t)sm_count * 8u) rows = (rows + 1) / 2;
  }
  k.tile_rows = rows;
  k.in_dense = (contiguous && rows > 1) ? 1 : 0;
  k.tiles_per_band = (uint32_t)((k.band_rows + rows - 1) / rows);
  const uint64_t n_tiles = (uint64_t)k.n_frames * (uint64_t)k.tiles_per_band * (uint64_t)k.nsplit;
  if (n_tiles >= (1ull << 31)) return false;
  k.n_tiles = (uint32_t)n_tiles;
  const uint32_t dense_out_row = planar ? Wo : S * (unit / 4u);
  k.out_dense = k.out_row_bytes == dense_out_row ? 1 : 0;

  auto up16 = [](uint32_t v) { return (v + 15u) & ~15u; };
  k.stage_stride = up16((k.in_dense ? std::max(len_in_max, k.in_row_bytes) : len_in_max) + 15u) + 16u;
  const uint32_t in_bytes = 2u * ((uint32_t)rows * k.stage_stride + 32u);   // two stages; + slack: lds_px reads one word ahead
  uint32_t stage_bytes = (uint32_t)rows * ((uint32_t)k.tile_px / 4u) * unit;
  if (planar) stage_bytes += 2u * (uint32_t)(rows / std::max(1, k.planar_vs) + 1) * (uint32_t)k.tile_px;
  k.out_buf_off = in_bytes;
  k.out_buf_stride = up16(stage_bytes + 32u);                               // + slack: span_store reads one word ahead
  k.meta_off = k.out_buf_off + k.out_buf_stride;
  k.smem_bytes = k.meta_off + 2u * (uint32_t)kMaxTileRows * 4u;           // held words of both stages
  if (k.smem_bytes > max_smem_optin) return false;
  k.block_threads = kFlexThreads;
  k.ctas_per_sm = (int32_t)std::max<uint32_t>(1u, std::min<uint32_t>(std::min<uint32_t>(8u, 2048u / kFlexThreads),
                                                                      227u * 1024u / (k.smem_bytes + 1024u)));
  return true;
}

namespace {
template <int FMT>
int launch_flex_fmt(const KPlan& k, unsigned grid, cudaStream_t st) {
  if (k.trunc) csic_flex_kernel<FMT, true><<<grid, kFlexThreads, k.smem_bytes, st>>>(k);
  else csic_flex_kernel<FMT, false><<<grid, kFlexThreads, k.smem_bytes, st>>>(k);
  return (int)cudaGetLastError();
}
template <int FMT>
cudaError_t flex_attr(size_t b) {
  cudaError_t e = cudaFuncSetAttribute(csic_flex_kernel<FMT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(csic_flex_kernel<FMT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b);
}
}  // namespace

int launch_flex(const KPlan& k, int sm_count, int force_ctas_per_sm, void* stream) {
  const int per_sm = force_ctas_per_sm > 0 ? force_ctas_per_sm : std::max(1, k.ctas_per_sm);
  const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)k.n_tiles, (uint64_t)sm_count * (uint64_t)per_sm);
  cudaStream_t st = (cudaStream_t)stream;
  switch (k.kformat) {
    case KF_YCC888: return launch_flex_fmt<KF_YCC888>(k, grid, st);
    case KF_RGB888: return launch_flex_fmt<KF_RGB888>(k, grid, st);
    case KF_SLOT8: return launch_flex_fmt<KF_SLOT8>(k, grid, st);
    case KF_SLOT16: return launch_flex_fmt<KF_SLOT16>(k, grid, st);
    case KF_PLANAR: return launch_flex_fmt<KF_PLANAR>(k, grid, st);
    default: return launch_flex_fmt<KF_SLOT32>(k, grid, st);
  }
}

int flex_set_attributes(size_t max_smem_optin) {
  cudaError_t e;
  if ((e = flex_attr<KF_YCC888>(max_smem_optin)) != cudaSuccess) return (int)e;
  if ((e = flex_attr<KF_RGB888>(max_smem_optin)) != cudaSuccess) return (int)e;
  if ((e = flex_attr<KF_SLOT8>(max_smem_optin)) != cudaSuccess) return (int)e;
  if ((e = flex_attr<KF_SLOT16>(max_smem_optin)) != cudaSuccess) return (int)e;
  if ((e = flex_attr<KF_SLOT32>(max_smem_optin)) != cudaSuccess) return (int)e;
  if ((e = flex_attr<KF_PLANAR>(max_smem_optin)) != cudaSuccess) return (int)e;
  return (int)cudaSuccess;
}

}  // namespace csic
