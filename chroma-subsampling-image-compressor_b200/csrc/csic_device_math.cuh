// Device-side integer arithmetic shared by both kernels (sm_100a).
#ifndef CSIC_DEVICE_MATH_CUH_
#define CSIC_DEVICE_MATH_CUH_
#include <cuda_runtime.h>
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

// Coefficient words for pixels stored R,G,B,(x).  For B,G,R,(x) input the host swaps bytes 0 and 2
// (KPlan::coef_*), so the kernels never permute pixel bytes.
constexpr uint32_t kCoefY = 0x001D964Du;     //  77, 150,  29, 0   (u8)
constexpr uint32_t kCoefNCb = 0x0080552Bu;   //  43,  85,-128, 0   (s8)  == -cb row
constexpr uint32_t kCoefNCr = 0x00156B80u;   //-128, 107,  21, 0   (s8)  == -cr row

// byte 1 of the result is Y
__device__ __forceinline__ uint32_t fwd_y16(uint32_t p, uint32_t coef_y) { return dp4a_uu(p, coef_y, 128u); }
// byte 1 of the result is ~Cb / ~Cr; result < 65536
template <bool TRUNC>
__device__ __forceinline__ uint32_t fwd_nc16(uint32_t p, uint32_t coef) {
  int32_t x = max(dp4a_us(p, coef, 32639), 0);
  if (TRUNC) x -= (x >> 15) * 255;
  return (uint32_t)x;
}


// YCbCrUtils.ycbcr2rgb with the -128 offsets folded into the constants.  Returns R | G<<8 | B<<16.
__device__ __forceinline__ uint32_t inverse_rgb(int y, int cb, int cr) {
  const int c = 298 * y;
  // clamp the 24.8 values to [0, 0xFFFF]: byte 1 is then the clamped channel -- no shift, PRMT picks the bytes
  const uint32_t r = (uint32_t)__vimin_s32_relu(c + 409 * cr - 52224, 0xFFFF);
  const uint32_t g = (uint32_t)__vimin_s32_relu(c - 100 * cb - 208 * cr + 39552, 0xFFFF);
  const uint32_t b = (uint32_t)__vimin_s32_relu(c + 516 * cb - 65920, 0xFFFF);
  return __byte_perm(__byte_perm(r, g, 0x3351), b, 0x3510);     // R | G << 8 | B << 16
}

// The same transform applied straight to the forward results: dy = fwd_y16 (byte 1 = Y), xb / xr = fwd_nc16 (byte 1 =
// ~Cb / ~Cr); my8 / mcb8 / mcr8 are the quantiser keep-masks shifted left by 8.  The channel values stay shifted left
// by 8 (one LOP3 each instead of shift + subtract + mask), the constants are scaled by 256 and the final shift is 16:
// floor(256 N / 65536) == floor(N / 256), the same integers with eight instructions fewer per pixel.
__device__ __forceinline__ uint32_t inverse_rgb_raw(uint32_t dy, uint32_t xb, uint32_t xr, uint32_t my8, uint32_t mcb8, uint32_t mcr8) {
  const int y8 = (int)(dy & my8), cb8 = (int)((xb ^ 0xFFFFu) & mcb8), cr8 = (int)((xr ^ 0xFFFFu) & mcr8);
  const int c = 298 * y8;
  // clamp the 16.16 values to [0, 0xFFFFFF]: byte 2 is then the clamped channel -- no shift, PRMT picks the bytes
  const uint32_t r = (uint32_t)__vimin_s32_relu(c + 409 * cr8 - 52224 * 256, 0x00FFFFFF);
  const uint32_t g = (uint32_t)__vimin_s32_relu(c - 100 * cb8 - 208 * cr8 + 39552 * 256, 0x00FFFFFF);
  const uint32_t b = (uint32_t)__vimin_s32_relu(c + 516 * cb8 - 65920 * 256, 0x00FFFFFF);
  return __byte_perm(__byte_perm(r, g, 0x3362), b, 0x3610);     // R | G << 8 | B << 16
}

// The same transform split where the subsampled pipeline lets it be shared: everything that depends on the chroma pair
// only (three multiply-adds with the -128 offsets and the rounding constants folded in) is computed once per chroma
// SAMPLE -- every second / fourth pixel under 4:2:x / 4:1:x, once per row on a held line -- and a pixel costs one mask,
// one multiply and three add-clamp instructions.  (The compiler cannot find this sharing by itself once held and
// sampled rows have merged in the control flow: 13 instructions per pixel became 7.)
struct InvChroma { int tr, tg, tb; };
__device__ __forceinline__ InvChroma inv_chroma_terms(uint32_t xb, uint32_t xr, uint32_t mcb8, uint32_t mcr8) {
  const int cb8 = (int)((xb ^ 0xFFFFu) & mcb8), cr8 = (int)((xr ^ 0xFFFFu) & mcr8);
  InvChroma t;
  t.tr = 409 * cr8 - 52224 * 256;
  t.tg = -208 * cr8 + (-100 * cb8 + 39552 * 256);
  t.tb = 516 * cb8 - 65920 * 256;
  return t;
}
// Four pixels -> the twelve RGB bytes of a granule (w0 = R0 G0 B0 R1, w1 = G1 B1 R2 G2, w2 = B2 R3 G3 B3).
// Each channel is one multiply-add in 16.16 fixed point; its UPPER half word is floor(value), a signed 16-bit number in
// [-560, 816].  Two channels are then packed into one register by a PRMT (upper halves), clamped to [0, 255] TOGETHER by
// one VIMNMX.S16x2.RELU (min against 0x00FF00FF, relu), and a third PRMT per output word gathers the four low bytes:
// 12 IMAD (FMA pipe) + 9 PRMT + 6 VIMNMX (ALU pipe) per granule instead of 12 + 11 + 12 with one clamp per channel.
__device__ __forceinline__ void inv_granule(const uint32_t (&dy)[4], uint32_t my8, const InvChroma& t0, const InvChroma& t1,
                                            const InvChroma& t2, const InvChroma& t3, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
  const int y0 = 298 * (int)(dy[0] & my8), y1 = 298 * (int)(dy[1] & my8), y2 = 298 * (int)(dy[2] & my8), y3 = 298 * (int)(dy[3] & my8);
  auto pair = [](int a, int b) {      // clamp(a >> 16) in byte 0, clamp(b >> 16) in byte 2
    return __vimin_s16x2_relu(__byte_perm((uint32_t)a, (uint32_t)b, 0x7632), 0x00FF00FFu);
  };
  const uint32_t p0 = pair(y0 + t0.tr, y0 + t0.tg), p1 = pair(y0 + t0.tb, y1 + t1.tr);
  const uint32_t p2 = pair(y1 + t1.tg, y1 + t1.tb), p3 = pair(y2 + t2.tr, y2 + t2.tg);
  const uint32_t p4 = pair(y2 + t2.tb, y3 + t3.tr), p5 = pair(y3 + t3.tg, y3 + t3.tb);
  w0 = __byte_perm(p0, p1, 0x6420);
  w1 = __byte_perm(p2, p3, 0x6420);
  w2 = __byte_perm(p4, p5, 0x6420);
}

// The PLANAR decoder's flavour: four pixels given as packed bytes (yw = Y0..Y3, cbw / crw = the chroma each pixel
// replays, already expanded by the hold rule) -> the twelve RGB bytes.  share = log2 of how many consecutive pixels share
// a chroma pair (0, 1 or 2; warp-uniform): the chroma terms are computed once per pair.  Same integers as inverse_rgb.
__device__ __forceinline__ void decode_rgb_granule(uint32_t yw, uint32_t cbw, uint32_t crw, uint32_t share, uint32_t& w0,
                                                   uint32_t& w1, uint32_t& w2) {
  auto terms = [&](uint32_t sel) {          // sel: PRMT selector placing the pixel's byte into byte 1, zeros elsewhere
    const int cb8 = (int)__byte_perm(cbw, 0, sel), cr8 = (int)__byte_perm(crw, 0, sel);
    InvChroma t;
    t.tr = 409 * cr8 - 52224 * 256;
    t.tg = -208 * cr8 + (-100 * cb8 + 39552 * 256);
    t.tb = 516 * cb8 - 65920 * 256;
    return t;
  };
  const uint32_t dy[4] = {__byte_perm(yw, 0, 0x4404), __byte_perm(yw, 0, 0x4414), __byte_perm(yw, 0, 0x4424), __byte_perm(yw, 0, 0x4434)};
  const InvChroma t0 = terms(0x4404);
  InvChroma t1 = t0, t2, t3;
  if (share == 0u) t1 = terms(0x4414);
  if (share <= 1u) t2 = terms(0x4424); else t2 = t0;
  t3 = t2;
  if (share == 0u) t3 = terms(0x4434);
  inv_granule(dy, 0xFF00u, t0, t1, t2, t3, w0, w1, w2);
}

// One decoded granule: interleaved Y,Cb,Cr bytes (eight PRMTs) or RGB (decode_rgb_granule).
__device__ __forceinline__ void decode_granule(uint32_t yw, uint32_t cbw, uint32_t crw, uint32_t share, int to_rgb, uint32_t& w0,
                                               uint32_t& w1, uint32_t& w2) {
  if (to_rgb) {
    decode_rgb_granule(yw, cbw, crw, share, w0, w1, w2);
  } else {        // Y0 Cb0 Cr0 Y1 | Cb1 Cr1 Y2 Cb2 | Cr2 Y3 Cb3 Cr3
    w0 = __byte_perm(__byte_perm(yw, cbw, 0x1040), crw, 0x3410);
    w1 = __byte_perm(__byte_perm(cbw, crw, 0x0051), __byte_perm(yw, cbw, 0x0062), 0x5410);
    w2 = __byte_perm(__byte_perm(crw, yw, 0x0072), __byte_perm(cbw, crw, 0x0073), 0x5410);
  }
}

// The same with the chroma sample of each pixel named by its byte index (c0..c3) inside cbw / crw, so that a caller whose
// hold pattern is known at compile time (after unrolling) needs no expanded words: the terms of a repeated index are
// reused, the YCC interleave is eight PRMTs whatever the pattern.
__device__ __forceinline__ InvChroma dec_terms(uint32_t cbw, uint32_t crw, uint32_t idx) {
  const uint32_t sel = 0x4404u | (idx << 4);           // the sample's byte into byte 1, zeros elsewhere
  const int cb8 = (int)__byte_perm(cbw, 0, sel), cr8 = (int)__byte_perm(crw, 0, sel);
  InvChroma t;
  t.tr = 409 * cr8 - 52224 * 256;
  t.tg = -208 * cr8 + (-100 * cb8 + 39552 * 256);
  t.tb = 516 * cb8 - 65920 * 256;
  return t;
}
__device__ __forceinline__ void rgb_granule_terms(uint32_t yw, const InvChroma& t0, const InvChroma& t1, const InvChroma& t2,
                                                  const InvChroma& t3, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
  const uint32_t dy[4] = {__byte_perm(yw, 0, 0x4404), __byte_perm(yw, 0, 0x4414), __byte_perm(yw, 0, 0x4424), __byte_perm(yw, 0, 0x4434)};
  inv_granule(dy, 0xFFFFFFFFu, t0, t1, t2, t3, w0, w1, w2);
}
template <bool RGB>
__device__ __forceinline__ void decode_granule_idx(uint32_t yw, uint32_t cbw, uint32_t crw, uint32_t c0, uint32_t c1, uint32_t c2,
                                                   uint32_t c3, uint32_t& w0, uint32_t& w1, uint32_t& w2) {
  if (RGB) {
    const InvChroma t0 = dec_terms(cbw, crw, c0);
    const InvChroma t1 = c1 == c0 ? t0 : dec_terms(cbw, crw, c1);
    const InvChroma t2 = c2 == c1 ? t1 : dec_terms(cbw, crw, c2);
    const InvChroma t3 = c3 == c2 ? t2 : dec_terms(cbw, crw, c3);
    rgb_granule_terms(yw, t0, t1, t2, t3, w0, w1, w2);
  } else {        // Y0 Cb0 Cr0 Y1 | Cb1 Cr1 Y2 Cb2 | Cr2 Y3 Cb3 Cr3
    w0 = __byte_perm(__byte_perm(yw, cbw, 0x1000u | ((4u + c0) << 4)), crw, 0x3010u | ((4u + c0) << 8));
    w1 = __byte_perm(__byte_perm(cbw, crw, c1 | ((4u + c1) << 4)), __byte_perm(yw, cbw, 0x2u | ((4u + c2) << 4)), 0x5410);
    w2 = __byte_perm(__byte_perm(crw, yw, c2 | 0x70u), __byte_perm(cbw, crw, c3 | ((4u + c3) << 4)), 0x5410);
  }
}

// Four 24-bit pixels (R | G << 8 | B << 16) -> the twelve bytes of a granule, three PRMTs.
__device__ __forceinline__ void pack_rgb_granule(const uint32_t (&v)[4], uint32_t& w0, uint32_t& w1, uint32_t& w2) {
  w0 = __byte_perm(v[0], v[1], 0x4210);     // R0 G0 B0 R1
  w1 = __byte_perm(v[1], v[2], 0x5421);     // G1 B1 R2 G2
  w2 = __byte_perm(v[2], v[3], 0x6542);     // B2 R3 G3 B3
}

}  // namespace csic
#endif  // CSIC_DEVICE_MATH_CUH_
