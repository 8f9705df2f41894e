// Internal declarations shared by the host half (csic_params.cpp, csic_api.cu) and the kernels.
#ifndef CSIC_INTERNAL_H_
#define CSIC_INTERNAL_H_

#include <cstddef>
#include <cstdint>

#include "../../include/csic.h"

namespace csic {

struct Geometry {
  int32_t out_w, out_h;
  size_t in_row_bytes, in_frame_bytes;
  size_t out_row_bytes, out_frame_bytes;
  int32_t out_px_bytes;      // 3 for YCC888/RGB888, slot bytes (1/2/4) for bundles
  int32_t in_px_bytes;       // 3 (RGB24) or 4 (RGBA32 / BGRA32, fourth byte ignored)
  bool chroma_first;         // ChromaSubsampling precedes SpatialSampling in op1..op3
  bool quant_first;          // ColorQuantization precedes SpatialSampling (only matters for AVERAGE)
  int32_t hf, vf;            // ChromaSubsampler.scala:26-27
  int32_t planar_hs, planar_vs, planar_cw, planar_ch;   // CSIC_OUT_PLANAR chroma decimation / plane size
};

int slot_bits(const csic_params& p);
Geometry geometry(const csic_params& p);

// Kernel-side output format after resolving the bundle slot width.
enum KFormat : int32_t { KF_YCC888 = 0, KF_RGB888 = 1, KF_SLOT8 = 2, KF_SLOT16 = 3, KF_SLOT32 = 4, KF_PLANAR = 5 };

// Everything a kernel needs, passed by value (lives in the constant bank).
struct KPlan {
  const uint8_t* in;
  uint8_t* out;
  uint64_t in_frame_bytes, out_frame_bytes;
  uint32_t in_row_bytes, out_row_bytes;  // row PITCHES in bytes (== dense row size unless the caller pitched the buffers)
  int32_t W, H, Wo, Ho;
  int32_t Wp;                            // row kernel: processing width in output pixels (Wo rounded up to 16)
  int32_t in_dense, out_dense, ragged;   // rows of a tile contiguous in memory (no pitch gap); Wp != Wo
  int32_t in_px_bytes;                   // 3 or 4
  uint32_t coef_y, coef_ncb, coef_ncr;   // dp4a coefficient words in the byte order of the input pixels
  int32_t f;
  int32_t hf, vf, last_sample_col;       // ((W-1)/hf)*hf : last chroma sample column of a line
  int32_t case_b;                        // spatial before chroma with f > 1 (misaligned counters)
  uint32_t caseb_row_add, caseb_col_bytes; // case B: held pixel = decimated-stream element (line-1)*W + last_sample_col
  int32_t quant_first, trunc, average;
  int32_t kformat, slot_bytes, slots_per_row;
  int32_t planar_hs, planar_vs, planar_cw, planar_ch;   // PLANAR: chroma decimation (output px) and plane size
  uint64_t planar_cb_off, planar_cr_off;                 // byte offsets of the Cb / Cr planes inside an output frame
  int32_t sy, scb, scr;                  // 8 - target bits
  int32_t cb_bits, cr_bits;
  int32_t row0, band_rows;               // output rows [row0, row0 + band_rows) of every frame
  int32_t compact;                       // input holds only the rows a DECIMATE pipeline reads (every f-th), densely
  int32_t row_step;                      // input rows between consecutive output rows: f, or 1 when compact
  uint32_t n_frames;
  // ---- TMA row kernel only ----
  int32_t hfe;                           // chroma hold width in *output* pixels inside a 4-pixel granule
  int32_t nsplit, tile_px;               // segments per output row, output pixels per row segment
  int32_t tile_rows;                     // output rows per tile (1 when nsplit > 1)
  uint32_t tiles_per_band;
  uint32_t tile_in_bytes, tile_out_bytes; // bytes of ONE row segment (input / output)
  uint32_t n_tiles;
  int32_t stages;
  int32_t block_threads;
  int32_t ctas_per_sm;
  uint32_t stage_stride, out_buf_off, out_buf_stride, meta_off, bar_off, smem_bytes;
  uint32_t qmask;                        // my | mcb<<8 | mcr<<16 (per-channel keep masks)
  int32_t tall_ho;                       // flex kernel: != 0 -> the batch is addressed as ONE image of n * Ho rows; rows per original frame
};

// Fills the TMA-row-kernel fields of `k`; returns false when the configuration is not eligible
// (then the generic kernel runs).  `sm_count`/`max_smem` come from the device.
bool plan_rows_kernel(KPlan& k, int sm_count, size_t max_smem_optin, int force_stages, uint32_t force_tile_bytes);

// Both return a cudaError_t as int.
int launch_generic(const KPlan& k, int sm_count, void* stream);
int launch_expand_planar(const KPlan& k, const uint8_t* planar, uint8_t* out, int to_rgb, int sm_count, size_t max_smem_optin, void* stream);
// TMA-staged decoder for any width / alignment (csic_decode_kernel.cu); -1 = not eligible, else a cudaError_t
int launch_decode_tma(const KPlan& k, const uint8_t* planar, uint8_t* out, int to_rgb, int sm_count, size_t max_smem_optin, void* stream);
int decode_set_attributes(size_t max_smem_optin);
int launch_rows(const KPlan& k, int sm_count, int force_ctas_per_sm, void* stream);
constexpr int kDefaultBlockThreads = 256;   // consumer threads; one producer warp is added at launch
constexpr int kMaxConsumerThreads = 512;
int rows_kernel_set_attributes(size_t max_smem_optin);

// AVERAGE extension: TMA-staged pooling kernel (csic_pool_kernel.cu), chroma-first orders only.
bool plan_pool_kernel(KPlan& k, int sm_count, size_t max_smem_optin);
int launch_pool(const KPlan& k, int sm_count, int force_ctas_per_sm, void* stream);
template <int F> int launch_pool_factor(const KPlan& k, unsigned grid, void* stream);
template <int F> int pool_set_attributes_factor(size_t max_smem_optin);

// Any width / pitch / base alignment (csic_flex_kernel.cu): DECIMATE pipelines the TMA kernels' 16-byte rules exclude.
bool plan_flex_kernel(KPlan& k, int sm_count, size_t max_smem_optin, int force_stages, uint32_t force_tile_bytes);
int launch_flex(const KPlan& k, int sm_count, int force_ctas_per_sm, void* stream);
int flex_set_attributes(size_t max_smem_optin);

// Implemented once per spatial factor in csic_rows_kernel.cu (explicit specialisations for F = 1, 2, 4, 8).
template <int F> int launch_rows_factor(const KPlan& k, unsigned grid, void* stream);
template <int F> int rows_set_attributes_factor(size_t max_smem_optin);
constexpr int kMaxTileRows = 64;        // rows per tile of the row kernel (small frames: 64 rows x 384 B still make a 24 KB tile; 256 rows
                                        // measured worse on 32x32 and 64x64 frames: the producer's per-row work, profiles/r2/rows_tall.txt)
constexpr uint32_t kTileMetaBytes = 32 + 4 * kMaxTileRows;   // sizeof(TileMeta) in csic_rows_kernel.cu
constexpr int kPoolMaxRows = 16;        // output rows per tile of the pooling kernel (each carries f input rows)
constexpr uint32_t kPoolMetaBytes = 32 + 4 * kPoolMaxRows;   // sizeof(PoolMeta) in csic_pool_kernel.cu
constexpr int kFlexMaxRows = 64;        // rows per tile of the flex kernel: the producer's lanes take two rows each

}  // namespace csic

#endif  // CSIC_INTERNAL_H_
