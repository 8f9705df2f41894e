// Minimal PNG reader / writer on top of zlib (the image has no libpng).  Host-side stand-in for scrimage in
// the reference's ImageProcessorModel.readImage / writeImage (src/test/scala/jpeg/ImageProcessorModel.scala:14-28).
// Reads 8-bit, non-interlaced gray / gray+alpha / RGB / RGBA / palette PNGs; writes 8-bit RGB.
#ifndef CSIC_PNG_IO_H_
#define CSIC_PNG_IO_H_
#include <cstdint>
#include <string>
#include <vector>

namespace csic_host {

struct Image {
  int width = 0, height = 0, channels = 0;   // channels: 3 (RGB) or 4 (RGBA), 8 bits each, row-major
  std::vector<uint8_t> pixels;
};

// Returns "" on success, else an error text.
std::string read_png(const std::string& path, Image& out);
std::string write_png_rgb(const std::string& path, const uint8_t* rgb, int width, int height);

}  // namespace csic_host
#endif
