#include "png_io.h"

#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace csic_host {
namespace {

uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)(x >> 24)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)x);
}
int paeth(int a, int b, int c) {
  const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
  return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
bool read_file(const std::string& path, std::vector<uint8_t>& data) {
  FILE* f = std::fopen(path.c_str(), "rb");
  if (!f) return false;
  std::fseek(f, 0, SEEK_END);
  long n = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  data.resize(n > 0 ? (size_t)n : 0);
  const bool ok = n >= 0 && std::fread(data.data(), 1, data.size(), f) == data.size();
  std::fclose(f);
  return ok;
}

}  // namespace

std::string read_png(const std::string& path, Image& out) {
  std::vector<uint8_t> d;
  if (!read_file(path, d)) return "cannot read " + path;
  static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  if (d.size() < 8 || std::memcmp(d.data(), sig, 8) != 0) return "not a PNG: " + path;
  int w = 0, h = 0, depth = 0, ctype = -1, interlace = 0;
  std::vector<uint8_t> idat, plte, trns;
  for (size_t pos = 8; pos + 12 <= d.size();) {
    const uint32_t len = be32(&d[pos]);
    const char* type = reinterpret_cast<const char*>(&d[pos + 4]);
    if (pos + 12 + len > d.size()) return "truncated PNG";
    const uint8_t* body = &d[pos + 8];
    if (!std::memcmp(type, "IHDR", 4)) {
      if (len < 13) return "malformed PNG (IHDR shorter than 13 bytes)";
      if (be32(body) > 0x7FFFFFFFu || be32(body + 4) > 0x7FFFFFFFu) return "malformed PNG (dimensions out of range)";
      w = (int)be32(body); h = (int)be32(body + 4); depth = body[8]; ctype = body[9]; interlace = body[12];
    } else if (!std::memcmp(type, "PLTE", 4)) {
      plte.assign(body, body + len);
    } else if (!std::memcmp(type, "tRNS", 4)) {
      trns.assign(body, body + len);
    } else if (!std::memcmp(type, "IDAT", 4)) {
      idat.insert(idat.end(), body, body + len);
    } else if (!std::memcmp(type, "IEND", 4)) {
      break;
    }
    pos += 12 + len;
  }
  if (w <= 0 || h <= 0 || depth != 8 || interlace != 0) return "unsupported PNG (need 8-bit, non-interlaced)";
  int spp;   // samples per pixel in the file
  switch (ctype) {
    case 0: spp = 1; break;
    case 2: spp = 3; break;
    case 3: spp = 1; break;
    case 4: spp = 2; break;
    case 6: spp = 4; break;
    default: return "unsupported PNG colour type";
  }
  const size_t stride = (size_t)w * spp;
  // a hostile IHDR must end in an error message, not in std::bad_alloc: 2^31 bytes of raw scanlines is the cap
  if (stride + 1 > ((size_t)1 << 31) / (size_t)h) return "PNG too large (more than 2 GiB of scanlines)";
  std::vector<uint8_t> raw;
  try {
    raw.resize((stride + 1) * (size_t)h);
  } catch (const std::bad_alloc&) {
    return "out of memory reading PNG";
  }
  uLongf rawlen = (uLongf)raw.size();
  if (uncompress(raw.data(), &rawlen, idat.data(), (uLong)idat.size()) != Z_OK || rawlen != raw.size()) return "PNG inflate failed";
  std::vector<uint8_t> img(stride * (size_t)h);
  for (int y = 0; y < h; ++y) {           // undo the per-row filters
    const uint8_t* src = &raw[(stride + 1) * (size_t)y];
    uint8_t* cur = &img[stride * (size_t)y];
    const uint8_t* up = y ? &img[stride * (size_t)(y - 1)] : nullptr;
    const int ft = src[0];
    for (size_t x = 0; x < stride; ++x) {
      const int a = x >= (size_t)spp ? cur[x - spp] : 0, b = up ? up[x] : 0, c = (up && x >= (size_t)spp) ? up[x - spp] : 0;
      int v = src[1 + x];
      switch (ft) {
        case 0: break;
        case 1: v += a; break;
        case 2: v += b; break;
        case 3: v += (a + b) / 2; break;
        case 4: v += paeth(a, b, c); break;
        default: return "bad PNG filter";
      }
      cur[x] = (uint8_t)v;
    }
  }
  const bool alpha = ctype == 4 || ctype == 6 || (ctype == 3 && !trns.empty());
  out.width = w; out.height = h; out.channels = alpha ? 4 : 3;
  out.pixels.resize((size_t)w * h * out.channels);
  for (size_t i = 0; i < (size_t)w * h; ++i) {
    uint8_t r, g, b, a = 255;
    const uint8_t* s = &img[i * spp];
    if (ctype == 0) { r = g = b = s[0]; }
    else if (ctype == 4) { r = g = b = s[0]; a = s[1]; }
    else if (ctype == 2) { r = s[0]; g = s[1]; b = s[2]; }
    else if (ctype == 6) { r = s[0]; g = s[1]; b = s[2]; a = s[3]; }
    else {
      const size_t k = s[0];
      if (3 * k + 2 >= plte.size()) return "palette index out of range";
      r = plte[3 * k]; g = plte[3 * k + 1]; b = plte[3 * k + 2];
      a = k < trns.size() ? trns[k] : 255;
    }
    uint8_t* o = &out.pixels[i * out.channels];
    o[0] = r; o[1] = g; o[2] = b;
    if (alpha) o[3] = a;
  }
  return "";
}

std::string write_png_rgb(const std::string& path, const uint8_t* rgb, int width, int height) {
  std::vector<uint8_t> raw(((size_t)width * 3 + 1) * (size_t)height);
  for (int y = 0; y < height; ++y) {
    uint8_t* row = &raw[((size_t)width * 3 + 1) * (size_t)y];
    row[0] = 0;   // filter: none
    std::memcpy(row + 1, rgb + (size_t)y * width * 3, (size_t)width * 3);
  }
  uLongf clen = compressBound((uLong)raw.size());
  std::vector<uint8_t> comp(clen);
  if (compress2(comp.data(), &clen, raw.data(), (uLong)raw.size(), 6) != Z_OK) return "PNG deflate failed";
  comp.resize(clen);
  std::vector<uint8_t> f = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
  auto chunk = [&](const char* type, const std::vector<uint8_t>& body) {
    put32(f, (uint32_t)body.size());
    const size_t start = f.size();
    f.insert(f.end(), type, type + 4);
    f.insert(f.end(), body.begin(), body.end());
    put32(f, (uint32_t)crc32(0L, &f[start], (uInt)(f.size() - start)));
  };
  std::vector<uint8_t> ihdr;
  put32(ihdr, (uint32_t)width); put32(ihdr, (uint32_t)height);
  ihdr.push_back(8); ihdr.push_back(2); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
  chunk("IHDR", ihdr);
  chunk("IDAT", comp);
  chunk("IEND", {});
  FILE* out = std::fopen(path.c_str(), "wb");
  if (!out) return "cannot write " + path;
  const bool ok = std::fwrite(f.data(), 1, f.size(), out) == f.size();
  std::fclose(out);
  return ok ? "" : "short write to " + path;
}

}  // namespace csic_host
