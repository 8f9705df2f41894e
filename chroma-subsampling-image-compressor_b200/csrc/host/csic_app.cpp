// csic_app -- native host program with the command line of the reference's `ImageCompressionApp`
// (src/test/scala/jpeg/ImageCompressorTopApp.scala:147-216):
//
//   csic_app --input test_images/in128x128.png --a 2 --b 0 --yq 3 --cbq 3 --crq 2 --sf 1
//            --op1 chroma --op2 color --op3 spatial
//
// Same flags, defaults (:164-173), banner (:177-185) and output naming (:187-190).  Where the reference elaborates
// a Chisel DUT and simulates it pixel by pixel (:53-131), this calls libcsic.so once: PNG -> csic_process_host
// (fused RGB2YCbCr / chroma / spatial / quantiser / ycbcr2rgb on the GPU) -> PNG.  Only the C ABI of include/csic.h
// is used; an RGBA PNG is handed over as RGBA32 (alpha ignored, like pixel.red/green/blue, :86-89).
// Extra: --outdir DIR (default APP_OUTPUT), --device N, --selftest-png IN OUT (re-encode a PNG; no GPU).
#include <sys/stat.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../../include/csic.h"
#include "png_io.h"

namespace {

const char* step_name(int s) {
  return s == CSIC_STEP_SPATIAL ? "SpatialSampling" : s == CSIC_STEP_COLOR ? "ColorQuantization" : "ChromaSubsampling";
}

void mkdirs(const std::string& dir) {
  std::string cur;
  for (size_t i = 0; i <= dir.size(); ++i) {
    if (i == dir.size() || dir[i] == '/') {
      if (!cur.empty()) mkdir(cur.c_str(), 0755);
    }
    if (i < dir.size()) cur.push_back(dir[i]);
  }
}

// ImageCompressionApp.processImage, ImageCompressorTopApp.scala:23-145
int process_image(const std::string& in_path, const std::string& out_path, int a, int b, int yq, int cbq, int crq, int sf,
                  int op1, int op2, int op3, int device) {
  csic_host::Image img;
  std::string err = csic_host::read_png(in_path, img);                                   // :39
  if (!err.empty()) { std::fprintf(stderr, "[ERROR] %s\n", err.c_str()); return 2; }
  const int W = img.width, H = img.height;
  const bool spatial = op1 == CSIC_STEP_SPATIAL || op2 == CSIC_STEP_SPATIAL || op3 == CSIC_STEP_SPATIAL;   // :43
  if (spatial && sf == 0) {   // the reference computes w / spatialFactorToUse first (:44): Scala throws before any `require`
    std::fprintf(stderr, "java.lang.ArithmeticException: / by zero\n");
    return 3;
  }
  const int out_w = spatial ? W / sf : W, out_h = spatial ? H / sf : H;                   // :44-45
  if (spatial && sf > 0 && (W % sf != 0 || H % sf != 0))
    std::printf("[WARN] Image dimensions (%dx%d) are not perfectly divisible by spatialFactor (%d). SpatialDownsampler might truncate.\n", W, H, sf);

  csic_params p;
  std::memset(&p, 0, sizeof p);
  p.width = W; p.height = H; p.chroma_a = a; p.chroma_b = b; p.y_bits = yq; p.cb_bits = cbq; p.cr_bits = crq;
  p.factor = sf; p.op[0] = op1; p.op[1] = op2; p.op[2] = op3;
  p.out_format = CSIC_OUT_RGB888;                                                        // fused ycbcr2rgb (:118)
  p.in_format = img.channels == 4 ? CSIC_IN_RGBA32 : CSIC_IN_RGB24;
  char msg[256];
  int rc = csic_validate(&p, msg, sizeof msg);
  if (rc != CSIC_OK) {   // what the Scala constructor's require(...) throws
    std::fprintf(stderr, "Exception in thread \"main\" java.lang.IllegalArgumentException: requirement failed: %s\n", msg);
    return 3;
  }
  int32_t ow = 0, oh = 0; size_t fb = 0;
  csic_out_shape(&p, &ow, &oh, nullptr, &fb);
  csic_ctx* ctx = nullptr;
  if ((rc = csic_create(device, &ctx)) != CSIC_OK) {
    std::fprintf(stderr, "[ERROR] %s (%s)\n", csic_strerror(rc), csic_last_error());
    return 4;
  }
  std::vector<uint8_t> stream(fb);
  rc = csic_process_host(ctx, &p, img.pixels.data(), 1, stream.data());
  csic_destroy(ctx);
  if (rc != CSIC_OK) { std::fprintf(stderr, "[ERROR] %s (%s)\n", csic_strerror(rc), csic_last_error()); return 5; }

  // :108-142 -- the first out_w*out_h emitted pixels, row-major, on a magenta canvas
  std::vector<uint8_t> canvas((size_t)out_w * out_h * 3);
  for (size_t i = 0; i < (size_t)out_w * out_h; ++i) { canvas[3 * i] = 255; canvas[3 * i + 1] = 0; canvas[3 * i + 2] = 255; }
  const size_t n = std::min((size_t)out_w * out_h, (size_t)ow * oh);
  std::memcpy(canvas.data(), stream.data(), n * 3);
  const size_t slash = out_path.find_last_of('/');
  if (slash != std::string::npos) mkdirs(out_path.substr(0, slash));                      // getParentFile().mkdirs()
  err = csic_host::write_png_rgb(out_path, canvas.data(), out_w, out_h);                  // :144
  if (!err.empty()) { std::fprintf(stderr, "[ERROR] %s\n", err.c_str()); return 6; }
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  std::map<std::string, std::string> args;                     // args.sliding(2, 2), :149-151
  if (argc >= 4 && std::string(argv[1]) == "--selftest-png") {
    csic_host::Image img;
    std::string err = csic_host::read_png(argv[2], img);
    if (!err.empty()) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
    std::vector<uint8_t> rgb((size_t)img.width * img.height * 3);
    for (size_t i = 0; i < (size_t)img.width * img.height; ++i)
      std::memcpy(&rgb[3 * i], &img.pixels[i * img.channels], 3);
    err = csic_host::write_png_rgb(argv[3], rgb.data(), img.width, img.height);
    if (!err.empty()) { std::fprintf(stderr, "%s\n", err.c_str()); return 1; }
    std::printf("%dx%d channels=%d\n", img.width, img.height, img.channels);
    return 0;
  }
  for (int i = 1; i + 1 < argc; i += 2)
    if (std::strncmp(argv[i], "--", 2) == 0) args[argv[i]] = argv[i + 1];
  auto get = [&](const char* k, const char* d) { auto it = args.find(k); return it == args.end() ? std::string(d) : it->second; };
  const std::string input = get("--input", "test_images/in128x128.png");
  const int a = std::atoi(get("--a", "4").c_str()), b = std::atoi(get("--b", "4").c_str());
  const int yq = std::atoi(get("--yq", "8").c_str()), cbq = std::atoi(get("--cbq", "8").c_str()), crq = std::atoi(get("--crq", "8").c_str());
  const int sf = std::atoi(get("--sf", "8").c_str());
  int ops[3];
  const char* keys[3] = {"--op1", "--op2", "--op3"};
  const char* defs[3] = {"spatial", "color", "chroma"};
  for (int i = 0; i < 3; ++i) {
    const std::string name = get(keys[i], defs[i]);
    ops[i] = csic_parse_step(name.c_str());                    // :154-161
    if (ops[i] < 0) {
      std::fprintf(stderr, "Exception in thread \"main\" java.lang.IllegalArgumentException: Unknown processing step: %s. "
                           "Use 'spatial', 'color', or 'chroma'.\n", name.c_str());
      return 3;
    }
  }
  std::string base = input.substr(input.find_last_of('/') == std::string::npos ? 0 : input.find_last_of('/') + 1);
  base = base.substr(0, base.find('.'));                        // getName.takeWhile(_ != '.'), :175
  const char* bar = "----------------------------------------------------";
  std::printf("%s\nImage Compressor Application Parameters:\n%s\n", bar, bar);
  std::printf("Input Image: %s\n", input.c_str());
  std::printf("Selected Chroma Subsampling (J:a:b): 4:%d:%d\n", a, b);
  std::printf("Selected Quantization Bits (Y/Cb/Cr): %d/%d/%d\n", yq, cbq, crq);
  std::printf("Selected Spatial Downsampling Factor: %d\n", sf);
  std::printf("Selected Pipeline Order: %s -> %s -> %s\n%s\n", step_name(ops[0]), step_name(ops[1]), step_name(ops[2]), bar);
  const std::string outdir = get("--outdir", "APP_OUTPUT");
  char suffix[256];
  std::snprintf(suffix, sizeof suffix, "chroma4-%d-%d_Y%dCb%dCr%d_sf%d_order-%.2s-%.2s-%.2s", a, b, yq, cbq, crq, sf,
                step_name(ops[0]), step_name(ops[1]), step_name(ops[2]));                 // :188-189
  const std::string out_path = outdir + "/" + base + "_processed_" + suffix + ".png";
  FILE* probe = std::fopen(input.c_str(), "rb");
  if (!probe) { std::printf("[ERROR] Input image not found: %s\n", input.c_str()); return 1; }   // :198-199
  std::fclose(probe);
  const int rc = process_image(input, out_path, a, b, yq, cbq, crq, sf, ops[0], ops[1], ops[2], std::atoi(get("--device", "0").c_str()));
  if (rc == 0) std::printf("Image processing complete. Output saved to: %s\n", out_path.c_str());
  return rc;
}
