// jpeg_fuzz.cpp — TEST INFRASTRUCTURE ONLY.  Memory-safety fuzzer for the code the device JPEG decoder is built from
// (csrc/jpeg_fixed.cuh through tests/jpeg_host.cpp's two decode flows): mutates seed JPEG files (bit flips, byte
// splices, truncation, marker injection, random tails) and decodes them from exact-size heap buffers.  Build with
//   g++ -O1 -g -fwrapv -fsanitize=address,undefined -fno-sanitize-recover=all -I <csrc> tests/jpeg_fuzz.cpp tests/jpeg_host.cpp
// (-fwrapv: the 32-bit IDCT wraps on absurd coefficients of corrupt files, like libjpeg-turbo's SIMD and the GPU)
// and run:  ./jpeg_fuzz SEED SECONDS file1.jpg file2.jpg ...   (tools/fuzz_jpeg.sh does both).
// An out-of-bounds read, a misaligned access or signed overflow in the shared header aborts the run: on the GPU the
// same defect would be a fault on a corrupt input file.
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

extern "C" int jpeg_host_decode_gray(const uint8_t* data, int len, int H, int W, uint8_t* dst);
extern "C" int jpeg_host_decode_gray_parallel(const uint8_t* data, int len, int H, int W, uint8_t* dst, int* rounds);

static uint64_t s_rng = 88172645463325252ull;
static uint32_t rnd() {
  s_rng ^= s_rng << 13;
  s_rng ^= s_rng >> 7;
  s_rng ^= s_rng << 17;
  return (uint32_t)(s_rng >> 11);
}

static void frame_size(const std::vector<uint8_t>& f, int* H, int* W) {
  *H = *W = 8;
  for (size_t p = 2; p + 9 < f.size();) {
    if (f[p] != 0xFF) { ++p; continue; }
    const int m = f[p + 1];
    if (m == 0xC0 || m == 0xC1) { *H = (f[p + 5] << 8) | f[p + 6]; *W = (f[p + 7] << 8) | f[p + 8]; return; }
    if (m == 0xFF || m == 0 || (m >= 0xD0 && m <= 0xD8)) { p += (m == 0xFF) ? 1 : 2; continue; }
    p += 2 + ((f[p + 2] << 8) | f[p + 3]);
  }
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  s_rng ^= (uint64_t)atoll(argv[1]) * 0x9E3779B97F4A7C15ull;
  const double seconds = atof(argv[2]);
  std::vector<std::vector<uint8_t>> seeds;
  for (int i = 3; i < argc; ++i) {
    FILE* fp = fopen(argv[i], "rb");
    if (!fp) continue;
    std::vector<uint8_t> b;
    uint8_t tmp[4096];
    size_t n;
    while ((n = fread(tmp, 1, sizeof tmp, fp)) > 0) b.insert(b.end(), tmp, tmp + n);
    fclose(fp);
    if (b.size() > 4) seeds.push_back(b);
  }
  if (seeds.empty()) return 2;
  const auto t0 = std::chrono::steady_clock::now();
  long cases = 0, statuses[8] = {0};
  while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds) {
    std::vector<uint8_t> f = seeds[rnd() % seeds.size()];
    int H, W;
    frame_size(f, &H, &W);                                    // the size the caller expects: from the pristine file
    const int nmut = 1 + rnd() % 6;
    for (int m = 0; m < nmut && !f.empty(); ++m) {
      switch (rnd() % 7) {
        case 0: f[rnd() % f.size()] ^= (uint8_t)(1u << (rnd() % 8)); break;                 // bit flip
        case 1: f[rnd() % f.size()] = (uint8_t)rnd(); break;                                // random byte
        case 2: f.resize(1 + rnd() % f.size()); break;                                      // truncate
        case 3: { size_t p = rnd() % f.size(); f[p] = 0xFF; if (p + 1 < f.size()) f[p + 1] = (uint8_t)(0xC0 + rnd() % 0x40); break; }  // marker
        case 4: { size_t p = rnd() % f.size(), n = rnd() % 64; for (size_t i = p; i < f.size() && i < p + n; ++i) f[i] = (uint8_t)rnd(); break; }
        case 5: { size_t p = rnd() % f.size(), n = 1 + rnd() % 32; f.insert(f.begin() + p, n, (uint8_t)(rnd() % 3 == 0 ? 0xFF : rnd())); break; }
        default: { size_t p = rnd() % f.size(), n = rnd() % 16; f.erase(f.begin() + p, f.begin() + (p + n < f.size() ? p + n : f.size())); break; }
      }
    }
    if (f.empty()) continue;
    uint8_t* data = new uint8_t[f.size()];                    // exact size: any over-read trips the sanitizer
    memcpy(data, f.data(), f.size());
    uint8_t* dst = new uint8_t[(size_t)H * W];
    int rounds = 0;
    const int a = jpeg_host_decode_gray(data, (int)f.size(), H, W, dst);
    const int b = jpeg_host_decode_gray_parallel(data, (int)f.size(), H, W, dst, &rounds);
    if (a >= 0 && a < 8) ++statuses[a];
    (void)b;
    delete[] data;
    delete[] dst;
    ++cases;
  }
  printf("cases %ld statuses ok=%ld notjpeg=%ld unsupported=%ld table=%ld size=%ld code=%ld\n", cases, statuses[0],
         statuses[1], statuses[2], statuses[3], statuses[4], statuses[5]);
  return 0;
}
