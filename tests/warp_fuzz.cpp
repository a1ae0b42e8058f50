// warp_fuzz.cpp — TEST INFRASTRUCTURE ONLY.  Memory-safety fuzzer for the arithmetic of the arbitrary-angle loader
// (csrc/warp_fixed.cuh through tests/warp_host.cpp): random, degenerate and non-finite 2x3 matrices on small images read
// from exact-size heap buffers.  Whatever the matrix, every tap must stay inside the image.  Build like jpeg_fuzz.cpp
// (g++ -O1 -g -fwrapv -ffp-contract=off -fsanitize=address,undefined ...); tools/fuzz_jpeg.sh runs both.
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <limits>

extern "C" void warp_host_planar_u8(const uint8_t* src, int H, int W, const double* M, int vflip, int hflip, int y0,
                                    int x0, int Ho, int Wo, uint8_t* dst);

static uint64_t s_rng = 0x2545F4914F6CDD1Dull;
static uint32_t rnd() {
  s_rng ^= s_rng << 13;
  s_rng ^= s_rng >> 7;
  s_rng ^= s_rng << 17;
  return (uint32_t)(s_rng >> 11);
}
static double pick() {
  switch (rnd() % 10) {
    case 0: return 0.0;
    case 1: return std::numeric_limits<double>::quiet_NaN();
    case 2: return (rnd() & 1) ? INFINITY : -INFINITY;
    case 3: return ((int)(rnd() % 2001) - 1000) * 1e6;
    case 4: return ((int)(rnd() % 2001) - 1000) * 1e-9;
    case 5: return ((int)(rnd() % 2001) - 1000) * 3.3e4;
    default: return ((int)(rnd() % 4001) - 2000) / 500.0;
  }
}

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  s_rng ^= (uint64_t)atoll(argv[1]) * 0x9E3779B97F4A7C15ull;
  const double seconds = atof(argv[2]);
  const auto t0 = std::chrono::steady_clock::now();
  long cases = 0;
  while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds) {
    const int H = 1 + rnd() % 24, W = 1 + rnd() % 24;
    const int Ho = 1 + rnd() % H, Wo = 1 + rnd() % W;
    const int y0 = rnd() % (H - Ho + 1), x0 = rnd() % (W - Wo + 1);
    uint8_t* src = new uint8_t[6 * H * W];
    for (int i = 0; i < 6 * H * W; ++i) src[i] = (uint8_t)rnd();
    uint8_t* dst = new uint8_t[6 * Ho * Wo];
    double M[6];
    for (double& m : M) m = pick();
    warp_host_planar_u8(src, H, W, M, rnd() & 1, rnd() & 1, y0, x0, Ho, Wo, dst);
    delete[] src;
    delete[] dst;
    ++cases;
  }
  printf("warp cases %ld\n", cases);
  return 0;
}
