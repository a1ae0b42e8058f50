// common.cu — error state, device check, launch counter.
#include <stdlib.h>
#include <vector>
#include "common.cuh"

namespace rxb {

static thread_local char t_err[1024] = "";
long long g_launches = 0;

char* err_buf() { return t_err; }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
  return code;
}

bool g_prof_on = false;
bool g_dbg_sync = getenv("RXB_DBG_SYNC") && atoi(getenv("RXB_DBG_SYNC")) != 0;
bool g_pdl = getenv("RXB_PDL") && atoi(getenv("RXB_PDL")) != 0;   // measured: no gain on the DenseNet step
namespace {
struct ProfRec { cudaEvent_t a, b; int cat; };
std::vector<ProfRec> g_prof;
}  // namespace

ProfScope::ProfScope(cudaStream_t s, int cat) : st(s), slot(-1) {
  if (!g_prof_on) return;
  ProfRec r;
  r.cat = cat;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, st);
  slot = (int)g_prof.size();
  g_prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].b, st);
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

}  // namespace rxb

extern "C" {

int rxb_version(void) { return 100; }

const char* rxb_last_error(void) { return rxb::err_buf(); }

int rxb_check_device(void) {
  int dev = 0, major = 0, minor = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess)
    return rxb::set_error(RXB_ERR_NO_DEVICE, "no CUDA device: %s (librxb has no CPU fallback)",
                          cudaGetErrorString(e));
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10)
    return rxb::set_error(RXB_ERR_NO_DEVICE,
                          "device %d is sm_%d%d; librxb is built for sm_100a only (no fallback)", dev,
                          major, minor);
  return RXB_OK;
}

void rxb_profile_enable(int on) { rxb::g_prof_on = on != 0; }

// Sums the recorded launches per category (ms and count), then clears the record.  Synchronises the device.
int rxb_profile_collect(float* ms, long long* launches, int ncat) {
  if (!ms || !launches || ncat < rxb::PROF_NCAT) return rxb::set_error(RXB_ERR_INVALID, "rxb_profile_collect: need %d slots", (int)rxb::PROF_NCAT);
  for (int i = 0; i < ncat; ++i) { ms[i] = 0.f; launches[i] = 0; }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) return rxb::set_error(RXB_ERR_CUDA, "rxb_profile_collect: %s", cudaGetErrorString(e));
  for (auto& r : rxb::g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms[r.cat] += t; launches[r.cat] += 1; }
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  rxb::g_prof.clear();
  return RXB_OK;
}

int64_t rxb_launch_count(void) { return rxb::g_launches; }
void rxb_launch_count_reset(void) { rxb::g_launches = 0; }

}  // extern "C"
