// common.cu — error state, device check, launch counter.
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

int64_t rxb_launch_count(void) { return rxb::g_launches; }
void rxb_launch_count_reset(void) { rxb::g_launches = 0; }

}  // extern "C"
