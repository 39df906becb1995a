// Error plumbing + small utilities of the C-ABI (see include/genrl_b200.h).
#include "common.cuh"
#include <stdarg.h>
#include <stdlib.h>

namespace genrl {

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return 1;
  }
  return 0;
}

static long long g_launches = 0;

bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("GENRL_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

}  // namespace genrl

extern "C" const char* genrl_last_error() { return genrl::g_err; }

extern "C" int genrl_abi_version() { return 1; }

// device properties the host side sizes its grids with
extern "C" int genrl_device_info(int* sm_count, int* cc_major, int* cc_minor, long long* smem_optin) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { genrl::set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return 1; }
  cudaDeviceProp p;
  e = cudaGetDeviceProperties(&p, dev);
  if (e != cudaSuccess) { genrl::set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return 1; }
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (smem_optin) *smem_optin = (long long)p.sharedMemPerBlockOptin;
  return 0;
}
