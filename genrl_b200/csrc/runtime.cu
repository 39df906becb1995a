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
    // Measured on B200 (profiles/r02_notes.md): inside the CUDA-graph replays the step does not get faster with programmatic edges
    // (25.87 ms with, 25.60 ms without) - the 4-20 us kernels are bound by their own ramp, not by the gap between them - so the
    // attribute is opt-in (GENRL_PDL=1); the in-kernel griddepcontrol instructions are no-ops without it.
    const char* e = getenv("GENRL_PDL");
    on = (e && e[0] == '1') ? 1 : 0;
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
