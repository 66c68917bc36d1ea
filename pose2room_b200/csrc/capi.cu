// capi.cu -- process-wide pieces of the C ABI declared in include/p2r_b200.h.
#include "p2r_common.cuh"
#include <stdio.h>
#include <string.h>

static thread_local char g_last_error[512] = "";

extern "C" void p2r_set_last_error(const char* where, int code) {
  if (code > 0)
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s (cudaError %d)", where,
             cudaGetErrorString((cudaError_t)code), code);
  else
    snprintf(g_last_error, sizeof(g_last_error), "%s", where);
}

extern "C" const char* p2r_last_error(void) { return g_last_error; }

extern "C" int p2r_abi_version(void) { return 1; }

// Compiled-for architecture, so the host side can refuse to run on anything but sm_100.
extern "C" int p2r_compiled_arch(void) { return 100; }

extern "C" int p2r_device_sm_count(int device, int* sm_count, int* cc_major, int* cc_minor) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) { p2r_set_last_error("p2r_device_sm_count", (int)e); return (int)e; }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (cc_major) *cc_major = prop.major;
  if (cc_minor) *cc_minor = prop.minor;
  return 0;
}
