// Error plumbing + device queries for the C ABI (include/b200audio.h).
#include "common.cuh"
#include <stdlib.h>
#include <stdarg.h>

static thread_local char g_err[512] = "";

void b2a_set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}

extern "C" const char* b2a_last_error(void) { return g_err; }
extern "C" int32_t b2a_version(void) { return 100; }
extern "C" int32_t b2a_device_sm_count(void) {
  int dev = 0, n = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
  return n;
}

bool b2a_pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("B2A_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}
