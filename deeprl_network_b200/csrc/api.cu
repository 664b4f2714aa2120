// api.cu -- error reporting and ABI self-description for libnmarl.
#include <stdarg.h>
#include "common.cuh"

static thread_local char g_err[512] = "";

void nmarl_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* nmarl_last_error(void) { return g_err; }
extern "C" int nmarl_version(void) { return 100; }
extern "C" int nmarl_sizeof_model(void) { return (int)sizeof(nmarl_model); }
extern "C" int nmarl_sizeof_agent(void) { return (int)sizeof(nmarl_agent); }
extern "C" int nmarl_sizeof_cacc_cfg(void) { return (int)sizeof(nmarl_cacc_cfg); }

// debug hook (not part of the public ABI): device buffer of >= 128 int64 receiving clock64() stamps from
// CTA (0,0) of the tensor-core forward kernel
long long* g_nmarl_prof = nullptr;
extern "C" __attribute__((visibility("default"))) void nmarl_debug_set_prof(long long* p) { g_nmarl_prof = p; }
