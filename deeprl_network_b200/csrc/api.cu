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
