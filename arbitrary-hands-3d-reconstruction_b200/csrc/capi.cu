// Error reporting + version for the C ABI (include/acr_b200.h).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace acr {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace acr

extern "C" const char* acr_b200_last_error(void) { return acr::g_err; }
extern "C" const char* acr_b200_version(void) { return "acr_b200 r2 sm_100a"; }
