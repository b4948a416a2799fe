#include <stdarg.h>
#include "common.cuh"

namespace dcb {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace dcb

extern "C" int dc_version(void) { return 100; }
extern "C" const char* dc_last_error(void) { return dcb::g_err; }
