// Error reporting and version for the C ABI (include/quadfield.h).
#include <stdarg.h>

#include "common.cuh"

namespace qf {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace qf

extern "C" const char* qf_last_error(void) { return qf::g_err; }
extern "C" int qf_version(void) { return 100; }
