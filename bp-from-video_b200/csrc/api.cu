// libbpv: error plumbing + version.
#include <stdarg.h>
#include "common.cuh"

namespace bpv {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace bpv

extern "C" int bpv_version(void) { return BPV_VERSION; }
extern "C" const char* bpv_last_error(void) { return bpv::g_err; }
