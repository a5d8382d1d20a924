// Error reporting, ABI version and launch accounting of libpia_b200.so.
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace pia {
static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace pia

extern "C" const char *pia_last_error(void) { return pia::g_err; }
extern "C" int pia_abi_version(void) { return PIA_ABI_VERSION; }
extern "C" unsigned long long pia_launch_count(void) { return pia::g_launches.load(); }
