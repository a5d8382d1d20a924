// Error reporting, ABI version and launch accounting of libpia_b200.so.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"

namespace pia {
static thread_local char g_err[512] = "";
std::atomic<unsigned long long> g_launches{0};

thread_local int g_pdl_off = 0;
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("PIA_PDL"); v = (e && e[0] == '0') ? 0 : 1; }
  return v == 1;
}

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
