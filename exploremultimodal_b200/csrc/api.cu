// libmome: version / error / accounting entry points.
#include <stdarg.h>

#include "common.cuh"

namespace mome {
static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace mome

extern "C" int mome_version(void) { return MOME_ABI_VERSION; }
extern "C" const char* mome_last_error(void) { return mome::g_err; }
extern "C" int mome_sm_count(void) { return mome::sm_count(); }
extern "C" int64_t mome_launch_count(void) { return mome::g_launches.load(); }
