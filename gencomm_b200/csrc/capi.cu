// Error reporting + version for the C ABI declared in include/gencomm_b200.h.
#include <stdarg.h>

#include "common.cuh"

namespace gc {
static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace gc

extern "C" int gc_version(void) { return 100; /* 0.1.0 */ }

extern "C" const char *gc_last_error(void) { return gc::g_err; }
