// api.cu — library-level entry points of libclasr_sm100.so (version, error text, launch counter).
#include <stdarg.h>
#include <atomic>

#include "common.cuh"

namespace clasr {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // namespace clasr

extern "C" int clasr_version(void) { return 100; }
extern "C" const char* clasr_last_error(void) { return clasr::g_err; }
extern "C" int64_t clasr_launch_count(void) { return (int64_t)clasr::g_launches.load(std::memory_order_relaxed); }
