// api.cu — library-level entry points of libclasr_sm100.so (version, error text, launch counter).
#include <stdarg.h>
#include <string.h>
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

// ---- optional per-kernel timing (CUDA events on the launch stream), used by bench.py for the roofline numbers
struct ProfSlot {
  char name[32];
  cudaEvent_t ev[2 * 64];
  int n;       // recorded pairs
  int made;    // created events
};
static ProfSlot g_slots[16];
static int g_nslots = 0;
static int g_profiling = 0;

static ProfSlot* slot(const char* name) {
  for (int i = 0; i < g_nslots; ++i)
    if (!strncmp(g_slots[i].name, name, 31)) return &g_slots[i];
  if (g_nslots == 16) return nullptr;
  ProfSlot* sl = &g_slots[g_nslots++];
  strncpy(sl->name, name, 31);
  sl->name[31] = 0;
  sl->n = 0;
  sl->made = 0;
  return sl;
}

void prof_begin(const char* name, cudaStream_t s) {
  if (!g_profiling) return;
  ProfSlot* sl = slot(name);
  if (!sl || sl->n >= 64) return;
  while (sl->made < 2 * (sl->n + 1)) cudaEventCreate(&sl->ev[sl->made++]);
  cudaEventRecord(sl->ev[2 * sl->n], s);
}
void prof_end(const char* name, cudaStream_t s) {
  if (!g_profiling) return;
  ProfSlot* sl = slot(name);
  if (!sl || sl->n >= 64 || sl->made < 2 * (sl->n + 1)) return;
  cudaEventRecord(sl->ev[2 * sl->n + 1], s);
  sl->n++;
}

}  // namespace clasr

extern "C" int clasr_version(void) { return 100; }
extern "C" const char* clasr_last_error(void) { return clasr::g_err; }
extern "C" int64_t clasr_launch_count(void) { return (int64_t)clasr::g_launches.load(std::memory_order_relaxed); }

extern "C" void clasr_set_profiling(int on) { clasr::g_profiling = on; }
extern "C" void clasr_profile_reset(void) {
  for (int i = 0; i < clasr::g_nslots; ++i) clasr::g_slots[i].n = 0;
}
// mean duration (ms) of the kernel(s) recorded under `name` since the last reset; -1 if none.  Synchronises.
extern "C" float clasr_profile_ms(const char* name, int* count) {
  for (int i = 0; i < clasr::g_nslots; ++i) {
    clasr::ProfSlot* sl = &clasr::g_slots[i];
    if (strncmp(sl->name, name, 31) || sl->n == 0) continue;
    double tot = 0;
    for (int j = 0; j < sl->n; ++j) {
      cudaEventSynchronize(sl->ev[2 * j + 1]);
      float ms = 0;
      cudaEventElapsedTime(&ms, sl->ev[2 * j], sl->ev[2 * j + 1]);
      tot += ms;
    }
    if (count) *count = sl->n;
    return (float)(tot / sl->n);
  }
  if (count) *count = 0;
  return -1.f;
}
