// runtime.cu -- error state, version, device check.
#include "common.cuh"
#include <cstdarg>

namespace acan {
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static long long g_launches = 0;
void note_launch() { ++g_launches; }

// ---- event profiling: begin/end marks per kind, resolved at acan_profile_end
struct ProfPair { cudaEvent_t a, b; int kind; };
static bool g_prof_on = false;
static bool g_prof_coarse = false;   // coarse: one bracket around a whole attention call (PROF_CORE) instead of one per pass -- events between
                                     // the passes keep them from overlapping (programmatic dependent launch), so the two views are separate runs
static ProfPair* g_pairs = nullptr;
static int g_pairs_cap = 0, g_pairs_n = 0;
static int g_open[PROF_KINDS];

void prof_mark(int kind, bool begin, cudaStream_t st) {
  if (!g_prof_on) return;
  if (g_prof_coarse ? (kind <= PROF_PV) : (kind == PROF_CORE)) return;
  if (begin) {
    if (g_pairs_n >= g_pairs_cap) { g_open[kind] = -1; return; }
    ProfPair& p = g_pairs[g_pairs_n];
    p.kind = kind;
    cudaEventRecord(p.a, st);
    g_open[kind] = g_pairs_n++;
  } else if (g_open[kind] >= 0) {
    cudaEventRecord(g_pairs[g_open[kind]].b, st);
    g_open[kind] = -1;
  }
}
}  // namespace acan

extern "C" long long acan_launch_count(void) { return acan::g_launches; }

extern "C" int acan_profile_begin(int max_marks) {
  using namespace acan;
  if (max_marks <= 0) return fail(ACAN_ERR_BAD_ARG, "acan_profile_begin: max_marks must be positive");
  if (g_pairs_cap < max_marks) {
    ProfPair* np = new ProfPair[max_marks];
    for (int i = 0; i < g_pairs_cap; ++i) np[i] = g_pairs[i];
    for (int i = g_pairs_cap; i < max_marks; ++i) { ACAN_CUDA(cudaEventCreate(&np[i].a)); ACAN_CUDA(cudaEventCreate(&np[i].b)); np[i].kind = 0; }
    delete[] g_pairs;
    g_pairs = np;
    g_pairs_cap = max_marks;
  }
  g_pairs_n = 0;
  for (int k = 0; k < PROF_KINDS; ++k) g_open[k] = -1;
  g_prof_coarse = false;
  g_prof_on = true;
  return 0;
}

extern "C" int acan_profile_begin_coarse(int max_marks) {
  if (int e = acan_profile_begin(max_marks)) return e;
  acan::g_prof_coarse = true;
  return 0;
}

extern "C" int acan_profile_end(float* ms_by_kind, int* count_by_kind) {
  using namespace acan;
  ACAN_CHECK_ARG(ms_by_kind && count_by_kind, "acan_profile_end: null pointer");
  g_prof_on = false;
  for (int k = 0; k < PROF_KINDS; ++k) { ms_by_kind[k] = 0.f; count_by_kind[k] = 0; }
  for (int i = 0; i < g_pairs_n; ++i) {
    ACAN_CUDA(cudaEventSynchronize(g_pairs[i].b));
    float ms = 0.f;
    ACAN_CUDA(cudaEventElapsedTime(&ms, g_pairs[i].a, g_pairs[i].b));
    ms_by_kind[g_pairs[i].kind] += ms;
    count_by_kind[g_pairs[i].kind] += 1;
  }
  g_pairs_n = 0;
  return 0;
}

extern "C" int acan_abi_version(void) { return ACAN_ABI_VERSION; }
extern "C" const char* acan_last_error(void) { return acan::g_err; }

extern "C" int acan_device_check(void) {
  int dev = 0;
  ACAN_CUDA(cudaGetDevice(&dev));
  int major = 0;
  ACAN_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if (major != 10) return acan::fail(ACAN_ERR_ARCH, "device %d has compute capability %d.x; libacan_b200 is sm_100a only", dev, major);
  return 0;
}
