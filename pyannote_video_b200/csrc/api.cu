// api.cu — error plumbing and library-level entry points of libpvb200.so.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include "../../include/pv_b200.h"
#include "pv_common.cuh"

std::atomic<long long> g_pv_launches{0};
static thread_local char g_err[1024] = "";

void pv_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* pv_last_error(void) { return g_err; }
extern "C" int pv_version(void) { return 100; }
extern "C" int64_t pv_launch_count(void) { return g_pv_launches.load(); }
