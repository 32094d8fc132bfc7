// Error reporting, version string, launch accounting.
#include <atomic>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace mgp {

static thread_local char g_err[1024] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("MGP_PDL");
    return e != nullptr && e[0] == '1';       // opt-in until measured (MGP_PDL=1)
  }();
  return on;
}

}  // namespace mgp

extern "C" {

const char* mgp_last_error(void) { return mgp::g_err; }

const char* mgp_version(void) { return "mgp_b200 0.1.0 sm_100a"; }

int64_t mgp_launch_count(void) { return mgp::g_launches.load(std::memory_order_relaxed); }

void mgp_reset_launch_count(void) { mgp::g_launches.store(0, std::memory_order_relaxed); }

void mgp_add_launch_count(int64_t n) { mgp::g_launches.fetch_add(n, std::memory_order_relaxed); }

}  // extern "C"
