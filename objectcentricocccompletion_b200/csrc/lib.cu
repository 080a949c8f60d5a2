// lib.cu -- error reporting, launch accounting, ABI version.
#include <stdarg.h>
#include <stdio.h>

#include <atomic>

#include "common.cuh"

namespace occb200 {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace occb200

extern "C" int occb200_abi_version(void) { return 7; }
extern "C" void occb200_struct_sizes(int64_t *out4) {
  out4[0] = (int64_t)sizeof(occb200_pose_t);
  out4[1] = (int64_t)sizeof(occb200_sensor_t);
  out4[2] = (int64_t)sizeof(occb200_annotate_args_t);
  out4[3] = (int64_t)sizeof(occb200_ri_desc_t);
}
extern "C" const char *occb200_last_error(void) { return occb200::g_err; }
extern "C" int64_t occb200_launch_count(void) { return occb200::g_launches.load(); }
