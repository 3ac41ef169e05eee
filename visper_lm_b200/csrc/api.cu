// Library-level plumbing of the C ABI: thread-local error string, launch counter, version.
#include "common.cuh"
#include "visper_b200.h"
#include <atomic>
#include <cstdarg>
#include <cstdio>

namespace vpb {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
// defaults: the variants measured faster on B200 in round 2 (profiles/r02_variants_ab.txt) are ON —
// WIN_ATTN_V2 (10) = 1, DWCONV_FFMA2 (11), ATTN_FWD_TC64 (12), GEMM_EPI8 (13), GATHER_FLAT (14); 0 selects the old kernel
static std::atomic<int> g_options[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 0};
int get_option(int key) { return (key >= 0 && key < 16) ? g_options[key].load() : 0; }
}  // namespace vpb

extern "C" int vpb_abi_version(void) { return VPB_ABI_VERSION; }
extern "C" const char* vpb_last_error(void) { return vpb::g_err; }
extern "C" int64_t vpb_launch_count(void) { return vpb::g_launches.load(); }
extern "C" void vpb_reset_launch_count(void) { vpb::g_launches.store(0); }
namespace vpb { void set_trace_buffer(void* p); }
extern "C" void vpb_set_trace_buffer(void* device_ptr) { vpb::set_trace_buffer(device_ptr); }
extern "C" int vpb_set_option(int key, int value) {
  if (key < 0 || key >= 16) return -1;
  vpb::g_options[key].store(value);
  return 0;
}
