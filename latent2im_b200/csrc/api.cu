// Library-level entry points: ABI version, thread-local error string, launch counter.
#include "common.cuh"

namespace l2i {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

}  // namespace l2i

extern "C" int l2i_abi_version(void) { return 1; }
extern "C" const char* l2i_last_error_string(void) { return l2i::g_err; }
extern "C" int64_t l2i_launch_count(void) { return l2i::g_launches.load(std::memory_order_relaxed); }
