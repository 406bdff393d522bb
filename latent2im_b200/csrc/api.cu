// Library-level entry points: ABI version, thread-local error string, launch counter.
#include <cstdlib>

#include "conv_common.cuh"

namespace l2i {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

KernelSwitches g_switches;

void refresh_kernel_switches() {
  auto flag = [](const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v == nullptr ? dflt : std::atoi(v);
  };
  KernelSwitches s;
  s.halo = flag("L2I_HALO", 1) != 0;
  s.halo_mask = flag("L2I_HALO_MASK", 15);
  s.halo_base_offset = flag("L2I_HALO_BASE_OFFSET", 0) != 0;
  s.quad = flag("L2I_QUAD", 1) != 0;
  s.ares = flag("L2I_ARES", 1) != 0;
  s.vpair = flag("L2I_VPAIR", 1) != 0;
  s.fir_simt = flag("L2I_FIR_SIMT", 0) != 0;
  s.uprow = flag("L2I_UPROW", 1) != 0;
  s.uprow_mask = flag("L2I_UPROW_MASK", 7);
  s.cluster = flag("L2I_CLUSTER", 0);
  s.ares_pair = flag("L2I_ARES_PAIR", 1) != 0;
  s.hring = flag("L2I_HRING", 1) != 0;
  s.hring_store = flag("L2I_HRING_STORE", 1) != 0;
  g_switches = s;
}

}  // namespace l2i

extern "C" int l2i_abi_version(void) { return 1; }
extern "C" const char* l2i_last_error_string(void) { return l2i::g_err; }
extern "C" int64_t l2i_launch_count(void) { return l2i::g_launches.load(std::memory_order_relaxed); }
