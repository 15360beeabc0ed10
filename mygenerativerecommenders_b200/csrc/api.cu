// api.cu — library-level entry points: version, error string, launch counter.
#include "common.cuh"
#include <cstring>
#include <mutex>

namespace grb {

static thread_local char t_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(t_err, sizeof(t_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

}  // namespace grb

extern "C" {

int grb_version(void) { return GRB_VERSION; }
const char* grb_last_error_string(void) { return grb::t_err; }
int64_t grb_launch_count(void) { return grb::g_launches.load(); }

}
