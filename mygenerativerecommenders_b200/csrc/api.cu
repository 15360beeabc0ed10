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

int grb_bucket_octaves(const int64_t* thr, int32_t nb, uint32_t* out) {
  GRB_REQUIRE(thr != nullptr && out != nullptr && nb > 0, GRB_ERR_INVALID_ARG,
              "bucket_octaves: bad arguments");
  auto bucket = [&](int64_t d) {  // #{t : thr[t] <= d}
    int lo = 0, hi = nb;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (thr[mid] <= d) lo = mid + 1; else hi = mid; }
    return lo;
  };
  bool bad = false;
  const int b0 = bucket(0);
  for (int e = 0; e < 32; ++e) {
    const int64_t lo = 1ll << e, hi = (1ll << (e + 1)) - 1;
    const int base = bucket(lo);
    out[4 * e] = (uint32_t) base;
    for (int i = 0; i < 3; ++i) {
      const int idx = base + i;
      out[4 * e + 1 + i] = (idx < nb && thr[idx] <= hi) ? (uint32_t) thr[idx] : 0xffffffffu;
    }
    if (base + 3 < nb && thr[base + 3] <= hi) bad = true;
    if (e == 0 && b0 != base) bad = true;
  }
  out[128] = bad ? 1u : 0u;
  out[129] = (uint32_t) b0;
  return GRB_OK;
}

}
