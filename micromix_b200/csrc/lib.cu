// lib.cu -- process-wide plumbing of libmicromix_b200.so: error text, options, launch counter, SF geometry.
#include <cstring>
#include <mutex>

#include "common.h"

namespace mmx {

static thread_local char g_err[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

Options& options() {
  static Options o;
  return o;
}

// Per-device facts, cached per device slot (one process may drive several GPUs, common.h) behind one mutex.
namespace {
struct DevInfo {
  int sms = 0;
  int major = -1;
};
DevInfo g_dev[kMaxDevices];
std::mutex g_dev_mu;

const DevInfo& dev_info() {
  const int slot = current_device_slot();
  std::lock_guard<std::mutex> lk(g_dev_mu);
  DevInfo& d = g_dev[slot];
  if (d.major < 0) {
    int dev = 0, n = 0, major = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) major = 0;
    d.sms = n;
    d.major = major;
  }
  return d;
}
}  // namespace

int sm_count() { return dev_info().sms; }

bool device_is_sm100() { return dev_info().major == 10; }

}  // namespace mmx

extern "C" __attribute__((visibility("default"))) int mmx_version(void) { return 100; }

extern "C" __attribute__((visibility("default"))) const char* mmx_last_error(void) { return mmx::g_err; }

extern "C" __attribute__((visibility("default"))) int64_t mmx_sf_bytes_act(int64_t M, int64_t Kseg) { return (M / 128 + 1) * 128 * Kseg / 32; }

extern "C" __attribute__((visibility("default"))) int64_t mmx_sf_bytes_wgt(int64_t N, int64_t Kseg) { return ((N + 127) / 128) * 128 * Kseg / 32; }

extern "C" __attribute__((visibility("default"))) int64_t mmx_sf_offset(int64_t r, int64_t g, int64_t Kseg) {
  const int64_t katoms = (Kseg + 127) / 128;
  return (r / 128) * katoms * 512 + (g / 4) * 512 + (r % 32) * 16 + ((r / 32) % 4) * 4 + (g % 4);
}

extern "C" __attribute__((visibility("default"))) int64_t mmx_launch_count(void) { return mmx::g_launches.load(std::memory_order_relaxed); }

extern "C" __attribute__((visibility("default"))) int mmx_set_option(const char* key, int64_t value) {
  if (!key) return MMX_ERR_INVALID;
  mmx::Options& o = mmx::options();
  if (!std::strcmp(key, "gemm_watchdog")) o.gemm_watchdog = value;
  else if (!std::strcmp(key, "gemm_tx_mode")) o.gemm_tx_mode = value;
  else if (!std::strcmp(key, "quant_rows")) o.quant_rows = value;
  else if (!std::strcmp(key, "quant_variant")) o.quant_variant = value;
  else if (!std::strcmp(key, "quant_ctas")) o.quant_ctas = value;
  else if (!std::strcmp(key, "gemm_ctas")) o.gemm_ctas = value;
  else if (!std::strcmp(key, "tp_debug")) o.tp_debug = value;
  else if (!std::strcmp(key, "gemm_splitk")) o.gemm_splitk = value;
  else if (!std::strcmp(key, "gemm_cta_group")) o.gemm_cta_group = value;
  else if (!std::strcmp(key, "gemm_debug_flags")) o.gemm_debug_flags = value;
  else if (!std::strcmp(key, "gemm_raster")) o.gemm_raster = value;
  else if (!std::strcmp(key, "pdl")) o.pdl = value;
  else if (!std::strcmp(key, "tp_reduce_ctas")) o.tp_reduce_ctas = value;
  else if (!std::strcmp(key, "tp_timeout_ms")) o.tp_timeout_ms = value;
  else {
    mmx::set_error("mmx_set_option: unknown key '%s'", key);
    return MMX_ERR_INVALID;
  }
  return MMX_OK;
}
