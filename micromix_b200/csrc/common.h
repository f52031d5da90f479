// common.h -- shared host-side helpers of libmicromix_b200.so (error text, launch counter, options).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "micromix_b200.h"

namespace mmx {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

struct Options {
  int64_t gemm_watchdog = 0;  // 1: bounded mbarrier spins, status words written to the debug buffer
  int64_t gemm_tx_mode = 0;   // 0: TMA tx bytes = packed gmem bytes (FP4 64B/row, FP6 96B/row); 1: smem footprint
  int64_t quant_rows = 0;     // 0: auto (two rows per item), 4: force the four-rows-per-item kernel for K <= 4096
  int64_t quant_variant = 0; // tuning sweeps only: alternative table encoding / occupancy bound of the quantize kernel
  int64_t quant_ctas = 0;     // 0: SMs x occupancy, else force the persistent grid size
  int64_t gemm_ctas = 0;      // 0: one CTA per SM, else force the persistent grid size
  int64_t gemm_debug_flags = 0;  // watchdog build: timing experiments (results are wrong), see GemmParams::flags
  int64_t pdl = 1;            // 1: kernels are launched with programmatic stream serialization (prologue overlap)
  int64_t gemm_raster = 0;    // 0: auto, 1: force M-fastest tile order, 2: force N-fastest
  int64_t gemm_cta_group = 0; // 0: auto (pairs when M > 128), 1: force the single-CTA kernel
};
Options& options();

inline int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: %s", what, cudaGetErrorString(e));
  return MMX_ERR_CUDA;
}

#define MMX_CUDA_TRY(expr)                                    \
  do {                                                        \
    cudaError_t _e = (expr);                                  \
    if (_e != cudaSuccess) return ::mmx::cuda_fail(_e, #expr); \
  } while (0)

int sm_count();       // cached multiprocessor count of the current device
bool device_is_sm100();

}  // namespace mmx
