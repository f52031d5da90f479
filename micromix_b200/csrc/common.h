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
  int64_t gemm_splitk = 0;    // M <= 128: 0 auto, 1 never split K, 2 | 4 | 8 force that cluster size (if it fits)
  int64_t tp_reduce_ctas = 0; // 0: one reducer CTA per SM, else cap the tile_allreduce_kernel grid (single-GPU tests)
  int64_t tp_debug = 0;       // timing experiments of the fused all-reduce (results wrong): see ReduceParams::dbg
  int64_t tp_timeout_ms = 10000;  // bound on every cross-rank spin of tile_allreduce_kernel (a lost peer cannot hang the GPU)
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

constexpr int kMaxTp = 8;  // ranks of one NVSwitch box

// Function attributes and __device__ symbol addresses are PER DEVICE: a process that drives several GPUs (the
// reference's own model/parallel_utils.py places decoder layers on different GPUs of one process) needs them cached per
// device, not per process.
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= kMaxDevices) d = 0;
  return d;
}

// Fused row-parallel GEMM -> all-reduce: what tp_reduce.cu hands to gemm.cu's matmul_impl (see RsParams in gemm.cu).
struct RsLaunch {
  const void* dst_maps;      // CUtensorMap[tp]: MY slot in rank d's staging buffer as bf16 [own_tiles_cap * 256, 256]
  uint32_t* tile_flags[kMaxTp];  // rank d's arrival counters for the tiles it owns (this call's parity)
  int tp, rank;
  int64_t own_tiles_cap;     // staging tiles (256 rows each) per (rank, slot)
  // filled in by matmul_impl: the tile geometry the reducer must mirror
  int cg, m_tiles, n_tiles, n_fastest;
  int rot_s;                 // owner rotation period, see RsParams in gemm.cu
  int pull;                  // 1: in-switch reduction -- the epilogue stores the partial tile into c_local (this rank's own
  void* c_local;             //    C of the call's parity, ordinary [M, N] coordinates) and only the arrival goes to the owner
  int shard;                 // 1: REDUCE-SCATTER -- owner-interleaved tile order (GemmParams::n_fastest == 2), rank o owns the
  int m_per;                 //    m-tiles [o * m_per, (o+1) * m_per) (filled in by matmul_impl) and keeps only those rows
};

// Optional extras of matmul_impl (all null / zero = the plain GEMM).
struct MatmulExtra {
  // grouped GEMM: A rows sorted by group and padded per group to whole m-tiles of `grp_tile_rows` (128 | 256) rows, B = the
  // groups' weights stacked on N (grp_n rows each); grp_mblk[m-tile] = group or -1 (device memory, written on the stream)
  const int* grp_mblk = nullptr;
  int grp_n = 0;
  int grp_count = 0;
  int grp_tile_rows = 0;
  const int* rows_dev = nullptr;  // optional (device memory): rows of A that exist, a multiple of the m-tile
  // gathered A (sequence-parallel hand-over): see GemmParams::ag_*
  const uint32_t* ag_arrived = nullptr;
  uint32_t* ag_taken = nullptr;
  int ag_rows = 0;
  uint32_t* ag_ticket = nullptr;
  uint32_t* ag_consumed[kMaxTp] = {};
  int ag_tp = 0;
  uint32_t* ag_err = nullptr;  // local error word (mmx_tp_status): bit 3 = a source rank's rows never arrived
  // fused SiLU(gate) * up + MX quantize (see GemmParams::act_*): outputs and the FP4 | FP6 | FP8 split of the activation
  // (N == 2 * (act_k[0] + act_k[1] + act_k[2]), B rows interleaved gate | up per 128 channels); act_q[0] != nullptr enables it
  uint8_t* act_q[3] = {};
  uint8_t* act_sf[3] = {};
  int act_k[3] = {};
  const void* residual = nullptr;  // bf16 [M, N]: c = bf16(residual + product), see GemmParams::residual
  const void* rope_cos = nullptr;  // rotary embedding in the epilogue, see GemmParams::rope_cos
  const void* rope_sin = nullptr;
  int rope_S = 0;
  int rope_cols = 0;
};
int matmul_impl(const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs, const uint8_t* ao,
                const uint8_t* bo, const uint8_t* sfan, const uint8_t* sfbn, const uint8_t* sfas, const uint8_t* sfbs,
                const uint8_t* sfao, const uint8_t* sfbo, int64_t M, int64_t N, int KN, int KS, int KO, int w4,
                const void* bias, void* c, void* stream, RsLaunch* rsl, const MatmulExtra* ex = nullptr);
// Sequence-parallel hand-over: the quantizer's outputs are multicast addresses; see QuantParams::ag_* in quantize.cu.
struct QuantGather {
  const uint32_t* consumed;  // local: bumped once per rank when that rank is done reading the previous gather
  uint32_t* issued;          // local: gathers this rank has issued so far
  uint32_t* arrived[kMaxTp]; // rank d's (peer-mapped) count of gathers whose rows from THIS rank have landed there
  uint32_t* err;             // local error word (mmx_tp_status): bit 2 = the wait for the consumers timed out
  int tp;
  // all-to-all form (see QuantParams::a2a_per): rows per destination rank (0 = gather form), per-destination code / scale
  // base pointers (offset to this rank's columns / atoms), packed-row pitch and scale atoms per row block of the destination
  int a2a_per;
  uint8_t* qd[kMaxTp][3];
  uint8_t* sfd[kMaxTp][3];
  uint32_t a2a_pitch[3];
  int a2a_katoms[3];
};
// quantize.cu: the reorder+quantize launcher behind mmx_reorder_quantize_* (fmt = bits per segment; norm_w: fused RMSNorm)
int reorder_quantize(const void* x, int64_t rows, int K, const int16_t* idx, int KN, int KS, int KO, const int fmt[3],
                     uint8_t* q0, uint8_t* q1, uint8_t* q2, uint8_t* s0, uint8_t* s1, uint8_t* s2, void* stream,
                     const void* norm_w, float eps, bool norm, const QuantGather* ag, const int* grp_rowblk = nullptr,
                     const int* row_src = nullptr, const int* rows_dev = nullptr);
int encode_store_tmap(void* ptr, int64_t rows, int64_t cols, void* out /* CUtensorMap* */);

int sm_count();       // cached multiprocessor count of the current device
bool device_is_sm100();

}  // namespace mmx
