// moe.cu -- the expert-combine of a Mixtral sparse-MoE block, for sm_100a (HBM-bound byte work).
//
// Replaces the tail of the reference's per-expert loop (/root/reference/model/qMixtralLayer.py:446-450):
//     current_hidden_states = expert(current_state) * routing_weights[top_x, idx, None]
//     final_hidden_states.index_add_(0, top_x, current_hidden_states.to(hidden_states.dtype))
// i.e. for every token, in ascending expert order, out = bf16(out + bf16(y * w)) starting from zero.  With the experts'
// outputs produced by ONE grouped GEMM over the expert-sorted token matrix (gemm.cu, mmx_matmul_grouped), the combine is
// a destination-driven gather-reduce: one warp per token row reads the <= top_k rows of y that belong to it (only those
// that live on this rank) with 128-bit loads and writes the output row once -- no atomics, no zero-fill pass, no
// [tokens * top_k, hidden] temporary, and exactly the reference's rounding sequence.
#include <cuda_bf16.h>

#include "common.h"

namespace mmx {

constexpr int kMaxTopK = 8;

struct CombineParams {
  const uint4* y;
  const int* row;
  const int* expert;
  const uint16_t* w;
  uint4* out;
  int64_t T;
  int top_k, H8;  // H8 = hidden / 8 (uint4 per row)
};

__device__ __forceinline__ uint4 ld_nc_u4(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// acc = bf16(acc + bf16(v * w)) on both halves of the word, the reference's two roundings
__device__ __forceinline__ uint32_t mul_add_bf16x2(uint32_t acc, uint32_t v, float w, bool first) {
  const float v0 = __uint_as_float(v << 16), v1 = __uint_as_float(v & 0xffff0000u);
  const __nv_bfloat162 p = __floats2bfloat162_rn(__fmul_rn(v0, w), __fmul_rn(v1, w));
  const uint32_t pb = *reinterpret_cast<const uint32_t*>(&p);
  if (first) {
    // 0 + p: exact, but -0 + 0 = +0 as in index_add_ into a zero buffer
    const float s0 = __fadd_rn(0.0f, __uint_as_float(pb << 16)), s1 = __fadd_rn(0.0f, __uint_as_float(pb & 0xffff0000u));
    const __nv_bfloat162 r = __floats2bfloat162_rn(s0, s1);
    return *reinterpret_cast<const uint32_t*>(&r);
  }
  const float a0 = __uint_as_float(acc << 16), a1 = __uint_as_float(acc & 0xffff0000u);
  const __nv_bfloat162 r = __floats2bfloat162_rn(__fadd_rn(a0, __uint_as_float(pb << 16)),
                                                 __fadd_rn(a1, __uint_as_float(pb & 0xffff0000u)));
  return *reinterpret_cast<const uint32_t*>(&r);
}

__global__ void __launch_bounds__(256) moe_combine_kernel(const __grid_constant__ CombineParams p) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < p.T; t += warps) {
    // this token's slots, sorted by expert id (insertion sort, top_k <= 8): the loop order of the reference
    int rows[kMaxTopK], exps[kMaxTopK];
    float ws[kMaxTopK];
    int n = 0;
    for (int s = 0; s < p.top_k; ++s) {
      const int r = __ldg(p.row + t * p.top_k + s);
      if (r < 0) continue;  // that expert lives on another rank
      const int e = __ldg(p.expert + t * p.top_k + s);
      const float w = __uint_as_float((uint32_t)__ldg(p.w + t * p.top_k + s) << 16);
      int i = n++;
      while (i > 0 && exps[i - 1] > e) {
        rows[i] = rows[i - 1];
        exps[i] = exps[i - 1];
        ws[i] = ws[i - 1];
        --i;
      }
      rows[i] = r;
      exps[i] = e;
      ws[i] = w;
    }
    uint4* o = p.out + t * p.H8;
    for (int c = lane; c < p.H8; c += 32) {
      uint4 acc = make_uint4(0, 0, 0, 0);
      for (int i = 0; i < n; ++i) {
        const uint4 v = ld_nc_u4(p.y + (int64_t)rows[i] * p.H8 + c);
        acc.x = mul_add_bf16x2(acc.x, v.x, ws[i], i == 0);
        acc.y = mul_add_bf16x2(acc.y, v.y, ws[i], i == 0);
        acc.z = mul_add_bf16x2(acc.z, v.z, ws[i], i == 0);
        acc.w = mul_add_bf16x2(acc.w, v.w, ws[i], i == 0);
      }
      o[c] = acc;
    }
  }
}

}  // namespace mmx

extern "C" __attribute__((visibility("default"))) int mmx_moe_combine(const void* y, const int32_t* row, const int32_t* expert,
                                                                   const void* w, int64_t T, int top_k, int H, void* out,
                                                                   void* stream) {
  using namespace mmx;
  if (!y || !row || !expert || !w || !out || T < 0 || top_k < 1 || top_k > kMaxTopK || H <= 0 || (H % 8)) {
    set_error("moe_combine: bad arguments (top_k <= %d, hidden a multiple of 8)", kMaxTopK);
    return MMX_ERR_INVALID;
  }
  if (((uintptr_t)y | (uintptr_t)out) & 15) {
    set_error("moe_combine: y and out must be 16-byte aligned");
    return MMX_ERR_INVALID;
  }
  if (T == 0) return MMX_OK;
  CombineParams p;
  p.y = static_cast<const uint4*>(y);
  p.row = row;
  p.expert = expert;
  p.w = static_cast<const uint16_t*>(w);
  p.out = static_cast<uint4*>(out);
  p.T = T;
  p.top_k = top_k;
  p.H8 = H / 8;
  int64_t blocks = (T + 7) / 8;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3(256);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = options().pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MMX_CUDA_TRY(cudaLaunchKernelEx(&cfg, moe_combine_kernel, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MMX_OK;
}

// ------------------------------------------------------------------------------------------------ routing tables
// The device-side routing tables of the grouped expert path (qMixtralLayer.route_tables restated as ONE kernel; the torch
// version is a stable argsort, three scatters, two cumsums and two searchsorted -- about twenty launches per block).
// A stable counting sort by local expert slot: the (token, slot) pairs of each local expert keep their token order (the
// order torch.where(sel == e) gives the reference's loop, /root/reference/model/qMixtralLayer.py:437-450), every expert's
// rows are padded to whole m-tiles.  One CTA: thread t owns a contiguous run of pairs; per-slot exclusive scans over the
// threads give every run its first row.
namespace mmx {

constexpr int kRouteThreads = 1024;
constexpr int kRouteWarps = kRouteThreads / 32;
constexpr int kRouteMaxLocal = 64;
constexpr int kRouteMaxExperts = 256;

struct RouteParams {
  const long long* sel;         // [T * k] expert ids
  const long long* local_slot;  // [E]: slot of expert e on this rank, n_local = elsewhere
  int n_pairs, k, n_local, n_experts, tile, Mp;
  int* row_src;     // [Mp]
  int* pair_row;    // [T * k]
  int* grp_rowblk;  // [Mp / 128]
  int* grp_mtile;   // [Mp / tile]
  int* rows_used;   // [1]
};

// Warp w owns the contiguous pairs [w * per_warp, (w + 1) * per_warp) and walks them 32 at a time (coalesced loads and
// stores); the stable rank of a pair inside its slot = pairs of that slot in earlier warps + in earlier iterations of this
// warp + in lower lanes of this iteration (match_any + popc).
__global__ void __launch_bounds__(kRouteThreads) moe_route_kernel(const __grid_constant__ RouteParams p) {
  __shared__ int s_w[kRouteWarps][kRouteMaxLocal];  // pass 1: pairs per (warp, slot); then the warp's running first rank
  __shared__ int s_start[kRouteMaxLocal], s_end[kRouteMaxLocal];
  __shared__ int s_slot[kRouteMaxExperts];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  for (int i = t; i < kRouteWarps * kRouteMaxLocal; i += kRouteThreads) (&s_w[0][0])[i] = 0;
  // the expert -> slot map in shared memory; experts beyond the table live elsewhere
  for (int e = t; e < kRouteMaxExperts; e += kRouteThreads) s_slot[e] = e < p.n_experts ? (int)__ldg(p.local_slot + e) : p.n_local;
  asm volatile("griddepcontrol.wait;" ::: "memory");  // sel comes from the router's top-k
  __syncthreads();
  const int per_warp = ((p.n_pairs + kRouteWarps - 1) / kRouteWarps + 31) & ~31;
  const int lo = min(p.n_pairs, warp * per_warp), hi = min(p.n_pairs, lo + per_warp);
  auto slot_of = [&](int i) {
    if (i >= hi) return p.n_local;
    const long long e = __ldg(p.sel + i);
    return (e >= 0 && e < kRouteMaxExperts) ? s_slot[(int)e] : p.n_local;
  };
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int slot = slot_of(i0 + lane);
    const unsigned same = __match_any_sync(0xffffffffu, slot);
    if (slot < p.n_local && (same & ((1u << lane) - 1u)) == 0) s_w[warp][slot] += __popc(same);  // the slot's first lane
    __syncwarp();
  }
  __syncthreads();
  // per slot: exclusive scan over the warps, padded extent
  if (t < p.n_local) {
    int acc = 0;
    for (int w = 0; w < kRouteWarps; ++w) {
      const int c = s_w[w][t];
      s_w[w][t] = acc;
      acc += c;
    }
    s_end[t] = acc;  // pairs of slot t, for now
  }
  __syncthreads();
  if (t == 0) {
    int acc = 0;
    for (int s = 0; s < p.n_local; ++s) {
      const int padded = (s_end[s] + p.tile - 1) / p.tile * p.tile;
      s_start[s] = acc;
      acc += padded;
      s_end[s] = acc;  // end of the slot's padded rows
    }
    *p.rows_used = acc;
  }
  __syncthreads();
  for (int r = t; r < p.Mp; r += kRouteThreads) p.row_src[r] = 0;  // padding rows read token 0 (any valid row)
  for (int b = t; b < p.Mp / 128; b += kRouteThreads) {
    int g = 0;
    while (g < p.n_local && s_end[g] <= b * 128) ++g;
    p.grp_rowblk[b] = min(g, p.n_local - 1);
  }
  for (int m = t; m < p.Mp / p.tile; m += kRouteThreads) {
    int g = 0;
    while (g < p.n_local && s_end[g] <= m * p.tile) ++g;
    p.grp_mtile[m] = g < p.n_local ? g : -1;
  }
  __syncthreads();  // the zero fill of row_src is ordered before the scatter below (same CTA)
  for (int i0 = lo; i0 < hi; i0 += 32) {
    const int i = i0 + lane;
    const int slot = slot_of(i);
    const unsigned same = __match_any_sync(0xffffffffu, slot);
    const unsigned below = same & ((1u << lane) - 1u);
    if (slot < p.n_local) {
      const int dst = s_start[slot] + s_w[warp][slot] + __popc(below);
      p.row_src[dst] = i / p.k;
      p.pair_row[i] = dst;
    } else if (i < hi) {
      p.pair_row[i] = -1;
    }
    __syncwarp();
    if (slot < p.n_local && below == 0) s_w[warp][slot] += __popc(same);
    __syncwarp();
  }
}

}  // namespace mmx

extern "C" __attribute__((visibility("default"))) int mmx_moe_route(const int64_t* sel, const int64_t* local_slot, int64_t T,
                                                                 int top_k, int experts, int n_local, int tile, int64_t Mp,
                                                                 int32_t* row_src,
                                                                 int32_t* pair_row, int32_t* grp_rowblk, int32_t* grp_mtile,
                                                                 int32_t* rows_used, void* stream) {
  using namespace mmx;
  if (!sel || !local_slot || !row_src || !pair_row || !grp_rowblk || !grp_mtile || !rows_used || T < 0 || top_k < 1 ||
      n_local < 1 || n_local > kRouteMaxLocal || experts < 1 || experts > kRouteMaxExperts || (tile != 128 && tile != 256) || Mp <= 0 || (Mp % tile) ||
      T * top_k > 0x7fffffff || Mp > 0x7fffffff || Mp < (T * top_k + (int64_t)n_local * (tile - 1)) / tile * tile) {
    set_error("moe_route: bad arguments (1 <= n_local <= %d, experts <= %d, tile 128 | 256, Mp a multiple of the tile that "
              "holds every padded expert)", kRouteMaxLocal, kRouteMaxExperts);
    return MMX_ERR_INVALID;
  }
  RouteParams p;
  p.sel = reinterpret_cast<const long long*>(sel);
  p.local_slot = reinterpret_cast<const long long*>(local_slot);
  p.n_pairs = (int)(T * top_k);
  p.k = top_k;
  p.n_local = n_local;
  p.n_experts = experts;
  p.tile = tile;
  p.Mp = (int)Mp;
  p.row_src = row_src;
  p.pair_row = pair_row;
  p.grp_rowblk = grp_rowblk;
  p.grp_mtile = grp_mtile;
  p.rows_used = rows_used;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(kRouteThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = options().pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MMX_CUDA_TRY(cudaLaunchKernelEx(&cfg, moe_route_kernel, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MMX_OK;
}
