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
