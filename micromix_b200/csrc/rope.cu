// rope.cu -- rotary position embedding applied IN PLACE to the q and k columns of the fused qkv GEMM output (sm_100a).
//
// The reference keeps RoPE in PyTorch (/root/reference/model/qLlamaLayer.py:25-54, 271-272: HF's apply_rotary_pos_emb,
//     q_embed = q * cos + rotate_half(q) * sin                     rotate_half(x) = cat(-x[..., d/2:], x[..., :d/2])
// on [b, heads, s, d] VIEWS of the projection outputs).  On transposed views those are eight un-vectorised elementwise
// kernels plus two torch.cat copies per layer -- a third of a Llama-3-8B prefill layer's device time on B200
// (profiles/r02_prefill_breakdown.txt) for what is one read and one write of q and k.  This kernel does exactly the same
// arithmetic, with the same bf16 roundings (each product and the sum are rounded to bf16, as the three torch ops do), on
// the rows of the GEMM output where they lie: HBM-bound byte work, one 16-byte load / store per lane and half-head.
#include <cuda_bf16.h>

#include "common.h"

namespace mmx {

struct RopeParams {
  uint16_t* y;          // bf16 [M, ld]: the first heads * d columns of every row are rotated
  const uint16_t* cos;  // bf16 [S, d]
  const uint16_t* sin;  // bf16 [S, d]
  int64_t M, ld, S;
  int heads, d;
};

// (the _rn forms: nvcc must NOT contract a product and the sum into one fused multiply-add -- the three torch ops round
// three times)
__device__ __forceinline__ uint32_t bf16x2_mul(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hmul2_rn(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ uint32_t bf16x2_add(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hadd2_rn(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}

// one thread = 8 channels i .. i+7 of the first half of one head AND their partners i + d/2 .. : both are read, both written
__global__ void __launch_bounds__(256) rope_inplace_kernel(const __grid_constant__ RopeParams p) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int half8 = p.d / 16;                       // 16-byte chunks per half head
  const int64_t per_row = (int64_t)p.heads * half8;  // threads per row
  const int64_t total = p.M * per_row;
  for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = w / per_row;
    const int r = (int)(w - m * per_row);
    const int h = r / half8, c = r - h * half8;
    uint16_t* lo = p.y + m * p.ld + (int64_t)h * p.d + c * 8;
    uint16_t* hi = lo + p.d / 2;
    const int64_t pos = m % p.S;
    const uint4 x1 = *reinterpret_cast<const uint4*>(lo), x2 = *reinterpret_cast<const uint4*>(hi);
    const uint4 c1 = __ldg(reinterpret_cast<const uint4*>(p.cos + pos * p.d + c * 8));
    const uint4 c2 = __ldg(reinterpret_cast<const uint4*>(p.cos + pos * p.d + p.d / 2 + c * 8));
    const uint4 s1 = __ldg(reinterpret_cast<const uint4*>(p.sin + pos * p.d + c * 8));
    const uint4 s2 = __ldg(reinterpret_cast<const uint4*>(p.sin + pos * p.d + p.d / 2 + c * 8));
    const uint32_t a1[4] = {x1.x, x1.y, x1.z, x1.w}, a2[4] = {x2.x, x2.y, x2.z, x2.w};
    const uint32_t k1[4] = {c1.x, c1.y, c1.z, c1.w}, k2[4] = {c2.x, c2.y, c2.z, c2.w};
    const uint32_t n1[4] = {s1.x, s1.y, s1.z, s1.w}, n2[4] = {s2.x, s2.y, s2.z, s2.w};
    uint32_t o1[4], o2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // first half : x1 * cos + (-x2) * sin      second half : x2 * cos + x1 * sin   (bf16 product, bf16 product, bf16 sum)
      o1[i] = bf16x2_add(bf16x2_mul(a1[i], k1[i]), bf16x2_mul(a2[i] ^ 0x80008000u, n1[i]));
      o2[i] = bf16x2_add(bf16x2_mul(a2[i], k2[i]), bf16x2_mul(a1[i], n2[i]));
    }
    *reinterpret_cast<uint4*>(lo) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    *reinterpret_cast<uint4*>(hi) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
  }
}

}  // namespace mmx

extern "C" __attribute__((visibility("default"))) int mmx_rope_inplace(void* y, int64_t ld, int64_t M, int heads, int head_dim,
                                                                    const void* cos, const void* sin, int64_t S, void* stream) {
  using namespace mmx;
  if (!y || !cos || !sin || M < 0 || heads <= 0 || head_dim <= 0 || (head_dim % 16) || S <= 0 || ld < (int64_t)heads * head_dim ||
      (ld % 8)) {
    set_error("rope_inplace: bad arguments (head_dim a multiple of 16, ld >= heads * head_dim and a multiple of 8)");
    return MMX_ERR_INVALID;
  }
  if (((uintptr_t)y | (uintptr_t)cos | (uintptr_t)sin) & 15) {
    set_error("rope_inplace: pointers must be 16-byte aligned");
    return MMX_ERR_INVALID;
  }
  if (M == 0) return MMX_OK;
  RopeParams p;
  p.y = static_cast<uint16_t*>(y);
  p.cos = static_cast<const uint16_t*>(cos);
  p.sin = static_cast<const uint16_t*>(sin);
  p.M = M;
  p.ld = ld;
  p.S = S;
  p.heads = heads;
  p.d = head_dim;
  const int64_t total = M * heads * (head_dim / 16);
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)sm_count() * 16;
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
  MMX_CUDA_TRY(cudaLaunchKernelEx(&cfg, rope_inplace_kernel, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MMX_OK;
}
