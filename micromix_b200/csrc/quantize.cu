// quantize.cu -- fused reorder (gather) + per-32 absmax + E8M0 scale + FP4/FP6/FP8 convert + pack, for sm_100a.
//
// Replaces /root/reference/mgemm/src/reorder.cu:94-269 (reorder_quantize_mixed_kernel) and :271-432
// (reorder_quantize_mxfp4_kernel) behind mmx_reorder_quantize_{x,w,w4}; results are bit-identical
// (codes, packing, scale bytes, scale-factor swizzle) -- see oracle/mmx_oracle.c for the restated semantics.
//
// Design (HBM-bound, nothing GEMM-shaped here):
//   * Persistent CTAs.  A work item is R rows {rb*128 + l + 32*(R*h+j), j<R}: the R rows whose scale bytes share
//     one 16-byte line of the 512-byte scale-factor atom, so scales leave the SM as 16-byte (R=4) / 8-byte (R=2)
//     stores instead of 1-byte scatters.
//   * The R rows are read with coalesced 32-bit loads and stored to shared memory channel-major, row-interleaved
//     (xs[c][j]); the permuted gather x[r, idx[c]] then costs ONE 8-byte (R=4) shared load per channel for all R
//     rows, i.e. R times fewer bank-conflicted gathers than a row-at-a-time kernel.  reorder_index is staged in
//     shared memory once per CTA (the reference re-reads it from global memory for every row).
//   * A thread owns 8 consecutive permuted channels (a quarter of a 32-group) of all R rows: absmax over packed
//     bf16x2 (two rows per instruction), two warp shuffles finish the group reduction.
//   * Scale: integer exponent arithmetic on the bf16 absmax -- byte = exp(amax) - log2floor(QMAX) + (mant > mant(QMAX)),
//     0x7E for an all-zero group -- provably equal to the reference's ceil(log2(amax/QMAX)) for every bf16 amax
//     (tests/test_oracle_pin.py::test_scale_rule_exhaustive).  Elements: x * 2^-e is exact in fp32, then the
//     hardware cvt.rn.satfinite.{e2m1x2,e3m2x2,e4m3x2}.f32 (RNE, saturating == the reference's clamp + software RNE).
//   * Packed codes are staged in shared memory per pass and leave as fully coalesced 16-byte stores.
#include <cuda_bf16.h>

#include "common.h"

namespace mmx {

struct QuantParams {
  const uint16_t* x;
  const int16_t* idx;
  int64_t rows;
  int64_t num_items;
  int K;
  int kseg[3];      // channels per segment
  int fmt[3];       // bits per code: 4 | 6 | 8
  int cend[3];      // cumulative channel ends
  int vend[3];      // cumulative packed-byte ends of the "virtual packed row" (all three segments back to back)
  int katoms[3];    // kseg / 128
  int64_t rowbytes[3];
  uint8_t* q[3];
  uint8_t* sf[3];
};

__device__ __forceinline__ uint32_t ld_stream_u32(const void* p) {
  uint32_t v;
  asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}

__device__ __forceinline__ void st_stream_v4(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

// 8 floats -> 8 E2M1 codes, element 0 in the low nibble of byte 0 (reorder.cu:30-33 PackFp4).
__device__ __forceinline__ uint32_t cvt8_e2m1(const float (&f)[8]) {
  uint32_t r;
  asm("{\n"
      ".reg .b8 b0, b1, b2, b3;\n"
      "cvt.rn.satfinite.e2m1x2.f32 b0, %2, %1;\n"
      "cvt.rn.satfinite.e2m1x2.f32 b1, %4, %3;\n"
      "cvt.rn.satfinite.e2m1x2.f32 b2, %6, %5;\n"
      "cvt.rn.satfinite.e2m1x2.f32 b3, %8, %7;\n"
      "mov.b32 %0, {b0, b1, b2, b3};\n"
      "}"
      : "=r"(r)
      : "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]), "f"(f[4]), "f"(f[5]), "f"(f[6]), "f"(f[7]));
  return r;
}

// 2 floats -> two 6-bit E3M2 codes, each in the low 6 bits of a byte (a low byte, b high byte).
__device__ __forceinline__ uint32_t cvt2_e3m2(float a, float b) {
  uint16_t h;
  asm("cvt.rn.satfinite.e3m2x2.f32 %0, %2, %1;" : "=h"(h) : "f"(a), "f"(b));
  return h;
}

__device__ __forceinline__ uint32_t cvt2_e4m3(float a, float b) {
  uint16_t h;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %2, %1;" : "=h"(h) : "f"(a), "f"(b));
  return h;
}

// four 6-bit codes (c0|c1<<8 in h01, c2|c3<<8 in h23) -> 24 bits, little-endian bit-contiguous (reorder.cu:54-63)
__device__ __forceinline__ uint32_t pack4_fp6(uint32_t h01, uint32_t h23) {
  return (h01 & 0x3fu) | ((h01 >> 2) & 0xfc0u) | ((h23 & 0x3fu) << 12) | ((h23 & 0x3f00u) << 10);
}

__device__ __forceinline__ int voff(const QuantParams& p, int c) {
  int v = 0;
#pragma unroll
  for (int s = 0; s < 3; ++s) {
    int n = min(max(c, 0), p.kseg[s]);
    v += (n * p.fmt[s]) >> 3;
    c -= p.kseg[s];
  }
  return v;
}

// One pass-step of a thread: 8 permuted channels x R rows -> packed codes in the staging buffer.
//   g[e][k]   : gathered data, channel e, row pair k (rows 2k | 2k+1 in the low | high half), bf16 bits
//   mult[k]   : bf16x2 multiplier 2^(127-byte) per row -> x * mult is exact (power of two), so HMUL2.BF16 is bit-safe
template <int FMT, int R>
__device__ __forceinline__ void convert_and_stage(const uint32_t (&g)[8][R / 2], const uint32_t (&mult)[R / 2],
                                                  uint8_t* dst, int row_stride) {
#pragma unroll
  for (int k = 0; k < R / 2; ++k) {
    uint32_t h[8];
    const __nv_bfloat162 m2 = *reinterpret_cast<const __nv_bfloat162*>(&mult[k]);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const __nv_bfloat162 v = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&g[e][k]), m2);
      h[e] = *reinterpret_cast<const uint32_t*>(&v);
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) f[e] = __uint_as_float(half ? (h[e] & 0xffff0000u) : (h[e] << 16));
      uint8_t* d = dst + (2 * k + half) * row_stride;
      if constexpr (FMT == 4) {
        *reinterpret_cast<uint32_t*>(d) = cvt8_e2m1(f);
      } else if constexpr (FMT == 6) {
        const uint32_t lo = pack4_fp6(cvt2_e3m2(f[0], f[1]), cvt2_e3m2(f[2], f[3]));
        const uint32_t hi = pack4_fp6(cvt2_e3m2(f[4], f[5]), cvt2_e3m2(f[6], f[7]));
        uint16_t* d16 = reinterpret_cast<uint16_t*>(d);
        d16[0] = (uint16_t)lo;
        d16[1] = (uint16_t)((lo >> 16) | (hi << 8));
        d16[2] = (uint16_t)(hi >> 8);
      } else {
        uint2 o;
        o.x = cvt2_e4m3(f[0], f[1]) | (cvt2_e4m3(f[2], f[3]) << 16);
        o.y = cvt2_e4m3(f[4], f[5]) | (cvt2_e4m3(f[6], f[7]) << 16);
        *reinterpret_cast<uint2*>(d) = o;
      }
    }
  }
}

__device__ __forceinline__ uint2 ld_stream_u64(const void* p) {
  uint2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint2 lds64(uint32_t a) {
  uint2 v;
  asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}

// R rows per item, T threads, NLD = register-prefetch depth (each thread holds NLD x R 8-byte loads of the NEXT item
// while it computes the current one; needs NLD * T * 4 >= K).
//
// Shared memory: idx_s[K] i16 | tabA[K/8] u32 | tabB[K/16] u32 | tabV[npass+1] u32 | xs | stage | sfs
//   tabA[o]  per channel octet o: format code (bits 30..31: 0 FP4, 1 FP6, 2 FP8) | byte offset of the octet's packed
//            codes relative to the start of its pass in the "virtual packed row" (all three segments back to back)
//   tabB[q]  per 16-byte chunk q of the virtual packed row: segment (bits 28..29) | byte offset inside that segment's row
//   tabV[ps] first 16-byte chunk of pass ps (tabV[npass] = total chunks)
// The tables are built once per CTA, so the per-item / per-pass code does no segment bookkeeping at all.
template <int R, int T, int NLD, int MINB>
__global__ void __launch_bounds__(T, MINB) reorder_quantize_kernel(const __grid_constant__ QuantParams p) {
  static_assert(R == 4 || R == 2, "rows per item");
  constexpr int HS = 4 / R;        // items per (128-row block, lane row l)
  constexpr int RW = R / 2;        // 32-bit words per gathered channel (two rows per word)
  constexpr int PASS_CH = 8 * T;   // channels per pass
  constexpr int TPR = T / R;       // copy-out threads per row
  extern __shared__ __align__(16) uint8_t smem[];
  const int K = p.K;
  const int npass = (K + PASS_CH - 1) / PASS_CH;
  int16_t* idx_s = reinterpret_cast<int16_t*>(smem);
  uint32_t* tabA = reinterpret_cast<uint32_t*>(smem + ((K * 2 + 15) & ~15));
  uint32_t* tabB = tabA + K / 8;
  uint32_t* tabV = tabB + K / 16;
  uint8_t* xs = reinterpret_cast<uint8_t*>(tabV) + 64;
  uint8_t* stage = xs + (size_t)K * R * 2;
  uint8_t* sfs = stage + 2 * R * PASS_CH;
  const uint32_t xs_a = smem_addr(xs);
  const int t = threadIdx.x;
  const int K4 = K >> 2;

  // ---- one-time per CTA: permutation and lookup tables
  for (int i = t; i < K / 8; i += T) reinterpret_cast<uint4*>(idx_s)[i] = reinterpret_cast<const uint4*>(p.idx)[i];
  for (int o = t; o < K / 8; o += T) {
    const int c = o * 8;
    const int sg = (c >= p.cend[1]) ? 2 : (c >= p.cend[0] ? 1 : 0);
    const int C0 = (c / PASS_CH) * PASS_CH;
    tabA[o] = ((uint32_t)((p.fmt[sg] - 4) >> 1) << 30) | (uint32_t)(voff(p, c) - voff(p, C0));  // 4|6|8 bits -> 0|1|2
  }
  for (int q = t; q < p.vend[2] / 16; q += T) {
    const int v = q * 16;
    const int sg = (v >= p.vend[1]) ? 2 : (v >= p.vend[0] ? 1 : 0);
    tabB[q] = ((uint32_t)sg << 28) | (uint32_t)(v - (sg ? p.vend[sg - 1] : 0));
  }
  if (t <= npass) tabV[t] = (uint32_t)(voff(p, min(t * PASS_CH, K)) >> 4);
  // (the first __syncthreads of the item loop publishes the tables)

  uint2 pre[NLD][R];  // prefetched rows of the next item

  // item -> first row (the other R-1 rows follow at +32 each), all in 32-bit arithmetic
  auto item_row0 = [&](int item, int& l, int& h, int& rb) -> int {
    l = item & 31;
    const int it2 = item >> 5;
    h = it2 % HS;
    rb = it2 / HS;
    return rb * 128 + l + 32 * (R * h);
  };
  auto prefetch = [&](int item) {
    int l, h, rb;
    const int row0 = item_row0(item, l, h, rb);
    if (row0 >= (int)p.rows) return;  // block-uniform: nothing to fetch for an item past the last row
    const uint16_t* b0 = p.x + (int64_t)row0 * K + 4 * t;
    const int64_t rstride = (int64_t)32 * K;
    const int nvalid = min(R, ((int)p.rows - 1 - row0) / 32 + 1);
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      if (i * T + t < K4) {
#pragma unroll
        for (int j = 0; j < R; ++j)  // rows past the end re-read row0; they are zeroed when stored to shared memory
          pre[i][j] = ld_stream_u64(b0 + (j < nvalid ? j * rstride : 0) + (size_t)i * T * 4);
      }
    }
  };

  const int num_items = (int)p.num_items;
  int item = blockIdx.x;
  if (item < num_items) prefetch(item);

  for (; item < num_items; item += gridDim.x) {
    int l, h, rb;
    const int row0 = item_row0(item, l, h, rb);
    if (row0 >= (int)p.rows) break;  // block-uniform; items are ordered by row, nothing valid follows for this CTA
    const int nvalid = min(R, ((int)p.rows - 1 - row0) / 32 + 1);

    // ---- prefetched registers -> shared memory, channel-major / row-interleaved
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int c4 = i * T + t;
      if (c4 < K4) {
        uint2 w[R];
#pragma unroll
        for (int j = 0; j < R; ++j) w[j] = pre[i][j];
        if (nvalid < R) {  // block-uniform; only the last, partly filled row block takes this path
#pragma unroll
          for (int j = 1; j < R; ++j)
            if (j >= nvalid) w[j] = make_uint2(0u, 0u);
        }
        if constexpr (R == 4) {
          uint4 o0, o1;
          o0.x = __byte_perm(w[0].x, w[1].x, 0x5410);
          o0.y = __byte_perm(w[2].x, w[3].x, 0x5410);
          o0.z = __byte_perm(w[0].x, w[1].x, 0x7632);
          o0.w = __byte_perm(w[2].x, w[3].x, 0x7632);
          o1.x = __byte_perm(w[0].y, w[1].y, 0x5410);
          o1.y = __byte_perm(w[2].y, w[3].y, 0x5410);
          o1.z = __byte_perm(w[0].y, w[1].y, 0x7632);
          o1.w = __byte_perm(w[2].y, w[3].y, 0x7632);
          uint4* d = reinterpret_cast<uint4*>(xs + (size_t)c4 * 32);
          d[0] = o0;
          d[1] = o1;
        } else {
          uint4 o;
          o.x = __byte_perm(w[0].x, w[1].x, 0x5410);
          o.y = __byte_perm(w[0].x, w[1].x, 0x7632);
          o.z = __byte_perm(w[0].y, w[1].y, 0x5410);
          o.w = __byte_perm(w[0].y, w[1].y, 0x7632);
          *reinterpret_cast<uint4*>(xs + (size_t)c4 * 16) = o;
        }
      }
    }
    __syncthreads();
    if (item + (int)gridDim.x < num_items) prefetch(item + (int)gridDim.x);  // in flight during the compute below

    // this thread's copy-out row and its three destination row pointers
    const int jrow = t / TPR;
    const bool jvalid = jrow < nvalid;
    const int64_t rj = (int64_t)row0 + 32 * jrow;
    uint8_t* const d0 = p.q[0] + rj * p.rowbytes[0];
    uint8_t* const d1 = p.q[1] + rj * p.rowbytes[1];
    uint8_t* const d2 = p.q[2] + rj * p.rowbytes[2];

    for (int ps = 0; ps < npass; ++ps) {
      const int oct = ps * T + t;
      const bool active = oct * 8 < K;
      const int oc = active ? oct : 0;
      const uint32_t ta = tabA[oc];
      const uint32_t fmtc = ta >> 30;
      uint8_t* stagebuf = stage + (ps & 1) * (R * PASS_CH);

      // ---- gather 8 permuted channels x R rows
      const uint4 iv = *reinterpret_cast<const uint4*>(idx_s + oc * 8);
      const uint32_t ivw[4] = {iv.x, iv.y, iv.z, iv.w};
      uint32_t g[8][RW];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const uint32_t ch = (e & 1) ? (ivw[e >> 1] >> 16) : (ivw[e >> 1] & 0xffffu);
        if constexpr (R == 4) {
          const uint2 v = lds64(xs_a + ch * 8);
          g[e][0] = v.x;
          g[e][1] = v.y;
        } else {
          g[e][0] = lds32(xs_a + ch * 4);
        }
      }

      // ---- absmax per row over the 32-group (8 local channels, then the 4 lanes of the team), scale byte, multiplier
      const uint32_t q2 = (fmtc == 0) ? 0x00020002u : ((fmtc == 1) ? 0x00040004u : 0x00080008u);  // log2floor(QMAX)
      const uint32_t add2 = (fmtc == 0) ? (63u * 0x00010001u) : (31u * 0x00010001u);            // 127 - thr, thr = 64 | 96
      uint32_t mult[RW];
      uint32_t sfb[RW];
#pragma unroll
      for (int k = 0; k < RW; ++k) {
        uint32_t u = g[0][k] & 0x7fff7fffu;
#pragma unroll
        for (int e = 1; e < 8; ++e) u = __vmaxu2(u, g[e][k] & 0x7fff7fffu);  // |bf16| orders like u16
        u = __vmaxu2(u, __shfl_xor_sync(0xffffffffu, u, 1));
        u = __vmaxu2(u, __shfl_xor_sync(0xffffffffu, u, 2));
        // byte = max(exp - qexp, 0) + (mant > thr), both rows of the pair at once; 0x7E for an all-zero group
        const uint32_t ex2 = (u >> 7) & 0x00ff00ffu;
        const uint32_t gt2 = (((u & 0x007f007fu) + add2) >> 7) & 0x00010001u;
        const uint32_t b2 = __vmaxu2(ex2, q2) - q2 + gt2;
        mult[k] = (0x00fe00feu - b2) << 7;  // bf16x2 of 2^(127-byte)
        const uint32_t z = __vcmpeq2(u, 0u);
        sfb[k] = (b2 & ~z) | (0x007e007eu & z);
      }

      if (active) {
        uint8_t* dst = stagebuf + (ta & 0x3fffffffu);
        if (fmtc == 0) convert_and_stage<4, R>(g, mult, dst, PASS_CH);
        else if (fmtc == 1) convert_and_stage<6, R>(g, mult, dst, PASS_CH);
        else convert_and_stage<8, R>(g, mult, dst, PASS_CH);
        if ((t & 3) == 0) {
          const int G = oct >> 2;  // 32-channel group index
          uint8_t* d = sfs + (G >> 2) * 16 + (R * h) * 4 + (G & 3);
#pragma unroll
          for (int k = 0; k < RW; ++k) {
            d[8 * k] = (uint8_t)sfb[k];
            d[8 * k + 4] = (uint8_t)(sfb[k] >> 16);
          }
        }
      }
      __syncthreads();

      // ---- coalesced 16-byte copy-out of this pass' packed codes: T/R threads per row
      if (jvalid) {
        const int q0 = (int)tabV[ps];
        const int n16 = (int)tabV[ps + 1] - q0;
        const uint8_t* src = stagebuf + jrow * PASS_CH;
        for (int ch = t - jrow * TPR; ch < n16; ch += TPR) {
          const uint32_t tb = tabB[q0 + ch];
          const uint32_t sg = tb >> 28;
          uint8_t* d = (sg == 0) ? d0 : ((sg == 1) ? d1 : d2);
          st_stream_v4(d + (tb & 0x0fffffffu), *reinterpret_cast<const uint4*>(src + 16 * ch));
        }
      }
    }

    // ---- scale factors: one 16-byte (R=4) / 8-byte (R=2) store per 128 channels
    for (int ch = t; ch < K / 128; ch += T) {
      const int c = ch * 128;
      const int sg = (c >= p.cend[1]) ? 2 : (c >= p.cend[0] ? 1 : 0);
      const int cb = (sg == 0) ? 0 : p.cend[sg - 1];
      const int ka = (c - cb) >> 7;
      uint8_t* dst = p.sf[sg] + ((int64_t)rb * p.katoms[sg] + ka) * 512 + l * 16;
      if constexpr (R == 4) {
        *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(sfs + ch * 16);
      } else {
        *reinterpret_cast<uint2*>(dst + 8 * h) = *reinterpret_cast<const uint2*>(sfs + ch * 16 + 8 * h);
      }
    }
    // the next item's __syncthreads (after its xs stores) orders these reads before sfs/stage reuse; xs itself is
    // only rewritten after the last pass' barrier, i.e. after every gather of this item
  }
}

template <int R, int T>
static size_t quant_smem_bytes(int K) {
  return ((size_t)(K * 2 + 15) & ~(size_t)15) + (size_t)(K / 8) * 4 + (size_t)(K / 16) * 4 + 64 /*tabV*/ +
         (size_t)K * R * 2 + (size_t)2 * R * 8 * T + (size_t)(K / 128) * 16;
}

template <int R, int T, int NLD, int MINB>
static int launch_quant(const QuantParams& p, cudaStream_t stream) {
  const size_t smem = quant_smem_bytes<R, T>(p.K);
  if (smem > 227 * 1024 || (int64_t)NLD * T * 4 < p.K) {
    set_error("reorder_quantize: K=%d does not fit the <%d,%d,%d> kernel (%zu bytes of shared memory)", p.K, R, T, NLD,
              smem);
    return MMX_ERR_INVALID;
  }
  auto kern = reorder_quantize_kernel<R, T, NLD, MINB>;
  static size_t attr_set = 0;
  if (smem > attr_set) {
    MMX_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = smem;
  }
  int occ = 0;
  MMX_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
  if (occ < 1) occ = 1;
  int64_t grid = (int64_t)sm_count() * occ;
  if (grid > p.num_items) grid = p.num_items;
  if (grid < 1) return MMX_OK;
  kern<<<(unsigned)grid, T, smem, stream>>>(p);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  MMX_CUDA_TRY(cudaGetLastError());
  return MMX_OK;
}

static int reorder_quantize(const void* x, int64_t rows, int K, const int16_t* idx, int KN, int KS, int KO,
                            const int fmt[3], uint8_t* q0, uint8_t* q1, uint8_t* q2, uint8_t* s0, uint8_t* s1,
                            uint8_t* s2, void* stream) {
  if (rows < 0 || K <= 0 || KN < 0 || KS < 0 || KO < 0 || KN + KS + KO != K) {
    set_error("reorder_quantize: bad shape rows=%lld K=%d (KN,KS,KO)=(%d,%d,%d)", (long long)rows, K, KN, KS, KO);
    return MMX_ERR_INVALID;
  }
  if ((KN % 128) || (KS % 128) || (KO % 128)) {
    set_error("reorder_quantize: KN, KS, KO must be multiples of 128, got (%d,%d,%d)", KN, KS, KO);
    return MMX_ERR_INVALID;
  }
  if (K > 32767) {
    set_error("reorder_quantize: K=%d exceeds the int16 reorder_index range", K);
    return MMX_ERR_INVALID;
  }
  uint8_t* q[3] = {q0, q1, q2};
  uint8_t* s[3] = {s0, s1, s2};
  const int ks[3] = {KN, KS, KO};
  if (!x || !idx) {
    set_error("reorder_quantize: null input pointer");
    return MMX_ERR_INVALID;
  }
  for (int i = 0; i < 3; ++i)
    if (ks[i] && (!q[i] || !s[i])) {
      set_error("reorder_quantize: null output pointer for non-empty segment %d", i);
      return MMX_ERR_INVALID;
    }
  if (((uintptr_t)x | (uintptr_t)idx | (uintptr_t)q0 | (uintptr_t)q1 | (uintptr_t)q2 | (uintptr_t)s0 | (uintptr_t)s1 |
       (uintptr_t)s2) & 15) {
    set_error("reorder_quantize: pointers must be 16-byte aligned");
    return MMX_ERR_INVALID;
  }
  if (rows == 0) return MMX_OK;
  QuantParams p;
  p.x = static_cast<const uint16_t*>(x);
  p.idx = idx;
  p.rows = rows;
  p.K = K;
  int cacc = 0, vacc = 0;
  for (int i = 0; i < 3; ++i) {
    p.kseg[i] = ks[i];
    p.fmt[i] = fmt[i];
    cacc += ks[i];
    vacc += ks[i] * fmt[i] / 8;
    p.cend[i] = cacc;
    p.vend[i] = vacc;
    p.katoms[i] = ks[i] / 128;
    p.rowbytes[i] = (int64_t)ks[i] * fmt[i] / 8;
    p.q[i] = q[i];
    p.sf[i] = s[i];
  }
  const int64_t rblocks = (rows + 127) / 128;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int force = (int)options().quant_rows;
  // configuration by K: rows per item R, threads T, prefetch depth NLD (NLD*T*4 >= K); smem grows with K*R
  if (force != 2 && K <= 4096) {
    p.num_items = rblocks * 32;
    return launch_quant<4, 256, 4, 3>(p, st);
  }
  if (force != 2 && K <= 8192) {
    p.num_items = rblocks * 32;
    return launch_quant<4, 512, 4, 2>(p, st);
  }
  p.num_items = rblocks * 64;
  if (K <= 4096) return launch_quant<2, 256, 4, 4>(p, st);
  if (K <= 16384) return launch_quant<2, 512, 8, 2>(p, st);
  return launch_quant<2, 512, 16, 1>(p, st);
}

}  // namespace mmx

extern "C" __attribute__((visibility("default"))) int mmx_reorder_quantize_x(const void* x, int64_t M, int K, const int16_t* idx, int KN, int KS, int KO,
                                      uint8_t* xn, uint8_t* xs, uint8_t* xo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo,
                                      void* stream) {
  const int fmt[3] = {4, 6, 8};
  return mmx::reorder_quantize(x, M, K, idx, KN, KS, KO, fmt, xn, xs, xo, sfn, sfs, sfo, stream);
}

extern "C" __attribute__((visibility("default"))) int mmx_reorder_quantize_w(const void* w, int64_t N, int K, const int16_t* idx, int KN, int KS, int KO,
                                      uint8_t* wn, uint8_t* ws, uint8_t* wo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo,
                                      void* stream) {
  const int fmt[3] = {4, 6, 8};
  return mmx::reorder_quantize(w, N, K, idx, KN, KS, KO, fmt, wn, ws, wo, sfn, sfs, sfo, stream);
}

extern "C" __attribute__((visibility("default"))) int mmx_reorder_quantize_w4(const void* w, int64_t N, int K, const int16_t* idx, int KN, int KS, int KO,
                                       uint8_t* wn, uint8_t* ws, uint8_t* wo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo,
                                       void* stream) {
  const int fmt[3] = {4, 4, 4};
  return mmx::reorder_quantize(w, N, K, idx, KN, KS, KO, fmt, wn, ws, wo, sfn, sfs, sfo, stream);
}
