// quantize.cu -- fused reorder (gather) + per-32 absmax + E8M0 scale + FP4/FP6/FP8 convert + pack, for sm_100a.
//
// Replaces /root/reference/mgemm/src/reorder.cu:94-269 (reorder_quantize_mixed_kernel) and :271-432
// (reorder_quantize_mxfp4_kernel) behind mmx_reorder_quantize_{x,w,w4}; results are bit-identical
// (codes, packing, scale bytes, scale-factor swizzle) -- see oracle/mmx_oracle.c for the restated semantics.
//
// Design (HBM-bound byte work; the enemy is the instruction count per element, not arithmetic):
//   * Persistent CTAs.  A work item is R rows {rb*128 + l + 32*(R*h+j), j<R}: the R rows whose scale bytes share
//     one 16-byte line of the 512-byte scale-factor atom.
//   * SCATTER ON WRITE.  The permutation is applied while the rows go into shared memory, not when they are read
//     back: each thread loads 16 bytes (8 original channels) of each of the R rows with coalesced 128-bit loads,
//     interleaves the rows with PRMT (one slot = the R values of one channel, 2R bytes) and stores every slot at the
//     position of its PERMUTED channel (inverse permutation, built once per CTA as a table of swizzled byte
//     offsets).  The read side is then perfectly sequential: a thread owns 16 consecutive permuted channels (half a
//     32-group) of all R rows and fetches them with conflict-free 128-bit shared loads (XOR swizzle on the 16-byte
//     chunk index), with no index loads and no address arithmetic in the loop.
//   * Register prefetch: the global loads of the NEXT item are issued right after the scatter of the current one and
//     stay in flight during its compute phase.
//   * absmax: max.xorsign.abs.bf16x2 (one instruction per two elements, two rows at once), one warp shuffle finishes
//     the 32-group.  Scale: integer exponent arithmetic on the bf16 absmax --
//     byte = exp(amax) - log2floor(QMAX) + (mant > mant(QMAX)), 0x7E for an all-zero group -- provably equal to the
//     reference's ceil(log2(amax/QMAX)) for every bf16 amax (tests/test_oracle_pin.py::test_scale_rule_exhaustive).
//     Elements: x * 2^-e is exact (bf16x2 multiply by a power of two), then the hardware
//     cvt.rn.satfinite.{e2m1x2,e3m2x2,e4m3x2}.f32 (RNE, saturating == the reference's clamp + software RNE).
//   * Packed codes leave straight from registers: 8 (FP4) / 12 (FP6) / 16 (FP8) contiguous bytes per thread and row,
//     so a warp writes 256 / 384 / 512 contiguous bytes per row -- full sectors, no staging pass.
#include <cuda_bf16.h>

#include "common.h"

namespace mmx {

struct QuantParams {
  const uint16_t* x;
  const int16_t* idx;
  int64_t rows;
  int num_items;
  int K;
  unsigned int* sched;  // {next dynamic item, finished CTAs}: both zero at launch, reset by the last CTA to finish
  int fmt[3];       // bits per code: 4 | 6 | 8
  int cend[3];      // cumulative channel ends
  int katoms[3];    // kseg / 128
  int64_t rowbytes[3];
  uint8_t* q[3];
  uint8_t* sf[3];
};

__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sts64(uint32_t a, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(a), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t x) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(a), "r"(x) : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void stg_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void stg_v2(void* p, uint32_t a, uint32_t b) {
  asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void stg_b32(void* p, uint32_t a) {
  asm volatile("st.global.L1::no_allocate.b32 [%0], %1;" ::"l"(p), "r"(a) : "memory");
}

// max(|a|, |b|) on both bf16 halves; the sign bits of the result are garbage (xor of the input signs)
__device__ __forceinline__ uint32_t absmax2(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("max.xorsign.abs.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}

// 8 floats -> 8 E2M1 codes, element 0 in the low nibble of byte 0 (reorder.cu:30-33 PackFp4).
__device__ __forceinline__ uint32_t cvt8_e2m1(const float* f) {
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

// 4 floats -> four E3M2 codes, one per byte (6 bits each, upper two bits zero)
__device__ __forceinline__ uint32_t cvt4_e3m2(const float* f) {
  uint32_t r;
  asm("{\n"
      ".reg .b16 h0, h1;\n"
      "cvt.rn.satfinite.e3m2x2.f32 h0, %2, %1;\n"
      "cvt.rn.satfinite.e3m2x2.f32 h1, %4, %3;\n"
      "mov.b32 %0, {h0, h1};\n"
      "}"
      : "=r"(r)
      : "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]));
  return r;
}

// 4 floats -> four E4M3 codes, one per byte
__device__ __forceinline__ uint32_t cvt4_e4m3(const float* f) {
  uint32_t r;
  asm("{\n"
      ".reg .b16 h0, h1;\n"
      "cvt.rn.satfinite.e4m3x2.f32 h0, %2, %1;\n"
      "cvt.rn.satfinite.e4m3x2.f32 h1, %4, %3;\n"
      "mov.b32 %0, {h0, h1};\n"
      "}"
      : "=r"(r)
      : "f"(f[0]), "f"(f[1]), "f"(f[2]), "f"(f[3]));
  return r;
}

// four 6-bit codes, one per byte -> 24 bits, little-endian bit-contiguous (reorder.cu:54-63): c0 | c1<<6 | c2<<12 | c3<<18
__device__ __forceinline__ uint32_t squeeze4_fp6(uint32_t w) {
  const uint32_t a = w & 0x00ff00ffu;          // c0, c2
  const uint32_t b = (w >> 8) & 0x00ff00ffu;   // c1, c3
  const uint32_t x = b * 64u + a;              // [c0 | c1<<6] in bits 0..11, [c2 | c3<<6] in bits 16..27
  return ((x >> 4) & 0xfffff000u) | (x & 0xfffu);
}

// One row of a thread's 16 channels: 16 floats -> packed codes -> global memory (8 | 12 | 16 contiguous bytes).
template <int FMT>
__device__ __forceinline__ void convert_store_row(const float (&f)[16], uint8_t* dst) {
  if constexpr (FMT == 4) {
    stg_v2(dst, cvt8_e2m1(&f[0]), cvt8_e2m1(&f[8]));
  } else if constexpr (FMT == 6) {
    const uint32_t y0 = squeeze4_fp6(cvt4_e3m2(&f[0])), y1 = squeeze4_fp6(cvt4_e3m2(&f[4]));
    const uint32_t y2 = squeeze4_fp6(cvt4_e3m2(&f[8])), y3 = squeeze4_fp6(cvt4_e3m2(&f[12]));
    stg_b32(dst, __byte_perm(y0, y1, 0x4210));
    stg_b32(dst + 4, __byte_perm(y1, y2, 0x5421));
    stg_b32(dst + 8, __byte_perm(y2, y3, 0x6542));
  } else {
    stg_v4(dst, cvt4_e4m3(&f[0]), cvt4_e4m3(&f[4]), cvt4_e4m3(&f[8]), cvt4_e4m3(&f[12]));
  }
}

// 16 channels x R rows of one thread: scale (x * 2^-e, exact), convert, pack, store.
//   g[c][k]  : channel c, row pair k (rows 2k | 2k+1 in the low | high half), bf16 bits
//   mult[k]  : bf16x2 multiplier 2^(127-byte) per row -> x * mult is exact (power of two), so HMUL2.BF16 is bit-safe
//   dst      : first row's destination; row j lives at dst + j * row_stride; rows >= nstore are not stored
template <int FMT, int R, bool FULL>
__device__ __forceinline__ void convert_store(const uint32_t (&g)[16][R / 2], const uint32_t (&mult)[R / 2], uint8_t* dst,
                                              int64_t row_stride, int nstore) {
#pragma unroll
  for (int k = 0; k < R / 2; ++k) {
    uint32_t h[16];
    const __nv_bfloat162 m2 = *reinterpret_cast<const __nv_bfloat162*>(&mult[k]);
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const __nv_bfloat162 v = __hmul2(*reinterpret_cast<const __nv_bfloat162*>(&g[c][k]), m2);
      h[c] = *reinterpret_cast<const uint32_t*>(&v);
    }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float f[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) f[c] = __uint_as_float(half ? (h[c] & 0xffff0000u) : (h[c] << 16));
      if (FULL || 2 * k + half < nstore) convert_store_row<FMT>(f, dst);
      dst += row_stride;
    }
  }
}

// What a thread needs to know about one of its compute units (16 permuted channels); item-invariant, so it is
// computed once per CTA and kept in three registers per pass.
struct UnitCtx {
  uint32_t meta;  // bits 0..3 fmt (4|6|8), bits 4..5 segment, bit 6 active, bit 7 writes the group's scale bytes
  uint32_t qoff;  // byte offset of the unit's codes inside a packed row of its segment
  uint32_t sfo;   // byte offset of the unit's group inside the row block's scale atoms: (G/4)*512 + G%4
  uint32_t xoff;  // byte offset of the unit's first (swizzled) 16-byte chunk inside xs
};

// R rows per item, T threads, NLD = 16-byte chunks per thread and row (NLD * T * 8 >= K), NP = compute passes
// (NP * T * 16 >= K), WIDE = table offsets in 4-byte units (K * 2R > 65536).
//
// Shared memory: tab[K] u16 | xs[K] slots of 2R bytes
//   tab[c]  byte offset (WIDE: /4) inside xs of the slot of ORIGINAL channel c, i.e. of permuted position j with
//           idx[j] == c, after the chunk swizzle.  Built once per CTA.
//   xs      slot j holds the R rows of permuted channel j; 16-byte chunk p = j * 2R / 16 is stored at chunk
//           p ^ ((p >> 3) & SWM) so that the lane-strided 128-bit reads of the compute phase are conflict-free.
template <int R, int NLD, int NP, bool WIDE>
struct QuantKernel {
  static constexpr int RW = R / 2;          // 32-bit words per slot (two rows per word)
  static constexpr int SLOT = 2 * R;        // bytes per slot
  static constexpr int CPT = SLOT;          // 16-byte chunks holding a thread's 16 channels: 8 (R=4) | 4 (R=2)
  static constexpr uint32_t SWM = (R == 4) ? 7u : 3u;

  // item -> first row: 32 consecutive items share a block of 32R rows, item i takes rows row0 + 32j, j < R
  static __device__ __forceinline__ int item_row0(int item) { return (item >> 5) * (32 * R) + (item & 31); }

  static __device__ __forceinline__ UnitCtx make_ctx(const QuantParams& p, int ps, int t, int T, int nunits) {
    const bool active = ps * T + t < nunits;
    const int u = active ? ps * T + t : 0;
    const int c0 = u << 4;
    const int sg = (c0 >= p.cend[1]) ? 2 : (c0 >= p.cend[0] ? 1 : 0);
    const int fmt = p.fmt[sg];
    const int cb = (sg == 0) ? 0 : p.cend[sg - 1];
    const int G = (c0 - cb) >> 5;  // 32-channel group inside the segment
    const uint32_t p0 = (uint32_t)u * CPT;
    UnitCtx cx;
    cx.meta = (uint32_t)fmt | ((uint32_t)sg << 4) | (active ? 64u : 0u) | ((active && !(u & 1)) ? 128u : 0u);
    cx.qoff = (uint32_t)(((c0 - cb) * fmt) >> 3);
    cx.sfo = (uint32_t)((G >> 2) * 512 + (G & 3));
    cx.xoff = (p0 ^ ((p0 >> 3) & SWM)) << 4;
    return cx;
  }

  template <bool FULL>
  static __device__ __forceinline__ void prefetch(const QuantParams& p, int row0, int t, int T, int K8,
                                                  uint4 (&pre)[NLD][R]) {
    const int K = p.K;
    const uint16_t* b0 = p.x + (int64_t)row0 * K + 8 * t;
    const int64_t rs = (int64_t)32 * K;  // elements between two rows of the item
    const int nvalid = FULL ? R : min(R, ((int)p.rows - 1 - row0) / 32 + 1);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      // rows past the end re-read row0; they are zeroed when stored to shared memory
      const uint16_t* bj = (FULL || j < nvalid) ? b0 + j * rs : b0;
#pragma unroll
      for (int i = 0; i < NLD; ++i)
        if (i * T + t < K8) pre[i][j] = ld_stream_v4(bj + (size_t)i * T * 8);
    }
  }

  template <bool FULL>
  static __device__ __forceinline__ void process(const QuantParams& p, int row0, int t, int T, int K8, int nunits, uint32_t xs_a,
                                                 const uint16_t* tab, uint4 (&pre)[NLD][R], int next_row0, int* s_next,
                                                 const UnitCtx& ctx0) {
    const int nvalid = FULL ? R : min(R, ((int)p.rows - 1 - row0) / 32 + 1);
    // ---- scatter: prefetched registers -> shared memory at the PERMUTED channel position, rows interleaved
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int c8 = i * T + t;
      if (c8 < K8) {
        const uint4 ov = *reinterpret_cast<const uint4*>(tab + 8 * c8);
        const uint32_t o[4] = {ov.x, ov.y, ov.z, ov.w};
        uint32_t w[R][4];
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const bool keep = FULL || j < nvalid;
          w[j][0] = keep ? pre[i][j].x : 0u;
          w[j][1] = keep ? pre[i][j].y : 0u;
          w[j][2] = keep ? pre[i][j].z : 0u;
          w[j][3] = keep ? pre[i][j].w : 0u;
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) {  // channels 8*c8 + 2m (low halves) and + 2m+1 (high halves)
          const uint32_t a_lo = xs_a + (WIDE ? ((o[m] & 0xffffu) << 2) : (o[m] & 0xffffu));
          const uint32_t a_hi = xs_a + (WIDE ? ((o[m] >> 16) << 2) : (o[m] >> 16));
          if constexpr (R == 4) {
            sts64(a_lo, __byte_perm(w[0][m], w[1][m], 0x5410), __byte_perm(w[2][m], w[3][m], 0x5410));
            sts64(a_hi, __byte_perm(w[0][m], w[1][m], 0x7632), __byte_perm(w[2][m], w[3][m], 0x7632));
          } else {
            sts32(a_lo, __byte_perm(w[0][m], w[1][m], 0x5410));
            sts32(a_hi, __byte_perm(w[0][m], w[1][m], 0x7632));
          }
        }
      }
    }
    __syncthreads();
    // the next item's rows: in flight during the compute below
    if (next_row0 >= 0) {
      if (next_row0 + 32 * (R - 1) < (int)p.rows) prefetch<true>(p, next_row0, t, T, K8, pre);
      else prefetch<false>(p, next_row0, t, T, K8, pre);
    }

    // dynamic schedule: claim the item after next now; the answer is only needed one whole item later
    unsigned int claimed = 0;
    if (t == 0) claimed = atomicAdd(p.sched, 1u);

    // ---- compute: thread owns permuted channels [16u, 16u+16) of all R rows.  The loop is warp-uniform (full-mask
    // shuffles inside); lanes past the last unit compute on unit 0 and store nothing.
    const uint32_t sfrow = (uint32_t)(row0 >> 7) * 512u;                                  // x katoms: the row block
    const uint32_t sfin = (uint32_t)(row0 & 31) * 16u + (uint32_t)((row0 >> 5) & 3) * 4u;  // (row0, group 0) in an atom
#pragma unroll
    for (int ps = 0; ps < NP; ++ps) {
      if (ps * T >= nunits) break;
      // one pass: the context lives in registers; more: recomputed per item (it would spill otherwise)
      const UnitCtx cx = (NP == 1) ? ctx0 : make_ctx(p, ps, t, T, nunits);
      const uint32_t meta = cx.meta;
      const bool active = (meta & 64u) != 0;
      const int fmt = (int)(meta & 15u);
      const int sg = (int)((meta >> 4) & 3u);
      const uint32_t base = xs_a + cx.xoff;
      uint32_t g[16][RW];
#pragma unroll
      for (int e = 0; e < CPT; ++e) {
        const uint4 v = lds128(base ^ ((uint32_t)e << 4));
        if constexpr (R == 4) {
          g[2 * e][0] = v.x;
          g[2 * e][1] = v.y;
          g[2 * e + 1][0] = v.z;
          g[2 * e + 1][1] = v.w;
        } else {
          g[4 * e][0] = v.x;
          g[4 * e + 1][0] = v.y;
          g[4 * e + 2][0] = v.z;
          g[4 * e + 3][0] = v.w;
        }
      }

      // ---- absmax per row over the 32-group (16 local channels + the partner lane), scale byte, multiplier
      const uint32_t q2 = (fmt == 4) ? 0x00020002u : ((fmt == 6) ? 0x00040004u : 0x00080008u);  // log2floor(QMAX)
      const uint32_t add2 = (fmt == 4) ? (63u * 0x00010001u) : (31u * 0x00010001u);             // 127 - thr, thr = 64 | 96
      uint32_t mult[RW];
      uint32_t sfb[RW];
#pragma unroll
      for (int k = 0; k < RW; ++k) {
        uint32_t m = absmax2(g[0][k], g[1][k]);
#pragma unroll
        for (int c = 2; c < 16; ++c) m = absmax2(m, g[c][k]);
        m = absmax2(m, __shfl_xor_sync(0xffffffffu, m, 1));
        const uint32_t a = m & 0x7fff7fffu;  // |bf16| bits of the group's absmax, two rows
        // byte = max(exp - qexp, 0) + (mant > thr), both rows of the pair at once; 0x7E for an all-zero group
        const uint32_t ex2 = (a >> 7) & 0x00ff00ffu;
        const uint32_t gt2 = (((a & 0x007f007fu) + add2) >> 7) & 0x00010001u;
        const uint32_t b2 = __vmaxu2(ex2, q2) - q2 + gt2;
        mult[k] = (0x00fe00feu - b2) << 7;  // bf16x2 of 2^(127-byte)
        const uint32_t z = __vcmpeq2(a, 0u);
        sfb[k] = (b2 & ~z) | (0x007e007eu & z);
      }

      // ---- scale bytes: the even lane of the pair writes the R bytes of its group (one per row)
      if (meta & 128u) {
        uint8_t* d = p.sf[sg] + (sfrow * (uint32_t)p.katoms[sg] + sfin + cx.sfo);
#pragma unroll
        for (int k = 0; k < RW; ++k) {
          d[8 * k] = (uint8_t)sfb[k];
          d[8 * k + 4] = (uint8_t)(sfb[k] >> 16);
        }
      }

      // ---- convert + store: 16 codes per row, contiguous bytes
      const int64_t rbytes = p.rowbytes[sg];
      uint8_t* dst = p.q[sg] + ((int64_t)row0 * rbytes + cx.qoff);
      const int64_t rstride = 32 * rbytes;
      const int nstore = active ? nvalid : 0;
      if (FULL && !active) {
      } else if (fmt == 4) convert_store<4, R, FULL>(g, mult, dst, rstride, nstore);
      else if (fmt == 6) convert_store<6, R, FULL>(g, mult, dst, rstride, nstore);
      else convert_store<8, R, FULL>(g, mult, dst, rstride, nstore);
    }
    if (t == 0) *s_next = (int)(claimed + 2u * gridDim.x);
    __syncthreads();  // every read of xs is done before the next item's scatter overwrites it; publishes *s_next
  }
};

template <int R, int TMAX, int NLD, int NP, int MINB, bool WIDE>
__global__ void __launch_bounds__(TMAX, MINB) reorder_quantize_kernel(const __grid_constant__ QuantParams p) {
  static_assert(R == 4 || R == 2, "rows per item");
  using QK = QuantKernel<R, NLD, NP, WIDE>;
  const int T = blockDim.x;  // a multiple of 32 chosen by the launcher so that NP passes of T threads cover K/16 units
  extern __shared__ __align__(128) uint8_t smem[];
  const int K = p.K;
  const int K8 = K >> 3;
  const int nunits = K >> 4;
  uint16_t* tab = reinterpret_cast<uint16_t*>(smem);
  uint8_t* xs = smem + ((K * 2 + 127) & ~127);
  const uint32_t xs_a = smem_addr(xs);
  const int t = threadIdx.x;
  const int rows = (int)p.rows;

  // programmatic dependent launch: let the next kernel in the stream begin its own prologue as SMs free up ...
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  uint4 pre[NLD][R];  // prefetched rows of the next item
  __shared__ int s_next;
  const int num_items = p.num_items;
  // schedule: the first two rounds are static (item = blockIdx.x, blockIdx.x + gridDim.x), later items are claimed
  // from a global counter one item ahead of their prefetch, so SMs that run ahead simply take more items
  int item = blockIdx.x;
  int row0 = item < num_items ? QK::item_row0(item) : rows;

  // ---- one-time per CTA: inverse permutation as swizzled slot offsets, eight entries per step
  for (int j8 = t; j8 < K8; j8 += T) {
    const uint4 iv = __ldg(reinterpret_cast<const uint4*>(p.idx) + j8);
    const uint32_t ivw[4] = {iv.x, iv.y, iv.z, iv.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t j = (uint32_t)j8 * 8u + e;
      const uint32_t c = (e & 1) ? (ivw[e >> 1] >> 16) : (ivw[e >> 1] & 0xffffu);
      const uint32_t pc = (j * QK::SLOT) >> 4;  // 16-byte chunk of slot j
      const uint32_t off = ((pc ^ ((pc >> 3) & QK::SWM)) << 4) | ((j * QK::SLOT) & 15u);
      tab[c] = (uint16_t)(WIDE ? (off >> 2) : off);
    }
  }
  // ---- one-time per thread: what it needs to know about its compute units
  const UnitCtx ctx0 = QK::make_ctx(p, 0, t, T, nunits);
  // ... and wait for the previous kernel (which may still be producing X) only now: the table above depends on
  // reorder_index alone, which no kernel of this library writes, so it was built under the previous kernel's tail.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (row0 < rows) {
    if (row0 + 32 * (R - 1) < rows) QK::template prefetch<true>(p, row0, t, T, K8, pre);
    else QK::template prefetch<false>(p, row0, t, T, K8, pre);
  }
  __syncthreads();

  // items are ordered by row: the first item past the last row ends this CTA's work (block-uniform)
  int nitem = item + (int)gridDim.x;
  while (row0 < rows) {
    int nrow0 = nitem < num_items ? QK::item_row0(nitem) : rows;
    if (nrow0 >= rows) nrow0 = -1;
    if (row0 + 32 * (R - 1) < rows) QK::template process<true>(p, row0, t, T, K8, nunits, xs_a, tab, pre, nrow0, &s_next, ctx0);
    else QK::template process<false>(p, row0, t, T, K8, nunits, xs_a, tab, pre, nrow0, &s_next, ctx0);
    if (nrow0 < 0) break;
    row0 = nrow0;
    nitem = s_next;
  }
  // the last CTA to leave resets the schedule for the next launch that uses this slot
  if (t == 0) {
    const unsigned int done = atomicInc(p.sched + 1, gridDim.x - 1);  // wraps to 0 by itself
    if (done == gridDim.x - 1) atomicExch(p.sched, 0u);
  }
}

__device__ unsigned int g_quant_sched[64][2];  // rotating schedule slots (zero-initialised, self-resetting)

template <int R>
static size_t quant_smem_bytes(int K) {
  return ((size_t)(K * 2 + 127) & ~(size_t)127) + (size_t)K * 2 * R;
}

template <int R, int TMAX, int NLD, int NP, int MINB, bool WIDE>
static int launch_quant(QuantParams& p, cudaStream_t stream) {
  // threads: NP passes of T threads cover the K/16 compute units exactly (T a multiple of 32)
  const int T = ((p.K / 16 + NP - 1) / NP + 31) & ~31;
  const size_t smem = quant_smem_bytes<R>(p.K);
  if (smem > 227 * 1024 || T > TMAX || (int64_t)NLD * T * 8 < p.K || (!WIDE && (int64_t)p.K * 2 * R > 65536)) {
    set_error("reorder_quantize: K=%d does not fit the <%d,%d,%d,%d> kernel (%zu bytes of shared memory)", p.K, R, TMAX,
              NLD, NP, smem);
    return MMX_ERR_INVALID;
  }
  auto kern = reorder_quantize_kernel<R, TMAX, NLD, NP, MINB, WIDE>;
  static size_t attr_set = 0;
  if (smem > attr_set) {
    MMX_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = smem;
  }
  static int occ_cached = 0;
  static size_t occ_smem = 0;
  static int occ_T = 0;
  if (occ_cached == 0 || occ_smem != smem || occ_T != T) {
    occ_T = T;
    int occ = 0;
    MMX_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
    occ_cached = occ < 1 ? 1 : occ;
    occ_smem = smem;
  }
  const int64_t rblocks = (p.rows + 127) / 128;
  const int64_t items = rblocks * 32 * (4 / R);
  p.num_items = (int)items;
  int64_t grid = (int64_t)sm_count() * occ_cached;  // every SM full; items beyond two rounds are claimed dynamically
  if (options().quant_ctas > 0) grid = options().quant_ctas;
  if (grid > items) grid = items;
  if (grid < 1) return MMX_OK;
  static unsigned int* sched_base = nullptr;
  if (!sched_base) MMX_CUDA_TRY(cudaGetSymbolAddress(reinterpret_cast<void**>(&sched_base), g_quant_sched));
  static std::atomic<unsigned int> seq{0};
  p.sched = sched_base + 2 * (seq.fetch_add(1, std::memory_order_relaxed) & 63u);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = options().pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MMX_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
  g_launches.fetch_add(1, std::memory_order_relaxed);
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
  if (rows > (int64_t)1 << 24) {
    set_error("reorder_quantize: rows=%lld exceeds the supported 2^24", (long long)rows);
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
  int cacc = 0;
  for (int i = 0; i < 3; ++i) {
    p.fmt[i] = fmt[i];
    cacc += ks[i];
    p.cend[i] = cacc;
    p.katoms[i] = ks[i] / 128;
    p.rowbytes[i] = (int64_t)ks[i] * fmt[i] / 8;
    p.q[i] = q[i];
    p.sf[i] = s[i];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int force = (int)options().quant_rows;
  // configuration by K: rows per item R, thread bound, 16-byte chunks per thread and row NLD, compute passes NP
  if (force == 4 && K <= 4096) return launch_quant<4, 256, 2, 1, 3, false>(p, st);
  if (K <= 4096) return launch_quant<2, 256, 2, 1, 4, false>(p, st);
  if (K <= 8192) return launch_quant<2, 512, 2, 1, 2, false>(p, st);
  if (K <= 14336) return launch_quant<2, 448, 4, 2, 2, false>(p, st);  // 72 registers at two CTAs per SM
  if (K <= 16384) return launch_quant<2, 512, 4, 2, 2, false>(p, st);
  return launch_quant<2, 1024, 4, 2, 1, true>(p, st);
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
