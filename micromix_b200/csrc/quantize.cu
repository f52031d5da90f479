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

#include <cstring>
#include <mutex>

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
  const uint16_t* nw;  // RMSNorm weight (bf16 [K], ORIGINAL channel order); NORM kernels only
  float eps;
  // MC kernels (sequence-parallel hand-over, tp_reduce.cu): q[] / sf[] are NVSwitch MULTICAST addresses -- every store
  // lands in the same place of every rank's gather buffer.  The kernel first waits until all ranks are done reading the
  // previous gather (ag_consumed >= ag_tp * ag_issued) and, after its last store, bumps ag_arrived[] on every rank.
  const uint32_t* ag_consumed;
  uint32_t* ag_issued;
  uint32_t* ag_arrived[kMaxTp];
  int ag_tp;
  // ALL-TO-ALL form of the MC kernel (token-parallel row linears): rows [d * a2a_per, (d+1) * a2a_per) go to rank d's
  // buffer (unicast peer pointers qd / sfd, already offset to THIS rank's columns / scale atoms inside the destination's
  // [a2a_per, K_total] activation), whose packed rows are a2a_pitch bytes long and whose scale row blocks hold the atoms
  // of all ranks (katoms[] then counts those).  a2a_per == 0: the gather form (q[] / sf[] are multicast addresses).
  int a2a_per;
  uint8_t* qd[kMaxTp][3];
  uint8_t* sfd[kMaxTp][3];
  uint32_t a2a_pitch[3];
  unsigned long long ag_timeout_ns;  // bound on the wait (option tp_timeout_ms): a lost peer must not hang the GPU
  uint32_t* ag_err;                  // local error word of the tp context: bit 2 = this wait timed out
  uint32_t ag_dbg;                   // timing experiments (option tp_debug >> 5)
  // GROUPED (Mixtral experts): rows are sorted by group, every 128-row block belongs to one group, and group g permutes
  // its rows with idx + g * K (same split for all groups).  grp_rowblk[row / 128] = group (device memory, written on the
  // stream by the router); a CTA rebuilds its scatter table when the group of its next item changes.
  const int* grp_rowblk;
  // ... and row r of the sorted matrix is row row_src[r] of x (the token of that (token, slot) pair; padding rows carry
  // any valid token): the gather of the routed tokens happens in this kernel's loads, not in a staging copy
  const int* row_src;
  const int* rows_dev;  // optional, device memory: only the first *rows_dev rows exist (the padded row count of the sorted
                        // matrix is known on the device only); p.rows is the static upper bound the buffers were sized for
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
// MC = the destination is a multicast address: multimem.st (one store, replicated to every rank by the switch)
template <bool MC = false>
__device__ __forceinline__ void stg_v4(void* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  if constexpr (MC)
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(a)),
                 "f"(__uint_as_float(b)), "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
                 : "memory");
  else
    asm volatile("st.global.L1::no_allocate.v4.b32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
template <bool MC = false>
__device__ __forceinline__ void stg_v2(void* p, uint32_t a, uint32_t b) {
  if constexpr (MC)
    asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(__uint_as_float(a)),
                 "f"(__uint_as_float(b))
                 : "memory");
  else
    asm volatile("st.global.L1::no_allocate.v2.b32 [%0], {%1, %2};" ::"l"(p), "r"(a), "r"(b) : "memory");
}
template <bool MC = false>
__device__ __forceinline__ void stg_b32(void* p, uint32_t a) {
  if constexpr (MC)
    asm volatile("multimem.st.relaxed.sys.global.b32 [%0], %1;" ::"l"(p), "r"(a) : "memory");
  else
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
// MC (multicast destination): the codes are STAGED in shared memory -- `dst` is then a shared-memory address carried in a
// pointer -- and leave later as whole 16-byte lines per lane (QuantKernel::copy_out): an NVLink multicast write of 4 or
// 8 bytes costs a packet of its own, a warp's 512 contiguous bytes cost four.
template <int FMT, bool MC>
__device__ __forceinline__ void convert_store_row(const float (&f)[16], uint8_t* dst) {
  if constexpr (MC) {
    const uint32_t a = (uint32_t)(uintptr_t)dst;
    if constexpr (FMT == 4) {
      sts64(a, cvt8_e2m1(&f[0]), cvt8_e2m1(&f[8]));
    } else if constexpr (FMT == 6) {
      const uint32_t y0 = squeeze4_fp6(cvt4_e3m2(&f[0])), y1 = squeeze4_fp6(cvt4_e3m2(&f[4]));
      const uint32_t y2 = squeeze4_fp6(cvt4_e3m2(&f[8])), y3 = squeeze4_fp6(cvt4_e3m2(&f[12]));
      sts32(a, __byte_perm(y0, y1, 0x4210));
      sts32(a + 4, __byte_perm(y1, y2, 0x5421));
      sts32(a + 8, __byte_perm(y2, y3, 0x6542));
    } else {
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(cvt4_e4m3(&f[0])), "r"(cvt4_e4m3(&f[4])),
                   "r"(cvt4_e4m3(&f[8])), "r"(cvt4_e4m3(&f[12]))
                   : "memory");
    }
  } else if constexpr (FMT == 4) {
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
template <int FMT, int R, bool FULL, bool MC>
__device__ __forceinline__ void convert_store(const uint32_t (&g)[16][R / 2], const uint32_t (&mult)[R / 2], uint8_t* dst,
                                              int64_t row_stride, int nstore) {
#pragma unroll
  for (int k = 0; k < R / 2; ++k) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      // x * 2^-e straight into fp32: ONE mixed-precision FMA per element (FHFMA.BF16 with a half selector on both packed
      // operands) replaces the bf16x2 multiply + the unpack of its halves.  The product of a bf16 and a power of two is
      // exact in fp32; the addend is -0.0 so that a -0.0 input keeps its sign bit (reorder.cu gives -0 a sign-bit code).
      float f[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) {
        if (half == 0)
          asm("{.reg .b16 xl, xh, ml, mh;\n"
              "mov.b32 {xl, xh}, %1;\n"
              "mov.b32 {ml, mh}, %2;\n"
              "fma.rn.f32.bf16 %0, xl, ml, 0f80000000;}"
              : "=f"(f[c])
              : "r"(g[c][k]), "r"(mult[k]));
        else
          asm("{.reg .b16 xl, xh, ml, mh;\n"
              "mov.b32 {xl, xh}, %1;\n"
              "mov.b32 {ml, mh}, %2;\n"
              "fma.rn.f32.bf16 %0, xh, mh, 0f80000000;}"
              : "=f"(f[c])
              : "r"(g[c][k]), "r"(mult[k]));
      }
      if (FULL || 2 * k + half < nstore) convert_store_row<FMT, MC>(f, dst);
      dst += row_stride;
    }
  }
}

// What a thread needs to know about one of its compute units (16 permuted channels of all R rows).  Item-invariant:
// computed once per CTA; kept in registers when a thread has one unit (NP == 1), in shared memory otherwise.
struct UnitCtx {
  uint32_t meta;  // bits 0..3 fmt (4|6|8), bits 4..5 segment, bit 6 active, bit 7 writes the group's scale bytes
  uint32_t qoff;  // byte offset of the unit's codes inside a packed row of its segment
  uint32_t sfo;   // byte offset of the unit's group inside the row block's scale atoms: (G/4)*512 + G%4
  uint32_t xoff;  // byte offset of the unit's first (swizzled) 16-byte chunk inside xs
};

// Everything else a unit needs, derived from its context and the launch parameters.  For NP == 1 this lives in
// registers for the whole kernel, so the per-item address work is one multiply-add per pointer.
struct UnitPtrs {
  uint8_t* qp;      // p.q[sg] + qoff
  uint8_t* sfp;     // p.sf[sg] + sfo
  uint32_t rbytes;  // bytes of one packed row of the segment
  uint32_t ka512;   // scale bytes of one 128-row block of the segment: katoms * 512
  uint32_t q2;      // log2floor(QMAX) in both 16-bit lanes: 2 | 4 | 8
  uint32_t add2;    // 127 - mantissa threshold (0x40 for QMAX = 6 = 1.5 * 4, 0x60 for 28 | 448 = 1.75 * 2^n) in both lanes
  uint32_t stoff;   // MC: byte offset of the unit's codes inside a STAGED row (the three segments' packed rows back to back)
};

// R rows per item, NLD = 16-byte chunks per thread and row (NLD * T * 8 >= K), NP = compute passes
// (NP * T * 16 >= K), TABMODE = encoding of the scatter table: 0 = u16 byte offset into xs, 1 = u16 offset / 4
// (K * 2R > 65536), 2 = u32 absolute shared-memory address (no unpacking in the scatter loop).
//
// Shared memory: tab[K] | xs[K] slots of 2R bytes | NP > 1: ctx[K/16]
//   tab[c]  where the slot of ORIGINAL channel c lives inside xs, i.e. of permuted position j with idx[j] == c,
//           after the chunk swizzle.  Built once per CTA.
//   xs      slot j holds the R rows of permuted channel j; 16-byte chunk p = j * 2R / 16 is stored at chunk
//           p ^ ((p >> 3) & SWM) so that the lane-strided 128-bit reads of the compute phase are conflict-free.
//
// NORM (mmx_rmsnorm_quantize_x): the rows are RMS-normalised on the way through.  The sum of squares of a row is taken
// in a FIXED order that does not depend on the launch shape (oracle/mmx_oracle.c restates it): an fp32 fma chain over
// each aligned 8-channel chunk (exactly the 16 bytes one thread loads), then a perfect binary tree over the chunk
// index -- five shuffle levels inside a warp, the warp sums through shared memory, seven more levels redone by every
// warp.  y = bf16((x * w) * rinv) is applied in the compute phase with the weight staged in PERMUTED order.
//
// DEPTH > 0 (EXPERIMENT, option quant_variant = 2 | 3; not the default): the rows of the next DEPTH items are in flight at
// any time, copied by cp.async into a per-thread ring of raw 16-byte chunks in shared memory (a thread reads back only
// what it copied itself: no barrier, no bank conflicts) instead of ONE item held in registers.  Measured SLOWER than the
// register prefetch at every shape (profiles/r02_quantize_depth.log: 8192 x 4096 22.9 vs 21.0 us, 65536 x 4096 144.9 vs
// 128.3 us): the kernel is not short of bytes in flight -- four CTAs per SM with one 16 KB item each already cover the
// DRAM latency; the extra shared-memory round trip costs more than the deeper queue gains.
template <int R, int NLD, int NP, int TABMODE, bool NORM = false, bool MC = false, int DEPTH = 0>
struct QuantKernel {
  static constexpr int RW = R / 2;          // 32-bit words per slot (two rows per word)
  static constexpr int SLOT = 2 * R;        // bytes per slot
  static constexpr int CPT = SLOT;          // 16-byte chunks holding a thread's 16 channels: 8 (R=4) | 4 (R=2)
  static constexpr uint32_t SWM = (R == 4) ? 7u : 3u;
  static constexpr int TABW = (TABMODE == 2) ? 4 : 2;

  // item -> first row: 32 consecutive items share a block of 32R rows, item i takes rows row0 + 32j, j < R
  static __device__ __forceinline__ int item_row0(int item, int num_items, int rows) {
    return item < num_items ? (item >> 5) * (32 * R) + (item & 31) : rows;
  }

  static __device__ __forceinline__ UnitCtx make_ctx(const QuantParams& p, int u, int nunits) {
    const bool active = u < nunits;
    if (!active) u = 0;
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

  static __device__ __forceinline__ UnitPtrs make_ptrs(const QuantParams& p, const UnitCtx& cx) {
    const int sg = (int)((cx.meta >> 4) & 3u);
    const int fmt = (int)(cx.meta & 15u);
    UnitPtrs up;
    up.qp = p.q[sg] + cx.qoff;
    up.sfp = p.sf[sg] + cx.sfo;
    up.rbytes = (uint32_t)p.rowbytes[sg];
    up.ka512 = (uint32_t)p.katoms[sg] * 512u;
    up.q2 = (fmt == 4) ? 0x00020002u : ((fmt == 6) ? 0x00040004u : 0x00080008u);
    up.add2 = (fmt == 4) ? 0x003f003fu : 0x001f001fu;
    up.stoff = cx.qoff + (sg == 0 ? 0u : (uint32_t)p.rowbytes[0]) + (sg == 2 ? (uint32_t)p.rowbytes[1] : 0u);
    return up;
  }

  // The global loads of one item: thread t takes the 16-byte chunks t, t + T, ... of each of the R rows.
  // EXACT: NLD * T * 8 == K, no chunk predicate.  FULL: all R rows exist.
  template <bool FULL, bool EXACT>
  static __device__ __forceinline__ void prefetch(const QuantParams& p, const uint16_t* xt, int row0, int t, int T, int K8,
                                                  uint4 (&pre)[NLD][R], uint32_t raw_a = 0) {
    const int K = p.K;
    const uint16_t* b0 = xt + (int64_t)row0 * K;
    const int64_t rs = (int64_t)32 * K;  // elements between two rows of the item
    const int nvalid = FULL ? R : min(R, ((int)p.rows - 1 - row0) / 32 + 1);
#pragma unroll
    for (int j = 0; j < R; ++j) {
      // rows past the end re-read row0; they are zeroed when stored to shared memory
      const uint16_t* bj = (FULL || j < nvalid) ? b0 + j * rs : b0;
      if (p.row_src != nullptr)  // grouped: the row is gathered through the routing table (uniform branch)
        bj = xt + (int64_t)__ldg(p.row_src + ((FULL || j < nvalid) ? row0 + 32 * j : row0)) * K;
#pragma unroll
      for (int i = 0; i < NLD; ++i)
        if (EXACT || i * T + t < K8) {
          if constexpr (DEPTH > 0)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(raw_a + 16u * (uint32_t)((j * NLD + i) * T + t)),
                         "l"(bj + (size_t)i * T * 8)
                         : "memory");
          else
            pre[i][j] = ld_stream_v4(bj + (size_t)i * T * 8);
        }
    }
  }

  // DEPTH > 0: this thread's chunks of the oldest item in the ring -> registers
  template <bool EXACT>
  static __device__ __forceinline__ void load_raw(uint32_t raw_a, int t, int T, int K8, uint4 (&pre)[NLD][R]) {
#pragma unroll
    for (int j = 0; j < R; ++j)
#pragma unroll
      for (int i = 0; i < NLD; ++i)
        if (EXACT || i * T + t < K8) pre[i][j] = lds128(raw_a + 16u * (uint32_t)((j * NLD + i) * T + t));
  }

  // ---- scatter: prefetched registers -> shared memory at the PERMUTED channel position, rows interleaved
  template <bool FULL, bool EXACT>
  static __device__ __forceinline__ void scatter(int nvalid, int t, int T, int K8, uint32_t xs_a, uint32_t tab_a,
                                                 const uint4 (&pre)[NLD][R], uint32_t ss_a = 0) {
    if constexpr (NORM) {
      // sum of squares: chunk c8 = i*T + t is this thread's, the warp owns 32 consecutive chunks (T % 32 == 0)
#pragma unroll
      for (int i = 0; i < NLD; ++i) {
        const int c8 = i * T + t;
        float ssq[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
          float acc = 0.0f;
          if ((EXACT || c8 < K8) && (FULL || j < nvalid)) {
            const uint32_t wv[4] = {pre[i][j].x, pre[i][j].y, pre[i][j].z, pre[i][j].w};
#pragma unroll
            for (int m = 0; m < 4; ++m) {
              // acc = fma(lo, lo, acc); acc = fma(hi, hi, acc) with the bf16 halves as operands of the mixed-precision FMA
              // (no unpack instructions; a bf16 square is exact in fp32, so the rounding is the plain fp32 FMA's)
              asm("{.reg .b16 l, h;\n"
                  "mov.b32 {l, h}, %1;\n"
                  "fma.rn.f32.bf16 %0, l, l, %0;\n"
                  "fma.rn.f32.bf16 %0, h, h, %0;}"
                  : "+f"(acc)
                  : "r"(wv[m]));
            }
          }
          ssq[j] = acc;
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
#pragma unroll
          for (int j = 0; j < R; ++j) ssq[j] = __fadd_rn(ssq[j], __shfl_xor_sync(0xffffffffu, ssq[j], d));
        }
        if ((t & 31) == 0) {
#pragma unroll
          for (int j = 0; j < R; ++j)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(ss_a + 4u * (uint32_t)(j * 128 + ((i * T + t) >> 5))), "f"(ssq[j]) : "memory");
        }
      }
    }
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int c8 = i * T + t;
      if (EXACT || c8 < K8) {
        uint32_t adr[8];  // shared-memory address of the slot of original channel 8*c8 + e
        if constexpr (TABMODE == 2) {
          const uint4 o0 = lds128(tab_a + 32u * (uint32_t)c8), o1 = lds128(tab_a + 32u * (uint32_t)c8 + 16u);
          adr[0] = o0.x; adr[1] = o0.y; adr[2] = o0.z; adr[3] = o0.w;
          adr[4] = o1.x; adr[5] = o1.y; adr[6] = o1.z; adr[7] = o1.w;
        } else {
          const uint4 ov = lds128(tab_a + 16u * (uint32_t)c8);
          const uint32_t o[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
          for (int m = 0; m < 4; ++m) {
            adr[2 * m] = xs_a + ((TABMODE == 1) ? ((o[m] & 0xffffu) << 2) : (o[m] & 0xffffu));
            adr[2 * m + 1] = xs_a + ((TABMODE == 1) ? ((o[m] >> 16) << 2) : (o[m] >> 16));
          }
        }
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
          if constexpr (R == 4) {
            sts64(adr[2 * m], __byte_perm(w[0][m], w[1][m], 0x5410), __byte_perm(w[2][m], w[3][m], 0x5410));
            sts64(adr[2 * m + 1], __byte_perm(w[0][m], w[1][m], 0x7632), __byte_perm(w[2][m], w[3][m], 0x7632));
          } else {
            sts32(adr[2 * m], __byte_perm(w[0][m], w[1][m], 0x5410));
            sts32(adr[2 * m + 1], __byte_perm(w[0][m], w[1][m], 0x7632));
          }
        }
      }
    }
  }

  // ---- compute: one unit = permuted channels [16u, 16u+16) of all R rows.  Warp-uniform (full-mask shuffle inside);
  // lanes past the last unit compute on unit 0 and store nothing.
  template <bool FULL>
  static __device__ __forceinline__ void compute_unit(const UnitCtx& cx, const UnitPtrs& up, uint32_t xs_a, int row0,
                                                      int nvalid, uint32_t wp_a = 0, const float* rinv = nullptr,
                                                      uint32_t stage_a = 0, uint32_t stage_row = 0, int a2a_per = 0,
                                                      uint8_t* const* a2a_sf = nullptr) {
    const uint32_t meta = cx.meta;
    const bool active = (meta & 64u) != 0;
    const int fmt = (int)(meta & 15u);
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

    if constexpr (NORM) {
      // y = bf16((x * w) * rinv): wp_a points at the 16 permuted weights of this unit (bf16, 32 bytes)
      const uint4 wa = lds128(wp_a), wb = lds128(wp_a + 16u);
      const uint32_t ww[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
      for (int c = 0; c < 16; ++c) {
#pragma unroll
        for (int k = 0; k < RW; ++k) {
          // x * w: the product of two bf16 values is exact in fp32, so one mixed-precision FMA on the packed halves (addend
          // -0.0: the sign of a zero product is kept) equals fmul(float(x), float(w)) without the three unpack instructions
          float t0, t1;
          if (c & 1)
            asm("{.reg .b16 xl, xh, wl, wh;\n"
                "mov.b32 {xl, xh}, %2;\n"
                "mov.b32 {wl, wh}, %3;\n"
                "fma.rn.f32.bf16 %0, xl, wh, 0f80000000;\n"
                "fma.rn.f32.bf16 %1, xh, wh, 0f80000000;}"
                : "=f"(t0), "=f"(t1)
                : "r"(g[c][k]), "r"(ww[c >> 1]));
          else
            asm("{.reg .b16 xl, xh, wl, wh;\n"
                "mov.b32 {xl, xh}, %2;\n"
                "mov.b32 {wl, wh}, %3;\n"
                "fma.rn.f32.bf16 %0, xl, wl, 0f80000000;\n"
                "fma.rn.f32.bf16 %1, xh, wl, 0f80000000;}"
                : "=f"(t0), "=f"(t1)
                : "r"(g[c][k]), "r"(ww[c >> 1]));
          const __nv_bfloat162 y = __floats2bfloat162_rn(__fmul_rn(t0, rinv[2 * k]), __fmul_rn(t1, rinv[2 * k + 1]));
          g[c][k] = *reinterpret_cast<const uint32_t*>(&y);
        }
      }
    }

    // ---- absmax per row over the 32-group (16 local channels + the partner lane), scale byte, multiplier
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
      const uint32_t gt2 = (((a & 0x007f007fu) + up.add2) >> 7) & 0x00010001u;
      const uint32_t b2 = __vmaxu2(ex2, up.q2) - up.q2 + gt2;
      mult[k] = (0x00fe00feu - b2) << 7;  // bf16x2 of 2^(127-byte)
      // an all-zero group has b2 == 0 (exp 0 < qexp, mant 0): add 0x7E exactly there
      const uint32_t nz = __vminu2(a, 0x00010001u);  // 1 where the row's absmax is non-zero
      sfb[k] = b2 + 0x007e007eu - nz * 0x7eu;
    }

    // ---- scale bytes: the even lane of the pair writes the R bytes of its group (one per row, 4 bytes apart)
    if constexpr (MC) {
      // multicast stores should be as wide as possible (every one is an NVLink packet): lanes 8q .. 8q+7 hold the four
      // groups of ONE scale atom (128 channels, same segment: segments are multiples of 128); the item's two rows
      // l + 64h and l + 64h + 32 own bytes [8h, 8h + 8) of the atom's 16-byte line for row l -- lane 8q collects the four
      // groups from lanes 8q+2, +4, +6 and writes those 8 bytes with one store
      static_assert(RW == 1, "two rows per item");
      uint32_t sfoff = (uint32_t)(row0 >> 7) * up.ka512 + (uint32_t)(row0 & 31) * 16u + (uint32_t)((row0 >> 5) & 3) * 4u;
      uint8_t* d = up.sfp + sfoff;
      if (a2a_per > 0) {  // all-to-all: the row block lives in its owner's buffer (a2a_per is a multiple of 256)
        const int dst = row0 / a2a_per, rl = row0 - dst * a2a_per;
        sfoff = (uint32_t)(rl >> 7) * up.ka512 + (uint32_t)(row0 & 31) * 16u + (uint32_t)((row0 >> 5) & 3) * 4u;
        d = a2a_sf[dst * 3 + ((meta >> 4) & 3u)] + cx.sfo + sfoff;
      }
      const uint32_t v1 = __shfl_down_sync(0xffffffffu, sfb[0], 2), v2 = __shfl_down_sync(0xffffffffu, sfb[0], 4);
      const uint32_t v3 = __shfl_down_sync(0xffffffffu, sfb[0], 6);
      const uint32_t lo01 = __byte_perm(sfb[0], v1, 0x0040), lo23 = __byte_perm(v2, v3, 0x0040);  // first row : bytes 0
      const uint32_t hi01 = __byte_perm(sfb[0], v1, 0x0062), hi23 = __byte_perm(v2, v3, 0x0062);  // second row: bytes 2
      if ((meta & 64u) && (threadIdx.x & 7) == 0) {
        const uint32_t w0 = __byte_perm(lo01, lo23, 0x5410), w1 = __byte_perm(hi01, hi23, 0x5410);
        if (FULL || nvalid > 1) stg_v2<true>(d, w0, w1);
        else stg_b32<true>(d, w0);
      }
    } else if (meta & 128u) {
      const uint32_t sfoff = (uint32_t)(row0 >> 7) * up.ka512 + (uint32_t)(row0 & 31) * 16u + (uint32_t)((row0 >> 5) & 3) * 4u;
      uint8_t* d = up.sfp + sfoff;
#pragma unroll
      for (int k = 0; k < RW; ++k) {
        d[8 * k] = (uint8_t)sfb[k];
        d[8 * k + 4] = (uint8_t)(sfb[k] >> 16);
      }
    }

    // ---- convert + store: 16 codes per row, contiguous bytes
    uint8_t* dst = up.qp + (uint64_t)(uint32_t)row0 * up.rbytes;
    int64_t rstride = (int64_t)(32u * up.rbytes);
    if constexpr (MC) {  // staged: row j of the item at stage_a + j * stage_row
      dst = reinterpret_cast<uint8_t*>((uintptr_t)(stage_a + up.stoff));
      rstride = (int64_t)stage_row;
    }
    const int nstore = active ? nvalid : 0;
    if (FULL && !active) {
    } else if (fmt == 4) convert_store<4, R, FULL, MC>(g, mult, dst, rstride, nstore);
    else if (fmt == 6) convert_store<6, R, FULL, MC>(g, mult, dst, rstride, nstore);
    else convert_store<8, R, FULL, MC>(g, mult, dst, rstride, nstore);
  }

  // MC: the staged codes of one item -> the multicast address, 16 bytes per lane, a warp writes 512 contiguous bytes
  static __device__ __forceinline__ void copy_out(const QuantParams& p, uint32_t stage_a, uint32_t stage_row, int row0,
                                                  int nvalid, int t, int T) {
    const uint32_t rb0 = (uint32_t)p.rowbytes[0], rb1 = (uint32_t)p.rowbytes[1], rb2 = (uint32_t)p.rowbytes[2];
    const uint32_t n16 = stage_row >> 4;
    for (int j = 0; j < nvalid; ++j) {
      const int64_t row = row0 + 32 * j;
      for (uint32_t c = (uint32_t)t; c < n16; c += (uint32_t)T) {
        const uint32_t o = c << 4;
        const uint4 v = lds128(stage_a + (uint32_t)j * stage_row + o);
        uint8_t* g;
        if (p.a2a_per > 0) {  // all-to-all: a peer-mapped unicast address in the row's owner
          const int dst = (int)(row / p.a2a_per);
          const int64_t rl = row - (int64_t)dst * p.a2a_per;
          if (o < rb0) g = p.qd[dst][0] + rl * p.a2a_pitch[0] + o;
          else if (o < rb0 + rb1) g = p.qd[dst][1] + rl * p.a2a_pitch[1] + (o - rb0);
          else g = p.qd[dst][2] + rl * p.a2a_pitch[2] + (o - rb0 - rb1);
        } else if (o < rb0) g = p.q[0] + row * rb0 + o;
        else if (o < rb0 + rb1) g = p.q[1] + row * rb1 + (o - rb0);
        else g = p.q[2] + row * rb2 + (o - rb0 - rb1);
        stg_v4<true>(g, v.x, v.y, v.z, v.w);
      }
    }
  }
};

// One item, start to finish.  `pre` holds the item's rows on entry and the next item's rows on exit; returns the
// next item's first row.  Schedule protocol: s_next[(n+1)&1] holds the first row of item n+1 (written during item
// n-1, or before the loop for n == 0); thread 0 claims item n+2 and publishes it in s_next[n&1].
// NBUF == 2: xs is double-buffered and ONE barrier per item suffices (the buffer written by item n+1 was last read by
// item n-1, which every thread has left before it passes item n's barrier).
// NORM: ss_a = this item's warp sums of squares [R][128] fp32 (double-buffered with xs), wp_unit = this thread's 16
// permuted RMSNorm weights.
// MC: stg_a = the two staging buffers of the packed codes; the PREVIOUS item's codes leave right after this item's barrier
// (every thread has finished the previous compute by then, and the buffer is not written again before the next barrier).
// DEPTH > 0: s_next is a ring of DEPTH + 1 entries (item n + k at slot (n + k) % (DEPTH + 1)); the rows of items n ..
// n + DEPTH - 1 are in flight on entry (cp.async groups, oldest first), item n + DEPTH is issued here.
template <typename QK, int R, int NLD, int NP, int NBUF, bool FULL, bool EXACT, bool NORM, bool MC, int DEPTH>
__device__ __forceinline__ int quant_process(const QuantParams& p, const uint16_t* xt, int row0, int rows, int n, int t, int T,
                                             int K8, int nunits, uint32_t xs_a, uint32_t tab_a, uint32_t ctx_a,
                                             uint4 (&pre)[NLD][R], int* s_next, const UnitCtx& ctx0, const UnitPtrs& up0,
                                             uint32_t ss_a, uint32_t wp_unit, uint32_t stg_a, uint32_t stage_row,
                                             int& prev_row0, int& prev_nvalid, uint32_t raw_a, uint32_t raw_bytes) {
  constexpr int DE = DEPTH > 0 ? DEPTH : 1, RING = DE + 1;
  const int nvalid = FULL ? R : min(R, (rows - 1 - row0) / 32 + 1);
  const uint32_t raw_cur = raw_a + (uint32_t)(n % DE) * raw_bytes;
  if constexpr (DEPTH > 0) {
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    QK::template load_raw<EXACT>(raw_cur, t, T, K8, pre);
  }
  QK::template scatter<FULL, EXACT>(nvalid, t, T, K8, xs_a, tab_a, pre, ss_a);
  __syncthreads();
  const uint32_t stage_cur = stg_a + (uint32_t)(n & 1) * (uint32_t)R * stage_row;
  if constexpr (MC) {
    if (prev_row0 >= 0)
      QK::copy_out(p, stg_a + (uint32_t)((n + 1) & 1) * (uint32_t)R * stage_row, stage_row, prev_row0, prev_nvalid, t, T);
    prev_row0 = row0;
    prev_nvalid = nvalid;
  }
  float rinv[R];
  if constexpr (NORM) {
    // the top seven levels of the tree over the chunk index: 128 warp sums per row (unused ones are zero), four
    // consecutive ones per lane, then five shuffle levels.  Every warp redoes it: cheaper than a second barrier.
#pragma unroll
    for (int j = 0; j < R; ++j) {
      const uint4 v = lds128(ss_a + (uint32_t)(j * 512 + (t & 31) * 16));
      float sum = __fadd_rn(__fadd_rn(__uint_as_float(v.x), __uint_as_float(v.y)),
                            __fadd_rn(__uint_as_float(v.z), __uint_as_float(v.w)));
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) sum = __fadd_rn(sum, __shfl_xor_sync(0xffffffffu, sum, d));
      rinv[j] = __frcp_rn(__fsqrt_rn(__fadd_rn(__fdiv_rn(sum, (float)p.K), p.eps)));
    }
  }
  // the rows of item n + DE: in flight during the compute below (and, DEPTH > 1, during the next DEPTH - 1 items)
  const int next_row0 = s_next[(n + 1) % RING];
  const int pf_row0 = (DE == 1) ? next_row0 : s_next[(n + DE) % RING];
  if (pf_row0 < rows) {
    if (pf_row0 + 32 * (R - 1) < rows) QK::template prefetch<true, EXACT>(p, xt, pf_row0, t, T, K8, pre, raw_cur);
    else QK::template prefetch<false, EXACT>(p, xt, pf_row0, t, T, K8, pre, raw_cur);
  }
  if constexpr (DEPTH > 0) asm volatile("cp.async.commit_group;" ::: "memory");
  // dynamic schedule: thread 0 claims item n + DE + 1 now; the answer is needed one whole item later
  unsigned int claimed = 0;
  if (t == 0) claimed = atomicAdd(p.sched, 1u);

  if constexpr (NP == 1) {
    QK::template compute_unit<FULL>(ctx0, up0, xs_a, row0, nvalid, wp_unit, rinv, stage_cur, stage_row, MC ? p.a2a_per : 0,
                                    MC ? &p.sfd[0][0] : nullptr);
  } else {
#pragma unroll
    for (int ps = 0; ps < NP; ++ps) {
      if (ps * T >= nunits) break;
      const uint4 cv = lds128(ctx_a + 16u * (uint32_t)(ps * T + t));
      UnitCtx cx;
      cx.meta = cv.x;
      cx.qoff = cv.y;
      cx.sfo = cv.z;
      cx.xoff = cv.w;
      const UnitPtrs up = QK::make_ptrs(p, cx);
      QK::template compute_unit<FULL>(cx, up, xs_a, row0, nvalid);
    }
  }
  if (t == 0) s_next[n % RING] = QK::item_row0((int)(claimed + (unsigned)RING * gridDim.x), p.num_items, (int)p.rows);
  if constexpr (NBUF == 1) __syncthreads();  // every read of xs is done before the next item's scatter overwrites it
  return next_row0;
}

template <int R, int TMAX, int NLD, int NP, int MINB, int TABMODE, int NBUF, bool EXACT, bool NORM, bool MC = false, int DEPTH = 0>
__global__ void __launch_bounds__(TMAX, MINB) reorder_quantize_kernel(const __grid_constant__ QuantParams p) {
  static_assert(R == 4 || R == 2, "rows per item");
  static_assert(NBUF == 1 || (NBUF == 2 && TABMODE != 2), "absolute table addresses cannot follow a second xs buffer");
  static_assert(!NORM || NP == 1, "the fused RMSNorm keeps a thread's weights addressable by thread index");
  static_assert(!MC || (R == 2 && NP == 1), "the multicast path is written for two rows per item, one unit per thread");
  using QK = QuantKernel<R, NLD, NP, TABMODE, NORM, MC, DEPTH>;
  constexpr int DE = DEPTH > 0 ? DEPTH : 1, RING = DE + 1;
  const int T = blockDim.x;  // a multiple of 32 chosen by the launcher so that NP passes of T threads cover K/16 units
  extern __shared__ __align__(128) uint8_t smem[];
  const int K = p.K;
  const int K8 = K >> 3;
  const int nunits = K >> 4;
  const uint32_t tab_a = smem_addr(smem);
  const uint32_t xs_a = tab_a + (uint32_t)((K * QK::TABW + 127) & ~127);
  const uint32_t xs_bytes = (uint32_t)K * QK::SLOT;
  const uint32_t ctx_a = xs_a + NBUF * xs_bytes;  // NP > 1 only
  const uint32_t wp_a = ctx_a;                    // NORM only (NP == 1): bf16 weights in permuted order [K]
  const uint32_t ss_a = wp_a + (uint32_t)K * 2u;  // NORM only: warp sums of squares [NBUF][R][128] fp32
  // MC only: two staging buffers of R packed rows (the three segments back to back), behind everything else
  const uint32_t stage_row = (uint32_t)(p.rowbytes[0] + p.rowbytes[1] + p.rowbytes[2]);
  const uint32_t stg_a = ctx_a + (NP > 1 ? (uint32_t)(NP * T * 16) : 0u) +
                         (NORM ? (uint32_t)K * 2u + (uint32_t)(NBUF * R * 512) : 0u);
  // DEPTH > 0 only: the ring of raw rows, R * NLD * T chunks of 16 bytes per item, behind everything else
  const uint32_t raw_a = stg_a + (MC ? 2u * (uint32_t)R * stage_row : 0u);
  const uint32_t raw_bytes = (uint32_t)(R * NLD * 16) * (uint32_t)T;
  int prev_row0 = -1, prev_nvalid = 0;
  const int t = threadIdx.x;
  int rows = (int)p.rows;

  // programmatic dependent launch: let the next kernel in the stream begin its own prologue as SMs free up ...
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  uint4 pre[NLD][R];  // prefetched rows of the next item
  __shared__ int s_next[RING];
  const int num_items = p.num_items;
  // schedule: the first RING rounds are static (item = blockIdx.x + k * gridDim.x), later items are claimed from a global
  // counter one item ahead of their prefetch, so SMs that run ahead simply take more items
  int row0 = QK::item_row0((int)blockIdx.x, num_items, rows);
  if (t >= 1 && t < RING) s_next[t] = QK::item_row0((int)(blockIdx.x + (unsigned)t * gridDim.x), num_items, rows);

  // ---- one-time per CTA: inverse permutation as swizzled slot positions.  Eight consecutive permuted positions
  // j = 8*j8 + e occupy 8*SLOT contiguous bytes of xs whose chunks share one swizzle mask, so position e is at
  // base ^ (e * SLOT).
  auto build_tab = [&](const int16_t* idxp) {
  for (int j8 = t; j8 < K8; j8 += T) {
    const uint4 iv = __ldg(reinterpret_cast<const uint4*>(idxp) + j8);
    const uint32_t ivw[4] = {iv.x, iv.y, iv.z, iv.w};
    const uint32_t pc0 = (uint32_t)j8 * (QK::SLOT / 2);  // first 16-byte chunk of the eight slots
    const uint32_t base = ((pc0 ^ ((pc0 >> 3) & QK::SWM)) << 4) + (TABMODE == 2 ? xs_a : 0u);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const uint32_t c = (e & 1) ? (ivw[e >> 1] >> 16) : (ivw[e >> 1] & 0xffffu);
      const uint32_t pos = base ^ (uint32_t)(e * QK::SLOT);
      if constexpr (TABMODE == 2) sts32(tab_a + 4u * c, pos);
      else asm volatile("st.shared.u16 [%0], %1;" ::"r"(tab_a + 2u * c), "h"((uint16_t)(TABMODE == 1 ? (pos >> 2) : pos)) : "memory");
    }
    if constexpr (NORM) {  // the weights of permuted channels 8*j8 .. 8*j8+7
      uint32_t wq[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const uint32_t lo = __ldg(p.nw + (ivw[e] & 0xffffu)), hi = __ldg(p.nw + (ivw[e] >> 16));
        wq[e] = lo | (hi << 16);
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wp_a + 16u * (uint32_t)j8), "r"(wq[0]), "r"(wq[1]),
                   "r"(wq[2]), "r"(wq[3])
                   : "memory");
    }
  }
  };
  int cur_grp = 0;
  if (p.grp_rowblk == nullptr) build_tab(p.idx);
  else cur_grp = -1;  // grouped: the first item's group is only known after griddepcontrol.wait (the router wrote it)
  if constexpr (NORM) {
    for (int i = t; i < NBUF * R * 128; i += T) sts32(ss_a + 4u * (uint32_t)i, 0u);
  }
  // ---- one-time per thread (NP == 1) or per CTA (NP > 1, in shared memory): the compute units' contexts
  UnitCtx ctx0 = QK::make_ctx(p, t, nunits);
  UnitPtrs up0 = QK::make_ptrs(p, ctx0);
  if constexpr (NP > 1) {
    for (int u = t; u < NP * T; u += T) {
      const UnitCtx cx = QK::make_ctx(p, u, nunits);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ctx_a + 16u * (uint32_t)u), "r"(cx.meta), "r"(cx.qoff),
                   "r"(cx.sfo), "r"(cx.xoff)
                   : "memory");
    }
  }
  const uint32_t wp_unit = wp_a + ((ctx0.meta & 64u) ? 32u * (uint32_t)t : 0u);
  const uint16_t* xt = p.x + 8 * t;
  // ... and wait for the previous kernel (which may still be producing X) only now: the table above depends on
  // reorder_index alone, which no kernel of this library writes, so it was built under the previous kernel's tail.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p.rows_dev != nullptr) {
    // (whole 128-row blocks: items past the last used block end this CTA's work like items past the last row)
    rows = min(rows, __ldg(p.rows_dev));
    if (row0 >= rows) row0 = rows;
  }
#pragma unroll
  for (int k = 0; k < DE; ++k) {  // the first DE items' rows
    const int rk = (k == 0) ? row0 : QK::item_row0((int)(blockIdx.x + (unsigned)k * gridDim.x), num_items, (int)p.rows);
    if (rk < rows) {
      if (rk + 32 * (R - 1) < rows) QK::template prefetch<true, EXACT>(p, xt, rk, t, T, K8, pre, raw_a + (uint32_t)k * raw_bytes);
      else QK::template prefetch<false, EXACT>(p, xt, rk, t, T, K8, pre, raw_a + (uint32_t)k * raw_bytes);
    }
    if constexpr (DEPTH > 0) asm volatile("cp.async.commit_group;" ::: "memory");
  }
  if constexpr (MC) {
    // the gather buffers may be overwritten once EVERY rank has finished reading the previous gather (normally long ago)
    if (t == 0) {
      uint32_t issued, seen;
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(issued) : "l"(p.ag_issued) : "memory");
      const uint32_t need = issued * (uint32_t)p.ag_tp;
      unsigned long long t0 = 0;
      for (uint32_t it = 1;; ++it) {
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.ag_consumed) : "memory");
        if ((int32_t)(seen - need) >= 0) break;
        __nanosleep(64);
        if ((it & 1023u) == 0) {  // bounded: a lost peer must not hang the GPU
          unsigned long long now;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
          if (t0 == 0) t0 = now;
          else if (now - t0 > p.ag_timeout_ns) {
            atomicOr(p.ag_err, 4u);
            break;
          }
        }
      }
      // acquire WITHOUT a fence (fence.acq_rel.sys = MEMBAR.ALL.SYS + ERRBAR: ~8 us here, with the first rows' loads in
      // flight and the whole CTA waiting at the barrier below -- profiles/r02_gather_membar_stalls.txt)
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.ag_consumed) : "memory");
    }
  }
  __syncthreads();

  // items are ordered by row: the first item past the last row ends this CTA's work (block-uniform)
  int n = 0;
  while (row0 < rows) {
    if (p.grp_rowblk != nullptr) {
      const int g = __ldg(p.grp_rowblk + (row0 >> 7));  // block-uniform
      if (g != cur_grp) {
        __syncthreads();  // (NORM: the previous item's compute still reads the permuted weights)
        build_tab(p.idx + (int64_t)g * K);
        cur_grp = g;
        __syncthreads();
      }
    }
    const uint32_t xs_cur = xs_a + ((NBUF == 2 && (n & 1)) ? xs_bytes : 0u);
    const uint32_t ss_cur = ss_a + ((NBUF == 2 && (n & 1)) ? (uint32_t)(R * 512) : 0u);
    if (row0 + 32 * (R - 1) < rows)
      row0 = quant_process<QK, R, NLD, NP, NBUF, true, EXACT, NORM, MC, DEPTH>(p, xt, row0, rows, n, t, T, K8, nunits, xs_cur, tab_a, ctx_a, pre, s_next, ctx0, up0, ss_cur, wp_unit, stg_a, stage_row, prev_row0, prev_nvalid, raw_a, raw_bytes);
    else
      row0 = quant_process<QK, R, NLD, NP, NBUF, false, EXACT, NORM, MC, DEPTH>(p, xt, row0, rows, n, t, T, K8, nunits, xs_cur, tab_a, ctx_a, pre, s_next, ctx0, up0, ss_cur, wp_unit, stg_a, stage_row, prev_row0, prev_nvalid, raw_a, raw_bytes);
    ++n;
  }
  if constexpr (MC) {
    __syncthreads();                      // the last item's codes are staged
    if (prev_row0 >= 0) QK::copy_out(p, stg_a + (uint32_t)((n + 1) & 1) * (uint32_t)R * stage_row, stage_row, prev_row0, prev_nvalid, t, T);
    __syncthreads();                      // every thread's multicast stores are issued ...
    if (t == 0) {                         // ... and performed at system scope before this CTA counts as finished
      if (p.ag_dbg & 1u) __threadfence();  // (timing experiment: device scope only -- NOT sufficient for peer visibility)
      else __threadfence_system();
    }
  }
  // the last CTA to leave resets the schedule for the next launch that uses this slot
  if (t == 0) {
    const unsigned int done = atomicInc(p.sched + 1, gridDim.x - 1);  // wraps to 0 by itself
    if (done == gridDim.x - 1) {
      atomicExch(p.sched, 0u);
      if constexpr (MC) {
        // all CTAs' stores have landed everywhere: tell every rank that this rank's rows of the gather are complete
        __threadfence_system();
        uint32_t issued;
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(issued) : "l"(p.ag_issued) : "memory");
        asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p.ag_issued), "r"(issued + 1u) : "memory");
        for (int d = 0; d < p.ag_tp; ++d)
          asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(p.ag_arrived[d]) : "memory");
      }
    }
  }
}

__device__ unsigned int g_quant_sched[64][2];  // rotating schedule slots (zero-initialised, self-resetting)

template <int R, int TMAX, int NLD, int NP, int MINB, int TABMODE, int NBUF, bool NORM = false, bool MC = false, int DEPTH = 0>
static int launch_quant(QuantParams& p, cudaStream_t stream) {
  // threads: NP passes of T threads cover the K/16 compute units exactly (T a multiple of 32)
  const int T = ((p.K / 16 + NP - 1) / NP + 31) & ~31;
  constexpr int TABW = (TABMODE == 2) ? 4 : 2;
  const size_t smem = ((size_t)(p.K * TABW + 127) & ~(size_t)127) + (size_t)NBUF * p.K * 2 * R + (NP > 1 ? (size_t)NP * T * 16 : 0) +
                      (NORM ? (size_t)p.K * 2 + (size_t)NBUF * R * 512 : 0) +
                      (MC ? (size_t)2 * R * (size_t)(p.rowbytes[0] + p.rowbytes[1] + p.rowbytes[2]) : 0) +
                      (size_t)DEPTH * R * NLD * 16 * T;
  const int64_t xs_bytes = (int64_t)p.K * 2 * R;
  if (smem > 227 * 1024 || T > TMAX || (int64_t)NLD * T * 8 < p.K || (TABMODE == 0 && xs_bytes > 65536) ||
      (TABMODE == 1 && xs_bytes > 4 * 65536)) {
    set_error("reorder_quantize: K=%d does not fit the <%d,%d,%d,%d,%d> kernel (%zu bytes of shared memory)", p.K, R, TMAX,
              NLD, NP, TABMODE * 10 + NBUF, smem);
    return MMX_ERR_INVALID;
  }
  const bool exact = (int64_t)NLD * T * 8 == p.K && (int64_t)NP * T * 16 == p.K;
  auto kern = exact ? reorder_quantize_kernel<R, TMAX, NLD, NP, MINB, TABMODE, NBUF, true, NORM, MC, DEPTH>
                    : reorder_quantize_kernel<R, TMAX, NLD, NP, MINB, TABMODE, NBUF, false, NORM, MC, DEPTH>;
  // per-device, per-variant launch facts (dynamic shared memory opt-in, occupancy), guarded: callers may be threads
  struct Cache {
    size_t attr_smem = 0;  // largest opt-in granted so far
    size_t occ_smem = 0;
    int occ_T = 0;
    int occ = 0;
  };
  static Cache cache[kMaxDevices][2];
  static std::mutex mu;
  const int dev = current_device_slot();
  int occ_cached;
  {
    std::lock_guard<std::mutex> lk(mu);
    Cache& c = cache[dev][exact];
    if (smem > c.attr_smem) {
      MMX_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      c.attr_smem = smem;
    }
    if (c.occ == 0 || c.occ_smem != smem || c.occ_T != T) {
      int occ = 0;
      MMX_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, T, smem));
      c.occ = occ < 1 ? 1 : occ;
      c.occ_smem = smem;
      c.occ_T = T;
    }
    occ_cached = c.occ;
  }
  const int64_t rblocks = (p.rows + 127) / 128;
  const int64_t items = rblocks * 32 * (4 / R);
  p.num_items = (int)items;
  int64_t grid = (int64_t)sm_count() * occ_cached;  // every SM full; items beyond two rounds are claimed dynamically
  if (options().quant_ctas > 0) grid = options().quant_ctas;
  if (grid > items) grid = items;
  if (grid < 1) {
    if (!MC) return MMX_OK;
    grid = 1;  // no rows on this rank: one CTA still waits, then announces the (empty) shard
  }
  static unsigned int* sched_base[kMaxDevices] = {};  // the symbol's address differs from device to device
  if (!sched_base[dev]) MMX_CUDA_TRY(cudaGetSymbolAddress(reinterpret_cast<void**>(&sched_base[dev]), g_quant_sched));
  static std::atomic<unsigned int> seq{0};
  p.sched = sched_base[dev] + 2 * (seq.fetch_add(1, std::memory_order_relaxed) & 63u);
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

int reorder_quantize(const void* x, int64_t rows, int K, const int16_t* idx, int KN, int KS, int KO,
                     const int fmt[3], uint8_t* q0, uint8_t* q1, uint8_t* q2, uint8_t* s0, uint8_t* s1,
                     uint8_t* s2, void* stream, const void* norm_w, float eps, bool norm, const QuantGather* ag,
                     const int* grp_rowblk, const int* row_src, const int* rows_dev) {
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
  if (norm && (!norm_w || ((uintptr_t)norm_w & 15))) {
    set_error("rmsnorm_quantize_x: the norm weight must be a non-null 16-byte aligned pointer");
    return MMX_ERR_INVALID;
  }
  if (norm && K > 16384) {
    set_error("rmsnorm_quantize_x: K=%d exceeds the supported 16384", K);
    return MMX_ERR_INVALID;
  }
  if (rows == 0 && ag == nullptr) return MMX_OK;  // (a gather takes part with zero rows too: its arrival is awaited)
  QuantParams p;
  memset(&p, 0, sizeof(p));
  if (ag != nullptr) {
    p.ag_consumed = ag->consumed;
    p.ag_issued = ag->issued;
    p.ag_tp = ag->tp;
    p.ag_timeout_ns = (unsigned long long)options().tp_timeout_ms * 1000000ull;
    p.ag_err = ag->err;
    p.a2a_per = ag->a2a_per;
    if (ag->a2a_per > 0) {
      for (int d = 0; d < ag->tp && d < kMaxTp; ++d)
        for (int i = 0; i < 3; ++i) {
          p.qd[d][i] = ag->qd[d][i];
          p.sfd[d][i] = ag->sfd[d][i];
        }
    }
    p.ag_dbg = (uint32_t)(options().tp_debug >> 5);
    for (int d = 0; d < ag->tp && d < kMaxTp; ++d) p.ag_arrived[d] = ag->arrived[d];
  }
  p.grp_rowblk = grp_rowblk;
  p.row_src = row_src;
  p.rows_dev = rows_dev;
  p.nw = static_cast<const uint16_t*>(norm_w);
  p.eps = eps;
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
  if (ag != nullptr && ag->a2a_per > 0) {
    for (int i = 0; i < 3; ++i) {
      p.katoms[i] = ag->a2a_katoms[i];       // scale row blocks of the DESTINATION hold the atoms of all ranks
      p.a2a_pitch[i] = ag->a2a_pitch[i];
    }
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int force = (int)options().quant_rows;
  const int var = (int)options().quant_variant;
  // configuration by K: rows per item R, thread bound, 16-byte chunks per thread and row NLD, compute passes NP,
  // minimum CTAs per SM (register bound), table encoding
  // (measured on B200, profiles/r01_quantize_sweep.log: all K <= 4096 variants are within 3 % of each other; the
  // double-buffered form wins by 2-3 % where one CTA owns the whole SM)
  if (ag != nullptr) {  // sequence-parallel hand-over: multicast stores + arrival (one configuration per K range)
    if (K > 16384) {
      set_error("quantize_allgather: K=%d exceeds the supported 16384", K);
      return MMX_ERR_INVALID;
    }
    if (norm) {
      if (K <= 4096) return launch_quant<2, 256, 2, 1, 3, 0, 1, true, true>(p, st);
      if (K <= 8192) return launch_quant<2, 512, 2, 1, 1, 0, 2, true, true>(p, st);
      return launch_quant<2, 1024, 2, 1, 1, 0, 2, true, true>(p, st);
    }
    if (K <= 4096) return launch_quant<2, 256, 2, 1, 4, 0, 1, false, true>(p, st);
    if (K <= 8192) return launch_quant<2, 512, 2, 1, 2, 0, 2, false, true>(p, st);
    return launch_quant<2, 1024, 2, 1, 1, 0, 2, false, true>(p, st);
  }
  if (norm) {
    if (K <= 4096) return launch_quant<2, 256, 2, 1, 3, 0, 1, true>(p, st);
    if (K <= 8192) return launch_quant<2, 512, 2, 1, 1, 0, 2, true>(p, st);
    return launch_quant<2, 1024, 2, 1, 1, 0, 2, true>(p, st);
  }
  // quant_variant: 0 = default, 1 = the other xs buffering, 2 / 3 = cp.async ring of 2 / 3 items (experiments)
  if (K <= 4096) {
    if (force == 4) return launch_quant<4, 256, 2, 1, 2, 0, 2>(p, st);
    if (var == 1) return launch_quant<2, 256, 2, 1, 4, 0, 2>(p, st);
    if (var == 2) return launch_quant<2, 256, 2, 1, 4, 0, 1, false, false, 2>(p, st);
    if (var == 3) return launch_quant<2, 256, 2, 1, 3, 0, 1, false, false, 3>(p, st);
    return launch_quant<2, 256, 2, 1, 4, 0, 1>(p, st);
  }
  if (K <= 8192) {
    if (force == 4) return launch_quant<4, 512, 2, 1, 1, 0, 2>(p, st);
    if (var == 2) return launch_quant<2, 512, 2, 1, 2, 0, 1, false, false, 2>(p, st);
    return var == 1 ? launch_quant<2, 512, 2, 1, 2, 0, 1>(p, st) : launch_quant<2, 512, 2, 1, 2, 0, 2>(p, st);
  }
  if (K <= 16384) {
    if (var == 2) return launch_quant<2, 1024, 2, 1, 1, 0, 1, false, false, 2>(p, st);
    // two units per thread, two CTAs per SM: 9 % faster at K = 11008, 19 % slower at 14336 (profiles/r02_quantize_depth.log)
    if (var == 4 || (var == 0 && K <= 11264)) return launch_quant<2, 512, 4, 2, 2, 0, 1>(p, st);
    return var == 1 ? launch_quant<2, 1024, 2, 1, 1, 0, 1>(p, st) : launch_quant<2, 1024, 2, 1, 1, 0, 2>(p, st);
  }
  return launch_quant<2, 1024, 4, 2, 1, 1, 1>(p, st);
}

}  // namespace mmx

extern "C" __attribute__((visibility("default"))) int mmx_reorder_quantize_x(const void* x, int64_t M, int K, const int16_t* idx, int KN, int KS, int KO,
                                      uint8_t* xn, uint8_t* xs, uint8_t* xo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo,
                                      void* stream) {
  const int fmt[3] = {4, 6, 8};
  return mmx::reorder_quantize(x, M, K, idx, KN, KS, KO, fmt, xn, xs, xo, sfn, sfs, sfo, stream, nullptr, 0.0f, false, nullptr, nullptr);
}

extern "C" __attribute__((visibility("default"))) int mmx_reorder_quantize_w(const void* w, int64_t N, int K, const int16_t* idx, int KN, int KS, int KO,
                                      uint8_t* wn, uint8_t* ws, uint8_t* wo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo,
                                      void* stream) {
  const int fmt[3] = {4, 6, 8};
  return mmx::reorder_quantize(w, N, K, idx, KN, KS, KO, fmt, wn, ws, wo, sfn, sfs, sfo, stream, nullptr, 0.0f, false, nullptr, nullptr);
}

extern "C" __attribute__((visibility("default"))) int mmx_reorder_quantize_w4(const void* w, int64_t N, int K, const int16_t* idx, int KN, int KS, int KO,
                                       uint8_t* wn, uint8_t* ws, uint8_t* wo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo,
                                       void* stream) {
  const int fmt[3] = {4, 4, 4};
  return mmx::reorder_quantize(w, N, K, idx, KN, KS, KO, fmt, wn, ws, wo, sfn, sfs, sfo, stream, nullptr, 0.0f, false, nullptr, nullptr);
}

extern "C" __attribute__((visibility("default"))) int mmx_rmsnorm_quantize_x(const void* x, const void* w, float eps, int64_t M, int K, const int16_t* idx,
                                      int KN, int KS, int KO, uint8_t* xn, uint8_t* xs, uint8_t* xo, uint8_t* sfn,
                                      uint8_t* sfs, uint8_t* sfo, void* stream) {
  const int fmt[3] = {4, 6, 8};
  return mmx::reorder_quantize(x, M, K, idx, KN, KS, KO, fmt, xn, xs, xo, sfn, sfs, sfo, stream, w, eps, true, nullptr, nullptr);
}

// Grouped form (Mixtral experts, extension; the reference loops over experts in Python, model/qMixtralLayer.py:437-450):
// rows sorted by group, each 128-row block owned by ONE group; idx = int16 [groups, K], grp_rowblk = int32 [ceil(M/128)]
// in device memory (written on the stream -- no host synchronisation), the same (KN, KS, KO) for every group.
// row_src (optional) = int32 [M]: sorted row r is row row_src[r] of x -- the gather of the routed tokens is fused.
// rows_dev (optional) = int32 in device memory: rows that actually exist (a multiple of 128); M is the static upper bound.
extern "C" __attribute__((visibility("default"))) int mmx_reorder_quantize_x_grouped(
    const void* x, int64_t M, int K, const int16_t* idx, const int32_t* grp_rowblk, const int32_t* row_src,
    const int32_t* rows_dev, int KN, int KS, int KO, uint8_t* xn, uint8_t* xs, uint8_t* xo, uint8_t* sfn, uint8_t* sfs,
    uint8_t* sfo, void* stream) {
  if (!grp_rowblk) {
    mmx::set_error("reorder_quantize_x_grouped: null group table");
    return MMX_ERR_INVALID;
  }
  const int fmt[3] = {4, 6, 8};
  return mmx::reorder_quantize(x, M, K, idx, KN, KS, KO, fmt, xn, xs, xo, sfn, sfs, sfo, stream, nullptr, 0.0f, false, nullptr,
                               grp_rowblk, row_src, rows_dev);
}
