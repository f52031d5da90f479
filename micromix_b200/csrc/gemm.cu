// gemm.cu -- three-segment mixed-MX GEMM as ONE persistent tcgen05 kernel for sm_100a.
//
// Replaces /root/reference/mgemm/src/gemm.cu:26-78 (matmul_host / matmul_w4_host: three CUTLASS Sm120 mma.sync
// launches chained through a bf16 D in HBM, w4a4.cu:176 / w4a6.cu:178 / w4a8.cu:178) behind mmx_matmul.
//
//   C[M,N] = sum_seg (A_seg o SFA_seg) (B_seg o SFB_seg)^T ,  seg in {FP4xFP4, FP6xFP4|FP6, FP8xFP4|FP8}
//
// Kernel shape
//   * persistent, one CTA per SM (CTA pairs, cta_group::2, when M > 128), 192 threads: warp 0 = TMA producer,
//     warp 1 = MMA issuer (+TMEM owner), warps 2..5 = epilogue (TMEM lane quarter = warp % 4).
//   * tile 128 x 256 per CTA (256 x 256 per pair); all three K segments accumulate into ONE fp32 TMEM accumulator,
//     so the output is rounded to bf16 exactly once and C is written exactly once (no memset, no beta=1 re-reads).
//   * TMEM: two 256-column accumulators that overlap in 96 columns + 4 rotating scale-factor slots (see kAccOverlap):
//     tile i+1 starts as soon as the epilogue of tile i has drained the shared columns, the rest of the drain
//     (TMEM reads run at ~64 B/clk) hides behind the next tile's MMAs.
//   * smem pipeline of 6 stages (4 for the single-CTA kernel); a stage is [A 128 x 128 B][B 128|256 x 128 B][SFA][SFB],
//     128B-swizzled:
//       FP4xFP4 segment : kind::mxf4.block_scale.block32, packed nibbles, 256 K per stage, 4 MMAs of K=64
//       other segments  : kind::mxf8f6f4.block_scale, TMA expands FP6/FP4 to the 16-byte-aligned "unpacked"
//                         smem form (CU_TENSOR_MAP_DATA_TYPE_16U6_ALIGN16B / 16U4_ALIGN16B), 128 K per stage,
//                         4 MMAs of K=32
//     so both kinds advance the smem descriptors by 32 bytes per MMA and use identical stage geometry.
//   * scale factors: the gmem layout (512-byte SfKMajorAtom per 128 rows x 128 K) is already the layout
//     tcgen05.cp.32x128b.warpx4 wants, so SF tiles go gmem -> smem by TMA (plain u32 boxes) and
//     smem -> TMEM with tcgen05.cp issued by the MMA thread right before the stage's MMAs.
//   * the MMA-issuing warp is the kernel's critical path (a stage is only ~512 tensor-pipe cycles): its stage body is
//     straight-line code specialised on (kind, scale atoms) with running descriptor words -- see the warp == 1 branch.
//   * epilogue: tcgen05.ld 32x32b.x32 -> fp32 -> bf16 (+ optional bias, rounded like the reference's separate add)
//     -> 64B-swizzled smem staging -> TMA stores (clipped at M and N by the tensor map).
#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>
#include <type_traits>
#include <unordered_map>

#include "act_math.h"
#include "common.h"

namespace mmx {

// Tile geometry.  CG = CTAs cooperating on one MMA (tcgen05 cta_group): 1 -> 128x256 tile per CTA,
// 2 -> 256x256 tile per CTA pair (each CTA holds its own 128 A rows and HALF of the B rows, so L2->SM operand
// traffic per flop drops by a third; the kernel is bound by that traffic, see DESIGN.md / profiles/).
constexpr int BM = 128;   // A rows per CTA
constexpr int BN = 256;   // N per tile
constexpr int kThreads = 192;
constexpr uint32_t kTmemCols = 512;
// TMEM map (512 columns): two fp32 accumulators of 256 columns that OVERLAP in kAccOverlap columns, then the scale
// factors.  Tile i accumulates in acc[i & 1]; its epilogue drains the overlap columns first and hands them back, so
// the MMAs of tile i+1 start after ~kAccOverlap/256 of a drain instead of a whole one (TMEM cannot hold two full
// 128x256 accumulators plus scales).
constexpr uint32_t kSFSlot = 24;     // scale columns of one stage: SFA 8 = 2 atoms x 4, SFB 16 = 2 n-blocks x 2 atoms x 4
// (Round 2 experiment: ONE slot and a 32-column overlap -- tcgen05.cp is ordered behind earlier MMAs of the same thread, the
// way CUTLASS's block-scaled sm100 mainloop uses a single scale buffer -- passes every GEMM test and changes nothing:
// qkv 104.3 / o 68.3 / gate_up 428.6 us against 103.3 / 68.4 / 425.5 us, profiles/r02_gemm_overlap32.txt.  The step runs
// under sw_power_cap: the tensor pipe's duty cycle is set by the power limit, not by the drain the next tile waits for.)
constexpr uint32_t kNumSFSlots = 4;  // rotating slots: a stage's tcgen05.cp never lands on scales that MMAs in flight read
constexpr uint32_t kAccOverlap = 96;  // >= kSFSlot * kNumSFSlots, multiple of 32 (epilogue chunk)
constexpr uint32_t kColAcc1 = 256 - kAccOverlap;
constexpr uint32_t kColSF = 512 - kAccOverlap;
static_assert(kSFSlot * kNumSFSlots <= kAccOverlap && kAccOverlap % 32 == 0, "TMEM map");
constexpr int kEpiBuf = 2048;      // 32 rows x 32 bf16 (64-byte rows, 64B swizzle): source of one TMA store
constexpr int kEpiStage = 2 * kEpiBuf;  // two buffers per epilogue warp: stage chunk i+1 while chunk i is read

// RS (fused row-parallel mode): one stage less, so that a tile_allreduce_kernel CTA (no dynamic shared memory, but
// every resident CTA reserves 1 KB) fits on the SM NEXT TO this kernel's CTA -- with six stages the pair kernel leaves
// 768 bytes and the reducer could only start once the GEMM had left.
template <int CG, bool RS = false>
struct Geo {
  static constexpr int kStages = (CG == 1) ? 4 : (RS ? 5 : 6);
  static constexpr int kBRows = BN / CG;          // B rows this CTA stages
  static constexpr int kStageA = BM * 128;        // 16 KB
  static constexpr int kStageB = kBRows * 128;    // 32 | 16 KB
  static constexpr int kStageSFA = 2 * 512;       // [atom][512]
  static constexpr int kStageSFB = 2 * 2 * 512;   // [n-block][atom][512], all 256 N columns in every CTA
  static constexpr int kStageBytes = kStageA + kStageB + kStageSFA + kStageSFB;
  static constexpr int kEpiOff = kStages * kStageBytes;        // 4 x kEpiStage, 1024-aligned
  static constexpr int kBarOff = kEpiOff + 4 * kEpiStage;
  static constexpr int kSmemBytes = kBarOff + 256 /*barriers*/;
};

struct GemmSeg {
  int ktiles;          // pipeline stages this segment contributes per output tile
  int kind;            // 0: kind::mxf4 (256 K per stage)   1: kind::mxf8f6f4 (128 K per stage)
  int kelems;          // K elements per stage
  int atoms_per_tile;  // 128-K scale atoms per stage: 2 (mxf4) or 1
  int last_atoms;      // atoms in the last stage (1 when an mxf4 segment has Kseg % 256 == 128)
  uint32_t idesc;      // instruction descriptor, sf ids zero
  uint32_t tx_a;       // bytes one CTA's A tile load completes on the stage barrier
  uint32_t tx_b;       // ... B tile load (per CTA)
};

struct GemmParams {
  GemmSeg seg[3];
  int nseg;
  int m_tiles, n_tiles;  // in units of (CG*128) x 256
  int n_fastest;         // tile order: 1 = consecutive tiles walk N (all of B stays L2-resident per wave), 0 = walk M,
                         // 2 = OWNER-INTERLEAVED (reduce-scatter): tile t belongs to rank t % own_tp, whose j = t / own_tp-th
                         // tile is (m_blk = o * m_per + j / n_tiles, n_blk = j % n_tiles) -- every rank walks the same order,
                         // so all owners' tiles complete at an even pace and every owner's rows are contiguous
  int own_tp, m_per;     // raster 2: ranks and m-tiles per rank (ceil); tiles with m_blk >= m_tiles do not exist
  int num_tiles;         // loop bound of the tile walk (raster 2: own_tp * m_per * n_tiles, else m_tiles * n_tiles)
  // GROUPED GEMM (Mixtral experts): A = token rows sorted by group and padded per group to whole m-tiles; the B tensors
  // hold the groups' weights stacked on N.  grp_mblk[m_blk] = group of that m-tile, or -1 = padding tile (skipped).
  const int* grp_mblk;
  int grp_n;             // B rows per group
  const int* rows_dev;   // grouped, optional (device memory): rows of A that exist; bounds the tile walk
  // GATHERED A (sequence-parallel hand-over, tp_reduce.cu): the A rows of source rank s = row / ag_rows are valid once
  // ag_arrived[32 * s] (one counter per 128-byte line) has reached ag_taken[0] + 1 (the peers' quantizers multicast them
  // and then bump the counter)
  const uint32_t* ag_arrived;
  uint32_t* ag_taken;
  int ag_rows;
  // ... and once the LAST CTA of this grid is done with A it bumps ag_taken and tells every rank (ag_consumed[d], peer
  // mapped) that this rank no longer reads the gathered buffers, so that the next gather may overwrite them
  uint32_t* ag_ticket;
  uint32_t* ag_consumed[kMaxTp];
  int ag_tp;
  uint32_t* ag_err;
  unsigned long long ag_timeout_ns;
  // FUSED ACTIVATION (ACT kernels, mmx_matmul_activate_quantize): B = gate / up rows INTERLEAVED in blocks of 128 channels
  // (tile n_blk: columns 0..127 = gate of channels [128 n_blk, +128), columns 128..255 = up of the same channels), and the
  // epilogue does not write C: it rounds both accumulators to bf16 (what the plain GEMM would have stored), evaluates
  // v = silu(gate) * up in fp32 and MX-quantizes v over its 32-channel groups exactly like mmx_activate_quantize_x
  // (rowquant.cu) -- codes and scale bytes of the down projection's operand leave the kernel directly.
  uint8_t* act_q[3];
  uint8_t* act_sf[3];
  int act_cend[3];       // cumulative channel ends of the FP4 | FP6 | FP8 segments of the activation
  int act_katoms[3];     // scale atoms per row block of each segment
  int64_t act_rowbytes[3];
  int64_t M, N;
  __nv_bfloat16* c;
  const __nv_bfloat16* bias;
  // optional bf16 [M, N] (row stride N), added AFTER the product was rounded to bf16 (and after the bias):
  // c = bf16(residual + bf16(acc)) -- the decoder layer's `residual + o_proj(...)` / `residual + down_proj(...)`
  // (model/qLlamaLayer.py:116-158) without a separate elementwise kernel; may alias c
  const __nv_bfloat16* residual;
  // ROPE kernels (mmx_matmul_rope): the first rope_cols columns of C are q / k heads of 128 channels whose B rows were stored
  // PAIR-ADJACENT (row 2j of a head = channel j, row 2j + 1 = channel j + 64), so a lane holds both partners of the rotary
  // embedding in neighbouring accumulator columns; the epilogue applies HF's apply_rotary_pos_emb with its three bf16
  // roundings (rope.cu) and stores every value at its ORIGINAL column -- C is what matmul + mmx_rope_inplace produce.
  const __nv_bfloat16* rope_cos;  // bf16 [rope_S, 128]; row m uses table row m % rope_S
  const __nv_bfloat16* rope_sin;
  int rope_cols, rope_S;
  uint32_t* dbg;
  uint32_t flags;  // watchdog build only: 1 = skip SF copies, 2 = skip MMAs, 4 = skip C stores
};

struct alignas(64) TmapSet {
  CUtensorMap a[3];
  CUtensorMap b[3];
  CUtensorMap sfa[3];  // 3-D [row-block][k-atom][128 x u32] views of the 512-byte scale atoms
  CUtensorMap sfb[3];
  CUtensorMap c;       // bf16 [M, N], box 32 x 32, 64B swizzle: epilogue TMA stores
};

// Fused row-parallel mode (RS = true, driven by tp_reduce.cu): instead of writing C, the epilogue PUSHES every partial
// tile into the staging buffer of the rank that owns it (owner = tile % tp, slot = this rank) -- a peer-mapped address,
// so the TMA store travels over NVLink -- and then bumps the owner's per-tile counter with a system-scope release.
// The owner's tile_allreduce_kernel (co-resident with this kernel through programmatic dependent launch) sums the tp
// slots as soon as a tile's counter is full and writes the bf16 result to every rank's C.
struct alignas(64) RsParams {
  CUtensorMap dst[kMaxTp];      // bf16 [own_tiles_cap * tile_rows, 256] views of MY slot in rank d's staging buffer
  uint32_t* tile_flags[kMaxTp];  // rank d's per-owned-tile arrival counters (peer-mapped)
  int tp, rank;
  int rot_s;  // owner rotation period (a multiple of tp): owner(tile) = (tile + tile / tp * tp / rot_s) % tp
  int pull;   // 1: partial tiles go into this rank's own C (tmaps.c) and are reduced in the switch by the owner
};
struct NoRsParams {
  int unused;
};

__device__ uint32_t g_gemm_dbg[64];

// tile index -> (m_blk, n_blk, group); false = the tile does not exist (raster-2 round-up, grouped padding tile)
__device__ __forceinline__ bool tile_coords(const GemmParams& p, int tile, int& m_blk, int& n_blk, int& grp) {
  grp = 0;
  if (p.n_fastest == 2) {
    const int o = tile % p.own_tp, j = tile / p.own_tp;
    m_blk = o * p.m_per + j / p.n_tiles;
    n_blk = j % p.n_tiles;
    return m_blk < p.m_tiles;
  }
  m_blk = p.n_fastest ? tile / p.n_tiles : tile % p.m_tiles;
  n_blk = p.n_fastest ? tile % p.n_tiles : tile / p.m_tiles;
  if (p.grp_mblk != nullptr) {
    grp = __ldg(p.grp_mblk + m_blk);
    return grp >= 0;
  }
  return true;
}

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n.reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n}" ::"r"(bar), "r"(rank)
      : "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ uint32_t mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// WD = watchdog build: bounded spin, returns false on timeout so a wrong descriptor cannot hang the GPU.
template <bool WD>
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  if constexpr (WD) {
    for (uint32_t spins = 0; spins < (1u << 24); ++spins)
      if (mbar_test_wait(bar, parity)) return true;
    return false;
  } else {
    while (!mbar_try_wait(bar, parity)) {
    }
    return true;
  }
}

// true in exactly one lane of the (converged) warp; ptxas knows an elect.sync region is single-lane, so the
// uniform-register operands of tcgen05 / TMA instructions inside it need no per-lane "waterfall" loop
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n.reg .pred p;\n"
      "elect.sync _|p, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, p;\n}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ long long clk() {
  long long c;
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(c));
  return c;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// TMA tile loads.  CG == 2: the .cta_group::2 form lets the transaction bytes land on the LEADER CTA's barrier
// (`bar` is then a shared::cluster address produced by mapa).
template <int CG>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  if constexpr (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
        "[%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                            int c2) {
  if constexpr (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
        "%5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  }
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int CG>
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t cols) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// MMA-completion arrive; CG == 2 multicasts the arrive to the barrier at the same offset in both CTAs of the pair
template <int CG>
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    const uint16_t mask = 3;
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
        "h"(mask)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tc_cp_sf(uint32_t tmem_dst, uint64_t desc) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(desc) : "memory");
  else
    asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(desc) : "memory");
}
template <int CG>
__device__ __forceinline__ void mma_mxf4(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t sfa,
                                         uint32_t sfb, uint32_t acc) {
  if constexpr (CG == 1) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n}" ::"r"(d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.block32 [%0], %1, %2, %3, [%5], [%6], p;\n}" ::"r"(d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void mma_mxf8f6f4(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t sfa,
                                             uint32_t sfb, uint32_t acc) {
  if constexpr (CG == 1) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n}" ::"r"(d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb)
        : "memory");
  } else {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n}" ::"r"(d),
        "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(sfa), "r"(sfb)
        : "memory");
  }
}

// K-major, 128B-swizzled operand tile (rows of 128 bytes, 8-row groups 1024 bytes apart).
// bits: [0,14) addr>>4 | [16,30) LBO>>4 (unused for swizzled K-major) | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Scale-factor chunk: 32 rows x 16 bytes, no swizzle; 8-row core matrices 128 bytes apart (SBO), one atom along K.
__device__ __forceinline__ uint64_t make_desc_sf(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)(128 >> 4) << 32) | (1ull << 46);
}

// The same descriptors from a precomputed low word ((smem address >> 4) & 0x3fff, advanced by immediates) and the
// constant high word, so the per-stage code is one 32-bit add per descriptor.
constexpr uint32_t kDescHiOp = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024, version 1, SWIZZLE_128B
constexpr uint32_t kDescHiSF = (128u >> 4) | (1u << 14);                // SBO 128, version 1, no swizzle
__device__ __forceinline__ uint64_t desc_from_lo(uint32_t lo, uint32_t hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
  return d;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// 32 fp32 accumulator columns of one row -> 16 packed bf16x2 words (+bias, rounded like the reference's separate add)
// y = bf16(y + r) on packed pairs (torch's bf16 add: fp32 sum, one rounding)
__device__ __forceinline__ uint32_t add_bf16x2(uint32_t y, uint32_t r) {
  return pack_bf16(__uint_as_float(y << 16) + __uint_as_float(r << 16),
                   __uint_as_float(y & 0xffff0000u) + __uint_as_float(r & 0xffff0000u));
}
__device__ __forceinline__ void add_residual(uint32_t* o, const uint4 (&res)[4]) {
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    o[4 * v] = add_bf16x2(o[4 * v], res[v].x);
    o[4 * v + 1] = add_bf16x2(o[4 * v + 1], res[v].y);
    o[4 * v + 2] = add_bf16x2(o[4 * v + 2], res[v].z);
    o[4 * v + 3] = add_bf16x2(o[4 * v + 3], res[v].w);
  }
}

__device__ __forceinline__ void pack_chunk(const uint32_t (&r)[32], uint32_t* o, const __nv_bfloat16* bias) {
  if (bias != nullptr) {
    const uint4* bp = reinterpret_cast<const uint4*>(bias);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const uint4 bv = __ldg(bp + v);
      const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = v * 8 + u * 2;
        // y = bf16(acc); y = bf16(y + bias): the reference's matmul followed by `y + self.bias`
        const uint32_t y = pack_bf16(__uint_as_float(r[e]), __uint_as_float(r[e + 1]));
        const float y0 = __uint_as_float(y << 16) + __uint_as_float(bw[u] << 16);
        const float y1 = __uint_as_float(y & 0xffff0000u) + __uint_as_float(bw[u] & 0xffff0000u);
        o[v * 4 + u] = pack_bf16(y0, y1);
      }
    }
  } else {
#pragma unroll
    for (int e = 0; e < 16; ++e) o[e] = pack_bf16(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
  }
}

// 16 packed words (32 bf16 of one row) -> four 16-byte chunks of a 32x32 staging buffer.  Row r is 64 bytes; chunk c
// lives at ((c ^ ((r >> 1) & 3)) << 4): the 64B-swizzle pattern of the C tensor map, conflict-free for the warp.
__device__ __forceinline__ void stage_chunk(const uint32_t* o, uint32_t sbuf, int lane) {
  const uint32_t srow = sbuf + (uint32_t)lane * 64u;
  const uint32_t sw = (uint32_t)(lane >> 1) & 3u;
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + (((uint32_t)v ^ sw) << 4)), "r"(o[4 * v]),
                 "r"(o[4 * v + 1]), "r"(o[4 * v + 2]), "r"(o[4 * v + 3])
                 : "memory");
  }
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ the kernel
// SK > 1 (small M, single-CTA tiles only): SPLIT-K over a cluster of SK CTAs.  A 128 x 256 tile of a decode-sized GEMM
// is bound by streaming its B panel, and N / 256 CTAs cannot pull the weights at HBM rate; here SK CTAs each accumulate a
// contiguous 1/SK of the tile's K stages in their own TMEM, park the fp32 accumulator in their (now idle) stage buffers,
// and CTA r of the cluster sums column slice r of all SK copies through distributed shared memory -- in CTA order, so the
// result does not depend on timing -- and writes bf16.  One tile per cluster, no global workspace, no atomics.
template <int CG, bool WD, bool RS, int SK = 1, bool ACT = false, bool ROPE = false>
__global__ void __launch_bounds__(kThreads, 1)
mixed_gemm_kernel(const __grid_constant__ TmapSet tmaps, const __grid_constant__ GemmParams p,
                  const __grid_constant__ std::conditional_t<RS, RsParams, NoRsParams> rs) {
  static_assert(SK == 1 || (CG == 1 && !WD && !RS), "split-K is a variant of the plain single-CTA kernel");
  static_assert(!ACT || (SK == 1 && !WD && !RS), "the fused activation is a variant of the plain kernels");
  static_assert(!ROPE || (SK == 1 && !WD && !RS && !ACT), "the fused rotary embedding is a variant of the plain kernels");
  using G = Geo<CG, RS>;
  constexpr int kStages = G::kStages;
  extern __shared__ __align__(1024) uint8_t smem_raw[];  // no static smem in this kernel: offset 0 of the window
  const uint32_t smem_base = smem_u32(smem_raw);
  const uint32_t bar_base = smem_base + G::kBarOff;
  // barriers: full[kStages] | empty[kStages] | tmem_full | tmem_empty | tmem_ptr(u32)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kStages);
  const uint32_t tmem_empty_bar = bar_base + 8u * (2 * kStages + 1);
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * kStages + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;  // 0 = leader of the pair
  const int group = (CG == 2) ? (blockIdx.x >> 1) : (SK > 1 ? (int)(blockIdx.x / SK) : (int)blockIdx.x);
  const int ngroups = (CG == 2) ? (gridDim.x >> 1) : (SK > 1 ? (int)(gridDim.x / SK) : (int)gridDim.x);
  int num_tiles = p.num_tiles;
  // split-K: this CTA's share [sk_lo, sk_hi) of the tile's concatenated stage list (all segments, in order)
  int sk_lo = 0, sk_hi = 0x7fffffff;
  uint32_t krank = 0;
  if constexpr (SK > 1) {
    krank = cluster_ctarank();
    int total = 0;
    for (int s = 0; s < p.nseg; ++s) total += p.seg[s].ktiles;
    sk_lo = (int)((long long)krank * total / SK);
    sk_hi = (int)((long long)(krank + 1) * total / SK);
  }

  // programmatic dependent launch: the next kernel in the stream may start its prologue as SMs free up
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);   // the leader's producer arrives once (with the expected bytes of the whole pair)
      mbar_init(empty_bar(s), 1);  // one (multicast) tcgen05.commit
    }
    mbar_init(tmem_full_bar, 1);
    mbar_init(tmem_empty_bar, 4 * CG);  // one elected lane per epilogue warp, both CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc<CG>(tmem_ptr_smem, kTmemCols);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      prefetch_tmap(&tmaps.a[s]);
      prefetch_tmap(&tmaps.b[s]);
      prefetch_tmap(&tmaps.sfa[s]);
      prefetch_tmap(&tmaps.sfb[s]);
    }
    if constexpr (RS) {
      for (int d = 0; d < rs.tp; ++d) prefetch_tmap(&rs.dst[d]);
    } else {
      prefetch_tmap(&tmaps.c);
    }
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  tmem_base = __shfl_sync(0xffffffffu, tmem_base, 0);  // provably warp-uniform
  // everything above overlapped the previous kernel's tail; its outputs (our operands) are complete after this
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p.rows_dev != nullptr) {
    // grouped: only the first *rows_dev rows of A exist (whole m-tiles, a prefix of the walk because tiles walk N
    // fastest); the m-tiles behind them are padding of the static upper bound and are not even looked at
    const int used_m = __ldg(p.rows_dev) / (CG * BM);
    num_tiles = min(num_tiles, used_m * p.n_tiles);
  }

  if (warp == 0) {
    // ======================================================================== TMA producer (one lane per CTA)
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      bool ok = true;
      long long t_wait = 0, t_begin = WD ? clk() : 0;
      uint32_t ag_target = 0;
      if (p.ag_arrived != nullptr) {
        asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(ag_target) : "l"(p.ag_taken) : "memory");
        ag_target += 1u;
      }
      uint32_t ag_ok_mask = 0;  // gathered A: source ranks whose arrival this CTA has already observed
      auto ag_wait = [&](int src) {
        if ((ag_ok_mask >> src) & 1u) return;
        uint32_t v;
        unsigned long long t0 = 0;
        for (uint32_t it = 1;; ++it) {
          asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ag_arrived + 32 * src) : "memory");
          if ((int32_t)(v - ag_target) >= 0) break;
          __nanosleep(64);
          if ((it & 1023u) == 0) {  // bounded: a lost peer costs wrong rows (flagged), never a hung GPU
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > p.ag_timeout_ns) {
              atomicOr(p.ag_err, 8u);
              break;
            }
          }
        }
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p.ag_arrived + 32 * src) : "memory");
        asm volatile("fence.proxy.async.global;" ::: "memory");  // the TMA loads below read what the peers wrote
        ag_ok_mask |= 1u << src;
      };
      // exchanged A (all-to-all of K slices: ag_rows == 0): every source rank contributes columns to EVERY row
      if (p.ag_arrived != nullptr && p.ag_rows == 0)
        for (int src = 0; src < p.ag_tp; ++src) ag_wait(src);
      for (int tile = group; tile < num_tiles && ok; tile += ngroups) {
        int m_blk, n_blk, grp;
        if (!tile_coords(p, tile, m_blk, n_blk, grp)) continue;
        const int a_row = (m_blk * CG + (int)rank) * BM;          // this CTA's A rows
        const int b_row = grp * p.grp_n + n_blk * BN + (int)rank * G::kBRows;  // this CTA's share of the B rows
        const int sfb_blk = (grp * p.grp_n) / 128 + n_blk * 2;
        // gathered A: the rows of this m-tile were quantized by rank a_row / ag_rows -- wait (once per source) until they
        // have landed
        if (p.ag_arrived != nullptr && p.ag_rows > 0) ag_wait(a_row / p.ag_rows);
        int seg_off = 0;  // split-K: stages of the segments before this one
        for (int s = 0; s < p.nseg && ok; ++s) {
          const GemmSeg& sg = p.seg[s];
          const uint32_t sf_bytes = (uint32_t)sg.atoms_per_tile * 512u * 3u;  // SFA + two SFB row blocks
          const int kt_lo = SK > 1 ? max(0, sk_lo - seg_off) : 0;
          const int kt_hi = SK > 1 ? min(sg.ktiles, sk_hi - seg_off) : sg.ktiles;
          seg_off += sg.ktiles;
          for (int kt = kt_lo; kt < kt_hi; ++kt) {
            const long long tw0 = WD ? clk() : 0;
            if (!mbar_wait<WD>(empty_bar(stage), phase ^ 1)) {
              if (WD) atomicOr(&p.dbg[0], 0x1u | (uint32_t)(stage << 8) | (uint32_t)(s << 16) | (rank << 24));
              ok = false;
              break;
            }
            if (WD) t_wait += clk() - tw0;
            const uint32_t sbase = smem_base + stage * G::kStageBytes;
            uint32_t bar = full_bar(stage);
            if (rank == 0) mbar_arrive_expect_tx(bar, (sg.tx_a + sg.tx_b + sf_bytes) * CG);
            if constexpr (CG == 2) bar = mapa_rank(bar, 0);  // every byte of the pair is counted on the leader
            tma_load_2d<CG>(sbase, &tmaps.a[s], bar, kt * sg.kelems, a_row);
            tma_load_2d<CG>(sbase + G::kStageA, &tmaps.b[s], bar, kt * sg.kelems, b_row);
            const int ka = kt * sg.atoms_per_tile;
            tma_load_3d<CG>(sbase + G::kStageA + G::kStageB, &tmaps.sfa[s], bar, 0, ka, m_blk * CG + (int)rank);
            tma_load_3d<CG>(sbase + G::kStageA + G::kStageB + G::kStageSFA, &tmaps.sfb[s], bar, 0, ka, sfb_blk);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
      if (WD && blockIdx.x == 0) {
        p.dbg[16] = (uint32_t)((clk() - t_begin) >> 4);  // producer: total cycles / 16
        p.dbg[17] = (uint32_t)(t_wait >> 4);             // producer: cycles waiting for a free slot / 16
      }
    }
  } else if (warp == 1) {
    // ======================================================================== MMA issuer (leader CTA)
    // The whole warp walks the loop (converged control flow, warp-uniform operands); one elected lane issues.
    // This warp is the critical path of the kernel: one stage is only ~512 tensor-pipe cycles (four MMAs), and
    // every instruction this single warp executes per stage is serial latency in front of the next MMA.  Hence:
    // all per-stage values are warp-uniform running registers (uniform datapath, no R2UR), descriptors are a
    // constant high word + a low word advanced by immediates, and the stage body is specialised at compile time
    // for (kind, scale atoms) so that it is straight-line code.
    if (rank == 0) {
      bool ok = true;
      long long t_full = 0, t_tmem = 0, t_begin = WD ? clk() : 0;
      uint32_t n_stage = 0;
      const bool do_cp = !WD || !(p.flags & 1u);
      const bool do_mma = !WD || !(p.flags & 2u);
      // low descriptor words of stage 0; a stage adds kStageBytes >> 4
      const uint32_t a_lo0 = (smem_base >> 4) & 0x3fffu;
      const uint32_t b_lo0 = ((smem_base + G::kStageA) >> 4) & 0x3fffu;
      const uint32_t sf_lo0 = ((smem_base + G::kStageA + G::kStageB) >> 4) & 0x3fffu;
      constexpr uint32_t kStageInc = G::kStageBytes >> 4;
      const uint32_t t_sf0 = tmem_base + kColSF;
      uint32_t stage = 0, soff = 0, phase = 0, slot_off = 0;  // slot_off = kSFSlot * (rotating scale slot)
      uint32_t tphase = 0, tcount = 0;

      // one pipeline stage: KIND 0 = kind::mxf4 (256 K, NATOMS in {1, 2}), KIND 1 = kind::mxf8f6f4 (128 K, one atom)
      auto issue_stage = [&](auto kind_c, auto natoms_c, uint32_t idesc, uint32_t d_acc, bool first_of_tile,
                             bool last_of_tile) {
        constexpr int KIND = decltype(kind_c)::value, NATOMS = decltype(natoms_c)::value;
        constexpr int APT = (KIND == 0) ? 2 : 1;
        const long long tf0 = WD ? clk() : 0;
        if (!mbar_wait<WD>(bar_base + 8u * stage, phase)) {
          if (WD && lane == 0) atomicOr(&p.dbg[1], 0x4u | (stage << 8));
          ok = false;
          return;
        }
        if (WD) {
          t_full += clk() - tf0;
          ++n_stage;
        }
        tc_fence_after();
        if (elect_one()) {
          const uint32_t t_sfa = t_sf0 + slot_off;
          const uint32_t t_sfb = t_sfa + 8;
          const uint32_t sf_lo = sf_lo0 + soff;
          if (do_cp) {
#pragma unroll
            for (int a = 0; a < NATOMS; ++a) {
              tc_cp_sf<CG>(t_sfa + 4 * a, desc_from_lo(sf_lo + 32u * a, kDescHiSF));
              tc_cp_sf<CG>(t_sfb + 8 * a, desc_from_lo(sf_lo + (G::kStageSFA >> 4) + 32u * a, kDescHiSF));  // n-block 0
              tc_cp_sf<CG>(t_sfb + 8 * a + 4,
                           desc_from_lo(sf_lo + (G::kStageSFA >> 4) + 32u * (APT + a), kDescHiSF));          // n-block 1
            }
          }
          const uint32_t a_lo = a_lo0 + soff, b_lo = b_lo0 + soff;
          if (do_mma) {
            if constexpr (KIND == 0) {
#pragma unroll
              for (int j = 0; j < 2 * NATOMS; ++j) {  // K=64 each; two scales per row per MMA: sf id 0 or 2
                constexpr uint32_t kSid[2] = {0u, (2u << 29) | (2u << 4)};
                mma_mxf4<CG>(d_acc, desc_from_lo(a_lo + 2u * j, kDescHiOp), desc_from_lo(b_lo + 2u * j, kDescHiOp),
                             idesc | kSid[j & 1], t_sfa + 4 * (j >> 1), t_sfb + 8 * (j >> 1),
                             (j == 0 && first_of_tile) ? 0u : 1u);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {  // K=32 each; sf id = k-block within the 128-K atom
                mma_mxf8f6f4<CG>(d_acc, desc_from_lo(a_lo + 2u * j, kDescHiOp), desc_from_lo(b_lo + 2u * j, kDescHiOp),
                                 idesc | ((uint32_t)j << 29) | ((uint32_t)j << 4), t_sfa, t_sfb,
                                 (j == 0 && first_of_tile) ? 0u : 1u);
              }
            }
          }
          tc_commit<CG>(bar_base + 8u * (kStages + stage));  // frees the smem slot (both CTAs) once it has been read
          if (last_of_tile) tc_commit<CG>(tmem_full_bar);
        }
        __syncwarp();
        ++stage;
        soff += kStageInc;
        if (stage == (uint32_t)kStages) {
          stage = 0;
          soff = 0;
          phase ^= 1;
        }
        slot_off += kSFSlot;
        if (slot_off == kSFSlot * kNumSFSlots) slot_off = 0;
      };
      using std::integral_constant;
      for (int tile = group; tile < num_tiles && ok; tile += ngroups) {
        {
          int m_blk, n_blk, grp;
          if (!tile_coords(p, tile, m_blk, n_blk, grp)) continue;
        }
        // the epilogue must have drained the columns this tile's accumulator shares with the previous one
        const long long tt0 = WD ? clk() : 0;
        if (!mbar_wait<WD>(tmem_empty_bar, tphase ^ 1)) {
          if (WD && lane == 0) atomicOr(&p.dbg[1], 0x2u);
          ok = false;
          break;
        }
        if (WD) t_tmem += clk() - tt0;
        tc_fence_after();
        const uint32_t d_acc = tmem_base + ((tcount & 1u) ? kColAcc1 : 0u);
        if constexpr (SK > 1) {
          // this CTA's slice of the stage list; per-stage dispatch is fine here (the kernel streams weights)
          int seg_off = 0, issued = 0;
          const int mine = sk_hi - sk_lo;
          for (int sidx = 0; sidx < p.nseg && ok; ++sidx) {
            const GemmSeg& sg = p.seg[sidx];
            const int kt_lo = max(0, sk_lo - seg_off), kt_hi = min(sg.ktiles, sk_hi - seg_off);
            seg_off += sg.ktiles;
            for (int kt = kt_lo; kt < kt_hi && ok; ++kt, ++issued) {
              const bool first = issued == 0, last = issued == mine - 1;
              if (sg.kind == 0) {
                if (kt == sg.ktiles - 1 && sg.last_atoms == 1)
                  issue_stage(integral_constant<int, 0>{}, integral_constant<int, 1>{}, sg.idesc, d_acc, first, last);
                else
                  issue_stage(integral_constant<int, 0>{}, integral_constant<int, 2>{}, sg.idesc, d_acc, first, last);
              } else {
                issue_stage(integral_constant<int, 1>{}, integral_constant<int, 1>{}, sg.idesc, d_acc, first, last);
              }
            }
          }
        } else
        for (int sidx = 0; sidx < p.nseg && ok; ++sidx) {
          const GemmSeg& sg = p.seg[sidx];
          const int ktiles = sg.ktiles;
          const uint32_t idesc = sg.idesc;
          const bool first_seg = sidx == 0, last_seg = sidx == p.nseg - 1;
          if (sg.kind == 0) {
            for (int kt = 0; kt < ktiles - 1 && ok; ++kt)
              issue_stage(integral_constant<int, 0>{}, integral_constant<int, 2>{}, idesc, d_acc, first_seg && kt == 0,
                          false);
            if (!ok) break;
            if (sg.last_atoms == 2)
              issue_stage(integral_constant<int, 0>{}, integral_constant<int, 2>{}, idesc, d_acc,
                          first_seg && ktiles == 1, last_seg);
            else
              issue_stage(integral_constant<int, 0>{}, integral_constant<int, 1>{}, idesc, d_acc,
                          first_seg && ktiles == 1, last_seg);
          } else {
            for (int kt = 0; kt < ktiles - 1 && ok; ++kt)
              issue_stage(integral_constant<int, 1>{}, integral_constant<int, 1>{}, idesc, d_acc, first_seg && kt == 0,
                          false);
            if (!ok) break;
            issue_stage(integral_constant<int, 1>{}, integral_constant<int, 1>{}, idesc, d_acc, first_seg && ktiles == 1,
                        last_seg);
          }
        }
        tphase ^= 1;
        ++tcount;
      }
      if (WD && blockIdx.x == 0 && lane == 0) {
        p.dbg[18] = (uint32_t)((clk() - t_begin) >> 4);  // MMA warp: total cycles / 16
        p.dbg[19] = (uint32_t)(t_full >> 4);             // ... blocked on TMA data / 16
        p.dbg[20] = (uint32_t)(t_tmem >> 4);             // ... waiting for the epilogue to hand TMEM back / 16
        p.dbg[21] = n_stage;
      }
    }
  } else {
    // ======================================================================== epilogue warps 2..5 (every CTA)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    uint32_t tphase = 0, tcount = 0;
    long long t_ewait = 0, t_ework = 0;
    uint32_t n_tiles_done = 0;
    const uint32_t sbuf = smem_base + G::kEpiOff + (uint32_t)(warp - 2) * kEpiStage;
    uint32_t nstore = 0;  // TMA stores issued by this warp: picks the staging buffer
    uint32_t* prev_flag = nullptr;  // RS: arrival counter of the previous tile, bumped one tile late (lane 0)
    if (RS && blockIdx.x == 0 && warp == 2 && lane == 0) {  // timeline probe (tools/tp_fused_probe.py)
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[40] = (uint32_t)t;
      p.dbg[41] = (uint32_t)(t >> 32);
    }
    for (int tile = group; tile < num_tiles; tile += ngroups) {
      int m_blk, n_blk, grp;
      if (!tile_coords(p, tile, m_blk, n_blk, grp)) continue;
      const long long te0 = WD ? clk() : 0;
      if (!mbar_wait<WD>(tmem_full_bar, tphase)) {
        if (WD && lane == 0) atomicOr(&p.dbg[2], 0x8u | (uint32_t)(q << 8) | (rank << 24));
        break;
      }
      const long long te1 = WD ? clk() : 0;
      tc_fence_after();
      const int row0 = (m_blk * CG + (int)rank) * BM + q * 32;  // first C row of this warp's 32-row band
      const bool odd = (tcount & 1u) != 0;
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16) + (odd ? kColAcc1 : 0u);
      if constexpr (SK > 1) {
        // park the fp32 partial accumulator in the stage buffers (every MMA has read them: tmem_full follows the last
        // commit): row r = 1 KB, 16-byte chunk c at (c ^ (r & 7)) -- conflict-free for the row-per-lane writes here
        // and for the chunk-per-lane reads of the reduction
        const uint32_t srow = smem_base + (uint32_t)(q * 32 + lane) * 1024u;
        const uint32_t sw = (uint32_t)lane & 7u;
        const bool row_live = row0 + lane < p.M;  // decode: most of the 128 rows do not exist and are never reduced
        if (row0 < p.M) {                         // warp-uniform (tcgen05.ld is warp-collective)
#pragma unroll 1
          for (int i = 0; i < BN / 32; ++i) {
            if (n_blk * BN + i * 32 >= p.N) break;
            uint32_t r[32];
            tmem_ld32(tbase + (uint32_t)(i * 32), r);
            tmem_ld_wait();
            if (row_live) {
#pragma unroll
              for (int v = 0; v < 8; ++v)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(srow + ((((uint32_t)(i * 8 + v)) ^ sw) << 4)),
                             "r"(r[4 * v]), "r"(r[4 * v + 1]), "r"(r[4 * v + 2]), "r"(r[4 * v + 3])
                             : "memory");
            }
          }
        }
        tc_fence_before();
        break;  // one tile per cluster; the reduction follows the role branches
      }
      if constexpr (ACT) {
        // ---- fused SiLU(gate) * up + MX quantize: this lane owns row row0 + lane; 32-column chunk i < 4 holds the gate,
        // chunk i + 4 the up accumulators of channels [128 n_blk + 32 i, + 32) -- one scale group per (lane, i)
        constexpr int kShared = (int)kAccOverlap / 32;
        static_assert(kShared <= 4, "the shared columns lie inside one half of the tile");
        const int ch0 = n_blk * 128;
        const int sg = (ch0 >= p.act_cend[1]) ? 2 : (ch0 >= p.act_cend[0] ? 1 : 0);
        const int cb = (sg == 0) ? 0 : p.act_cend[sg - 1];
        const int fmt = 4 + 2 * sg;
        const float qmax = (sg == 0) ? 6.0f : (sg == 1 ? 28.0f : 448.0f);
        const bool row_ok = row0 + lane < p.M;
        uint8_t* qrow = p.act_q[sg] + (int64_t)(row0 + lane) * p.act_rowbytes[sg] + (((ch0 - cb) * fmt) >> 3);
        uint32_t sfword = 0;
        // one scale group: gw / uw = the bf16-rounded gate / up accumulators of 32 channels, two per word
        auto act_group = [&](const uint32_t (&gw)[16], const uint32_t (&uw)[16], int pr) {
          // fast_silu on all 32 channels as straight-line code (this warp is alone on its scheduler: only instruction-
          // level parallelism hides the ~100-cycle chain of one element); the rare group with a gate value outside the
          // range fast_silu is exact on takes the reference sequence instead
          uint32_t lo2 = 0xffffffffu, hi2 = 0u;
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const uint32_t ab = gw[e] & 0x7fff7fffu;
            lo2 = __vminu2(lo2, ab);
            hi2 = __vmaxu2(hi2, ab);
          }
          float v[32];
          if (min(lo2 & 0xffffu, lo2 >> 16) >= kFastSiluLo && max(hi2 & 0xffffu, hi2 >> 16) <= kFastSiluHi) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              v[2 * e] = __fmul_rn(fast_silu(__uint_as_float(gw[e] << 16)), __uint_as_float(uw[e] << 16));
              v[2 * e + 1] = __fmul_rn(fast_silu(__uint_as_float(gw[e] & 0xffff0000u)), __uint_as_float(uw[e] & 0xffff0000u));
            }
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              v[2 * e] = __fmul_rn(ref_silu(__uint_as_float(gw[e] << 16)), __uint_as_float(uw[e] << 16));
              v[2 * e + 1] = __fmul_rn(ref_silu(__uint_as_float(gw[e] & 0xffff0000u)), __uint_as_float(uw[e] & 0xffff0000u));
            }
          }
          float m = 0.0f;
#pragma unroll
          for (int e = 0; e < 32; ++e) m = fmaxf(m, fabsf(v[e]));
          int n = 0;
          if (m > 1e-6f) n = ref_scale_exp(m, qmax);
          const float rsc = __uint_as_float((uint32_t)(127 - n) << 23);  // 2^-n, exact multiplier
          sfword |= (uint32_t)(n + 127) << (8 * pr);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __fmul_rn(v[e], rsc);
          if (row_ok) {
            if (sg == 0) {
              uint32_t w[4];
#pragma unroll
              for (int k = 0; k < 4; ++k)
                w[k] = rq_cvt4_e2m1(v[8 * k], v[8 * k + 1], v[8 * k + 2], v[8 * k + 3]) |
                       (rq_cvt4_e2m1(v[8 * k + 4], v[8 * k + 5], v[8 * k + 6], v[8 * k + 7]) << 16);
              *reinterpret_cast<uint4*>(qrow + pr * 16) = make_uint4(w[0], w[1], w[2], w[3]);
            } else if (sg == 1) {
              uint32_t y[8];  // 24 bits each
#pragma unroll
              for (int k = 0; k < 8; ++k)
                y[k] = rq_squeeze4_fp6(rq_cvt4_e3m2(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
#pragma unroll
              for (int k = 0; k < 2; ++k) {  // four 24-bit words -> three 32-bit words
                const uint32_t a0 = y[4 * k] | (y[4 * k + 1] << 24);
                const uint32_t a1 = (y[4 * k + 1] >> 8) | (y[4 * k + 2] << 16);
                const uint32_t a2 = (y[4 * k + 2] >> 16) | (y[4 * k + 3] << 8);
                uint32_t* d = reinterpret_cast<uint32_t*>(qrow + pr * 24 + k * 12);
                d[0] = a0;
                d[1] = a1;
                d[2] = a2;
              }
            } else {
              uint32_t w[8];
#pragma unroll
              for (int k = 0; k < 8; ++k) w[k] = rq_cvt4_e4m3(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
              *reinterpret_cast<uint4*>(qrow + pr * 32) = make_uint4(w[0], w[1], w[2], w[3]);
              *reinterpret_cast<uint4*>(qrow + pr * 32 + 16) = make_uint4(w[4], w[5], w[6], w[7]);
            }
          }
        };
        auto pack32 = [&](const uint32_t (&r)[32], uint32_t (&o)[16]) {
#pragma unroll
          for (int e = 0; e < 16; ++e) o[e] = pack_bf16(__uint_as_float(r[2 * e]), __uint_as_float(r[2 * e + 1]));
        };
        // the columns shared with the other accumulator leave TMEM first (the top chunks of acc0 = up chunks, the bottom
        // chunks of acc1 = gate chunks), already rounded to bf16; then the MMA warp may start the next tile
        uint32_t sh[kShared][16];
        {
          uint32_t raw[32];
#pragma unroll
          for (int i = 0; i < kShared; ++i) {
            tmem_ld32(tbase + (uint32_t)((odd ? i : 7 - i) * 32), raw);
            tmem_ld_wait();
            pack32(raw, sh[i]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(tmem_empty_bar, 0); else mbar_arrive(tmem_empty_bar);
        }
        // pairs in the order that uses the pre-loaded chunks first (acc0: pairs 3, 2, 1, 0 -- up chunk 7 - k is sh[k];
        // acc1: pairs 0, 1, 2, 3 -- gate chunk k is sh[k]); one copy of the group code, the loop is not unrolled
        uint32_t gw[16], uw[16];
#pragma unroll 1
        for (int k = 0; k < 4; ++k) {
          const int pr = odd ? k : 3 - k;
          const bool pre = k < kShared;
          uint32_t raw[32];
          if (!(pre && odd)) {
            tmem_ld32(tbase + (uint32_t)(pr * 32), raw);
            tmem_ld_wait();
            pack32(raw, gw);
          }
          if (!(pre && !odd)) {
            tmem_ld32(tbase + (uint32_t)((pr + 4) * 32), raw);
            tmem_ld_wait();
            pack32(raw, uw);
          }
          if (pre) {
#pragma unroll
            for (int j = 0; j < kShared; ++j) {
              if (k == j) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                  if (odd) gw[i] = sh[j][i];
                  else uw[i] = sh[j][i];
                }
              }
            }
          }
          act_group(gw, uw, pr);
        }
        // the row block's scale bytes of this tile's atom: four groups of row (lane, q) are one aligned 32-bit word
        const int rowblk = m_blk * CG + (int)rank;
        if ((int64_t)rowblk * BM < p.M) {
          uint8_t* d = p.act_sf[sg] + ((int64_t)rowblk * p.act_katoms[sg] + ((ch0 - cb) >> 7)) * 512 + lane * 16 + q * 4;
          *reinterpret_cast<uint32_t*>(d) = sfword;
        }
        tc_fence_before();
        tphase ^= 1;
        ++tcount;
        continue;
      }
      const bool do_store = row0 < p.M && !(WD && (p.flags & 4u));
      const uint32_t nstore_tile0 = nstore;
      // where this warp's 32-row band goes: C itself, or (RS) the owner rank's staging tile, tile-local coordinates
      const CUtensorMap* cmap = &tmaps.c;
      int st_row0 = row0, st_col_base = n_blk * BN;
      int owner = 0, own_idx = 0;
      if constexpr (RS) {
        // blocks of tp consecutive tiles go to the tp ranks, rotated by one every rot_s tiles: a CTA group's tiles
        // (stride = #groups) then go to DIFFERENT owners in turn instead of always the same one, so every SM pushes its
        // share over NVLink rather than half of them pushing everything (tp = 2, even #groups)
        own_idx = tile / rs.tp;
        owner = (tile + (own_idx * rs.tp) / rs.rot_s) % rs.tp;
        if (!rs.pull) {
          cmap = &rs.dst[(p.flags & 8u) ? rs.rank : owner];  // flag 8: timing experiment, partials stay local
          st_row0 = own_idx * (CG * BM) + (int)rank * BM + q * 32;
          st_col_base = 0;
        }
      }
      // 32-column chunks, the columns shared with the other accumulator first: the top ones of acc0, the bottom
      // ones of acc1.  Once those are in registers the MMA warp may start the next tile.
      constexpr int kChunks = BN / 32, kShared = (int)kAccOverlap / 32;
      auto chunk_col = [&](int i) { return odd ? i : (kChunks - 1 - i); };
      // residual rows: this lane's 64 bytes of chunk number i (in emit order) travel two chunks ahead of their use
      auto res_issue = [&](int i, uint4 (&dst)[4]) {
        const int col0 = n_blk * BN + chunk_col(i) * 32;
        if (row0 + lane < p.M && col0 < p.N) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.residual + (int64_t)(row0 + lane) * p.N + col0);
#pragma unroll
          for (int v = 0; v < 4; ++v) dst[v] = __ldg(rp + v);
        }
      };
      uint4 rres0[4] = {}, rres1[4] = {};
      const bool has_res = !RS && !ROPE && p.residual != nullptr;
      if (has_res) {
        res_issue(0, rres0);
        res_issue(1, rres1);
      }
      auto emit = [&](const uint32_t (&r)[32], int sc, uint4 (&res)[4], int inext) {  // 32 rows x 32 columns: bf16, staged, TMA-stored
        const int col0 = n_blk * BN + sc * 32;
        if (col0 >= p.N) {  // warp-uniform: a column block past N (last tile of an N % 256 == 128 matrix)
          // the residual ring still advances: chunk inext travels in the slot this (skipped) chunk would have freed
          if (has_res && inext < kChunks) res_issue(inext, res);
          return;
        }
        {
          uint32_t o[16];
          pack_chunk(r, o, p.bias ? p.bias + col0 : nullptr);
          if constexpr (ROPE) {
            if (col0 < p.rope_cols) {  // warp-uniform: a q / k chunk -- word e = (channel j0 + e, channel j0 + e + 64)
              const int64_t row = row0 + lane;
              if (row < p.M) {
                const int hcol = col0 & ~127, j0 = (col0 & 127) >> 1;  // head's first column; first channel of the chunk
                const int64_t toff = (row % p.rope_S) * 128 + j0;
                const uint4* cl = reinterpret_cast<const uint4*>(p.rope_cos + toff);
                const uint4* ch = reinterpret_cast<const uint4*>(p.rope_cos + toff + 64);
                const uint4* sl = reinterpret_cast<const uint4*>(p.rope_sin + toff);
                const uint4* sh = reinterpret_cast<const uint4*>(p.rope_sin + toff + 64);
                uint32_t lo[8], hi[8];
#pragma unroll
                for (int v = 0; v < 2; ++v) {
                  const uint4 c1 = __ldg(cl + v), c2 = __ldg(ch + v), s1 = __ldg(sl + v), s2 = __ldg(sh + v);
                  const uint32_t k1[4] = {c1.x, c1.y, c1.z, c1.w}, k2[4] = {c2.x, c2.y, c2.z, c2.w};
                  const uint32_t n1[4] = {s1.x, s1.y, s1.z, s1.w}, n2[4] = {s2.x, s2.y, s2.z, s2.w};
#pragma unroll
                  for (int u = 0; u < 4; ++u) {  // table word u: channels 2u, 2u + 1 of this half-run (both halves)
                    uint32_t out[2];
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                      const uint32_t w = o[8 * v + 2 * u + b];                       // (x1, x2)
                      const uint32_t cw = __byte_perm(k1[u], k2[u], b ? 0x7632 : 0x5410);  // (cos[j], cos[j + 64])
                      const uint32_t sw = __byte_perm(n1[u], n2[u], b ? 0x7632 : 0x5410);
                      const uint32_t rot = __byte_perm(w, w, 0x1032) ^ 0x00008000u;  // (-x2, x1)
                      const __nv_bfloat162 t1 = __hmul2_rn(*reinterpret_cast<const __nv_bfloat162*>(&w),
                                                           *reinterpret_cast<const __nv_bfloat162*>(&cw));
                      const __nv_bfloat162 t2 = __hmul2_rn(*reinterpret_cast<const __nv_bfloat162*>(&rot),
                                                           *reinterpret_cast<const __nv_bfloat162*>(&sw));
                      const __nv_bfloat162 sum = __hadd2_rn(t1, t2);
                      out[b] = *reinterpret_cast<const uint32_t*>(&sum);
                    }
                    lo[4 * v + u] = __byte_perm(out[0], out[1], 0x5410);  // channels j, j + 1
                    hi[4 * v + u] = __byte_perm(out[0], out[1], 0x7632);  // channels j + 64, j + 65
                  }
                }
                __nv_bfloat16* d = p.c + row * p.N + hcol + j0;
                *reinterpret_cast<uint4*>(d) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<uint4*>(d + 8) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
                *reinterpret_cast<uint4*>(d + 64) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(d + 72) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
              }
              return;
            }
          }
          if (has_res) {
            add_residual(o, res);
            if (inext < kChunks) res_issue(inext, res);
          }
          const uint32_t buf = sbuf + (nstore & 1u) * kEpiBuf;
          // the store issued from this buffer two chunks ago must have finished READING it
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          stage_chunk(o, buf, lane);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to TMA
          __syncwarp();
          if (lane == 0 && do_store) {
            // C: clipped at M and N by the map.  RS: the staging slot is BOX-MAJOR -- every 32 x 32 box is 2 KB of
            // contiguous peer memory (64-byte rows of a row-major tile travel over NVLink at less than half the rate)
            bool box_major = false;
            if constexpr (RS) box_major = !rs.pull;
            if (box_major) tma_store_2d(cmap, buf, 0, st_row0 * 8 + sc * 32);
            else tma_store_2d(cmap, buf, st_col_base + sc * 32, st_row0);
          }
          ++nstore;
        }
      };
      // Software pipeline: the TMEM load of chunk i+1 is in flight while chunk i is packed, staged and handed to TMA
      // (a chunk is latency-, not bandwidth-bound: ld -> wait -> 16 cvt -> 4 st.shared -> fence -> TMA issue).
      uint32_t ra[32], rb[32];
      {
        uint32_t rs[kShared][32];
#pragma unroll
        for (int i = 0; i < kShared; ++i) tmem_ld32(tbase + (uint32_t)(chunk_col(i) * 32), rs[i]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if constexpr (CG == 2) mbar_arrive_cluster(tmem_empty_bar, 0); else mbar_arrive(tmem_empty_bar);
        }
        tmem_ld32(tbase + (uint32_t)(chunk_col(kShared) * 32), ra);  // first private chunk: under the shared chunks' emit
#pragma unroll
        for (int i = 0; i < kShared; ++i) {
          if (i & 1) emit(rs[i], chunk_col(i), rres1, i + 2);
          else emit(rs[i], chunk_col(i), rres0, i + 2);
        }
      }
#pragma unroll
      for (int i = kShared; i < kChunks; i += 2) {
        tmem_ld_wait();  // ra = chunk i
        if (i + 1 < kChunks) tmem_ld32(tbase + (uint32_t)(chunk_col(i + 1) * 32), rb);
        if (i & 1) emit(ra, chunk_col(i), rres1, i + 2);
        else emit(ra, chunk_col(i), rres0, i + 2);
        if (i + 1 < kChunks) {
          tmem_ld_wait();  // rb = chunk i + 1
          if (i + 2 < kChunks) tmem_ld32(tbase + (uint32_t)(chunk_col(i + 2) * 32), ra);
          if ((i + 1) & 1) emit(rb, chunk_col(i + 1), rres1, i + 3);
          else emit(rb, chunk_col(i + 1), rres0, i + 3);
        }
      }
      if constexpr (RS) {
        // DEFERRED arrival: once this tile's stores are issued, the stores of the PREVIOUS tile have long LANDED in its
        // owner's memory (wait_group N without .read = all but the N most recent groups are complete), so the wait costs
        // no NVLink round trip; then ONE system-scope release on the owner's counter (4 * CG arrivals per source rank
        // complete a tile).  The last tile's arrival follows the loop.
        if (lane == 0) {
          // stores this warp issued for this tile: 8, or 4 on a 128-wide last column block, none for rows past M
          const int cnt = do_store ? (int)(nstore - nstore_tile0) : 0;
          if (prev_flag != nullptr) {
            if (cnt == 8) asm volatile("cp.async.bulk.wait_group 8;" ::: "memory");
            else if (cnt == 4) asm volatile("cp.async.bulk.wait_group 4;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.global;" ::: "memory");
            if (p.flags & 32u) {  // timing experiment: no arrival at all
            } else if (p.flags & 16u) {
              asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(prev_flag) : "memory");
            } else if (rs.pull && !(p.flags & 64u)) {
              // in-switch mode: the partial tile stays in THIS GPU's memory and the owner reads it through NVLink, i.e. out
              // of this GPU's L2 -- a gpu-scope fence makes the (completed) TMA stores visible there; a system-scope release
              // would also wait for this warp's previous arrival to be acknowledged over NVLink, one round trip per tile
              asm volatile("fence.acq_rel.gpu;" ::: "memory");
              asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(prev_flag) : "memory");
            } else {
              asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(prev_flag) : "memory");
            }
          }
          prev_flag = rs.tile_flags[owner] + own_idx;
        }
      }
      tphase ^= 1;
      ++tcount;
      if (WD) {
        t_ewait += te1 - te0;
        t_ework += clk() - te1;
        ++n_tiles_done;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores done before smem goes away
    if (RS && blockIdx.x == 0 && warp == 2 && lane == 0) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      p.dbg[42] = (uint32_t)t;
      p.dbg[43] = (uint32_t)(t >> 32);
    }
    if constexpr (RS) {
      if (lane == 0 && prev_flag != nullptr) {
        asm volatile("fence.proxy.async.global;" ::: "memory");
        if (p.flags & 32u) {
        } else if (rs.pull && !(p.flags & 64u)) {
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
          asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(prev_flag) : "memory");
        } else {
          asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(prev_flag) : "memory");
        }
      }
    }
    if (WD && blockIdx.x == 0 && warp == 2 && lane == 0) {
      p.dbg[22] = (uint32_t)(t_ewait >> 4);  // epilogue warp: cycles waiting for an accumulator / 16
      p.dbg[23] = (uint32_t)(t_ework >> 4);  // ... draining TMEM and storing C / 16
      p.dbg[24] = n_tiles_done;
    }
  }

  if constexpr (SK > 1) {
    cluster_sync_all();  // every CTA of the cluster has parked its partial tile
    if (warp >= 2 && group < num_tiles) {
      constexpr int CH = (BN / SK) / 4;  // 16-byte fp32 chunks per row of this CTA's column slice
      const int m_blk = p.n_fastest ? group / p.n_tiles : group % p.m_tiles;
      const int n_blk = p.n_fastest ? group % p.n_tiles : group / p.m_tiles;
      const int live_rows = (int)min((int64_t)BM, p.M - (int64_t)m_blk * BM);
      for (int idx = (int)threadIdx.x - 64; idx < live_rows * CH; idx += kThreads - 64) {
        const int row = idx / CH;
        const int c = (int)krank * CH + idx % CH;  // chunk index inside the 256-column row
        const int64_t grow = (int64_t)m_blk * BM + row;
        const int gcol = n_blk * BN + c * 4;
        if (grow >= p.M || gcol >= p.N) continue;
        const uint32_t local = smem_base + (uint32_t)row * 1024u + ((((uint32_t)c) ^ ((uint32_t)row & 7u)) << 4);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int sidx = 0; sidx < SK; ++sidx) {  // CTA order: the sum does not depend on timing
          float x0, x1, x2, x3;
          asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
                       : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3)
                       : "r"(mapa_rank(local, (uint32_t)sidx)));
          a0 += x0;
          a1 += x1;
          a2 += x2;
          a3 += x3;
        }
        uint32_t w0 = pack_bf16(a0, a1), w1 = pack_bf16(a2, a3);
        if (p.bias != nullptr) {  // y = bf16(acc); y = bf16(y + bias), as in pack_chunk
          const uint2 bv = __ldg(reinterpret_cast<const uint2*>(p.bias + gcol));
          w0 = pack_bf16(__uint_as_float(w0 << 16) + __uint_as_float(bv.x << 16),
                         __uint_as_float(w0 & 0xffff0000u) + __uint_as_float(bv.x & 0xffff0000u));
          w1 = pack_bf16(__uint_as_float(w1 << 16) + __uint_as_float(bv.y << 16),
                         __uint_as_float(w1 & 0xffff0000u) + __uint_as_float(bv.y & 0xffff0000u));
        }
        if (p.residual != nullptr) {
          const uint2 rv = __ldg(reinterpret_cast<const uint2*>(p.residual + grow * p.N + gcol));
          w0 = add_bf16x2(w0, rv.x);
          w1 = add_bf16x2(w1, rv.y);
        }
        if (p.rope_cos != nullptr && gcol < p.rope_cols) {
          // pair-adjacent q / k columns (see GemmParams::rope_cos): w0 = (x1[j], x2[j]), w1 = (x1[j + 1], x2[j + 1])
          const int hcol = gcol & ~127, j = (gcol & 127) >> 1;
          const int64_t toff = (grow % p.rope_S) * 128 + j;
          const uint32_t c1 = __ldg(reinterpret_cast<const uint32_t*>(p.rope_cos + toff));
          const uint32_t c2 = __ldg(reinterpret_cast<const uint32_t*>(p.rope_cos + toff + 64));
          const uint32_t s1 = __ldg(reinterpret_cast<const uint32_t*>(p.rope_sin + toff));
          const uint32_t s2 = __ldg(reinterpret_cast<const uint32_t*>(p.rope_sin + toff + 64));
          uint32_t out[2];
#pragma unroll
          for (int b = 0; b < 2; ++b) {
            const uint32_t w = b ? w1 : w0;
            const uint32_t cw = __byte_perm(c1, c2, b ? 0x7632 : 0x5410), sw = __byte_perm(s1, s2, b ? 0x7632 : 0x5410);
            const uint32_t rot = __byte_perm(w, w, 0x1032) ^ 0x00008000u;
            const __nv_bfloat162 t1 = __hmul2_rn(*reinterpret_cast<const __nv_bfloat162*>(&w),
                                                 *reinterpret_cast<const __nv_bfloat162*>(&cw));
            const __nv_bfloat162 t2 = __hmul2_rn(*reinterpret_cast<const __nv_bfloat162*>(&rot),
                                                 *reinterpret_cast<const __nv_bfloat162*>(&sw));
            const __nv_bfloat162 sum = __hadd2_rn(t1, t2);
            out[b] = *reinterpret_cast<const uint32_t*>(&sum);
          }
          __nv_bfloat16* d = p.c + grow * p.N + hcol + j;
          *reinterpret_cast<uint32_t*>(d) = __byte_perm(out[0], out[1], 0x5410);
          *reinterpret_cast<uint32_t*>(d + 64) = __byte_perm(out[0], out[1], 0x7632);
          continue;
        }
        *reinterpret_cast<uint2*>(p.c + grow * p.N + gcol) = make_uint2(w0, w1);
      }
    }
    cluster_sync_all();  // nobody leaves while a peer may still read its shared memory
    if (warp == 1) {
      tc_fence_after();
      tmem_dealloc<CG>(tmem_base, kTmemCols);
    }
    return;
  }
  tc_fence_before();
  if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<CG>(tmem_base, kTmemCols);
  }
  if (p.ag_arrived != nullptr && threadIdx.x == 0) {
    // every TMA load of this CTA has completed (its full-barriers were consumed); the last CTA releases the gather
    __threadfence();
    if (atomicInc(p.ag_ticket, gridDim.x - 1) == gridDim.x - 1) {
      uint32_t taken;
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(taken) : "l"(p.ag_taken) : "memory");
      asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(p.ag_taken), "r"(taken + 1u) : "memory");
      __threadfence_system();
      for (int d = 0; d < p.ag_tp; ++d)
        asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(p.ag_consumed[d]) : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  int64_t rows;
  int kseg;
  int dtype;
  int box_rows;
  int box_k;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && kseg == o.kseg && dtype == o.dtype && box_rows == o.box_rows &&
           box_k == o.box_k;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    h ^= std::hash<int64_t>()(k.rows * 1000003 + k.kseg) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    h ^= std::hash<int>()(k.dtype * 31 + k.box_rows * 7 + k.box_k) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    return h;
  }
};

// operand tile map: 2-D tensor [rows, Kseg] of sub-byte / byte codes, K fastest, box = box_k x box_rows, swizzle 128B
static int get_tmap(const void* ptr, int64_t rows, int kseg, int bits, bool unpack, int box_rows, CUtensorMap* out) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  CUtensorMapDataType dt;
  int box_k;
  if (bits == 4 && !unpack) { dt = CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN8B; box_k = 256; }
  else if (bits == 4) { dt = CU_TENSOR_MAP_DATA_TYPE_16U4_ALIGN16B; box_k = 128; }
  else if (bits == 6) { dt = CU_TENSOR_MAP_DATA_TYPE_16U6_ALIGN16B; box_k = 128; }
  else { dt = CU_TENSOR_MAP_DATA_TYPE_UINT8; box_k = 128; }
  TmapKey key{ptr, rows, kseg, (int)dt, box_rows, box_k};
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) {
      *out = it->second;
      return MMX_OK;
    }
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return MMX_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)kseg, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)kseg * bits / 8};
  const cuuint32_t box[2] = {(cuuint32_t)box_k, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) ptr=%p rows=%lld kseg=%d bits=%d unpack=%d box=%dx%d", (int)r, ptr,
              (long long)rows, kseg, bits, (int)unpack, box_k, box_rows);
    return MMX_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    if (cache.size() > 8192) cache.clear();
    cache.emplace(key, *out);
  }
  return MMX_OK;
}

// Scale-factor and output maps are cached like the operand maps (an encode costs ~1 us of host time per map and a
// decode-sized GEMM is only a few us): key = (kind, pointer, three shape words).
struct AuxKey {
  int kind;
  const void* ptr;
  int64_t a, b, c;
  bool operator==(const AuxKey& o) const { return kind == o.kind && ptr == o.ptr && a == o.a && b == o.b && c == o.c; }
};
struct AuxKeyHash {
  size_t operator()(const AuxKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr) ^ (size_t)k.kind * 0x9e3779b97f4a7c15ULL;
    h ^= std::hash<int64_t>()(k.a * 1000003 + k.b * 8191 + k.c) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    return h;
  }
};
static std::unordered_map<AuxKey, CUtensorMap, AuxKeyHash>& aux_cache() {
  static std::unordered_map<AuxKey, CUtensorMap, AuxKeyHash> c;
  return c;
}
static std::mutex g_aux_mu;
static bool aux_lookup(const AuxKey& key, CUtensorMap* out) {
  std::lock_guard<std::mutex> lk(g_aux_mu);
  auto it = aux_cache().find(key);
  if (it == aux_cache().end()) return false;
  *out = it->second;
  return true;
}
static void aux_store(const AuxKey& key, const CUtensorMap& m) {
  std::lock_guard<std::mutex> lk(g_aux_mu);
  if (aux_cache().size() > 8192) aux_cache().clear();
  aux_cache().emplace(key, m);
}

// scale-factor map: the 512-byte atoms of one SF buffer viewed as u32[rblocks][katoms][128]; box = nrb x atoms x 128
static int get_sf_tmap(const void* ptr, int64_t rblocks, int katoms, int box_atoms, int box_rblocks, CUtensorMap* out) {
  const AuxKey key{1, ptr, rblocks, katoms, box_atoms * 16 + box_rblocks};
  if (aux_lookup(key, out)) return MMX_OK;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return MMX_ERR_CUDA;
  }
  const cuuint64_t gdim[3] = {128, (cuuint64_t)katoms, (cuuint64_t)rblocks};
  const cuuint64_t gstride[2] = {512, (cuuint64_t)512 * katoms};
  const cuuint32_t box[3] = {128, (cuuint32_t)box_atoms, (cuuint32_t)box_rblocks};
  const cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (scale factors) failed (%d) ptr=%p rblocks=%lld katoms=%d", (int)r, ptr,
              (long long)rblocks, katoms);
    return MMX_ERR_CUDA;
  }
  aux_store(key, *out);
  return MMX_OK;
}

// output map: bf16 C[M, N] row-major, box = 32 columns x 32 rows, 64B swizzle (matches stage_chunk's layout)
static int get_c_tmap(void* ptr, int64_t M, int64_t N, CUtensorMap* out) {
  const AuxKey key{2, ptr, M, N, 0};
  if (aux_lookup(key, out)) return MMX_OK;
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled is not available from this driver");
    return MMX_ERR_CUDA;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)M};
  const cuuint64_t gstride[1] = {(cuuint64_t)N * 2};
  const cuuint32_t box[2] = {32, 32};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (C) failed (%d) ptr=%p M=%lld N=%lld", (int)r, ptr, (long long)M, (long long)N);
    return MMX_ERR_CUDA;
  }
  aux_store(key, *out);
  return MMX_OK;
}

static uint32_t make_idesc(int kind, int a_bits, int b_bits, int mma_m) {
  auto fmt_f8f6f4 = [](int bits) -> uint32_t { return bits == 8 ? 0u : (bits == 6 ? 4u : 5u); };  // E4M3, E3M2, E2M1
  const uint32_t afmt = kind == 0 ? 1u : fmt_f8f6f4(a_bits);  // MXF4Format::E2M1 = 1
  const uint32_t bfmt = kind == 0 ? 1u : fmt_f8f6f4(b_bits);
  uint32_t d = 0;
  d |= afmt << 7;
  d |= bfmt << 10;
  // bits 13,14 negate = 0; bits 15,16 major = 0 (K-major)
  d |= (uint32_t)(BN >> 3) << 17;
  d |= 1u << 23;  // scale format UE8M0
  d |= (uint32_t)(mma_m >> 4) << 24;
  return d;
}

template <int CG>
static int launch_gemm(const TmapSet& tm, GemmParams& p, cudaStream_t st, const RsParams* rs, bool act = false,
                       bool rope = false) {
  const int smem_bytes = rs != nullptr ? Geo<CG, true>::kSmemBytes : Geo<CG, false>::kSmemBytes;
  p.m_tiles = (int)((p.M + BM * CG - 1) / (BM * CG));
  p.n_tiles = (int)((p.N + BN - 1) / BN);
  int64_t tiles = (int64_t)p.m_tiles * p.n_tiles;
  if (p.n_fastest == 2) {  // reduce-scatter raster, requested by the caller (own_tp set)
    p.m_per = (p.m_tiles + p.own_tp - 1) / p.own_tp;
    tiles = (int64_t)p.own_tp * p.m_per * p.n_tiles;
  } else if (p.grp_mblk != nullptr) {
    p.n_fastest = 1;  // a group's B panel stays L2-resident while its m-tiles are walked
  } else {
    // one wave of CTAs should re-use the SMALLER operand panel from L2 and stream the other exactly once
    p.n_fastest = options().gemm_raster == 1 ? 0 : (options().gemm_raster == 2 ? 1 : (p.n_tiles <= p.m_tiles ? 1 : 0));
  }
  p.num_tiles = (int)tiles;
  int64_t groups = options().gemm_ctas > 0 ? options().gemm_ctas / CG : sm_count() / CG;
  if (groups < 1) groups = 1;
  if (groups > tiles) groups = tiles;
  const bool wd = options().gemm_watchdog != 0 && rs == nullptr;
  if (wd) MMX_CUDA_TRY(cudaMemsetAsync(p.dbg, 0, 64 * sizeof(uint32_t), st));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(groups * CG));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = options().pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  static bool attr_done[kMaxDevices][5] = {};
  static std::mutex attr_mu;
  const int dev = current_device_slot();
  auto prepare = [&](auto kern, int slot) -> cudaError_t {
    std::lock_guard<std::mutex> lk(attr_mu);
    if (attr_done[dev][slot]) return cudaSuccess;
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e == cudaSuccess) attr_done[dev][slot] = true;
    return e;
  };
  if (rs != nullptr) {
    auto kern = mixed_gemm_kernel<CG, false, true>;
    MMX_CUDA_TRY(prepare(kern, 2));
    MMX_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tm, p, *rs));
  } else if (act) {
    const NoRsParams none = {0};
    auto kern = mixed_gemm_kernel<CG, false, false, 1, true>;
    MMX_CUDA_TRY(prepare(kern, 3));
    MMX_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tm, p, none));
  } else if (rope) {
    const NoRsParams none = {0};
    auto kern = mixed_gemm_kernel<CG, false, false, 1, false, true>;
    MMX_CUDA_TRY(prepare(kern, 4));
    MMX_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tm, p, none));
  } else {
    const NoRsParams none = {0};
    auto kern = wd ? mixed_gemm_kernel<CG, true, false> : mixed_gemm_kernel<CG, false, false>;
    MMX_CUDA_TRY(prepare(kern, wd ? 1 : 0));
    MMX_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tm, p, none));
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MMX_OK;
}

// Split-K launch (M <= 128): one cluster of SK CTAs per 128 x 256 tile.
template <int SK>
static int launch_gemm_splitk(const TmapSet& tm, GemmParams& p, cudaStream_t st, bool probe_only, int* max_clusters) {
  using G = Geo<1, false>;
  auto kern = mixed_gemm_kernel<1, false, false, SK>;
  static bool attr_done[kMaxDevices] = {};
  static std::mutex attr_mu;
  const int dev = current_device_slot();
  {
    std::lock_guard<std::mutex> lk(attr_mu);
    if (!attr_done[dev]) {
      MMX_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmemBytes));
      attr_done[dev] = true;
    }
  }
  const int64_t tiles = (int64_t)p.m_tiles * p.n_tiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(tiles * SK));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = G::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = SK;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = options().pdl ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (probe_only) {  // how many clusters of this size can be resident at once (cached by the caller)
    cfg.numAttrs = 1;
    cfg.gridDim = dim3((unsigned)(SK * 64));
    int n = 0;
    MMX_CUDA_TRY(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    *max_clusters = n;
    return MMX_OK;
  }
  const NoRsParams none = {0};
  MMX_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tm, p, none));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return MMX_OK;
}

// Largest split (8, 4, 2) whose clusters -- one per tile -- are all resident in ONE wave and get at least two stages each.
static int choose_splitk(int64_t tiles, int total_stages) {
  const int64_t forced = options().gemm_splitk;
  if (forced == 1) return 1;
  // resident clusters per size and device, -1 = not probed yet
  static int caps[kMaxDevices][9];
  static bool caps_init = false;
  static std::mutex caps_mu;
  std::lock_guard<std::mutex> lk(caps_mu);
  if (!caps_init) {
    for (auto& c : caps)
      for (int i = 0; i < 9; ++i) c[i] = (i == 2 || i == 4 || i == 8) ? -1 : 0;
    caps_init = true;
  }
  int* cap = caps[current_device_slot()];
  TmapSet tm_none;
  GemmParams p_none;
  memset(&p_none, 0, sizeof(p_none));
  for (int sk = 8; sk >= 2; sk >>= 1) {
    if (forced > 1 && forced != sk) continue;
    if (total_stages < 2 * sk && forced != sk) continue;
    if (total_stages < sk) continue;
    if (cap[sk] < 0) {
      int n = 0;
      const int rc = sk == 8 ? launch_gemm_splitk<8>(tm_none, p_none, nullptr, true, &n)
                             : (sk == 4 ? launch_gemm_splitk<4>(tm_none, p_none, nullptr, true, &n)
                                        : launch_gemm_splitk<2>(tm_none, p_none, nullptr, true, &n));
      cap[sk] = rc == MMX_OK ? n : 0;
    }
    if (tiles <= cap[sk]) return sk;
  }
  return 1;
}

int matmul_impl(const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs, const uint8_t* ao,
                const uint8_t* bo, const uint8_t* sfan, const uint8_t* sfbn, const uint8_t* sfas, const uint8_t* sfbs,
                const uint8_t* sfao, const uint8_t* sfbo, int64_t M, int64_t N, int KN, int KS, int KO, int w4,
                const void* bias, void* c, void* stream, RsLaunch* rsl, const MatmulExtra* ex) {
  if (M < 0 || N <= 0 || KN < 0 || KS < 0 || KO < 0 || (KN % 128) || (KS % 128) || (KO % 128) || KN + KS + KO == 0) {
    set_error("matmul: bad shape M=%lld N=%lld (KN,KS,KO)=(%d,%d,%d)", (long long)M, (long long)N, KN, KS, KO);
    return MMX_ERR_INVALID;
  }
  if (N % 128) {
    set_error("matmul: N=%lld must be a multiple of 128", (long long)N);
    return MMX_ERR_INVALID;
  }
  const bool act = ex != nullptr && ex->act_q[0] != nullptr;
  if (act) {
    const int ka = ex->act_k[0] + ex->act_k[1] + ex->act_k[2];
    if (ex->act_k[0] <= 0 || ex->act_k[1] < 0 || ex->act_k[2] < 0 || (ex->act_k[0] % 128) || (ex->act_k[1] % 128) ||
        (ex->act_k[2] % 128) || N != 2 * (int64_t)ka) {
      set_error("matmul_activate_quantize: N=%lld must be twice the activation width (%d,%d,%d), segments multiples of 128",
                (long long)N, ex->act_k[0], ex->act_k[1], ex->act_k[2]);
      return MMX_ERR_INVALID;
    }
    if (bias != nullptr || rsl != nullptr || ex->ag_arrived != nullptr) {
      set_error("matmul_activate_quantize: no bias, no tensor-parallel form");
      return MMX_ERR_INVALID;
    }
    for (int i = 0; i < 3; ++i)
      if (ex->act_k[i] && (!ex->act_q[i] || !ex->act_sf[i] || (((uintptr_t)ex->act_q[i] | (uintptr_t)ex->act_sf[i]) & 15))) {
        set_error("matmul_activate_quantize: outputs of segment %d must be non-null 16-byte aligned pointers", i);
        return MMX_ERR_INVALID;
      }
  } else if (rsl == nullptr && (!c || ((uintptr_t)c & 15))) {
    set_error("matmul: output must be a non-null 16-byte aligned pointer");
    return MMX_ERR_INVALID;
  }
  if (M == 0) return MMX_OK;
  if (!device_is_sm100()) {
    set_error("matmul: this library only runs on sm_100 (B200) devices");
    return MMX_ERR_ARCH;
  }
  // Small problems (M <= 512 and fewer 128-row tiles than half the SMs): single-CTA tiles with K split over a cluster,
  // so that the whole machine streams the operands; everything else: CTA pairs (one 128-row tile: the single-CTA kernel).
  int sk = 1;
  const bool grouped = ex != nullptr && ex->grp_mblk != nullptr;
  const bool gathered = ex != nullptr && ex->ag_arrived != nullptr;
  if (grouped && (ex->grp_n <= 0 || ex->grp_count <= 0 || ex->grp_n != N || (ex->grp_n % 256) || (ex->grp_tile_rows != 128 && ex->grp_tile_rows != 256) ||
                  (M % ex->grp_tile_rows))) {
    set_error("matmul (grouped): rows per group of B must be a multiple of 256, the padded M a multiple of the m-tile");
    return MMX_ERR_INVALID;
  }
  if (rsl == nullptr && !grouped && !gathered && !act && options().gemm_watchdog == 0 && M <= 512 &&
      options().gemm_cta_group != 2) {
    const int64_t tiles1 = ((M + BM - 1) / BM) * ((N + BN - 1) / BN);
    const int stages = (KN + 255) / 256 + KS / 128 + KO / 128;
    // measured (profiles/r01_config2_m_sweep.log): above one row of tiles the split only pays for long K (down_proj,
    // 62 stages: 28.0 -> 16.4 us at M = 256); at K = 4096 the pair kernel is as fast or faster
    if (tiles1 * 2 <= sm_count() && (M <= BM || stages >= 40 || options().gemm_splitk > 1)) sk = choose_splitk(tiles1, stages);
  }
  int cg = (M > 128 && options().gemm_cta_group != 1 && sk == 1) ? 2 : 1;
  if (grouped) cg = ex->grp_tile_rows / BM;  // the caller padded the groups to this m-tile
  if (rsl != nullptr && rsl->shard) cg = 2;  // reduce-scatter ownership is defined on 256-row m-tiles
  const uint8_t* A[3] = {an, as, ao};
  const uint8_t* B[3] = {bn, bs, bo};
  const uint8_t* SA[3] = {sfan, sfas, sfao};
  const uint8_t* SB[3] = {sfbn, sfbs, sfbo};
  const int ks[3] = {KN, KS, KO};
  const int abits[3] = {4, 6, 8};
  const int bbits[3] = {4, w4 ? 4 : 6, w4 ? 4 : 8};
  TmapSet tm;
  GemmParams p;
  memset(&tm, 0, sizeof(tm));
  memset(&p, 0, sizeof(p));
  int ns = 0;
  for (int s = 0; s < 3; ++s) {
    if (ks[s] == 0) continue;
    if (!A[s] || !B[s] || !SA[s] || !SB[s]) {
      set_error("matmul: null pointer for non-empty segment %d", s);
      return MMX_ERR_INVALID;
    }
    if (((uintptr_t)A[s] | (uintptr_t)B[s]) & 31 || ((uintptr_t)SA[s] | (uintptr_t)SB[s]) & 15) {
      set_error("matmul: segment %d operands must be 32-byte aligned (scale factors 16)", s);
      return MMX_ERR_INVALID;
    }
    GemmSeg& g = p.seg[ns];
    g.kind = (s == 0) ? 0 : 1;
    g.kelems = (s == 0) ? 256 : 128;
    const int katoms = ks[s] / 128;
    g.atoms_per_tile = (s == 0) ? 2 : 1;
    g.ktiles = (ks[s] + g.kelems - 1) / g.kelems;
    g.last_atoms = katoms - (g.ktiles - 1) * g.atoms_per_tile;
    g.idesc = make_idesc(g.kind, abits[s], bbits[s], BM * cg);
    const bool unpack = g.kind == 1;
    // bytes TMA reports per row of 128 smem bytes: packed gmem bytes (mode 0) or the smem footprint (mode 1)
    auto row_tx = [&](int bits) -> uint32_t {
      if (!unpack) return 128u;
      if (options().gemm_tx_mode == 1) return 128u;
      return (uint32_t)(128 * bits / 8);
    };
    const int b_rows = BN / cg;
    g.tx_a = (uint32_t)BM * row_tx(abits[s]);
    g.tx_b = (uint32_t)b_rows * row_tx(bbits[s]);
    int rc = get_tmap(A[s], M, ks[s], abits[s], unpack, BM, &tm.a[ns]);
    if (rc) return rc;
    const int64_t n_b = grouped ? (int64_t)ex->grp_n * ex->grp_count : N;  // grouped: the groups' weights stacked on N
    rc = get_tmap(B[s], n_b, ks[s], bbits[s], unpack, b_rows, &tm.b[ns]);
    if (rc) return rc;
    rc = get_sf_tmap(SA[s], (M + 127) / 128, katoms, g.atoms_per_tile, 1, &tm.sfa[ns]);
    if (rc) return rc;
    rc = get_sf_tmap(SB[s], (n_b + 127) / 128, katoms, g.atoms_per_tile, 2, &tm.sfb[ns]);
    if (rc) return rc;
    ++ns;
  }
  RsParams rs;
  if (rsl != nullptr) {
    // fused row-parallel mode: the epilogue stores into the owners' staging tiles (maps prepared by tp_reduce.cu)
    int64_t tiles = ((M + BM * cg - 1) / (BM * cg)) * ((N + BN - 1) / BN);
    if (rsl->shard && rsl->tp >= 1) {
      const int64_t mt = (M + BM * cg - 1) / (BM * cg);
      tiles = (mt + rsl->tp - 1) / rsl->tp * rsl->tp * ((N + BN - 1) / BN);
    }
    if (rsl->tp < 1 || rsl->tp > kMaxTp || (tiles + rsl->tp - 1) / rsl->tp * (BM * cg) > rsl->own_tiles_cap * 256) {
      set_error("matmul_allreduce: %lld tiles over tp=%d exceed the workspace (%lld tiles per rank)", (long long)tiles,
                rsl->tp, (long long)rsl->own_tiles_cap);
      return MMX_ERR_INVALID;
    }
    memset(&rs, 0, sizeof(rs));
    const CUtensorMap* maps = static_cast<const CUtensorMap*>(rsl->dst_maps);
    for (int d = 0; d < rsl->tp; ++d) {
      rs.dst[d] = maps[d];
      rs.tile_flags[d] = rsl->tile_flags[d];
    }
    rs.tp = rsl->tp;
    rs.rank = rsl->rank;
    int64_t groups = options().gemm_ctas > 0 ? options().gemm_ctas / cg : sm_count() / cg;
    if (groups < 1) groups = 1;
    rs.rot_s = (int)((groups + rsl->tp - 1) / rsl->tp * rsl->tp);
    if (rsl->shard) {
      rs.rot_s = 1 << 30;  // no owner rotation: owner(tile) = tile % tp, the raster itself interleaves the owners
      p.n_fastest = 2;
      p.own_tp = rsl->tp;
    }
    rsl->rot_s = rs.rot_s;
    rs.pull = rsl->pull;
    if (rsl->pull) {
      if (int rc = get_c_tmap(rsl->c_local, M, N, &tm.c)) return rc;
    }
  } else if (act) {
    int cacc = 0;
    for (int i = 0; i < 3; ++i) {
      cacc += ex->act_k[i];
      p.act_cend[i] = cacc;
      p.act_katoms[i] = ex->act_k[i] / 128;
      p.act_rowbytes[i] = (int64_t)ex->act_k[i] * (4 + 2 * i) / 8;
      p.act_q[i] = ex->act_q[i];
      p.act_sf[i] = ex->act_sf[i];
    }
    tm.c = tm.a[0];  // never used: the ACT epilogue stores through plain pointers (a valid map keeps the prefetch happy)
  } else if (int rc = get_c_tmap(c, M, N, &tm.c)) {
    return rc;
  }
  p.nseg = ns;
  p.M = M;
  p.N = N;
  p.c = static_cast<__nv_bfloat16*>(c);
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  if (ex != nullptr && ex->residual != nullptr) {
    if (act || rsl != nullptr || grouped || ((uintptr_t)ex->residual & 15)) {
      set_error("matmul: the residual input (16-byte aligned bf16 [M, N]) goes with the plain GEMM only");
      return MMX_ERR_INVALID;
    }
    p.residual = static_cast<const __nv_bfloat16*>(ex->residual);
  }
  if (ex != nullptr && ex->rope_cos != nullptr) {
    if (act || rsl != nullptr || grouped || gathered || ex->residual != nullptr || options().gemm_watchdog != 0 ||
        !ex->rope_sin || (((uintptr_t)ex->rope_cos | (uintptr_t)ex->rope_sin) & 15) || ex->rope_S <= 0 || ex->rope_cols <= 0 ||
        (ex->rope_cols % 128) || ex->rope_cols > N) {
      set_error("matmul_rope: plain GEMM only; cos / sin = 16-byte aligned bf16 [S, 128] tables, rope_cols a multiple of 128 <= N");
      return MMX_ERR_INVALID;
    }
    p.rope_cos = static_cast<const __nv_bfloat16*>(ex->rope_cos);
    p.rope_sin = static_cast<const __nv_bfloat16*>(ex->rope_sin);
    p.rope_cols = ex->rope_cols;
    p.rope_S = ex->rope_S;
  }
  static uint32_t* dbg_addr[kMaxDevices] = {};  // the symbol's address differs from device to device
  {
    const int dev = current_device_slot();
    uint32_t* dbg = dbg_addr[dev];
    if (!dbg) {
      MMX_CUDA_TRY(cudaGetSymbolAddress(reinterpret_cast<void**>(&dbg), g_gemm_dbg));
      dbg_addr[dev] = dbg;
    }
    p.dbg = dbg;
  }
  if (grouped) {
    p.grp_mblk = ex->grp_mblk;
    p.grp_n = ex->grp_n;
    p.rows_dev = ex->rows_dev;
  }
  if (gathered) {
    p.ag_arrived = ex->ag_arrived;
    p.ag_taken = ex->ag_taken;
    p.ag_rows = ex->ag_rows;
    p.ag_ticket = ex->ag_ticket;
    p.ag_tp = ex->ag_tp;
    p.ag_err = ex->ag_err;
    p.ag_timeout_ns = (unsigned long long)options().tp_timeout_ms * 1000000ull;
    for (int d = 0; d < ex->ag_tp && d < kMaxTp; ++d) p.ag_consumed[d] = ex->ag_consumed[d];
  }
  p.flags = (uint32_t)options().gemm_debug_flags;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const RsParams* rsp = rsl != nullptr ? &rs : nullptr;
  if (sk > 1) {
    p.m_tiles = (int)((M + BM - 1) / BM);
    p.n_tiles = (int)((N + BN - 1) / BN);
    p.n_fastest = 1;
    p.num_tiles = p.m_tiles * p.n_tiles;
    if (sk == 8) return launch_gemm_splitk<8>(tm, p, st, false, nullptr);
    if (sk == 4) return launch_gemm_splitk<4>(tm, p, st, false, nullptr);
    return launch_gemm_splitk<2>(tm, p, st, false, nullptr);
  }
  const bool rope = p.rope_cos != nullptr;
  const int rc = cg == 2 ? launch_gemm<2>(tm, p, st, rsp, act, rope) : launch_gemm<1>(tm, p, st, rsp, act, rope);
  if (rc == MMX_OK && rsl != nullptr) {
    rsl->cg = cg;
    rsl->m_tiles = p.m_tiles;
    rsl->n_tiles = p.n_tiles;
    rsl->n_fastest = p.n_fastest;
    rsl->m_per = p.m_per;
  }
  return rc;
}

// bf16 [rows, cols] row-major view for the epilogue's 32x32 TMA stores (C itself, or a staging slot of tile rows)
int encode_store_tmap(void* ptr, int64_t rows, int64_t cols, void* out) {
  return get_c_tmap(ptr, rows, cols, static_cast<CUtensorMap*>(out));
}

}  // namespace mmx

extern "C" __attribute__((visibility("default"))) int mmx_matmul(const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs,
                          const uint8_t* ao, const uint8_t* bo, const uint8_t* sfan, const uint8_t* sfbn,
                          const uint8_t* sfas, const uint8_t* sfbs, const uint8_t* sfao, const uint8_t* sfbo, int64_t M,
                          int64_t N, int KN, int KS, int KO, int w4, const void* bias, void* c, void* stream) {
  return mmx::matmul_impl(an, bn, as, bs, ao, bo, sfan, sfbn, sfas, sfbs, sfao, sfbo, M, N, KN, KS, KO, w4, bias, c,
                          stream, nullptr);
}

// mmx_matmul with a residual input (extension): c = bf16(residual + y), y = mmx_matmul's result (bias included) -- the
// `residual + self_attn(...)` / `residual + mlp(...)` adds of the decoder layer (model/qLlamaLayer.py:116-158) in the GEMM's
// epilogue, with torch's rounding (one fp32 add, one rounding to bf16).  residual: bf16 [M, N], may be c itself.
extern "C" __attribute__((visibility("default"))) int mmx_matmul_residual(
    const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs, const uint8_t* ao, const uint8_t* bo,
    const uint8_t* sfan, const uint8_t* sfbn, const uint8_t* sfas, const uint8_t* sfbs, const uint8_t* sfao,
    const uint8_t* sfbo, int64_t M, int64_t N, int KN, int KS, int KO, int w4, const void* bias, const void* residual, void* c,
    void* stream) {
  mmx::MatmulExtra ex;
  ex.residual = residual;
  return mmx::matmul_impl(an, bn, as, bs, ao, bo, sfan, sfbn, sfas, sfbs, sfao, sfbo, M, N, KN, KS, KO, w4, bias, c, stream,
                          nullptr, &ex);
}

// mmx_matmul with the rotary position embedding of the q / k heads in the epilogue (extension; the reference applies HF's
// apply_rotary_pos_emb with torch ops on the projection outputs, model/qLlamaLayer.py:25-54, 271-272).  The first rope_cols
// columns of C are heads of 128 channels whose B (and bias) rows are stored PAIR-ADJACENT -- row 2j of a head = channel j,
// row 2j + 1 = channel j + 64 (mixedgemm.pair_adjacent_rows) -- so both partners of a rotation sit in one lane; the epilogue
// computes q * cos + rotate_half(q) * sin with the three bf16 roundings of the torch ops and stores every value at its
// ORIGINAL column: C equals mmx_matmul (on the unpermuted weight) followed by mmx_rope_inplace, bit for bit.  cos / sin:
// bf16 [S, 128], row m of C uses table row m % S.  Columns >= rope_cols (v) are plain.
extern "C" __attribute__((visibility("default"))) int mmx_matmul_rope(
    const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs, const uint8_t* ao, const uint8_t* bo,
    const uint8_t* sfan, const uint8_t* sfbn, const uint8_t* sfas, const uint8_t* sfbs, const uint8_t* sfao,
    const uint8_t* sfbo, int64_t M, int64_t N, int KN, int KS, int KO, int w4, const void* bias, const void* cos,
    const void* sin, int64_t S, int rope_cols, void* c, void* stream) {
  if (!cos || S <= 0 || S > 0x7fffffff) {
    mmx::set_error("matmul_rope: null tables or bad S");
    return MMX_ERR_INVALID;
  }
  mmx::MatmulExtra ex;
  ex.rope_cos = cos;
  ex.rope_sin = sin;
  ex.rope_S = (int)S;
  ex.rope_cols = rope_cols;
  return mmx::matmul_impl(an, bn, as, bs, ao, bo, sfan, sfbn, sfas, sfbs, sfao, sfbo, M, N, KN, KS, KO, w4, bias, c, stream,
                          nullptr, &ex);
}

// Fused gate_up GEMM + SiLU(gate) * up + MX quantize (extension; the reference runs matmul, then activate_quantize_x,
// /root/reference/mgemm/src/activate.cu:510-552, on the bf16 result): B holds the gate and up rows INTERLEAVED per 128
// channels of the activation (rows [256 t, 256 t + 128) = gate of channels [128 t, +128), the next 128 rows = up of the same
// channels), N = 2 * (DN + DS + DO).  Outputs are exactly those of mmx_activate_quantize_x(matmul's gate half, up half).
extern "C" __attribute__((visibility("default"))) int mmx_matmul_activate_quantize(
    const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs, const uint8_t* ao, const uint8_t* bo,
    const uint8_t* sfan, const uint8_t* sfbn, const uint8_t* sfas, const uint8_t* sfbs, const uint8_t* sfao,
    const uint8_t* sfbo, int64_t M, int64_t N, int KN, int KS, int KO, int w4, int DN, int DS, int DO, uint8_t* xn, uint8_t* xs,
    uint8_t* xo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo, void* stream) {
  mmx::MatmulExtra ex;
  ex.act_q[0] = xn;
  ex.act_q[1] = xs;
  ex.act_q[2] = xo;
  ex.act_sf[0] = sfn;
  ex.act_sf[1] = sfs;
  ex.act_sf[2] = sfo;
  ex.act_k[0] = DN;
  ex.act_k[1] = DS;
  ex.act_k[2] = DO;
  if (xn == nullptr) {
    mmx::set_error("matmul_activate_quantize: the FP4 segment of the activation must not be empty");
    return MMX_ERR_INVALID;
  }
  return mmx::matmul_impl(an, bn, as, bs, ao, bo, sfan, sfbn, sfas, sfbs, sfao, sfbo, M, N, KN, KS, KO, w4, nullptr, nullptr,
                          stream, nullptr, &ex);
}

// ------------------------------------------------------------------------------------------------ tensor-pipe peak probe
// The roofline denominator of mixed_gemm_kernel, MEASURED: one CTA per SM issues back-to-back block-scaled MMAs of the
// shapes the GEMM uses (M = 128, N = 256; kind::mxf4 K = 64 or kind::mxf8f6f4 K = 32) on operands that never leave
// shared memory -- no TMA, no epilogue, nothing but the tensor pipe (and the power cap).  kind: 0 = mxf4 (E2M1 x E2M1),
// 1 = mxf8f6f4 E3M2 x E2M1, 2 = mxf8f6f4 E4M3 x E2M1.  sf_copies: also issue the GEMM's per-stage tcgen05.cp of the scales.
namespace mmx {
__global__ void __launch_bounds__(128, 1) mma_peak_kernel(int kind, int stages, int sf_copies, uint32_t idesc) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t smem_base = smem_u32(smem_raw);
  constexpr int kA = BM * 128, kB = BN * 128, kSF = 3 * 1024;  // one stage of the single-CTA GEMM
  const uint32_t bar = smem_base + kA + kB + kSF, tmem_slot = bar + 8;
  const int warp = threadIdx.x >> 5;
  // operands: pseudo-random bytes without Inf/NaN patterns (E4M3: exponent 1111 + mantissa 111); scales 2^0
  for (int i = threadIdx.x; i < (kA + kB) / 4; i += blockDim.x) {
    uint32_t h = (uint32_t)i * 2654435761u + blockIdx.x * 40503u;
    h ^= h >> 15;
    h *= 2246822519u;
    h ^= h >> 13;
    reinterpret_cast<uint32_t*>(smem_raw)[i] = h & 0xb7b7b7b7u;
  }
  for (int i = threadIdx.x; i < kSF / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + kA + kB)[i] = 0x7f7f7f7fu;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 1) tmem_alloc<1>(tmem_slot, kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t a_lo = (smem_base >> 4) & 0x3fffu, b_lo = ((smem_base + kA) >> 4) & 0x3fffu;
      const uint32_t sf_lo = ((smem_base + kA + kB) >> 4) & 0x3fffu;
      const uint32_t t_sfa = tmem_base + 256, t_sfb = t_sfa + 8;
      const int natoms = kind == 0 ? 2 : 1;
      auto copy_sf = [&]() {
        for (int a = 0; a < natoms; ++a) {
          tc_cp_sf<1>(t_sfa + 4 * a, desc_from_lo(sf_lo + 32u * a, kDescHiSF));
          tc_cp_sf<1>(t_sfb + 8 * a, desc_from_lo(sf_lo + 64u + 32u * a, kDescHiSF));
          tc_cp_sf<1>(t_sfb + 8 * a + 4, desc_from_lo(sf_lo + 64u + 32u * (natoms + a), kDescHiSF));
        }
      };
      copy_sf();
      for (int s = 0; s < stages; ++s) {
        if (sf_copies && s) copy_sf();
        if (kind == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t sid = (j & 1) ? ((2u << 29) | (2u << 4)) : 0u;
            mma_mxf4<1>(tmem_base, desc_from_lo(a_lo + 2u * j, kDescHiOp), desc_from_lo(b_lo + 2u * j, kDescHiOp), idesc | sid,
                        t_sfa + 4 * (j >> 1), t_sfb + 8 * (j >> 1), (s | j) ? 1u : 0u);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            mma_mxf8f6f4<1>(tmem_base, desc_from_lo(a_lo + 2u * j, kDescHiOp), desc_from_lo(b_lo + 2u * j, kDescHiOp),
                            idesc | ((uint32_t)j << 29) | ((uint32_t)j << 4), t_sfa, t_sfb, (s | j) ? 1u : 0u);
        }
      }
      tc_commit<1>(bar);
    }
    __syncwarp();
    while (!mbar_try_wait(bar, 0)) {
    }
    tc_fence_after();
  }
  __syncthreads();
  if (warp == 1) tmem_dealloc<1>(tmem_base, kTmemCols);
}
}  // namespace mmx

// Runs the probe on every SM and returns the achieved dense TFLOP/s in *tflops (2 * 128 * 256 * K flops per MMA,
// 4 MMAs per stage).  Synchronises the device: a measurement tool (bench.py's roofline peak), not part of the hot path.
extern "C" __attribute__((visibility("default"))) int mmx_debug_mma_peak(int kind, int stages, int sf_copies, int reps,
                                                                      double* tflops, double* ms_out) {
  using namespace mmx;
  if (kind < 0 || kind > 2 || stages <= 0 || reps == 0 || !tflops) {
    set_error("mmx_debug_mma_peak: bad arguments");
    return MMX_ERR_INVALID;
  }
  if (!device_is_sm100()) {
    set_error("mmx_debug_mma_peak: this library only runs on sm_100 (B200) devices");
    return MMX_ERR_ARCH;
  }
  const int smem = 160 * 1024;  // one CTA per SM (each CTA takes the whole TMEM)
  MMX_CUDA_TRY(cudaFuncSetAttribute(mma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const uint32_t idesc = make_idesc(kind == 0 ? 0 : 1, kind == 0 ? 4 : (kind == 1 ? 6 : 8), 4, BM);
  cudaEvent_t e0, e1;
  MMX_CUDA_TRY(cudaEventCreate(&e0));
  MMX_CUDA_TRY(cudaEventCreate(&e1));
  const int grid = sm_count();
  mma_peak_kernel<<<grid, 128, smem>>>(kind, 64, sf_copies, idesc);  // warm-up
  float best = 1e30f;
  if (reps < 0) {
    // SUSTAINED: -reps launches back to back, timed as ONE interval (pick stages x launches >= ~0.5 s so that the power
    // management has settled: a single 10 ms kernel still runs at the boost clock)
    cudaEventRecord(e0);
    for (int r = 0; r < -reps; ++r) mma_peak_kernel<<<grid, 128, smem>>>(kind, stages, sf_copies, idesc);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (e == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) return cuda_fail(e, "mma_peak_kernel");
    const double k_per = kind == 0 ? 64.0 : 32.0;
    *tflops = 2.0 * BM * BN * k_per * 4.0 * stages * grid * (double)(-reps) / (ms * 1e-3) / 1e12;
    if (ms_out) *ms_out = ms;
    return MMX_OK;
  }
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    mma_peak_kernel<<<grid, 128, smem>>>(kind, stages, sf_copies, idesc);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (e != cudaSuccess) {
      cudaEventDestroy(e0);
      cudaEventDestroy(e1);
      return cuda_fail(e, "mma_peak_kernel");
    }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double k_per_mma = kind == 0 ? 64.0 : 32.0;
  const double flops = 2.0 * BM * BN * k_per_mma * 4.0 * stages * grid;
  *tflops = flops / (best * 1e-3) / 1e12;
  if (ms_out) *ms_out = best;
  return MMX_OK;
}

// Grouped mixed GEMM (Mixtral experts, extension; the reference loops over the experts in Python,
// model/qMixtralLayer.py:437-450): ONE persistent launch over all groups.
extern "C" __attribute__((visibility("default"))) int mmx_matmul_grouped(
    const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs, const uint8_t* ao, const uint8_t* bo,
    const uint8_t* sfan, const uint8_t* sfbn, const uint8_t* sfas, const uint8_t* sfbs, const uint8_t* sfao,
    const uint8_t* sfbo, int64_t M, int64_t N, int KN, int KS, int KO, int w4, int groups, int tile_rows,
    const int32_t* grp_mblk, const int32_t* rows_dev, void* c, void* stream) {
  if (!grp_mblk || groups <= 0) {
    mmx::set_error("matmul_grouped: null group table or no groups");
    return MMX_ERR_INVALID;
  }
  mmx::MatmulExtra ex;
  ex.grp_mblk = grp_mblk;
  ex.grp_n = (int)N;
  ex.grp_count = groups;
  ex.grp_tile_rows = tile_rows;
  ex.rows_dev = rows_dev;
  return mmx::matmul_impl(an, bn, as, bs, ao, bo, sfan, sfbn, sfas, sfbs, sfao, sfbo, M, N, KN, KS, KO, w4, nullptr, c, stream,
                          nullptr, &ex);
}

// mmx_matmul_grouped with the SiLU(gate) * up + MX quantize epilogue of mmx_matmul_activate_quantize: every group's B block
// holds its gate (w1) and up (w3) rows interleaved per 128 channels, N = rows per group = 2 * (DN + DS + DO); the outputs are
// the operand tensors of the grouped down (w2) GEMM over the same sorted rows.
extern "C" __attribute__((visibility("default"))) int mmx_matmul_grouped_activate_quantize(
    const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs, const uint8_t* ao, const uint8_t* bo,
    const uint8_t* sfan, const uint8_t* sfbn, const uint8_t* sfas, const uint8_t* sfbs, const uint8_t* sfao,
    const uint8_t* sfbo, int64_t M, int64_t N, int KN, int KS, int KO, int w4, int groups, int tile_rows,
    const int32_t* grp_mblk, const int32_t* rows_dev, int DN, int DS, int DO, uint8_t* xn, uint8_t* xs, uint8_t* xo,
    uint8_t* sfn, uint8_t* sfs, uint8_t* sfo, void* stream) {
  if (!grp_mblk || groups <= 0 || !xn) {
    mmx::set_error("matmul_grouped_activate_quantize: null group table, no groups or an empty FP4 segment");
    return MMX_ERR_INVALID;
  }
  mmx::MatmulExtra ex;
  ex.grp_mblk = grp_mblk;
  ex.grp_n = (int)N;
  ex.grp_count = groups;
  ex.grp_tile_rows = tile_rows;
  ex.rows_dev = rows_dev;
  ex.act_q[0] = xn;
  ex.act_q[1] = xs;
  ex.act_q[2] = xo;
  ex.act_sf[0] = sfn;
  ex.act_sf[1] = sfs;
  ex.act_sf[2] = sfo;
  ex.act_k[0] = DN;
  ex.act_k[1] = DS;
  ex.act_k[2] = DO;
  return mmx::matmul_impl(an, bn, as, bs, ao, bo, sfan, sfbn, sfas, sfbs, sfao, sfbo, M, N, KN, KS, KO, w4, nullptr, nullptr,
                          stream, nullptr, &ex);
}

extern "C" __attribute__((visibility("default"))) int mmx_gemm_debug_status(uint32_t* out, int n) {
  if (!out || n <= 0) return MMX_ERR_INVALID;
  uint32_t tmp[64];
  cudaError_t e = cudaMemcpyFromSymbol(tmp, mmx::g_gemm_dbg, sizeof(tmp));
  if (e != cudaSuccess) return mmx::cuda_fail(e, "cudaMemcpyFromSymbol");
  for (int i = 0; i < n && i < 64; ++i) out[i] = tmp[i];
  return MMX_OK;
}
