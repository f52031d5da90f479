/*
 * micromix_b200.h -- C ABI of libmicromix_b200.so, the B200 (sm_100a) drop-in for MicroMix's hot path.
 *
 * Every entry point replaces one pybind op of the reference's `mixedgemm` module
 * (/root/reference/mgemm/src/bindings.cpp:682-742) and is what a maintainer's binding would call
 * (see INTEGRATION.md for the ctypes / pybind stub).  Plain pointers and sizes only: no torch types.
 *
 * Conventions
 *   - all data pointers are DEVICE pointers, 16-byte aligned, contiguous row-major; the library never allocates
 *     or frees caller-visible memory and keeps no state between calls except a cache of TMA descriptors;
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream, which is what the reference
 *     launches on: reorder.cu:455, w4a4.cu:182); calls are asynchronous and CUDA-graph capturable;
 *   - return value 0 = success, negative = error (mmx_last_error() gives the text; nothing throws or exits,
 *     unlike the reference's CHECK_CUDA -> exit(), gemm_utils.h:53-61);
 *   - K = KN + KS + KO, each a multiple of 128 (model/qLinearLayer.py:40), K <= 32767 (int16 reorder_index).
 *     Unlike the reference (bindings.cpp:134-147: ten hard-coded K) any such K is accepted.
 *
 * Data formats (bit-identical to the reference):
 *   codes   FP4 E2M1 two per byte, even channel in the low nibble  [rows, Kseg/2]      (reorder.cu:30-33)
 *           FP6 E3M2 four codes in three bytes, little-endian bits  [rows, Kseg*3/4]    (reorder.cu:54-63)
 *           FP8 E4M3 one per byte                                   [rows, Kseg]
 *   scales  UE8M0, one per 32 channels, in the 512-byte "SfKMajorAtom" swizzle
 *           offset(r,g) = (r/128)*ceil(Kseg/128)*512 + (g/4)*512 + (r%32)*16 + ((r/32)%4)*4 + g%4
 *           (cutlass/detail/sm100_blockscaled_layout.hpp:54-55,93 via mgemm/include/reorder.cuh:120-125)
 */
#ifndef MICROMIX_B200_H_
#define MICROMIX_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMX_OK 0
#define MMX_ERR_INVALID (-1)  /* bad shape / alignment / null pointer */
#define MMX_ERR_CUDA (-2)     /* a CUDA runtime or driver call failed */
#define MMX_ERR_ARCH (-3)     /* device is not sm_100 */

/* Library version (major*100 + minor). */
int mmx_version(void);

/* Text of the last error on the calling thread ("" if none). */
const char* mmx_last_error(void);

/* Bytes of a scale-factor buffer, exactly as the reference allocates them:
 *   activations (M/128+1)*128*Kseg/32          bindings.cpp:120-123
 *   weights     ceil(N/128)*128*Kseg/32        bindings.cpp:170-172 (== N*Kseg/32 for N % 128 == 0) */
int64_t mmx_sf_bytes_act(int64_t M, int64_t Kseg);
int64_t mmx_sf_bytes_wgt(int64_t N, int64_t Kseg);

/* Byte offset of the scale of (row, 32-channel group) in a segment of Kseg channels. */
int64_t mmx_sf_offset(int64_t row, int64_t group, int64_t Kseg);

/*
 * mixedgemm.reorder_quantize_x(X, reorder_index, KN, KS, KO)            bindings.cpp:104-151
 *   -> run_reorder_quantize_x<32,K> -> reorder_quantize_mixed_kernel      reorder.cu:434-469, 94-269
 * Per row: gather X[r, idx[j]], per-32 absmax, E8M0 scale, convert to FP4 | FP6 | FP8, pack, store.
 *   x   bf16 [M, K]          idx int16 [K]
 *   xn  u8 [M, KN/2]   xs u8 [M, KS*3/4]   xo u8 [M, KO]
 *   sfn/sfs/sfo  u8, mmx_sf_bytes_act(M, Kseg) bytes each; rows >= M of a partly filled 128-row block are
 *   written with defined bytes (the reference leaves them uninitialised).
 * Pointers of an empty segment (Kseg == 0) may be NULL.
 */
int mmx_reorder_quantize_x(const void* x, int64_t M, int K, const int16_t* idx, int KN, int KS, int KO, uint8_t* xn,
                           uint8_t* xs, uint8_t* xo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo, void* stream);

/*
 * mixedgemm.reorder_quantize_w(W, reorder_index, KN, KS, KO)            bindings.cpp:155-202
 *   -> run_reorder_quantize_w<32,K>                                       reorder.cu:471-506
 * Same as _x on weight rows (FP4 | FP6 | FP8), SF buffers of mmx_sf_bytes_wgt(N, Kseg) bytes.
 */
int mmx_reorder_quantize_w(const void* w, int64_t N, int K, const int16_t* idx, int KN, int KS, int KO, uint8_t* wn,
                           uint8_t* ws, uint8_t* wo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo, void* stream);

/*
 * mixedgemm.reorder_quantize_w4(W, reorder_index, KN, KS, KO)           bindings.cpp:206-253
 *   -> run_reorder_quantize_w4<32,K> -> reorder_quantize_mxfp4_kernel     reorder.cu:508-543, 271-432
 * All three segments FP4:  wn u8 [N, KN/2]   ws u8 [N, KS/2]   wo u8 [N, KO/2].
 */
int mmx_reorder_quantize_w4(const void* w, int64_t N, int K, const int16_t* idx, int KN, int KS, int KO, uint8_t* wn,
                            uint8_t* ws, uint8_t* wo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo, void* stream);

/*
 * mixedgemm.rmsnorm_quantize_x(X, W, eps, reorder_index, KN, KS, KO)    bindings.cpp:257-303
 *   -> run_rmsnorm_bf16_mixed<32,K> -> rmsnorm_bf16_mixed_kernel          rmsnorm.cu:96-310
 * RMSNorm fused into reorder+quantize:  y[r,c] = bf16((float(x[r,c]) * float(w[c])) * rinv[r]),
 * rinv[r] = 1 / sqrt(sum_c x[r,c]^2 / K + eps), then exactly mmx_reorder_quantize_x on y.  Any K that
 * mmx_reorder_quantize_x accepts (the reference: K in {3072,3584,4096,5120} and a reduction that is only
 * right for 128 threads, rmsnorm.cu:166-173); codes are RNE like every other op (the reference rounds
 * x/scale to an INTEGER first, rmsnorm.cu:264 -- a defect that is not reproduced).
 * The sum of squares is taken in a fixed order so that the result does not depend on the launch shape:
 * fp32 fma chain over each aligned 8-channel chunk, then a perfect binary tree over the chunk index.
 *   x bf16 [M, K]   w bf16 [K]   outputs as mmx_reorder_quantize_x
 */
int mmx_rmsnorm_quantize_x(const void* x, const void* w, float eps, int64_t M, int K, const int16_t* idx, int KN, int KS,
                           int KO, uint8_t* xn, uint8_t* xs, uint8_t* xo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo,
                           void* stream);

/*
 * mixedgemm.activate_quantize_x(A, B, KN, KS, KO)                       bindings.cpp:307-334
 *   -> run_activate_bf16_mixed -> activate_quantize_kernel_with_cute_layout   activate.cu:510-552, 40-202
 * v = silu(float(a)) * float(b) in fp32, NO channel permutation (the producer already emits down_proj's order),
 * per-32 absmax, scale 2^ceil(log2(amax/QMAX)) (1.0 when amax <= 1e-6), codes RNE straight from fp32.
 *   a, b bf16 [M, K], K = KN+KS+KO     outputs as mmx_reorder_quantize_x
 */
int mmx_activate_quantize_x(const void* a, const void* b, int64_t M, int KN, int KS, int KO, uint8_t* xn, uint8_t* xs,
                            uint8_t* xo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo, void* stream);
/* The same op on column slices of a wider matrix (extension; the reference op takes dense tensors only): row r of a / b
 * starts at a + r * ld elements, ld >= K, ld % 8 == 0.  Used on the fused gate_up GEMM output [M, 2 * intermediate]. */
int mmx_activate_quantize_x_strided(const void* a, const void* b, int64_t ld, int64_t M, int KN, int KS, int KO, uint8_t* xn,
                                    uint8_t* xs, uint8_t* xo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo, void* stream);

/* mmx_matmul with a residual input (extension; the reference adds the residual with a torch op after the linear,
 * model/qLlamaLayer.py:116-158): c = bf16(residual + y) with y = mmx_matmul's bf16 result (bias included) -- torch's bf16
 * add, bit for bit, in the GEMM's epilogue.  residual: bf16 [M, N], 16-byte aligned, may alias c. */
int mmx_matmul_residual(const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs, const uint8_t* ao,
                        const uint8_t* bo, const uint8_t* sfan, const uint8_t* sfbn, const uint8_t* sfas, const uint8_t* sfbs,
                        const uint8_t* sfao, const uint8_t* sfbo, int64_t M, int64_t N, int KN, int KS, int KO, int w4,
                        const void* bias, const void* residual, void* c, void* stream);

/* mmx_matmul with the rotary position embedding of the q / k heads in the epilogue (extension; the reference applies HF's
 * apply_rotary_pos_emb with torch ops, model/qLlamaLayer.py:25-54, 271-272).  The first rope_cols columns of C (a multiple of
 * 128) are heads of 128 channels whose B and bias rows are stored PAIR-ADJACENT: row 2j of a head = channel j, row 2j + 1 =
 * channel j + 64.  The epilogue computes q * cos + rotate_half(q) * sin with the torch ops' three bf16 roundings and stores
 * every value at its ORIGINAL column: C = mmx_matmul on the unpermuted weight followed by mmx_rope_inplace, bit for bit.
 * cos / sin: bf16 [S, 128], 16-byte aligned; row m uses table row m % S.  Columns >= rope_cols are plain. */
int mmx_matmul_rope(const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs, const uint8_t* ao,
                    const uint8_t* bo, const uint8_t* sfan, const uint8_t* sfbn, const uint8_t* sfas, const uint8_t* sfbs,
                    const uint8_t* sfao, const uint8_t* sfbo, int64_t M, int64_t N, int KN, int KS, int KO, int w4,
                    const void* bias, const void* cos, const void* sin, int64_t S, int rope_cols, void* c, void* stream);

/* matmul + activate_quantize_x in ONE kernel (extension; the reference runs mixedgemm.matmul for gate_proj / up_proj,
 * model/qLlamaLayer.py:324-387, then activate_quantize_x on the two bf16 results): the GEMM epilogue rounds its accumulators
 * to bf16 (what mmx_matmul would have stored), evaluates silu(gate) * up and MX-quantizes it -- the [M, 2 * inter] bf16
 * intermediate never exists in memory.  Weight layout: the B tensors hold gate and up rows INTERLEAVED per 128 activation
 * channels: rows [256 t, 256 t + 128) = gate rows of channels [128 t, 128 t + 128) (in down_proj's channel order), rows
 * [256 t + 128, 256 t + 256) = the up rows of the same channels; N = 2 * (DN + DS + DO).  (DN, DS, DO) = FP4 | FP6 | FP8
 * split of the activation, multiples of 128, DN > 0.  No bias.  Outputs are bit-identical to
 * mmx_activate_quantize_x(gate half, up half of mmx_matmul's result), including the 0x7F scale bytes of padding rows. */
int mmx_matmul_activate_quantize(const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs, const uint8_t* ao,
                                 const uint8_t* bo, const uint8_t* sfan, const uint8_t* sfbn, const uint8_t* sfas,
                                 const uint8_t* sfbs, const uint8_t* sfao, const uint8_t* sfbo, int64_t M, int64_t N, int KN,
                                 int KS, int KO, int w4, int DN, int DS, int DO, uint8_t* xn, uint8_t* xs, uint8_t* xo,
                                 uint8_t* sfn, uint8_t* sfs, uint8_t* sfo, void* stream);

/* Test hook for the epilogue above: fp32 bits of silu(x) for EVERY bf16 x (index = bit pattern, 65536 entries, device
 * memory) through the branch-free sequence the epilogue uses (`fast`) and through the reference sequence (`ref`); the two
 * must agree for 2^-60 <= |x| <= 32, the range the epilogue uses the fast one on. */
int mmx_debug_silu_table(uint32_t* fast, uint32_t* ref, void* stream);

/*
 * mixedgemm.downproj_quantize_w(W, KN, KS, KO)                          bindings.cpp:336-360
 *   -> run_downproj_bf16_mixed -> downproj_quantize_kernel_with_cute_layout   activate.cu:554-592, 204-349
 * mixedgemm.downproj_quantize_w4(W, KN, KS, KO)                         bindings.cpp:362-387
 *   -> run_downproj_bf16_mxfp4 -> downproj_quantize_kernel_with_cute_layout_w4  activate.cu:594-632, 351-507
 * Weight rows that are already in down_proj's channel order: the activate op's quantizer on float(w), to
 * FP4 | FP6 | FP8 (_w) or FP4 in all three segments (_w4).  SF buffers are sized like activations,
 * mmx_sf_bytes_act(N, Kseg) (bindings.cpp:346-348).
 */
int mmx_downproj_quantize_w(const void* w, int64_t N, int KN, int KS, int KO, uint8_t* wn, uint8_t* ws, uint8_t* wo,
                            uint8_t* sfn, uint8_t* sfs, uint8_t* sfo, void* stream);
int mmx_downproj_quantize_w4(const void* w, int64_t N, int KN, int KS, int KO, uint8_t* wn, uint8_t* ws, uint8_t* wo,
                             uint8_t* sfn, uint8_t* sfs, uint8_t* sfo, void* stream);

/*
 * mixedgemm.matmul(AN,BN,AS,BS,AO,BO,SFAN,SFBN,SFAS,SFBS,SFAO,SFBO)      bindings.cpp:50-102
 *   -> matmul_w4_host (w4 != 0: W4A4 + W4A6 + W4A8)                        gemm.cu:53-78
 *   -> matmul_host    (w4 == 0: W4A4 + W6A6 + W8A8)                        gemm.cu:26-51
 * C[M,N] (bf16) = sum over the three K segments of (A_seg o SFA_seg)(B_seg o SFB_seg)^T.
 * One persistent tcgen05 kernel; all segments accumulate in one fp32 TMEM accumulator and are rounded to bf16
 * once (the reference rounds to bf16 after each of its three launches).  C needs no pre-zeroing.
 *   bias  optional bf16 [N] (NULL = none): C = bf16(float(bf16(acc)) + bias), i.e. exactly the reference's
 *         separate `y + self.bias` (model/qLinearLayer.py:70-71) fused into the epilogue.
 *   N must be a multiple of 128 (the reference's SF sizing assumes it, bindings.cpp:170-172).
 */
int mmx_matmul(const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs, const uint8_t* ao,
               const uint8_t* bo, const uint8_t* sfan, const uint8_t* sfbn, const uint8_t* sfas, const uint8_t* sfbs,
               const uint8_t* sfao, const uint8_t* sfbo, int64_t M, int64_t N, int KN, int KS, int KO, int w4,
               const void* bias, void* c, void* stream);

/*
 * Row-parallel linear fused with its all-reduce over NVLink peer memory (tensor parallelism inside one NVSwitch box).
 * The reference has no counterpart: model/parallel_utils.py:89-163 only places whole layers on different GPUs.
 * BASELINE.json's north_star asks for row-parallel o_proj / down_proj whose bf16 partials are summed across ranks;
 * mmx_matmul + ncclAllReduce is the plain way, this is the fused one (micromix_b200/csrc/tp_reduce.cu):
 * the GEMM epilogue pushes each partial tile to the rank that owns it while the tensor cores run the next tile, and
 * a co-resident reducer kernel sums a tile as soon as all tp partials have landed and writes it to every rank's C.
 *
 *   mmx_peer_alloc / _open / _close / _free   one zero-filled device workspace per rank + its 64-byte cudaIpc handle;
 *       the host code exchanges the handles (torch.distributed all_gather_object) and opens the peers' workspaces.
 *       (Any other allocator that yields peer-mapped pointers works too: the Python host code uses torch's symmetric
 *       memory when it also wants the multicast mapping, see mmx_tp_ctx_set_multicast.)
 *   mmx_tp_workspace_bytes(M_cap, N_cap, tp)   size of that workspace for outputs up to M_cap x N_cap.
 *   mmx_tp_ctx_create(ws, tp, rank, ...)       ws[d] = rank d's workspace as mapped in THIS process (ws[rank] = own).
 *   mmx_matmul_allreduce(ctx, <mmx_matmul operands of this rank's K shard>, M, N, KN, KS, KO, w4, bias, &c, stream)
 *       C = sum over ranks of this rank's partial product; *c receives a pointer INTO the own workspace holding the
 *       full bf16 [M, N] result (valid until the second next call on this context); bias is added by the rank
 *       that passes it (pass it on rank 0 only).  All ranks must issue the same call sequence on one stream.
 *   mmx_tp_status(ctx, &w)                     0 = clean; bit 0 / bit 1 = a cross-rank wait timed out
 *       (option "tp_timeout_ms", default 10 s) -- a lost peer produces an error word, never a hung GPU.
 */
#define MMX_PEER_HANDLE_BYTES 64
int mmx_peer_alloc(int64_t bytes, void** ptr, uint8_t* handle);
int mmx_peer_open(const uint8_t* handle, void** ptr);
int mmx_peer_close(void* ptr);
int mmx_peer_free(void* ptr);
int64_t mmx_tp_workspace_bytes(int64_t M_cap, int64_t N_cap, int tp);
int mmx_tp_ctx_create(void* const* ws, int tp, int rank, int64_t M_cap, int64_t N_cap, void** ctx);
/* Optional: mc_ws = an NVSwitch MULTICAST mapping of the same workspaces (cuMulticast*; the Python host code gets it from
 * torch's symmetric-memory rendezvous).
 *   in_switch_reduce = 0: the reducer writes each result tile with one multimem.st instead of tp peer stores (same bits).
 *   in_switch_reduce = 1: partial tiles stay in every rank's own C; the owner's reducer sums them INSIDE the switch
 *     (multimem.ld_reduce, fp32 accumulation) and multicasts the bf16 result: per rank and direction NVLink carries the
 *     output once instead of 2(tp-1)/tp times.  All ranks receive identical bits; the summation order is the switch's,
 *     so the result may differ from the rank-ordered fp32 sum in the last bf16 bit. */
int mmx_tp_ctx_set_multicast(void* ctx, void* mc_ws, int in_switch_reduce);
int mmx_tp_ctx_destroy(void* ctx);
int mmx_tp_status(void* ctx, uint32_t* out);
int mmx_matmul_allreduce(void* ctx, const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs,
                         const uint8_t* ao, const uint8_t* bo, const uint8_t* sfan, const uint8_t* sfbn,
                         const uint8_t* sfas, const uint8_t* sfbs, const uint8_t* sfao, const uint8_t* sfbo, int64_t M,
                         int64_t N, int KN, int KS, int KO, int w4, const void* bias, void** c_out, void* stream);

/*
 * Sequence-parallel form of the same tensor parallelism (Megatron "SP"; no reference counterpart either).  The all-reduce
 * above leaves the whole bf16 [M, N] on every rank, after which every rank repeats the same residual / RMSNorm / quantize.
 * Here the row-parallel linear ends in a REDUCE-SCATTER and the next column-parallel linear starts with an ALL-GATHER of
 * PACKED MX CODES:
 *   mmx_tp_shard_rows(M, tp)            rows per rank: ceil(ceil(M/256)/tp)*256 (whole 256-row GEMM tiles)
 *   mmx_matmul_reduce_scatter(...)       like mmx_matmul_allreduce, but rank r ends with rows
 *       [r*shard, min(M,(r+1)*shard)) only: *c_out points at its first row (inside the own workspace), *row0 / *rows
 *       say which rows.  In switch mode the owner pulls its rows through multimem.ld_reduce and nothing is sent back
 *       (NVLink ingress per rank: C/tp instead of C); tiles are walked owner-interleaved so that all owners' rows complete
 *       at an even pace while the GEMM is still running.
 *   mmx_tp_quantize_allgather(ctx, x_shard, M, K, idx, KN, KS, KO, norm_w, eps, views, stream)
 *       mmx_reorder_quantize_x (norm_w == NULL) or mmx_rmsnorm_quantize_x on THIS rank's rows of the [M, K] activation
 *       (x_shard = its first row); every store goes to the NVSwitch multicast address of the context's gather channel,
 *       so the codes and scales of all M rows land in every rank's workspace: 0.53-0.66 bytes per element on the wire
 *       instead of 2.  views[0..5] receive the LOCAL addresses of (XN, XS, XO, SFXN, SFXS, SFXO) of the gathered
 *       activation.  Needs a context created with mmx_tp_ctx_create_ex(..., ag_M, ag_K) and a multicast mapping.
 *   mmx_tp_matmul_gathered(ctx, <weights>, M, N, KN, KS, KO, w4, bias, c, stream)
 *       mmx_matmul whose A operand is the gathered activation: the TMA producer waits, m-tile by m-tile, for the arrival
 *       counter of the rank that quantized those rows; the last CTA tells every rank that the channel may be reused.
 *       Exactly ONE gathered matmul must follow each gather on every rank.
 * All counters of this protocol live in device memory and only ever count up: the sequence is CUDA-graph replayable.
 */
int64_t mmx_tp_workspace_bytes_ex(int64_t M_cap, int64_t N_cap, int tp, int64_t ag_M, int64_t ag_K);
int mmx_tp_ctx_create_ex(void* const* ws, int tp, int rank, int64_t M_cap, int64_t N_cap, int64_t ag_M, int64_t ag_K,
                         void** ctx);
int64_t mmx_tp_shard_rows(int64_t M, int tp);
int mmx_matmul_reduce_scatter(void* ctx, const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs,
                              const uint8_t* ao, const uint8_t* bo, const uint8_t* sfan, const uint8_t* sfbn,
                              const uint8_t* sfas, const uint8_t* sfbs, const uint8_t* sfao, const uint8_t* sfbo, int64_t M,
                              int64_t N, int KN, int KS, int KO, int w4, const void* bias, void** c_out, int64_t* row0,
                              int64_t* rows, void* stream);
int mmx_tp_quantize_allgather(void* ctx, const void* x_shard, int64_t M, int K, const int16_t* idx, int KN, int KS, int KO,
                              const void* norm_w, float eps, void** views, void* stream);
int mmx_tp_matmul_gathered(void* ctx, const uint8_t* bn, const uint8_t* bs, const uint8_t* bo, const uint8_t* sfbn,
                           const uint8_t* sfbs, const uint8_t* sfbo, int64_t M, int64_t N, int KN, int KS, int KO, int w4,
                           const void* bias, void* c, void* stream);

/*
 * TOKEN-PARALLEL row linears (the alternative to the K-sharded GEMM + collective above; no reference counterpart).  The
 * MXFP4 weights of o_proj / down_proj are replicated; the ranks exchange the PACKED MX CODES of the activation (an
 * all-to-all of (tp-1)/tp * M * K/tp * ~0.66 bytes per rank) instead of bf16 partial sums ((tp-1)/tp * M * N * 2 bytes):
 *   mmx_tp_quantize_alltoall(ctx, x_local, M, K_local, idx_local, KN, KS, KO, seg_tot, seg_off, views, stream)
 *       mmx_reorder_quantize_x of this rank's K slice (x_local bf16 [M, K_local], rank-local permutation and split, as in
 *       the row-parallel form); the codes and scales of rows [d*shard, (d+1)*shard) are written into rank d's exchange
 *       buffer (the gather channel), at channel offset seg_off[i] inside segment i of the [shard, K_total] activation whose
 *       segments hold seg_tot[i] channels (all multiples of 128).  views[0..5]: local addresses of that activation.
 *   mmx_tp_matmul_exchanged(ctx, <replicated weights>, M, N, KN_tot, KS_tot, KO_tot, w4, bias, c, &row0, &rows, stream)
 *       the three-segment GEMM over the full K on this rank's rows (c: bf16 [rows, N]) once every rank's columns have
 *       landed.  Bit-identical to mmx_reorder_quantize_x + mmx_matmul on one GPU with the rank-blocked permutation.
 * Exactly one mmx_tp_matmul_exchanged must follow each mmx_tp_quantize_alltoall on every rank.
 */
int mmx_tp_quantize_alltoall(void* ctx, const void* x_local, int64_t M, int K_local, const int16_t* idx_local, int KN, int KS,
                             int KO, const int32_t* seg_tot, const int32_t* seg_off, void** views, void* stream);
int mmx_tp_matmul_exchanged(void* ctx, const uint8_t* bn, const uint8_t* bs, const uint8_t* bo, const uint8_t* sfbn,
                            const uint8_t* sfbs, const uint8_t* sfbo, int64_t M, int64_t N, int KN, int KS, int KO, int w4,
                            const void* bias, void* c, int64_t* row0, int64_t* rows, void* stream);

/*
 * Grouped forms for Mixtral's experts (extension).  The reference runs one Python iteration per expert -- quantize,
 * matmul, quantize, matmul, index_add_ (model/qMixtralLayer.py:437-450, 502-519).  Here the (token, slot) pairs are sorted
 * by expert once, each expert's rows padded to whole m-tiles, and ONE quantize launch + ONE GEMM launch serve all experts:
 *   mmx_reorder_quantize_x_grouped   x bf16 [M, K] (sorted, padded), idx int16 [groups, K] (one permutation per group, the
 *       same (KN, KS, KO) for all), grp_rowblk int32 [ceil(M/128)] in DEVICE memory = group of each 128-row block
 *       (padding blocks carry any valid group); row_src int32 [M] (optional, DEVICE memory): sorted row r is row
 *       row_src[r] of x, i.e. the gather of the routed tokens is fused into the quantizer's loads (padding rows carry any
 *       valid row); rows_dev int32 (optional, DEVICE memory): rows that actually exist, a multiple of 128 -- row blocks
 *       past it are not touched.  Outputs as mmx_reorder_quantize_x.
 *   mmx_matmul_grouped               B tensors = the groups' weights stacked on N ([groups*N, Kseg*bits/8], scales likewise);
 *       grp_mblk int32 [M / tile_rows] in DEVICE memory = group of each m-tile or -1 (padding tile: skipped);
 *       tile_rows = 256 (CTA pairs) or 128; rows_dev int32 (optional, DEVICE memory) = rows of A that exist (whole
 *       m-tiles): the tile walk stops there; C bf16 [M, N].  No host synchronisation anywhere: the tables are written
 *       on the stream by the router.
 */
int mmx_reorder_quantize_x_grouped(const void* x, int64_t M, int K, const int16_t* idx, const int32_t* grp_rowblk,
                                   const int32_t* row_src, const int32_t* rows_dev, int KN, int KS, int KO, uint8_t* xn,
                                   uint8_t* xs, uint8_t* xo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo, void* stream);
/* mmx_matmul_grouped with the epilogue of mmx_matmul_activate_quantize: every group's block of B rows holds its gate (w1) and
 * up (w3) rows interleaved per 128 channels, N = B rows per group = 2 * (DN + DS + DO); outputs = the operand tensors of the
 * grouped down (w2) GEMM over the same sorted rows (rows of skipped padding tiles are not written). */
int mmx_matmul_grouped_activate_quantize(const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs,
                                         const uint8_t* ao, const uint8_t* bo, const uint8_t* sfan, const uint8_t* sfbn,
                                         const uint8_t* sfas, const uint8_t* sfbs, const uint8_t* sfao, const uint8_t* sfbo,
                                         int64_t M, int64_t N, int KN, int KS, int KO, int w4, int groups, int tile_rows,
                                         const int32_t* grp_mblk, const int32_t* rows_dev, int DN, int DS, int DO, uint8_t* xn,
                                         uint8_t* xs, uint8_t* xo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo, void* stream);
/* mmx_activate_quantize_x_strided on the first *rows_dev rows only (rows_dev: int32 in DEVICE memory, a multiple of 128;
 * M = the static upper bound the buffers were sized for): the expert-sorted matrix of the grouped path is padded to a
 * bound known on the host, its used length only on the device. */
int mmx_activate_quantize_x_rows(const void* a, const void* b, int64_t ld, int64_t M, const int32_t* rows_dev, int KN, int KS,
                                 int KO, uint8_t* xn, uint8_t* xs, uint8_t* xo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo,
                                 void* stream);
/* The routing tables of the grouped expert path in ONE launch (extension; the reference loops over the experts on the host,
 * model/qMixtralLayer.py:437-450): a stable counting sort of the (token, slot) pairs by local expert, every expert's rows
 * padded to whole m-tiles.  sel int64 [T, top_k] expert ids, local_slot int64 [experts] = slot of the expert on this rank or
 * n_local (elsewhere), tile 128 | 256, Mp = a multiple of tile >= T * top_k + n_local * (tile - 1) rounded up.  Outputs (device,
 * int32): row_src [Mp] token of each padded row (padding: 0), pair_row [T, top_k] padded row of each pair or -1,
 * grp_rowblk [Mp / 128], grp_mtile [Mp / tile] expert slot per row block / m-tile (-1 = unused m-tile), rows_used [1].
 * n_local <= 64, experts <= 256. */
int mmx_moe_route(const int64_t* sel, const int64_t* local_slot, int64_t T, int top_k, int experts, int n_local, int tile, int64_t Mp,
                  int32_t* row_src, int32_t* pair_row, int32_t* grp_rowblk, int32_t* grp_mtile, int32_t* rows_used,
                  void* stream);
/* out[t] = sum over the top_k slots of token t, in ascending expert order, of bf16(y[row[t,s]] * w[t,s]) with a bf16
 * rounding after every add -- exactly what the reference's per-expert `index_add_` loop leaves in its bf16 buffer
 * (model/qMixtralLayer.py:446-450).  row int32 [T, top_k] = row of y holding the pair's expert output, or -1 (expert not on
 * this rank: skipped); expert int32 [T, top_k]; w bf16 [T, top_k]; y bf16 [*, H]; out bf16 [T, H]. */
int mmx_moe_combine(const void* y, const int32_t* row, const int32_t* expert, const void* w, int64_t T, int top_k, int H,
                    void* out, void* stream);
int mmx_matmul_grouped(const uint8_t* an, const uint8_t* bn, const uint8_t* as, const uint8_t* bs, const uint8_t* ao,
                       const uint8_t* bo, const uint8_t* sfan, const uint8_t* sfbn, const uint8_t* sfas, const uint8_t* sfbs,
                       const uint8_t* sfao, const uint8_t* sfbo, int64_t M, int64_t N, int KN, int KS, int KO, int w4,
                       int groups, int tile_rows, const int32_t* grp_mblk, const int32_t* rows_dev, void* c, void* stream);

/*
 * Rotary position embedding IN PLACE on the first heads*head_dim columns of every row of y (bf16 [M, ld]) -- the q and k
 * columns of the fused qkv GEMM output.  Extension: the reference keeps RoPE in PyTorch (model/qLlamaLayer.py:25-54,
 * 271-272); same arithmetic and the same bf16 roundings as HF's apply_rotary_pos_emb:
 *   y[.., :d/2] = bf16(bf16(x1*cos1) + bf16(-x2*sin1)),   y[.., d/2:] = bf16(bf16(x2*cos2) + bf16(x1*sin2)).
 * cos, sin bf16 [S, head_dim]; row m uses table row m % S (S = seq_len for a batch of equal-length prefills, S = M for a
 * table that is already per token).
 */
int mmx_rope_inplace(void* y, int64_t ld, int64_t M, int heads, int head_dim, const void* cos, const void* sin, int64_t S,
                     void* stream);

/* Tensor-pipe peak probe (measurement tool behind bench.py's roofline denominator): every SM issues `stages` x 4
 * back-to-back block-scaled MMAs (M=128, N=256) on shared-memory-resident operands; kind 0 = kind::mxf4 (K=64),
 * 1 = kind::mxf8f6f4 E3M2 x E2M1, 2 = kind::mxf8f6f4 E4M3 x E2M1 (K=32).  reps > 0: best of `reps` launches (burst);
 * reps < 0: -reps launches back to back timed as one interval (sustained, under the power cap).  Synchronises the device. */
int mmx_debug_mma_peak(int kind, int stages, int sf_copies, int reps, double* tflops, double* ms);

/* Number of kernels launched by this library since load (bench.py's gpu_launches counter). */
int64_t mmx_launch_count(void);

/* Tuning / bring-up knobs: key in {"gemm_ctas" (cap the persistent GEMM grid), "gemm_cta_group", "gemm_splitk" (M <= 512:
 * 0 auto, 1 never, 2|4|8 force), "pdl", "tp_timeout_ms", "tp_reduce_ctas"} and, for tests / timing experiments only (results
 * may be wrong): {"gemm_watchdog", "gemm_tx_mode", "gemm_debug_flags", "gemm_raster", "quant_rows", "quant_variant",
 * "quant_ctas", "tp_debug"}.  Returns non-zero for an unknown key. */
int mmx_set_option(const char* key, int64_t value);

/* After a GEMM launched with the watchdog on: copies the kernel's status words (0 = clean) to out[0..n).
 * Words 40..43: globaltimer (lo, hi) at the start / end of CTA 0's epilogue of the last fused (RS) GEMM. */
int mmx_gemm_debug_status(uint32_t* out, int n);

/* Timeline probe of the last tile_allreduce_kernel launch (reducer CTA 0, globaltimer ns): entry, first owned tile
 * complete, last unit reduced, final cross-rank arrival seen.  tools/tp_fused_probe.py; not part of the product path. */
int mmx_tp_debug_times(uint64_t* out, int n);

#ifdef __cplusplus
}
#endif
#endif /* MICROMIX_B200_H_ */
