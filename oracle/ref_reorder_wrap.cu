// ref_reorder_wrap.cu -- TEST INFRASTRUCTURE.  C-ABI shim over the REFERENCE's own quantize launchers
// (run_reorder_quantize_{x,w,w4}<32,K>, declared in /root/reference/mgemm/include/reorder.cuh:296-337 and
// explicitly instantiated in /root/reference/mgemm/src/reorder.cu:545-694).  Linked with the reference
// reorder.cu compiled in place for sm_100a into oracle/_ref/libref_reorder.so (see oracle/Makefile).
// Used on the B200 box by tests (-m gpu) and tools/make_golden_ref.py as the bit-exact GPU oracle.
#include <cstdint>
#include <cuda_runtime.h>
#include "cutlass/numeric_types.h"

typedef cutlass::float_ue8m0_t sf_t;
typedef cutlass::bfloat16_t bf16_t;

template <int group_size, int hidden_dim>
void run_reorder_quantize_x(bf16_t*, int, int16_t*, uint8_t*, uint8_t*, uint8_t*, sf_t*, sf_t*, sf_t*, int, int, int);
template <int group_size, int hidden_dim>
void run_reorder_quantize_w(bf16_t*, int, int16_t*, uint8_t*, uint8_t*, uint8_t*, sf_t*, sf_t*, sf_t*, int, int, int);
template <int group_size, int hidden_dim>
void run_reorder_quantize_w4(bf16_t*, int, int16_t*, uint8_t*, uint8_t*, uint8_t*, sf_t*, sf_t*, sf_t*, int, int, int);

#define REF_CASE(FN, KV)                                                                                      \
  case KV:                                                                                                    \
    FN<32, KV>((bf16_t*)x, rows, idx, qn, qs, qo, (sf_t*)sfn, (sf_t*)sfs, (sf_t*)sfo, KN, KS, KO);            \
    break;

#define REF_SWITCH(FN)                                                                                        \
  switch (KN + KS + KO) {                                                                                     \
    REF_CASE(FN, 3072) REF_CASE(FN, 3584) REF_CASE(FN, 4096) REF_CASE(FN, 5120) REF_CASE(FN, 8192)            \
    REF_CASE(FN, 11008) REF_CASE(FN, 12288) REF_CASE(FN, 13824) REF_CASE(FN, 14336) REF_CASE(FN, 18944)       \
    default: return -1; /* the reference throws on any other K (bindings.cpp:145-147) */                      \
  }                                                                                                           \
  return (int)cudaGetLastError();

// mode 0: reorder_quantize_x, 1: reorder_quantize_w, 2: reorder_quantize_w4.  Device pointers; default stream
// (the reference launches on the legacy default stream, reorder.cu:455).
extern "C" int ref_reorder_quantize(int mode, void* x, int rows, int16_t* idx, int KN, int KS, int KO, uint8_t* qn,
                                    uint8_t* qs, uint8_t* qo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo) {
  if (mode == 0) { REF_SWITCH(run_reorder_quantize_x) }
  if (mode == 1) { REF_SWITCH(run_reorder_quantize_w) }
  if (mode == 2) { REF_SWITCH(run_reorder_quantize_w4) }
  return -2;
}
