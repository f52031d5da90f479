// ref_activate_wrap.cu -- TEST INFRASTRUCTURE.  C-ABI shim over the REFERENCE's own launchers of the ops that
// quantize without a permutation (/root/reference/mgemm/src/activate.cu:510-632: run_activate_bf16_mixed,
// run_downproj_bf16_mixed, run_downproj_bf16_mxfp4; declared in mgemm/include/reorder.cuh).  Linked with the
// reference activate.cu compiled in place for sm_100a into oracle/_ref/libref_activate.so (oracle/Makefile).
// Used on the B200 box by tests (-m gpu) and tools/make_golden_ref.py as the bit-exact GPU oracle.
//
// The reference kernel calls __syncthreads() inside its FP6 branch (activate.cu:187,196); a 128-thread block covers
// 512 consecutive channels, so callers must keep KN and KN+KS multiples of 512 or the barrier is divergent.
#include <cstdint>
#include <cuda_runtime.h>
#include "cutlass/numeric_types.h"

typedef cutlass::float_ue8m0_t sf_t;
typedef cutlass::bfloat16_t bf16_t;

void run_activate_bf16_mixed(bf16_t*, bf16_t*, int, int, uint8_t*, uint8_t*, uint8_t*, sf_t*, sf_t*, sf_t*, int, int, int);
void run_downproj_bf16_mixed(bf16_t*, int, int, uint8_t*, uint8_t*, uint8_t*, sf_t*, sf_t*, sf_t*, int, int, int);
void run_downproj_bf16_mxfp4(bf16_t*, int, int, uint8_t*, uint8_t*, uint8_t*, sf_t*, sf_t*, sf_t*, int, int, int);

// mode 0: activate_quantize_x(a, b), 1: downproj_quantize_w(a), 2: downproj_quantize_w4(a).  Device pointers,
// legacy default stream (activate.cu:541).
extern "C" int ref_rowwise_quantize(int mode, void* a, void* b, int rows, int KN, int KS, int KO, uint8_t* qn, uint8_t* qs,
                                    uint8_t* qo, uint8_t* sfn, uint8_t* sfs, uint8_t* sfo) {
  if ((KN % 512) || ((KN + KS) % 512)) return -3;  // see the header comment
  const int K = KN + KS + KO;
  if (mode == 0) run_activate_bf16_mixed((bf16_t*)a, (bf16_t*)b, rows, K, qn, qs, qo, (sf_t*)sfn, (sf_t*)sfs, (sf_t*)sfo, KN, KS, KO);
  else if (mode == 1) run_downproj_bf16_mixed((bf16_t*)a, rows, K, qn, qs, qo, (sf_t*)sfn, (sf_t*)sfs, (sf_t*)sfo, KN, KS, KO);
  else if (mode == 2) run_downproj_bf16_mxfp4((bf16_t*)a, rows, K, qn, qs, qo, (sf_t*)sfn, (sf_t*)sfs, (sf_t*)sfo, KN, KS, KO);
  else return -2;
  return (int)cudaGetLastError();
}
