// ref_sz_literal.h -- TEST INFRASTRUCTURE.  Force-included (-include) when compiling the reference's
// mgemm/src/reorder.cu in place: that file does not compile as shipped because of a typo,
// `#define FP6_MAX 28sz` (reorder.cu:18), which the compiler parses as a user-defined literal.
// Supplying the literal operator makes `28sz` evaluate to 28 without touching or copying the source.
#pragma once
__host__ __device__ constexpr int operator""sz(unsigned long long v) { return static_cast<int>(v); }
