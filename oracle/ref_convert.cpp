// ref_convert.cpp -- TEST INFRASTRUCTURE.  Runs the REFERENCE's own element converters and scale-factor layout
// (vendored NVIDIA/cutlass @ a1aaf230 under /root/reference/cutlass, used by mgemm/src/reorder.cu:138-143 and
// mgemm/include/reorder.cuh:120-125) on the host and dumps their outputs, so the CPU oracle can be pinned to them.
//
// Built only where /root/reference exists (oracle/Makefile -> oracle/_ref/ref_convert); the dumped table is
// committed as tests/golden/cutlass_convert_table.npz by tools/make_golden_cutlass.py.
//
// Output (binary, little endian) to argv[1]:
//   u8 e2m1[65536], u8 e3m2[65536], u8 e4m3[65536], u8 ue8m0[65536]    -- indexed by the bf16 bit pattern of the input
//   i64 sf_off[M*G]  for M=300, K=1024 (G=32), row-major (r,g)          -- CuTe SFA layout offsets
//   i64 sfb_off[N*G] for N=384, K=640  (G=20)                           -- CuTe SFB layout offsets
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "cute/tensor.hpp"
#include "cutlass/cutlass.h"
#include "cutlass/detail/sm100_blockscaled_layout.hpp"
#include "cutlass/numeric_conversion.h"
#include "cutlass/numeric_types.h"

using namespace cute;

static float bf16_bits_to_float(uint16_t h) {
  uint32_t u = uint32_t(h) << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* fp = std::fopen(argv[1], "wb");
  if (!fp) return 3;
  cutlass::NumericConverter<cutlass::float_e2m1_t, float, cutlass::FloatRoundStyle::round_to_nearest> c4;
  cutlass::NumericConverter<cutlass::float_e3m2_t, float, cutlass::FloatRoundStyle::round_to_nearest> c6;
  cutlass::NumericConverter<cutlass::float_e4m3_t, float, cutlass::FloatRoundStyle::round_to_nearest> c8;
  cutlass::NumericConverter<cutlass::float_ue8m0_t, float, cutlass::FloatRoundStyle::round_to_nearest> csf;
  std::vector<uint8_t> t4(65536), t6(65536), t8(65536), tsf(65536);
  for (uint32_t i = 0; i < 65536; ++i) {
    float f = bf16_bits_to_float(uint16_t(i));
    t4[i] = c4(f).storage;
    t6[i] = c6(f).storage;
    t8[i] = c8(f).storage;
    tsf[i] = csf(f).storage;
  }
  std::fwrite(t4.data(), 1, 65536, fp);
  std::fwrite(t6.data(), 1, 65536, fp);
  std::fwrite(t8.data(), 1, 65536, fp);
  std::fwrite(tsf.data(), 1, 65536, fp);

  using Cfg = cutlass::detail::Sm1xxBlockScaledConfig<32>;
  {
    int M = 300, K = 1024;
    auto layout = filter_zeros(Cfg::tile_atom_to_shape_SFA(make_shape(M, 128, K, 1)));
    std::vector<int64_t> off(size_t(M) * (K / 32));
    for (int r = 0; r < M; ++r)
      for (int g = 0; g < K / 32; ++g) {
        // the coordinate the reference kernel builds, reorder.cu:182-185
        auto c0 = make_coord(make_coord(r % 32, (r / 32) % 4), r / 128);
        auto c1 = make_coord(make_coord(0, g % 4), g / 4);
        auto c2 = make_coord(0, 0);
        off[size_t(r) * (K / 32) + g] = int64_t(layout(make_coord(c0, c1, c2)));
      }
    std::fwrite(off.data(), 8, off.size(), fp);
  }
  {
    int N = 384, K = 640;
    auto layout = filter_zeros(Cfg::tile_atom_to_shape_SFB(make_shape(128, N, K, 1)));
    std::vector<int64_t> off(size_t(N) * (K / 32));
    for (int r = 0; r < N; ++r)
      for (int g = 0; g < K / 32; ++g) {
        auto c0 = make_coord(make_coord(r % 32, (r / 32) % 4), r / 128);
        auto c1 = make_coord(make_coord(0, g % 4), g / 4);
        auto c2 = make_coord(0, 0);
        off[size_t(r) * (K / 32) + g] = int64_t(layout(make_coord(c0, c1, c2)));
      }
    std::fwrite(off.data(), 8, off.size(), fp);
  }
  std::fclose(fp);
  return 0;
}
