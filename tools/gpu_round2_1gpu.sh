#!/bin/bash
# Round-2 one-GPU pass: parity tests, bench line (with the 32-layer prefill leg), ncu launch list + full capture of ONE bench
# step, quantize sweep at M large enough to defeat L2, config-2 M sweep, MX peak probe.
# usage (under gpurun): bash tools/gpu_round2_1gpu.sh <tag> [tests|notests]
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt
if [ "${2:-tests}" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
fi
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err
timeout 120 python tools/mx_peak.py > $OUT/mx_peak.json 2> $OUT/mx_peak.err
timeout 200 python tools/m_sweep.py > $OUT/m_sweep.log 2>&1   # BASELINE config 2: M = 1 .. 8192 per linear
timeout 200 python tools/quant_sweep.py --no-parity --shapes 8192x4096,8192x14336,16384x4096,32768x4096,16384x14336,65536x4096 > $OUT/quant_sweep.log 2>&1
# launch list of the bench command itself (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-prefill --no-moe --no-e2e > $OUT/bench_under_ncu.log 2>&1
# full capture of ONE bench step: weights' quantization (4 launches) + 3 warm-up steps (24) precede it
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"mixed_gemm|reorder_quantize" -s 28 -c 8 -o $OUT/prof_step \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-prefill --no-moe --no-e2e > $OUT/prof_step.log 2>&1
tail -3 $OUT/pytest.log; cut -c1-400 $OUT/bench.json; tail -8 $OUT/quant_sweep.log
