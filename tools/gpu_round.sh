#!/bin/bash
# One GPU-box pass: parity tests, bench line, ncu launch list, ncu --set full of the two hot kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [tests|notests]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > $OUT/gpu.txt
if [ "${2:-tests}" = "tests" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/pytest.log
fi
timeout 600 python bench.py --steps 30 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2>> $OUT/bench.err
timeout 200 python tools/m_sweep.py > $OUT/m_sweep.log 2>&1   # BASELINE config 2: M = 1 .. 8192 per linear
# launch list of the bench command itself (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
# full capture of the dominant kernels on the same workload (M=8192): gate_up GEMM, down quantize
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mixed_gemm -s 9 -c 4 -o $OUT/prof_gemm \
  python tools/prof_workload.py 8192 4 > $OUT/prof_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:reorder_quantize -s 16 -c 4 -o $OUT/prof_quant \
  python tools/prof_workload.py 8192 4 > $OUT/prof_quant.log 2>&1
tail -3 $OUT/pytest.log; cat $OUT/bench.json | cut -c1-600
