#!/bin/bash
# Measurements this round ran out of GPU minutes for -- the first multi-GPU call of the next round.
# usage (under `gpurun --gpus N`, N = 4 or 8): bash tools/gpu_round_next.sh <tag> <N>
TAG=${1:-next}
N=${2:-4}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
# 1. push vs switch data path of the fused all-reduce at this tp (round 1 measured switch only at tp 4 / 8)
for mode in push switch; do
  MMX_TP_MODE=$mode timeout 200 $TR --master-port 29701 tools/tp_fused_check.py --iters 20 > $OUT/check_${mode}_tp$N.log 2>&1
  MMX_TP_MODE=$mode timeout 200 $TR --master-port 29702 bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --no-cpu-baseline \
    > $OUT/bench_${mode}_tp$N.json 2> $OUT/bench_${mode}_tp$N.err
done
timeout 200 $TR --master-port 29703 bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --tp-reduce nccl \
  > $OUT/bench_nccl_tp$N.json 2> $OUT/bench_nccl_tp$N.err
# 2. BASELINE config 4 (Qwen2.5-32B layer, fused vs NCCL) and config 5 (Mixtral experts, expert parallel)
for red in fused nccl; do
  timeout 200 $TR --master-port 29704 tools/bench_models.py qwen_tp --iters 3 --tp-reduce $red > $OUT/qwen_${red}_tp$N.json 2> $OUT/qwen_${red}_tp$N.err
done
timeout 200 $TR --master-port 29705 tools/bench_models.py mixtral_ep --iters 3 --fused > $OUT/mixtral_ep$N.json 2> $OUT/mixtral_ep$N.err
# 3. one GPU: quantize at M large enough to defeat L2 residency (the >= 80 % of HBM target), with a committed log
timeout 200 python tools/quant_sweep.py --no-parity --shapes 16384x4096,32768x4096,65536x4096,32768x14336 > $OUT/quant_sweep_largeM.log 2>&1
grep -h "^{" $OUT/*.json $OUT/check_*.log 2>/dev/null | cut -c1-300
