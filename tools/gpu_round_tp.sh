#!/bin/bash
# Multi-GPU pass (under `gpurun --gpus N`): fused GEMM->all-reduce parity + timing against NCCL, then the TP bench both ways.
# usage: bash tools/gpu_round_tp.sh <tag> <N> [tokens]
TAG=${1:-tp}
N=${2:-2}
TOK=${3:-8192}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tools/tp_fused_check.py --tokens $TOK > $OUT/fused_check_tp$N.log 2>&1
echo "fused_check rc=$?" >> $OUT/fused_check_tp$N.log
timeout 300 $TR --master-port 29512 bench.py --gpus $N --steps 30 --warmup 5 --tokens $TOK --tp-reduce nccl > $OUT/bench_tp${N}_nccl.json 2> $OUT/bench_tp${N}_nccl.err
timeout 300 $TR --master-port 29513 bench.py --gpus $N --steps 30 --warmup 5 --tokens $TOK --tp-reduce fused > $OUT/bench_tp${N}_fused.json 2> $OUT/bench_tp${N}_fused.err
grep -h "^{" $OUT/fused_check_tp$N.log | cut -c1-400
for f in $OUT/bench_tp${N}_nccl.json $OUT/bench_tp${N}_fused.json; do cut -c1-200 $f; done
tail -3 $OUT/bench_tp${N}_fused.err
