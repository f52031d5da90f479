#!/bin/bash
# Round-2 multi-GPU measurements: usage (under `gpurun --gpus N`): bash tools/gpu_round2_tp.sh <tag> <N> [quick]
TAG=${1:-tp}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L > $OUT/gpus.txt
# 1. real-rank parity: layers (NCCL / fused all-reduce / sequence parallel / expert parallel) against the 1-GPU layer
timeout 300 $TR --master-port 29711 tools/tp_layer_check.py --tokens 1024 > $OUT/layer_check_tp$N.json 2> $OUT/layer_check_tp$N.err
echo "layer_check rc=$?" >> $OUT/rc.txt
# 2. the bench step: sequence parallel (default), all-reduce, NCCL
for mode in sp ar; do
  timeout 300 $TR --master-port 29712 bench.py --gpus $N --steps 30 --warmup 5 --tp-mode $mode --no-cpu-baseline \
    > $OUT/bench_${mode}_tp$N.json 2> $OUT/bench_${mode}_tp$N.err
  echo "bench $mode rc=$?" >> $OUT/rc.txt
done
if [ "$3" != "quick" ]; then
  timeout 300 $TR --master-port 29713 bench.py --gpus $N --steps 30 --warmup 5 --tp-reduce nccl --no-e2e --no-cpu-baseline \
    > $OUT/bench_nccl_tp$N.json 2> $OUT/bench_nccl_tp$N.err
  echo "bench nccl rc=$?" >> $OUT/rc.txt
  MMX_TP_MODE=switch timeout 300 $TR --master-port 29714 bench.py --gpus $N --steps 30 --warmup 5 --tp-mode ar --no-e2e --no-cpu-baseline \
    > $OUT/bench_ar_switch_tp$N.json 2> $OUT/bench_ar_switch_tp$N.err
  echo "bench ar switch rc=$?" >> $OUT/rc.txt
fi
cat $OUT/rc.txt
for f in $OUT/*.json; do echo "== $f"; cut -c1-400 $f; done
tail -5 $OUT/*.err | cut -c1-300
