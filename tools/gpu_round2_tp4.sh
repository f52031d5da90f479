#!/bin/bash
# 4-GPU pass: layer parity, bench (all three TP forms), NCCL arm, Qwen2.5-32B layer in the four forms
TAG=${1:-tp4}; N=4; OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29711 tools/tp_layer_check.py --tokens 4096 > $OUT/layer_check_tp$N.json 2> $OUT/layer_check_tp$N.err; echo "layer_check rc=$?" >> $OUT/rc.txt
timeout 300 $TR --master-port 29712 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench_tp$N.json 2> $OUT/bench_tp$N.err; echo "bench rc=$?" >> $OUT/rc.txt
timeout 200 $TR --master-port 29713 bench.py --gpus $N --steps 30 --warmup 5 --tp-reduce nccl --no-e2e --no-cpu-baseline > $OUT/bench_nccl_tp$N.json 2> $OUT/bench_nccl_tp$N.err; echo "bench nccl rc=$?" >> $OUT/rc.txt
for v in "ar --fused" "sp --sp" "tpr --tpr" "nccl --fused --tp-reduce nccl"; do
  set -- $v; name=$1; shift
  timeout 200 $TR --master-port 29714 tools/bench_models.py qwen_tp --iters 3 "$@" > $OUT/qwen_${name}_tp$N.json 2> $OUT/qwen_${name}_tp$N.err; echo "qwen $name rc=$?" >> $OUT/rc.txt
done
cat $OUT/rc.txt
for f in $OUT/*.json; do echo "== $f"; grep "^{" $f | cut -c1-400; done
