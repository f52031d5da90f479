#!/bin/bash
# Round-2 multi-GPU measurements: usage (under `gpurun --gpus N`): bash tools/gpu_round2_tp.sh <tag> <N> [quick]
TAG=${1:-tp}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi -L > $OUT/gpus.txt
# 1. real-rank parity: layers (NCCL / fused all-reduce / sequence parallel / expert parallel) against the 1-GPU layer
timeout 200 $TR --master-port 29711 tools/tp_layer_check.py --tokens 4096 > $OUT/layer_check_tp$N.json 2> $OUT/layer_check_tp$N.err
echo "layer_check rc=$?" >> $OUT/rc.txt
# 2. the bench step: default flags (measures the all-reduce and the sequence-parallel form, reports the faster), then NCCL
timeout 300 $TR --master-port 29712 bench.py --gpus $N --steps 30 --warmup 5 > $OUT/bench_tp$N.json 2> $OUT/bench_tp$N.err
echo "bench rc=$?" >> $OUT/rc.txt
timeout 200 $TR --master-port 29713 bench.py --gpus $N --steps 30 --warmup 5 --tp-reduce nccl --no-e2e --no-cpu-baseline \
  > $OUT/bench_nccl_tp$N.json 2> $OUT/bench_nccl_tp$N.err
echo "bench nccl rc=$?" >> $OUT/rc.txt
if [ "$3" != "quick" ]; then
  # 3. BASELINE config 4: Qwen2.5-32B-shaped layer -- fused all-reduce, sequence parallel, NCCL
  timeout 200 $TR --master-port 29714 tools/bench_models.py qwen_tp --iters 3 --fused > $OUT/qwen_ar_tp$N.json 2> $OUT/qwen_ar_tp$N.err
  echo "qwen ar rc=$?" >> $OUT/rc.txt
  timeout 200 $TR --master-port 29715 tools/bench_models.py qwen_tp --iters 3 --sp > $OUT/qwen_sp_tp$N.json 2> $OUT/qwen_sp_tp$N.err
  echo "qwen sp rc=$?" >> $OUT/rc.txt
  timeout 200 $TR --master-port 29716 tools/bench_models.py qwen_tp --iters 3 --fused --tp-reduce nccl > $OUT/qwen_nccl_tp$N.json 2> $OUT/qwen_nccl_tp$N.err
  echo "qwen nccl rc=$?" >> $OUT/rc.txt
  timeout 200 $TR --master-port 29719 tools/bench_models.py qwen_tp --iters 3 --tpr > $OUT/qwen_tpr_tp$N.json 2> $OUT/qwen_tpr_tp$N.err
  echo "qwen tpr rc=$?" >> $OUT/rc.txt
  # 4. BASELINE config 5: Mixtral 8x7B expert FFN, expert parallel N: grouped path (fused SiLU*up) and the expert loop
  timeout 200 $TR --master-port 29717 tools/bench_models.py mixtral_ep --iters 3 --fused > $OUT/mixtral_grouped_ep$N.json 2> $OUT/mixtral_grouped_ep$N.err
  echo "mixtral grouped rc=$?" >> $OUT/rc.txt
  timeout 200 $TR --master-port 29718 tools/bench_models.py mixtral_ep --iters 3 --fused --loop > $OUT/mixtral_loop_ep$N.json 2> $OUT/mixtral_loop_ep$N.err
  echo "mixtral loop rc=$?" >> $OUT/rc.txt
fi
cat $OUT/rc.txt
for f in $OUT/*.json; do echo "== $f"; grep "^{" $f | cut -c1-500; done
