#!/bin/bash
# TP schedule sweep (under `gpurun --gpus N`): token chunks / CUDA graph / GEMM grid cap / NCCL CTA cap.
# usage: bash tools/gpu_tp_sweep.sh <tag> <N> "<variant>;<variant>;..."   variant = "ENV=.. -- bench flags"
TAG=${1:-tpsweep}
N=${2:-2}
VARS=${3:-"-- --tp-chunks 1 --no-graph;-- --tp-chunks 1;-- --tp-chunks 2;-- --tp-chunks 2 --gemm-ctas 132;-- --tp-chunks 4"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
IFS=';' read -ra VV <<< "$VARS"
for v in "${VV[@]}"; do
  envs="${v%%--*}"; flags="${v#*--}"
  port=$((29600 + i))
  env $envs timeout 200 $TR --master-port $port bench.py --gpus $N --steps 30 --warmup 5 --no-e2e --no-cpu-baseline $flags \
    > $OUT/v$i.json 2> $OUT/v$i.err
  echo "rc=$? [$envs] [$flags]" >> $OUT/summary.txt
  python - $OUT/v$i.json >> $OUT/summary.txt <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    pl = d["per_linear"]
    print("   value %.0f TFLOP/s  ms/step %.3f  tp=%s  kernels: %s" % (d["value"], d["ms_per_step"], d.get("tp"),
          " ".join("%s q%.0f g%.0f" % (k, v["quant_us"], v["gemm_us"]) for k, v in pl.items())))
except Exception as e:
    print("   no line:", e)
PY
  i=$((i + 1))
done
cat $OUT/summary.txt
