#!/bin/bash
# reducer / gather tuning at tp = N: usage bash tools/gpu_rs_sweep.sh <tag> <N>
TAG=$1; N=$2; OUT=gpurun_out/$TAG; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
python -m pytest tests/test_tp_fused_gpu.py -q -m gpu -x -k "gather or reduce_scatter or back_to_back or matches_sum" > $OUT/pytest.log 2>&1; tail -3 $OUT/pytest.log
run() { # name, env, extra args
  name=$1; shift
  timeout 150 $TR --master-port 29720 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline "$@" \
    > $OUT/$name.json 2> $OUT/$name.err
  python - $OUT/$name.json $name <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    pl=d['per_linear']
    print(sys.argv[2], d['tp']['mode'], d['tp']['fused_mode'], 'ms/step', round(d['ms_per_step'],3), 'parity', d['tp_parity']['ok'], 'status', d.get('tp_fused_status'),
          ' '.join(f"{k}:{v['quant_us']:.0f}/{v['gemm_us']:.0f}" for k,v in pl.items()), 'e2e', d.get('e2e') and round(d['e2e']['ms_per_step'],2))
    if 'tp_modes_measured' in d: print('   modes', {m:(round(v['ms_per_step'],3), v['e2e_ms_per_step'] and round(v['e2e_ms_per_step'],2)) for m,v in d['tp_modes_measured'].items()})
except Exception as e:
    print(sys.argv[2], 'ERR', e)
PY
}
run auto
MMX_TP_MODE=switch run switch_sp --tp-mode sp --no-e2e
MMX_TP_MODE=switch run switch_ar --tp-mode ar --no-e2e
timeout 100 $TR --master-port 29731 tools/tp_gather_probe.py --N 768 2>$OUT/gp.err | grep "^{"
