OUT=gpurun_out/r2tpr2; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29711 tools/tp_layer_check.py --tokens 4096 > $OUT/layer_check.json 2> $OUT/layer_check.err; echo "layer rc=$?"
grep "^{" $OUT/layer_check.json | cut -c1-1500
timeout 300 $TR --master-port 29712 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2tpr2/bench.json'))
print('line mode', d['tp']['mode'], d['ms_per_step'], 'parity', d['tp_parity']['ok'])
print('modes', {m:(round(v['ms_per_step'],3), v['e2e_ms_per_step'] and round(v['e2e_ms_per_step'],2), v['tp_parity_ok'], v['per_linear_us']) for m,v in d['tp_modes_measured'].items()})
a=d.get('alt_token_parallel'); print('alt', a and (a['ms_per_step'], a['tp_parity']['ok'], a['per_linear_us'], a['e2e'] and a['e2e']['ms_per_step']))
PY
tail -3 $OUT/bench.err
timeout 200 $TR --master-port 29715 tools/bench_models.py qwen_tp --iters 3 --tpr > $OUT/qwen_tpr.json 2> $OUT/qwen_tpr.err; echo "qwen rc=$?"; grep "^{" $OUT/qwen_tpr.json | cut -c1-600
