#!/bin/bash
# e2e leg with and without NUMA binding: usage (under `gpurun --gpus N`): bash tools/gpu_e2e_numa.sh <tag> <N>
TAG=${1:-e2e}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
{ nproc; cat /sys/devices/system/node/online; for n in /sys/devices/system/node/node*; do echo "$n $(cat $n/cpulist)"; done
  nvidia-smi topo -m; python -c "import os; print(sorted(os.sched_getaffinity(0)))"; } > $OUT/host_topology.txt 2>&1
timeout 250 $TR --master-port 29731 bench.py --gpus $N --steps 20 --warmup 4 --tp-mode sp --no-prefill --no-cpu-baseline \
  > $OUT/bench_bind_tp$N.json 2> $OUT/bench_bind_tp$N.err
echo "bind rc=$?" >> $OUT/rc.txt
timeout 250 $TR --master-port 29732 bench.py --gpus $N --steps 20 --warmup 4 --tp-mode sp --no-prefill --no-cpu-baseline --no-numa-bind \
  > $OUT/bench_nobind_tp$N.json 2> $OUT/bench_nobind_tp$N.err
echo "nobind rc=$?" >> $OUT/rc.txt
cat $OUT/rc.txt; cat $OUT/host_topology.txt | head -40
for f in $OUT/*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
for l in open(sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']
        print(d.get('host_numa'), d['ms_per_step'], e['ms_per_step'], e.get('copies_only_ms'), e.get('bound'), e.get('pcie_gbps_per_rank'))
PY
done
