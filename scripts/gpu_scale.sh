#!/bin/bash
# scripts/gpu_scale.sh <ngpus> [tag] -- on an N-GPU box: the driver's scaling invocation (20 steps) and a long run for K1, and the
# 3840x2160 sequence workload (BASELINE.json configs[4]) frame-sharded over the N GPUs.  Outputs under gpurun_out/.
N=$1; tag=${2:-r2}; O=gpurun_out
run() { # name, bench args...
  name=$1; shift
  if [ $N -gt 1 ]; then
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N "$@" 2> $O/${name}.err | tail -1 > $O/${name}.json
  else
    timeout 400 python bench.py --gpus 1 "$@" 2> $O/${name}.err | tail -1 > $O/${name}.json
  fi
  python -c "
import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],4), 'parity', d['parity_checked'], d['ms_per_step_per_rank'])" $O/${name}.json
}
run bench_${tag}_k1_${N}gpu_20steps --steps 20 --warmup 5 --no-extras --no-cpu-baseline
run bench_${tag}_k1_${N}gpu --steps 1000 --warmup 20 --no-extras --no-cpu-baseline
run bench_${tag}_k4_${N}gpu --workload k4 --steps 300 --warmup 10 --no-extras --no-cpu-baseline
