#!/bin/bash
# usage: tools/ab_env_dist.sh NGPU VAR v1 v2 ...  -- time the data-parallel c2 step under each value of an env knob
n=$1; var=$2; shift; shift
port=29600
for v in "$@"; do
  port=$((port+1))
  env $var=$v timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $n --skip-cpu --skip-retrieval --steps 50 --warmup 5 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$var=$v', 'n=$n', round(d['ms_per_step'],4), 'ms/step  value', round(d['value']), ' e2e', round(d['e2e']['value']))"
done
