#!/bin/bash
# usage: tools/ab_env.sh VAR v1 v2 ...   -- time the c2 step under each value of an env knob, on one box
var=$1; shift
for v in "$@"; do
  env $var=$v python bench.py --skip-cpu --skip-retrieval --steps 50 --warmup 5 2>/dev/null | tail -1 | \
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$var=$v', round(d['ms_per_step'],4), 'ms/step  e2e', round(d['e2e']['value']))"
done
