#!/bin/bash
# GEMM CTA-pair mode: parity tests, then the c2 bench with the mode off (0), selective (1, default) and forced (2).
out=gpurun_out/pair; mkdir -p $out
timeout 400 python -m pytest tests -m gpu -q > $out/pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -6 $out/pytest_all.log
for P in ${MODES:-0 1}; do
  TT_GEMM_PAIR=$P timeout 300 python bench.py --steps 40 --warmup 5 --skip-cpu --skip-retrieval > $out/bench_pair$P.json 2> $out/bench_pair$P.err; echo "bench pair=$P rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/pair/bench_pair$P.json'))
r=d['roofline']
print('pair=$P', 'ms_per_step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'gemm frac', round(r['frac'],3), 'us', round(r['us_per_step'],1), 'all gemm us', round(r['all_gemm_launches']['us_per_step'],1))
PY
done
