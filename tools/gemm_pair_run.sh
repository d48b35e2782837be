#!/bin/bash
# GEMM CTA-pair modes: parity tests under TT_GEMM_PAIR=$TESTMODE, then the c2 bench for each mode in $MODES
# (0 = off, 1 = deep-K launches only (default), 2 = every eligible launch).
out=gpurun_out/pair; mkdir -p $out
TT_GEMM_PAIR=${TESTMODE:-1} timeout 400 python -m pytest tests/test_gemm.py tests/test_engine.py -m gpu -q > $out/pytest.log 2>&1; echo "pytest (TT_GEMM_PAIR=${TESTMODE:-1}) rc=$?"; tail -6 $out/pytest.log
for P in ${MODES:-0 1}; do
  TT_GEMM_PAIR=$P timeout 300 python bench.py --steps 40 --warmup 5 --skip-cpu --skip-retrieval > $out/bench_pair$P.json 2> $out/bench_pair$P.err; echo "bench pair=$P rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/pair/bench_pair$P.json'))
r=d['roofline']
print('pair=$P', 'ms_per_step', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), 'gemm frac', round(r['frac'],3), 'us', round(r['us_per_step'],1), 'all gemm us', round(r['all_gemm_launches']['us_per_step'],1))
PY
done
if [ -n "$LOGMODE" ]; then TT_GEMM_PAIR=$LOGMODE timeout 200 python tools/gemm_log.py 2>&1 | grep "M= 51200\|total"; fi
