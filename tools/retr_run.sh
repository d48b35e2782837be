#!/bin/bash
# Retrieval-only check: parity tests, a short bench, launch list, one full capture of score_topk.
out=gpurun_out/retr; mkdir -p $out
timeout 300 python -m pytest tests/test_retrieval.py tests/test_api.py -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?" | tee $out/status.txt
tail -3 $out/pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --skip-cpu > $out/bench.json 2> $out/bench.err; echo "bench rc=$?" | tee -a $out/status.txt
python - <<'PY'
import json
d=json.load(open('gpurun_out/retr/bench.json'))
print(d['ms_per_step'], json.dumps(d['retrieval'])[:600])
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/launches_retrieval.csv python tools/profile_step.py --what retrieval > $out/ncu_lr.log 2>&1; echo "launches rc=$?" | tee -a $out/status.txt
python tools/summarize_launches.py $out/launches_retrieval.csv 2>/dev/null | head -8
timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:'score_topk|topk_finalize' -o $out/topk_full -f python tools/profile_step.py --what retrieval > $out/ncu_ft.log 2>&1; echo "full rc=$?" | tee -a $out/status.txt
python tools/ncu_summary.py $out/topk_full.ncu-rep > $out/ncu_full_topk.summary.txt 2>&1
ls -la $out
