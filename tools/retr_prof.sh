#!/bin/bash
out=gpurun_out/retr; mkdir -p $out
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/launches_retrieval.csv python tools/profile_step.py --what retrieval > $out/ncu_lr.log 2>&1; echo "launches rc=$?"
python tools/summarize_launches.py $out/launches_retrieval.csv 2>/dev/null | head -7
timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:'score_topk' -o $out/topk_full -f python tools/profile_step.py --what retrieval > $out/ncu_ft.log 2>&1; echo "full rc=$?"
python tools/ncu_summary.py $out/topk_full.ncu-rep 2>&1 | cut -c1-400
