#!/bin/bash
# Copies the evidence of tools/final_run.sh (gpurun_out/final/) into profiles/ under this round's names.
r=${1:-r02}; f=gpurun_out/final; p=profiles
cp $f/bench.json $p/${r}_bench_1gpu.json
cp $f/bench_reference.json $p/${r}_bench_reference.json
cp $f/launches_train.csv $p/${r}_launches_train.csv
cp $f/launches_retrieval.csv $p/${r}_launches_retrieval.csv
cp $f/launches_retrieval_shard8.csv $p/${r}_launches_retrieval_shard8.csv
for w in train retrieval retrieval_shard8; do python tools/summarize_launches.py $p/${r}_launches_$w.csv > $p/${r}_launches_$w.summary.txt; done
cp $f/ncu_full_gemm.summary.txt $p/${r}_ncu_full_gemm.summary.txt
cp $f/ncu_full_topk.summary.txt $p/${r}_ncu_full_topk.summary.txt
cp $f/ncu_full_attn.summary.txt $p/${r}_ncu_full_attn.summary.txt
cp $f/ncu_stalls_attn_bwd.txt $p/${r}_ncu_stalls_attn_bwd.txt
cp $f/ncu_stalls_attn_fwd.txt $p/${r}_ncu_stalls_attn_fwd.txt
cp $f/attention_times.txt $p/${r}_attention_times.txt
cp $f/gemm_dram_traffic.json $p/gemm_dram_traffic.json
cp $f/gemm_log.txt $p/${r}_gemm_per_launch.txt
grep -v "^/opt\|_warn" $f/trace.txt > $p/${r}_step_timeline_1gpu.txt
cp gpurun_out/parity_report.jsonl $p/${r}_parity_report.jsonl 2>/dev/null
ls -la $p | grep ${r}_ | wc -l
