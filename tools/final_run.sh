#!/bin/bash
# One-GPU verification + evidence run: parity tests, smoke, bench (both arms), ncu launch lists and
# full captures of the dominant kernels. Everything lands under gpurun_out/final/. Each stage has its own
# timeout; a profiler stage runs only after the same command exited 0 without the profiler.
out=gpurun_out/final; mkdir -p $out
timeout 400 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/status.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/status.txt
timeout 600 python bench.py --steps 50 --warmup 5 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?" | tee -a $out/status.txt
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err; echo "bench-ref rc=$?" | tee -a $out/status.txt
timeout 200 python tools/profile_step.py > /dev/null 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/launches_train.csv python tools/profile_step.py > $out/ncu_lt.log 2>&1; echo "launches-train rc=$?" | tee -a $out/status.txt
timeout 200 python tools/profile_step.py --what retrieval > /dev/null 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/launches_retrieval.csv python tools/profile_step.py --what retrieval > $out/ncu_lr.log 2>&1; echo "launches-retrieval rc=$?" | tee -a $out/status.txt
if [ "${FULL:-1}" = "1" ]; then
timeout 500 ncu --set full --clock-control none --profile-from-start off -k regex:gemm_bf16 -o $out/gemm_full -f python tools/profile_step.py > $out/ncu_fg.log 2>&1; echo "full-gemm rc=$?" | tee -a $out/status.txt
timeout 300 ncu --set full --clock-control none --profile-from-start off -k regex:'score_topk|topk_finalize|sample_threshold' -o $out/topk_full -f python tools/profile_step.py --what retrieval > $out/ncu_ft.log 2>&1; echo "full-topk rc=$?" | tee -a $out/status.txt
# .ncu-rep files are too large to travel back: summarise on the box, keep the text
python tools/ncu_summary.py $out/gemm_full.ncu-rep > $out/ncu_full_gemm.summary.txt 2>&1
python tools/ncu_summary.py --gemm-traffic $out/gemm_full.ncu-rep $out/gemm_dram_traffic.json > /dev/null 2>&1
python tools/ncu_summary.py $out/topk_full.ncu-rep > $out/ncu_full_topk.summary.txt 2>&1
ncu -i $out/gemm_full.ncu-rep --page raw --csv 2>/dev/null | cut -d, -f1-60 | head -60 > $out/ncu_full_gemm.raw.head.csv
rm -f $out/gemm_full.ncu-rep $out/topk_full.ncu-rep
# attention (persistent forward / backward): full capture with source, summarised on the box
timeout 100 python tools/time_attention.py > $out/attention_times.txt 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:'attn_(fwd|bwd)_persist' -s 6 -c 8 -o $out/attn_full -f python tools/time_attention.py > $out/ncu_fa.log 2>&1; echo "full-attn rc=$?" | tee -a $out/status.txt
python tools/ncu_summary.py $out/attn_full.ncu-rep > $out/ncu_full_attn.summary.txt 2>&1
python tools/ncu_stalls.py $out/attn_full.ncu-rep attn_bwd_persist > $out/ncu_stalls_attn_bwd.txt 2>&1
python tools/ncu_stalls.py $out/attn_full.ncu-rep attn_fwd_persist > $out/ncu_stalls_attn_fwd.txt 2>&1
rm -f $out/attn_full.ncu-rep
fi   # FULL=0 skips the --set full captures (the slow part)
timeout 200 python tools/profile_step.py --what shard > /dev/null 2>&1 && timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/launches_retrieval_shard8.csv python tools/profile_step.py --what shard > $out/ncu_ls.log 2>&1; echo "launches-shard rc=$?" | tee -a $out/status.txt
timeout 200 python tools/gemm_log.py > $out/gemm_log.txt 2>&1
timeout 200 python tools/trace_step.py > $out/trace.txt 2>&1
tail -3 $out/pytest_gpu.log; cut -c1-300 $out/bench.json; cut -c1-300 $out/bench_reference.json
