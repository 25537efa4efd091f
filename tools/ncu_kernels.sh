#!/bin/bash
# Development aid (run on the GPU box through gpurun): one `ncu --set full` capture per hot kernel of a guided step at the bench
# shape, plus the launch list of a short bench run.  Reports land in gpurun_out/ (tag = $1).
#   gpurun -- 'bash tools/ncu_kernels.sh r2a'
tag=${1:-r2}
mkdir -p gpurun_out
for spec in "bwd:tc_pred_edge_bwd_kernel" "fwd:tc_pred_edge_fwd_kernel" "den:tc_den_edge_kernel" "lin:tc_lin_kernel" "red:pred_bwd_reduce_kernel"; do
    name=${spec%%:*}; kern=${spec##*:}
    cnt=1; [ "$name" = "den" ] && cnt=2
    timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$kern" -s 20 -c $cnt -f -o gpurun_out/prof_${tag}_${name} \
        python bench.py --profile-only --steps 2 --warmup 3 > gpurun_out/prof_${tag}_${name}.log 2>&1
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --profile-only --steps 2 --warmup 3 > gpurun_out/launches_${tag}.log 2>&1
ls -la gpurun_out/prof_${tag}_*.ncu-rep gpurun_out/launches_${tag}.csv
