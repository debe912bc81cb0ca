#!/bin/bash
# End-of-round validation on the GPU box: full GPU test suite, smoke(), the benchmark, the in-graph op timings, the ncu launch
# list of one eager step and `--set full` captures of the two attention kernels.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
timeout 260 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/pytest_gpu_final.log; cat gpurun_out/pytest_gpu_final.log
timeout 80 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 150 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench.err; cut -c1-260 gpurun_out/bench_final.json
timeout 150 python tools/bench_ops.py > gpurun_out/bench_ops_final.txt 2>&1; tail -2 gpurun_out/bench_ops_final.txt
export HCP_SIDE_STREAM=0
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv \
    python tools/profile_step.py > gpurun_out/profile_step.log 2>&1; tail -1 gpurun_out/profile_step.log
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on"
timeout 120 $NCU -k regex:attn_fwd2_kernel --launch-skip 0 --launch-count 1 -f -o gpurun_out/ncu_attn_fwd_final python tools/profile_step.py > gpurun_out/ncu_attn_fwd_final.log 2>&1; echo "ncu fwd rc=$?"
timeout 120 $NCU -k regex:attn_bwd_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/ncu_attn_bwd_final python tools/profile_step.py > gpurun_out/ncu_attn_bwd_final.log 2>&1; echo "ncu bwd rc=$?"
