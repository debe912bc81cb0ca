#!/bin/bash
# End-of-round validation on the GPU box: full GPU test suite, smoke(), the benchmark (configs 2 and 3), the in-graph op timings, the
# ncu launch list of one eager step and `--set full` captures of the two attention kernels.  Everything lands in gpurun_out/ with the
# prefix given as $1 (default r02_final); copy what should be judged into profiles/.
set -u
P=${1:-r02_final}
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/${P}_pytest_gpu.log; cat gpurun_out/${P}_pytest_gpu.log
timeout 80 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/bench.err; cut -c1-260 gpurun_out/${P}_bench.json
timeout 200 python bench.py --config 3 --no-cpu-baseline > gpurun_out/${P}_bench_config3.json 2> gpurun_out/bench3.err; cut -c1-260 gpurun_out/${P}_bench_config3.json; tail -2 gpurun_out/bench3.err
timeout 150 python tools/bench_ops.py > gpurun_out/${P}_bench_ops.txt 2>&1; tail -2 gpurun_out/${P}_bench_ops.txt
timeout 100 python tools/probe_attn_trace.py > gpurun_out/${P}_probe_attn_trace.txt 2>&1
timeout 150 python tools/bench_sdxl.py > gpurun_out/${P}_bench_sdxl_config4.json 2> gpurun_out/sdxl.err; cut -c1-200 gpurun_out/${P}_bench_sdxl_config4.json
export HCP_SIDE_STREAM=0
timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${P}_launches.csv \
    python tools/profile_step.py > gpurun_out/profile_step.log 2>&1; tail -1 gpurun_out/profile_step.log
cp gpurun_out/launch_shapes.json gpurun_out/${P}_launch_shapes.json
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on"
timeout 120 $NCU -k regex:attn_fwd2_kernel --launch-skip 0 --launch-count 1 -f -o gpurun_out/${P}_ncu_attn_fwd python tools/profile_step.py > gpurun_out/ncu_attn_fwd.log 2>&1; echo "ncu fwd rc=$?"
timeout 120 $NCU -k regex:attn_bwd_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/${P}_ncu_attn_bwd python tools/profile_step.py > gpurun_out/ncu_attn_bwd.log 2>&1; echo "ncu bwd rc=$?"
