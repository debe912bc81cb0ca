#!/bin/bash
# ncu --set full captures of the dominant kernels of one eager training step (tools/profile_step.py), one kernel per invocation.
# Run on the GPU box:  bash tools/ncu_captures.sh   -> gpurun_out/ncu_*.ncu-rep (read here with `ncu -i ... --page raw --csv`).
# The launch-skip indices select, among the launches whose name matches, the instance named in the comment (they follow the
# launch order of the step: see profiles/r01_step_time_by_shape_v7.txt / tools/join_launches.py).
set -u
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on"
run() {  # name, regex, skip, count
    timeout 300 $NCU -k "regex:$2" --launch-skip "$3" --launch-count "$4" -f -o "gpurun_out/ncu_$1" python tools/profile_step.py > "gpurun_out/ncu_$1.log" 2>&1
    echo "$1 rc=$?"
}
run gemm_first8 'gemm_tc_kernel' 0 8          # #0 conv 320->320 @64x64 (MSUB 2) ... #6 linear M16384 K320+8 N320
run conv8x8 'gemm_tc_kernel' 101 1            # conv 1280->1280 @8x8, split-K
run attn_fwd 'attn_fwd2_kernel' 0 1           # self-attention L4096 d40
run attn_bwd 'attn_bwd_kernel' 1 1            # self-attention L4096 d40 backward
HCP_GN_TWO_PASS=1 run gn 'gn8_' 0 2           # two-pass GroupNorm C320 HW4096: partial + apply
run gnf 'gnf_kernel' 0 2                      # single-pass GroupNorm (cluster) C320 HW4096: first two launches
