#!/bin/bash
# ncu --set full captures of the dominant kernels of one eager training step (tools/profile_step.py), one kernel family per
# invocation.  Run on the GPU box:  bash tools/ncu_captures.sh   -> gpurun_out/ncu_*.ncu-rep (read here with
# `ncu -i ... --page raw --csv`, tools/ncu_hot.py).  The side stream is switched off so that the launch order is the single-stream
# order of profiles/r01_ncu_launches_step_v14.csv; the launch-skip indices select, among the launches whose name matches, the
# instance named in the comment (tools/join_launches.py prints the order).
set -u
export HCP_SIDE_STREAM=0
mkdir -p gpurun_out
NCU="ncu --profile-from-start off --set full --clock-control none --import-source on"
run() {  # name, regex, skip, count
    timeout 300 $NCU -k "regex:$2" --launch-skip "$3" --launch-count "$4" -f -o "gpurun_out/ncu_$1" python tools/profile_step.py > "gpurun_out/ncu_$1.log" 2>&1
    echo "$1 rc=$?"
}
run gemm_first14 'gemm_tc_kernel' 0 14        # #0 conv 320->320 @64x64 (MSUB 2) ... #6 linear M16384 K320+8 N320 ... #13 linear M16384 K320 N2560
run conv8x8 'gemm_tc_kernel' 101 1            # conv 1280->1280 @8x8, split-K
run attn_fwd 'attn_fwd2_kernel' 0 1           # self-attention L4096 d40
run attn_bwd 'attn_bwd_kernel' 1 1            # self-attention L4096 d40 backward
run gnf_fwd 'gnf_kernel' 0 1                  # single-pass GroupNorm forward C320 HW4096 (first launch of the step)
run gnf_bwd 'gnf_kernel' 57 1                 # single-pass GroupNorm backward C320 HW4096 (the 57 forward launches come first)
run lora_grad 'lora_grad_tc_kernel' 0 1       # LoRA gradients of the first group of the backward
