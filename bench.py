#!/usr/bin/env python
"""Benchmark of the hot path: SD1.5 LoRA (rank 8, every attn1/attn2 Linear) training step, 512x512 (64x64 latents),
batch 4 per GPU, bf16 kernels / fp32 master weights -- BASELINE.json `configs[1]`, metric "LoRA-train images/sec".

  python bench.py [--gpus N] [--steps K] [--warmup W]            product arm (B200 kernels)
  python bench.py --impl reference [--gpus N] ...                CPU arm: the oracle restatement of the reference path
                                                                 (diffusers UNet semantics + hcpdiff LoRA operator + train_ac
                                                                 step order) on the host cores -- the reference itself cannot
                                                                 run here (diffusers/accelerate/hydra are not installable)
For N > 1 launch with torchrun (one rank per GPU); rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

F_FWD = 803.27e9                 # algorithmic FLOP per image, forward (BASELINE.md section 3 / SURVEY.md App. B)
F_ATTN = 126.05e9
F_STEP = 2 * F_FWD + F_ATTN      # LoRA training step with frozen base: 1732.6 GFLOP / image
METRIC = "LoRA-train images/sec SD1.5 512px"
PER_GPU_BATCH = 4
LORA_RANK = 8


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1423.3), d.get("bf16_tflops", 1713.4), d.get("hbm_gbs", 6567.7), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


def usable_cores() -> int:
    """Host threads this process may really use: the scheduler affinity mask capped by the cgroup CPU quota (a box that
    shows 128 logical CPUs but grants a 16-CPU quota thrashes with 128 threads)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except (OSError, ValueError):
        pass
    return n


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); smax.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm (oracle port of the reference path)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference_steps(max_steps: int, warmup: int, budget_s: float):
    """One step = ONE image (B=1, 64x64 latent, 77 tokens) through the reference step order (train_ac.py:467-504): forward ->
    MSE(eps) -> backward -> clip 1.0 -> AdamW -> zero_grad, fp32, all host threads.  Steps stop early when `budget_s` is spent."""
    from oracle import unet_ref as U
    cores = usable_cores()
    torch.set_num_threads(cores)
    spec = U.SD15
    sd = U.init_params(spec)
    lora = U.init_lora(spec, rank=LORA_RANK, up_std=0.0)
    params = []
    for blocks in lora.values():
        for e in blocks:
            e.W_down.requires_grad_(True)
            e.W_up.requires_grad_(True)
            params += [e.W_down, e.W_up]
    opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-2)
    acp = U.ddpm_alphas_cumprod()
    lat, noise, t, ehs = U.synthetic_batch(1, spec, seed=1234)

    def one_step():
        x_t = U.add_noise(lat, noise, t, acp)
        pred = U.unet_forward(sd, x_t, t, ehs, lora=lora, spec=spec)
        loss = torch.nn.functional.mse_loss(pred.float(), noise.float(), reduction="none").mean()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        opt.zero_grad(set_to_none=False)
        return float(loss.detach())

    t_start = time.perf_counter()
    done_w = 0
    for _ in range(warmup):
        one_step()
        done_w += 1
        if time.perf_counter() - t_start > budget_s * 0.4:
            break
    t0 = time.perf_counter()
    done = 0
    while done < max_steps:
        one_step()
        done += 1
        if time.perf_counter() - t_start > budget_s:
            break
    dt = time.perf_counter() - t0
    return {"images_per_s": done / dt, "steps_run": done, "warmup_run": done_w, "seconds": dt, "cores": cores}


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    r = cpu_reference_steps(args.steps, args.warmup, budget_s=200.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["images_per_s"], "unit": "images/s", "n_gpus": args.gpus, "steps": r["steps_run"],
        "requested_steps": args.steps, "warmup": r["warmup_run"], "ms_per_step": 1e3 * r["seconds"] / r["steps_run"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SD1.5 UNet LoRA r=8 all attn1/attn2 Linear, 512x512 (64x64 latent), 77 tokens, train step",
                   "per_step_sample": "1 image (B=1): fwd + MSE + bwd + clip + AdamW", "global_batch": 1},
        "cpu_baseline": {"value": r["images_per_s"], "unit": "images/s", "cores": r["cores"], "kind": "port",
                         "sample": f"{r['steps_run']} timed step(s) of 1 image each after {r['warmup_run']} warm-up; oracle/unet_ref.py "
                                   "(restated diffusers UNet + reference LoRA operator), torch CPU fp32, time-bounded"},
        "e2e": {"value": r["images_per_s"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------------
# product arm
# ----------------------------------------------------------------------------------------------------------------------
def time_kernel(fn, iters=10):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / iters


def ncu_dram_traffic(summary="profiles/r01_ncu_attn_bwd_v23.summary.csv"):
    """DRAM bytes (read + write) of ONE launch of the dominant kernel, from the committed `ncu --set full` summary of the same
    kernel at the same shape (attention backward, B=4 H=8 L=4096 d=40).  None when the file is absent."""
    import csv
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), summary)
    try:
        rows = list(csv.reader(open(path)))
        hdr, unit, val = rows[0], rows[1], rows[2]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = 0.0
        for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(name)
            tot += float(val[i]) * scale[unit[i]]
        return {"bytes_per_launch": tot, "kernel": "attn_bwd_kernel B4 H8 L4096 d40", "source": summary}
    except (OSError, ValueError, KeyError, IndexError):
        return None


def dominant_kernel_roofline(peak_tflops):
    """Live CUDA-event timing of the kernels that dominate the step, at their benchmark shapes (B=4)."""
    from hcp_diffusion_b200 import ops
    B, H, L, d = PER_GPU_BATCH, 8, 4096, 40
    C = H * d
    qkv = torch.randn(B, L, 3 * C, device="cuda").to(torch.bfloat16).requires_grad_(True)
    o = ops.attention(H, C, (0, C, 2 * C), qkv)
    do = torch.randn_like(o)
    ms_f = time_kernel(lambda: ops.attention(H, C, (0, C, 2 * C), qkv.detach()))

    def bwd():
        oo = ops.attention(H, C, (0, C, 2 * C), qkv)
        oo.backward(do)
    ms_fb = time_kernel(bwd)
    ms_b = ms_fb - ms_f
    fl_f = 4.0 * B * H * L * L * d
    fl_b = 10.0 * B * H * L * L * d
    out = {
        "attn_fwd_L4096_d40": {"ms": ms_f, "tflops": fl_f / ms_f * 1e-9, "frac": fl_f / ms_f * 1e-9 / peak_tflops},
        "attn_bwd_L4096_d40": {"ms": ms_b, "tflops": fl_b / ms_b * 1e-9, "frac": fl_b / ms_b * 1e-9 / peak_tflops},
    }
    x = torch.randn(B, 4096, 320, device="cuda").to(torch.bfloat16)
    w = torch.randn(320, 320, 3, 3, device="cuda") * 0.02
    pack = ops.ConvPack(w, None, 1)
    ms_c = time_kernel(lambda: ops.conv3x3(pack, x, (B, 64, 64)))
    fl_c = 2.0 * B * 4096 * 320 * 9 * 320
    out["conv3x3_320_320_64x64"] = {"ms": ms_c, "tflops": fl_c / ms_c * 1e-9, "frac": fl_c / ms_c * 1e-9 / peak_tflops}
    return out


def run_product_arm(args, rank, world, local_rank):
    import torch.distributed as dist
    from hcp_diffusion_b200 import _lib
    from hcp_diffusion_b200.engine import LoraTrainStep
    from hcp_diffusion_b200.models import UNet2DConditionModel
    from hcp_diffusion_b200.utils.cfg_net_tools import make_hcpdiff
    from oracle import unet_ref as U   # used for the synthetic weight/input generators and the cpu_baseline leg only

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.check(_lib.lib().hcp_device_check(), "hcp_device_check")
    sustained, burst, hbm, peak_src = measured_peaks()

    spec = U.SD15
    unet = UNet2DConditionModel()
    unet.load_state_dict(U.init_params(spec, seed=0))
    unet = unet.to(dev).requires_grad_(False).eval()
    groups, lora = make_hcpdiff(unet, None, [{"lr": 1e-4, "rank": LORA_RANK, "alpha": 1.0, "dropout": 0.0, "layers": [r"re:.*\.attn.?$"]}])
    params = [p for g in groups for p in g["params"]]
    n_lora = sum(p.numel() for p in params)
    step = LoraTrainStep(unet, params, lr=1e-4, weight_decay=1e-2, max_grad_norm=1.0, use_cuda_graph=True)

    B = PER_GPU_BATCH
    lat, noise, t, ehs = U.synthetic_batch(B, spec, seed=1234 + rank)
    host = [x.pin_memory() for x in (lat, noise, t, ehs)]
    h2d = sum(x.numel() * x.element_size() for x in host)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    losses = []

    def e2e_step():
        loss = step.step(*host)
        losses.append(float(loss.cpu()))          # the per-step D2H read of the result (reference: loss.item(), train_ac.py:504)

    # warm-up (also captures the CUDA graphs)
    for _ in range(max(args.warmup, 3)):
        e2e_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _lib.launch_count
    ms_resident = timed(step.step_resident, args.steps)
    clocks = sampler.stop()
    ms_e2e = timed(e2e_step, args.steps)
    assert all(l == l and l < 1e4 for l in losses), "loss diverged / NaN"

    if rank != 0:
        return
    imgs = world * B * args.steps
    value = imgs / (ms_resident * 1e-3)
    e2e_value = imgs / (ms_e2e * 1e-3)
    per_gpu = value / world
    achieved = per_gpu * F_STEP * 1e-12
    kern = dominant_kernel_roofline(burst) if world == 1 else None
    line = {
        "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_resident / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "SD1.5 UNet LoRA r=8 on all attn1/attn2 Linear (128 layers, %d params), bs=4/GPU, 512x512 (64x64 latent), "
                               "77 tokens; step = add_noise + UNet fwd + MSE + bwd + grad all-reduce + clip + AdamW" % n_lora,
                   "global_batch": world * B, "per_gpu_batch": B, "lora_rank": LORA_RANK, "parallelism": f"dp{world}",
                   "l2": "working set (1.7 GB bf16 weights + activations) is far larger than the 126 MB L2; no explicit flush",
                   "cuda_graph": True, "grad_checkpointing": False,
                   "side_stream": os.environ.get("HCP_SIDE_STREAM", "1") != "0", "pdl": os.environ.get("HCP_PDL", "1") != "0"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "images/s", "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
        "gpu_launches": step.launches_per_step * args.steps,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained,
                     "traffic": ncu_dram_traffic(), "peak_source": f"{peak_src} bf16_tflops_sustained (step-level); kernels vs burst {burst}",
                     "flop_per_image": F_STEP, "kernels": kern},
        "final_loss": losses[-1],
    }
    if world == 1 and not args.no_cpu_baseline:
        r = cpu_reference_steps(1, 0, budget_s=30.0)
        line["cpu_baseline"] = {"value": r["images_per_s"], "unit": "images/s", "cores": r["cores"], "kind": "port",
                                "sample": "1 training step of 1 image (B=1, 64x64 latent): oracle/unet_ref.py fp32 on all host threads, no warm-up"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="hcpb200", choices=["hcpb200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if world != args.gpus:
        if args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # rank 0 prints ONE json line on stdout: NCCL's version banner / debug lines (whatever NCCL_DEBUG the box exports) go to stderr
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_product_arm(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
