"""Data-parallel semantics on CPU (gloo, world_size 2): the engine's recipe -- ONE all-reduce(SUM) of the flat LoRA gradient
followed by a 1/world scale inside the optimizer -- reproduces the single-process gradient of the concatenated batch
(reference: DDP averaging, hcpdiff/train_ac.py:117-123,175,482).  Arithmetic here is the CPU oracle; the flat-buffer plumbing
(`engine.FlatParams`) is the product's."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hcp_diffusion_b200.engine import FlatParams
from oracle import unet_ref as U


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    spec = U.TINY
    sd = U.init_params(spec)
    lora = U.init_lora(spec, rank=4)
    lat, noise, t, ehs = U.synthetic_batch(2 * world, spec, ctx_len=7)
    sl = slice(2 * rank, 2 * rank + 2)                                   # this rank's shard of the global batch
    _, _, grads = U.lora_step_loss_and_grads(sd, lora, lat[sl], noise[sl], t[sl], ehs[sl], spec)
    params = [torch.nn.Parameter(e.W_down.clone()) for bl in lora.values() for e in bl] + \
             [torch.nn.Parameter(e.W_up.clone()) for bl in lora.values() for e in bl]
    flat = FlatParams(params)
    gl = [g[0][0] for g in grads.values()] + [g[0][1] for g in grads.values()]
    for p, g in zip(params, gl):
        p.grad.copy_(g)                                                  # .grad is a view into the flat buffer
    dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM)                     # the only collective of the job
    flat.grad.mul_(1.0 / world)
    if rank == 0:
        torch.save({"flat": flat.grad.clone(), "offsets": flat.offsets, "numel": [p.numel() for p in params]}, os.path.join(out_dir, "dp.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_process_on_concatenated_batch(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = torch.load(os.path.join(tmp_path, "dp.pt"))
    spec = U.TINY
    sd = U.init_params(spec)
    lora = U.init_lora(spec, rank=4)
    lat, noise, t, ehs = U.synthetic_batch(2 * world, spec, ctx_len=7)
    _, _, grads = U.lora_step_loss_and_grads(sd, lora, lat, noise, t, ehs, spec)
    ref = [g[0][0] for g in grads.values()] + [g[0][1] for g in grads.values()]
    for off, n, r in zip(got["offsets"], got["numel"], ref):
        torch.testing.assert_close(got["flat"][off:off + n].view_as(r), r, rtol=1e-4, atol=1e-6)
