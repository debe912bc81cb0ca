"""GPU parity tests of the full fine-tune path (`unet:` config items, reference cfg_net_tools.py:96-106, DreamBooth.yaml:6-10;
BASELINE config 3): every weight-gradient kernel against fp32 torch autograd on bf16-rounded operands, then every parameter
gradient of the TINY UNet (and two whole optimizer steps) against the fp32 oracle.

Tolerances: per-op rel-L2 <= 1e-2 (bf16 operands, fp32 accumulation); UNet-level parameter gradients: global rel-L2 <= 5e-2."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():          # fp32 torch references must be real fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

from hcp_diffusion_b200 import _lib, ops  # noqa: E402
from hcp_diffusion_b200._lib import call, stream_ptr  # noqa: E402
from hcp_diffusion_b200.engine import LoraTrainStep  # noqa: E402
from hcp_diffusion_b200.models import UNet2DConditionModel  # noqa: E402
from hcp_diffusion_b200.utils.cfg_net_tools import make_hcpdiff  # noqa: E402
from oracle import step_ref as S  # noqa: E402
from oracle import unet_ref as U  # noqa: E402

DEV = "cuda"
BF = torch.bfloat16


def rel_l2(a, b):
    a, b = a.detach().double().flatten().cpu(), b.detach().double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


@pytest.mark.parametrize("M,K,N", [(256, 320, 320), (1000, 768, 640), (4096, 640, 5120), (77, 320, 64), (16384, 320, 960)])
def test_linear_wgrad_and_colsum(M, K, N):
    x, dy = rnd(M, K, seed=1).to(BF), rnd(M, N, seed=2).to(BF)
    dw = torch.full((N, K), 0.5, device=DEV)             # gradients ACCUMULATE
    call("hcp_wgrad_bf16", dy.data_ptr(), N, N, x.data_ptr(), K, K, M, 1.0, dw.data_ptr(), K, 1, stream_ptr())
    assert rel_l2(dw - 0.5, dy.float().t() @ x.float()) < 1e-2
    # a column slice of dY (fused QKV group: one host of three) into a column slice of dW (second input segment of a concat)
    if N >= 128:
        dw2 = torch.zeros((64, 2 * K), device=DEV)
        call("hcp_wgrad_bf16", dy.data_ptr() + 2 * 64, N, 64, x.data_ptr(), K, K, M, 1.0, dw2.data_ptr() + 4 * K, 2 * K, 1, stream_ptr())
        assert rel_l2(dw2[:, K:], dy[:, 64:128].float().t() @ x.float()) < 1e-2 and float(dw2[:, :K].abs().sum()) == 0.0
    db = torch.zeros(N, device=DEV)
    call("hcp_colsum_bf16", dy.data_ptr(), N, M, N, 0, 1.0, db.data_ptr(), N, stream_ptr())
    assert rel_l2(db, dy.float().sum(0)) < 1e-3


@pytest.mark.parametrize("B,H,Cin,Cout,stride", [(2, 16, 64, 128, 1), (2, 32, 320, 320, 2), (3, 8, 128, 64, 1), (4, 64, 320, 320, 1), (4, 8, 1280, 1280, 1)])
def test_conv3x3_wgrad_and_temb_grad(B, H, Cin, Cout, stride):
    W = H
    Ho = H // stride
    x = rnd(B, H * W, Cin, seed=1).to(BF)
    dy = rnd(B, Ho * Ho, Cout, seed=2).to(BF)
    dw = torch.zeros((Cout, Cin, 3, 3), device=DEV)
    call("hcp_wgrad_conv3x3_bf16", dy.data_ptr(), Cout, x.data_ptr(), B, H, W, Cin, stride, 1.0, dw.data_ptr(), stream_ptr())
    w = torch.zeros((Cout, Cin, 3, 3), device=DEV, requires_grad=True)
    xr = x.float().view(B, H, W, Cin).permute(0, 3, 1, 2)
    y = F.conv2d(xr, w, None, stride=stride, padding=1)
    y.backward(dy.float().view(B, Ho, Ho, Cout).permute(0, 3, 1, 2))
    assert rel_l2(dw, w.grad) < 1e-2
    # per-image column sums = the gradient of the time-embedding row bias of a ResnetBlock2D
    dt = torch.zeros((B, Cout), device=DEV)
    call("hcp_colsum_bf16", dy.data_ptr(), Cout, B * Ho * Ho, Cout, Ho * Ho, 1.0, dt.data_ptr(), Cout, stream_ptr())
    assert rel_l2(dt, dy.float().sum(1)) < 1e-3


@pytest.mark.parametrize("B,HW,C1,C2,silu", [(2, 256, 320, 0, True), (2, 64, 1280, 640, True), (3, 1024, 640, 0, False)])
def test_groupnorm_and_layernorm_affine_grads(B, HW, C1, C2, silu):
    C = C1 + C2
    x1 = (rnd(B, HW, C1, seed=1) * 2 + 0.5).to(BF).requires_grad_(True)
    x2 = (rnd(B, HW, C2, seed=2) - 0.3).to(BF).requires_grad_(True) if C2 else None
    gamma = (1 + 0.1 * rnd(C, seed=3)).requires_grad_(True)
    beta = (0.1 * rnd(C, seed=4)).requires_grad_(True)
    y = ops.group_norm(gamma, beta, 32, 1e-5, silu, x1, x2)[0]
    dy = rnd(B, HW, C, seed=5).to(BF)
    y.backward(dy)
    xr = torch.cat([x1.detach().float()] + ([x2.detach().float()] if C2 else []), -1)
    g2, b2 = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    yr = F.group_norm(xr.transpose(1, 2), 32, g2, b2, 1e-5)
    yr = (F.silu(yr) if silu else yr).transpose(1, 2)
    yr.backward(dy.float())
    assert rel_l2(gamma.grad, g2.grad) < 1e-2 and rel_l2(beta.grad, b2.grad) < 1e-2
    # LayerNorm
    x = (rnd(B * HW, C1, seed=6) * 1.5 + 0.2).to(BF).requires_grad_(True)
    gl, bl = (1 + 0.1 * rnd(C1, seed=7)).requires_grad_(True), (0.1 * rnd(C1, seed=8)).requires_grad_(True)
    yl, _ = ops.layer_norm(gl, bl, 1e-5, x)
    dyl = rnd(B * HW, C1, seed=9).to(BF)
    yl.backward(dyl)
    g3, b3 = gl.detach().clone().requires_grad_(True), bl.detach().clone().requires_grad_(True)
    F.layer_norm(x.detach().float(), (C1,), g3, b3, 1e-5).backward(dyl.float())
    assert rel_l2(gl.grad, g3.grad) < 1e-2 and rel_l2(bl.grad, b3.grad) < 1e-2


def test_boundary_conv_wgrads_and_small_linear():
    B, H, W = 3, 16, 16
    lat = rnd(B, 4, H, W, seed=1)
    dh = rnd(B, H * W, 320, seed=2).to(BF)
    dw, db = torch.zeros((320, 4, 3, 3), device=DEV), torch.zeros(320, device=DEV)
    call("hcp_conv_in_wgrad_f32", dh.data_ptr(), lat.data_ptr(), B, 4, H, W, 320, dw.data_ptr(), db.data_ptr(), stream_ptr())
    w = torch.zeros((320, 4, 3, 3), device=DEV, requires_grad=True)
    b = torch.zeros(320, device=DEV, requires_grad=True)
    F.conv2d(lat, w, b, padding=1).backward(dh.float().view(B, H, W, 320).permute(0, 3, 1, 2))
    assert rel_l2(dw, w.grad) < 1e-4 and rel_l2(db, b.grad) < 1e-4
    act = rnd(B, H * W, 320, seed=3).to(BF)
    dy = rnd(B, 4, H, W, seed=4)
    dw2, db2 = torch.zeros((4, 320, 3, 3), device=DEV), torch.zeros(4, device=DEV)
    call("hcp_conv_out_wgrad_f32", dy.data_ptr(), act.data_ptr(), B, H, W, 320, 4, dw2.data_ptr(), db2.data_ptr(), stream_ptr())
    w2 = torch.zeros((4, 320, 3, 3), device=DEV, requires_grad=True)
    b2 = torch.zeros(4, device=DEV, requires_grad=True)
    F.conv2d(act.float().view(B, H, W, 320).permute(0, 3, 1, 2), w2, b2, padding=1).backward(dy)
    assert rel_l2(dw2, w2.grad) < 1e-4 and rel_l2(db2, b2.grad) < 1e-4
    # small fp32 linear with SiLU (time-embedding MLP) through the autograd wrapper
    lin = torch.nn.Linear(320, 1280).to(DEV)
    x = rnd(B, 320, seed=5).requires_grad_(True)
    wb = lin.weight.detach().to(BF).contiguous()
    y = ops.small_linear(x, wb, lin.bias.detach(), True, [(lin.weight, lin.bias, 0, 1280)])
    dyy = rnd(B, 1280, seed=6)
    y.backward(dyy)
    xr = x.detach().clone().requires_grad_(True)
    wr, br = wb.float().requires_grad_(True), lin.bias.detach().clone().requires_grad_(True)
    F.silu(F.linear(xr, wr, br)).backward(dyy)
    assert rel_l2(y, F.silu(F.linear(xr, wr, br))) < 1e-5
    assert rel_l2(x.grad, xr.grad) < 1e-4 and rel_l2(lin.weight.grad, wr.grad) < 1e-4 and rel_l2(lin.bias.grad, br.grad) < 1e-4


def build_full_ft(spec, sd):
    unet = UNet2DConditionModel(sample_size=spec.sample_size, block_out_channels=spec.block_out_channels,
                                attention_head_dim=spec.num_heads, cross_attention_dim=spec.cross_attention_dim)
    unet.load_state_dict(sd)
    unet = unet.to(DEV).requires_grad_(False).eval()
    groups, lora = make_hcpdiff(unet, [{"lr": 1e-5, "layers": [""]}], None)        # DreamBooth.yaml:6-10: every layer of the UNet
    assert lora.empty() and len(groups) == 1
    return unet, groups


def test_tiny_unet_every_parameter_gradient_matches_oracle():
    spec = U.TINY
    sd = U.init_params(spec)
    unet, groups = build_full_ft(spec, sd)
    names = [n for n, _ in unet.named_parameters()]
    assert set(names) == set(sd) and all(p.requires_grad for p in unet.parameters())
    step = LoraTrainStep(unet, groups, use_cuda_graph=False, max_grad_norm=1.0)
    lat, noise, t, ehs = U.synthetic_batch(4, spec)
    step._forward_backward(lat.to(DEV), noise.to(DEV), t.to(DEV), ehs.to(DEV))
    torch.cuda.synchronize()
    ref_sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    x_t = U.add_noise(lat, noise, t, U.ddpm_alphas_cumprod())
    pred = U.unet_forward(ref_sd, x_t, t, ehs, spec=spec)
    loss = F.mse_loss(pred, noise, reduction="none").mean()
    loss.backward()
    assert abs(float(step.loss) - float(loss)) < 2e-2 * float(loss)
    num = den = 0.0
    worst = []
    for name, p in unet.named_parameters():
        ref = ref_sd[name].grad
        assert ref is not None, name
        e = rel_l2(p.grad, ref)
        worst.append((e, name, float(ref.norm())))
        num += float((p.grad.cpu().double() - ref.double()).pow(2).sum())
        den += float(ref.double().pow(2).sum())
    worst.sort(reverse=True)
    print("worst parameter gradients:", worst[:6])
    print("global relL2", math.sqrt(num / den))
    assert math.sqrt(num / den) < 5e-2
    assert all(e < 0.25 for e, _, _ in worst)


def test_tiny_full_finetune_steps_match_reference_loop():
    spec = U.TINY
    sd = U.init_params(spec)
    unet, groups = build_full_ft(spec, sd)
    ref_sd = {k: v.clone() for k, v in sd.items()}
    names = [n for n, _ in unet.named_parameters()]
    ref = S.ReferenceLoop(ref_sd, None, spec, lr=1e-5, train_base=names)
    p0 = {n: p.detach().clone() for n, p in unet.named_parameters()}
    step = LoraTrainStep(unet, groups, lr=1e-5, use_cuda_graph=True)
    for it in range(3):
        lat, noise, t, ehs = U.synthetic_batch(4, spec, seed=300 + it)
        l_ref = ref.micro_step(lat, noise, t, ehs)
        l_prod = float(step.step(lat, noise, t, ehs).cpu())
        assert abs(l_prod - l_ref) < 2e-2 * abs(l_ref), (it, l_prod, l_ref)
    du_p = torch.cat([(p.detach() - p0[n]).flatten().cpu() for n, p in unet.named_parameters()])
    du_r = torch.cat([(ref_sd[n].detach() - sd[n]).flatten() for n in names])
    cos = float((du_p.double() @ du_r.double()) / (du_p.double().norm() * du_r.double().norm()))
    print("update cosine", cos, "norm ratio", float(du_p.norm() / du_r.norm()))
    assert cos > 0.9 and 0.9 < float(du_p.norm() / du_r.norm()) < 1.1
