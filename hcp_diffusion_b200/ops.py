"""Autograd wrappers around the C-ABI kernels (libhcpb200).

Every function here enqueues hand-written sm_100a kernels on the current CUDA stream through `_lib.call`; torch only
allocates the buffers and records the autograd graph.  Activations are bf16, "NHWC": a feature map [B,H,W,C] and the
token matrix [B*H*W, C] are the same memory.

Gradient fan-in is folded into producer kernels instead of separate adds: the normalisation functions return an alias of
their input next to the normalised output; the residual consumer uses the alias, so the normalisation backward receives
both gradients and adds them inside its own kernel.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import AttnArgs, AttnBwdArgs, ConvArgs, GemmArgs, GroupNormArgs, call, ptr, stream_ptr

BF16 = torch.bfloat16


def _chk(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != BF16 or not t.is_cuda:
        raise _lib.HcpError(f"{name}: expected a CUDA bf16 tensor, got {t.dtype} on {t.device}")
    return t if t.is_contiguous() else t.contiguous()


# ----------------------------------------------------------------------------------------------------------------------
# raw launches
# ----------------------------------------------------------------------------------------------------------------------
def gemm_raw(a_list: Sequence[Tuple[torch.Tensor, int, int]], b_list: Sequence[Tuple[torch.Tensor, int, int, int]], M: int, N: int,
             out: torch.Tensor, ldo: int, bias: Optional[torch.Tensor] = None, rowbias: Optional[torch.Tensor] = None,
             rows_per_group: int = 0, residual: Optional[torch.Tensor] = None, ldr: int = 0) -> None:
    """out[M,N] = sum_s A_s . B_s^T (+bias +rowbias +residual).
    a_list: (tensor_or_ptr_holder, lda, k);  b_list: (tensor, ldb, n_rows_b, elem_offset)."""
    g = GemmArgs()
    g.nseg = len(a_list)
    for s, ((a, lda, k), (b, ldb, nrb, boff)) in enumerate(zip(a_list, b_list)):
        g.a[s] = a.data_ptr()
        g.lda[s] = lda
        g.k[s] = k
        g.b[s] = b.data_ptr() + 2 * boff
        g.ldb[s] = ldb
        g.n_rows_b[s] = nrb
    g.M, g.N = M, N
    g.bias = ptr(bias)
    g.rowbias = ptr(rowbias)
    g.rows_per_group = rows_per_group
    g.residual = ptr(residual)
    g.ldr = ldr
    g.out = out.data_ptr()
    g.ldo = ldo
    wsb = _lib.lib().hcp_splitk_workspace_bytes(M, N, sum(k for _, _, k in a_list))
    if wsb:
        ws = torch.empty((wsb // 4,), dtype=torch.float32, device=out.device)
        g.workspace, g.workspace_bytes = ws.data_ptr(), wsb
        _lib.launch_count += 1
    call("hcp_gemm_bf16", C.byref(g), stream_ptr())


def conv3x3_raw(x: torch.Tensor, w: torch.Tensor, B: int, Hin: int, Win: int, Cin: int, Cout: int, stride: int, mode: int,
                out: torch.Tensor, bias=None, rowbias=None, residual=None, rowbias_ld: int = 0) -> None:
    a = ConvArgs()
    a.x, a.w = x.data_ptr(), w.data_ptr()
    a.B, a.Hin, a.Win, a.Cin, a.Cout = B, Hin, Win, Cin, Cout
    a.stride, a.mode = stride, mode
    a.bias, a.rowbias, a.residual = ptr(bias), ptr(rowbias), ptr(residual)
    a.rowbias_ld = rowbias_ld
    a.out = out.data_ptr()
    if mode == 0:
        Mo = B * (Hin // stride) * (Win // stride)
        wsb = _lib.lib().hcp_splitk_workspace_bytes(Mo, Cout, 9 * Cin)
        if wsb:
            ws = torch.empty((wsb // 4,), dtype=torch.float32, device=out.device)
            a.workspace, a.workspace_bytes = ws.data_ptr(), wsb
            _lib.launch_count += 1
    call("hcp_conv3x3_bf16", C.byref(a), stream_ptr())
    if mode == 1:
        _lib.launch_count += 3


# ----------------------------------------------------------------------------------------------------------------------
# packed weights
# ----------------------------------------------------------------------------------------------------------------------
class LoraBlockRef:
    """One LoRA block inside a fused linear group (see hcp_lora_job in include/hcp_b200.h)."""
    __slots__ = ("w_down", "w_up", "alpha", "rank", "in_dim", "out_dim", "c0", "o0", "g_down", "g_up")

    def __init__(self, w_down, w_up, alpha, c0, o0):
        self.w_down, self.w_up, self.alpha = w_down, w_up, float(alpha)
        self.rank, self.in_dim = w_down.shape
        self.out_dim = w_up.shape[0]
        self.c0, self.o0 = c0, o0
        self.g_down = None   # optional fp32 views into a flat gradient buffer (direct accumulation)
        self.g_up = None


class LinearPack:
    """bf16 operands of one (possibly fused, possibly LoRA-patched) linear group  y = x . W^T + b."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor], k_splits: Optional[Sequence[int]] = None):
        # weight fp32/bf16 [N, K]
        self.N, self.K = weight.shape
        self.W = weight.detach().to(BF16).contiguous()
        self.WT = self.W.t().contiguous()
        self.bias = None if bias is None else bias.detach().float().contiguous()
        self.k_splits = list(k_splits) if k_splits else [self.K]
        self.lora: List[LoraBlockRef] = []
        self.r_tot = 0
        self.A = self.AT = self.Bl = self.BlT = None

    def attach_lora(self, blocks: List[LoraBlockRef]) -> None:
        self.lora = blocks
        self.r_tot = sum(b.rank for b in blocks)
        if self.r_tot > 64 or any(b.rank > 32 for b in blocks):
            raise _lib.HcpError(f"LoRA ranks of one fused linear group must sum to <= 64 (each <= 32), got {[b.rank for b in blocks]}")
        dev = self.W.device
        self.A = torch.zeros((self.r_tot, self.K), dtype=BF16, device=dev)
        self.AT = torch.zeros((self.K, 64), dtype=BF16, device=dev)
        self.Bl = torch.zeros((self.N, 64), dtype=BF16, device=dev)
        self.BlT = torch.zeros((self.r_tot, self.N), dtype=BF16, device=dev)

    def jobs(self) -> List[_lib.LoraJob]:
        out = []
        for b in self.lora:
            j = _lib.LoraJob()
            j.w_down, j.w_up, j.alpha = b.w_down.data_ptr(), b.w_up.data_ptr(), b.alpha
            j.rank, j.in_dim, j.out_dim = b.rank, b.in_dim, b.out_dim
            j.c0, j.o0, j.out_tot = b.c0, b.o0, self.N
            j.A, j.AT, j.Bl, j.BlT = self.A.data_ptr(), self.AT.data_ptr(), self.Bl.data_ptr(), self.BlT.data_ptr()
            out.append(j)
        return out


class ConvPack:
    """bf16 operands of one 3x3 convolution (weights [Cout,Cin,3,3] fp32 -> tap-major K-major matrices)."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor], stride: int):
        self.Cout, self.Cin = weight.shape[0], weight.shape[1]
        self.stride = stride
        w = weight.detach().to(BF16)
        self.W = w.permute(0, 2, 3, 1).contiguous()                          # [Cout, kh, kw, Cin]
        if stride == 1:
            self.Wd = w.flip(2, 3).permute(1, 2, 3, 0).contiguous()          # dgrad: [Cin, kh', kw', Cout], taps flipped
        else:
            self.Wd = w.permute(1, 2, 3, 0).contiguous()                     # stride-2 dgrad arrangement (not flipped)
        self.bias = None if bias is None else bias.detach().float().contiguous()


# ----------------------------------------------------------------------------------------------------------------------
# linear (+LoRA, + fused residual)
# ----------------------------------------------------------------------------------------------------------------------
class FusedLinearFn(torch.autograd.Function):
    """y = cat(xs, -1) . W^T + b (+ T . Bl^T, T = x . A^T)(+ residual).  Reference semantics:
    LoraPatchContainer.forward / LoraBlock.post_forward / LinearLayer.forward (hcpdiff/models/lora_base_patch.py:21-35,
    68-74, lora_layers_patch.py:44-57) without materialising W + alpha*W_up@W_down."""

    @staticmethod
    def forward(ctx, pack: LinearPack, residual: Optional[torch.Tensor], n_x: int, *tensors):
        xs = [_chk(t, "linear input") for t in tensors[:n_x]]
        ctx.n_extra = len(tensors) - n_x      # LoRA parameters: autograd inputs so the node exists even when x has no grad
        M = xs[0].numel() // xs[0].shape[-1]
        ks = [x.shape[-1] for x in xs]
        if ks != pack.k_splits:
            raise _lib.HcpError(f"linear: input widths {ks} do not match the packed weight splits {pack.k_splits}")
        N = pack.N
        out = torch.empty((*xs[0].shape[:-1], N), dtype=BF16, device=xs[0].device)
        a_list = [(x, k, k) for x, k in zip(xs, ks)]
        b_list, off = [], 0
        for k in ks:
            b_list.append((pack.W, pack.K, N, off))
            off += k
        T = None
        if pack.lora:
            if n_x != 1:
                raise _lib.HcpError("LoRA on a multi-input linear is not supported")
            T = torch.empty((M, 64), dtype=BF16, device=xs[0].device)
            gemm_raw([(xs[0], ks[0], ks[0])], [(pack.A, pack.K, pack.r_tot, 0)], M, 64, T, 64)
            a_list.append((T, 64, pack.r_tot))
            b_list.append((pack.Bl, 64, N, 0))
        res = None
        if residual is not None:
            res = _chk(residual, "linear residual")
        gemm_raw(a_list, b_list, M, N, out, N, bias=pack.bias, residual=res, ldr=N)
        ctx.pack, ctx.n_x, ctx.M, ctx.ks = pack, n_x, M, ks
        ctx.has_res = residual is not None
        saved = list(xs) if pack.lora else []
        if T is not None:
            saved.append(T)
        ctx.save_for_backward(*saved)
        ctx.x_shapes = [x.shape for x in xs]
        return out

    @staticmethod
    def backward(ctx, dy):
        pack, M, ks = ctx.pack, ctx.M, ctx.ks
        dy = _chk(dy, "linear grad")
        N = pack.N
        U = None
        if pack.lora:
            x, T = ctx.saved_tensors
            U = torch.empty((M, 64), dtype=BF16, device=dy.device)
            gemm_raw([(dy, N, N)], [(pack.BlT, N, pack.r_tot, 0)], M, 64, U, 64)
            nb = len(pack.lora)
            down = (_lib.LoraGradBlock * nb)()
            up = (_lib.LoraGradBlock * nb)()
            for i, b in enumerate(pack.lora):
                gd = b.g_down if b.g_down is not None else _acc_grad(b.w_down)
                gu = b.g_up if b.g_up is not None else _acc_grad(b.w_up)
                down[i].n_lo, down[i].n_hi, down[i].c0, down[i].rank = 0, ks[0], b.c0, b.rank
                down[i].scale, down[i].transpose_out, down[i].dst, down[i].dst_ld = 1.0, 0, gd.data_ptr(), ks[0]
                up[i].n_lo, up[i].n_hi, up[i].c0, up[i].rank = b.o0, b.o0 + b.out_dim, b.c0, b.rank
                up[i].scale, up[i].transpose_out, up[i].dst, up[i].dst_ld = b.alpha, 1, gu.data_ptr(), b.rank
            # dW_down = U^T x ;  dW_up = alpha * dY^T T   (tensor-core TN GEMMs, all blocks of the group per launch)
            if nb <= 8:
                call("hcp_lora_grad_pair", U.data_ptr(), x.data_ptr(), ks[0], ks[0], down, T.data_ptr(), dy.data_ptr(), N, N, up, nb, M,
                     stream_ptr())
            else:
                for b0 in range(0, nb, 8):
                    n8 = min(8, nb - b0)
                    call("hcp_lora_grad", U.data_ptr(), x.data_ptr(), ks[0], M, 0, ks[0], C.cast(C.byref(down[b0]), C.POINTER(_lib.LoraGradBlock)),
                         n8, stream_ptr())
                    call("hcp_lora_grad", T.data_ptr(), dy.data_ptr(), N, M, 0, N, C.cast(C.byref(up[b0]), C.POINTER(_lib.LoraGradBlock)), n8,
                         stream_ptr())
        grads = []
        off = 0
        for i, k in enumerate(ks):
            if ctx.needs_input_grad[3 + i]:
                dx = torch.empty(ctx.x_shapes[i], dtype=BF16, device=dy.device)
                a_list = [(dy, N, N)]
                b_list = [(pack.WT, N, k, off * N)]
                if U is not None:
                    a_list.append((U, 64, pack.r_tot))
                    b_list.append((pack.AT, 64, k, 0))
                gemm_raw(a_list, b_list, M, k, dx, k)
                grads.append(dx)
            else:
                grads.append(None)
            off += k
        dres = dy if (ctx.has_res and ctx.needs_input_grad[1]) else None
        return (None, dres, None, *grads, *([None] * ctx.n_extra))


def _acc_grad(p: torch.Tensor) -> torch.Tensor:
    """fp32 .grad of a LoRA parameter, created zeroed on first use; the kernels accumulate into it directly."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, dtype=torch.float32)
    return p.grad


def fused_linear(pack: LinearPack, xs: Sequence[torch.Tensor], residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    # The LoRA parameters are passed as autograd inputs so the node is recorded even when x carries no gradient (cross-attention
    # k/v on the text embedding); their gradients are accumulated in place by the kernels (fp32 .grad / flat grad buffer).
    extra = []
    for b in pack.lora:
        extra += [b.w_down, b.w_up]
    return FusedLinearFn.apply(pack, residual, len(xs), *xs, *extra)


# ----------------------------------------------------------------------------------------------------------------------
# 3x3 convolution
# ----------------------------------------------------------------------------------------------------------------------
class Conv3x3Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pack: ConvPack, geom: Tuple[int, int, int], rowbias: Optional[torch.Tensor], residual: Optional[torch.Tensor],
                x: torch.Tensor):
        B, H, W = geom
        x = _chk(x, "conv input")
        s = pack.stride
        Ho, Wo = H // s, W // s
        out = torch.empty((B, Ho * Wo, pack.Cout), dtype=BF16, device=x.device)
        res = None if residual is None else _chk(residual, "conv residual")
        rb_ld = 0
        if rowbias is not None:
            if rowbias.dtype != torch.float32 or rowbias.stride(-1) != 1:
                raise _lib.HcpError("conv rowbias must be fp32 with unit inner stride")
            rb_ld = rowbias.stride(0)
        conv3x3_raw(x, pack.W, B, H, W, pack.Cin, pack.Cout, s, 0, out, bias=pack.bias, rowbias=rowbias, residual=res, rowbias_ld=rb_ld)
        ctx.pack, ctx.geom = pack, geom
        ctx.has_res = residual is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        pack = ctx.pack
        B, H, W = ctx.geom
        dy = _chk(dy, "conv grad")
        dx = None
        if ctx.needs_input_grad[4]:
            dx = torch.empty((B, H * W, pack.Cin), dtype=BF16, device=dy.device)
            if pack.stride == 1:
                conv3x3_raw(dy, pack.Wd, B, H, W, pack.Cout, pack.Cin, 1, 0, dx)
            else:
                conv3x3_raw(dy, pack.Wd, B, H // 2, W // 2, pack.Cout, pack.Cin, 2, 1, dx)
        dres = dy if (ctx.has_res and ctx.needs_input_grad[3]) else None
        return None, None, None, dres, dx


def conv3x3(pack: ConvPack, x: torch.Tensor, geom, rowbias=None, residual=None) -> torch.Tensor:
    return Conv3x3Fn.apply(pack, geom, rowbias, residual, x)


# ----------------------------------------------------------------------------------------------------------------------
# normalisation
# ----------------------------------------------------------------------------------------------------------------------
class GroupNormFn(torch.autograd.Function):
    """(y, alias(x1)[, alias(x2)]) = GN(cat(x1, x2)) [+SiLU]; the aliases carry the residual-branch gradients back so the
    fan-in add happens inside the GroupNorm backward kernel."""

    @staticmethod
    def forward(ctx, gamma: torch.Tensor, beta: torch.Tensor, groups: int, eps: float, silu: bool, x1: torch.Tensor,
                x2: Optional[torch.Tensor]):
        x1 = _chk(x1, "groupnorm input")
        B, HW, C1 = x1.shape
        C2 = 0
        if x2 is not None:
            x2 = _chk(x2, "groupnorm input 2")
            C2 = x2.shape[-1]
        y = torch.empty((B, HW, C1 + C2), dtype=BF16, device=x1.device)
        stats = torch.empty((B, groups, 2), dtype=torch.float32, device=x1.device)
        wsb = _lib.lib().hcp_groupnorm_workspace_bytes(B, HW, groups)
        ws = torch.empty((max(wsb, 4) // 4,), dtype=torch.float32, device=x1.device)
        a = GroupNormArgs()
        a.x1, a.x2 = x1.data_ptr(), ptr(x2)
        a.B, a.HW, a.C1, a.C2, a.G = B, HW, C1, C2, groups
        a.gamma, a.beta, a.eps, a.silu = gamma.data_ptr(), beta.data_ptr(), eps, int(silu)
        a.stats, a.workspace, a.workspace_bytes = stats.data_ptr(), ws.data_ptr(), wsb
        a.y = y.data_ptr()
        call("hcp_groupnorm_fwd_bf16", C.byref(a), stream_ptr())
        ctx.save_for_backward(x1, x2, stats, gamma, beta)
        ctx.cfg = (groups, eps, silu)
        if x2 is None:
            return y, x1
        return y, x1, x2

    @staticmethod
    def backward(ctx, dy, d1, d2=None):
        x1, x2, stats, gamma, beta = ctx.saved_tensors
        groups, eps, silu = ctx.cfg
        B, HW, C1 = x1.shape
        C2 = 0 if x2 is None else x2.shape[-1]
        need1 = ctx.needs_input_grad[5]
        need2 = x2 is not None and ctx.needs_input_grad[6]
        if not (need1 or need2):
            return (None,) * 7
        if dy is None:
            return None, None, None, None, None, d1, d2
        dy = _chk(dy, "groupnorm grad")
        dx1 = torch.empty_like(x1)
        dx2 = None if x2 is None else torch.empty_like(x2)
        wsb = _lib.lib().hcp_groupnorm_workspace_bytes(B, HW, groups)
        ws = torch.empty((max(wsb, 4) // 4,), dtype=torch.float32, device=x1.device)
        a = GroupNormArgs()
        a.x1, a.x2 = x1.data_ptr(), ptr(x2)
        a.B, a.HW, a.C1, a.C2, a.G = B, HW, C1, C2, groups
        a.gamma, a.beta, a.eps, a.silu = gamma.data_ptr(), beta.data_ptr(), eps, int(silu)
        a.stats, a.workspace, a.workspace_bytes = stats.data_ptr(), ws.data_ptr(), wsb
        a.dy = dy.data_ptr()
        a.add1 = None if d1 is None else _chk(d1, "groupnorm alias grad").data_ptr()
        a.add2 = None if d2 is None else _chk(d2, "groupnorm alias grad 2").data_ptr()
        a.dx1, a.dx2 = dx1.data_ptr(), ptr(dx2)
        call("hcp_groupnorm_bwd_bf16", C.byref(a), stream_ptr())
        return None, None, None, None, None, dx1, dx2


def group_norm(gamma, beta, groups, eps, silu, x1, x2=None):
    return GroupNormFn.apply(gamma, beta, groups, eps, silu, x1, x2)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, gamma: torch.Tensor, beta: torch.Tensor, eps: float, x: torch.Tensor):
        x = _chk(x, "layernorm input")
        C_ = x.shape[-1]
        M = x.numel() // C_
        y = torch.empty_like(x)
        stats = torch.empty((M, 2), dtype=torch.float32, device=x.device)
        call("hcp_layernorm_fwd_bf16", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, M, C_, stats.data_ptr(), y.data_ptr(),
             stream_ptr())
        ctx.save_for_backward(x, stats, gamma)
        return y, x

    @staticmethod
    def backward(ctx, dy, dalias):
        x, stats, gamma = ctx.saved_tensors
        if not ctx.needs_input_grad[3]:
            return None, None, None, None
        if dy is None:
            return None, None, None, dalias
        dy = _chk(dy, "layernorm grad")
        C_ = x.shape[-1]
        M = x.numel() // C_
        dx = torch.empty_like(x)
        add = None if dalias is None else _chk(dalias, "layernorm alias grad")
        call("hcp_layernorm_bwd_bf16", x.data_ptr(), dy.data_ptr(), ptr(add), gamma.data_ptr(), stats.data_ptr(), M, C_, dx.data_ptr(),
             stream_ptr())
        return None, None, None, dx


def layer_norm(gamma, beta, eps, x):
    return LayerNormFn.apply(gamma, beta, eps, x)


# ----------------------------------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------------------------------
class AttentionFn(torch.autograd.Function):
    """softmax(q k^T * scale + bias) v over heads laid out as column blocks.
    q_src [B,Lq,ldq] holds q at column q_off; kv_src [B,Lkv,ldkv] holds k at k_off and v at v_off (q_src may be kv_src:
    the fused QKV projection output)."""

    @staticmethod
    def forward(ctx, heads: int, C_: int, offs: Tuple[int, int, int], kv_bias: Optional[torch.Tensor], q_src: torch.Tensor,
                kv_src: Optional[torch.Tensor]):
        q_src = _chk(q_src, "attention q")
        same = kv_src is None
        kvt = q_src if same else _chk(kv_src, "attention kv")
        B, Lq, ldq = q_src.shape
        _, Lkv, ldkv = kvt.shape
        d = C_ // heads
        scale = 1.0 / math.sqrt(d)
        o = torch.empty((B, Lq, C_), dtype=BF16, device=q_src.device)
        lse = torch.empty((B, heads, Lq), dtype=torch.float32, device=q_src.device)
        a = AttnArgs()
        a.q, a.ldq = q_src.data_ptr() + 2 * offs[0], ldq
        a.k, a.ldk = kvt.data_ptr() + 2 * offs[1], ldkv
        a.v, a.ldv = kvt.data_ptr() + 2 * offs[2], ldkv
        a.B, a.H, a.Lq, a.Lkv, a.d = B, heads, Lq, Lkv, d
        a.scale = scale
        a.kv_bias = ptr(kv_bias)
        a.o, a.ldo, a.lse = o.data_ptr(), C_, lse.data_ptr()
        call("hcp_attn_fwd_bf16", C.byref(a), stream_ptr())
        ctx.save_for_backward(q_src, kvt, o, lse, kv_bias)
        ctx.cfg = (heads, C_, offs, same, scale)
        return o

    @staticmethod
    def backward(ctx, do):
        q_src, kvt, o, lse, kv_bias = ctx.saved_tensors
        heads, C_, offs, same, scale = ctx.cfg
        do = _chk(do, "attention grad")
        B, Lq, ldq = q_src.shape
        _, Lkv, ldkv = kvt.shape
        d = C_ // heads
        # gradients are written straight into buffers with the layout of the sources
        dq_src = torch.empty_like(q_src)
        dkv = dq_src if same else torch.empty_like(kvt)
        wsb = _lib.lib().hcp_attn_bwd_workspace_bytes(B, heads, Lq, Lkv, d)
        ws = torch.empty((wsb // 4,), dtype=torch.float32, device=do.device)
        a = AttnBwdArgs()
        a.q, a.ldq = q_src.data_ptr() + 2 * offs[0], ldq
        a.k, a.ldk = kvt.data_ptr() + 2 * offs[1], ldkv
        a.v, a.ldv = kvt.data_ptr() + 2 * offs[2], ldkv
        a.o, a.ldo, a.dout, a.lddo = o.data_ptr(), C_, do.data_ptr(), C_
        a.B, a.H, a.Lq, a.Lkv, a.d = B, heads, Lq, Lkv, d
        a.scale, a.kv_bias, a.lse = scale, ptr(kv_bias), lse.data_ptr()
        a.dq, a.lddq = dq_src.data_ptr() + 2 * offs[0], ldq
        a.dk, a.lddk = dkv.data_ptr() + 2 * offs[1], ldkv
        a.dv, a.lddv = dkv.data_ptr() + 2 * offs[2], ldkv
        a.workspace, a.workspace_bytes = ws.data_ptr(), wsb
        call("hcp_attn_bwd_bf16", C.byref(a), stream_ptr())
        if d > 128:
            _lib.launch_count += 1
        if Lkv <= 128 and Lq > 128:
            _lib.launch_count += 1
        return None, None, None, None, dq_src, (None if same else dkv)


def attention(heads: int, C_: int, offs, q_src, kv_src=None, kv_bias=None):
    return AttentionFn.apply(heads, C_, tuple(offs), kv_bias, q_src, kv_src)


# ----------------------------------------------------------------------------------------------------------------------
# elementwise
# ----------------------------------------------------------------------------------------------------------------------
class GegluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u):
        u = _chk(u, "geglu input")
        F2 = u.shape[-1]
        F_ = F2 // 2
        M = u.numel() // F2
        h = torch.empty((*u.shape[:-1], F_), dtype=BF16, device=u.device)
        call("hcp_geglu_fwd_bf16", u.data_ptr(), M, F_, h.data_ptr(), stream_ptr())
        ctx.save_for_backward(u)
        return h

    @staticmethod
    def backward(ctx, dh):
        (u,) = ctx.saved_tensors
        dh = _chk(dh, "geglu grad")
        F2 = u.shape[-1]
        M = u.numel() // F2
        du = torch.empty_like(u)
        call("hcp_geglu_bwd_bf16", u.data_ptr(), dh.data_ptr(), M, F2 // 2, du.data_ptr(), stream_ptr())
        return du


class Upsample2xFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, geom, x):
        B, H, W = geom
        x = _chk(x, "upsample input")
        C_ = x.shape[-1]
        y = torch.empty((B, 4 * H * W, C_), dtype=BF16, device=x.device)
        call("hcp_upsample2x_fwd_bf16", x.data_ptr(), B, H, W, C_, y.data_ptr(), stream_ptr())
        ctx.geom = geom
        return y

    @staticmethod
    def backward(ctx, dy):
        B, H, W = ctx.geom
        dy = _chk(dy, "upsample grad")
        C_ = dy.shape[-1]
        dx = torch.empty((B, H * W, C_), dtype=BF16, device=dy.device)
        call("hcp_upsample2x_bwd_bf16", dy.data_ptr(), B, H, W, C_, dx.data_ptr(), stream_ptr())
        return None, dx


class Fork2Fn(torch.autograd.Function):
    """Two aliases of one tensor whose gradients are summed by our own kernel (skip connections)."""

    @staticmethod
    def forward(ctx, x):
        return x, x

    @staticmethod
    def backward(ctx, da, db):
        if da is None:
            return db
        if db is None:
            return da
        da, db = _chk(da, "fork grad"), _chk(db, "fork grad")
        out = torch.empty_like(da)
        call("hcp_add_bf16", da.data_ptr(), db.data_ptr(), da.numel(), out.data_ptr(), stream_ptr())
        return out


class ConvOutFn(torch.autograd.Function):
    """bf16 NHWC [B,HW,Cin] -> fp32 NCHW [B,4,H,W] 3x3 convolution at the module boundary."""

    @staticmethod
    def forward(ctx, w: torch.Tensor, bias: Optional[torch.Tensor], geom, x):
        B, H, W = geom
        x = _chk(x, "conv_out input")
        Cin, Cout = x.shape[-1], w.shape[0]
        y = torch.empty((B, Cout, H, W), dtype=torch.float32, device=x.device)
        call("hcp_conv_out_f32", x.data_ptr(), w.data_ptr(), ptr(bias), B, H, W, Cin, Cout, y.data_ptr(), stream_ptr())
        ctx.w, ctx.geom, ctx.Cin = w, geom, Cin
        return y

    @staticmethod
    def backward(ctx, dy):
        B, H, W = ctx.geom
        dy = dy.float().contiguous()
        dx = torch.empty((B, H * W, ctx.Cin), dtype=BF16, device=dy.device)
        call("hcp_conv_out_dgrad_f32", dy.data_ptr(), ctx.w.data_ptr(), B, H, W, ctx.Cin, ctx.w.shape[0], dx.data_ptr(), stream_ptr())
        return None, None, None, dx


def conv_in(x_nchw: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """fp32 NCHW latent -> bf16 NHWC [B, H*W, Cout] (no gradient: the latent is data, conv_in is frozen)."""
    x = x_nchw.float().contiguous()
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    y = torch.empty((B, H * W, Cout), dtype=BF16, device=x.device)
    call("hcp_conv_in_f32", x.data_ptr(), w.data_ptr(), ptr(bias), B, Cin, H, W, Cout, y.data_ptr(), stream_ptr())
    return y


def skinny_linear(x: torch.Tensor, w_bf16: torch.Tensor, bias: Optional[torch.Tensor], in_mode: int, out_silu: bool) -> torch.Tensor:
    """fp32 [M<=16, K] (or timesteps [M] when in_mode == 2) -> fp32 [M, N]."""
    N, K = w_bf16.shape
    M = x.shape[0]
    y = torch.empty((M, N), dtype=torch.float32, device=w_bf16.device)
    call("hcp_skinny_linear", x.data_ptr(), w_bf16.data_ptr(), ptr(bias), M, K, N, in_mode, int(out_silu), y.data_ptr(), stream_ptr())
    return y


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    x = x.float().contiguous()
    y = torch.empty(x.shape, dtype=BF16, device=x.device)
    call("hcp_cast_f32_to_bf16", x.data_ptr(), x.numel(), y.data_ptr(), stream_ptr())
    return y
