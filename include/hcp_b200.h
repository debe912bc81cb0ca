/* SPDX-License-Identifier: Apache-2.0
 *
 * libhcpb200 -- C ABI of the B200 (sm_100a) kernels behind the HCP-Diffusion UNet denoising hot path.
 *
 * The reference (IrisRainbowNeko/HCP-Diffusion @ 404f0e85) has NO native code and NO FFI: every op
 * below replaces a PyTorch call made from the pure-Python hot path.  Each entry point cites the
 * reference call it replaces (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 on success, a negative hcp_status otherwise; nothing throws across the ABI;
 *     hcp_last_error_string() describes the last failure on the calling thread.
 *   - the caller owns every buffer (activations, outputs, workspaces); the library never allocates device
 *     memory, never synchronises and only enqueues work on the `stream` argument (a cudaStream_t).
 *   - activations are bf16, row-major "NHWC": a [B,H,W,C] feature map is the same memory as the [B*H*W, C]
 *     token matrix the transformer blocks use.  Accumulation is fp32.  `ld*` are row pitches in ELEMENTS.
 *   - all device pointers must be 16-byte aligned, all row pitches multiples of 8 elements.
 */
#ifndef HCP_B200_H_
#define HCP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* hcp_stream_t; /* cudaStream_t */

enum hcp_status {
    HCP_OK = 0,
    HCP_ERR_INVALID = -1,   /* bad argument / unsupported shape */
    HCP_ERR_CUDA = -2,      /* CUDA runtime or driver error (see hcp_last_error_string) */
    HCP_ERR_NO_DEVICE = -3, /* no sm_100 device / driver entry point missing */
};

int hcp_version(void);                     /* ABI version, bumps on any signature change */
const char* hcp_last_error_string(void);   /* thread-local, never NULL */
int hcp_device_check(void);                /* HCP_OK iff the current device is compute capability 10.x */

/* ------------------------------------------------------------------------------------------------
 * GEMM family (tcgen05 + TMA).  out[M,N] = sum_s A_s[M,K_s] . B_s[N,K_s]^T  (+ epilogue)
 *
 * Replaces: LinearLayer.forward -> torch.mm(x2d, (W_host + dW).T)   hcpdiff/models/lora_layers_patch.py:50-57
 *           LinearLayer.get_weight -> alpha*mm(W_up, W_down)         hcpdiff/models/lora_layers_patch.py:44-45
 *           LoraPatchContainer.forward / LoraBlock.post_forward       hcpdiff/models/lora_base_patch.py:21-35,68-74
 *           nn.Linear / 1x1 nn.Conv2d inside diffusers' UNet2DConditionModel (structure: cfgs/unet_struct.txt)
 * The rank-r LoRA product is never materialised as a [out,in] matrix: segment 1 is (x.W_down^T)[M,rpad] times
 * (alpha*W_up)[N,rpad], i.e. extra K-blocks of the same tensor-core pipeline.
 * ---------------------------------------------------------------------------------------------- */
#define HCP_GEMM_MAX_SEG 3

typedef struct hcp_gemm_args {
    int32_t nseg;                          /* 1..3 K-segments */
    const void* a[HCP_GEMM_MAX_SEG];       /* bf16 [M, k[s]] row-major, pitch lda[s] */
    const void* b[HCP_GEMM_MAX_SEG];       /* bf16 [n_rows_b[s], k[s]] row-major, pitch ldb[s]; rows >= n_rows_b read as 0 */
    int64_t lda[HCP_GEMM_MAX_SEG];
    int64_t ldb[HCP_GEMM_MAX_SEG];
    int64_t k[HCP_GEMM_MAX_SEG];           /* reduction extent; the TMA box is 64 wide, tails read as zero */
    int64_t n_rows_b[HCP_GEMM_MAX_SEG];    /* valid rows of b[s] (== N normally; < N for the LoRA down-projection) */
    int64_t M, N;
    const float* bias;                     /* fp32 [N] or NULL */
    const float* rowbias;                  /* fp32 [ceil(M/rows_per_group), N] or NULL (time-embedding bias per image) */
    int64_t rows_per_group;
    const void* residual;                  /* bf16 [M,N] pitch ldr, or NULL: added in the epilogue */
    int64_t ldr;
    void* out;                             /* bf16 [M,N] pitch ldo */
    int64_t ldo;
    int32_t flags;                         /* reserved, 0 */
} hcp_gemm_args;

int hcp_gemm_bf16(const hcp_gemm_args* args, hcp_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * 3x3 convolution as implicit GEMM (tcgen05; the im2col gather is done by 4D/5D TMA boxes with
 * out-of-bounds zero fill standing in for the padding).
 *
 * Replaces: F.conv2d inside diffusers ResnetBlock2D.conv1/conv2, Downsample2D.conv (stride 2),
 *           Upsample2D.conv (after nearest x2) -- module shapes pinned by cfgs/unet_struct.txt;
 *           and its dgrad (autograd of the same call) for the backward pass.
 *   mode 0: y[b,ho,wo,:] = sum_{kh,kw} x[b, ho*s+kh-1, wo*s+kw-1, :] . w[:, kh, kw, :]^T   (s = stride, pad 1)
 *   mode 1: transposed conv of a stride-2 conv (dgrad): x is dY [B,H/2,W/2,Cin'], output [B,H,W,Cout'];
 *           w must already be the "dgrad" arrangement (see hcp_conv3x3_args.w).
 * ---------------------------------------------------------------------------------------------- */
typedef struct hcp_conv3x3_args {
    const void* x;       /* bf16 [B, Hin, Win, Cin] */
    const void* w;       /* bf16 [Cout, 3, 3, Cin] (tap-major, channel-minor "K-major" layout) */
    int64_t B, Hin, Win, Cin, Cout;
    int32_t stride;      /* 1 or 2 */
    int32_t mode;        /* 0 = forward conv, 1 = dgrad of the stride-2 conv */
    const float* bias;   /* fp32 [Cout] or NULL */
    const float* rowbias;/* fp32 [B, Cout] or NULL (time embedding projection, broadcast over H,W) */
    const void* residual;/* bf16 [B,Hout,Wout,Cout] or NULL */
    void* out;           /* bf16 [B,Hout,Wout,Cout] */
} hcp_conv3x3_args;

int hcp_conv3x3_bf16(const hcp_conv3x3_args* args, hcp_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HCP_B200_H_ */
