// SPDX-License-Identifier: Apache-2.0
// Halo-tile 3x3 convolution, stand-alone bring-up kernel + harness (test infrastructure, NOT part of the product path).
// DRAFT written at the end of round 1 without GPU access: it compiles for sm_100a, it has never run.  The hardware fact it rests on
// HAS been measured (profiles/r01_native_probe_shift.log): a K-major SWIZZLE_128B A operand may start at any multiple of 128 bytes
// inside a TMA-written buffer.
//
// Idea.  The product's implicit-GEMM convolution (csrc/gemm.cu) loads the A tile (128 output pixels x 64 input channels) once per
// TAP: nine 16 KB loads per 64-channel block, and the kernel is bound by the L2->SM operand stream.  Here the output pixels of a
// tile are taken in "virtual" order v = h (W + 2) + w (two junk pixels per image row), so that the input pixel of output v for tap
// (kh, kw) is v + kh (W + 2) + kw in the zero-padded image: for a FIXED tap the 128 rows of the A operand are 128 CONSECUTIVE rows
// of one shared-memory image of the padded input, and the nine taps differ only in the start row.  One 4-D TMA box
// (64 channels x (W + 2) x R image rows, borders zero-filled by the out-of-bounds rule) per 64-channel block replaces nine loads:
//     A bytes per (tile, 64 channels): 9 x 16 KB = 144 KB  ->  R (W + 2) 128 B = 42 KB at W = 64 (R = 5)
// while the B stream (weights, 9 x BN x 128 B) is unchanged.  Junk outputs (w >= W, h >= H) are computed and dropped by the epilogue:
// 3 % at W = 64, 12 % at W = 32 (tile quantisation included).
//
//   warp 0      TMA producer: A halo pipeline (2 stages) and B tap pipeline (4 stages)
//   warp 1      MMA issue: per 64-channel block, 9 taps x 4 k-steps on the same A stage with shifted start addresses
//   warps 2..5  epilogue: TMEM -> +bias -> bf16 -> NHWC global, rows masked by (h < H, w < W); two accumulators alternate
//
//   make -C tests/native conv_halo && tests/native/conv_halo
#include "../../hcp_diffusion_b200/csrc/common.cuh"
#include "../../hcp_diffusion_b200/csrc/host_util.h"
#include "../../include/hcp_b200.h"
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

using namespace hcp;

namespace {

constexpr int BM = 128, BK = 64, BN = 160;
constexpr int A_STAGES = 2, B_STAGES = 4;
constexpr int B_STAGE_BYTES = BN * 128;
constexpr int ACC_STRIDE = 256;                    // TMEM columns between the two accumulators
constexpr int kThreads = 6 * 32;

struct alignas(64) HaloParams {
    CUtensorMap tmX;        // bf16 NHWC activations as (C, W, H, B); box (64, P, R, 1)
    CUtensorMap tmW;        // bf16 weights [Cout, 9 * Cin] (tap-major: column = (kh * 3 + kw) * Cin + ci); box (64, BN)
    int B, H, W, Cin, Cout;
    int P, R;               // pitch W + 2; image rows per halo box
    int tiles_m_img, tiles_n;
    int a_stage_bytes;      // R * P * 128 rounded up to 1024
    const float* bias;      // [Cout] or nullptr
    __nv_bfloat16* out;     // [B, H, W, Cout]
};

// MSUB = 2: two virtually adjacent 128-pixel tiles per work item share every B (weight) tile -- the product kernel's answer to the
// operand-stream bound, and the configuration the halo has to beat (A 59 KB per 256 pixels instead of 2 x 9 x 16 KB).
template <int MSUB>
__global__ void __launch_bounds__(kThreads, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;                                            // [A_STAGES][a_stage_bytes]
    uint8_t* sB = sA + A_STAGES * p.a_stage_bytes;                 // [B_STAGES][BN][128 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sB + B_STAGES * B_STAGE_BYTES);
    uint64_t* a_full = bars;                    // [A_STAGES]
    uint64_t* a_empty = a_full + A_STAGES;      // [A_STAGES]
    uint64_t* b_full = a_empty + A_STAGES;      // [B_STAGES]
    uint64_t* b_empty = b_full + B_STAGES;      // [B_STAGES]
    uint64_t* tmem_full = b_empty + B_STAGES;   // [2]
    uint64_t* tmem_empty = tmem_full + 2;       // [2]
    uint32_t* slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_m = p.B * p.tiles_m_img;
    const int total_work = tiles_m * p.tiles_n;
    const int nkb = p.Cin / BK;

    if (threadIdx.x == 0) {
        for (int i = 0; i < A_STAGES; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
        for (int i = 0; i < B_STAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 128); }
        fence_mbar_init();
    }
    if (warp == 1) { tmem_alloc(slot, 512); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (elect_one()) {
            int sa = 0, sb = 0;
            uint32_t pa = 0, pb = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const int m = w / p.tiles_n, n0 = (w % p.tiles_n) * BN;
                const int b = m / p.tiles_m_img, v0 = (m % p.tiles_m_img) * BM * MSUB;
                const int r0 = v0 / p.P;                                   // first padded image row of the halo (image row r0 - 1)
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&a_empty[sa], pa ^ 1);
                    mbar_arrive_expect_tx(&a_full[sa], p.R * p.P * 128);
                    tma_load_4d(sA + sa * p.a_stage_bytes, &p.tmX, &a_full[sa], kb * BK, -1, r0 - 1, b);
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&b_empty[sb], pb ^ 1);
                        mbar_arrive_expect_tx(&b_full[sb], B_STAGE_BYTES);
                        tma_load_2d(sB + sb * B_STAGE_BYTES, &p.tmW, &b_full[sb], tap * p.Cin + kb * BK, n0);
                        if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
                    }
                    if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issue ------------------------------
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(BM, BN, 0, 0);
            const uint64_t adesc0 = make_smem_desc(smem_u32(sA), 16, 1024);
            const uint64_t bdesc0 = make_smem_desc(smem_u32(sB), 16, 1024);
            int sa = 0, sb = 0, item = 0;
            uint32_t pa = 0, pb = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++item) {
                const int v0 = ((w / p.tiles_n) % p.tiles_m_img) * BM * MSUB;
                const uint32_t off = (uint32_t)(v0 % p.P);                  // row of output pixel v0 inside the halo box, tap (0, 0)
                // MSUB == 1: two accumulators alternate between items; MSUB == 2: one item owns both
                const int as = (MSUB == 1) ? (item & 1) : 0;
                const uint32_t eph = (MSUB == 1) ? ((item >> 1) & 1) : (item & 1);
                mbar_wait(&tmem_empty[as], eph ^ 1);                        // epilogue drained the accumulator(s)
                tc_fence_after();
                const uint32_t acc = tmem + as * ACC_STRIDE;
                uint32_t accum = 0;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&a_full[sa], pa);
                    tc_fence_after();
                    const uint32_t a_rows = (uint32_t)(sa * p.a_stage_bytes) / 128u + off;      // in 128-byte rows from sA
#pragma unroll 1
                    for (int tap = 0; tap < 9; ++tap) {
                        mbar_wait(&b_full[sb], pb);
                        tc_fence_after();
                        // 128 bytes per row = 8 descriptor units; +2 units per 16-element k-step inside the row
                        const uint32_t a16 = (a_rows + (uint32_t)((tap / 3) * p.P + (tap % 3))) * 8u;
                        const uint32_t b16 = (uint32_t)(sb * B_STAGE_BYTES) >> 4;
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
#pragma unroll
                            for (int sub = 0; sub < MSUB; ++sub)            // the M sub-tiles are 128 rows (1024 units) apart in the halo
                                umma_ss(acc + sub * ACC_STRIDE, adesc0 + a16 + sub * 1024 + 2 * k, bdesc0 + b16 + 2 * k, idesc, accum);
                            accum = 1;
                        }
                        umma_commit(&b_empty[sb]);
                        if (++sb == B_STAGES) { sb = 0; pb ^= 1; }
                    }
                    umma_commit(&a_empty[sa]);                               // all nine taps have read this halo
                    if (++sa == A_STAGES) { sa = 0; pa ^= 1; }
                }
                umma_commit(&tmem_full[as]);
            }
        }
    } else {
        // ------------------------------ epilogue ------------------------------
        const int quarter = warp & 3;                                       // warps 2..5 -> TMEM lane quarters 2, 3, 0, 1
        const int row = quarter * 32 + lane;
        const uint32_t lb = static_cast<uint32_t>(quarter * 32) << 16;
        int item = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++item) {
            const int m = w / p.tiles_n, n0 = (w % p.tiles_n) * BN;
            const int b = m / p.tiles_m_img;
            const int as = (MSUB == 1) ? (item & 1) : 0;
            mbar_wait(&tmem_full[as], (MSUB == 1) ? ((item >> 1) & 1) : (item & 1));
            tc_fence_after();
#pragma unroll 1
            for (int sub = 0; sub < MSUB; ++sub) {
                const int v = (m % p.tiles_m_img) * BM * MSUB + sub * BM + row;
                const int h = v / p.P, wq = v % p.P;
                const bool valid = (h < p.H) && (wq < p.W);
                __nv_bfloat16* orow = p.out + ((static_cast<int64_t>(b) * p.H + h) * p.W + wq) * p.Cout + n0;
                const uint32_t acc = tmem + (as + sub) * ACC_STRIDE + lb;
#pragma unroll 1
                for (int c = 0; c < BN; c += 32) {
                    uint32_t r[32];
                    tmem_ld32(acc + c, r);
                    tmem_wait_ld();
                    if (valid) {
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            float f[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e)
                                f[e] = __uint_as_float(r[g * 8 + e]) + (p.bias ? __ldg(p.bias + n0 + c + g * 8 + e) : 0.f);
                            uint4 o;
                            o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
                            o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
                            *reinterpret_cast<uint4*>(orow + c + g * 8) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&tmem_empty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

// naive reference on a subset of the output pixels (every `stride`-th pixel, all output channels), fp32 accumulation
__global__ void conv_ref_kernel(const __nv_bfloat16* x, const __nv_bfloat16* wt, const float* bias, int B, int H, int W, int Cin, int Cout,
                                int stride, float* out_sample) {
    const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    const int64_t npix = ((int64_t)B * H * W + stride - 1) / stride;
    if (idx >= npix * Cout) return;
    const int co = (int)(idx % Cout);
    const int64_t pix = (idx / Cout) * stride;
    const int w = (int)(pix % W), h = (int)((pix / W) % H), b = (int)(pix / ((int64_t)W * H));
    float acc = bias ? bias[co] : 0.f;
    for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
            const int hh = h + kh - 1, ww = w + kw - 1;
            if (hh < 0 || hh >= H || ww < 0 || ww >= W) continue;
            const __nv_bfloat16* xp = x + (((int64_t)b * H + hh) * W + ww) * Cin;
            const __nv_bfloat16* wp = wt + ((int64_t)co * 9 + kh * 3 + kw) * Cin;
            for (int ci = 0; ci < Cin; ++ci) acc += __bfloat162float(xp[ci]) * __bfloat162float(wp[ci]);
        }
    out_sample[idx] = acc;
}

template <int MSUB>
bool run_case(int B, int H, int W, int Cin, int Cout, int stride, int iters) {
    const int64_t nx = (int64_t)B * H * W * Cin, nw = (int64_t)Cout * 9 * Cin, ny = (int64_t)B * H * W * Cout;
    std::vector<__nv_bfloat16> hx(nx), hw(nw);
    std::vector<float> hb(Cout);
    srand(7);
    for (auto& v : hx) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.f);
    for (auto& v : hw) v = __float2bfloat16((rand() % 2001 - 1000) / 1000.f * 0.05f);
    for (auto& v : hb) v = (rand() % 2001 - 1000) / 1000.f;
    __nv_bfloat16 *dx, *dw, *dy;
    float *db, *dref;
    cudaMalloc(&dx, nx * 2); cudaMalloc(&dw, nw * 2); cudaMalloc(&dy, ny * 2); cudaMalloc(&db, Cout * 4);
    cudaMemcpy(dx, hx.data(), nx * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, hw.data(), nw * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), Cout * 4, cudaMemcpyHostToDevice);
    cudaMemset(dy, 0xFF, ny * 2);

    HaloParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout;
    p.P = W + 2;
    p.R = 3 + (BM * MSUB) / p.P + 1;
    p.tiles_m_img = ((H - 1) * p.P + W + BM * MSUB - 1) / (BM * MSUB);
    p.tiles_n = Cout / BN;
    p.a_stage_bytes = (p.R * p.P * 128 + 1023) / 1024 * 1024;
    p.bias = db; p.out = dy;
    if (Cin % BK || Cout % BN || p.P > 256 || p.R > 256) { printf("[SKIP] unsupported shape\n"); return true; }
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {64, (uint32_t)p.P, (uint32_t)p.R, 1};
    if (make_tmap_nd(&p.tmX, dx, 4, dims, strides, box) || make_tmap_2d(&p.tmW, dw, 9 * (uint64_t)Cin, Cout, 9 * (uint64_t)Cin, 64, BN)) {
        printf("[FAIL] conv_halo: tensor map: %s\n", hcp_last_error_string());
        return false;
    }
    const int smem = A_STAGES * p.a_stage_bytes + B_STAGES * B_STAGE_BYTES + 256 + 1024;
    if (smem > 227 * 1024) { printf("[SKIP] shared memory %d\n", smem); return true; }
    cudaFuncSetAttribute(conv_halo_kernel<MSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int work = B * p.tiles_m_img * p.tiles_n;
    const int grid = work < sms ? work : sms;
    conv_halo_kernel<MSUB><<<grid, kThreads, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[FAIL] conv_halo: kernel error %s\n", cudaGetErrorString(e)); exit(3); }

    const int64_t npix = ((int64_t)B * H * W + stride - 1) / stride, nref = npix * Cout;
    cudaMalloc(&dref, nref * 4);
    conv_ref_kernel<<<(unsigned)((nref + 255) / 256), 256>>>(dx, dw, db, B, H, W, Cin, Cout, stride, dref);
    std::vector<float> ref(nref);
    std::vector<__nv_bfloat16> y(ny);
    cudaMemcpy(ref.data(), dref, nref * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(y.data(), dy, ny * 2, cudaMemcpyDeviceToHost);
    double err2 = 0, ref2 = 0;
    for (int64_t i = 0; i < nref; ++i) {
        const int64_t pix = (i / Cout) * stride;
        const float got = __bfloat162float(y[pix * Cout + i % Cout]);
        const double d = isnan(got) ? 1e3 : got - ref[i];
        err2 += d * d; ref2 += (double)ref[i] * ref[i];
    }
    const double rel = sqrt(err2 / (ref2 + 1e-30));

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) conv_halo_kernel<MSUB><<<grid, kThreads, smem>>>(p);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double us = 1e3 * ms / iters, tf = 2.0 * B * H * W * Cout * 9.0 * Cin / (us * 1e-6) * 1e-12;
    const bool ok = rel < 1e-2;                                        // bf16 output rounding: ~2e-3
    printf("[%s] conv_halo MSUB%d B%d %dx%d %4d->%4d  relL2=%.3e  %8.2f us  %7.1f TF/s  (tiles %d x %d, R=%d, A stage %d B)\n", ok ? "PASS" : "FAIL", MSUB,
           B, H, W, Cin, Cout, rel, us, tf, B * p.tiles_m_img, p.tiles_n, p.R, p.a_stage_bytes);
    fflush(stdout);
    cudaFree(dx); cudaFree(dw); cudaFree(dy); cudaFree(db); cudaFree(dref);
    return ok;
}

}  // namespace

int main() {
    if (hcp_device_check() != 0) { printf("no sm_100 device: %s\n", hcp_last_error_string()); return 2; }
    int fail = 0;
    fail += !run_case<1>(1, 16, 16, 64, 160, 1, 3);          // small, every pixel checked
    fail += !run_case<2>(1, 16, 16, 64, 160, 1, 3);
    fail += !run_case<1>(2, 32, 32, 128, 320, 1, 3);
    fail += !run_case<2>(2, 32, 32, 128, 320, 1, 3);
    // product shapes; bench_ops (nine A loads per block, MSUB 2 for the long reductions): 34.7 us / 870 TF/s, 55.8 / 1082, 35.7 / 846, 63.2 / 956
    fail += !run_case<1>(4, 64, 64, 320, 320, 13, 20);
    fail += !run_case<2>(4, 64, 64, 320, 320, 13, 20);
    fail += !run_case<2>(4, 64, 64, 640, 320, 13, 20);
    fail += !run_case<1>(4, 32, 32, 640, 640, 13, 20);
    fail += !run_case<2>(4, 32, 32, 640, 640, 13, 20);
    fail += !run_case<2>(4, 32, 32, 1280, 640, 13, 20);
    return fail ? 1 : 0;
}
