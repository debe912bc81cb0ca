// SPDX-License-Identifier: Apache-2.0
// Row-shifted A operand probe (test infrastructure, NOT part of the product path).  Result on B200 (profiles/r01_native_probe_shift.log):
// every shift is exact with the base-offset field LEFT AT 0 and wrong with it set -- the swizzle depends on the absolute address only.
//
// Question: can ONE shared-memory copy of an activation halo serve all nine taps of a 3x3 convolution?  The implicit-GEMM kernel
// loads the A tile (128 pixels x 64 channels, K-major, SWIZZLE_128B) nine times per 64-channel block, once per tap, and is bound by
// the L2->SM operand stream.  If the pixels of a tile are laid out in "virtual" order (image rows with pitch W + 2, borders zero
// filled by TMA), tap (dh, dw) of the same tile is the SAME buffer read from a start address shifted by (dh (W + 2) + dw) rows of
// 128 bytes.  Such a start is not aligned to the 1024-byte swizzle atom; the UMMA shared-memory descriptor has a 3-bit
// "matrix base offset" field (bits 49-51) for this.  The probe loads A' with 128 + PAD rows through one TMA box per 64-wide k block
// and checks  D = A'[s : s + 128] . B^T  for several shifts s, with the base offset set to (start_address >> 7) & 7 and, as a
// control, left at 0.
//
//   make -C tests/native probe_shift && tests/native/probe_shift
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

constexpr int PAD = 72;                 // extra rows below the tile: shifts up to 72 (a multiple of 8 keeps the box a whole number of atoms)
constexpr int ROWS = 128 + PAD;

struct alignas(64) ShiftParams {
    CUtensorMap tmA, tmB;               // A' [ROWS, K] box 64 x ROWS;  B [N, K] box 64 x N
    int N, K, shift, use_base_offset;
    float* out;                         // [128, N]
};

__global__ void __launch_bounds__(128, 1) shift_kernel(const __grid_constant__ ShiftParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nbox = p.K / 64;
    const int a_box = ROWS * 128, b_box = p.N * 128;
    uint8_t* sA = smem;                              // [nbox][ROWS][128 B]
    uint8_t* sB = smem + 2 * a_box;                  // [nbox][N][128 B]           (K <= 128)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 2 * 128 * 128);
    uint64_t* done = bar + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        mbar_init(done, 1);
        fence_mbar_init();
    }
    if (warp == 0) { tmem_alloc(slot, 128); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, nbox * (a_box + b_box));
        for (int i = 0; i < nbox; ++i) {
            tma_load_2d(sA + i * a_box, &p.tmA, bar, i * 64, 0);
            tma_load_2d(sB + i * b_box, &p.tmB, bar, i * 64, 0);
        }
        mbar_wait(bar, 0);
        tc_fence_after();
        const uint32_t idesc = make_idesc_bf16(128, p.N, 0, 0);
        for (int k = 0; k < p.K / 16; ++k) {
            const uint32_t a_addr = smem_u32(sA) + (k / 4) * a_box + p.shift * 128 + (k % 4) * 32;
            uint64_t adesc = make_smem_desc(a_addr, 16, 1024);
            if (p.use_base_offset) adesc |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;   // matrix base offset
            const uint64_t bdesc = make_smem_desc(smem_u32(sB) + (k / 4) * b_box + (k % 4) * 32, 16, 1024);
            umma_ss(tmem, adesc, bdesc, idesc, k > 0);
        }
        umma_commit(done);
    }
    __syncwarp();
    mbar_wait(done, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c = 0; c < p.N; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
        tmem_wait_ld();
        for (int j = 0; j < 16; ++j) p.out[row * p.N + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

float bf16_round(float f) { return __bfloat162float(__float2bfloat16(f)); }

bool run_case(int N, int K, int shift, int use_base_offset) {
    std::vector<float> A(ROWS * K), B(N * K);
    std::vector<__nv_bfloat16> Ag(ROWS * K), Bg(N * K);
    srand(99);
    for (int i = 0; i < ROWS * K; ++i) { A[i] = bf16_round((rand() % 2001 - 1000) / 1000.f); Ag[i] = __float2bfloat16(A[i]); }
    for (int i = 0; i < N * K; ++i) { B[i] = bf16_round((rand() % 2001 - 1000) / 1000.f); Bg[i] = __float2bfloat16(B[i]); }
    __nv_bfloat16 *dA, *dB;
    float* dOut;
    cudaMalloc(&dA, ROWS * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dOut, 128 * N * 4);
    cudaMemcpy(dA, Ag.data(), ROWS * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bg.data(), N * K * 2, cudaMemcpyHostToDevice);
    cudaMemset(dOut, 0xFF, 128 * N * 4);
    ShiftParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.K = K; p.shift = shift; p.use_base_offset = use_base_offset; p.out = dOut;
    if (make_tmap_2d(&p.tmA, dA, K, ROWS, K, 64, ROWS) || make_tmap_2d(&p.tmB, dB, K, N, K, 64, N)) {
        printf("[FAIL] shift probe: tensor map: %s\n", hcp_last_error_string());
        return false;
    }
    const int smem = 2 * ROWS * 128 + 2 * 128 * 128 + 1024 + 1024;
    cudaFuncSetAttribute(shift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    shift_kernel<<<1, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[FAIL] shift probe: kernel error %s\n", cudaGetErrorString(e)); exit(3); }
    std::vector<float> out(128 * N);
    cudaMemcpy(out.data(), dOut, 128 * N * 4, cudaMemcpyDeviceToHost);
    double err2 = 0, ref2 = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < K; ++k) r += (double)A[(m + shift) * K + k] * B[n * K + k];
            double d = isnan(out[m * N + n]) ? 1e3 : out[m * N + n] - r;
            err2 += d * d; ref2 += r * r;
        }
    const double rel = sqrt(err2 / (ref2 + 1e-30));
    printf("[%s] A rows [%2d, %3d)  N=%3d K=%3d  base_offset %s   relL2=%.3e\n", rel < 1e-3 ? "PASS" : "FAIL", shift, shift + 128, N, K,
           use_base_offset ? "set " : "zero", rel);
    fflush(stdout);
    cudaFree(dA); cudaFree(dB); cudaFree(dOut);
    return rel < 1e-3;
}

}  // namespace

int main() {
    if (hcp_device_check() != 0) { printf("no sm_100 device: %s\n", hcp_last_error_string()); return 2; }
    int pass = 0, total = 0;
    const int shifts[] = {0, 8, 1, 3, 7, 9, 65, 66, 67};        // 65..67 = (W + 2) +- 1 for a 64-wide image
    for (int s : shifts)
        for (int bo = 1; bo >= 0; --bo) {
            if (s % 8 == 0 && bo == 0) continue;                // aligned start: the field is 0 either way
            ++total; pass += run_case(128, 128, s, bo) ? 1 : 0;
        }
    ++total; pass += run_case(160, 64, 67, 1) ? 1 : 0;
    printf("%d of %d configurations match (informational: the 'zero' rows are the control)\n", pass, total);
    return 0;
}
