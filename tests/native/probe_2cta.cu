// SPDX-License-Identifier: Apache-2.0
// CTA-pair probe (test infrastructure, NOT part of the product path; written without GPU access at the end of round 1 and
// compile-checked only -- run it first thing before building the 2-CTA GEMM):
//
//   one cluster of two CTAs computes D[256, N] = A[256, K] . B[N, K]^T with tcgen05.mma.cta_group::2:
//     * CTA r loads A rows [128 r, 128 r + 128) and B rows [N/2 r, N/2 r + N/2) into ITS OWN shared memory (K-major,
//       SWIZZLE_128B boxes of 64 k, same offsets in both CTAs) -- the B tile is NOT replicated, which is the point: the L2->SM
//       operand stream per CTA drops from (128 + N) to (128 + N/2) rows per k-block;
//     * both CTAs' TMA transactions complete on the LEADER's (rank 0) mbarrier (`cp.async.bulk.tensor...cta_group::2`);
//     * the leader's elected thread issues K/16 MMAs of shape 256 x N x 16 and commits with the multicast form so that the
//       `done` barrier of BOTH CTAs is signalled;
//     * each CTA drains its own 128 accumulator rows from its own TMEM.
//
// What the probe pins down on the real part (each is an assumption of the design in DESIGN.md "next round" item 1):
//   (1) the B descriptor of a cta_group::2 MMA addresses N/2 rows per CTA at the same CTA-relative offset;
//   (2) instruction descriptor M = 256, N = full tile width;
//   (3) tcgen05.alloc.cta_group::2 is executed by one warp of each CTA and returns the same column base in both;
//   (4) remote complete_tx + multicast commit semantics (parities, counts).
//
//   make -C tests/native probe_2cta && tests/native/probe_2cta
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

struct alignas(64) PairParams {
    CUtensorMap tmA, tmB;      // A [256 rows, K], B [N rows, K]; boxes of 64 k x 128 rows (A) / 64 k x N/2 rows (B)
    int N, K;
    float* out;                // [256, N]
};

// (the cta_group::2 wrappers this probe introduced now live in csrc/common.cuh)

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_kernel(const __grid_constant__ PairParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int nbox = p.K / 64;
    const int a_box = 128 * 128;               // 128 rows x 64 k bf16
    const int b_box = (p.N / 2) * 128;         // N/2 rows x 64 k bf16 (a multiple of 1024 when N % 16 == 0)
    uint8_t* sA = smem;                        // [nbox][128 rows][128 B]
    uint8_t* sB = smem + 4 * a_box;            // [nbox][N/2 rows][128 B]      (K <= 256)
    uint64_t* full = reinterpret_cast<uint64_t*>(sB + 4 * 128 * 128);
    uint64_t* done = full + 1;
    uint32_t* slot = reinterpret_cast<uint32_t*>(full + 2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();

    if (threadIdx.x == 0) {
        mbar_init(full, 1);                    // leader: one arrive.expect_tx; complete_tx from both CTAs
        mbar_init(done, 1);                    // one multicast commit
        fence_mbar_init();
    }
    cluster_arrive();                          // barriers of BOTH CTAs are initialised before anybody signals them
    cluster_wait();
    if (warp == 0) {                           // one warp of EACH CTA of the pair
        tmem_alloc_pair(slot, 256);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;

    if (threadIdx.x == 0) {
        const uint32_t leader_full = mapa_u32(smem_u32(full), 0);
        const uint32_t bytes_per_cta = nbox * (a_box + b_box);
        if (rank == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(full)), "r"(2 * bytes_per_cta) : "memory");
        }
        for (int i = 0; i < nbox; ++i) {
            tma_load_2d_pair(sA + i * a_box, &p.tmA, leader_full, i * 64, (int)rank * 128);
            tma_load_2d_pair(sB + i * b_box, &p.tmB, leader_full, i * 64, (int)rank * (p.N / 2));
        }
        if (rank == 0) {
            mbar_wait(full, 0);
            tc_fence_after();
            const uint32_t idesc = make_idesc_bf16(256, p.N, 0, 0);
            for (int k = 0; k < p.K / 16; ++k) {
                const uint64_t adesc = make_smem_desc(smem_u32(sA) + (k / 4) * a_box + (k % 4) * 32, 16, 1024);
                const uint64_t bdesc = make_smem_desc(smem_u32(sB) + (k / 4) * b_box + (k % 4) * 32, 16, 1024);
                umma_ss_pair(tmem, adesc, bdesc, idesc, k > 0);
            }
            umma_commit_pair(done, 0b11);
        }
    }
    __syncwarp();
    mbar_wait(done, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    for (int c = 0; c < p.N; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
        tmem_wait_ld();
        for (int j = 0; j < 16; ++j)
            if (c + j < p.N) p.out[(rank * 128 + row) * p.N + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    cluster_arrive();                          // neither CTA frees TMEM / exits while the peer may still touch the pair's state
    cluster_wait();
    if (warp == 0) tmem_dealloc_pair(tmem, 256);
}

float bf16_round(float f) { return __bfloat162float(__float2bfloat16(f)); }

bool run_case(int N, int K) {
    const int M = 256;
    std::vector<float> A(M * K), B(N * K);
    std::vector<__nv_bfloat16> Ag(M * K), Bg(N * K);
    srand(4321);
    for (int i = 0; i < M * K; ++i) { A[i] = bf16_round((rand() % 2001 - 1000) / 1000.f); Ag[i] = __float2bfloat16(A[i]); }
    for (int i = 0; i < N * K; ++i) { B[i] = bf16_round((rand() % 2001 - 1000) / 1000.f); Bg[i] = __float2bfloat16(B[i]); }
    __nv_bfloat16 *dA, *dB;
    float* dOut;
    cudaMalloc(&dA, M * K * 2); cudaMalloc(&dB, N * K * 2); cudaMalloc(&dOut, M * N * 4);
    cudaMemcpy(dA, Ag.data(), M * K * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, Bg.data(), N * K * 2, cudaMemcpyHostToDevice);
    cudaMemset(dOut, 0xFF, M * N * 4);
    PairParams p;
    memset(&p, 0, sizeof(p));
    p.N = N; p.K = K; p.out = dOut;
    if (make_tmap_2d(&p.tmA, dA, K, M, K, 64, 128) || make_tmap_2d(&p.tmB, dB, K, N, K, 64, N / 2)) {
        printf("[FAIL] pair N=%d K=%d: tensor map: %s\n", N, K, hcp_last_error_string());
        return false;
    }
    const int smem = 4 * 128 * 128 * 2 + 1024 + 1024;
    cudaFuncSetAttribute(pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    pair_kernel<<<2, 128, smem>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[FAIL] pair N=%d K=%d: kernel error %s\n", N, K, cudaGetErrorString(e)); exit(3); }
    std::vector<float> out(M * N);
    cudaMemcpy(out.data(), dOut, M * N * 4, cudaMemcpyDeviceToHost);
    double err2 = 0, ref2 = 0;
    for (int m = 0; m < M; ++m)
        for (int n = 0; n < N; ++n) {
            double r = 0;
            for (int k = 0; k < K; ++k) r += (double)A[m * K + k] * B[n * K + k];
            double d = isnan(out[m * N + n]) ? 1e3 : out[m * N + n] - r;
            err2 += d * d; ref2 += r * r;
        }
    const double rel = sqrt(err2 / (ref2 + 1e-30));
    printf("[%s] cta_group::2  256 x %3d x %3d   relL2=%.3e\n", rel < 1e-3 ? "PASS" : "FAIL", N, K, rel);
    fflush(stdout);
    cudaFree(dA); cudaFree(dB); cudaFree(dOut);
    return rel < 1e-3;
}

}  // namespace

int main() {
    if (hcp_device_check() != 0) { printf("no sm_100 device: %s\n", hcp_last_error_string()); return 2; }
    int fail = 0;
    const int cases[][2] = {{64, 64}, {128, 64}, {128, 256}, {160, 128}, {256, 256}};
    for (const auto& c : cases) fail += run_case(c[0], c[1]) ? 0 : 1;
    return fail ? 1 : 0;
}
