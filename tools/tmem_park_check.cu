// Development aid for the next step of the fused PRN ("what comes next"): can a weight k-block be PARKED in
// tensor memory and used from there as the A operand?  One CTA: A [128 x 64] and B [N x 64] bf16, K-major, 128-byte
// swizzle in shared memory (the layout the kernel's TMA boxes have).
//   D0 = A B^T with both operands from shared memory (what prn_fused_kernel does today);
//   D1 = the same product with every 128 x 16 slice of A first copied into TMEM (tcgen05.cp.128x256b, the same descriptor
//        the MMA would take) and the MMA reading A from there (tcgen05.mma ... [d], [a], b-desc).
// Prints max |D0 - reference| and whether D1 == D0 bit for bit.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I multiposenet_b200/csrc -o tools/tmem_park_check tools/tmem_park_check.cu -lcuda
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>

#include "tcgen05_utils.cuh"

using namespace mpn::tc;

constexpr int kM = 128, kN = 80, kK = 64;

__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_cp_128x256b(uint32_t tmem_dst, uint64_t sdesc)
{
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_dst), "l"(sdesc) : "memory");
}

// element (row, k) of a K-major tile with 128-byte rows and the 128-byte swizzle (16-byte chunk index XOR row % 8)
__device__ __forceinline__ int sw128_offset(int row, int k)
{
    const int chunk = (k * 2) >> 4, within = (k * 2) & 15;
    return row * 128 + ((chunk ^ (row & 7)) << 4) + within;
}

__global__ void __launch_bounds__(128, 1) check_kernel(const __nv_bfloat16 *A, const __nv_bfloat16 *B, float *D0, float *D1)
{
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sA = smem, *sB = smem + kM * 128;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sB + 128 * 128);
    uint32_t *slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < kM * kK; i += 128) *reinterpret_cast<__nv_bfloat16 *>(sA + sw128_offset(i / kK, i % kK)) = A[i];
    for (int i = tid; i < kN * kK; i += 128) *reinterpret_cast<__nv_bfloat16 *>(sB + sw128_offset(i / kK, i % kK)) = B[i];
    if (tid == 0) { mbar_init(bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *slot;
    const uint32_t d0 = tmem, d1 = tmem + 128, ta = tmem + 256;
    if (tid == 0) {
        const uint32_t idesc = make_idesc_bf16(kM, kN);
        const uint64_t adesc = make_kmajor_sw128_desc(smem_u32(sA)), bdesc = make_kmajor_sw128_desc(smem_u32(sB));
        for (int k = 0; k < kK / UMMA_K; ++k) umma_bf16(d0, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, k > 0);
        for (int k = 0; k < kK / UMMA_K; ++k) tmem_cp_128x256b(ta + (uint32_t)(8 * k), adesc + (uint64_t)(2 * k));
        for (int k = 0; k < kK / UMMA_K; ++k) umma_bf16_ts(d1, ta + (uint32_t)(8 * k), bdesc + (uint64_t)(2 * k), idesc, k > 0);
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    for (int c0 = 0; c0 < kN; c0 += 16) {
        uint32_t r[16];
        tmem_ld16(d0 + lane_base + c0, r);
        for (int j = 0; j < 16; ++j) D0[(warp * 32 + lane) * kN + c0 + j] = __uint_as_float(r[j]);
        tmem_ld16(d1 + lane_base + c0, r);
        for (int j = 0; j < 16; ++j) D1[(warp * 32 + lane) * kN + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main()
{
    std::vector<__nv_bfloat16> hA(kM * kK), hB(kN * kK);
    std::vector<float> fA(kM * kK), fB(kN * kK), ref(kM * kN), h0(kM * kN), h1(kM * kN);
    srand(7);
    for (int i = 0; i < kM * kK; ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < kN * kK; ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.0f); fB[i] = __bfloat162float(hB[i]); }
    for (int m = 0; m < kM; ++m)
        for (int n = 0; n < kN; ++n) {
            double s = 0;
            for (int k = 0; k < kK; ++k) s += (double)fA[m * kK + k] * fB[n * kK + k];
            ref[m * kN + n] = (float)s;
        }
    __nv_bfloat16 *dA, *dB; float *d0, *d1;
    cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&d0, ref.size() * 4); cudaMalloc(&d1, ref.size() * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(d0, 0, ref.size() * 4); cudaMemset(d1, 0xff, ref.size() * 4);
    const int smem = kM * 128 + 128 * 128 + 64 + 1024;
    cudaFuncSetAttribute(check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    check_kernel<<<1, 128, smem>>>(dA, dB, d0, d1);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    cudaMemcpy(h0.data(), d0, ref.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h1.data(), d1, ref.size() * 4, cudaMemcpyDeviceToHost);
    double e0 = 0, e1 = 0; int diff = 0;
    for (size_t i = 0; i < ref.size(); ++i) {
        e0 = fmax(e0, fabs((double)h0[i] - ref[i])); e1 = fmax(e1, fabs((double)h1[i] - ref[i]));
        diff += memcmp(&h0[i], &h1[i], 4) != 0;
    }
    printf("A from shared memory: max |D - ref| = %.3g\nA parked in TMEM:     max |D - ref| = %.3g, %d of %zu elements differ from the shared-memory result\n",
           e0, e1, diff, ref.size());
    return 0;
}
