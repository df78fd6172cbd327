// Development aid: is  q' = fma(fma(-q, d, a), y, q)  with  y = rcp_rn(d), q = a * y  bit-identical to the IEEE quotient
// a / d on the domain the normalise kernel sees (0 <= a <= d <= 1, create_pb.py:93)?   nvcc -arch=sm_100a -o div_check ...
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long splitmix(unsigned long long &s)
{
    s += 0x9E3779B97F4A7C15ull;
    unsigned long long z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ float fast_div(float a, float d, float y)
{
    const float q = __fmul_rn(a, y);
    const float r = __fmaf_rn(-q, d, a);
    return __fmaf_rn(r, y, q);
}

__global__ void check(unsigned long long seed, int iters, float lo_guard, unsigned long long *mismatch, unsigned long long *tested,
                      float *example)
{
    unsigned long long s = seed + 0x1000003ull * (blockIdx.x * blockDim.x + threadIdx.x);
    unsigned long long bad = 0, n = 0;
    for (int it = 0; it < iters; ++it) {
        const unsigned long long r1 = splitmix(s), r2 = splitmix(s);
        // d: random float in (0, 1]: random exponent in [-60, 0), random mantissa (sometimes all ones / all zeros)
        unsigned mant = (unsigned)(r1 & 0x7FFFFFu);
        const unsigned sel = (unsigned)(r1 >> 60);
        if (sel == 0) mant = 0x7FFFFFu;
        if (sel == 1) mant = 0u;
        if (sel == 2) mant = 0x7FFFFEu;
        const int e = 127 - 1 - (int)((r1 >> 24) % 60);
        float d = __uint_as_float(((unsigned)e << 23) | mant);
        if (sel == 3) d = 1.0f;
        // a in [0, d]: either a random fraction of d, or d minus a few ulps, or a few ulps, or a random smaller float
        float a;
        const unsigned mode = (unsigned)(r2 >> 61);
        if (mode == 0) a = __uint_as_float(__float_as_uint(d) - (unsigned)(r2 & 7u));
        else if (mode == 1) a = d * __uint_as_float(0x3F000000u | (unsigned)(r2 & 0x7FFFFFu)) ;   // d * [0.5, 1)
        else if (mode == 2) a = d * (float)((r2 >> 8) & 0xFFFFFF) * (1.0f / 16777216.0f);
        else { const unsigned ua = (unsigned)(r2 % (unsigned long long)__float_as_uint(d)); a = __uint_as_float(ua); }
        if (!(a <= d) || a < 0.0f) continue;
        if (a != 0.0f && a < lo_guard) continue;          // the kernel sends these to the true division
        const float y = __frcp_rn(d);
        const float want = __fdiv_rn(a, d), got = fast_div(a, d, y);
        ++n;
        if (__float_as_uint(want) != __float_as_uint(got)) {
            if (bad == 0) { example[0] = a; example[1] = d; example[2] = want; example[3] = got; }
            ++bad;
        }
    }
    atomicAdd(mismatch, bad);
    atomicAdd(tested, n);
}

int main()
{
    unsigned long long *d_bad, *d_n, h_bad = 0, h_n = 0;
    float *d_ex, h_ex[4] = {0, 0, 0, 0};
    cudaMalloc(&d_bad, 8); cudaMalloc(&d_n, 8); cudaMalloc(&d_ex, 16);
    for (float guard : {0.0f, 1e-30f}) {
        cudaMemset(d_bad, 0, 8); cudaMemset(d_n, 0, 8); cudaMemset(d_ex, 0, 16);
        check<<<148 * 8, 256>>>(12345ull, 40000, guard, d_bad, d_n, d_ex);
        cudaDeviceSynchronize();
        cudaMemcpy(&h_bad, d_bad, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&h_n, d_n, 8, cudaMemcpyDeviceToHost);
        cudaMemcpy(h_ex, d_ex, 16, cudaMemcpyDeviceToHost);
        printf("guard %g: %llu pairs tested, %llu mismatches", guard, h_n, h_bad);
        if (h_bad) printf("  (e.g. a=%a d=%a want=%a got=%a)", h_ex[0], h_ex[1], h_ex[2], h_ex[3]);
        printf("\n");
    }
    return 0;
}
