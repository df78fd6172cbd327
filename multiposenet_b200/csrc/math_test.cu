// Element-wise hooks that expose the device exp / sigmoid of the path to the bit-level parity tests.
#include "common.cuh"
#include "mpn_math.cuh"

namespace mpn {
namespace {
__global__ void test_math_kernel(const float *__restrict__ x, float *__restrict__ y, const long long n, const int which)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    y[i] = which == 0 ? exact_expf(x[i]) : exact_sigmoidf(x[i]);
}

// ordered key -> float: keys below 2^31 are the negative floats in increasing order, the rest the positive ones
__device__ __forceinline__ float float_of_key(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

// Every thread walks 64 consecutive keys: counts neighbours x < x' (both finite) with sigmoid(x) > sigmoid(x'), and any
// value on which the packed-pair form of the recipe (heatmap kernels) differs from the scalar form.
__global__ void test_monotone_kernel(const unsigned key_begin, const unsigned long long count, unsigned long long *violations)
{
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned long long i0 = t * 64ull;
    if (i0 >= count) return;
    unsigned long long bad = 0;
    const unsigned long long i1 = i0 + 64ull < count ? i0 + 64ull : count;
    for (unsigned long long i = i0; i < i1; ++i) {
        const unsigned long long k = (unsigned long long)key_begin + i;
        if (k + 1ull > 0xffffffffull) break;
        const float x0 = float_of_key((unsigned)k), x1 = float_of_key((unsigned)(k + 1ull));
        if (!(fabsf(x0) <= 3.4028234e38f) || !(fabsf(x1) <= 3.4028234e38f)) continue;      // NaN / inf keys
        float p0, p1;
        exact_sigmoidf_pair(x0, x1, p0, p1);
        const float s0 = exact_sigmoidf(x0), s1 = exact_sigmoidf(x1);
        bad += (s0 > s1) + (__float_as_uint(p0) != __float_as_uint(s0)) + (__float_as_uint(p1) != __float_as_uint(s1));
    }
    if (bad) atomicAdd(violations, bad);
}
}  // namespace

int launch_test_monotone(unsigned key_begin, unsigned long long count, unsigned long long *violations, cudaStream_t s)
{
    if (count == 0) return 0;
    const unsigned long long threads = (count + 63ull) / 64ull;
    test_monotone_kernel<<<(unsigned)((threads + 255ull) / 256ull), 256, 0, s>>>(key_begin, count, violations);
    return 1;
}

int launch_test_math(const float *x, float *y, int64_t n, int which, cudaStream_t s)
{
    if (n == 0) return 0;
    test_math_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(x, y, (long long)n, which);
    return 1;
}
}  // namespace mpn
