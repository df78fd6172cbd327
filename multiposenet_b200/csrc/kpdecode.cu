// Spatial softmax + argmax keypoint decode.
//
// Replaces create_pb.py:115-142: logits [N, 56*36, 17] -> softmax over the 2016 positions of every (person, channel)
// (tf.nn.softmax(axis=1): p = exp(l - max) * (1 / sum)), keypoint_scores = max p, keypoint_positions =
// (argmax // 36 / 56, argmax % 36 / 36) with tf.argmax's first-index tie rule.
//
// One thread-block CLUSTER of 4 CTAs per person, each CTA of 17 x 32 threads owning a quarter of the 2016 positions
// (a contiguous 34 KB slab of the person's logit row, read exactly once, fully coalesced; 16 values per thread stay in
// registers for the second pass).  The per-channel maxima, exp-sums and first-argmax candidates of the four CTAs are
// combined through distributed shared memory (two cluster barriers); softmax probabilities are never written.
//
// argmax rule, bit-matched to the oracle without depending on the summation order of the denominator:
// p[i] = e[i] * r with r = 1/sum and e[i] = exp(l[i] - lmax) <= 1.  e[i] * r == r iff e[i] == 1.0f (for e[i] <= 1 - 2^-24
// the product is at least half an ulp below r and rounds to a smaller float), so the first index with maximal
// probability is the first index whose e[i] is exactly 1.0f.  The exp recipe is monotone near 0, so "exp(d) == 1.0f"
// is the comparison d >= x0 with x0 the most negative float whose recipe value is 1.0f (found once per device by
// bisection with the recipe itself, kpdecode_prepare): the decision costs one subtraction and one compare per logit.
// keypoint_scores = 1 / sum is only held to 1e-4, so the 2016-term sum uses the hardware ex2 approximation (~2 ulp per
// term) instead of the 20-instruction exact recipe -- the kernel was instruction-issue bound on it.
#include "common.cuh"
#include "mpn_math.cuh"

namespace mpn {

namespace {

constexpr int kNK = 17;
constexpr int kLanes = 32;
constexpr int kThreads = kNK * kLanes;   // 544
constexpr int kCluster = 4;
constexpr int kMaxPerThread = 16;        // positions per thread: ceil(2048 / 4 / 32)

__device__ float g_exp_one_x0;     // most negative x with exact_expf(x) == 1.0f

__global__ void exp_one_threshold_kernel()
{
    // negative floats are ordered by their magnitude bits: bisect the largest magnitude that still gives exactly 1.0f
    unsigned lo = 0x80000000u, hi = 0xBF800000u;      // -0.0f (gives 1) .. -1.0f (does not)
    while (hi - lo > 1u) {
        const unsigned mid = lo + (hi - lo) / 2u;
        if (exact_expf(__uint_as_float(mid)) == 1.0f) lo = mid; else hi = mid;
    }
    g_exp_one_x0 = __uint_as_float(lo);
}

__device__ __forceinline__ unsigned cluster_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// address of `p` (a shared-memory object of this CTA) in CTA `rank` of the cluster
template <typename T>
__device__ __forceinline__ const T *peer_shared(const T *p, unsigned rank)
{
    unsigned long long out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(reinterpret_cast<unsigned long long>(p)), "r"(rank));
    return reinterpret_cast<const T *>(out);
}

__global__ void __cluster_dims__(kCluster, 1, 1) __launch_bounds__(kThreads)
keypoint_decode_kernel(const float *__restrict__ logits, const int *__restrict__ n_dev, const int n_host, const int crop_h,
                       const int crop_w, float *__restrict__ scores, float *__restrict__ positions,
                       int *__restrict__ argmax_out)
{
    __shared__ float s_f[kNK][kLanes + 1];
    __shared__ int s_i[kNK][kLanes + 1];
    __shared__ float s_max[kNK];          // this CTA's per-channel maximum (read by the peers)
    __shared__ float s_sum[kNK];          // this CTA's per-channel sum of exp(l - global max)
    __shared__ int s_first[kNK];          // this CTA's first position with exp(l - global max) == 1
    __shared__ float s_gmax[kNK];
    const int n = blockIdx.x / kCluster;
    const unsigned rank = cluster_rank();
    pdl_trigger();
    pdl_wait();                           // the PRN has completed
    const int N = n_dev ? *n_dev : n_host;
    if (n >= N) return;                   // uniform over the cluster
    const int P = crop_h * crop_w;
    const int per = (P + kCluster - 1) / kCluster;
    const int p0 = (int)rank * per, p1 = min(P, p0 + per);
    const int tid = threadIdx.x, c = tid % kNK, q = tid / kNK;
    const float *row = logits + (size_t)n * P * kNK + (size_t)p0 * kNK;
    float v[kMaxPerThread];
    float lmax = -__int_as_float(0x7f800000);
#pragma unroll
    for (int i = 0; i < kMaxPerThread; ++i) {
        const int p = p0 + q + kLanes * i;
        v[i] = (p < p1) ? __ldcs(row + (size_t)tid + (size_t)kThreads * i) : -__int_as_float(0x7f800000);
        lmax = fmaxf(lmax, v[i]);
    }
    const int warp = tid >> 5, lane = tid & 31;       // 17 warps: warp w folds the 32 partials of channel w
    s_f[c][q] = lmax;
    __syncthreads();
    {
        float m = s_f[warp][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) s_max[warp] = m;
    }
    cluster_sync_all();                   // every CTA's s_max is visible cluster-wide
    if (tid < kNK) {
        float m = s_max[tid];
        for (unsigned r = 0; r < kCluster; ++r)
            if (r != rank) m = fmaxf(m, *peer_shared(&s_max[tid], r));
        s_gmax[tid] = m;
    }
    __syncthreads();
    lmax = s_gmax[c];
    const float x0 = g_exp_one_x0;
    float sum = 0.0f;
    int first = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < kMaxPerThread; ++i) {
        const int p = p0 + q + kLanes * i;
        if (p < p1) {
            const float d = fsub(v[i], lmax);
            sum = fadd(sum, __expf(d));
            if (d >= x0 && p < first) first = p;       // <=> exact_expf(d) == 1.0f
        }
    }
    s_f[c][q] = sum; s_i[c][q] = first;
    __syncthreads();
    {
        float S = s_f[warp][lane]; int best = s_i[warp][lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {          // butterfly: a fixed summation tree
            S = fadd(S, __shfl_xor_sync(0xffffffffu, S, o));
            best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        }
        if (lane == 0) { s_sum[warp] = S; s_first[warp] = best; }
    }
    cluster_sync_all();                   // every CTA's s_sum / s_first is visible cluster-wide
    if (rank == 0 && tid < kNK) {
        float S = s_sum[tid]; int best = s_first[tid];
        for (unsigned r = 1; r < kCluster; ++r) {          // fixed order: CTA 0, 1, 2, 3
            S = fadd(S, *peer_shared(&s_sum[tid], r));
            best = min(best, *peer_shared(&s_first[tid], r));
        }
        if (best == 0x7fffffff) best = 0;    // only with NaN logits
        const float rcp = fdiv(1.0f, S);
        const size_t o = (size_t)n * kNK + tid;
        scores[o] = fmul(1.0f, rcp);
        positions[o * 2 + 0] = fdiv((float)(best / crop_w), (float)crop_h);
        positions[o * 2 + 1] = fdiv((float)(best % crop_w), (float)crop_w);
        if (argmax_out) argmax_out[o] = best;
    }
    cluster_sync_all();                   // peers stay alive until CTA 0 has read their shared memory
}

}  // namespace

int kpdecode_prepare(cudaStream_t s)
{
    exp_one_threshold_kernel<<<1, 1, 0, s>>>();
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

int launch_keypoint_decode(const float *logits, const int *n_dev, int n_host, int n_max, int crop_h, int crop_w,
                           float *scores, float *positions, int *argmax, cudaStream_t s)
{
    if (n_max <= 0) return 0;
    if (crop_h * crop_w > kMaxPerThread * kLanes * kCluster) return -(int)cudaErrorInvalidValue;
    prof_mark(s, "keypoint_decode");
    launch_k(keypoint_decode_kernel, dim3(n_max * kCluster), dim3(kThreads), 0, s, true, logits, n_dev, n_host, crop_h, crop_w,
             scores, positions, argmax);
    return 1;
}

}  // namespace mpn
