// Spatial softmax + argmax keypoint decode.
//
// Replaces create_pb.py:115-142: logits [N, 56*36, 17] -> softmax over the 2016 positions of every (person, channel)
// (tf.nn.softmax(axis=1): p = exp(l - max) * (1 / sum)), keypoint_scores = max p, keypoint_positions =
// (argmax // 36 / 56, argmax % 36 / 36) with tf.argmax's first-index tie rule.
//
// One CTA of 17 x 32 threads per person; thread (c, q) owns positions q, q+32, ... of channel c, so the CTA reads the
// person's 137 KB logit row exactly once, fully coalesced (flat index = thread + 544 * i), and keeps its 63 values in
// registers for the second pass.  Softmax probabilities are never written.
//
// argmax rule, bit-matched to the oracle without depending on the summation order of the denominator:
// p[i] = e[i] * r with r = 1/sum and e[i] = exp(l[i] - lmax) <= 1.  e[i] * r == r iff e[i] == 1.0f (for e[i] <= 1 - 2^-24
// the product is at least half an ulp below r and rounds to a smaller float), so the first index with maximal
// probability is the first index whose e[i] is exactly 1.0f.  keypoint_scores = 1.0f * r.
#include "common.cuh"
#include "mpn_math.cuh"

namespace mpn {

namespace {

constexpr int kNK = 17;
constexpr int kLanes = 32;
constexpr int kThreads = kNK * kLanes;   // 544
constexpr int kMaxPerThread = 64;        // positions per thread held in registers (56*36/32 = 63)

__global__ void __launch_bounds__(kThreads) keypoint_decode_kernel(const float *__restrict__ logits,
                                                                   const int *__restrict__ n_dev, const int n_host,
                                                                   const int crop_h, const int crop_w,
                                                                   float *__restrict__ scores,
                                                                   float *__restrict__ positions,
                                                                   int *__restrict__ argmax_out)
{
    __shared__ float s_f[kNK][kLanes + 1];
    __shared__ int s_i[kNK][kLanes + 1];
    __shared__ float s_max[kNK];
    const int n = blockIdx.x;
    const int N = n_dev ? *n_dev : n_host;
    if (n >= N) return;
    const int P = crop_h * crop_w;
    const int tid = threadIdx.x, c = tid % kNK, q = tid / kNK;
    const float *row = logits + (size_t)n * P * kNK;
    float v[kMaxPerThread];
    float lmax = -__int_as_float(0x7f800000);
#pragma unroll
    for (int i = 0; i < kMaxPerThread; ++i) {
        const int p = q + kLanes * i;
        v[i] = (p < P) ? __ldg(row + (size_t)tid + (size_t)kThreads * i) : -__int_as_float(0x7f800000);
        lmax = fmaxf(lmax, v[i]);
    }
    s_f[c][q] = lmax;
    __syncthreads();
    if (tid < kNK) {
        float m = s_f[tid][0];
        for (int i = 1; i < kLanes; ++i) m = fmaxf(m, s_f[tid][i]);
        s_max[tid] = m;
    }
    __syncthreads();
    lmax = s_max[c];
    float sum = 0.0f;
    int first = 0x7fffffff;
#pragma unroll
    for (int i = 0; i < kMaxPerThread; ++i) {
        const int p = q + kLanes * i;
        if (p < P) {
            const float e = exact_expf(fsub(v[i], lmax));
            sum = fadd(sum, e);
            if (e == 1.0f && p < first) first = p;
        }
    }
    __syncthreads();
    s_f[c][q] = sum; s_i[c][q] = first;
    __syncthreads();
    if (tid < kNK) {
        float S = 0.0f; int best = 0x7fffffff;
        for (int i = 0; i < kLanes; ++i) { S = fadd(S, s_f[tid][i]); best = min(best, s_i[tid][i]); }
        if (best == 0x7fffffff) best = 0;    // only with NaN logits
        const float r = fdiv(1.0f, S);
        const size_t o = (size_t)n * kNK + tid;
        scores[o] = fmul(1.0f, r);
        positions[o * 2 + 0] = fdiv((float)(best / crop_w), (float)crop_h);
        positions[o * 2 + 1] = fdiv((float)(best % crop_w), (float)crop_w);
        if (argmax_out) argmax_out[o] = best;
    }
}

}  // namespace

int launch_keypoint_decode(const float *logits, const int *n_dev, int n_host, int n_max, int crop_h, int crop_w,
                           float *scores, float *positions, int *argmax, cudaStream_t s)
{
    if (n_max <= 0) return 0;
    if (crop_h * crop_w > kMaxPerThread * kLanes) return -(int)cudaErrorInvalidValue;
    prof_mark(s, "keypoint_decode");
    keypoint_decode_kernel<<<n_max, kThreads, 0, s>>>(logits, n_dev, n_host, crop_h, crop_w, scores, positions, argmax);
    return 1;
}

}  // namespace mpn
