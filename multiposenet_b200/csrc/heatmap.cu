// Heatmap activation + per-(image, channel) min/max, crop_and_resize with the min-max normalisation folded into the
// taps, and the stand-alone numpy decoder get_keypoints.
//
// Replaces (reference):
//   create_pb.py:73-76    keypoint_heatmaps = sigmoid(heatmaps[..., :17]); segmentation_masks = heatmaps[..., 17]
//   create_pb.py:90-94    M = reduce_max, m = reduce_min over (h, w); hm = (hm - m) / (M - m) * float(M > 0.2)
//   create_pb.py:106-109  tf.image.crop_and_resize(heatmaps, boxes, box_ind, crop_size=[56, 36])  (bilinear, extrapolation 0)
//   inference/utils.py:29-52  get_keypoints
//
// The heatmap kernels are HBM streams over 64-pixel tiles (64 x 18 = 1152 consecutive floats of the NHWC logits) by CTAs
// of 288 threads.  Thread t owns elements t, t + 288, t + 576, t + 864 of a tile: 288 = 16 x 18, so ALL of them belong to
// channel t mod 18 (pixels t / 18 + 16 k) -- one channel per thread, its running min / max or its normalisation constants
// in a handful of registers -- and a warp's 32 lanes touch 32 consecutive floats in every load and every store (one
// wavefront each; an earlier float4-per-thread map made every scalar store a stride-4 access of four wavefronts).
#include "bulk_copy.cuh"
#include "common.cuh"
#include "mpn_math.cuh"

namespace mpn {

namespace {

constexpr int kNK = 17;             // keypoint channels
constexpr int kCH = 18;             // channels of the subnet output
constexpr int kHmThreads = 288;     // 9 warps: 288 float4 = 64 pixels x 18 channels
constexpr int kHmPix = 64;
constexpr int kPadCh = 20;          // channels per pixel of the padded layouts (normalised map, conv kernel rows)
constexpr int kGroups = kPadCh / 4;

// Tail of the heatmap kernels.  publish_minmax: the CTA's per-channel (min, max) in s_min / s_max go to the partial array;
// the last CTA of the image (threadfence + counter) folds all CTAs' values into minmax[img] and re-arms the counter.
__device__ __forceinline__ void publish_minmax(int *s_min, int *s_max, int *s_last, int *__restrict__ partial,
                                               unsigned int *__restrict__ counter, int *__restrict__ minmax)
{
    const int img = blockIdx.y, tid = threadIdx.x;
    int *my = partial + ((size_t)img * gridDim.x + blockIdx.x) * kNK * 2;
    if (tid < kNK) { my[tid * 2] = s_min[tid]; my[tid * 2 + 1] = s_max[tid]; }
    __threadfence();
    __syncthreads();
    if (tid == 0) *s_last = (atomicAdd(counter + img, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (!*s_last) return;
    __threadfence();
    if (tid < kNK) { s_min[tid] = 0x7f800000; s_max[tid] = 0; }
    __syncthreads();
    if (tid < kNK * 8) {
        const int c = tid % kNK, slice = tid / kNK;
        int lo = 0x7f800000, hi = 0;
        for (int i = slice; i < (int)gridDim.x; i += 8) {
            const int *q = partial + ((size_t)img * gridDim.x + i) * kNK * 2 + c * 2;
            lo = min(lo, __ldcg(q)); hi = max(hi, __ldcg(q + 1));
        }
        atomicMin(&s_min[c], lo); atomicMax(&s_max[c], hi);
    }
    __syncthreads();
    if (tid < kNK) {
        minmax[((size_t)img * kNK + tid) * 2] = s_min[tid];
        minmax[((size_t)img * kNK + tid) * 2 + 1] = s_max[tid];
    }
    if (tid == 0) counter[img] = 0u;
}

constexpr int kPerThread = kHmPix * kCH / kHmThreads;      // 4 elements of a tile per thread
constexpr int kRingTiles = 6;                              // tiles per CTA in the bulk-copy ring (even: two per trip)

// One pass (maps that take the per-tap crop path): activation, split, per-(image, channel) min / max.  grid = (chunks per
// image, B).  Outputs are written straight from registers; per-CTA (min, max) go to a small partial array; the last CTA of
// each image (threadfence + counter) folds them into minmax[b] and re-arms the counter, so no reset kernel is needed.
__global__ void __launch_bounds__(kHmThreads) heatmap_kernel(const float *__restrict__ hml, const int npix,
                                                             const int tiles_per_img, float *__restrict__ kh,
                                                             float *__restrict__ seg, int *__restrict__ partial,
                                                             unsigned int *__restrict__ counter,
                                                             int *__restrict__ minmax)
{
    __shared__ int s_min[kNK], s_max[kNK];
    __shared__ int s_last;
    __shared__ __align__(128) float s_tile[kRingTiles][kHmPix * kCH];
    __shared__ unsigned long long s_bar[kRingTiles];
    const int img = blockIdx.y, tid = threadIdx.x;
    pdl_trigger();
    const int c = tid % kCH, p0 = tid / kCH;
    const bool is_kp = c < kNK;
    float mn = __int_as_float(0x7f800000), mx = 0.0f;
    if (tid < kNK) { s_min[tid] = 0x7f800000; s_max[tid] = 0; }
    // The CTA's tiles (blockIdx.x, + gridDim.x, ...) arrive through a ring of kRingTiles bulk copies (bulk_copy.cuh): four
    // to six tiles -- 18 to 27 KB per CTA, four CTAs per SM -- are in flight while two are evaluated, which is what the
    // stream needs against HBM latency (eight register loads per thread and trip kept it at half of that).
    // Output pointers run along (tile j and tile j + 1 of this CTA, advanced by two grid strides per trip): every store is
    // base + immediate.  A keypoint thread writes keypoint_heatmaps, a mask-channel thread segmentation_masks.
    const size_t g = gridDim.x;
    const float *img_src = hml + (size_t)img * npix * kCH;
    const int n_mine = (int)blockIdx.x < tiles_per_img ? (tiles_per_img - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (tid == 0) {
        bulk_barrier_init(s_bar, kRingTiles);
        for (int j = 0; j < kRingTiles && j < n_mine; ++j)
            bulk_fetch(s_tile[j], img_src + ((size_t)blockIdx.x + j * g) * (kHmPix * kCH), kHmPix * kCH * 4, &s_bar[j]);
    }
    float *oa = is_kp ? kh + ((size_t)img * npix + (size_t)blockIdx.x * kHmPix + p0) * kNK + c
                      : (seg ? seg + (size_t)img * npix + (size_t)blockIdx.x * kHmPix + p0 : nullptr);
    const size_t ostep = is_kp ? kHmPix * kNK : kHmPix;
    float *ob = oa ? oa + g * ostep : nullptr;
    __syncthreads();                                   // barriers initialised, s_min / s_max reset
    // npix is a multiple of 64 on this path (images are multiples of 128): every tile is full.  Two tiles per trip.
    for (int j = 0; j < n_mine; j += 2) {
        const bool two = j + 1 < n_mine;
        const int sl_a = j % kRingTiles, sl_b = (j + 1) % kRingTiles;
        bulk_wait(&s_bar[sl_a], (unsigned)(j / kRingTiles) & 1u);
        if (two) bulk_wait(&s_bar[sl_b], (unsigned)((j + 1) / kRingTiles) & 1u);
        float va[kPerThread], vb[kPerThread];
#pragma unroll
        for (int k = 0; k < kPerThread; ++k) {
            va[k] = s_tile[sl_a][k * kHmThreads + tid];
            vb[k] = two ? s_tile[sl_b][k * kHmThreads + tid] : va[k];      // no second tile: the first again (min / max unchanged)
        }
        __syncthreads();                               // both slots have been read by everybody: refill them
        if (tid == 0) {
            if (j + kRingTiles < n_mine)
                bulk_fetch(s_tile[sl_a], img_src + ((size_t)blockIdx.x + (j + kRingTiles) * g) * (kHmPix * kCH), kHmPix * kCH * 4,
                           &s_bar[sl_a]);
            if (j + 1 + kRingTiles < n_mine)
                bulk_fetch(s_tile[sl_b], img_src + ((size_t)blockIdx.x + (j + 1 + kRingTiles) * g) * (kHmPix * kCH),
                           kHmPix * kCH * 4, &s_bar[sl_b]);
        }
        float ya[kPerThread], yb[kPerThread];
#pragma unroll
        for (int k = 0; k < kPerThread; ++k) {
            exact_sigmoidf_pair(va[k], vb[k], ya[k], yb[k]);      // two tiles' values share the packed FFMA2 / FMUL2 stream
            mn = fminf(mn, fminf(ya[k], yb[k])); mx = fmaxf(mx, fmaxf(ya[k], yb[k]));
        }
        if (is_kp) {
#pragma unroll
            for (int k = 0; k < kPerThread; ++k) {
                __stcs(oa + k * (16 * kNK), ya[k]);               // outputs are not read again on the device: streaming stores
                if (two) __stcs(ob + k * (16 * kNK), yb[k]);
            }
        } else if (oa) {
#pragma unroll
            for (int k = 0; k < kPerThread; ++k) {
                __stcs(oa + k * 16, va[k]);
                if (two) __stcs(ob + k * 16, vb[k]);
            }
        }
        if (oa) { oa += 2 * g * ostep; ob += 2 * g * ostep; }
    }
    __syncthreads();
    if (is_kp) {
        atomicMin(&s_min[c], __float_as_int(mn));            // sigmoid output is >= 0: integer order == float order
        atomicMax(&s_max[c], __float_as_int(mx));
    }
    __syncthreads();
    publish_minmax(s_min, s_max, &s_last, partial, counter, minmax);
}

// SURVEY section 8(f) row 2 -- the tail of KeypointSubnet fused in front of the activation pass:
//   detector/keypoint_subnet.py:49-58   heatmaps = conv2d(x, 18, kernel_size=1) + bias, NCHW -> NHWC transpose
//   create_pb.py:73-76, 90, 92          sigmoid / split / per-(image, channel) min and max
// x is the 64-channel NCHW feature map after final_bn + ReLU.  One thread per pixel, 256 pixels per CTA tile: the tile of
// x is staged through shared memory 32 input channels at a time (coalesced rows of the channel planes), the [64][18]
// kernel sits in shared memory padded to 20 columns and is read as broadcast 16-byte vectors, and every thread keeps its
// pixel's 18 accumulators in registers: 18 fma per 6 shared-memory loads, ascending input channel.  The pixel's logits
// are optionally stored (72 contiguous bytes), then activated and written exactly like heatmap_kernel does; the logits
// tensor (72 B / pixel written and read back) never has to exist.
constexpr int kHeadCin = 64;
constexpr int kHeadPix = 256;          // pixels per tile = threads per CTA
constexpr int kHeadChunk = 32;         // input channels staged at a time
__global__ void __launch_bounds__(kHeadPix) heatmap_head_kernel(const float *__restrict__ x, const float *__restrict__ w,
                                                                const float *__restrict__ bias, const int npix,
                                                                const int tiles_per_img, float *__restrict__ logits,
                                                                float *__restrict__ kh, float *__restrict__ seg,
                                                                int *__restrict__ partial,
                                                                unsigned int *__restrict__ counter,
                                                                int *__restrict__ minmax)
{
    pdl_trigger();
    __shared__ __align__(16) float s_x[kHeadChunk][kHeadPix];
    __shared__ __align__(16) float s_w[kHeadCin][kPadCh];
    __shared__ float s_b[kCH];
    __shared__ int s_min[kNK], s_max[kNK];
    __shared__ int s_last;
    const int img = blockIdx.y, tid = threadIdx.x, lane = tid & 31;
    float mn[kNK], mx[kNK];
#pragma unroll
    for (int k = 0; k < kNK; ++k) { mn[k] = __int_as_float(0x7f800000); mx[k] = 0.0f; }
    if (tid < kNK) { s_min[tid] = 0x7f800000; s_max[tid] = 0; }
    for (int i = tid; i < kHeadCin * kPadCh; i += kHeadPix) {
        const int c = i / kPadCh, k = i - c * kPadCh;
        s_w[c][k] = k < kCH ? __ldg(w + c * kCH + k) : 0.0f;
    }
    if (tid < kCH) s_b[tid] = __ldg(bias + tid);

    const float *x_img = x + (size_t)img * kHeadCin * npix;
    for (int tile = blockIdx.x; tile < tiles_per_img; tile += gridDim.x) {
        const size_t pix = (size_t)tile * kHeadPix + tid;
        float acc[kPadCh];
        for (int half = 0; half < kHeadCin / kHeadChunk; ++half) {
            __syncthreads();                               // previous chunk consumed (and s_w / s_b ready)
            for (int i = tid; i < kHeadChunk * kHeadPix / 4; i += kHeadPix) {
                const int c = i / (kHeadPix / 4), q = i - c * (kHeadPix / 4);
                reinterpret_cast<float4 *>(&s_x[c][0])[q] = __ldcs(
                    reinterpret_cast<const float4 *>(x_img + (size_t)(half * kHeadChunk + c) * npix + (size_t)tile * kHeadPix) + q);
            }
            __syncthreads();
            if (half == 0) {
#pragma unroll
                for (int k = 0; k < kPadCh; ++k) acc[k] = k < kCH ? s_b[k] : 0.0f;
            }
#pragma unroll 4
            for (int c = 0; c < kHeadChunk; ++c) {
                const float xv = s_x[c][tid];
                const float4 *wr = reinterpret_cast<const float4 *>(&s_w[half * kHeadChunk + c][0]);
#pragma unroll
                for (int g = 0; g < kGroups; ++g) {
                    const float4 wv = wr[g];
                    acc[4 * g + 0] = __fmaf_rn(xv, wv.x, acc[4 * g + 0]);
                    acc[4 * g + 1] = __fmaf_rn(xv, wv.y, acc[4 * g + 1]);
                    acc[4 * g + 2] = __fmaf_rn(xv, wv.z, acc[4 * g + 2]);
                    acc[4 * g + 3] = __fmaf_rn(xv, wv.w, acc[4 * g + 3]);
                }
            }
        }
        if (logits) {
            float2 *lg = reinterpret_cast<float2 *>(logits + ((size_t)img * npix + pix) * kCH);
#pragma unroll
            for (int k = 0; k < kCH / 2; ++k) lg[k] = make_float2(acc[2 * k], acc[2 * k + 1]);
        }
        float *kp = kh + ((size_t)img * npix + pix) * kNK;
#pragma unroll
        for (int k = 0; k < kNK; ++k) {
            const float sg = exact_sigmoidf(acc[k]);
            mn[k] = fminf(mn[k], sg); mx[k] = fmaxf(mx[k], sg);
            kp[k] = sg;
        }
        if (seg) seg[(size_t)img * npix + pix] = acc[kNK];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < kNK; ++k) {
        float lo = mn[k], hi = mx[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) {
            atomicMin(&s_min[k], __float_as_int(lo));
            atomicMax(&s_max[k], __float_as_int(hi));
        }
    }
    __syncthreads();
    publish_minmax(s_min, s_max, &s_last, partial, counter, minmax);
}

// Division by a channel's constant range d = M - m (create_pb.py:93), one per heatmap value or crop tap -- the kernels
// that do it are bound by instruction issue.  With y = RN(1 / d) and q = RN(a y), the correction
// q' = fma(fma(-q, d, a), y, q) IS the correctly rounded a / d for 0 <= a <= d, d >= 2^-60 and a == 0 or a >= 1e-30
// (tools/div_check.cu: 10^10 pairs incl. the all-ones / all-zeros mantissas, 0 mismatches; subnormal a does differ).
// range_rcp() returns 0 for a range that must take the true division (tiny, NaN, or M == m, whose 0 / 0 = NaN the
// reference propagates).
__device__ __forceinline__ float range_rcp(float d) { return (d >= 8.67361738e-19f) ? __frcp_rn(d) : 0.0f; }   // 2^-60
__device__ __forceinline__ float div_by_range(float a, float d, float y)
{
    if (y == 0.0f || (a != 0.0f && a < 1e-30f)) return fdiv(a, d);
    const float q0 = fmul(a, y);
    return __fmaf_rn(__fmaf_rn(-q0, d, a), y, q0);
}

// create_pb.py:93-94 on one tap
__device__ __forceinline__ float normalise_tap(float v, float m, float d, float y, float mask)
{
    return fmul(div_by_range(fsub(v, m), d, y), mask);
}

// ---- the two-pass form used by mpn_run whenever the crops come from the padded normalised map -------------------------
// The sigmoid recipe of mpn_math.cuh is monotone non-decreasing over ALL finite floats (walked exhaustively, 2^32 values:
// tests/test_oracle_kat.py on the oracle's copy of the recipe, tests/test_gpu_parity.py on this one), so
//     max over pixels of sigmoid(l) == sigmoid(max over pixels of l)     bit for bit, and the same for min.
// Pass 1 (logit_minmax_kernel) therefore takes min / max of the raw LOGITS -- two instructions per value instead of the
// ~35 of the bit-exact sigmoid, a pure HBM stream that runs beside the candidate scan -- and its last CTA per image
// publishes (m, M) = (sigmoid(min), sigmoid(max)).  Pass 2 (heatmap_norm_kernel) then knows the range when it computes the
// activations and writes all three products in one go: keypoint_heatmaps (create_pb.py:74), segmentation_masks (:75) and
// the min-max normalised, masked map of create_pb.py:93-94 (padded to 20 floats per pixel for the crop kernel's 16-byte
// taps).  This replaces activation + min / max in one kernel followed by a normalisation kernel that read the 17-channel
// map back (13.9 MB at 640 x 640 x 8) and divided every value in a pass of its own.

// order-preserving float <-> unsigned (atomicMin / atomicMax on possibly negative logits)
__device__ __forceinline__ unsigned float_key(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_float(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }
constexpr unsigned kKeyPosInf = 0xff800000u;       // float_key(+inf)
constexpr unsigned kKeyNegInf = 0x007fffffu;       // float_key(-inf)

// Pass 1: per-CTA (min, max) of the logits of every channel, as ordered keys, to partial[img][chunk][17][2].  No fold
// across CTAs here (no fence, no counter, no last-CTA tail on the critical path): pass 2 folds the chunks of its image.
__global__ void __launch_bounds__(kHmThreads) logit_minmax_kernel(const float *__restrict__ hml, const int npix,
                                                                  const int tiles_per_img, unsigned *__restrict__ partial)
{
    __shared__ unsigned s_min[kNK], s_max[kNK];
    const int img = blockIdx.y, tid = threadIdx.x;
    pdl_trigger();
    const int c = tid % kCH;
    float mn = __int_as_float(0x7f800000), mx = -__int_as_float(0x7f800000);
    if (tid < kNK) { s_min[tid] = kKeyPosInf; s_max[tid] = kKeyNegInf; }
    const float *src = hml + (size_t)img * npix * kCH + tid;
    // four tiles per trip: sixteen independent loads in flight per thread (default caching: pass 2 re-reads the logits,
    // from L2 whenever the call's maps fit)
    for (int tile = blockIdx.x; tile < tiles_per_img; tile += 4 * gridDim.x) {
        float v[4][kPerThread];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = tile + u * gridDim.x;              // past the end: the trip's first tile again (min / max unchanged)
            const float *st = src + (size_t)(t < tiles_per_img ? t : tile) * (kHmPix * kCH);
#pragma unroll
            for (int k = 0; k < kPerThread; ++k) v[u][k] = __ldg(st + k * kHmThreads);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int k = 0; k < kPerThread; ++k) { mn = fminf(mn, v[u][k]); mx = fmaxf(mx, v[u][k]); }
    }
    __syncthreads();
    if (c < kNK) {
        atomicMin(&s_min[c], float_key(mn));
        atomicMax(&s_max[c], float_key(mx));
    }
    __syncthreads();
    unsigned *my = partial + ((size_t)img * gridDim.x + blockIdx.x) * kNK * 2;
    if (tid < kNK) { my[tid * 2] = s_min[tid]; my[tid * 2 + 1] = s_max[tid]; }
}

// Pass 2.  Every CTA first folds the `n_chunks` per-CTA extremes pass 1 left for its image (a few KB from L2) and turns
// them into its channel's constants with the monotone recipe: m = sigmoid(min logit), M = sigmoid(max logit)
// (create_pb.py:90,92), d = M - m, 1 / d, mask = float(M > 0.2) (:91).  The first trip's logits are requested before the
// wait on pass 1 (they do not depend on it).  CTA 0 of every image also publishes (m, M) to minmax[img].
__global__ void __launch_bounds__(kHmThreads, 4) heatmap_norm_kernel(const float *__restrict__ hml, const int npix,
                                                                     const int tiles_per_img, float *__restrict__ kh,
                                                                     float *__restrict__ seg,
                                                                     const unsigned *__restrict__ partial,
                                                                     const int n_chunks, float *__restrict__ minmax,
                                                                     float *__restrict__ nh)
{
    __shared__ unsigned s_min[kNK], s_max[kNK];
    __shared__ __align__(128) float s_tile[kRingTiles][kHmPix * kCH];
    __shared__ unsigned long long s_bar[kRingTiles];
    const int img = blockIdx.y, tid = threadIdx.x;
    pdl_trigger();
    const int c = tid % kCH, p0 = tid / kCH;
    const bool is_kp = c < kNK;
    if (tid < kNK) { s_min[tid] = kKeyPosInf; s_max[tid] = kKeyNegInf; }
    // tile ring and running output pointers as in heatmap_kernel; a keypoint thread also owns a slot of the padded
    // normalised map.  The first tiles are requested before the wait on pass 1 (they do not depend on it).
    const size_t g = gridDim.x;
    const float *img_src = hml + (size_t)img * npix * kCH;
    const int n_mine = (int)blockIdx.x < tiles_per_img ? (tiles_per_img - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    if (tid == 0) {
        bulk_barrier_init(s_bar, kRingTiles);
        for (int j = 0; j < kRingTiles && j < n_mine; ++j)
            bulk_fetch(s_tile[j], img_src + ((size_t)blockIdx.x + j * g) * (kHmPix * kCH), kHmPix * kCH * 4, &s_bar[j]);
    }
    float *oa = is_kp ? kh + ((size_t)img * npix + (size_t)blockIdx.x * kHmPix + p0) * kNK + c
                      : (seg ? seg + (size_t)img * npix + (size_t)blockIdx.x * kHmPix + p0 : nullptr);
    const size_t ostep = is_kp ? kHmPix * kNK : kHmPix;
    float *ob = oa ? oa + g * ostep : nullptr;
    float *na = nh + ((size_t)img * npix + (size_t)blockIdx.x * kHmPix + p0) * kPadCh + c, *nb = na + g * (kHmPix * kPadCh);
    __syncthreads();
    pdl_wait();                                        // pass 1 has completed
    if (is_kp) {
        unsigned lo = kKeyPosInf, hi = kKeyNegInf;
        const unsigned *q = partial + (size_t)img * n_chunks * (kNK * 2) + c * 2;
        for (int i = p0; i < n_chunks; i += kHmThreads / kCH) {
            const uint2 e = __ldcg(reinterpret_cast<const uint2 *>(q + (size_t)i * (kNK * 2)));
            lo = min(lo, e.x); hi = max(hi, e.y);
        }
        atomicMin(&s_min[c], lo); atomicMax(&s_max[c], hi);
    }
    __syncthreads();
    float m = 0.0f, d = 1.0f, rcp = 1.0f, mask = 0.0f;
    if (is_kp) {
        const float hi = exact_sigmoidf(key_float(s_max[c]));
        m = exact_sigmoidf(key_float(s_min[c]));
        d = fsub(hi, m);
        rcp = range_rcp(d);
        mask = hi > 0.2f ? 1.0f : 0.0f;
        if (blockIdx.x == 0 && p0 == 0) {
            minmax[((size_t)img * kNK + c) * 2] = m;
            minmax[((size_t)img * kNK + c) * 2 + 1] = hi;
        }
    }
    // The reciprocal shortcut of div_by_range holds for a == 0 or a >= 1e-30 (a = kh - m); with m >= 1e-22 the smallest
    // non-zero difference of two floats >= m is far above that, so the per-value test collapses to this per-thread flag
    // (false for ordinary channels: the loop body is then straight-line code).
    const bool careful = rcp == 0.0f || m < 1e-22f;
    for (int j = 0; j < n_mine; j += 2) {
        const bool two = j + 1 < n_mine;
        const int sl_a = j % kRingTiles, sl_b = (j + 1) % kRingTiles;
        bulk_wait(&s_bar[sl_a], (unsigned)(j / kRingTiles) & 1u);
        if (two) bulk_wait(&s_bar[sl_b], (unsigned)((j + 1) / kRingTiles) & 1u);
        float xa[kPerThread], xb[kPerThread];
#pragma unroll
        for (int k = 0; k < kPerThread; ++k) {
            xa[k] = s_tile[sl_a][k * kHmThreads + tid];
            xb[k] = two ? s_tile[sl_b][k * kHmThreads + tid] : xa[k];
        }
        __syncthreads();                               // both slots have been read by everybody: refill them
        if (tid == 0) {
            if (j + kRingTiles < n_mine)
                bulk_fetch(s_tile[sl_a], img_src + ((size_t)blockIdx.x + (j + kRingTiles) * g) * (kHmPix * kCH), kHmPix * kCH * 4,
                           &s_bar[sl_a]);
            if (j + 1 + kRingTiles < n_mine)
                bulk_fetch(s_tile[sl_b], img_src + ((size_t)blockIdx.x + (j + 1 + kRingTiles) * g) * (kHmPix * kCH),
                           kHmPix * kCH * 4, &s_bar[sl_b]);
        }
        float ya[kPerThread], yb[kPerThread];
#pragma unroll
        for (int k = 0; k < kPerThread; ++k) exact_sigmoidf_pair(xa[k], xb[k], ya[k], yb[k]);
        if (is_kp) {
#pragma unroll
            for (int k = 0; k < kPerThread; ++k) {
                float qa, qb;
                if (careful) {
                    qa = normalise_tap(ya[k], m, d, rcp, mask);
                    qb = normalise_tap(yb[k], m, d, rcp, mask);
                } else {                                     // div_by_range's shortcut, unconditionally
                    const float aa = fsub(ya[k], m), ab = fsub(yb[k], m);
                    const float q0a = fmul(aa, rcp), q0b = fmul(ab, rcp);
                    qa = fmul(__fmaf_rn(__fmaf_rn(-q0a, d, aa), rcp, q0a), mask);
                    qb = fmul(__fmaf_rn(__fmaf_rn(-q0b, d, ab), rcp, q0b), mask);
                }
                __stcs(oa + k * (16 * kNK), ya[k]);          // an output: not read again on the device
                na[k * (16 * kPadCh)] = qa;                  // read by the crop kernel next: default policy
                if (two) {
                    __stcs(ob + k * (16 * kNK), yb[k]);
                    nb[k * (16 * kPadCh)] = qb;
                }
            }
        } else if (oa) {
#pragma unroll
            for (int k = 0; k < kPerThread; ++k) {
                __stcs(oa + k * 16, xa[k]);
                if (two) __stcs(ob + k * 16, xb[k]);
            }
        }
        if (oa) { oa += 2 * g * ostep; ob += 2 * g * ostep; }
        na += 2 * g * (kHmPix * kPadCh); nb += 2 * g * (kHmPix * kPadCh);
    }
}

// ---- the person of a crop CTA (common.cuh: PersonList) -------------------------------------------------------------------
// Returns false (uniformly over the CTA) when the CTA has no person; otherwise the box, its image b and the output row.
// Explicit list: CTA n crops list entry n.  Derived mode: CTA n stands for SLOT (b, k) = (n / max_det, n % max_det) of the
// padded detection outputs; warp 0 scans num_boxes -- row = (boxes of the images before b) + k, valid when k <
// num_boxes[b], i.e. the flat order of create_pb.py:96-103 -- while every thread already has the slot's box in flight: one
// L2 round trip in all.  Contains one __syncthreads() in derived mode.
__device__ __forceinline__ bool find_person(const PersonList &pl, int n, int *s_p, float4 &box, int &b, int &row)
{
    if (pl.num_boxes == nullptr) {                     // explicit list: count, box and image index in ONE round trip
        box = __ldcg(reinterpret_cast<const float4 *>(pl.boxes) + n);
        const int b_raw = __ldcg(pl.box_ind + n);
        const int N = pl.n_dev ? __ldcg(pl.n_dev) : pl.n_host;
        b = b_raw;
        row = n;
        return n < N;
    }
    box = __ldcg(reinterpret_cast<const float4 *>(pl.det_boxes) + n);      // zero padding when the slot is empty
    b = n / pl.max_det;
    const int k = n - b * pl.max_det;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        int before = 0, mine = 0;
        for (int b0 = 0; b0 <= b; b0 += 32) {
            const int nb = (b0 + lane <= b) ? __ldcg(pl.num_boxes + b0 + lane) : 0;
            if (b0 + lane == b) mine = nb;
            int v = (b0 + lane < b) ? nb : 0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            before += v;
        }
        mine = __shfl_sync(0xffffffffu, mine, b & 31);
        if (lane == 0) { s_p[0] = before + k; s_p[1] = k < mine; }
    }
    __syncthreads();
    row = s_p[0];
    return s_p[1] != 0;
}

// The list CTA of a derived-mode crop grid: person_offsets = exclusive scan of num_boxes (+ the user's copy), and for every
// person row its box and image (create_pb.py:96-103).  `scratch`: >= min(B, 1024) ints of shared memory.
__device__ __forceinline__ void build_person_list(const PersonList &pl, int *scratch)
{
    const int tid = threadIdx.x, lane = tid & 31;
    if (tid < 32) {
        int running = 0;
        for (int b0 = 0; b0 < pl.B; b0 += 32) {
            const int b = b0 + lane;
            const int nb = (b < pl.B) ? __ldcg(pl.num_boxes + b) : 0;
            int incl = nb;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            if (b < pl.B) {
                const int off = running + incl - nb;
                if (b < 1024) scratch[b] = off;
                pl.person_offsets[b] = off;
                if (pl.person_offsets_out) pl.person_offsets_out[b] = off;
            }
            running += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) {
            pl.person_offsets[pl.B] = running;
            if (pl.person_offsets_out) pl.person_offsets_out[pl.B] = running;
        }
    }
    __syncthreads();
    if (!pl.person_box) return;
    for (int idx = tid; idx < pl.B * pl.max_det; idx += blockDim.x) {
        const int b = idx / pl.max_det, k = idx - b * pl.max_det;
        if (k < __ldcg(pl.num_boxes + b)) {
            const int row = (b < 1024 ? scratch[b] : pl.person_offsets[b]) + k;
            reinterpret_cast<float4 *>(pl.person_box)[row] = __ldcg(reinterpret_cast<const float4 *>(pl.det_boxes) + idx);
            pl.person_img[row] = b;
        }
    }
}

// crop_and_resize of the PADDED normalised map (the path of mpn_run wherever that map exists): one person per blockIdx.x,
// a band of ROWS crop rows per blockIdx.y.  The sampling geometry is separable -- in_y depends on the crop row only, in_x
// on the crop column only (the op's own fp32 operation order) -- and the work is laid out along that split: a thread owns
// one (crop column, group of 4 channels) for the whole band, keeps its column's two source offsets and lerp weight in
// registers, and walks the band's rows; per row it reads the row's entry (one broadcast 16-byte shared-memory load), issues
// four aligned 16-byte tap loads (base pointer + row offset), does the 12 lerps and drops the four results into a
// shared-memory image of the band, which is then written out in memory order as 16-byte (fp32) and 8-byte (bf16) vectors.
// No division, no per-item index arithmetic: crop size, band height and block size are compile-time constants (the
// runtime-sized one-item-per-thread version spent two thirds of its issue slots on index arithmetic).
struct __align__(16) AxisTab {
    unsigned lo, hi;                  // offsets of the two source rows (row * ww * 5), in 16-byte units
    float w;                          // lerp weight towards `hi`
    unsigned valid;                   // 0: outside the source map (extrapolation value 0)
};

template <int CH, int CW, int ROWS>
__global__ void __launch_bounds__((CW * kGroups + 31) / 32 * 32)
crop_padded_kernel(const float *__restrict__ src, const int hh, const int ww, const PersonList pl,
                   float *__restrict__ out_f32, __nv_bfloat16 *__restrict__ out_bf16)
{
    constexpr int kThreads = (CW * kGroups + 31) / 32 * 32;          // 192 for 36 columns: 180 of them own a (column, group)
    constexpr int kPix = ROWS * CW, kOut4 = kPix * kNK / 4;
    static_assert(CH % ROWS == 0 && ROWS % 4 == 0 && (kPix * kNK) % 4 == 0 && (CH * CW * kNK) % 4 == 0 && ROWS <= 32, "band geometry");
    __shared__ AxisTab s_y[ROWS];
    __shared__ __align__(16) float s_out[kPix * kNK];
    __shared__ int s_p[4];
    const int n = blockIdx.x, tid = threadIdx.x;
    pdl_trigger();
    pdl_wait();                                        // normalised map and detections are complete
    if (blockIdx.y == CH / ROWS) {                     // the extra row of the grid: one CTA writes the flat person list
        if (n == 0 && pl.num_boxes) build_person_list(pl, reinterpret_cast<int *>(s_out));
        return;
    }
    float4 box;
    int b, row;
    if (!find_person(pl, n, s_p, box, b, row)) return;
    const int cy0 = blockIdx.y * ROWS;
    const float hm1 = (float)(hh - 1), wm1 = (float)(ww - 1);
    if (tid < ROWS) {                                  // the band's rows (create_pb.py:106-109, crop_and_resize_op.cc)
        const float y1 = box.x, y2 = box.z;
        float in_y;
        if (CH > 1) {
            const float hs = fdiv(fmul(fsub(y2, y1), hm1), (float)(CH - 1));
            in_y = fadd(fmul(y1, hm1), fmul((float)(cy0 + tid), hs));
        } else {
            in_y = fmul(fmul(0.5f, fadd(y1, y2)), hm1);
        }
        const int top = (int)floorf(in_y), bot = (int)ceilf(in_y);
        AxisTab t;
        t.valid = !(in_y < 0.0f || in_y > hm1) ? 1u : 0u;
        t.lo = t.valid ? (unsigned)(top * ww * kGroups) : 0u; t.hi = t.valid ? (unsigned)(bot * ww * kGroups) : 0u;
        t.w = fsub(in_y, (float)top);
        s_y[tid] = t;
    }
    // this thread's column
    const int cx = min(tid / kGroups, CW - 1), g = tid - (tid / kGroups) * kGroups;
    const bool owner = tid < CW * kGroups;
    float in_x;
    {
        const float x1 = box.y, x2 = box.w;
        if (CW > 1) {
            const float ws = fdiv(fmul(fsub(x2, x1), wm1), (float)(CW - 1));
            in_x = fadd(fmul(x1, wm1), fmul((float)cx, ws));
        } else {
            in_x = fmul(fmul(0.5f, fadd(x1, x2)), wm1);
        }
    }
    const bool x_valid = !(in_x < 0.0f || in_x > wm1);
    const int left = (int)floorf(in_x), right = (int)ceilf(in_x);
    const float lx = fsub(in_x, (float)left);
    const float4 *img = reinterpret_cast<const float4 *>(src + (size_t)b * hh * ww * kPadCh);
    unsigned xl = x_valid ? (unsigned)(left * kGroups + g) : 0u, xr = x_valid ? (unsigned)(right * kGroups + g) : 0u;
    asm volatile("" : "+r"(xl), "+r"(xr));             // keep the two column offsets in registers (no rematerialisation)
    __syncthreads();
    if (owner) {
        float *so = s_out + cx * kNK + 4 * g;
        // Four crop rows at a time: ALL sixteen tap loads are issued before the first lerp (rows or columns outside the map
        // carry offset 0: the load is harmless and its result unused).  What the kernel needs is loads in flight: a variant
        // that reused horizontally interpolated source rows across crop rows (a third of the loads, same bits) was 20 %
        // SLOWER at 2 805 persons, because its loads sat behind row-dependent branches, two at a time.
#pragma unroll
        for (int r0 = 0; r0 < ROWS; r0 += 4) {
            uint4 ty[4];
            float4 tap[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                ty[i] = *reinterpret_cast<const uint4 *>(&s_y[r0 + i]);
                tap[i][0] = __ldg(img + (ty[i].x + xl)); tap[i][1] = __ldg(img + (ty[i].x + xr));
                tap[i][2] = __ldg(img + (ty[i].y + xl)); tap[i][3] = __ldg(img + (ty[i].y + xr));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float ly = __uint_as_float(ty[i].z);
                const float tl[4] = {tap[i][0].x, tap[i][0].y, tap[i][0].z, tap[i][0].w};
                const float tr[4] = {tap[i][1].x, tap[i][1].y, tap[i][1].z, tap[i][1].w};
                const float bl[4] = {tap[i][2].x, tap[i][2].y, tap[i][2].z, tap[i][2].w};
                const float br[4] = {tap[i][3].x, tap[i][3].y, tap[i][3].z, tap[i][3].w};
                const bool inside = ty[i].w != 0u && x_valid;
                float o[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float tpv = fadd(tl[k], fmul(fsub(tr[k], tl[k]), lx));
                    const float btv = fadd(bl[k], fmul(fsub(br[k], bl[k]), lx));
                    o[k] = inside ? fadd(tpv, fmul(fsub(btv, tpv), ly)) : 0.0f;
                }
                float *sr = so + (r0 + i) * (CW * kNK);
                sr[0] = o[0];
                if (g < kGroups - 1) { sr[1] = o[1]; sr[2] = o[2]; sr[3] = o[3]; }      // group 4 holds channel 16 only
            }
        }
    }
    __syncthreads();
    const size_t o0 = ((size_t)row * CH + cy0) * (CW * kNK);         // multiple of 4 floats
#pragma unroll
    for (int it = 0; it < (kOut4 + kThreads - 1) / kThreads; ++it) {
        const int f = it * kThreads + tid;
        if (f < kOut4) {
            const float4 v = reinterpret_cast<const float4 *>(s_out)[f];
            // Streaming (evict-first) stores: in a crowded call the crops are hundreds of MB written once and read by the
            // PRN much later, while the normalised map the taps come from (tens of MB, re-read by every person of an
            // image) should be what stays in L2; in a small call nothing is evicted either way.
            if (out_f32) __stcs(reinterpret_cast<float4 *>(out_f32 + o0) + f, v);
            if (out_bf16) {
                const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
                uint2 u;
                u.x = *reinterpret_cast<const unsigned *>(&lo);
                u.y = *reinterpret_cast<const unsigned *>(&hi);
                __stcs(reinterpret_cast<uint2 *>(out_bf16 + o0) + f, u);
            }
        }
    }
}

// tf.image.crop_and_resize (create_pb.py:106-109; TF 1.15 crop_and_resize_op.cc: bilinear, extrapolation 0) of one
// person per blockIdx.x, a band of crop rows per blockIdx.y.  The sampling geometry of every crop pixel (tap offsets and
// lerp weights, in the op's own fp32 operation order) is computed once into shared memory; the main loop then walks the
// output in memory order, four consecutive samples j = (cy*crop_w + cx)*17 + c per thread (j is also the PRN input
// column, detector/prn.py:17), so that the fp32 and bf16 rows are written as 16- and 8-byte vectors.
struct PixTab {
    int top, bot;        // element offsets of the two source rows (row * ww * 17)
    int left, right;     // element offsets of the two source columns (col * 17)
    float ly, lx;
    int valid;
};
constexpr int kCropBands = 14;   // 4 crop rows per CTA: enough CTAs in flight to hide the tap-gather latency
constexpr int kCropMaxPix = 1024;     // pixels of one band held in shared memory

__global__ void __launch_bounds__(256) crop_kernel(const float *__restrict__ src, const float *__restrict__ minmax,
                                                   const int hh, const int ww, const PersonList pl, const int crop_h,
                                                   const int crop_w, float *__restrict__ out_f32,
                                                   __nv_bfloat16 *__restrict__ out_bf16)
{
    __shared__ PixTab s_tab[kCropMaxPix];
    __shared__ float s_m[kNK], s_d[kNK], s_rcp[kNK], s_mask[kNK];
    __shared__ int s_p[4];
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.x;
    const int n_bands = gridDim.y - 1;                 // the last row of the grid: one CTA writes the flat person list
    if ((int)blockIdx.y == n_bands) {
        if (n == 0 && pl.num_boxes) build_person_list(pl, reinterpret_cast<int *>(s_tab));
        return;
    }
    float4 box;
    int b, row;
    if (!find_person(pl, n, s_p, box, b, row)) return;
    const int rows_per_band = (crop_h + n_bands - 1) / n_bands;
    const int cy0 = blockIdx.y * rows_per_band, cy1 = min(crop_h, cy0 + rows_per_band);
    if (cy0 >= cy1) return;
    const int npix = (cy1 - cy0) * crop_w;
    const float y1 = box.x, x1 = box.y, y2 = box.z, x2 = box.w;
    const float hm1 = (float)(hh - 1), wm1 = (float)(ww - 1);
    if (minmax != nullptr && threadIdx.x < kNK) {
        const float m = __ldg(minmax + ((size_t)b * kNK + threadIdx.x) * 2);
        const float M = __ldg(minmax + ((size_t)b * kNK + threadIdx.x) * 2 + 1);
        const float d = fsub(M, m);
        s_m[threadIdx.x] = m; s_d[threadIdx.x] = d; s_rcp[threadIdx.x] = range_rcp(d);
        s_mask[threadIdx.x] = (M > 0.2f) ? 1.0f : 0.0f;
    }
    for (int p = threadIdx.x; p < npix; p += blockDim.x) {
        const int cy = cy0 + p / crop_w, cx = p % crop_w;
        float in_y, in_x;
        if (crop_h > 1) {
            const float hs = fdiv(fmul(fsub(y2, y1), hm1), (float)(crop_h - 1));
            in_y = fadd(fmul(y1, hm1), fmul((float)cy, hs));
        } else {
            in_y = fmul(fmul(0.5f, fadd(y1, y2)), hm1);
        }
        if (crop_w > 1) {
            const float ws = fdiv(fmul(fsub(x2, x1), wm1), (float)(crop_w - 1));
            in_x = fadd(fmul(x1, wm1), fmul((float)cx, ws));
        } else {
            in_x = fmul(fmul(0.5f, fadd(x1, x2)), wm1);
        }
        PixTab t;
        t.valid = !(in_y < 0.0f || in_y > hm1 || in_x < 0.0f || in_x > wm1);
        const int top = (int)floorf(in_y), bot = (int)ceilf(in_y);
        const int left = (int)floorf(in_x), right = (int)ceilf(in_x);
        t.top = t.valid ? top * ww * kNK : 0; t.bot = t.valid ? bot * ww * kNK : 0;
        t.left = t.valid ? left * kNK : 0; t.right = t.valid ? right * kNK : 0;
        t.ly = fsub(in_y, (float)top); t.lx = fsub(in_x, (float)left);
        s_tab[p] = t;
    }
    __syncthreads();
    const float *img = src + (size_t)b * hh * ww * kNK;
    const int D = crop_h * crop_w * kNK;
    const int j_begin = cy0 * crop_w * kNK, n4 = npix * kNK / 4;     // crop_w * 17 * rows is a multiple of 4 for 36 columns
    const bool norm = minmax != nullptr;
    for (int f = threadIdx.x; f < n4; f += blockDim.x) {
        const int jl = 4 * f;
        int p = jl / kNK, c = jl - p * kNK;
        float r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const PixTab t = s_tab[p];
            float v = 0.0f;
            if (t.valid) {
                const float *q = img + c;
                float tl = __ldg(q + t.top + t.left), tr = __ldg(q + t.top + t.right);
                float bl = __ldg(q + t.bot + t.left), br = __ldg(q + t.bot + t.right);
                if (norm) {
                    const float m = s_m[c], d = s_d[c], y = s_rcp[c], mask = s_mask[c];
                    tl = normalise_tap(tl, m, d, y, mask); tr = normalise_tap(tr, m, d, y, mask);
                    bl = normalise_tap(bl, m, d, y, mask); br = normalise_tap(br, m, d, y, mask);
                }
                const float tp = fadd(tl, fmul(fsub(tr, tl), t.lx));
                const float bt = fadd(bl, fmul(fsub(br, bl), t.lx));
                v = fadd(tp, fmul(fsub(bt, tp), t.ly));
            }
            r[k] = v;
            if (++c == kNK) { c = 0; ++p; }
        }
        const size_t o = (size_t)row * D + j_begin + jl;
        if (out_f32) *reinterpret_cast<float4 *>(out_f32 + o) = make_float4(r[0], r[1], r[2], r[3]);
        if (out_bf16) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(r[0], r[1]), hi = __floats2bfloat162_rn(r[2], r[3]);
            uint2 u;
            u.x = *reinterpret_cast<const unsigned *>(&lo);
            u.y = *reinterpret_cast<const unsigned *>(&hi);
            *reinterpret_cast<uint2 *>(out_bf16 + o) = u;
        }
    }
}

// inference/utils.py:29-52.  One CTA of 17 x 32 threads: thread (c, q) scans positions q, q+32, ...
__global__ void __launch_bounds__(kNK * 32) get_keypoints_kernel(const float *__restrict__ hm, const int hh,
                                                                const int ww, const double ymin, const double xmin,
                                                                const double ymax, const double xmax,
                                                                const double threshold, int *__restrict__ out)
{
    __shared__ float s_val[kNK][32];
    __shared__ int s_idx[kNK][32];
    const int tid = threadIdx.x, c = tid % kNK, q = tid / kNK;
    const int P = hh * ww;
    float best = -__int_as_float(0x7f800000);
    int bidx = 0x7fffffff;
    for (int p = q; p < P; p += 32) {
        const float v = __ldg(hm + (size_t)p * kNK + c);
        if (v > best) { best = v; bidx = p; }
    }
    s_val[c][q] = best; s_idx[c][q] = bidx;
    __syncthreads();
    if (tid < kNK) {
        float bv = s_val[tid][0]; int bi = s_idx[tid][0];
        for (int i = 1; i < 32; ++i) {
            const float v = s_val[tid][i]; const int id = s_idx[tid][i];
            if (v > bv || (v == bv && id < bi)) { bv = v; bi = id; }
        }
        int x = 0, y = 0, vis = 0;
        if ((double)bv > threshold) {
            const double height = ymax - ymin, width = xmax - xmin;
            const int yy = bi / ww, xx = bi - yy * ww;
            double fy = trunc((double)yy * height / (double)hh);
            double fx = trunc((double)xx * width / (double)ww);
            fy = fy < 0.0 ? 0.0 : (fy > height ? height : fy);
            fx = fx < 0.0 ? 0.0 : (fx > width ? width : fx);
            x = (int)fx; y = (int)fy; vis = 1;
        }
        out[tid * 3 + 0] = x; out[tid * 3 + 1] = y; out[tid * 3 + 2] = vis;
    }
}

__global__ void minmax_copy_kernel(const float *ws, float *out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ws[i];
}

}  // namespace

// Resident CTA slots (CTAs per SM x SMs) of the three 288-thread streaming kernels on this handle's device.
int heatmap_prepare(HeatmapWaves *w)
{
    int sms = 0, dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) return -1;
    int a = 0, b = 0, c = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, heatmap_kernel, kHmThreads, 0) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, logit_minmax_kernel, kHmThreads, 0) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&c, heatmap_norm_kernel, kHmThreads, 0) != cudaSuccess || a < 1 || b < 1 ||
        c < 1)
        return -1;
    w->one_pass = a * sms; w->minmax = b * sms; w->norm = c * sms;
    return 0;
}

// Chunks per image of a heatmap-sized streaming grid: at most ONE resident wave of CTAs (`slots`) -- a grid of 1.35 or 1.5
// waves costs two full rounds of trips (measured: 800 CTAs x 2 trips at 640 x 640 x 8 took as long as 4 trips) -- minus room
// for the B single-CTA sort / NMS blocks that run beside it (an SM that holds one of them still takes two of these CTAs);
// `tiles_per_trip` 64-pixel tiles per CTA and trip.  Small calls end up with one trip per CTA.
static int wave_chunks_per_image(int slots, int B, int tiles, int tiles_per_trip)
{
    int cap = (slots - 2 * (B < 148 ? B : 148)) / B;
    if (cap < 1) cap = 1;
    const int trips = (tiles + tiles_per_trip * cap - 1) / (tiles_per_trip * cap);
    const int per_img = (tiles + tiles_per_trip * trips - 1) / (tiles_per_trip * trips);
    return per_img < 1 ? 1 : per_img;
}

int launch_heatmaps(const float *hml, int B, int hh, int ww, float *kh, float *seg, float *minmax_ws,
                    float *minmax_out, int *partial_ws, unsigned int *counter_ws, int slots, cudaStream_t s)
{
    const int npix = hh * ww;
    const int tiles = (npix + kHmPix - 1) / kHmPix;
    int launches = 0;
    dim3 grid(wave_chunks_per_image(slots, B, tiles, 2), B);
    prof_mark(s, "heatmap");
    heatmap_kernel<<<grid, kHmThreads, 0, s>>>(hml, npix, tiles, kh, seg, partial_ws, counter_ws,
                                               reinterpret_cast<int *>(minmax_ws));
    ++launches;
    if (minmax_out) {
        const int nmm = B * kNK * 2;
        minmax_copy_kernel<<<(nmm + 255) / 256, 256, 0, s>>>(minmax_ws, minmax_out, nmm);
        ++launches;
    }
    return launches;
}

int launch_heatmap_head(const float *x, const float *w, const float *bias, int B, int hh, int ww, float *logits, float *kh,
                        float *seg, float *minmax_ws, float *minmax_out, int *partial_ws, unsigned int *counter_ws,
                        cudaStream_t s)
{
    const int npix = hh * ww;
    if (npix % kHeadPix != 0) return -(int)cudaErrorInvalidValue;     // true for every image that is a multiple of 128
    const int tiles = npix / kHeadPix;
    int launches = 0;
    // the partial array holds ceil(npix / 128) chunks per image: tiles <= that
    int per_img = (148 * 6 + B - 1) / B;
    if (per_img > tiles) per_img = tiles;
    dim3 grid(per_img, B);
    prof_mark(s, "heatmap_head");
    heatmap_head_kernel<<<grid, kHeadPix, 0, s>>>(x, w, bias, npix, tiles, logits, kh, seg, partial_ws, counter_ws,
                                                  reinterpret_cast<int *>(minmax_ws));
    ++launches;
    if (minmax_out) {
        const int nmm = B * kNK * 2;
        minmax_copy_kernel<<<(nmm + 255) / 256, 256, 0, s>>>(minmax_ws, minmax_out, nmm);
        ++launches;
    }
    return launches;
}

// pass 1 of the two-pass form: per-CTA (min, max) of the logits -> partial_ws [B, *n_chunks, 17, 2] (ordered keys)
int launch_logit_minmax(const float *hml, int B, int hh, int ww, int *partial_ws, int partial_chunks_cap, int slots,
                        int *n_chunks, cudaStream_t s)
{
    const int npix = hh * ww, tiles = (npix + kHmPix - 1) / kHmPix;
    int per_img = wave_chunks_per_image(slots, B, tiles, 4);
    if (per_img > partial_chunks_cap) per_img = partial_chunks_cap;
    *n_chunks = per_img;
    prof_mark(s, "logit_minmax");
    logit_minmax_kernel<<<dim3(per_img, B), kHmThreads, 0, s>>>(hml, npix, tiles, reinterpret_cast<unsigned *>(partial_ws));
    return 1;
}

// pass 2: keypoint_heatmaps, segmentation_masks, the padded normalised map and minmax_ws [B, 17, 2] in one pass over the
// logits (n_chunks: what pass 1 reported)
int launch_heatmap_norm(const float *hml, int B, int hh, int ww, float *kh, float *seg, const int *partial_ws, int n_chunks,
                        float *minmax_ws, float *nh, float *minmax_out, int slots, cudaStream_t s)
{
    const int npix = hh * ww, tiles = (npix + kHmPix - 1) / kHmPix;
    const int per_img = wave_chunks_per_image(slots, B, tiles, 2);
    int launches = 1;
    prof_mark(s, "heatmap_norm");
    launch_k(heatmap_norm_kernel, dim3(per_img, B), dim3(kHmThreads), 0, s, true, hml, npix, tiles, kh, seg,
             reinterpret_cast<const unsigned *>(partial_ws), n_chunks, minmax_ws, nh);
    if (minmax_out) {
        const int nmm = B * kNK * 2;
        minmax_copy_kernel<<<(nmm + 255) / 256, 256, 0, s>>>(minmax_ws, minmax_out, nmm);
        ++launches;
    }
    return launches;
}

int launch_crop(const float *src, const float *minmax, int hh, int ww, const PersonList &pl, int n_max, int crop_h, int crop_w,
                float *crops_f32, __nv_bfloat16 *crops_bf16, cudaStream_t s)
{
    if (n_max <= 0) return 0;
    // (A version of this kernel compiled for the 56 x 36 crop, one (column, channel) per thread over 8 rows like
    // crop_padded_kernel, was measured at 1024 x 1024 x 64 / 594 persons: 0.130 ms against 0.115 ms for this one.)
    int bands = kCropBands;
    while (((crop_h + bands - 1) / bands) * crop_w > kCropMaxPix) ++bands;
    // every band must start at a multiple of 4 samples: rows_per_band * crop_w * 17 % 4 == 0
    while (bands > 1 && ((((crop_h + bands - 1) / bands) * crop_w * kNK) % 4 != 0)) --bands;
    if (((crop_h + bands - 1) / bands) * crop_w > kCropMaxPix || (crop_h * crop_w * kNK) % 4 != 0)
        return -(int)cudaErrorInvalidValue;
    dim3 grid(n_max, bands + 1);
    prof_mark(s, "crop");
    launch_k(crop_kernel, grid, dim3(256), 0, s, true, src, minmax, hh, ww, pl, crop_h, crop_w, crops_f32, crops_bf16);
    return 1;
}

// the padded crop kernel is compiled for the reference's crop size (create_pb.py:19); other sizes take crop_kernel
bool crop_padded_supported(int crop_h, int crop_w) { return crop_h == 56 && crop_w == 36; }

int launch_crop_padded(const float *nh, int hh, int ww, const PersonList &pl, int n_max, int crop_h, int crop_w,
                       float *crops_f32, __nv_bfloat16 *crops_bf16, cudaStream_t s)
{
    if (n_max <= 0) return 0;
    if (!crop_padded_supported(crop_h, crop_w)) return -(int)cudaErrorInvalidValue;
    // Rows per CTA: 4 (twice the CTAs) while the call is small enough to be a latency problem -- at most one resident wave of
    // CTAs -- 8 (half the per-CTA prologue per value) beyond.  (+ 1 grid row: the list CTA, see PersonList.)
    prof_mark(s, "crop");
    const dim3 block((36 * kGroups + 31) / 32 * 32);
    if (n_max <= 640)
        launch_k(crop_padded_kernel<56, 36, 4>, dim3(n_max, 14 + 1), block, 0, s, true, nh, hh, ww, pl, crops_f32, crops_bf16);
    else
        launch_k(crop_padded_kernel<56, 36, 8>, dim3(n_max, 7 + 1), block, 0, s, true, nh, hh, ww, pl, crops_f32, crops_bf16);
    return 1;
}

int launch_get_keypoints(const float *hm, int hh, int ww, double ymin, double xmin, double ymax, double xmax,
                         double threshold, int *out, cudaStream_t s)
{
    get_keypoints_kernel<<<1, kNK * 32, 0, s>>>(hm, hh, ww, ymin, xmin, ymax, xmax, threshold, out);
    return 1;
}

}  // namespace mpn
