// Heatmap activation + per-(image, channel) min/max, crop_and_resize with the min-max normalisation folded into the
// taps, and the stand-alone numpy decoder get_keypoints.
//
// Replaces (reference):
//   create_pb.py:73-76    keypoint_heatmaps = sigmoid(heatmaps[..., :17]); segmentation_masks = heatmaps[..., 17]
//   create_pb.py:90-94    M = reduce_max, m = reduce_min over (h, w); hm = (hm - m) / (M - m) * float(M > 0.2)
//   create_pb.py:106-109  tf.image.crop_and_resize(heatmaps, boxes, box_ind, crop_size=[56, 36])  (bilinear, extrapolation 0)
//   inference/utils.py:29-52  get_keypoints
//
// heatmap_kernel is a pure HBM stream: 72 B in, 72 B out per pixel.  The [.., 18] input is read as a flat float4
// stream; a CTA of 288 threads covers 64 pixels per step, so thread t always sees the same four channels
// ((4t + j) mod 18) and keeps its running min / max in registers.  The 17-channel and 1-channel outputs are
// re-packed through shared memory so that both are written as aligned float4 as well.
// The normalised heatmap is never written: crop_kernel normalises each bilinear tap on the fly (same arithmetic,
// same order, as normalising the whole map first).
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

// fold_minmax: every thread's four running (min, max) -> the CTA's, then publish.
__device__ __forceinline__ void fold_minmax(const int (&ch)[4], const float (&mn)[4], const float (&mx)[4], int *s_min,
                                            int *s_max, int *s_last, int *__restrict__ partial,
                                            unsigned int *__restrict__ counter, int *__restrict__ minmax)
{
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (ch[j] < kNK) {
            atomicMin(&s_min[ch[j]], __float_as_int(mn[j]));     // sigmoid output is >= 0: integer order == float order
            atomicMax(&s_max[ch[j]], __float_as_int(mx[j]));
        }
    }
    __syncthreads();
    publish_minmax(s_min, s_max, s_last, partial, counter, minmax);
}

// grid = (chunks per image, B).  Thread t of a CTA always sees the four channels (4 t + j) mod 18 of its 64-pixel tiles,
// so its running min / max live in registers.  Outputs are written straight from registers: within a warp the 17-channel
// rows are one contiguous run of addresses, so the scalar stores coalesce into full lines without a shared-memory
// repack or any barrier in the streaming loop.  Per-CTA (min, max) go to a small partial array; the last CTA of each
// image (threadfence + counter) folds them into minmax[b] and re-arms the counter, so no reset kernel is needed.
__global__ void __launch_bounds__(kHmThreads) heatmap_kernel(const float *__restrict__ hml, const int npix,
                                                             const int tiles_per_img, float *__restrict__ kh,
                                                             float *__restrict__ seg, int *__restrict__ partial,
                                                             unsigned int *__restrict__ counter,
                                                             int *__restrict__ minmax)
{
    __shared__ int s_min[kNK], s_max[kNK];
    __shared__ int s_last;
    const int img = blockIdx.y, tid = threadIdx.x;
    pdl_trigger();
    // Straight-line inner code: every element gets the sigmoid (1 in 18 is the mask channel and keeps its raw value by a
    // select), every element has ONE precomputed destination (its keypoint_heatmaps slot, or its segmentation_masks slot),
    // and the mask channel's min / max registers are simply never merged.  No divergence inside a warp.
    int ch[4];
    bool is_kp[4];
    float *dst[4];                         // destination of element j in tile 0 of this image
    int dst_step[4];                       // ... and its stride from one tile to the next
    float *kh_img = kh + (size_t)img * npix * kNK;
    float *seg_img = seg ? seg + (size_t)img * npix : nullptr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int e = 4 * tid + j, p = e / kCH;
        ch[j] = e - p * kCH;
        is_kp[j] = ch[j] < kNK;
        dst[j] = is_kp[j] ? kh_img + p * kNK + ch[j] : (seg_img ? seg_img + p : nullptr);
        dst_step[j] = is_kp[j] ? kHmPix * kNK : kHmPix;
    }
    float mn[4], mx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { mn[j] = __int_as_float(0x7f800000); mx[j] = 0.0f; }
    if (tid < kNK) { s_min[tid] = 0x7f800000; s_max[tid] = 0; }

    const float4 *src = reinterpret_cast<const float4 *>(hml + (size_t)img * npix * kCH);
    // npix is a multiple of 64 on this path (images are multiples of 128): every tile is full.  Two tiles per trip.
    for (int tile = blockIdx.x; tile < tiles_per_img; tile += 2 * gridDim.x) {
        const int tile2 = tile + gridDim.x;
        const bool two = tile2 < tiles_per_img;
        const float4 qa = __ldcs(src + (size_t)tile * kHmThreads + tid);
        const float4 qb = two ? __ldcs(src + (size_t)tile2 * kHmThreads + tid) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float va[4] = {qa.x, qa.y, qa.z, qa.w}, vb[4] = {qb.x, qb.y, qb.z, qb.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float sa, sb;
            exact_sigmoidf_pair(va[j], vb[j], sa, sb);      // two tiles' values share the packed FFMA2 / FMUL2 stream
            mn[j] = fminf(mn[j], sa); mx[j] = fmaxf(mx[j], sa);
            if (two) { mn[j] = fminf(mn[j], sb); mx[j] = fmaxf(mx[j], sb); }
            if (dst[j]) {
                dst[j][(size_t)tile * dst_step[j]] = is_kp[j] ? sa : va[j];
                if (two) dst[j][(size_t)tile2 * dst_step[j]] = is_kp[j] ? sb : vb[j];
            }
        }
    }
    fold_minmax(ch, mn, mx, s_min, s_max, &s_last, partial, counter, minmax);
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

__global__ void __launch_bounds__(kHmThreads) logit_minmax_kernel(const float *__restrict__ hml, const int npix,
                                                                  const int tiles_per_img, unsigned *__restrict__ partial,
                                                                  unsigned int *__restrict__ counter,
                                                                  float *__restrict__ minmax)
{
    __shared__ unsigned s_min[kNK], s_max[kNK];
    __shared__ int s_last;
    const int img = blockIdx.y, tid = threadIdx.x;
    pdl_trigger();
    int ch[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) ch[j] = (4 * tid + j) % kCH;
    float mn[4], mx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { mn[j] = __int_as_float(0x7f800000); mx[j] = -__int_as_float(0x7f800000); }
    if (tid < kNK) { s_min[tid] = kKeyPosInf; s_max[tid] = kKeyNegInf; }
    const float4 *src = reinterpret_cast<const float4 *>(hml + (size_t)img * npix * kCH);
    // four tiles per trip: four independent 16-byte loads in flight per thread (default caching: pass 2 re-reads the
    // logits, from L2 whenever the call's maps fit)
    for (int tile = blockIdx.x; tile < tiles_per_img; tile += 4 * gridDim.x) {
        float4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int t = tile + u * gridDim.x;          // past the end: the trip's first tile again (min / max unchanged)
            q[u] = __ldg(src + (size_t)(t < tiles_per_img ? t : tile) * kHmThreads + tid);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float v[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { mn[j] = fminf(mn[j], v[j]); mx[j] = fmaxf(mx[j], v[j]); }
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (ch[j] < kNK) {
            atomicMin(&s_min[ch[j]], float_key(mn[j]));
            atomicMax(&s_max[ch[j]], float_key(mx[j]));
        }
    }
    __syncthreads();
    unsigned *my = partial + ((size_t)img * gridDim.x + blockIdx.x) * kNK * 2;
    if (tid < kNK) { my[tid * 2] = s_min[tid]; my[tid * 2 + 1] = s_max[tid]; }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = (atomicAdd(counter + img, 1u) == gridDim.x - 1u);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (tid < kNK) { s_min[tid] = kKeyPosInf; s_max[tid] = kKeyNegInf; }
    __syncthreads();
    if (tid < kNK * 8) {
        const int c = tid % kNK, slice = tid / kNK;
        unsigned lo = kKeyPosInf, hi = kKeyNegInf;
        for (int i = slice; i < (int)gridDim.x; i += 8) {
            const unsigned *q = partial + ((size_t)img * gridDim.x + i) * kNK * 2 + c * 2;
            lo = min(lo, __ldcg(q)); hi = max(hi, __ldcg(q + 1));
        }
        atomicMin(&s_min[c], lo); atomicMax(&s_max[c], hi);
    }
    __syncthreads();
    if (tid < kNK) {
        // the monotone recipe: activations of the extreme logits ARE the extreme activations (create_pb.py:90,92)
        minmax[((size_t)img * kNK + tid) * 2] = exact_sigmoidf(key_float(s_min[tid]));
        minmax[((size_t)img * kNK + tid) * 2 + 1] = exact_sigmoidf(key_float(s_max[tid]));
    }
    if (tid == 0) counter[img] = 0u;
}

// Pass 2.  Same element-to-thread map as heatmap_kernel: thread t of a CTA always sees the channels (4 t + j) mod 18 of
// its 64-pixel tiles, so the four channels' (m, M - m, 1 / (M - m), mask) live in registers; every element has one
// precomputed destination per output (keypoint_heatmaps / segmentation_masks slot, normalised-map slot).  The first
// trip's logits are requested before the wait on pass 1 (they do not depend on it).
__global__ void __launch_bounds__(kHmThreads, 4) heatmap_norm_kernel(const float *__restrict__ hml, const int npix,
                                                                     const int tiles_per_img, float *__restrict__ kh,
                                                                     float *__restrict__ seg,
                                                                     const float *__restrict__ minmax,
                                                                     float *__restrict__ nh)
{
    const int img = blockIdx.y, tid = threadIdx.x;
    pdl_trigger();
    bool is_kp[4];
    int off[4], noff[4];                   // element offsets inside tile 0 of this image: 32-bit, four CTAs per SM
    int ch[4];
    float *kh_img = kh + (size_t)img * npix * kNK;
    float *seg_img = seg ? seg + (size_t)img * npix : nullptr;
    float *nh_img = nh + (size_t)img * npix * kPadCh;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int e = 4 * tid + j, p = e / kCH;
        ch[j] = e - p * kCH;
        is_kp[j] = ch[j] < kNK;
        off[j] = is_kp[j] ? p * kNK + ch[j] : p;
        noff[j] = p * kPadCh + ch[j];
    }
    const float4 *src = reinterpret_cast<const float4 *>(hml + (size_t)img * npix * kCH);
    int tile = blockIdx.x;
    float4 qa = make_float4(0.f, 0.f, 0.f, 0.f), qb = qa;
    if (tile < tiles_per_img) {
        qa = __ldcs(src + (size_t)tile * kHmThreads + tid);
        if (tile + (int)gridDim.x < tiles_per_img) qb = __ldcs(src + (size_t)(tile + gridDim.x) * kHmThreads + tid);
    }
    pdl_wait();                                        // pass 1 (min / max) has completed
    float m[4], d[4], rcp[4], mask[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float lo = is_kp[j] ? __ldg(minmax + ((size_t)img * kNK + ch[j]) * 2) : 0.0f;
        const float hi = is_kp[j] ? __ldg(minmax + ((size_t)img * kNK + ch[j]) * 2 + 1) : 1.0f;
        m[j] = lo;
        d[j] = fsub(hi, lo);
        rcp[j] = range_rcp(d[j]);
        mask[j] = hi > 0.2f ? 1.0f : 0.0f;
    }
    for (; tile < tiles_per_img; tile += 2 * gridDim.x) {
        const int tile2 = tile + gridDim.x;
        const bool two = tile2 < tiles_per_img;
        const float va[4] = {qa.x, qa.y, qa.z, qa.w}, vb[4] = {qb.x, qb.y, qb.z, qb.w};
        // next trip's loads in flight while this one is evaluated
        const int nt = tile + 2 * gridDim.x, nt2 = nt + gridDim.x;
        if (nt < tiles_per_img) qa = __ldcs(src + (size_t)nt * kHmThreads + tid);
        if (nt2 < tiles_per_img) qb = __ldcs(src + (size_t)nt2 * kHmThreads + tid);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float sa, sb;
            exact_sigmoidf_pair(va[j], vb[j], sa, sb);
            if (is_kp[j]) {
                kh_img[(size_t)tile * (kHmPix * kNK) + off[j]] = sa;
                nh_img[(size_t)tile * (kHmPix * kPadCh) + noff[j]] = normalise_tap(sa, m[j], d[j], rcp[j], mask[j]);
                if (two) {
                    kh_img[(size_t)tile2 * (kHmPix * kNK) + off[j]] = sb;
                    nh_img[(size_t)tile2 * (kHmPix * kPadCh) + noff[j]] = normalise_tap(sb, m[j], d[j], rcp[j], mask[j]);
                }
            } else if (seg_img) {
                seg_img[(size_t)tile * kHmPix + off[j]] = va[j];
                if (two) seg_img[(size_t)tile2 * kHmPix + off[j]] = vb[j];
            }
        }
    }
}

// crop_and_resize of the PADDED normalised map (the path of mpn_run): same geometry table as crop_kernel below, but every
// thread takes one (crop pixel, group of 4 channels): four aligned 16-byte tap loads, 12 lerps, and the four results go
// to a shared-memory image of the band, which is then written out in memory order as 16-byte (fp32) and 8-byte (bf16)
// vectors.
struct __align__(16) PixTabP {
    int p_tl, p_tr, p_bl, p_br;      // source pixel indices of the four taps; p_tl < 0: extrapolated (output 0)
    float lx, ly, pad0, pad1;
};
constexpr int kCropPBands = 14;
constexpr int kCropPMaxPix = 160;     // 4 rows x 36 columns = 144 pixels per band

__global__ void __launch_bounds__(256) crop_padded_kernel(const float *__restrict__ src, const int hh, const int ww,
                                                          const float *__restrict__ boxes, const int *__restrict__ box_ind,
                                                          const int *__restrict__ n_dev, const int n_host, const int crop_h,
                                                          const int crop_w, float *__restrict__ out_f32,
                                                          __nv_bfloat16 *__restrict__ out_bf16)
{
    __shared__ PixTabP s_tab[kCropPMaxPix];
    __shared__ __align__(16) float s_out[kCropPMaxPix * kNK];
    const int n = blockIdx.x;
    pdl_trigger();
    pdl_wait();                                        // normalised map and person list are complete
    const int N = n_dev ? *n_dev : n_host;
    if (n >= N) return;
    const int rows_per_band = (crop_h + gridDim.y - 1) / gridDim.y;
    const int cy0 = blockIdx.y * rows_per_band, cy1 = min(crop_h, cy0 + rows_per_band);
    if (cy0 >= cy1) return;
    const int npix = (cy1 - cy0) * crop_w;
    const float4 box = __ldg(reinterpret_cast<const float4 *>(boxes) + n);
    const int b = __ldg(box_ind + n);
    const float y1 = box.x, x1 = box.y, y2 = box.z, x2 = box.w;
    const float hm1 = (float)(hh - 1), wm1 = (float)(ww - 1);
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
        PixTabP t;
        const bool valid = !(in_y < 0.0f || in_y > hm1 || in_x < 0.0f || in_x > wm1);
        const int top = (int)floorf(in_y), bot = (int)ceilf(in_y);
        const int left = (int)floorf(in_x), right = (int)ceilf(in_x);
        t.p_tl = valid ? top * ww + left : -1;
        t.p_tr = valid ? top * ww + right : 0;
        t.p_bl = valid ? bot * ww + left : 0;
        t.p_br = valid ? bot * ww + right : 0;
        t.ly = fsub(in_y, (float)top); t.lx = fsub(in_x, (float)left);
        t.pad0 = 0.0f; t.pad1 = 0.0f;
        s_tab[p] = t;
    }
    __syncthreads();
    const float *img = src + (size_t)b * hh * ww * kPadCh;
    const int total = npix * kGroups;
    for (int f = threadIdx.x; f < total; f += blockDim.x) {
        const int p = f / kGroups, g = f - p * kGroups;
        const int4 tp = *reinterpret_cast<const int4 *>(&s_tab[p].p_tl);
        const float2 w = *reinterpret_cast<const float2 *>(&s_tab[p].lx);
        float r[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        if (tp.x >= 0) {
            const float4 a = __ldg(reinterpret_cast<const float4 *>(img + (size_t)tp.x * kPadCh) + g);
            const float4 bq = __ldg(reinterpret_cast<const float4 *>(img + (size_t)tp.y * kPadCh) + g);
            const float4 cq = __ldg(reinterpret_cast<const float4 *>(img + (size_t)tp.z * kPadCh) + g);
            const float4 d = __ldg(reinterpret_cast<const float4 *>(img + (size_t)tp.w * kPadCh) + g);
            const float tl[4] = {a.x, a.y, a.z, a.w}, tr[4] = {bq.x, bq.y, bq.z, bq.w};
            const float bl[4] = {cq.x, cq.y, cq.z, cq.w}, br[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float tpv = fadd(tl[k], fmul(fsub(tr[k], tl[k]), w.x));
                const float btv = fadd(bl[k], fmul(fsub(br[k], bl[k]), w.x));
                r[k] = fadd(tpv, fmul(fsub(btv, tpv), w.y));
            }
        }
        float *so = s_out + p * kNK + 4 * g;
        so[0] = r[0];
        if (g < kGroups - 1) { so[1] = r[1]; so[2] = r[2]; so[3] = r[3]; }      // group 4 holds channel 16 only
    }
    __syncthreads();
    const int D = crop_h * crop_w * kNK;
    const size_t o0 = (size_t)n * D + (size_t)cy0 * crop_w * kNK;     // multiple of 4 floats (launch_crop checks)
    const int n4 = npix * kNK / 4;
    for (int f = threadIdx.x; f < n4; f += blockDim.x) {
        const float4 v = reinterpret_cast<const float4 *>(s_out)[f];
        if (out_f32) reinterpret_cast<float4 *>(out_f32 + o0)[f] = v;
        if (out_bf16) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
            uint2 u;
            u.x = *reinterpret_cast<const unsigned *>(&lo);
            u.y = *reinterpret_cast<const unsigned *>(&hi);
            reinterpret_cast<uint2 *>(out_bf16 + o0)[f] = u;
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
                                                   const int hh, const int ww, const float *__restrict__ boxes,
                                                   const int *__restrict__ box_ind, const int *__restrict__ n_dev,
                                                   const int n_host, const int crop_h, const int crop_w,
                                                   float *__restrict__ out_f32, __nv_bfloat16 *__restrict__ out_bf16)
{
    __shared__ PixTab s_tab[kCropMaxPix];
    __shared__ float s_m[kNK], s_d[kNK], s_rcp[kNK], s_mask[kNK];
    pdl_trigger();
    pdl_wait();
    const int n = blockIdx.x;
    const int N = n_dev ? *n_dev : n_host;
    if (n >= N) return;
    const int rows_per_band = (crop_h + gridDim.y - 1) / gridDim.y;
    const int cy0 = blockIdx.y * rows_per_band, cy1 = min(crop_h, cy0 + rows_per_band);
    if (cy0 >= cy1) return;
    const int npix = (cy1 - cy0) * crop_w;
    const float4 box = __ldg(reinterpret_cast<const float4 *>(boxes) + n);
    const int b = __ldg(box_ind + n);
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
        const size_t o = (size_t)n * D + j_begin + jl;
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

// pass 1 of the two-pass form: (min, max) of the activations from the min / max of the logits -> minmax_ws [B, 17, 2]
int launch_logit_minmax(const float *hml, int B, int hh, int ww, float *minmax_ws, int *partial_ws, int partial_chunks_cap,
                        unsigned int *counter_ws, int slots, cudaStream_t s)
{
    const int npix = hh * ww, tiles = (npix + kHmPix - 1) / kHmPix;
    int per_img = wave_chunks_per_image(slots, B, tiles, 4);
    if (per_img > partial_chunks_cap) per_img = partial_chunks_cap;
    prof_mark(s, "logit_minmax");
    logit_minmax_kernel<<<dim3(per_img, B), kHmThreads, 0, s>>>(hml, npix, tiles, reinterpret_cast<unsigned *>(partial_ws),
                                                                counter_ws, minmax_ws);
    return 1;
}

// pass 2: keypoint_heatmaps, segmentation_masks and the padded normalised map in one pass over the logits
int launch_heatmap_norm(const float *hml, int B, int hh, int ww, float *kh, float *seg, const float *minmax_ws, float *nh,
                        float *minmax_out, int slots, cudaStream_t s)
{
    const int npix = hh * ww, tiles = (npix + kHmPix - 1) / kHmPix;
    const int per_img = wave_chunks_per_image(slots, B, tiles, 2);
    int launches = 1;
    prof_mark(s, "heatmap_norm");
    launch_k(heatmap_norm_kernel, dim3(per_img, B), dim3(kHmThreads), 0, s, true, hml, npix, tiles, kh, seg, minmax_ws, nh);
    if (minmax_out) {
        const int nmm = B * kNK * 2;
        minmax_copy_kernel<<<(nmm + 255) / 256, 256, 0, s>>>(minmax_ws, minmax_out, nmm);
        ++launches;
    }
    return launches;
}

int launch_crop(const float *src, const float *minmax, int hh, int ww, const float *boxes, const int *box_ind,
                const int *n_dev, int n_host, int n_max, int crop_h, int crop_w, float *crops_f32,
                __nv_bfloat16 *crops_bf16, cudaStream_t s)
{
    if (n_max <= 0) return 0;
    int bands = kCropBands;
    while (((crop_h + bands - 1) / bands) * crop_w > kCropMaxPix) ++bands;
    // every band must start at a multiple of 4 samples: rows_per_band * crop_w * 17 % 4 == 0
    while (bands > 1 && ((((crop_h + bands - 1) / bands) * crop_w * kNK) % 4 != 0)) --bands;
    if (((crop_h + bands - 1) / bands) * crop_w > kCropMaxPix || (crop_h * crop_w * kNK) % 4 != 0)
        return -(int)cudaErrorInvalidValue;
    dim3 grid(n_max, bands);
    prof_mark(s, "crop");
    launch_k(crop_kernel, grid, dim3(256), 0, s, true, src, minmax, hh, ww, boxes, box_ind, n_dev, n_host, crop_h, crop_w,
             crops_f32, crops_bf16);
    return 1;
}

bool crop_padded_supported(int crop_h, int crop_w)
{
    const int rows = (crop_h + kCropPBands - 1) / kCropPBands;
    return rows * crop_w <= kCropPMaxPix && (rows * crop_w * kNK) % 4 == 0 && (crop_h * crop_w * kNK) % 4 == 0;
}

int launch_crop_padded(const float *nh, int hh, int ww, const float *boxes, const int *box_ind, const int *n_dev, int n_host,
                       int n_max, int crop_h, int crop_w, float *crops_f32, __nv_bfloat16 *crops_bf16, cudaStream_t s)
{
    if (n_max <= 0) return 0;
    if (!crop_padded_supported(crop_h, crop_w)) return -(int)cudaErrorInvalidValue;
    dim3 grid(n_max, kCropPBands);
    prof_mark(s, "crop");
    launch_k(crop_padded_kernel, grid, dim3(256), 0, s, true, nh, hh, ww, boxes, box_ind, n_dev, n_host, crop_h, crop_w,
             crops_f32, crops_bf16);
    return 1;
}

int launch_get_keypoints(const float *hm, int hh, int ww, double ymin, double xmin, double ymax, double xmax,
                         double threshold, int *out, cudaStream_t s)
{
    get_keypoints_kernel<<<1, kNK * 32, 0, s>>>(hm, hh, ww, ymin, xmin, ymax, xmax, threshold, out);
    return 1;
}

}  // namespace mpn
