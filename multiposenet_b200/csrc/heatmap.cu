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

__global__ void heatmap_reset_kernel(int *minmax, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) minmax[i] = (i & 1) ? 0 : 0x7f800000;   // (min, max) pairs: +inf, 0  (sigmoid output is >= 0)
}

__global__ void __launch_bounds__(kHmThreads) heatmap_kernel(const float *__restrict__ hml, const int npix,
                                                             const int tiles_per_img, float *__restrict__ kh,
                                                             float *__restrict__ seg, int *__restrict__ minmax)
{
    __shared__ __align__(16) float s_val[kHmThreads * 4];
    __shared__ int s_min[kNK], s_max[kNK];
    const int img = blockIdx.y, tid = threadIdx.x;
    int ch[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) ch[j] = (4 * tid + j) % kCH;
    float mn[4], mx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { mn[j] = __int_as_float(0x7f800000); mx[j] = 0.0f; }
    if (tid < kNK) { s_min[tid] = 0x7f800000; s_max[tid] = 0; }

    for (int tile = blockIdx.x; tile < tiles_per_img; tile += gridDim.x) {
        const int pix0 = tile * kHmPix;
        const int npx = min(kHmPix, npix - pix0);
        const size_t gpix = (size_t)img * npix + pix0;
        float v[4];
        if (4 * tid + 3 < npx * kCH) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(hml + gpix * kCH) + tid);
            v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (ch[j] < kNK) {
                    v[j] = exact_sigmoidf(v[j]);
                    mn[j] = fminf(mn[j], v[j]);
                    mx[j] = fmaxf(mx[j], v[j]);
                }
            }
            *reinterpret_cast<float4 *>(s_val + 4 * tid) = make_float4(v[0], v[1], v[2], v[3]);
        }
        __syncthreads();
        if (tid < (kHmPix * kNK) / 4) {
            if (4 * tid + 3 < npx * kNK) {
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int e = 4 * tid + j;
                    const int p = e / kNK, c = e - p * kNK;
                    o[j] = s_val[p * kCH + c];
                }
                reinterpret_cast<float4 *>(kh + gpix * kNK)[tid] = make_float4(o[0], o[1], o[2], o[3]);
            }
        } else if (seg != nullptr) {
            const int t2 = tid - (kHmPix * kNK) / 4;
            if (4 * t2 + 3 < npx) {
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = s_val[(4 * t2 + j) * kCH + kNK];
                reinterpret_cast<float4 *>(seg + gpix)[t2] = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
        __syncthreads();
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (ch[j] < kNK) {
            atomicMin(&s_min[ch[j]], __float_as_int(mn[j]));
            atomicMax(&s_max[ch[j]], __float_as_int(mx[j]));
        }
    }
    __syncthreads();
    if (tid < kNK) {
        atomicMin(minmax + ((size_t)img * kNK + tid) * 2 + 0, s_min[tid]);
        atomicMax(minmax + ((size_t)img * kNK + tid) * 2 + 1, s_max[tid]);
    }
}

// create_pb.py:93-94 on one tap
__device__ __forceinline__ float normalise_tap(float v, float m, float M, float mask)
{
    return fmul(fdiv(fsub(v, m), fsub(M, m)), mask);
}

// One thread per output sample (n, cy, cx, c); j = (cy*crop_w + cx)*17 + c is also the PRN input column
// (detector/prn.py:17).  Sampling grid and lerp order of TF 1.15 CropAndResize (crop_and_resize_op.cc).
__global__ void __launch_bounds__(256) crop_kernel(const float *__restrict__ kh, const float *__restrict__ minmax,
                                                   const int hh, const int ww, const float *__restrict__ boxes,
                                                   const int *__restrict__ box_ind, const int *__restrict__ n_dev,
                                                   const int n_host, const int crop_h, const int crop_w,
                                                   float *__restrict__ out_f32, __nv_bfloat16 *__restrict__ out_bf16)
{
    const int n = blockIdx.x;
    const int N = n_dev ? *n_dev : n_host;
    if (n >= N) return;
    const int D = crop_h * crop_w * kNK;
    const int j = blockIdx.y * blockDim.x + threadIdx.x;
    if (j >= D) return;
    const int pos = j / kNK, c = j - pos * kNK;
    const int cy = pos / crop_w, cx = pos - cy * crop_w;
    const float4 box = __ldg(reinterpret_cast<const float4 *>(boxes) + n);
    const int b = __ldg(box_ind + n);
    const float y1 = box.x, x1 = box.y, y2 = box.z, x2 = box.w;
    const float hm1 = (float)(hh - 1), wm1 = (float)(ww - 1);
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
    float r = 0.0f;
    if (!(in_y < 0.0f || in_y > hm1 || in_x < 0.0f || in_x > wm1)) {
        const int top = (int)floorf(in_y), bot = (int)ceilf(in_y);
        const int left = (int)floorf(in_x), right = (int)ceilf(in_x);
        const float ly = fsub(in_y, (float)top), lx = fsub(in_x, (float)left);
        const float *img = kh + (size_t)b * hh * ww * kNK + c;
        float tl = __ldg(img + ((size_t)top * ww + left) * kNK), tr = __ldg(img + ((size_t)top * ww + right) * kNK);
        float bl = __ldg(img + ((size_t)bot * ww + left) * kNK), br = __ldg(img + ((size_t)bot * ww + right) * kNK);
        if (minmax != nullptr) {
            const float m = __ldg(minmax + ((size_t)b * kNK + c) * 2), M = __ldg(minmax + ((size_t)b * kNK + c) * 2 + 1);
            const float mask = (M > 0.2f) ? 1.0f : 0.0f;
            tl = normalise_tap(tl, m, M, mask); tr = normalise_tap(tr, m, M, mask);
            bl = normalise_tap(bl, m, M, mask); br = normalise_tap(br, m, M, mask);
        }
        const float t = fadd(tl, fmul(fsub(tr, tl), lx));
        const float bt = fadd(bl, fmul(fsub(br, bl), lx));
        r = fadd(t, fmul(fsub(bt, t), ly));
    }
    if (out_f32) out_f32[(size_t)n * D + j] = r;
    if (out_bf16) out_bf16[(size_t)n * D + j] = __float2bfloat16_rn(r);
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

int launch_heatmaps(const float *hml, int B, int hh, int ww, float *kh, float *seg, float *minmax_ws,
                    float *minmax_out, cudaStream_t s)
{
    const int npix = hh * ww;
    const int tiles = (npix + kHmPix - 1) / kHmPix;
    int per_img = (148 * 7 + B - 1) / B;
    if (per_img > tiles) per_img = tiles;
    if (per_img < 1) per_img = 1;
    int launches = 0;
    const int nmm = B * kNK * 2;
    prof_mark(s, "heatmap_reset");
    heatmap_reset_kernel<<<(nmm + 255) / 256, 256, 0, s>>>(reinterpret_cast<int *>(minmax_ws), nmm);
    ++launches;
    dim3 grid(per_img, B);
    prof_mark(s, "heatmap");
    heatmap_kernel<<<grid, kHmThreads, 0, s>>>(hml, npix, tiles, kh, seg, reinterpret_cast<int *>(minmax_ws));
    ++launches;
    if (minmax_out) {
        minmax_copy_kernel<<<(nmm + 255) / 256, 256, 0, s>>>(minmax_ws, minmax_out, nmm);
        ++launches;
    }
    return launches;
}

int launch_crop(const float *kh, const float *minmax, int hh, int ww, const float *boxes, const int *box_ind,
                const int *n_dev, int n_host, int n_max, int crop_h, int crop_w, float *crops_f32,
                __nv_bfloat16 *crops_bf16, cudaStream_t s)
{
    if (n_max <= 0) return 0;
    const int D = crop_h * crop_w * kNK;
    dim3 grid(n_max, (D + 255) / 256);
    prof_mark(s, "crop");
    crop_kernel<<<grid, 256, 0, s>>>(kh, minmax, hh, ww, boxes, box_ind, n_dev, n_host, crop_h, crop_w, crops_f32,
                                     crops_bf16);
    return 1;
}

int launch_get_keypoints(const float *hm, int hh, int ww, double ymin, double xmin, double ymax, double xmax,
                         double threshold, int *out, cudaStream_t s)
{
    get_keypoints_kernel<<<1, kNK * 32, 0, s>>>(hm, hh, ww, ymin, xmin, ymax, xmax, threshold, out);
    return 1;
}

}  // namespace mpn
