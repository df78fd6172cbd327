/*
 * oracle/mpn_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C, IEEE binary32, no FMA contraction) of the
 * post-backbone inference path of TropComplique/MultiPoseNet:
 *   anchors -> sigmoid/threshold -> decode -> clip -> greedy NMS -> heatmap
 *   sigmoid + min/max normalise -> crop_and_resize -> PRN -> softmax/argmax.
 * Each function cites the reference file:line it follows (paths relative to the
 * reference checkout).  TensorFlow-1.15 library ops the reference calls
 * (NonMaxSuppressionV3, CropAndResize, Softmax, ArgMax, Sigmoid, Exp) are a
 * third-party dependency that is NOT under /root/reference (README.md:15 pins
 * "tensorflow 1.15"); their published algorithms are restated here.
 *
 * PINNING.  The reference ships no tests, golden vectors or fixtures for this path and cannot be executed whole in this
 * environment (TensorFlow is not installable).  What pins this file to the reference ITSELF:
 *   - tests/golden/graph_goldens.npz: vectors produced by importing and EXECUTING the reference's own graph code
 *     (detector/anchor_generator.py, detector/utils/box_utils.py, detector/box_predictor.py reshape_and_concatenate,
 *     detector/retinanet.py get_predictions, detector/utils/nms.py, detector/prn.py, create_pb.py:90-142,
 *     inference/detector.py:49-59) under a numpy-backed TensorFlow stand-in (tests/golden/tf_numpy_shim.py, generator
 *     tests/golden/make_graph_goldens.py).  Everything the reference spells with element-wise / shape ops -- anchors,
 *     decode around exp, clip, thresholds, masks, gather, padding, min-max normalisation, person list, argmax -> positions,
 *     post-filter -- is reproduced by this file BIT FOR BIT (tests/test_graph_goldens.py);
 *   - tests/golden/get_keypoints.npz: outputs of the reference's inference/utils.py::get_keypoints (runs as is).
 * STILL UNPINNED ("parity unpinned" in the sense of the task): the last-ulp behaviour of the TensorFlow LIBRARY kernels
 * the reference calls (Exp, Sigmoid, Softmax, NonMaxSuppressionV3, CropAndResize, MatMul).  Their published algorithms
 * are restated here (and handed to the stand-in when the goldens are made), anchored on the reference's call sites; the
 * hand-derived known answers of tests/test_oracle_kat.py and the independent numpy restatement oracle/spec_np.py
 * cross-check them, torchvision.ops.nms and grid_sample give second opinions.
 *
 * Only tests/, bench.py (cpu_baseline / --impl reference) and
 * __graft_entry__.smoke() may load this library.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -fopenmp -shared -fPIC
 */
#include "exact_math.h"

#include <stdlib.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* exp / sigmoid exposed for vectorised use from Python and for the device-vs-oracle bit test */

ORC_API void orc_expf_array(const float *x, float *y, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) y[i] = orc_expf(x[i]);
}

ORC_API void orc_sigmoidf_array(const float *x, float *y, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) y[i] = orc_sigmoidf(x[i]);
}

/* Walks `count` neighbouring float pairs (x, next float above x) in increasing order from ordered key `key_begin`
 * (keys below 2^31: negative floats, bits ~key; the rest: positive floats, bits key - 2^31) and counts pairs of finite
 * floats with sigmoid(x) > sigmoid(next): the recipe is monotone iff the count over all 2^32 - 1 pairs is 0.  The product
 * relies on that (min / max of the activations from the min / max of the logits, create_pb.py:90,92). */
ORC_API int64_t orc_sigmoid_monotone_violations(uint32_t key_begin, uint64_t count)
{
    int64_t bad = 0;
#pragma omp parallel for reduction(+ : bad) schedule(static)
    for (int64_t blk = 0; blk < (int64_t)((count + 65535) / 65536); ++blk) {
        uint64_t i0 = (uint64_t)blk * 65536, i1 = i0 + 65536 < count ? i0 + 65536 : count;
        for (uint64_t i = i0; i < i1; ++i) {
            uint64_t k = (uint64_t)key_begin + i;
            if (k + 1 > 0xffffffffull) break;
            uint32_t k0 = (uint32_t)k, k1 = (uint32_t)(k + 1);
            float x0 = orc_bits_to_float((k0 & 0x80000000u) ? (k0 & 0x7fffffffu) : ~k0);
            float x1 = orc_bits_to_float((k1 & 0x80000000u) ? (k1 & 0x7fffffffu) : ~k1);
            if (!(fabsf(x0) <= 3.4028234e38f) || !(fabsf(x1) <= 3.4028234e38f)) continue;
            bad += orc_sigmoidf(x0) > orc_sigmoidf(x1);
        }
    }
    return bad;
}

ORC_API void orc_round_bf16_array(const float *x, float *y, int64_t n)
{
    for (int64_t i = 0; i < n; ++i) y[i] = orc_round_bf16(x[i]);
}

/* ------------------------------------------------------------------------- */
/* Anchors: detector/anchor_generator.py:53-114 (levels, offsets, concat, normalise)
 * and tile_anchors :140-165; instantiated at detector/retinanet.py:38-44.       */

static int grid_size(int image, int stride)
{
    /* tf.to_int32(tf.ceil(image_height/stride)), float division (:59-60) */
    return (int)ceilf((float)image / (float)stride);
}

ORC_API int orc_num_anchors(int H, int W, int n_levels, const int *strides, int n_loc)
{
    int a = 0;
    for (int i = 0; i < n_levels; ++i) a += grid_size(H, strides[i]) * grid_size(W, strides[i]) * n_loc;
    return a;
}

ORC_API void orc_anchors(int H, int W, int n_levels, const int *strides, const double *scales,
                         int n_mult, const double *mults, int n_ratio, const double *ratios,
                         float *out /* [A,4] */)
{
    const float fH = (float)H, fW = (float)W;
    int64_t a = 0;
    for (int i = 0; i < n_levels; ++i) {
        const float s = (float)strides[i];
        const int gh = grid_size(H, strides[i]), gw = grid_size(W, strides[i]);
        /* offset = 0.5 * (image - (float(h) - 1.0) * stride)   (:92-93) */
        const float oy = 0.5f * (fH - ((float)gh - 1.0f) * s);
        const float ox = 0.5f * (fW - ((float)gw - 1.0f) * s);
        for (int y = 0; y < gh; ++y) {
            const float cy = (float)y * s + oy;                 /* :148 */
            for (int x = 0; x < gw; ++x) {
                const float cx = (float)x * s + ox;             /* :149 */
                /* pairs = itertools.product(scale_multipliers, aspect_ratios)  (:70) */
                for (int im = 0; im < n_mult; ++im) {
                    for (int ir = 0; ir < n_ratio; ++ir) {
                        /* python double product, then tf.constant(..., float32)  (:75) */
                        const float sc = (float)(mults[im] * scales[i]);
                        const float rs = sqrtf((float)ratios[ir]);          /* :141 */
                        const float ah = sc / rs, aw = sc * rs;             /* :142-143 */
                        const float hh = 0.5f * ah, hw = 0.5f * aw;         /* :163 */
                        out[4 * a + 0] = (cy - hh) / fH;                    /* :110-114 */
                        out[4 * a + 1] = (cx - hw) / fW;
                        out[4 * a + 2] = (cy + hh) / fH;
                        out[4 * a + 3] = (cx + hw) / fW;
                        ++a;
                    }
                }
            }
        }
    }
}

/* ------------------------------------------------------------------------- */
/* decode: detector/utils/box_utils.py:112-139 with to_center_coordinates :63-76
 * and SCALE_FACTORS detector/constants.py:19; clip: detector/utils/nms.py:36      */

static float clip01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }

static void decode_one(const float *code, const float *anc, const float *sf, float *box)
{
    const float ha = anc[2] - anc[0], wa = anc[3] - anc[1];          /* :75 */
    const float cya = anc[0] + 0.5f * ha, cxa = anc[1] + 0.5f * wa;  /* :76 */
    const float ty = code[0] / sf[0], tx = code[1] / sf[1];          /* :127-128 */
    const float th = code[2] / sf[2], tw = code[3] / sf[3];          /* :129-130 */
    const float h = orc_expf(th) * ha, w = orc_expf(tw) * wa;        /* :132-133 */
    const float cy = ty * ha + cya, cx = tx * wa + cxa;              /* :134-135 */
    box[0] = clip01(cy - 0.5f * h);                                  /* :137-139, nms.py:36 */
    box[1] = clip01(cx - 0.5f * w);
    box[2] = clip01(cy + 0.5f * h);
    box[3] = clip01(cx + 0.5f * w);
}

ORC_API void orc_decode(const float *codes, const float *anchors, int64_t n, const float *scale_factors,
                        float *boxes)
{
    for (int64_t i = 0; i < n; ++i) decode_one(codes + 4 * i, anchors + 4 * i, scale_factors, boxes + 4 * i);
}

/* IoU of TensorFlow 1.15 NonMaxSuppressionV3 (tensorflow/core/kernels/non_max_suppression_op.cc,
 * third-party, restated): corners re-ordered, zero-area boxes give 0, no epsilon. */
static float nms_iou(const float *p, const float *q)
{
    const float ymin_i = fminf(p[0], p[2]), xmin_i = fminf(p[1], p[3]);
    const float ymax_i = fmaxf(p[0], p[2]), xmax_i = fmaxf(p[1], p[3]);
    const float ymin_j = fminf(q[0], q[2]), xmin_j = fminf(q[1], q[3]);
    const float ymax_j = fmaxf(q[0], q[2]), xmax_j = fmaxf(q[1], q[3]);
    const float area_i = (ymax_i - ymin_i) * (xmax_i - xmin_i);
    const float area_j = (ymax_j - ymin_j) * (xmax_j - xmin_j);
    if (area_i <= 0.0f || area_j <= 0.0f) return 0.0f;
    const float iymin = fmaxf(ymin_i, ymin_j), ixmin = fmaxf(xmin_i, xmin_j);
    const float iymax = fminf(ymax_i, ymax_j), ixmax = fminf(xmax_i, xmax_j);
    const float inter = fmaxf(iymax - iymin, 0.0f) * fmaxf(ixmax - ixmin, 0.0f);
    return inter / (area_i + area_j - inter);
}

ORC_API float orc_iou(const float *p, const float *q) { return nms_iou(p, q); }

typedef struct { float score; int32_t idx; } cand_t;

static int cand_cmp(const void *a, const void *b)
{
    const cand_t *x = (const cand_t *)a, *y = (const cand_t *)b;
    if (x->score > y->score) return -1;
    if (x->score < y->score) return 1;
    return (x->idx > y->idx) - (x->idx < y->idx);   /* matched tie-break: lower anchor index first */
}

/* One image of batch_non_max_suppression: detector/utils/nms.py:27-53, scores from
 * detector/retinanet.py:73.
 *   cls[A] logits, enc[A,4] codes, anchors[A,4]
 *   boxes_out[max_det,4], scores_out[max_det] zero padded (:47-52), *num_out (:45)
 *   sel_anchor_out[max_det] (optional): anchor index of every kept box, -1 padded
 *   n_conf_out (optional): number of anchors with score >= thr (:30)
 *   min_iou_margin_out (optional): min |IoU - iou_thr| over all IoUs evaluated, for guard-band reporting */
ORC_API void orc_detect_image(const float *cls, const float *enc, const float *anchors, int A,
                              const float *scale_factors, float thr, float iou_thr, int max_det,
                              float *boxes_out, float *scores_out, int32_t *num_out,
                              int32_t *sel_anchor_out, int32_t *n_conf_out, float *min_iou_margin_out)
{
    cand_t *cand = (cand_t *)malloc(sizeof(cand_t) * (size_t)(A > 0 ? A : 1));
    int n_conf = 0, nc = 0;
    for (int a = 0; a < A; ++a) {
        const float s = orc_sigmoidf(cls[a]);                 /* retinanet.py:73 */
        if (s >= thr) {                                       /* nms.py:30 (>=) */
            ++n_conf;
            if (s > thr) { cand[nc].score = s; cand[nc].idx = a; ++nc; }   /* NMS op: strict > */
        }
    }
    qsort(cand, (size_t)nc, sizeof(cand_t), cand_cmp);
    for (int i = 0; i < max_det; ++i) {
        boxes_out[4 * i] = boxes_out[4 * i + 1] = boxes_out[4 * i + 2] = boxes_out[4 * i + 3] = 0.0f;
        scores_out[i] = 0.0f;
        if (sel_anchor_out) sel_anchor_out[i] = -1;
    }
    int kept = 0;
    float margin = INFINITY;
    for (int i = 0; i < nc && kept < max_det; ++i) {
        float box[4];
        const int a = cand[i].idx;
        decode_one(enc + 4 * (int64_t)a, anchors + 4 * (int64_t)a, scale_factors, box);   /* nms.py:35-36 */
        int keep = 1;
        for (int j = kept - 1; j >= 0; --j) {                 /* TF scans selected newest -> oldest */
            const float v = nms_iou(box, boxes_out + 4 * j);
            const float d = fabsf(v - iou_thr);
            if (d < margin) margin = d;
            if (v > iou_thr) { keep = 0; break; }             /* strict > */
        }
        if (keep) {
            boxes_out[4 * kept + 0] = box[0]; boxes_out[4 * kept + 1] = box[1];
            boxes_out[4 * kept + 2] = box[2]; boxes_out[4 * kept + 3] = box[3];
            scores_out[kept] = cand[i].score;
            if (sel_anchor_out) sel_anchor_out[kept] = a;
            ++kept;
        }
    }
    *num_out = kept;
    if (n_conf_out) *n_conf_out = n_conf;
    if (min_iou_margin_out) *min_iou_margin_out = margin;
    free(cand);
}

/* Whole batch; images are independent (tf.map_fn, nms.py:55-60). */
ORC_API void orc_detect_batch(const float *cls, const float *enc, const float *anchors, int B, int A,
                              const float *scale_factors, float thr, float iou_thr, int max_det,
                              float *boxes_out, float *scores_out, int32_t *num_out,
                              int32_t *sel_anchor_out, int32_t *n_conf_out)
{
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b) {
        orc_detect_image(cls + (int64_t)b * A, enc + (int64_t)b * A * 4, anchors, A, scale_factors, thr, iou_thr,
                         max_det, boxes_out + (int64_t)b * max_det * 4, scores_out + (int64_t)b * max_det,
                         num_out + b, sel_anchor_out ? sel_anchor_out + (int64_t)b * max_det : NULL,
                         n_conf_out ? n_conf_out + b : NULL, NULL);
    }
}

/* Plain greedy NMS on given boxes/scores (the TF op alone, nms.py:38-41): returns indices
 * into the input arrays in selection order. */
ORC_API int orc_nms(const float *boxes, const float *scores, int n, float score_thr, float iou_thr, int max_out,
                    int32_t *selected)
{
    cand_t *cand = (cand_t *)malloc(sizeof(cand_t) * (size_t)(n > 0 ? n : 1));
    int nc = 0;
    for (int i = 0; i < n; ++i)
        if (scores[i] > score_thr) { cand[nc].score = scores[i]; cand[nc].idx = i; ++nc; }
    qsort(cand, (size_t)nc, sizeof(cand_t), cand_cmp);
    int kept = 0;
    for (int i = 0; i < nc && kept < max_out; ++i) {
        int keep = 1;
        for (int j = kept - 1; j >= 0; --j)
            if (nms_iou(boxes + 4 * (int64_t)cand[i].idx, boxes + 4 * (int64_t)selected[j]) > iou_thr) { keep = 0; break; }
        if (keep) selected[kept++] = cand[i].idx;
    }
    free(cand);
    return kept;
}

/* ------------------------------------------------------------------------- */
/* Heatmaps: create_pb.py:73-76 (sigmoid of channels 0..nk-1, raw last channel) and
 * the per-(image,channel) min / max of create_pb.py:90,92.
 *   hml[h,w,ch_in] -> kh[h,w,nk], seg[h,w], mn[nk], mx[nk]                         */
ORC_API void orc_heatmaps(const float *hml, int h, int w, int ch_in, int nk, float *kh, float *seg, float *mn,
                          float *mx)
{
    for (int c = 0; c < nk; ++c) { mn[c] = INFINITY; mx[c] = -INFINITY; }
    const int64_t npix = (int64_t)h * w;
    for (int64_t p = 0; p < npix; ++p) {
        const float *src = hml + p * ch_in;
        for (int c = 0; c < nk; ++c) {
            const float v = orc_sigmoidf(src[c]);
            kh[p * nk + c] = v;
            if (v < mn[c]) mn[c] = v;
            if (v > mx[c]) mx[c] = v;
        }
        seg[p] = src[nk];            /* create_pb.py:75: heatmaps[:, :, :, 17], raw logits */
    }
}

ORC_API void orc_heatmaps_batch(const float *hml, int B, int h, int w, int ch_in, int nk, float *kh, float *seg,
                                float *mn, float *mx)
{
    const int64_t npix = (int64_t)h * w;
#pragma omp parallel for schedule(dynamic, 1)
    for (int b = 0; b < B; ++b)
        orc_heatmaps(hml + b * npix * ch_in, h, w, ch_in, nk, kh + b * npix * nk, seg + b * npix, mn + b * nk,
                     mx + b * nk);
}

/* create_pb.py:91,93-94:  hm = (hm - m) / (M - m);  hm *= float(M > 0.2)
 * M == m gives 0/0 = NaN and NaN * 0 = NaN, exactly as the reference graph would. */
static inline float normalise_tap(float v, float m, float M)
{
    const float mask = (M > 0.2f) ? 1.0f : 0.0f;
    return ((v - m) / (M - m)) * mask;
}

/* tf.image.crop_and_resize, bilinear, extrapolation_value 0 (create_pb.py:106-109; TF 1.15
 * tensorflow/core/kernels/crop_and_resize_op.cc, third-party, restated).
 *   img[h,w,nk]; if mn/mx non-NULL every tap is min-max normalised first (create_pb.py:90-94)
 *   box = (y1,x1,y2,x2) normalised; out[crop_h,crop_w,nk]                                  */
ORC_API void orc_crop_and_resize(const float *img, int h, int w, int nk, const float *mn, const float *mx,
                                 const float *box, int crop_h, int crop_w, float *out)
{
    const float y1 = box[0], x1 = box[1], y2 = box[2], x2 = box[3];
    const float hs = (crop_h > 1) ? (y2 - y1) * (float)(h - 1) / (float)(crop_h - 1) : 0.0f;
    const float ws = (crop_w > 1) ? (x2 - x1) * (float)(w - 1) / (float)(crop_w - 1) : 0.0f;
    for (int cy = 0; cy < crop_h; ++cy) {
        const float in_y = (crop_h > 1) ? y1 * (float)(h - 1) + (float)cy * hs : 0.5f * (y1 + y2) * (float)(h - 1);
        float *orow = out + (int64_t)cy * crop_w * nk;
        if (in_y < 0.0f || in_y > (float)(h - 1)) {
            for (int i = 0; i < crop_w * nk; ++i) orow[i] = 0.0f;
            continue;
        }
        const int top = (int)floorf(in_y), bot = (int)ceilf(in_y);
        const float ly = in_y - (float)top;
        for (int cx = 0; cx < crop_w; ++cx) {
            const float in_x =
                (crop_w > 1) ? x1 * (float)(w - 1) + (float)cx * ws : 0.5f * (x1 + x2) * (float)(w - 1);
            float *o = orow + (int64_t)cx * nk;
            if (in_x < 0.0f || in_x > (float)(w - 1)) {
                for (int c = 0; c < nk; ++c) o[c] = 0.0f;
                continue;
            }
            const int left = (int)floorf(in_x), right = (int)ceilf(in_x);
            const float lx = in_x - (float)left;
            const float *ptl = img + ((int64_t)top * w + left) * nk, *ptr = img + ((int64_t)top * w + right) * nk;
            const float *pbl = img + ((int64_t)bot * w + left) * nk, *pbr = img + ((int64_t)bot * w + right) * nk;
            for (int c = 0; c < nk; ++c) {
                float tl = ptl[c], tr = ptr[c], bl = pbl[c], br = pbr[c];
                if (mn) {
                    tl = normalise_tap(tl, mn[c], mx[c]); tr = normalise_tap(tr, mn[c], mx[c]);
                    bl = normalise_tap(bl, mn[c], mx[c]); br = normalise_tap(br, mn[c], mx[c]);
                }
                const float t = tl + (tr - tl) * lx;
                const float b = bl + (br - bl) * lx;
                o[c] = t + (b - t) * ly;
            }
        }
    }
}

/* create_pb.py:96-109: persons of all images in image order; crops[N,crop_h,crop_w,nk] */
ORC_API void orc_crop_batch(const float *kh, int B, int h, int w, int nk, const float *mn, const float *mx,
                            const float *boxes /*[N,4]*/, const int32_t *box_ind /*[N]*/, int N, int crop_h,
                            int crop_w, float *crops)
{
    (void)B;
    const int64_t D = (int64_t)crop_h * crop_w * nk, img = (int64_t)h * w * nk;
#pragma omp parallel for schedule(dynamic, 1)
    for (int n = 0; n < N; ++n) {
        const int b = box_ind[n];
        orc_crop_and_resize(kh + b * img, h, w, nk, mn ? mn + b * nk : NULL, mx ? mx + b * nk : NULL, boxes + 4 * n,
                            crop_h, crop_w, crops + n * D);
    }
}

/* ------------------------------------------------------------------------- */
/* PRN: detector/prn.py:15-25.  x[N,D] -> logits[N,D] = x + relu(relu(x W1 + b1) W2 + b2)
 *   mode 0: fp32 operands, fp64 accumulation (the "true" fp32 answer up to one rounding)
 *   mode 1: bf16 emulation: x, W1, y1, W2 rounded to bf16 (RNE) before each product,
 *           fp64 accumulation, fp32 bias / ReLU / residual -- what a bf16 tensor-core
 *           GEMM with fp32 accumulate computes up to summation order.
 * Weights are [in,out] row-major as slim.fully_connected stores them (prn.py:20,22). */

static void dense_f64(const float *A, int64_t lda, int N, int K, const float *W, int M, const float *bias,
                      int round_ops, float *out, int64_t ldo)
{
    enum { NB = 8, JB = 256 };
    const int n_blocks = (N + NB - 1) / NB, j_blocks = (M + JB - 1) / JB;
#pragma omp parallel for collapse(2) schedule(dynamic, 1)
    for (int nb = 0; nb < n_blocks; ++nb) {
        for (int jb = 0; jb < j_blocks; ++jb) {
            double acc[NB][JB];
            float wrow[JB];
            const int n0 = nb * NB, n1 = (n0 + NB < N) ? n0 + NB : N;
            const int j0 = jb * JB, j1 = (j0 + JB < M) ? j0 + JB : M;
            const int jn = j1 - j0;
            for (int n = 0; n < NB; ++n)
                for (int j = 0; j < JB; ++j) acc[n][j] = 0.0;
            for (int k = 0; k < K; ++k) {
                const float *w = W + (int64_t)k * M + j0;
                if (round_ops) { for (int j = 0; j < jn; ++j) wrow[j] = orc_round_bf16(w[j]); w = wrow; }
                for (int n = n0; n < n1; ++n) {
                    float a = A[(int64_t)n * lda + k];
                    if (round_ops) a = orc_round_bf16(a);
                    if (a == 0.0f) continue;
                    const double ad = (double)a;
                    double *ac = acc[n - n0];
                    for (int j = 0; j < jn; ++j) ac[j] += ad * (double)w[j];
                }
            }
            for (int n = n0; n < n1; ++n)
                for (int j = 0; j < jn; ++j) {
                    const float v = (float)(acc[n - n0][j] + (double)bias[j0 + j]);
                    out[(int64_t)n * ldo + j0 + j] = v > 0.0f ? v : 0.0f;      /* ReLU (prn.py:20,22) */
                }
        }
    }
}

ORC_API void orc_prn(const float *x, int N, int D, int hidden, const float *W1, const float *b1, const float *W2,
                     const float *b2, int mode, float *logits, float *y1_out /* optional [N,hidden] */)
{
    float *y1 = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1) * hidden);
    float *y2 = (float *)malloc(sizeof(float) * (size_t)(N > 0 ? N : 1) * D);
    dense_f64(x, D, N, D, W1, hidden, b1, mode, y1, hidden);          /* prn.py:20 */
    dense_f64(y1, hidden, N, hidden, W2, D, b2, mode, y2, D);         /* prn.py:22 */
    for (int64_t i = 0; i < (int64_t)N * D; ++i) logits[i] = x[i] + y2[i];   /* prn.py:24 */
    if (y1_out) memcpy(y1_out, y1, sizeof(float) * (size_t)N * hidden);
    free(y1); free(y2);
}

/* ------------------------------------------------------------------------- */
/* Keypoint decode: create_pb.py:115-142.  logits[N, crop_h*crop_w, nk]
 *   softmax over positions per (person, channel): p = exp(l - max) * (1 / sum)   (:116-117, Eigen softmax)
 *   idx = first argmax of p (:131), score = max p (:138), pos = (idx // W / H, idx % W / W)  (:133-142)
 *   gap_out (optional) = lmax - (largest logit at any other position), for the parity test's
 *   "decision margin" bookkeeping.                                                               */
ORC_API void orc_keypoint_decode(const float *logits, int N, int crop_h, int crop_w, int nk, float *scores,
                                 float *positions, int32_t *argmax_out, float *gap_out)
{
    const int P = crop_h * crop_w;
#pragma omp parallel for schedule(static)
    for (int n = 0; n < N; ++n) {
        const float *L = logits + (int64_t)n * P * nk;
        float *prob = (float *)malloc(sizeof(float) * (size_t)P);
        for (int c = 0; c < nk; ++c) {
            float lmax = -INFINITY;
            for (int p = 0; p < P; ++p) { const float v = L[(int64_t)p * nk + c]; if (v > lmax) lmax = v; }
            float sum = 0.0f;
            for (int p = 0; p < P; ++p) { prob[p] = orc_expf(L[(int64_t)p * nk + c] - lmax); sum += prob[p]; }
            const float r = 1.0f / sum;
            int best = 0; float pbest = -INFINITY;
            for (int p = 0; p < P; ++p) { const float v = prob[p] * r; if (v > pbest) { pbest = v; best = p; } }
            scores[(int64_t)n * nk + c] = pbest;
            positions[((int64_t)n * nk + c) * 2 + 0] = (float)(best / crop_w) / (float)crop_h;
            positions[((int64_t)n * nk + c) * 2 + 1] = (float)(best % crop_w) / (float)crop_w;
            if (argmax_out) argmax_out[(int64_t)n * nk + c] = best;
            if (gap_out) {
                float second = -INFINITY;
                for (int p = 0; p < P; ++p) { if (p == best) continue; const float v = L[(int64_t)p * nk + c]; if (v > second) second = v; }
                gap_out[(int64_t)n * nk + c] = lmax - second;
            }
        }
        free(prob);
    }
}

/* ------------------------------------------------------------------------- */
/* inference/utils.py:29-52 get_keypoints(heatmaps[h,w,17], box, threshold) -> int32 [17,3] (x, y, visible).
 * Python arithmetic: y * height / h in float64, int() truncates toward zero, np.clip(., 0, height),
 * assignment into an int32 array truncates again.                                               */
ORC_API void orc_get_keypoints(const float *heatmaps, int h, int w, int nk, const double *box, double threshold,
                               int32_t *out)
{
    const double height = box[2] - box[0], width = box[3] - box[1];   /* :40-41 */
    for (int j = 0; j < nk; ++j) {
        float mval = -INFINITY; int best = 0;
        for (int p = 0; p < h * w; ++p) { const float v = heatmaps[(int64_t)p * nk + j]; if (v > mval) { mval = v; best = p; } }
        out[3 * j] = out[3 * j + 1] = out[3 * j + 2] = 0;
        if ((double)mval > threshold) {                               /* :46 strict */
            const int yy = best / w, xx = best % w;                   /* :47 unravel_index, first max */
            double y = trunc((double)yy * height / (double)h);        /* :48 */
            double x = trunc((double)xx * width / (double)w);         /* :49 */
            y = y < 0.0 ? 0.0 : (y > height ? height : y);
            x = x < 0.0 ? 0.0 : (x > width ? width : x);
            out[3 * j] = (int32_t)x; out[3 * j + 1] = (int32_t)y; out[3 * j + 2] = 1;   /* :50 */
        }
    }
}

/* inference/detector.py:54-59 (B == 1): keep rows with score > score_threshold among the first n */
ORC_API int orc_detector_filter(const float *scores, int n, float score_threshold, int32_t *keep_idx)
{
    int k = 0;
    for (int i = 0; i < n; ++i) if (scores[i] > score_threshold) keep_idx[k++] = i;
    return k;
}

ORC_API int orc_version(void) { return 1; }
