// C ABI of libmpn_b200.so (see include/mpn_b200.h).  Host-side orchestration only: argument checking, workspace
// ownership, anchor tables, stream-ordered kernel launches.  There is no CPU implementation of any stage in this
// library: without a CUDA device every entry point fails with MPN_ERR_CUDA.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"
#include "handle.cuh"

using namespace mpn;

namespace mpn {
thread_local Profiler *g_prof = nullptr;
thread_local int g_pdl = 0;
}


namespace {

thread_local char g_create_error[512] = "";

int fail(mpn_handle *h, int code, const char *fmt, ...)
{
    char *dst = h ? h->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

#define MPN_CUDA(h, call)                                                                                  \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess)                                                                             \
            return fail(h, MPN_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int grid_size(int image, int stride) { return (int)std::ceil((float)image / (float)stride); }

int count_anchors(const mpn_config &c, int H, int W)
{
    int a = 0;
    for (int i = 0; i < c.num_levels; ++i) a += grid_size(H, c.strides[i]) * grid_size(W, c.strides[i]);
    return a * c.num_multipliers * c.num_ratios;
}

// detector/anchor_generator.py:53-93,141-145: everything about the anchors that does not depend on (y, x)
void build_anchor_table(const mpn_config &c, int H, int W, AnchorTable *t)
{
    memset(t, 0, sizeof(*t));
    t->n_levels = c.num_levels;
    t->n_loc = c.num_multipliers * c.num_ratios;
    t->fH = (float)H;
    t->fW = (float)W;
    int off = 0;
    for (int i = 0; i < c.num_levels; ++i) {
        const float s = (float)c.strides[i];
        t->gh[i] = grid_size(H, c.strides[i]);
        t->gw[i] = grid_size(W, c.strides[i]);
        t->off[i] = off;
        off += t->gh[i] * t->gw[i] * t->n_loc;
        t->stride[i] = s;
        // offset = 0.5 * (image - (float(h) - 1.0) * stride); host code is built with -ffp-contract=off (no fma)
        const float ty = ((float)t->gh[i] - 1.0f) * s;
        const float tx = ((float)t->gw[i] - 1.0f) * s;
        t->oy[i] = 0.5f * (t->fH - ty);
        t->ox[i] = 0.5f * (t->fW - tx);
        int k = 0;
        for (int im = 0; im < c.num_multipliers; ++im)
            for (int ir = 0; ir < c.num_ratios; ++ir, ++k) {
                const float sc = (float)(c.multipliers[im] * c.scales[i]);   // python double product -> float32 constant
                const float rs = sqrtf((float)c.ratios[ir]);
                const float ah = sc / rs, aw = sc * rs;
                t->half_h[i][k] = 0.5f * ah;
                t->half_w[i][k] = 0.5f * aw;
            }
    }
    t->off[c.num_levels] = off;
    t->num_anchors = off;
    for (int i = 0; i < 4; ++i) t->sf[i] = c.scale_factors[i];
}

// conservative logit pre-filter: every x below it has computed sigmoid(x) <= thr (sigmoid error < 3e-7 relative)
float prefilter_logit(float thr)
{
    if (!(thr > 0.0f)) return -INFINITY;
    if (thr >= 1.0f) return INFINITY;
    const double t = (double)thr * (1.0 - 4e-6);
    const double l = std::log(t / (1.0 - t));
    return (float)(l - std::fabs(l) * 1e-6 - 1e-6);
}

template <typename T>
cudaError_t dalloc(T **p, size_t n)
{
    *p = nullptr;
    if (n == 0) n = 1;
    return cudaMalloc(reinterpret_cast<void **>(p), n * sizeof(T));
}

int check_image_size(mpn_handle *h, int B, int H, int W)
{
    if (B < 1 || H < 1 || W < 1) return fail(h, MPN_ERR_INVALID_ARGUMENT, "batch/height/width must be positive");
    if (H % 128 != 0 || W % 128 != 0)   // inference/detector.py:45, detector/constants.py:4
        return fail(h, MPN_ERR_INVALID_ARGUMENT, "image height and width must be divisible by 128 (got %dx%d)", H, W);
    if (B > h->cfg.max_batch || H > h->cfg.max_height || W > h->cfg.max_width)
        return fail(h, MPN_ERR_CAPACITY, "call %dx%dx%d exceeds handle capacity %dx%dx%d", B, H, W, h->cfg.max_batch,
                    h->cfg.max_height, h->cfg.max_width);
    return MPN_OK;
}

int check_params(mpn_handle *h, const mpn_params *p)
{
    if (!p) return fail(h, MPN_ERR_INVALID_ARGUMENT, "params is NULL");
    if (p->max_detections < 1 || p->max_detections > h->cfg.max_detections)
        return fail(h, MPN_ERR_CAPACITY, "max_detections %d outside [1, %d]", p->max_detections, h->cfg.max_detections);
    if (!(p->iou_threshold >= 0.0f && p->iou_threshold <= 1.0f))
        return fail(h, MPN_ERR_INVALID_ARGUMENT, "iou_threshold must be in [0, 1]");   // as the TF op requires
    if (p->prn_mode != MPN_PRN_FP32 && p->prn_mode != MPN_PRN_BF16)
        return fail(h, MPN_ERR_INVALID_ARGUMENT, "unknown prn_mode %d", p->prn_mode);
    return MPN_OK;
}

void count(mpn_handle *h, int n, bool first)
{
    if (first) h->last_launches = 0;
    h->last_launches += n;
    h->total_launches += n;
}

int launched(mpn_handle *h, int n, bool first, const char *what)
{
    if (n < 0) return fail(h, MPN_ERR_CUDA, "%s: %s", what, cudaGetErrorString((cudaError_t)(-n)));
    count(h, n, first);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return fail(h, MPN_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
    return MPN_OK;
}

// Arms the per-kernel profiler for the duration of one run call and records the closing event.
struct ProfScope {
    mpn_handle *h;
    cudaStream_t s;
    ProfScope(mpn_handle *h_, cudaStream_t s_) : h(h_), s(s_)
    {
        if (h->prof.on) { h->prof.n = 0; g_prof = &h->prof; }
    }
    ~ProfScope()
    {
        if (h->prof.on) { cudaEventRecord(h->prof.ev[h->prof.n], s); g_prof = nullptr; }
    }
};

int do_detect(mpn_handle *h, const mpn_inputs *in, const mpn_params *p, float *boxes, float *scores, int *num_boxes,
              int *sel_anchor, int *n_candidates, cudaStream_t s, bool first, cudaEvent_t after_candidates = nullptr)
{
    int rc = check_image_size(h, in->batch, in->height, in->width);
    if (rc) return rc;
    rc = check_params(h, p);
    if (rc) return rc;
    const bool flat = in->class_logits != nullptr && in->encoded_boxes != nullptr;
    const bool levels = in->level_class != nullptr && in->level_boxes != nullptr;
    if (!flat && !levels)
        return fail(h, MPN_ERR_INVALID_ARGUMENT, "need class_logits+encoded_boxes or level_class+level_boxes");
    if (!boxes || !scores || !num_boxes) return fail(h, MPN_ERR_INVALID_ARGUMENT, "boxes/scores/num_boxes is NULL");
    AnchorTable t;
    build_anchor_table(h->cfg, in->height, in->width, &t);
    DetectArgs a;
    memset(&a, 0, sizeof(a));
    if (flat) {
        if (reinterpret_cast<uintptr_t>(in->encoded_boxes) % 16 != 0)
            return fail(h, MPN_ERR_INVALID_ARGUMENT, "encoded_boxes must be 16-byte aligned");
        a.cls = in->class_logits;
        a.enc = in->encoded_boxes;
        a.enc_ind = h->enc_indirect;
    } else {
        for (int i = 0; i < t.n_levels; ++i) {
            if (!in->level_class[i] || !in->level_boxes[i])
                return fail(h, MPN_ERR_INVALID_ARGUMENT, "level %d pointer is NULL", i);
            a.lv.cls[i] = in->level_class[i];
            a.lv.box[i] = in->level_boxes[i];
        }
    }
    a.B = in->batch;
    a.thr = p->score_threshold;
    a.iou_thr = p->iou_threshold;
    a.pre_thr = prefilter_logit(p->score_threshold);
    a.max_det = p->max_detections;
    a.cand_keys = h->cand_keys;
    a.key_cap = h->key_cap;
    a.cand_count = h->cand_count;
    a.boxes = boxes;
    a.scores = scores;
    a.num_boxes = num_boxes;
    a.sel_anchor = sel_anchor;
    a.n_candidates = n_candidates;
    a.trace = h->nms_trace;
    return launched(h, launch_detect(t, a, s, after_candidates), first, "detect");
}

int do_prn(mpn_handle *h, const float *x_f32, const __nv_bfloat16 *x_bf16, const int *n_dev, int n_host, int n_max,
           int mode, float *logits, cudaStream_t s, bool first, const FusedCropCall *fc = nullptr)
{
    if (!h->have_weights) return fail(h, MPN_ERR_NO_WEIGHTS, "mpn_set_prn_weights has not been called");
    PrnWeights w;
    w.D = h->D; w.hidden = h->cfg.prn_hidden;
    w.W1 = h->W1; w.b1 = h->b1; w.W2 = h->W2; w.b2 = h->b2; w.W1t = h->W1t; w.W2t = h->W2t;
    if (mode == MPN_PRN_FP32) {
        if (!(h->cfg.prn_modes & 1)) return fail(h, MPN_ERR_UNSUPPORTED, "handle was created without the fp32 PRN");
        // <= 240 persons: the fp32-accurate tensor-core kernel (prn_split3.cu: three bf16 parts per number, one pass over
        // the weights per group of 80 persons); more: the SIMT kernels.  The count is only known on the device: both are launched when the call's
        // capacity exceeds 240, and each exits at once outside its regime.
        int skip_le = 0;
        if (h->split3) {
            int rc = launched(h, launch_prn_split3(h, x_f32, n_dev, n_host, n_max, logits, s), first, "prn split3");
            if (rc || n_max <= kPrnSplit3MaxRows) return rc;
            first = false;
            skip_le = kPrnSplit3MaxRows;
        }
        return launched(h, launch_prn_fp32(w, h->prn_ws, x_f32, n_dev, n_host, n_max, logits, skip_le, s), first, "prn fp32");
    }
    if (!(h->cfg.prn_modes & 2)) return fail(h, MPN_ERR_UNSUPPORTED, "handle was created without the bf16 PRN");
    // <= 256 persons: one persistent kernel (prn_fused.cu); above: the tiled GEMM kernels.  The person count is only
    // known on the device, so when the call's capacity exceeds 256 both are launched and each exits at once outside
    // its regime.
    if (h->fused) {
        int rc = launched(h, launch_prn_fused(h, x_f32, n_dev, n_host, n_max, logits, s, fc), first, "prn fused");
        if (rc || n_max <= kPrnFusedMaxRows) return rc;
        first = false;
    }
    if (h->big)
        return launched(h, launch_prn_big(h, x_f32, n_dev, n_host, n_max, logits, h->fused ? kPrnFusedMaxRows : 0, s), first,
                        "prn big");
    return launched(h, launch_prn_bf16(w, h->prn_ws, x_f32, x_bf16, n_dev, n_host, n_max, logits, h->tmaps,
                                       h->fused ? kPrnFusedMaxRows : 0, s), first, "prn bf16");
}

}  // namespace

extern "C" {

int mpn_version(void) { return 100; }

const char *mpn_last_error(const mpn_handle *h) { return h ? h->err : g_create_error; }

int mpn_default_config(mpn_config *c)
{
    if (!c) return MPN_ERR_INVALID_ARGUMENT;
    memset(c, 0, sizeof(*c));
    c->struct_size = (int32_t)sizeof(mpn_config);
    c->device = 0;
    c->max_batch = 1;
    c->max_height = 640;
    c->max_width = 640;
    c->max_detections = 25;                           // create_pb.py:35
    c->num_levels = 5;                                // detector/retinanet.py:38-43
    const int strides[5] = {8, 16, 32, 64, 128};
    const double scales[5] = {32, 64, 128, 256, 512};
    for (int i = 0; i < 5; ++i) { c->strides[i] = strides[i]; c->scales[i] = scales[i]; }
    c->num_multipliers = 2; c->multipliers[0] = 1.0; c->multipliers[1] = 1.4142;
    c->num_ratios = 3; c->ratios[0] = 1.0; c->ratios[1] = 2.0; c->ratios[2] = 0.5;
    c->scale_factors[0] = 10.0f; c->scale_factors[1] = 10.0f; c->scale_factors[2] = 5.0f; c->scale_factors[3] = 5.0f;
    c->crop_height = 56; c->crop_width = 36;          // create_pb.py:19
    c->num_keypoints = 17; c->downsample = 4;         // detector/constants.py:10,13
    c->prn_hidden = 1024;                             // detector/prn.py:20
    c->prn_modes = 3;
    return MPN_OK;
}

int mpn_create(const mpn_config *cfg, mpn_handle **out)
{
    if (!cfg || !out) return fail(nullptr, MPN_ERR_INVALID_ARGUMENT, "cfg/out is NULL");
    *out = nullptr;
    if (cfg->struct_size != (int32_t)sizeof(mpn_config))
        return fail(nullptr, MPN_ERR_INVALID_ARGUMENT, "mpn_config size mismatch (%d vs %d)", cfg->struct_size,
                    (int)sizeof(mpn_config));
    if (cfg->num_levels < 1 || cfg->num_levels > kMaxLevels || cfg->num_multipliers < 1 || cfg->num_ratios < 1 ||
        cfg->num_multipliers * cfg->num_ratios > kMaxShapes)
        return fail(nullptr, MPN_ERR_INVALID_ARGUMENT, "bad anchor specification");
    if (cfg->num_keypoints != 17 || cfg->downsample != 4)
        return fail(nullptr, MPN_ERR_UNSUPPORTED, "kernels are built for 17 keypoints (+1 mask channel), downsample 4");
    if (cfg->crop_height < 1 || cfg->crop_width < 1 || cfg->crop_height * cfg->crop_width > 2048)
        return fail(nullptr, MPN_ERR_UNSUPPORTED, "crop size must have at most 2048 positions");
    if (cfg->max_batch < 1 || cfg->max_detections < 1 || cfg->max_detections > kMaxDetCap)
        return fail(nullptr, MPN_ERR_INVALID_ARGUMENT, "max_batch >= 1 and 1 <= max_detections <= %d required", kMaxDetCap);
    if (cfg->max_height % 128 || cfg->max_width % 128 || cfg->max_height < 128 || cfg->max_width < 128)
        return fail(nullptr, MPN_ERR_INVALID_ARGUMENT, "max_height/max_width must be positive multiples of 128");
    if (cfg->prn_hidden % 64 != 0) return fail(nullptr, MPN_ERR_UNSUPPORTED, "prn_hidden must be a multiple of 64");
    for (int i = 0; i < cfg->num_levels; ++i)
        if (cfg->strides[i] < 1) return fail(nullptr, MPN_ERR_INVALID_ARGUMENT, "stride %d is not positive", i);

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, MPN_ERR_CUDA, "no CUDA device (%s); this library has no CPU path",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, MPN_ERR_INVALID_ARGUMENT, "device %d of %d", cfg->device, ndev);
    MPN_CUDA(nullptr, cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    MPN_CUDA(nullptr, cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(nullptr, MPN_ERR_UNSUPPORTED, "kernels are built for sm_100a only; device is sm_%d%d", prop.major, prop.minor);

    mpn_handle *h = new (std::nothrow) mpn_handle;
    if (!h) return fail(nullptr, MPN_ERR_CUDA, "out of host memory");
    memset(h, 0, sizeof(*h));
    h->cfg = *cfg;
    h->D = cfg->crop_height * cfg->crop_width * cfg->num_keypoints;
    if (h->D % 32 != 0) {
        const int d = h->D;
        delete h;
        return fail(nullptr, MPN_ERR_UNSUPPORTED, "PRN width %d must be a multiple of 32", d);
    }
    h->max_anchors = count_anchors(*cfg, cfg->max_height, cfg->max_width);
    h->key_cap = 1;
    while (h->key_cap < h->max_anchors) h->key_cap <<= 1;
    h->max_hm_pix = (cfg->max_height / 4) * (cfg->max_width / 4);
    h->max_persons = cfg->max_batch * cfg->max_detections;
    const size_t B = cfg->max_batch, NP = h->max_persons, D = h->D, Hd = cfg->prn_hidden;
    const size_t NPpad = (NP + 127) / 128 * 128;

#define MPN_ALLOC(ptr, n)                                                                                   \
    do {                                                                                                    \
        cudaError_t e2 = dalloc(&(ptr), (n));                                                               \
        if (e2 != cudaSuccess) {                                                                            \
            fail(nullptr, MPN_ERR_CUDA, "cudaMalloc(%s, %zu elements) failed: %s", #ptr, (size_t)(n), cudaGetErrorString(e2)); \
            mpn_destroy(h);                                                                                 \
            return MPN_ERR_CUDA;                                                                            \
        }                                                                                                   \
    } while (0)

    MPN_ALLOC(h->cand_keys, B * h->key_cap);
    MPN_ALLOC(h->cand_count, B);
    MPN_ALLOC(h->person_box, NP * 4);
    MPN_ALLOC(h->person_img, NP);
    MPN_ALLOC(h->person_offsets, B + 1);
    MPN_ALLOC(h->kh_ws, B * h->max_hm_pix * 17);
    MPN_ALLOC(h->nh_ws, B * h->max_hm_pix * 20);   // padded pixels (heatmap.cu: kPadCh)
    MPN_ALLOC(h->minmax_ws, B * 17 * 2);
    h->hm_partial_chunks = (h->max_hm_pix / 64 + 1) / 2 + 1;      // per image; every heatmap launcher stays within it
    MPN_ALLOC(h->hm_partial, B * (size_t)h->hm_partial_chunks * 17 * 2);
    MPN_ALLOC(h->hm_partial2, B * (size_t)h->hm_partial_chunks * 17 * 2);
    MPN_ALLOC(h->hm_counter, B);
    cudaMemset(h->hm_counter, 0, B * sizeof(unsigned int));
    cudaMemset(h->nh_ws, 0, B * h->max_hm_pix * 20 * sizeof(float));     // pad channels 17..19 stay zero for good
    MPN_ALLOC(h->crops_f32, NPpad * D);
    MPN_ALLOC(h->logits, NPpad * D);
    MPN_ALLOC(h->b1, Hd);
    MPN_ALLOC(h->b2, D);
    if (cfg->prn_modes & 1) {
        MPN_ALLOC(h->W1, D * Hd);
        MPN_ALLOC(h->W2, Hd * D);
        h->prn_ws.partial_floats = (NPpad > 2400 ? NPpad : 2400) * Hd;     // mpn_prn accepts up to NPpad rows
        MPN_ALLOC(h->prn_ws.partial, h->prn_ws.partial_floats);
        MPN_ALLOC(h->prn_ws.y1, NPpad * Hd);
    }
    if (cfg->prn_modes & 2) {
        MPN_ALLOC(h->W1t, D * Hd);
        MPN_ALLOC(h->W2t, Hd * D);
        MPN_ALLOC(h->crops_bf16, NPpad * D);
        MPN_ALLOC(h->prn_ws.y1_bf16, NPpad * Hd);
        if (!h->prn_ws.partial) {
            h->prn_ws.partial_floats = (NPpad > 2432 ? NPpad : 2432) * Hd;
            MPN_ALLOC(h->prn_ws.partial, h->prn_ws.partial_floats);
        }
        cudaMemset(h->crops_bf16, 0, NPpad * D * sizeof(__nv_bfloat16));
        cudaMemset(h->prn_ws.y1_bf16, 0, NPpad * Hd * sizeof(__nv_bfloat16));
    }
    h->prn_ws.n_max = (int)NPpad;
    cudaMemset(h->cand_count, 0, B * sizeof(int));
    cudaMemset(h->person_offsets, 0, (B + 1) * sizeof(int));
    e = cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);     // numerically lower = higher priority
        e = cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, prio_hi);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->own_event, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_cand, cudaEventDisableTiming);
    h->cfg_use_graphs = getenv("MPN_NO_GRAPH") == nullptr;
    h->use_pdl = getenv("MPN_NO_PDL") == nullptr;
    {   // crop_and_resize inside the single-kernel PRN (prn_fused.cu: CropFuse); MPN_FUSE_CROP=0/1 overrides the default
        const char *fc = getenv("MPN_FUSE_CROP");
        h->fuse_crop = fc ? fc[0] == '1' : false;
    }
    if (e != cudaSuccess) {
        fail(nullptr, MPN_ERR_CUDA, "stream/event creation failed: %s", cudaGetErrorString(e));
        mpn_destroy(h);
        return MPN_ERR_CUDA;
    }
    if (heatmap_prepare(&h->waves) != 0) {
        fail(nullptr, MPN_ERR_CUDA, "heatmap kernel setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        mpn_destroy(h);
        return MPN_ERR_CUDA;
    }
    if (detect_prepare() != 0) {
        fail(nullptr, MPN_ERR_CUDA, "sort/NMS kernel setup failed: %s", cudaGetErrorString(cudaGetLastError()));
        mpn_destroy(h);
        return MPN_ERR_CUDA;
    }
    {
        const int rk = kpdecode_prepare(h->own_stream, cfg->crop_height, cfg->crop_width, &h->decode_clusters);
        if (rk == -2) {
            fail(nullptr, MPN_ERR_UNSUPPORTED, "crop size %dx%d is not covered by the keypoint decode kernel", cfg->crop_height,
                 cfg->crop_width);
            mpn_destroy(h);
            return MPN_ERR_UNSUPPORTED;
        }
        if (rk != 0 || cudaStreamSynchronize(h->own_stream) != cudaSuccess) {
            fail(nullptr, MPN_ERR_CUDA, "keypoint decode setup failed: %s", cudaGetErrorString(cudaGetLastError()));
            mpn_destroy(h);
            return MPN_ERR_CUDA;
        }
    }
    if (cfg->prn_modes & 2) {
        int rc = prn_bf16_prepare(h);
        if (rc == MPN_OK) rc = prn_fused_prepare(h);
        if (rc == MPN_OK && getenv("MPN_NO_BIG_GEMM") == nullptr) rc = prn_big_prepare(h);
        if (rc != MPN_OK) {
            snprintf(g_create_error, sizeof(g_create_error), "%s", h->err);
            mpn_destroy(h);
            return rc;
        }
    }
    if (cfg->prn_modes & 1) {
        const int rc = prn_split3_prepare(h);
        if (rc != MPN_OK) {
            snprintf(g_create_error, sizeof(g_create_error), "%s", h->err);
            mpn_destroy(h);
            return rc;
        }
    }
    *out = h;
    return MPN_OK;
}

void mpn_destroy(mpn_handle *h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    prn_bf16_release(h);
    prn_fused_release(h);
    prn_split3_release(h);
    prn_big_release(h);
    void *ptrs[] = {h->cand_keys, h->cand_count, h->person_box, h->person_img, h->person_offsets,
                    h->kh_ws, h->nh_ws, h->minmax_ws, h->hm_partial, h->hm_partial2, h->hm_counter, h->nms_trace, h->crops_f32, h->logits, h->crops_bf16, h->W1, h->b1, h->W2, h->b2, h->W1t,
                    h->W2t, h->prn_ws.partial, h->prn_ws.y1, h->prn_ws.y1_bf16};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (HostSlot &sl : h->slots) {
        void *sp[] = {sl.cls, sl.enc, sl.hml, sl.kh, sl.seg, sl.small_dev, (void *)sl.enc_word};
        for (void *p : sp)
            if (p) cudaFree(p);
        if (sl.small_host) cudaFreeHost(sl.small_host);
        if (sl.ev_in) cudaEventDestroy(sl.ev_in);
        if (sl.ev_comp) cudaEventDestroy(sl.ev_comp);
        if (sl.ev_out) cudaEventDestroy(sl.ev_out);
    }
    if (h->in_stream) cudaStreamDestroy(h->in_stream);
    if (h->out_stream) cudaStreamDestroy(h->out_stream);
    if (h->prof_events_ready)
        for (int i = 0; i <= kMaxMarks; ++i) cudaEventDestroy(h->prof.ev[i]);
    for (GraphEntry &g : h->graphs)
        if (g.exec) cudaGraphExecDestroy(g.exec);
    if (h->own_event) cudaEventDestroy(h->own_event);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_cand) cudaEventDestroy(h->ev_cand);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    if (h->aux_stream) cudaStreamDestroy(h->aux_stream);
    delete h;
}

int mpn_num_anchors(const mpn_handle *h, int32_t height, int32_t width)
{
    if (!h || height < 1 || width < 1) return MPN_ERR_INVALID_ARGUMENT;
    return count_anchors(h->cfg, height, width);
}

int mpn_set_prn_weights(mpn_handle *h, const float *W1, const float *b1, const float *W2, const float *b2)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!W1 || !b1 || !W2 || !b2) return fail(h, MPN_ERR_INVALID_ARGUMENT, "weight pointer is NULL");
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    const size_t D = h->D, Hd = h->cfg.prn_hidden;
    cudaStream_t s = h->own_stream;
    // fp32 master copies; when only the bf16 PRN is configured they are staged through the logits workspace
    float *dW1 = h->W1, *dW2 = h->W2;
    const bool temp = !(h->cfg.prn_modes & 1);
    if (temp) {
        MPN_CUDA(h, dalloc(&dW1, D * Hd));
        cudaError_t e = dalloc(&dW2, Hd * D);
        if (e != cudaSuccess) { cudaFree(dW1); return fail(h, MPN_ERR_CUDA, "cudaMalloc failed: %s", cudaGetErrorString(e)); }
    }
    MPN_CUDA(h, cudaMemcpyAsync(dW1, W1, D * Hd * sizeof(float), cudaMemcpyHostToDevice, s));
    MPN_CUDA(h, cudaMemcpyAsync(dW2, W2, Hd * D * sizeof(float), cudaMemcpyHostToDevice, s));
    MPN_CUDA(h, cudaMemcpyAsync(h->b1, b1, Hd * sizeof(float), cudaMemcpyHostToDevice, s));
    MPN_CUDA(h, cudaMemcpyAsync(h->b2, b2, D * sizeof(float), cudaMemcpyHostToDevice, s));
    if (h->cfg.prn_modes & 2) {
        // bf16 B operands, K-major: W1t [hidden, D], W2t [D, hidden]
        launch_transpose_to_bf16(dW1, (int)D, (int)Hd, h->W1t, s);
        launch_transpose_to_bf16(dW2, (int)Hd, (int)D, h->W2t, s);
    }
    if (h->cfg.prn_modes & 1) prn_split3_set_weights(h, dW1, dW2, s);
    MPN_CUDA(h, cudaStreamSynchronize(s));
    if (temp) { cudaFree(dW1); cudaFree(dW2); }
    MPN_CUDA(h, cudaGetLastError());
    h->have_weights = true;
    return MPN_OK;
}

int mpn_anchors(mpn_handle *h, int32_t height, int32_t width, float *anchors_out, void *stream)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!anchors_out || height < 1 || width < 1) return fail(h, MPN_ERR_INVALID_ARGUMENT, "bad arguments");
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    AnchorTable t;
    build_anchor_table(h->cfg, height, width, &t);
    return launched(h, launch_anchors(t, anchors_out, (cudaStream_t)stream), true, "anchors");
}

int mpn_detect(mpn_handle *h, const mpn_inputs *in, const mpn_params *p, float *boxes, float *scores,
               int32_t *num_boxes, int32_t *sel_anchor, int32_t *n_candidates, void *stream)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!in) return fail(h, MPN_ERR_INVALID_ARGUMENT, "inputs is NULL");
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    return do_detect(h, in, p, boxes, scores, num_boxes, sel_anchor, n_candidates, (cudaStream_t)stream, true);
}

int mpn_heatmaps(mpn_handle *h, const float *heatmap_logits, int32_t batch, int32_t hm_height, int32_t hm_width,
                 float *keypoint_heatmaps, float *segmentation_masks, float *minmax, void *stream)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!heatmap_logits || !keypoint_heatmaps) return fail(h, MPN_ERR_INVALID_ARGUMENT, "heatmap pointer is NULL");
    if (batch < 1 || batch > h->cfg.max_batch) return fail(h, MPN_ERR_CAPACITY, "batch %d outside [1, %d]", batch, h->cfg.max_batch);
    if (hm_height < 1 || hm_width < 1 || (hm_height * hm_width) % 64 != 0)
        return fail(h, MPN_ERR_UNSUPPORTED, "heatmap pixel count must be a multiple of 64 (it is for images divisible by 128)");
    if ((long long)hm_height * hm_width > h->max_hm_pix)     // the per-CTA min / max array is sized for the handle's capacity
        return fail(h, MPN_ERR_CAPACITY, "heatmap %dx%d exceeds the handle's capacity of %d pixels", hm_height, hm_width, h->max_hm_pix);
    if (reinterpret_cast<uintptr_t>(heatmap_logits) % 16 != 0)          // tiles arrive by 16-byte-granular bulk copies
        return fail(h, MPN_ERR_INVALID_ARGUMENT, "heatmap_logits must be 16-byte aligned");
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    return launched(h, launch_heatmaps(heatmap_logits, batch, hm_height, hm_width, keypoint_heatmaps, segmentation_masks,
                                       h->minmax_ws, minmax, h->hm_partial, h->hm_counter, h->waves.one_pass, (cudaStream_t)stream),
                    true, "heatmaps");
}

int mpn_heatmaps_normalised(mpn_handle *h, const float *heatmap_logits, int32_t batch, int32_t hm_height, int32_t hm_width,
                            float *keypoint_heatmaps, float *segmentation_masks, float *minmax, float *normalised, void *stream)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!heatmap_logits || !keypoint_heatmaps) return fail(h, MPN_ERR_INVALID_ARGUMENT, "heatmap pointer is NULL");
    if (batch < 1 || batch > h->cfg.max_batch) return fail(h, MPN_ERR_CAPACITY, "batch %d outside [1, %d]", batch, h->cfg.max_batch);
    if (hm_height < 1 || hm_width < 1 || (hm_height * hm_width) % 64 != 0)
        return fail(h, MPN_ERR_UNSUPPORTED, "heatmap pixel count must be a multiple of 64 (it is for images divisible by 128)");
    if ((long long)hm_height * hm_width > h->max_hm_pix)
        return fail(h, MPN_ERR_CAPACITY, "heatmap %dx%d exceeds the handle's capacity of %d pixels", hm_height, hm_width, h->max_hm_pix);
    if (reinterpret_cast<uintptr_t>(heatmap_logits) % 16 != 0 || (normalised && reinterpret_cast<uintptr_t>(normalised) % 16 != 0))
        return fail(h, MPN_ERR_INVALID_ARGUMENT, "heatmap_logits / normalised must be 16-byte aligned");
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    int n_chunks = 0;
    int rc = launched(h, launch_logit_minmax(heatmap_logits, batch, hm_height, hm_width, h->hm_partial2, h->hm_partial_chunks,
                                             h->waves.minmax, &n_chunks, s), true, "logit min/max");
    if (rc) return rc;
    return launched(h, launch_heatmap_norm(heatmap_logits, batch, hm_height, hm_width, keypoint_heatmaps, segmentation_masks,
                                           h->hm_partial2, n_chunks, h->minmax_ws, normalised ? normalised : h->nh_ws, minmax,
                                           h->waves.norm, s), false, "heatmaps + normalise");
}

int mpn_crop_padded(mpn_handle *h, const float *normalised, int32_t batch, int32_t hm_height, int32_t hm_width,
                    const float *boxes, const int32_t *box_ind, int32_t n, float *crops_f32, uint16_t *crops_bf16, void *stream)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!normalised || !boxes || !box_ind || (!crops_f32 && !crops_bf16)) return fail(h, MPN_ERR_INVALID_ARGUMENT, "pointer is NULL");
    if (n < 0 || batch < 1 || hm_height < 1 || hm_width < 1) return fail(h, MPN_ERR_INVALID_ARGUMENT, "bad sizes");
    if (reinterpret_cast<uintptr_t>(normalised) % 16 != 0 || reinterpret_cast<uintptr_t>(boxes) % 16 != 0 ||
        reinterpret_cast<uintptr_t>(crops_f32) % 16 != 0 || reinterpret_cast<uintptr_t>(crops_bf16) % 8 != 0)
        return fail(h, MPN_ERR_INVALID_ARGUMENT, "normalised / boxes / crops must be 16-byte aligned");
    if (!crop_padded_supported(h->cfg.crop_height, h->cfg.crop_width))
        return fail(h, MPN_ERR_UNSUPPORTED, "the padded crop kernel does not cover this crop size");
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    if (n == 0) { h->last_launches = 0; return MPN_OK; }
    PersonList pl;
    memset(&pl, 0, sizeof(pl));
    pl.boxes = boxes; pl.box_ind = box_ind; pl.n_host = n;
    return launched(h, launch_crop_padded(normalised, hm_height, hm_width, pl, n, h->cfg.crop_height, h->cfg.crop_width,
                                          crops_f32, reinterpret_cast<__nv_bfloat16 *>(crops_bf16), (cudaStream_t)stream), true,
                    "crop (padded)");
}

int mpn_heatmap_head(mpn_handle *h, const float *features, const float *weight, const float *bias, int32_t batch,
                     int32_t hm_height, int32_t hm_width, float *heatmap_logits, float *keypoint_heatmaps,
                     float *segmentation_masks, float *minmax, void *stream)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!features || !weight || !bias || !keypoint_heatmaps) return fail(h, MPN_ERR_INVALID_ARGUMENT, "pointer is NULL");
    if (batch < 1 || batch > h->cfg.max_batch) return fail(h, MPN_ERR_CAPACITY, "batch %d outside [1, %d]", batch, h->cfg.max_batch);
    if (hm_height < 1 || hm_width < 1 || (hm_height * hm_width) % 256 != 0 || hm_height * hm_width > h->max_hm_pix)
        return fail(h, MPN_ERR_UNSUPPORTED, "heatmap pixel count must be a multiple of 256 within the handle's capacity");
    if (reinterpret_cast<uintptr_t>(features) % 16 != 0 || (heatmap_logits && reinterpret_cast<uintptr_t>(heatmap_logits) % 8 != 0))
        return fail(h, MPN_ERR_INVALID_ARGUMENT, "features must be 16-byte aligned (heatmap_logits 8-byte)");
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    return launched(h, launch_heatmap_head(features, weight, bias, batch, hm_height, hm_width, heatmap_logits,
                                           keypoint_heatmaps, segmentation_masks, h->minmax_ws, minmax, h->hm_partial,
                                           h->hm_counter, (cudaStream_t)stream), true, "heatmap head");
}

int mpn_crop(mpn_handle *h, const float *keypoint_heatmaps, const float *minmax, int32_t batch, int32_t hm_height,
             int32_t hm_width, const float *boxes, const int32_t *box_ind, int32_t n, float *crops, void *stream)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!keypoint_heatmaps || !boxes || !box_ind || !crops) return fail(h, MPN_ERR_INVALID_ARGUMENT, "pointer is NULL");
    if (n < 0 || batch < 1 || hm_height < 1 || hm_width < 1) return fail(h, MPN_ERR_INVALID_ARGUMENT, "bad sizes");
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    if (n == 0) { h->last_launches = 0; return MPN_OK; }
    PersonList pl;
    memset(&pl, 0, sizeof(pl));
    pl.boxes = boxes; pl.box_ind = box_ind; pl.n_host = n;
    return launched(h, launch_crop(keypoint_heatmaps, minmax, hm_height, hm_width, pl, n, h->cfg.crop_height, h->cfg.crop_width,
                                   crops, nullptr, (cudaStream_t)stream), true, "crop");
}

int mpn_prn(mpn_handle *h, const float *crops, int32_t n, int32_t prn_mode, float *logits, void *stream)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!crops || !logits) return fail(h, MPN_ERR_INVALID_ARGUMENT, "pointer is NULL");
    if (n < 0 || n > h->prn_ws.n_max) return fail(h, MPN_ERR_CAPACITY, "n %d outside [0, %d]", n, h->prn_ws.n_max);
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    if (n == 0) { h->last_launches = 0; return MPN_OK; }
    if (crops == logits && (prn_mode != MPN_PRN_BF16 || !(h->big || (h->fused && n <= kPrnFusedMaxRows))))
        return fail(h, MPN_ERR_UNSUPPORTED, "in-place PRN (logits == crops) needs bf16 mode and a covered shape");
    ProfScope prof_scope(h, s);                        // per-kernel times of this call when profiling is on
    bool first = true;
    if (prn_mode == MPN_PRN_BF16) {
        if (!h->crops_bf16) return fail(h, MPN_ERR_UNSUPPORTED, "handle was created without the bf16 PRN");
        int rc = launched(h, launch_f32_to_bf16(crops, h->crops_bf16, nullptr, n, h->D, n, s), true, "bf16 convert");
        if (rc) return rc;
        first = false;
    }
    return do_prn(h, crops, h->crops_bf16, nullptr, n, n, prn_mode, logits, s, first);
}

int mpn_keypoint_decode(mpn_handle *h, const float *logits, int32_t n, float *scores, float *positions,
                        int32_t *argmax, void *stream)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!logits || !scores || !positions) return fail(h, MPN_ERR_INVALID_ARGUMENT, "pointer is NULL");
    if (n < 0) return fail(h, MPN_ERR_INVALID_ARGUMENT, "n is negative");
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    return launched(h, launch_keypoint_decode(logits, nullptr, n, n, h->cfg.crop_height, h->cfg.crop_width,
                                              h->decode_clusters, scores, positions, argmax, (cudaStream_t)stream), true, "keypoint decode");
}

int mpn_get_keypoints(mpn_handle *h, const float *heatmaps, int32_t hh, int32_t ww, const double box[4],
                      double threshold, int32_t *out, void *stream)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!heatmaps || !box || !out || hh < 1 || ww < 1) return fail(h, MPN_ERR_INVALID_ARGUMENT, "bad arguments");
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    return launched(h, launch_get_keypoints(heatmaps, hh, ww, box[0], box[1], box[2], box[3], threshold, out,
                                            (cudaStream_t)stream), true, "get_keypoints");
}

int mpn_test_exp(mpn_handle *h, const float *x, float *y, int64_t n, void *stream)
{
    if (!h || !x || !y || n < 0) return MPN_ERR_INVALID_ARGUMENT;
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    return launched(h, launch_test_math(x, y, n, 0, (cudaStream_t)stream), true, "test exp");
}

int mpn_test_sigmoid(mpn_handle *h, const float *x, float *y, int64_t n, void *stream)
{
    if (!h || !x || !y || n < 0) return MPN_ERR_INVALID_ARGUMENT;
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    return launched(h, launch_test_math(x, y, n, 1, (cudaStream_t)stream), true, "test sigmoid");
}

// Enqueues every kernel of the path.  With `fork` the two independent front halves run concurrently: heatmap activation
// -> normalisation (thousands of short CTAs) on `s`, candidates -> sort/NMS -> person list (a few latency-bound CTAs) on
// the handle's HIGH-PRIORITY auxiliary stream so that its CTAs are placed as soon as heatmap CTAs retire instead of
// queueing behind the whole heatmap grid; joined before the crops.  The same code is what gets captured into a CUDA graph
// (kernel nodes inherit the stream priority).
static int enqueue_path(mpn_handle *h, const mpn_inputs *in, const mpn_params *p, const mpn_outputs *out, cudaStream_t s,
                        bool fork)
{
    cudaStream_t sd = fork ? h->aux_stream : s;     // detect branch
    cudaStream_t sa = s;                            // heatmap branch
    // programmatic dependent launch along each branch (common.cuh); off while per-kernel events are being recorded
    struct PdlScope {
        PdlScope(int on) { g_pdl = on; }
        ~PdlScope() { g_pdl = 0; }
    } pdl_scope(h->use_pdl && !h->prof.on ? 1 : 0);
    if (fork) {
        MPN_CUDA(h, cudaEventRecord(h->ev_fork, s));
        MPN_CUDA(h, cudaStreamWaitEvent(sd, h->ev_fork, 0));
    }
    const unsigned skip = h->debug_skip;
    int rc = MPN_OK;
    // Two ways to the normalised taps: normalise the WHOLE map once into a padded workspace and crop from that (one
    // division per heatmap value, 16-byte tap loads), or crop from keypoint_heatmaps and normalise every tap (4 divisions
    // per crop sample, scalar tap loads, but no pass over the map).  Large maps with few persons take the second way
    // (1024 x 1024 x 64 with <= 25 boxes per image: 667 -> 507 us per call), everything else the first (crowded 640 x 640
    // x 32: 887 against 1134 us; one 512 x 512 image: 58.3 against 62.4 us); so do crop sizes the padded kernel does
    // not cover.
    const int hh = in->height / h->cfg.downsample, ww = in->width / h->cfg.downsample;
    const long long map_pixels = (long long)in->batch * hh * ww;
    const bool per_tap = map_pixels > 1500LL * in->batch * p->max_detections;
    const bool padded = crop_padded_supported(h->cfg.crop_height, h->cfg.crop_width) && !per_tap;
    float *kh = out->keypoint_heatmaps ? out->keypoint_heatmaps : h->kh_ws;
    // 1a. padded path, pass 1 of the heatmap stage: min / max of the logits (create_pb.py:90,92 through the monotone
    //     sigmoid).  A short HBM stream with no dependency on anything: it runs BESIDE the candidate scan.
    int n_chunks = 0;
    if (padded && !(skip & 2u))
        rc = launched(h, launch_logit_minmax(in->heatmap_logits, in->batch, hh, ww, h->hm_partial2, h->hm_partial_chunks,
                                             h->waves.minmax, &n_chunks, sa), true, "logit min/max");
    if (rc) return rc;
    // 1. scores, threshold, decode, NMS, person list        (retinanet.py:56-81, nms.py:6-61, create_pb.py:96-103)
    // The big heatmap grid (thousands of CTAs) is released only once the candidate scan has finished, i.e. at the moment
    // sort/NMS becomes ready as well: its few big CTAs (512 threads, 100 KB) are placed first on an empty GPU and the
    // heatmap CTAs fill the other SMs.  Released together with the candidate scan they would occupy every SM for their
    // whole life and sort/NMS would start only when they drain (measured: no overlap at all).
    if (!(skip & 1u))
        rc = do_detect(h, in, p, out->boxes, out->scores, out->num_boxes, nullptr, nullptr, sd, !(padded && !(skip & 2u)),
                       fork ? h->ev_cand : nullptr);
    if (rc) return rc;
    if (fork) {
        MPN_CUDA(h, cudaEventRecord(h->ev_join, sd));
        if (!(skip & 1u)) MPN_CUDA(h, cudaStreamWaitEvent(sa, h->ev_cand, 0));
    }
    // 2. heatmap activation / split (create_pb.py:73-76) and normalisation (:90-94): pass 2 of the padded path writes
    //    keypoint_heatmaps, segmentation_masks and the padded normalised map at once; the per-tap path needs the
    //    activations and their min / max only (one pass).
    if (!(skip & 2u))
        rc = padded ? launched(h, launch_heatmap_norm(in->heatmap_logits, in->batch, hh, ww, kh, out->segmentation_masks,
                                                      h->hm_partial2, n_chunks, h->minmax_ws, h->nh_ws, nullptr, h->waves.norm,
                                                      sa), false, "heatmaps + normalise")
                    : launched(h, launch_heatmaps(in->heatmap_logits, in->batch, hh, ww, kh, out->segmentation_masks,
                                                  h->minmax_ws, nullptr, h->hm_partial, h->hm_counter, h->waves.one_pass, sa), false, "heatmaps");
    if (rc) return rc;
    if (fork) MPN_CUDA(h, cudaStreamWaitEvent(s, h->ev_join, 0));
    // 3. crop_and_resize                                     (create_pb.py:106-109)
    const int n_max = in->batch * p->max_detections;
    const int *n_dev = h->person_offsets + in->batch;
    const bool bf16 = p->prn_mode == MPN_PRN_BF16;
    // every crop CTA derives its person from num_boxes / boxes (create_pb.py:96-103); one CTA writes the flat list
    PersonList pl;
    memset(&pl, 0, sizeof(pl));
    pl.num_boxes = out->num_boxes; pl.det_boxes = out->boxes; pl.B = in->batch; pl.max_det = p->max_detections;
    pl.person_box = h->person_box; pl.person_img = h->person_img; pl.person_offsets = h->person_offsets;
    pl.person_offsets_out = out->person_offsets;
    // Calls that fit the single-kernel PRN (capacity <= 256 persons, padded map, in place) can leave the crop to it: its
    // epilogue warps sample while the weights stream (prn_fused.cu: CropFuse), and the crop kernel is not launched.
    // Skip bit 64 (development aid): that kernel samples and stops, so that the crops can be fetched.
    const bool fuse = h->fuse_crop && padded && bf16 && n_max <= kPrnFusedMaxRows && prn_fused_can_crop(h, in->batch) &&
                      !(skip & (8u | 16u));
    if (!(skip & 8u) && !fuse)
        rc = launched(h, padded ? launch_crop_padded(h->nh_ws, hh, ww, pl, n_max, h->cfg.crop_height, h->cfg.crop_width,
                                                     h->crops_f32, bf16 ? h->crops_bf16 : nullptr, s)
                                : launch_crop(kh, h->minmax_ws, hh, ww, pl, n_max, h->cfg.crop_height, h->cfg.crop_width,
                                              h->crops_f32, bf16 ? h->crops_bf16 : nullptr, s),
                      false, "crop");
    if (rc) return rc;
    // 4. PRN                                                 (detector/prn.py:5-25)
    // bf16 mode: in place -- the logits overwrite the fp32 crops (the large-batch fc2 then adds the residual in L2 with a
    // TMA reduce-add instead of loading it; the single-kernel PRN loads and stores every element in the same thread)
    float *prn_out = (bf16 && (h->big || (h->fused && n_max <= kPrnFusedMaxRows))) ? h->crops_f32 : h->logits;
    h->last_run.batch = in->batch; h->last_run.hh = hh; h->last_run.ww = ww; h->last_run.n_max = n_max;
    h->last_run.padded = padded; h->last_run.prn_out = prn_out; h->last_run.valid = true;
    FusedCropCall fc;
    if (fuse) {
        fc.nh = h->nh_ws; fc.hh = hh; fc.ww = ww; fc.pl = pl; fc.only = (skip & 64u) ? 1 : 0;
    }
    if (!(skip & 16u))
        rc = do_prn(h, h->crops_f32, h->crops_bf16, n_dev, 0, n_max, p->prn_mode, prn_out, s, false, fuse ? &fc : nullptr);
    if (rc) return rc;
    // 5. softmax / argmax                                    (create_pb.py:115-142)
    if (skip & 32u) return MPN_OK;
    return launched(h, launch_keypoint_decode(prn_out, n_dev, 0, n_max, h->cfg.crop_height, h->cfg.crop_width,
                                              h->decode_clusters, out->keypoint_scores, out->keypoint_positions, nullptr, s), false,
                    "keypoint decode");
}

// A call is fully described by its pointers, shapes and parameters: the kernel sequence for one such description is
// captured once into a CUDA graph and replayed afterwards (one driver call instead of eight launches + four event ops).
static void make_graph_key(const mpn_handle *h, const mpn_inputs *in, const mpn_params *p, const mpn_outputs *out,
                           GraphKey *k)
{
    memset(k, 0, sizeof(*k));
    k->v[0] = (uint64_t)in->batch; k->v[1] = (uint64_t)in->height; k->v[2] = (uint64_t)in->width;
    k->v[3] = (uint64_t)(uintptr_t)in->class_logits;
    // host calls read the box codes through a per-slot device word: the graph does not depend on the caller's buffer
    k->v[4] = h->enc_indirect ? (uint64_t)(uintptr_t)h->enc_indirect : (uint64_t)(uintptr_t)in->encoded_boxes;
    k->v[5] = (uint64_t)(uintptr_t)in->heatmap_logits;
    const bool levels = !(in->class_logits && in->encoded_boxes) && in->level_class && in->level_boxes;
    for (int i = 0; i < h->cfg.num_levels && levels; ++i) {
        k->v[6 + i] = (uint64_t)(uintptr_t)in->level_class[i];
        k->v[6 + kMaxLevels + i] = (uint64_t)(uintptr_t)in->level_boxes[i];
    }
    int o = 6 + 2 * kMaxLevels;
    uint32_t f[2];
    memcpy(&f[0], &p->score_threshold, 4); memcpy(&f[1], &p->iou_threshold, 4);
    k->v[o++] = ((uint64_t)f[0] << 32) | f[1];
    k->v[o++] = ((uint64_t)(uint32_t)p->max_detections << 32) | (uint32_t)p->prn_mode;
    const void *ptrs[8] = {out->boxes, out->scores, out->num_boxes, out->keypoint_heatmaps, out->segmentation_masks,
                           out->keypoint_scores, out->keypoint_positions, out->person_offsets};
    for (int i = 0; i < 8; ++i) k->v[o++] = (uint64_t)(uintptr_t)ptrs[i];
    k->v[0] |= (uint64_t)h->debug_skip << 32;
}

static int run_graphed(mpn_handle *h, const mpn_inputs *in, const mpn_params *p, const mpn_outputs *out, cudaStream_t s)
{
    GraphKey key;
    make_graph_key(h, in, p, out, &key);
    GraphEntry *hit = nullptr, *victim = &h->graphs[0];
    for (GraphEntry &e : h->graphs) {
        if (e.exec && memcmp(&e.key, &key, sizeof(key)) == 0) { hit = &e; break; }
        if (!e.exec) { victim = &e; }
        else if (victim->exec && e.last_used < victim->last_used) victim = &e;
    }
    if (!hit) {
        // Capturing + instantiating costs far more than eight direct launches: a caller that keeps passing new pointers
        // (a whole cache worth of misses in a row) is served by direct launches from then on.
        if (++h->graph_miss_streak > kGraphCache) { h->graphs_disabled = true; return 1; }
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            return 1;      // not capturable (legacy default stream): caller falls back to direct launches
        }
        const int64_t before = h->total_launches;
        const int rc = enqueue_path(h, in, p, out, s, true);
        const cudaError_t ec = cudaStreamEndCapture(s, &graph);
        const int n_launches = (int)(h->total_launches - before);
        h->total_launches = before;
        if (rc != MPN_OK || ec != cudaSuccess || !graph) {
            cudaGetLastError();
            if (graph) cudaGraphDestroy(graph);
            if (rc != MPN_OK) return rc;
            h->graphs_disabled = true;
            return 1;
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ei = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ei != cudaSuccess) { cudaGetLastError(); h->graphs_disabled = true; return 1; }
        if (victim->exec) cudaGraphExecDestroy(victim->exec);
        victim->exec = exec; victim->key = key; victim->launches = n_launches;
        hit = victim;
    }
    else h->graph_miss_streak = 0;
    hit->last_used = ++h->graph_clock;
    MPN_CUDA(h, cudaGraphLaunch(hit->exec, s));
    h->last_launches = hit->launches;
    h->total_launches += hit->launches;
    return MPN_OK;
}

int mpn_run(mpn_handle *h, const mpn_inputs *in, const mpn_params *p, const mpn_outputs *out, void *stream)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!in || !p || !out) return fail(h, MPN_ERR_INVALID_ARGUMENT, "inputs/params/outputs is NULL");
    if (!in->heatmap_logits) return fail(h, MPN_ERR_INVALID_ARGUMENT, "heatmap_logits is NULL");
    if (!out->keypoint_scores || !out->keypoint_positions || !out->boxes || !out->scores || !out->num_boxes)
        return fail(h, MPN_ERR_INVALID_ARGUMENT, "an output pointer is NULL");
    if (!h->have_weights) return fail(h, MPN_ERR_NO_WEIGHTS, "mpn_set_prn_weights has not been called");
    int rc = check_image_size(h, in->batch, in->height, in->width);
    if (rc) return rc;
    rc = check_params(h, p);
    if (rc) return rc;
    if (p->prn_mode == MPN_PRN_BF16 && !h->crops_bf16) return fail(h, MPN_ERR_UNSUPPORTED, "handle was created without the bf16 PRN");
    if (p->prn_mode == MPN_PRN_FP32 && !(h->cfg.prn_modes & 1)) return fail(h, MPN_ERR_UNSUPPORTED, "handle was created without the fp32 PRN");
    const bool flat = in->class_logits != nullptr && in->encoded_boxes != nullptr;
    const bool levels = in->level_class != nullptr && in->level_boxes != nullptr;
    if (!flat && !levels) return fail(h, MPN_ERR_INVALID_ARGUMENT, "need class_logits+encoded_boxes or level_class+level_boxes");
    // per-call pointer preconditions are checked HERE: a graph replay never reaches the checks of the launchers
    if (flat && reinterpret_cast<uintptr_t>(in->encoded_boxes) % 16 != 0)
        return fail(h, MPN_ERR_INVALID_ARGUMENT, "encoded_boxes must be 16-byte aligned");
    if (reinterpret_cast<uintptr_t>(in->heatmap_logits) % 16 != 0)
        return fail(h, MPN_ERR_INVALID_ARGUMENT, "heatmap_logits must be 16-byte aligned");
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    cudaStream_t s = (cudaStream_t)stream;
    if (h->prof.on) {              // per-kernel timing: one stream, direct launches
        ProfScope prof_scope(h, s);
        return enqueue_path(h, in, p, out, s, false);
    }
    // a caller who is capturing this stream into a graph of their own gets the plain launch sequence recorded into it
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (s != nullptr && s != cudaStreamLegacy && cudaStreamIsCapturing(s, &cap) != cudaSuccess) { cudaGetLastError(); cap = cudaStreamCaptureStatusNone; }
    const bool capturable = s != nullptr && s != cudaStreamLegacy && h->cfg_use_graphs && !h->graphs_disabled &&
                            cap == cudaStreamCaptureStatusNone;
    if (capturable) {
        rc = run_graphed(h, in, p, out, s);
        if (rc <= 0) return rc;
    }
    return enqueue_path(h, in, p, out, s, s != nullptr && s != cudaStreamLegacy);
}

// Offsets of boxes, scores, keypoint_scores, keypoint_positions, num_boxes, person_offsets inside a slot's small-output
// block for a call of B images and NP = B * max_detections person rows (16-byte aligned sub-ranges); offs[6] = total.
static void small_layout(size_t B, size_t NP, size_t offs[7])
{
    const size_t bytes[6] = {NP * 16, NP * 4, NP * 68, NP * 136, B * 4, (B + 1) * 4};
    size_t o = 0;
    for (int i = 0; i < 6; ++i) { offs[i] = o; o += (bytes[i] + 15) / 16 * 16; }
    offs[6] = o;
}

static int ensure_staging(mpn_handle *h)
{
    if (h->staging_ready) return MPN_OK;
    const size_t B = h->cfg.max_batch, A = h->max_anchors, P = h->max_hm_pix, NP = h->max_persons;
    MPN_CUDA(h, cudaStreamCreateWithFlags(&h->in_stream, cudaStreamNonBlocking));
    MPN_CUDA(h, cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
    for (int k = 0; k < kHostSlots; ++k) {
        HostSlot &sl = h->slots[k];
        MPN_CUDA(h, dalloc(&sl.cls, B * A));
        MPN_CUDA(h, dalloc(&sl.hml, B * P * 18));
        MPN_CUDA(h, dalloc(&sl.kh, B * P * 17));
        MPN_CUDA(h, dalloc(&sl.seg, B * P));
        // one block for the six small outputs (laid out per call by small_layout)
        size_t offs[7];
        small_layout(B, NP, offs);
        MPN_CUDA(h, dalloc(&sl.small_dev, offs[6]));
        MPN_CUDA(h, cudaHostAlloc(reinterpret_cast<void **>(&sl.small_host), offs[6], cudaHostAllocDefault));
        MPN_CUDA(h, cudaEventCreateWithFlags(&sl.ev_in, cudaEventDisableTiming));
        MPN_CUDA(h, cudaEventCreateWithFlags(&sl.ev_comp, cudaEventDisableTiming));
        MPN_CUDA(h, cudaEventCreateWithFlags(&sl.ev_out, cudaEventDisableTiming));
    }
    h->staging_ready = true;
    return MPN_OK;
}

// Device-visible alias of a host buffer that the GPU can read in place (pinned / registered memory), else NULL.
static const float *mapped_alias(const float *host)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (at.type == cudaMemoryTypeHost && at.devicePointer) return static_cast<const float *>(at.devicePointer);
    return nullptr;
}

// Waits for the slot's copy-out and hands the small outputs from the pinned block to the caller's buffers (once).
static int finish_slot(mpn_handle *h, HostSlot &sl)
{
    if (!sl.used) return MPN_OK;
    MPN_CUDA(h, cudaEventSynchronize(sl.ev_out));
    if (sl.scatter_pending) {
        for (int i = 0; i < sl.n_scatter; ++i)
            if (sl.scatter[i].dst && sl.scatter[i].bytes) memcpy(sl.scatter[i].dst, sl.small_host + sl.scatter[i].off, sl.scatter[i].bytes);
        sl.scatter_pending = false;
    }
    return MPN_OK;
}

int mpn_submit_host(mpn_handle *h, const mpn_inputs *in, const mpn_params *p, const mpn_outputs *out, int64_t *ticket)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!in || !p || !out) return fail(h, MPN_ERR_INVALID_ARGUMENT, "inputs/params/outputs is NULL");
    if (!in->class_logits || !in->encoded_boxes || !in->heatmap_logits)
        return fail(h, MPN_ERR_INVALID_ARGUMENT, "the host path takes the concatenated layout: class_logits, encoded_boxes, heatmap_logits");
    if (!out->boxes || !out->scores || !out->num_boxes || !out->keypoint_scores || !out->keypoint_positions)
        return fail(h, MPN_ERR_INVALID_ARGUMENT, "output pointer is NULL");
    int rc = check_image_size(h, in->batch, in->height, in->width);
    if (rc) return rc;
    rc = check_params(h, p);
    if (rc) return rc;
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    rc = ensure_staging(h);
    if (rc) return rc;
    HostSlot &sl = h->slots[h->next_ticket % kHostSlots];
    rc = finish_slot(h, sl);           // a caller that never waited for this slot's previous ticket still gets its outputs
    if (rc) return rc;
    const size_t B = in->batch, A = count_anchors(h->cfg, in->height, in->width);
    const size_t P = (size_t)(in->height / 4) * (in->width / 4), NP = B * p->max_detections;
    // ---- feed (inference/detector.py:47) on the copy-in stream.  The box codes are only ever gathered for the few
    // hundred confident anchors, so when the caller's buffer is pinned the NMS kernel reads those 16-byte rows in
    // place over PCIe instead of copying all B*A*16 bytes.
    cudaStream_t si = h->in_stream, sc = h->own_stream, so = h->out_stream;
    if (sl.used) MPN_CUDA(h, cudaStreamWaitEvent(si, sl.ev_out, 0));     // slot drained by its previous call
    MPN_CUDA(h, cudaMemcpyAsync(sl.cls, in->class_logits, B * A * 4, cudaMemcpyHostToDevice, si));
    MPN_CUDA(h, cudaMemcpyAsync(sl.hml, in->heatmap_logits, B * P * 72, cudaMemcpyHostToDevice, si));
    // gathered in place only when the device sees the buffer AND its rows are 16-byte aligned (the kernels load a code as
    // one float4; a graph replay does not re-run the alignment check of the capture); anything else is copied
    const float *enc = mapped_alias(in->encoded_boxes);
    if (enc && reinterpret_cast<uintptr_t>(enc) % 16 != 0) enc = nullptr;
    h->last_h2d_bytes = (int64_t)(B * A * 4 + B * P * 72);
    if (!enc) {
        if (!sl.enc) MPN_CUDA(h, dalloc(&sl.enc, (size_t)h->cfg.max_batch * h->max_anchors * 4));
        MPN_CUDA(h, cudaMemcpyAsync(sl.enc, in->encoded_boxes, B * A * 16, cudaMemcpyHostToDevice, si));
        enc = sl.enc;
        h->last_h2d_bytes += (int64_t)(B * A * 16);
    }
    MPN_CUDA(h, cudaEventRecord(sl.ev_in, si));
    // ---- the path on the compute stream.  The kernels find this call's box codes through a per-slot device word, so
    // that ONE captured graph per slot serves every buffer the caller passes (a graph keyed on the caller's pointer was
    // re-captured for every new pinned buffer: 0.28 instead of 0.07 ms per single-image call over a ring of 91 buffers).
    // The word is written on the compute stream BEFORE it waits for the copies: the copy-in stream carries nothing but
    // copies (a kernel between them cost 5 % of the PCIe-bound throughput).
    if (!sl.enc_word) MPN_CUDA(h, dalloc(&sl.enc_word, 1));
    if (launch_set_pointer(sl.enc_word, enc, sc) < 0 || cudaPeekAtLastError() != cudaSuccess)
        return fail(h, MPN_ERR_CUDA, "host path: %s", cudaGetErrorString(cudaGetLastError()));
    MPN_CUDA(h, cudaStreamWaitEvent(sc, sl.ev_in, 0));
    mpn_inputs din = *in;
    din.class_logits = sl.cls; din.encoded_boxes = enc; din.heatmap_logits = sl.hml;
    din.level_class = nullptr; din.level_boxes = nullptr;
    size_t offs[7];
    small_layout(B, NP, offs);
    sl.small_bytes = offs[6];
    sl.boxes = reinterpret_cast<float *>(sl.small_dev + offs[0]);
    sl.scores = reinterpret_cast<float *>(sl.small_dev + offs[1]);
    sl.kscores = reinterpret_cast<float *>(sl.small_dev + offs[2]);
    sl.kpos = reinterpret_cast<float *>(sl.small_dev + offs[3]);
    sl.num = reinterpret_cast<int *>(sl.small_dev + offs[4]);
    sl.offsets = reinterpret_cast<int *>(sl.small_dev + offs[5]);
    mpn_outputs dout;
    dout.boxes = sl.boxes; dout.scores = sl.scores; dout.num_boxes = sl.num;
    dout.keypoint_heatmaps = sl.kh; dout.segmentation_masks = out->segmentation_masks ? sl.seg : nullptr;
    dout.keypoint_scores = sl.kscores; dout.keypoint_positions = sl.kpos; dout.person_offsets = sl.offsets;
    h->enc_indirect = sl.enc_word;
    rc = mpn_run(h, &din, p, &dout, sc);
    h->enc_indirect = nullptr;
    if (rc) return rc;
    MPN_CUDA(h, cudaEventRecord(sl.ev_comp, sc));
    // ---- fetch (inference/detector.py:48) on the copy-out stream
    MPN_CUDA(h, cudaStreamWaitEvent(so, sl.ev_comp, 0));
    MPN_CUDA(h, cudaMemcpyAsync(sl.small_host, sl.small_dev, sl.small_bytes, cudaMemcpyDeviceToHost, so));
    h->last_d2h_bytes = (int64_t)sl.small_bytes;
    {
        const unsigned char *base = sl.small_dev;
        auto rec = [&](int i, void *dst, const void *dev, size_t bytes) {
            sl.scatter[i].dst = dst;
            sl.scatter[i].off = (size_t)(static_cast<const unsigned char *>(dev) - base);
            sl.scatter[i].bytes = bytes;
        };
        rec(0, out->boxes, sl.boxes, NP * 16);
        rec(1, out->scores, sl.scores, NP * 4);
        rec(2, out->keypoint_scores, sl.kscores, NP * 68);
        rec(3, out->keypoint_positions, sl.kpos, NP * 136);
        rec(4, out->num_boxes, sl.num, B * 4);
        rec(5, out->person_offsets, sl.offsets, out->person_offsets ? (B + 1) * 4 : 0);
        sl.n_scatter = 6;
        sl.scatter_pending = true;
    }
    if (out->keypoint_heatmaps) {
        MPN_CUDA(h, cudaMemcpyAsync(out->keypoint_heatmaps, sl.kh, B * P * 68, cudaMemcpyDeviceToHost, so));
        h->last_d2h_bytes += (int64_t)(B * P * 68);
    }
    if (out->segmentation_masks) {
        MPN_CUDA(h, cudaMemcpyAsync(out->segmentation_masks, sl.seg, B * P * 4, cudaMemcpyDeviceToHost, so));
        h->last_d2h_bytes += (int64_t)(B * P * 4);
    }
    MPN_CUDA(h, cudaEventRecord(sl.ev_out, so));
    sl.used = true;
    if (ticket) *ticket = h->next_ticket;
    ++h->next_ticket;
    return MPN_OK;
}

int mpn_wait(mpn_handle *h, int64_t ticket)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (ticket < 0 || ticket >= h->next_ticket) return fail(h, MPN_ERR_INVALID_ARGUMENT, "unknown ticket %lld", (long long)ticket);
    // if the slot has been re-used since, this waits for the later call, which is ordered after this one
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    return finish_slot(h, h->slots[ticket % kHostSlots]);
}

int mpn_run_host(mpn_handle *h, const mpn_inputs *in, const mpn_params *p, const mpn_outputs *out)
{
    return mpn_submit_host(h, in, p, out, nullptr);
}

int mpn_synchronize(mpn_handle *h)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    if (h->staging_ready) MPN_CUDA(h, cudaStreamSynchronize(h->in_stream));
    MPN_CUDA(h, cudaStreamSynchronize(h->own_stream));
    if (h->staging_ready) {
        MPN_CUDA(h, cudaStreamSynchronize(h->out_stream));
        for (HostSlot &sl : h->slots) {
            const int rc = finish_slot(h, sl);
            if (rc) return rc;
        }
    }
    return MPN_OK;
}

int mpn_host_traffic(const mpn_handle *h, int64_t *h2d_bytes, int64_t *d2h_bytes)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (h2d_bytes) *h2d_bytes = h->last_h2d_bytes;
    if (d2h_bytes) *d2h_bytes = h->last_d2h_bytes;
    return MPN_OK;
}

int mpn_debug_fused_trace(mpn_handle *h, int32_t enable, uint64_t *host_out, int32_t capacity, int32_t *grid_out)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    MPN_CUDA(h, cudaDeviceSynchronize());
    int g = 0;
    // the trace pointer is a kernel argument baked into the captured graphs: drop them whenever tracing is switched
    for (GraphEntry &e : h->graphs)
        if (e.exec) { cudaGraphExecDestroy(e.exec); e.exec = nullptr; }
    h->graph_miss_streak = 0;
    const int rc = prn_fused_trace(h, enable, reinterpret_cast<unsigned long long *>(host_out), capacity, &g);
    if (grid_out) *grid_out = g;
    return rc == MPN_OK ? MPN_OK : fail(h, rc, "fused PRN trace unavailable");
}

int mpn_debug_fetch(mpn_handle *h, int32_t what, void *dst, int64_t capacity_bytes, int64_t *bytes_out)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!h->last_run.valid) return fail(h, MPN_ERR_INVALID_ARGUMENT, "no mpn_run has been issued on this handle");
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    MPN_CUDA(h, cudaDeviceSynchronize());
    const size_t B = h->last_run.batch, pix = (size_t)h->last_run.hh * h->last_run.ww, n = h->last_run.n_max, D = h->D;
    const void *src = nullptr;
    size_t bytes = 0;
    switch (what) {
    case MPN_DEBUG_NORMALISED:
        if (!h->last_run.padded) return fail(h, MPN_ERR_UNSUPPORTED, "the last run took the per-tap crop path: no normalised map");
        src = h->nh_ws; bytes = B * pix * 20 * 4; break;
    case MPN_DEBUG_CROPS_F32: src = h->crops_f32; bytes = n * D * 4; break;
    case MPN_DEBUG_CROPS_BF16:
        if (!h->crops_bf16) return fail(h, MPN_ERR_UNSUPPORTED, "handle was created without the bf16 PRN");
        src = h->crops_bf16; bytes = n * D * 2; break;
    case MPN_DEBUG_LOGITS: src = h->last_run.prn_out; bytes = n * D * 4; break;
    case MPN_DEBUG_MINMAX: src = h->minmax_ws; bytes = B * 17 * 2 * 4; break;
    case MPN_DEBUG_PERSON_BOX: src = h->person_box; bytes = n * 16; break;
    case MPN_DEBUG_PERSON_IMAGE: src = h->person_img; bytes = n * 4; break;
    default: return fail(h, MPN_ERR_INVALID_ARGUMENT, "unknown buffer %d", what);
    }
    if (bytes_out) *bytes_out = (int64_t)bytes;
    if (!dst) return MPN_OK;                           // size query
    if (capacity_bytes < 0) return fail(h, MPN_ERR_INVALID_ARGUMENT, "negative capacity");
    if ((size_t)capacity_bytes < bytes) bytes = (size_t)capacity_bytes;
    MPN_CUDA(h, cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    return MPN_OK;
}

int mpn_test_sigmoid_monotone(mpn_handle *h, uint32_t key_begin, uint64_t count, uint64_t *violations, void *stream)
{
    if (!h || !violations) return MPN_ERR_INVALID_ARGUMENT;
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    return launched(h, launch_test_monotone(key_begin, count, reinterpret_cast<unsigned long long *>(violations),
                                            (cudaStream_t)stream), true, "test sigmoid monotone");
}

int mpn_debug_nms_trace(mpn_handle *h, int32_t enable, uint64_t *host_out, int32_t capacity)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    MPN_CUDA(h, cudaDeviceSynchronize());
    // the trace pointer is a kernel argument baked into the captured graphs: drop them whenever tracing is switched
    for (GraphEntry &e : h->graphs)
        if (e.exec) { cudaGraphExecDestroy(e.exec); e.exec = nullptr; }
    h->graph_miss_streak = 0;
    const size_t n = (size_t)h->cfg.max_batch * 16;
    if (enable && !h->nms_trace) {
        MPN_CUDA(h, cudaMalloc(reinterpret_cast<void **>(&h->nms_trace), n * sizeof(unsigned long long)));
        MPN_CUDA(h, cudaMemset(h->nms_trace, 0, n * sizeof(unsigned long long)));
    }
    if (host_out && h->nms_trace) {
        const size_t m = (size_t)capacity < n ? (size_t)capacity : n;
        MPN_CUDA(h, cudaMemcpy(host_out, h->nms_trace, m * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    if (!enable && h->nms_trace) { cudaFree(h->nms_trace); h->nms_trace = nullptr; }
    return MPN_OK;
}

int mpn_debug_skip(mpn_handle *h, uint32_t mask)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    h->debug_skip = mask;
    return MPN_OK;
}

int mpn_set_profiling(mpn_handle *h, int32_t enable)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    if (enable && !h->prof_events_ready) {
        for (int i = 0; i <= kMaxMarks; ++i) MPN_CUDA(h, cudaEventCreate(&h->prof.ev[i]));
        h->prof_events_ready = true;
    }
    h->prof.on = enable != 0;
    h->prof.n = 0;
    return MPN_OK;
}

int mpn_get_profile(mpn_handle *h, int32_t capacity, const char **names, float *ms, int32_t *count)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (!names || !ms || !count || capacity < 0) return fail(h, MPN_ERR_INVALID_ARGUMENT, "bad arguments");
    if (!h->prof.on) return fail(h, MPN_ERR_INVALID_ARGUMENT, "profiling is off (mpn_set_profiling)");
    MPN_CUDA(h, cudaSetDevice(h->cfg.device));
    const int n = h->prof.n < capacity ? h->prof.n : capacity;
    if (h->prof.n > 0) MPN_CUDA(h, cudaEventSynchronize(h->prof.ev[h->prof.n]));
    for (int i = 0; i < n; ++i) {
        names[i] = h->prof.name[i];
        MPN_CUDA(h, cudaEventElapsedTime(&ms[i], h->prof.ev[i], h->prof.ev[i + 1]));
    }
    *count = n;
    return MPN_OK;
}

int mpn_launch_count(const mpn_handle *h, int64_t *last_call, int64_t *total)
{
    if (!h) return MPN_ERR_INVALID_ARGUMENT;
    if (last_call) *last_call = h->last_launches;
    if (total) *total = h->total_launches;
    return MPN_OK;
}

}  // extern "C"
