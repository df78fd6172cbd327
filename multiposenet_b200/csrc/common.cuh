// Internal declarations shared by the kernels and the C ABI (api.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <string.h>

#include "../../include/mpn_b200.h"

namespace mpn {

constexpr int kMaxLevels = MPN_MAX_LEVELS;
constexpr int kMaxShapes = MPN_MAX_ANCHOR_SHAPES;
constexpr int kSortSmemCap = 8192;      // candidates per image sorted in shared memory; above: global scratch
constexpr int kMaxDetCap = 1024;        // kept boxes staged in shared memory
constexpr int kPrnFusedMaxRows = 256;   // persons per call handled by the single-kernel PRN (prn_fused.cu)
constexpr int kPrnSplit3GroupRows = 80; // persons per launch of the fp32-accurate tensor-core PRN (prn_split3.cu)
constexpr int kPrnSplit3MaxRows = 240;  //   ... and per call (three launches); more: the SIMT kernels

// Everything the kernels need to rebuild anchor `a` from its index (detector/anchor_generator.py:53-116)
// without an anchor tensor in HBM.  Passed by value (~1.3 KB of kernel parameters).
struct AnchorTable {
    int n_levels, n_loc, num_anchors;
    int gh[kMaxLevels], gw[kMaxLevels];
    int off[kMaxLevels + 1];              // first anchor index of each level
    float stride[kMaxLevels], oy[kMaxLevels], ox[kMaxLevels];
    float half_h[kMaxLevels][kMaxShapes]; // 0.5 * scale / sqrt(ratio)
    float half_w[kMaxLevels][kMaxShapes]; // 0.5 * scale * sqrt(ratio)
    float fH, fW;
    float sf[4];                          // SCALE_FACTORS (detector/constants.py:19)
};

// Per-level NCHW head outputs (detector/box_predictor.py:53-90 fused away); NULL => concatenated layout.
struct LevelPtrs {
    const float *cls[kMaxLevels];
    const float *box[kMaxLevels];
};

struct DetectArgs {
    const float *cls;        // [B, A] or NULL
    const float *enc;        // [B, A, 4] or NULL
    const float *const *enc_ind;   // host path: device word holding the box-code pointer of this call (NULL: use enc), so
                                   // that one captured graph serves every pinned buffer the caller passes
    LevelPtrs lv;            // used when cls == NULL
    int B;
    float thr, iou_thr, pre_thr;
    int max_det;
    // workspace
    unsigned long long *cand_keys;   // [B, key_cap]
    int key_cap;                     // power of two >= A
    int *cand_count;                 // [B]
    // outputs
    float *boxes;            // [B, max_det, 4]
    float *scores;           // [B, max_det]
    int *num_boxes;          // [B]
    int *sel_anchor;         // [B, max_det] or NULL
    int *n_candidates;       // [B] or NULL
    unsigned long long *trace;   // optional [B, 16] globaltimer stamps of the sort / NMS phases (mpn_debug_nms_trace)
};

// How a crop CTA finds the person of its row n (create_pb.py:96-103: boxes = concat_i predicted_boxes[i][:num_boxes[i]],
// box_ind likewise).  Either an explicit flat list (the stage entry points), or -- inside mpn_run -- derived by every CTA
// from the detection outputs themselves: row n is box k of image b where the exclusive scan of num_boxes brackets n.  The
// sort / NMS kernel therefore ends when its last image is resolved; an earlier version had its last CTA build the flat
// list behind a fence and an atomic counter, a tail of ~4.5 us on the critical chain of every call.  One extra CTA of the
// crop grid (blockIdx.x == 0 of the last blockIdx.y) writes the list and the offsets for whoever wants them afterwards
// (person count for the PRN / decode kernels, person_offsets output, mpn_debug_fetch).
struct PersonList {
    const float *boxes;      // explicit: [N, 4]
    const int *box_ind;      // explicit: [N]
    const int *n_dev;        // explicit: device person count (NULL: n_host)
    int n_host;
    const int *num_boxes;    // derived: [B] (NULL: explicit list)
    const float *det_boxes;  // derived: [B, max_det, 4]
    int B, max_det;
    float *person_box;       // outputs of the list CTA (derived mode): [B * max_det, 4], [B * max_det], [B + 1], user copy or NULL
    int *person_img;
    int *person_offsets;
    int *person_offsets_out;
};

// Optional per-kernel timing (mpn_set_profiling): every launcher marks the stream before each kernel it launches.
constexpr int kMaxMarks = 24;
struct Profiler {
    bool on;
    int n;
    cudaEvent_t ev[kMaxMarks + 1];
    const char *name[kMaxMarks];
};
extern thread_local Profiler *g_prof;
inline void prof_mark(cudaStream_t s, const char *name)
{
    Profiler *p = g_prof;
    if (p && p->on && p->n < kMaxMarks) {
        cudaEventRecord(p->ev[p->n], s);
        p->name[p->n++] = name;
    }
}

// ---- programmatic dependent launch -----------------------------------------------------------------------------------
// Every kernel of the path starts with pdl_trigger() + pdl_wait() (prn_fused_kernel waits later, after its prologue and
// its first weight loads): launched with the programmatic-serialization attribute, the CTAs of kernel N+1 are placed and
// run their prologue while the last wave of kernel N drains, and block in griddepcontrol.wait until kernel N has
// completed and flushed.  Because EVERY kernel waits before it exits, completion is transitive along a stream.  Without
// the attribute both instructions are no-ops.  g_pdl is set by api.cu around the enqueue of one call.
extern thread_local int g_pdl;
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool pdl,
                            Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && g_pdl) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
#endif

// ---- launchers (each returns the number of kernels it launched, or a negative cudaError) ----
int detect_prepare();   // once per device (handle creation)
int launch_anchors(const AnchorTable &t, float *out, cudaStream_t s);
int launch_set_pointer(const float **slot, const float *value, cudaStream_t s);     // *slot = value, stream-ordered
// after_candidates (optional): event recorded on `s` between the candidate scan and the sort / NMS kernel
int launch_detect(const AnchorTable &t, const DetectArgs &a, cudaStream_t s, cudaEvent_t after_candidates = nullptr);

struct HeatmapWaves { int one_pass, minmax, norm; };   // resident CTA slots of the three heatmap-sized streaming kernels
int heatmap_prepare(HeatmapWaves *w);                 // once per handle
int launch_heatmaps(const float *hml, int B, int hh, int ww, float *kh, float *seg, float *minmax_ws,
                    float *minmax_out, int *partial_ws, unsigned int *counter_ws, int slots, cudaStream_t s);
int launch_heatmap_head(const float *x, const float *w, const float *bias, int B, int hh, int ww, float *logits, float *kh,
                        float *seg, float *minmax_ws, float *minmax_out, int *partial_ws, unsigned int *counter_ws,
                        cudaStream_t s);
// the two-pass form (padded crop path): min / max from the logits, then activation + normalisation in one pass
int launch_logit_minmax(const float *hml, int B, int hh, int ww, int *partial_ws, int partial_chunks_cap, int slots,
                        int *n_chunks, cudaStream_t s);
int launch_heatmap_norm(const float *hml, int B, int hh, int ww, float *kh, float *seg, const int *partial_ws, int n_chunks,
                        float *minmax_ws, float *nh, float *minmax_out, int slots, cudaStream_t s);
// crop of the padded (20 floats / pixel) normalised map written by launch_heatmap_norm
bool crop_padded_supported(int crop_h, int crop_w);
int launch_crop_padded(const float *nh, int hh, int ww, const PersonList &pl, int n_max, int crop_h, int crop_w,
                       float *crops_f32, __nv_bfloat16 *crops_bf16, cudaStream_t s);
int launch_crop(const float *kh, const float *minmax, int hh, int ww, const PersonList &pl, int n_max, int crop_h, int crop_w,
                float *crops_f32, __nv_bfloat16 *crops_bf16, cudaStream_t s);
int launch_get_keypoints(const float *hm, int hh, int ww, double ymin, double xmin, double ymax, double xmax,
                         double threshold, int *out, cudaStream_t s);

// once per handle before the first decode: bisects the exp == 1 threshold, reports how many clusters are resident at once;
// -2: crop size not covered by the decode kernel
int kpdecode_prepare(cudaStream_t s, int crop_h, int crop_w, int *resident_clusters);
int launch_keypoint_decode(const float *logits, const int *n_dev, int n_host, int n_max, int crop_h, int crop_w,
                           int resident_clusters, float *scores, float *positions, int *argmax, cudaStream_t s);

struct PrnWeights {
    int D, hidden;
    const float *W1, *b1, *W2, *b2;                 // fp32, [in,out] row-major
    const __nv_bfloat16 *W1t, *W2t;                 // bf16, transposed to [out,in] (K-major B operands)
};
struct PrnWorkspace {
    float *partial;          // split-K partial sums of fc1
    size_t partial_floats;
    float *y1;               // [n_max, hidden] fp32
    __nv_bfloat16 *y1_bf16;  // [n_max_pad, hidden]
    int n_max;
};
// skip_le: the kernels exit at once when the person count is <= skip_le (prn_split3.cu handles those)
int launch_prn_fp32(const PrnWeights &w, const PrnWorkspace &ws, const float *x, const int *n_dev, int n_host,
                    int n_max, float *logits, int skip_le, cudaStream_t s);
// skip_le: the launched kernels exit at once when the person count is <= skip_le (the fused kernel handles those)
int launch_prn_bf16(const PrnWeights &w, const PrnWorkspace &ws, const float *x_f32, const __nv_bfloat16 *x_bf16,
                    const int *n_dev, int n_host, int n_max, float *logits, void *tmaps, int skip_le, cudaStream_t s);
int launch_fc1_reduce(const float *partial, int splits, size_t split_stride, const float *bias, int hidden,
                      const int *m_dev, int m_host, int m_max, float *y1, __nv_bfloat16 *y1_bf16, int skip_le,
                      cudaStream_t s);
int launch_f32_to_bf16(const float *x, __nv_bfloat16 *y, const int *n_rows_dev, int n_rows_host, int row_len,
                       int n_rows_max, cudaStream_t s);
int launch_transpose_to_bf16(const float *w, int rows, int cols, __nv_bfloat16 *wt, cudaStream_t s);

int launch_test_math(const float *x, float *y, int64_t n, int which, cudaStream_t s);
int launch_test_monotone(unsigned key_begin, unsigned long long count, unsigned long long *violations, cudaStream_t s);

}  // namespace mpn
