// The handle behind the C ABI (opaque to callers).
#pragma once
#include "common.cuh"

namespace mpn {
constexpr int kHostSlots = 3;
// Device staging of one in-flight host call.
struct HostSlot {
    float *cls, *enc, *hml, *kh, *seg, *boxes, *scores, *kscores, *kpos;
    const float **enc_word;      // device word: where this call's box codes are (slot copy or the caller's pinned buffer)
    int *num, *offsets;
    // The six small outputs (boxes, scores, keypoint_scores, keypoint_positions, num_boxes, person_offsets) are sub-ranges
    // of ONE device block and come back in ONE copy into a pinned block of the same layout; mpn_wait hands them to the
    // caller's buffers (a few tens of KB of host memcpy) -- three device-to-host copies per call instead of eight.
    unsigned char *small_dev, *small_host;
    size_t small_bytes;
    struct { void *dst; size_t off, bytes; } scatter[6];
    int n_scatter;
    bool scatter_pending;
    cudaEvent_t ev_in, ev_comp, ev_out;
    bool used;
};
}  // namespace mpn

namespace mpn {
constexpr int kGraphCache = 160;   // distinct (pointers, shapes, parameters) descriptions kept as instantiated graphs
struct GraphKey { uint64_t v[6 + 2 * kMaxLevels + 2 + 8]; };
struct GraphEntry {
    GraphKey key;
    cudaGraphExec_t exec;
    int launches;
    uint64_t last_used;
};
}  // namespace mpn

struct mpn_handle {
    mpn_config cfg;
    int D;                  // crop_h * crop_w * num_keypoints
    int max_anchors, key_cap, max_hm_pix, max_persons;
    char err[512];
    cudaStream_t own_stream, aux_stream;
    cudaEvent_t own_event, ev_fork, ev_join, ev_cand;
    // CUDA graphs of mpn_run, keyed by the call's pointers / shapes / parameters
    mpn::GraphEntry graphs[mpn::kGraphCache];
    uint64_t graph_clock;
    int graph_miss_streak;  // consecutive calls that had to capture: callers that never repeat a description get direct launches
    bool cfg_use_graphs, graphs_disabled;
    const float *const *enc_indirect;   // set around the mpn_run of a host call (see DetectArgs::enc_ind)
    bool use_pdl;           // programmatic dependent launch between consecutive kernels of a branch
    unsigned debug_skip;    // mpn_debug_skip: bit i set = stage i is not launched (timing experiments only)
    bool fuse_crop;         // crop_and_resize inside the single-kernel PRN where the call allows it (MPN_FUSE_CROP)
    // detect workspace
    unsigned long long *cand_keys;
    int *cand_count;
    float *person_box;
    int *person_img;
    int *person_offsets;
    unsigned long long *nms_trace;   // mpn_debug_nms_trace (NULL: off)
    // heatmap workspace
    float *kh_ws;
    float *nh_ws;           // normalised heatmaps (create_pb.py:93-94), never returned
    float *minmax_ws;
    int *hm_partial;        // [B, chunks, 17, 2] per-CTA (min, max) of the heatmap kernels
    int *hm_partial2;       // the same for the logit min / max pass of the two-pass form (ordered keys): a buffer of its own,
                            // because that pass has no dependency on the previous call's kernels of the other branch
    int hm_partial_chunks;  // chunks per image either array holds
    mpn::HeatmapWaves waves;
    unsigned int *hm_counter;   // [B] CTAs finished per image (self re-arming)
    // PRN workspace / weights
    float *crops_f32, *logits;
    __nv_bfloat16 *crops_bf16;
    float *W1, *b1, *W2, *b2;
    __nv_bfloat16 *W1t, *W2t;
    mpn::PrnWorkspace prn_ws;
    void *tmaps;            // opaque: prn_tcgen05.cu
    void *fused;            // opaque: prn_fused.cu (NULL when the shape is not covered)
    void *split3;           // opaque: prn_split3.cu (fp32 mode on the tensor cores, <= 80 persons; NULL: SIMT kernels only)
    void *big;              // opaque: prn_big.cu (NULL when the shape is not covered or the capacity is <= 256 persons)
    bool have_weights;
    int decode_clusters;    // keypoint decode: clusters resident at once (kpdecode_prepare)
    // host path (mpn_submit_host): kHostSlots calls in flight, copy-in / compute / copy-out on three streams
    mpn::HostSlot slots[mpn::kHostSlots];
    cudaStream_t in_stream, out_stream;
    bool staging_ready;
    int64_t next_ticket;
    int64_t last_h2d_bytes, last_d2h_bytes;
    int64_t last_launches, total_launches;
    mpn::Profiler prof;
    bool prof_events_ready;
    // what the most recent mpn_run left in the workspace (mpn_debug_fetch)
    struct {
        bool valid, padded;
        int batch, hh, ww, n_max;
        const float *prn_out;
    } last_run;
};

namespace mpn {
int prn_bf16_prepare(mpn_handle *h);   // prn_tcgen05.cu: TMA tensor maps for the bf16 GEMMs
void prn_bf16_release(mpn_handle *h);
int prn_big_prepare(mpn_handle *h);    // prn_big.cu: persistent tcgen05 GEMMs for > kPrnFusedMaxRows persons
void prn_big_release(mpn_handle *h);
int launch_prn_big(mpn_handle *h, const float *x_f32, const int *n_dev, int n_host, int n_max, float *logits, int skip_le,
                   cudaStream_t s);
int prn_split3_prepare(mpn_handle *h); // prn_split3.cu: fp32-accurate PRN on the tensor cores (three bf16 parts per number)
void prn_split3_release(mpn_handle *h);
int prn_split3_set_weights(mpn_handle *h, const float *dW1, const float *dW2, cudaStream_t s);
int launch_prn_split3(mpn_handle *h, const float *x_f32, const int *n_dev, int n_host, int n_max, float *logits, cudaStream_t s);
int prn_fused_prepare(mpn_handle *h);  // prn_fused.cu: persistent single-kernel PRN for <= kPrnFusedMaxRows persons
void prn_fused_release(mpn_handle *h);
int prn_fused_trace(mpn_handle *h, int enable, unsigned long long *host_out, int capacity, int *grid_out);
// fc != NULL: crop_and_resize of the padded normalised map runs inside the kernel (prn_fused.cu: CropFuse); x_f32 must be
// logits (in place), the person list is derived from the detection outputs and written by the kernel
struct FusedCropCall {
    const float *nh;
    int hh, ww;
    PersonList pl;
    int only;               // development aid: sample the crops (bf16 and fp32) and stop
};
bool prn_fused_can_crop(const mpn_handle *h, int batch);
int launch_prn_fused(mpn_handle *h, const float *x_f32, const int *n_dev, int n_host, int n_max, float *logits, cudaStream_t s,
                     const FusedCropCall *fc = nullptr);
}
