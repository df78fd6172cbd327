/*
 * mpn_b200.h -- C ABI of libmpn_b200.so: the post-backbone inference path of
 * TropComplique/MultiPoseNet as hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI / plugin interface for this path: the boundary a user
 * sees is the Python class inference/detector.py:5-61 (`Detector`), which runs a
 * frozen TensorFlow graph built by create_pb.py:44-153.  This library replaces the
 * part of that graph downstream of the networks (create_pb.py:73-147).  Each entry
 * point below cites the reference code it stands in for; INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C types only; every tensor is a raw pointer to C-contiguous memory.
 *   - "device pointer" = CUDA device memory on the handle's GPU; "host pointer" =
 *     host memory (pinned for the *_host entry points to be asynchronous).
 *   - boxes are (ymin, xmin, ymax, xmax), normalised to [0,1] (detector/utils/box_utils.py:5-11).
 *   - every function returns MPN_OK (0) or a negative mpn_status; a message is
 *     available from mpn_last_error().  Nothing aborts, nothing falls back to the CPU.
 *   - a handle is bound to one GPU and is not re-entrant; calls are ordered on the
 *     given CUDA stream (`stream` is a cudaStream_t passed as void*, NULL = default stream).
 *   - caller owns all input / output buffers; the handle owns workspace and PRN weights;
 *     nothing is allocated inside the run calls.
 */
#ifndef MPN_B200_H
#define MPN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MPN_API __attribute__((visibility("default")))
#else
#define MPN_API
#endif

#define MPN_MAX_LEVELS 8
#define MPN_MAX_ANCHOR_SHAPES 16

typedef enum mpn_status {
    MPN_OK = 0,
    MPN_ERR_INVALID_ARGUMENT = -1,  /* bad shape / NULL pointer / H or W not divisible by 128 (inference/detector.py:45) */
    MPN_ERR_CUDA = -2,              /* a CUDA runtime call failed */
    MPN_ERR_UNSUPPORTED = -3,       /* configuration outside what the kernels implement */
    MPN_ERR_NO_WEIGHTS = -4,        /* PRN stage requested before mpn_set_prn_weights */
    MPN_ERR_CAPACITY = -5           /* call exceeds max_batch / max_height / max_width / max_detections of the handle */
} mpn_status;

typedef enum mpn_prn_mode {
    MPN_PRN_FP32 = 0,   /* fp32-accurate (parity mode, 1e-4; measured ~1e-6): <= 240 persons per call on the tensor cores with every
                         * number as three bf16 parts and fp32 accumulation (prn_split3.cu), more persons on FFMA (prn_simt.cu) */
    MPN_PRN_BF16 = 1    /* bf16 weights and activations on tcgen05 tensor cores, fp32 accumulate (1e-2) */
} mpn_prn_mode;

/* Static description of the model head; mirrors the constants of the reference:
 * AnchorGenerator(strides, scales, scale_multipliers, aspect_ratios)  detector/retinanet.py:38-43,
 * SCALE_FACTORS detector/constants.py:19, CROP_SIZE create_pb.py:19, NUM_KEYPOINTS / DOWNSAMPLE
 * detector/constants.py:10,13, PRN hidden width detector/prn.py:20.                                 */
typedef struct mpn_config {
    int32_t struct_size;            /* sizeof(mpn_config), for ABI checking */
    int32_t device;                 /* CUDA device ordinal */
    int32_t max_batch;              /* capacity: images per call */
    int32_t max_height, max_width;  /* capacity: input image size (multiples of 128) */
    int32_t max_detections;         /* capacity: max_boxes of create_pb.py:35 */
    int32_t num_levels;
    int32_t strides[MPN_MAX_LEVELS];
    double scales[MPN_MAX_LEVELS];
    int32_t num_multipliers;
    double multipliers[MPN_MAX_ANCHOR_SHAPES];
    int32_t num_ratios;
    double ratios[MPN_MAX_ANCHOR_SHAPES];
    float scale_factors[4];
    int32_t crop_height, crop_width;    /* 56, 36 */
    int32_t num_keypoints;              /* 17 (heatmap tensor has num_keypoints + 1 channels) */
    int32_t downsample;                 /* 4 */
    int32_t prn_hidden;                 /* 1024 */
    int32_t prn_modes;                  /* bit 0: allocate fp32 PRN, bit 1: allocate bf16 PRN */
} mpn_config;

/* Per-call parameters: PARAMS of create_pb.py:31-36 / RetinaNet.get_predictions(detector/retinanet.py:56). */
typedef struct mpn_params {
    float score_threshold;
    float iou_threshold;
    int32_t max_detections;   /* <= config.max_detections */
    int32_t prn_mode;         /* mpn_prn_mode */
} mpn_params;

/* Inputs at the graph cut (all fp32, C-contiguous):
 *   class_logits   [B, A]          retinanet.raw_predictions['class_predictions']  detector/retinanet.py:47-54
 *   encoded_boxes  [B, A, 4]       retinanet.raw_predictions['encoded_boxes']
 *   heatmap_logits [B, H/4, W/4, K+1]  subnet.heatmaps (NHWC)                      detector/keypoint_subnet.py:49-58
 * with A = sum over levels of ceil(H/stride)*ceil(W/stride)*n_loc in the order of
 * detector/box_predictor.py:53-90.
 * Alternatively (level_class / level_boxes non-NULL) the raw per-level NCHW head outputs
 * [B, n_loc, h_l, w_l] and [B, 4*n_loc, h_l, w_l] can be given and the reshape_and_concatenate
 * of detector/box_predictor.py:53-90 is fused away.                                                */
typedef struct mpn_inputs {
    int32_t batch;
    int32_t height, width;          /* image size the anchors are generated for */
    const float *class_logits;
    const float *encoded_boxes;
    const float *heatmap_logits;
    const float *const *level_class;    /* optional: num_levels pointers, NCHW */
    const float *const *level_boxes;    /* optional: num_levels pointers, NCHW */
} mpn_inputs;

/* Outputs = the seven named tensors of create_pb.py:23-27,149-152 (padded to max_detections):
 *   boxes [B, max_det, 4] f32, scores [B, max_det] f32, num_boxes [B] i32   (zero padded, detector/utils/nms.py:47-52)
 *   keypoint_heatmaps [B, H/4, W/4, K] f32 (sigmoid, un-normalised), segmentation_masks [B, H/4, W/4] f32 (raw logits)
 *   keypoint_scores [B*max_det, K] f32, keypoint_positions [B*max_det, K, 2] f32: rows 0..N-1 valid, persons of
 *   all images concatenated in image order (create_pb.py:96-103), N = sum(num_boxes)
 *   person_offsets [B+1] i32: row range of image b is [person_offsets[b], person_offsets[b+1]); person_offsets[B] = N.
 * keypoint_heatmaps / segmentation_masks may be NULL (not echoed back).                             */
typedef struct mpn_outputs {
    float *boxes;
    float *scores;
    int32_t *num_boxes;
    float *keypoint_heatmaps;
    float *segmentation_masks;
    float *keypoint_scores;
    float *keypoint_positions;
    int32_t *person_offsets;
} mpn_outputs;

typedef struct mpn_handle mpn_handle;

MPN_API int mpn_version(void);
MPN_API const char *mpn_last_error(const mpn_handle *h);      /* h may be NULL: last error of a failed mpn_create */
MPN_API int mpn_default_config(mpn_config *cfg);              /* the reference's constants; capacity 1 x 640 x 640 x 25 */

MPN_API int mpn_create(const mpn_config *cfg, mpn_handle **out);
MPN_API void mpn_destroy(mpn_handle *h);
MPN_API int mpn_num_anchors(const mpn_handle *h, int32_t height, int32_t width);   /* >0, or negative status */

/* PRN variables PRN/fc1/{weights,biases}, PRN/fc2/{weights,biases} (detector/prn.py:13,20,22; restored at
 * create_pb.py:182-185).  Host pointers, fp32, [in,out] row-major as slim stores them:
 * W1 [D, hidden], b1 [hidden], W2 [hidden, D], b2 [D], D = crop_height*crop_width*num_keypoints.    */
MPN_API int mpn_set_prn_weights(mpn_handle *h, const float *W1, const float *b1, const float *W2, const float *b2);

/* The whole path, device pointers in and out: create_pb.py:73-147 (+ detector/retinanet.py:56-81,
 * detector/utils/nms.py:6-61, detector/prn.py:5-25).  Asynchronous on `stream`.                      */
MPN_API int mpn_run(mpn_handle *h, const mpn_inputs *in, const mpn_params *p, const mpn_outputs *out, void *stream);

/* Same with HOST pointers for inputs and outputs: what Detector.__call__'s feed_dict / fetch does at
 * inference/detector.py:47-48.  Copy-in, the path and copy-out run on three streams of the handle, and up to
 * MPN_HOST_DEPTH calls may be in flight (the copy-in of call i+1 overlaps the kernels of call i and the copy-out of
 * call i-1).  mpn_submit_host returns after enqueueing (immediately when the buffers are pinned) and hands back a
 * ticket; mpn_wait(ticket) blocks until that call's outputs are in the caller's buffers.  The caller must not touch
 * the input and output buffers of a call before its ticket has been waited for, and must wait for ticket i before
 * re-using its buffers for call i + MPN_HOST_DEPTH.  When encoded_boxes is pinned (cudaHostAlloc /
 * cudaHostRegister) it is NOT copied: the NMS kernel gathers the 16-byte codes of the confident anchors in place.
 * mpn_run_host = mpn_submit_host without a ticket; mpn_synchronize waits for everything in flight.
 * mpn_host_traffic reports the bytes the most recent submit moved over PCIe with copy engines.              */
#define MPN_HOST_DEPTH 3
MPN_API int mpn_submit_host(mpn_handle *h, const mpn_inputs *in, const mpn_params *p, const mpn_outputs *out,
                            int64_t *ticket);
MPN_API int mpn_wait(mpn_handle *h, int64_t ticket);
MPN_API int mpn_run_host(mpn_handle *h, const mpn_inputs *in, const mpn_params *p, const mpn_outputs *out);
MPN_API int mpn_synchronize(mpn_handle *h);
MPN_API int mpn_host_traffic(const mpn_handle *h, int64_t *h2d_bytes, int64_t *d2h_bytes);

/* ---- single stages (device pointers), used by the parity tests and by the PRN-only sweep ---- */

/* anchors [A,4]: detector/anchor_generator.py:40-116 (the run path never materialises them) */
MPN_API int mpn_anchors(mpn_handle *h, int32_t height, int32_t width, float *anchors_out, void *stream);

/* detector/retinanet.py:73 + detector/utils/nms.py:6-61.  sel_anchor [B,max_det] i32 (optional): anchor index of
 * every kept box (-1 padded); n_candidates [B] i32 (optional): anchors with score > threshold.      */
MPN_API int mpn_detect(mpn_handle *h, const mpn_inputs *in, const mpn_params *p, float *boxes, float *scores,
               int32_t *num_boxes, int32_t *sel_anchor, int32_t *n_candidates, void *stream);

/* create_pb.py:73-76 and the min / max of :90,92.  minmax [B, K, 2] f32 (optional) = (min, max) per image, channel */
MPN_API int mpn_heatmaps(mpn_handle *h, const float *heatmap_logits, int32_t batch, int32_t hm_height, int32_t hm_width,
                 float *keypoint_heatmaps, float *segmentation_masks, float *minmax, void *stream);

/* The same stage in the two-pass form mpn_run uses whenever the crops are taken from the padded normalised map (every
 * BASELINE configuration except 1024 x 1024 x 64): create_pb.py:73-76 AND :90-94.  Pass 1 takes min / max of the LOGITS
 * (the sigmoid recipe is monotone over all floats -- verified exhaustively -- so the extreme activations are the
 * activations of the extreme logits, bit for bit); pass 2 writes keypoint_heatmaps, segmentation_masks and the normalised,
 * masked map `normalised` [B, hh, ww, 20]: channels 0..16 = (kh - m) / (M - m) * float(M > 0.2), channels 17..19 are
 * padding and are NOT written (NULL: the map stays in the handle's workspace).  minmax as in mpn_heatmaps.            */
MPN_API int mpn_heatmaps_normalised(mpn_handle *h, const float *heatmap_logits, int32_t batch, int32_t hm_height,
                                    int32_t hm_width, float *keypoint_heatmaps, float *segmentation_masks, float *minmax,
                                    float *normalised, void *stream);

/* SURVEY.md section 8(f) row 2, the step in front of the path: the tail of KeypointSubnet fused with mpn_heatmaps --
 * detector/keypoint_subnet.py:49-58 (heatmaps = conv2d(x, 18, kernel_size=1) + bias, then NCHW -> NHWC) followed by
 * create_pb.py:73-76,90,92.  features [B, 64, hh, ww] f32 NCHW (x after final_bn + ReLU), weight [64, 18] f32 (the
 * [1,1,64,18] HWIO kernel of the layer 'heatmaps'), bias [18].  heatmap_logits [B, hh, ww, 18] is optional (NULL: the
 * logits tensor is never materialised); the other outputs are those of mpn_heatmaps.                            */
MPN_API int mpn_heatmap_head(mpn_handle *h, const float *features, const float *weight, const float *bias, int32_t batch,
                             int32_t hm_height, int32_t hm_width, float *heatmap_logits, float *keypoint_heatmaps,
                             float *segmentation_masks, float *minmax, void *stream);

/* create_pb.py:90-94 + tf.image.crop_and_resize create_pb.py:106-109.  boxes [N,4], box_ind [N] i32,
 * minmax [B,K,2] or NULL (no normalisation) -> crops [N, crop_h, crop_w, K] f32                        */
MPN_API int mpn_crop(mpn_handle *h, const float *keypoint_heatmaps, const float *minmax, int32_t batch, int32_t hm_height,
             int32_t hm_width, const float *boxes, const int32_t *box_ind, int32_t n, float *crops, void *stream);

/* tf.image.crop_and_resize create_pb.py:106-109 of an already normalised map in the padded layout of
 * mpn_heatmaps_normalised ([B, hh, ww, 20]; the crop kernel of mpn_run wherever that map exists): crops_f32 [N, crop_h,
 * crop_w, K] f32 and / or crops_bf16 (the same values rounded to nearest even, bfloat16 bits) -- either may be NULL.   */
MPN_API int mpn_crop_padded(mpn_handle *h, const float *normalised, int32_t batch, int32_t hm_height, int32_t hm_width,
                            const float *boxes, const int32_t *box_ind, int32_t n, float *crops_f32, uint16_t *crops_bf16,
                            void *stream);

/* detector/prn.py:5-25: crops [N, D] f32 -> logits [N, D] f32.  logits == crops (bf16 mode only) computes in place:
 * x += relu(fc2(relu(fc1(x)))), which is how mpn_run uses it (the residual addition then happens in L2 by a TMA
 * reduce-add and the layer never loads x).                                                                   */
MPN_API int mpn_prn(mpn_handle *h, const float *crops, int32_t n, int32_t prn_mode, float *logits, void *stream);

/* create_pb.py:115-142: logits [N, crop_h*crop_w, K] -> scores [N,K], positions [N,K,2], argmax [N,K] i32 (optional) */
MPN_API int mpn_keypoint_decode(mpn_handle *h, const float *logits, int32_t n, float *scores, float *positions,
                        int32_t *argmax, void *stream);

/* inference/utils.py:29-52 get_keypoints: heatmaps [hh, ww, K] f32 (device), box (ymin,xmin,ymax,xmax) and threshold as
 * Python floats -> out [K,3] i32 (device) rows (x, y, visible)                                             */
MPN_API int mpn_get_keypoints(mpn_handle *h, const float *heatmaps, int32_t hh, int32_t ww, const double box[4],
                      double threshold, int32_t *out, void *stream);

/* Element-wise device exp / sigmoid of the path (bit-level test hooks): y[i] = f(x[i]) */
MPN_API int mpn_test_exp(mpn_handle *h, const float *x, float *y, int64_t n, void *stream);
MPN_API int mpn_test_sigmoid(mpn_handle *h, const float *x, float *y, int64_t n, void *stream);

/* Bit-level test hook: walks `count` consecutive floats in increasing order starting at ordered key `key_begin` (key k <
 * 2^31 is the negative float with bits ~k, k >= 2^31 the positive float with bits k - 2^31) and adds to *violations (a
 * DEVICE counter the caller zeroes) the number of neighbours x < x' with sigmoid(x) > sigmoid(x').  Zero over the whole
 * range is what lets mpn_run take min / max of the logits instead of the activations.                        */
MPN_API int mpn_test_sigmoid_monotone(mpn_handle *h, uint32_t key_begin, uint64_t count, uint64_t *violations, void *stream);

/* Test hook: copies one internal buffer of the most recent mpn_run to `dst` (host or device memory; waits for the device).
 * NORMALISED [B, hh, ww, 20] f32 (padded crop path only), CROPS_F32 [n_max, D] f32 and CROPS_BF16 [n_max, D] bfloat16 (the
 * PRN inputs; in bf16 mode the PRN runs in place, so fetch them from a run with the PRN skipped: mpn_debug_skip(16 | 32)),
 * LOGITS [n_max, D] f32 (what the keypoint decode read), MINMAX [B, K, 2], PERSON_BOX [n_max, 4], PERSON_IMAGE [n_max];
 * rows >= person_offsets[B] are stale.  dst == NULL only reports the size in *bytes_out.                       */
enum { MPN_DEBUG_NORMALISED = 0, MPN_DEBUG_CROPS_F32 = 1, MPN_DEBUG_CROPS_BF16 = 2, MPN_DEBUG_LOGITS = 3, MPN_DEBUG_MINMAX = 4,
       MPN_DEBUG_PERSON_BOX = 5, MPN_DEBUG_PERSON_IMAGE = 6 };
MPN_API int mpn_debug_fetch(mpn_handle *h, int32_t what, void *dst, int64_t capacity_bytes, int64_t *bytes_out);

/* Per-kernel device times of the most recent mpn_run (CUDA events recorded on the run's stream between the kernels;
 * used by bench.py for the roofline figures -- leave it off in production, every event costs a little launch time).
 * mpn_get_profile waits for the run to finish; names[i] are static strings, ms[i] the time from the launch of kernel i
 * to the launch of kernel i+1 (or the end of the run).                                                     */
MPN_API int mpn_set_profiling(mpn_handle *h, int32_t enable);
MPN_API int mpn_get_profile(mpn_handle *h, int32_t capacity, const char **names, float *ms, int32_t *count);

/* Development aid for timing experiments: stages whose bit is set are NOT launched by mpn_run (their outputs keep the
 * values of the previous call): 1 detect, 2 heatmap stage, 8 crop, 16 PRN, 32 keypoint decode.  64 (only with
 * MPN_FUSE_CROP=1, the crop inside the single-kernel PRN): that kernel samples the crops and stops, so that they can be
 * fetched.  0 = normal. */
MPN_API int mpn_debug_skip(mpn_handle *h, uint32_t mask);

/* Development aid: globaltimer stamps (ns) of the phases of the most recent single-kernel PRN launch, 16 slots per CTA:
 * 0 prologue, 7 crops ready, 1 fc1 loads issued, 2 / 14 MMAs of fc1 wave A / B issued, 3 / 13 accumulators of the wave complete,
 * 4 / 11 its partial sums stored, 5 / 12 every CTA's partial sums there, 8 / 15 y1 slice of the wave stored, 6 producer saw
 * the last y1 barrier, 9 fc2 accumulators complete, 10 logits stored.  Synchronises the device.                */
MPN_API int mpn_debug_fused_trace(mpn_handle *h, int32_t enable, uint64_t *host_out, int32_t capacity, int32_t *grid_out);

/* Development aid: globaltimer stamps (ns) of the sort / NMS kernel of the most recent call, 16 slots per image: 0 CTA
 * resident, 1 candidate scan complete, 2 keys sorted, 3 first chunk decoded, 4 all chunks resolved, 5 results published,
 * 6 person list built (the last CTA only).  Synchronises the device.                                         */
MPN_API int mpn_debug_nms_trace(mpn_handle *h, int32_t enable, uint64_t *host_out, int32_t capacity);

/* Counters of the most recent run (valid after the stream has been synchronised): number of kernels launched by the
 * last mpn_run / stage call, and the sum over calls since creation.                                       */
MPN_API int mpn_launch_count(const mpn_handle *h, int64_t *last_call, int64_t *total);

#ifdef __cplusplus
}
#endif
#endif /* MPN_B200_H */
