// Anchors, score threshold, candidate compaction, sort and greedy NMS.
//
// Replaces (reference, TensorFlow graph ops):
//   detector/anchor_generator.py:40-116      anchors -- never materialised here, rebuilt from the anchor index
//   detector/box_predictor.py:53-90          reshape_and_concatenate -- fused into the loads (NCHW variant)
//   detector/retinanet.py:73                 scores = sigmoid(class_predictions)
//   detector/utils/nms.py:27-53              threshold, boolean_mask, decode, clip, NonMaxSuppressionV3, gather, pad
//   detector/utils/box_utils.py:112-139      decode
//
// Kernel 1 (candidates_*): HBM-bound scan of the class logits.  A cheap conservative logit test rejects the
//   ~99.7 % background anchors; survivors get the exact sigmoid and a strict `> thr` test and are appended to a
//   per-image candidate list as 64-bit keys (score bits << 32 | ~anchor), so that a descending key order is
//   (score desc, anchor index asc) -- the matched tie-break of the oracle.
// Kernel 2 (sort_nms): one CTA per image.  Bitonic sort of the keys (shared memory up to 8192 candidates, the
//   image's global scratch above that -- exact for any candidate count), then greedy NMS in chunks of 1024
//   candidates: every thread decodes its candidate's box (4 gathered codes + anchor from the index), tests it
//   against the boxes kept so far, and the chunk is resolved with ballots, one barrier per kept box.  (The flat person
//   list of create_pb.py:96-103 is derived from num_boxes by the crop kernel: common.cuh, PersonList.)
#include "common.cuh"
#include "mpn_math.cuh"

namespace mpn {

namespace {

constexpr int kWideNmsAbove = 64;      // max_detections above this: 1024-thread sort / NMS CTAs

struct Anchor { float ymin, xmin, ymax, xmax; };

__device__ __forceinline__ int level_of(const AnchorTable &t, int a)
{
    int l = 0;
#pragma unroll
    for (int i = 1; i < kMaxLevels; ++i)
        if (i < t.n_levels && a >= t.off[i]) l = i;
    return l;
}

// detector/anchor_generator.py:92-93,141-165,110-114 in the reference's operation order
__device__ __forceinline__ Anchor anchor_from_index(const AnchorTable &t, int a, int *level_out, int *loc_out,
                                                    int *k_out)
{
    const int l = level_of(t, a);
    const int r = a - t.off[l];
    const int k = r % t.n_loc, loc = r / t.n_loc;
    const int y = loc / t.gw[l], x = loc - y * t.gw[l];
    const float cy = fadd(fmul((float)y, t.stride[l]), t.oy[l]);
    const float cx = fadd(fmul((float)x, t.stride[l]), t.ox[l]);
    const float hh = t.half_h[l][k], hw = t.half_w[l][k];
    Anchor an;
    an.ymin = fdiv(fsub(cy, hh), t.fH);
    an.xmin = fdiv(fsub(cx, hw), t.fW);
    an.ymax = fdiv(fadd(cy, hh), t.fH);
    an.xmax = fdiv(fadd(cx, hw), t.fW);
    if (level_out) { *level_out = l; *loc_out = loc; *k_out = k; }
    return an;
}

// detector/utils/box_utils.py:73-76,124-139 + clip of detector/utils/nms.py:36
__device__ __forceinline__ float4 decode_box(const Anchor an, const float4 code, const float *sf)
{
    const float ha = fsub(an.ymax, an.ymin), wa = fsub(an.xmax, an.xmin);
    const float cya = fadd(an.ymin, fmul(0.5f, ha)), cxa = fadd(an.xmin, fmul(0.5f, wa));
    const float ty = fdiv(code.x, sf[0]), tx = fdiv(code.y, sf[1]);
    const float th = fdiv(code.z, sf[2]), tw = fdiv(code.w, sf[3]);
    const float h = fmul(exact_expf(th), ha), w = fmul(exact_expf(tw), wa);
    const float cy = fadd(fmul(ty, ha), cya), cx = fadd(fmul(tx, wa), cxa);
    const float hh = fmul(0.5f, h), hw = fmul(0.5f, w);
    return make_float4(clip01(fsub(cy, hh)), clip01(fsub(cx, hw)), clip01(fadd(cy, hh)), clip01(fadd(cx, hw)));
}

// A confident anchor's sort key (score bits << 32 | ~anchor: a descending sort gives score desc, anchor asc), or 0 when the
// anchor does not pass (a score == thr passes nms.py:30 but can never be selected by the NMS op: strict >).
__device__ __forceinline__ unsigned long long candidate_key(const DetectArgs &a, int anchor, float logit)
{
    const float s = exact_sigmoidf(logit);
    return s > a.thr ? ((unsigned long long)__float_as_uint(s) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)anchor) : 0ULL;
}

// Appending the candidates of a CTA to their images' lists: ONE global atomic per CTA and image.  The slots inside the CTA
// come from a shared-memory counter; the order of a list does not matter (it is sorted, the keys are distinct).  One
// global atomic per candidate serialises on the image's counter: 2 100 candidates per image cost 47 us of a crowded call.
// `which` = the candidate's image minus the CTA's first image (a CTA touches at most kCandImgs images).
constexpr int kCandImgs = 2;
struct CandSlots { int cnt[kCandImgs], base[kCandImgs]; };

template <int PER>
__device__ __forceinline__ void append_candidates(const DetectArgs &a, CandSlots &sm, const int img0, const int n_imgs,
                                                  const unsigned long long (&key)[PER], const int (&which)[PER])
{
    // (sm.cnt zeroed and a __syncthreads() passed by the caller)
    int slot[PER];
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        slot[j] = -1;
        if (key[j] && which[j] < kCandImgs) slot[j] = atomicAdd(&sm.cnt[which[j]], 1);
        else if (key[j])                               // anchor sets so small that a CTA spans more images: one by one
            a.cand_keys[(size_t)(img0 + which[j]) * a.key_cap + atomicAdd(a.cand_count + img0 + which[j], 1)] = key[j];
    }
    __syncthreads();
    if (threadIdx.x < n_imgs && sm.cnt[threadIdx.x] > 0)
        sm.base[threadIdx.x] = atomicAdd(a.cand_count + img0 + threadIdx.x, sm.cnt[threadIdx.x]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < PER; ++j)
        if (slot[j] >= 0) a.cand_keys[(size_t)(img0 + which[j]) * a.key_cap + sm.base[which[j]] + slot[j]] = key[j];
}

__global__ void anchors_kernel(const AnchorTable t, float4 *out)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= t.num_anchors) return;
    const Anchor an = anchor_from_index(t, a, nullptr, nullptr, nullptr);
    out[a] = make_float4(an.ymin, an.xmin, an.ymax, an.xmax);
}

// Concatenated [B, A] logits: kCandVec 16-byte loads per thread over the flat array (a CTA covers
// 256 * 4 * kCandVec consecutive elements: at most two images when A >= that; smaller anchor sets still work, see
// append_candidates).  A CTA without a confident anchor -- almost all of them -- leaves after one barrier.
constexpr int kCandVec = 1;      // (4 loads per thread and a quarter of the CTAs: 9.5 -> 13.2 us at c2, 13.4 -> 20.0 at c3)
__global__ void __launch_bounds__(256) candidates_flat_kernel(const DetectArgs a, const int A, const long long total,
                                                              const int vec_ok)
{
    __shared__ CandSlots sm;
    pdl_trigger();
    if (threadIdx.x < kCandImgs) sm.cnt[threadIdx.x] = 0;
    const long long blk0 = (long long)blockIdx.x * (256 * 4 * kCandVec);
    const int img0 = (int)(blk0 / A);
    float v[kCandVec][4];
    long long e0[kCandVec];
    int n[kCandVec];
#pragma unroll
    for (int u = 0; u < kCandVec; ++u) {                 // element blk0 + (u * 256 + tid) * 4: coalesced 16-byte loads
        e0[u] = blk0 + ((long long)u * 256 + threadIdx.x) * 4;
        n[u] = e0[u] < total ? (int)min(4LL, total - e0[u]) : 0;
        if (vec_ok && n[u] == 4) {
            const float4 q = __ldg(reinterpret_cast<const float4 *>(a.cls + e0[u]));
            v[u][0] = q.x; v[u][1] = q.y; v[u][2] = q.z; v[u][3] = q.w;
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) v[u][j] = (j < n[u]) ? __ldg(a.cls + e0[u] + j) : -__int_as_float(0x7f800000);
        }
    }
    bool any = false;
#pragma unroll
    for (int u = 0; u < kCandVec; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) any |= (j < n[u]) && (v[u][j] > a.pre_thr);
    if (!__syncthreads_or(any)) return;                  // (also orders the zeroing of sm.cnt before the appends)
    unsigned long long key[4 * kCandVec];
    int which[4 * kCandVec];
#pragma unroll
    for (int u = 0; u < kCandVec; ++u)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            key[u * 4 + j] = 0ULL;
            which[u * 4 + j] = 0;
            if (j < n[u] && v[u][j] > a.pre_thr) {
                const long long e = e0[u] + j;
                const int img = (int)(e / A);
                key[u * 4 + j] = candidate_key(a, (int)(e - (long long)img * A), v[u][j]);
                which[u * 4 + j] = img - img0;
            }
        }
    append_candidates<4 * kCandVec>(a, sm, img0, min(kCandImgs, a.B - img0), key, which);
}

// Per-level NCHW logits [B, n_loc, gh, gw]: threads walk memory order, anchor index = off + (y*gw+x)*n_loc + k.
__global__ void __launch_bounds__(256) candidates_nchw_kernel(const AnchorTable t, const DetectArgs a)
{
    __shared__ CandSlots sm;
    pdl_trigger();
    if (threadIdx.x < kCandImgs) sm.cnt[threadIdx.x] = 0;
    const int img = blockIdx.y;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;   // position inside the image's concatenated levels
    unsigned long long key[1] = {0ULL};
    const int which[1] = {0};
    if (e < t.num_anchors) {
        const int l = level_of(t, e);
        const int r = e - t.off[l];
        const int plane = t.gh[l] * t.gw[l];
        const float v = __ldg(a.lv.cls[l] + (size_t)img * plane * t.n_loc + r);
        if (v > a.pre_thr) {
            const int k = r / plane, loc = r - k * plane;
            key[0] = candidate_key(a, t.off[l] + loc * t.n_loc + k, v);
        }
    }
    if (!__syncthreads_or(key[0] != 0ULL)) return;       // (also orders the zeroing of sm.cnt before the appends)
    append_candidates<1>(a, sm, img, 1, key, which);
}

__device__ __forceinline__ float4 load_code(const AnchorTable &t, const DetectArgs &a, int img, int anchor, int l,
                                            int loc, int k)
{
    if (a.enc) {
        const float *enc = a.enc_ind ? reinterpret_cast<const float *>(__ldcg(reinterpret_cast<const unsigned long long *>(a.enc_ind)))
                                     : a.enc;
        return __ldg(reinterpret_cast<const float4 *>(enc) + (size_t)img * t.num_anchors + anchor);
    }
    const int plane = t.gh[l] * t.gw[l];
    const float *p = a.lv.box[l] + ((size_t)img * t.n_loc * 4 + (size_t)k * 4) * plane + loc;   // channel k*4+coord
    return make_float4(__ldg(p), __ldg(p + plane), __ldg(p + 2 * plane), __ldg(p + 3 * plane));
}

// One CTA per image.
//   1. candidate keys -> descending order (= score desc, anchor asc).  <= 1024 candidates: rank sort in shared memory
//      (every key counts the keys above it; no barriers); more: bitonic sort, in shared memory up to 8192 keys and in the
//      image's global scratch beyond that -- exact for any candidate count.
//   2. greedy NMS in chunks of 512 candidates of that order:
//        a. every thread decodes one box (4 gathered codes + the anchor rebuilt from its index) into shared memory
//        b. ... and tests it against the boxes kept by earlier chunks
//        c. the chunk is resolved in score order with one barrier per KEPT box (not per candidate): ballots of the live
//           candidates -> first live one -> everybody tests itself against that box
//      until max_detections boxes are kept or the candidates run out.  TensorFlow's NonMaxSuppressionV3 semantics:
//      strict `IoU > thr`, kept boxes in descending score order.
// T threads per CTA: 512 by default; 1024 for calls that may keep many boxes per image (crowded scenes: thousands of
// candidates per image, where the candidate-versus-kept tests and the sort dominate and one CTA per image is all the
// parallelism there is).  Chunk = one candidate per thread, rank sort for up to two keys per thread.

template <int T>
struct NmsSmem {
    static constexpr int kChunk = T, kMaskWords = T / 32, kRankSortCap = 2 * T;
    unsigned long long keys[kSortSmemCap];     // 64 KB: unsorted (rank sort source) or bitonic work area
    unsigned long long sorted[kRankSortCap];   //  8 / 16 KB: rank sort destination
    float4 box[kChunk];                        //  8 / 16 KB: decoded boxes of the current chunk
    NmsBox nbox[kChunk];                       // 10 / 20 KB: ... with ordered corners and area (pair tests)
    NmsBox kept_nbox[kMaxDetCap];              // 20 KB
    int kept_local[kChunk];                    //  2 KB: chunk-local indices kept by step d
    unsigned alive[2][kMaskWords];             // live-candidate ballots, double buffered
    int n_kept;                                // boxes kept by the current chunk
};

__device__ __forceinline__ void nms_stamp(const DetectArgs &a, int slot)
{
    if (a.trace && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        a.trace[(size_t)blockIdx.x * 16 + slot] = t;
    }
}

template <int T>
__global__ void __launch_bounds__(T, 1) sort_nms_kernel(const AnchorTable t, const DetectArgs a)
{
    constexpr int kNmsThreads = T, kChunk = T, kMaskWords = T / 32, kRankSortCap = 2 * T;
    // The one-pass rank sort costs C * C / 2T key comparisons per thread: past one key per thread (~3 k instructions per
    // thread, all 32 warps of the SM) the bucketed sort below is cheaper, whatever the count (MPN_NMS_RANK_CUT for measurements).
#ifdef MPN_NMS_RANK_CUT
    constexpr int kRankCut = MPN_NMS_RANK_CUT < kRankSortCap ? MPN_NMS_RANK_CUT : kRankSortCap;
#else
    constexpr int kRankCut = T;
#endif
    extern __shared__ __align__(16) unsigned char smem_raw[];
    NmsSmem<T> &sm = *reinterpret_cast<NmsSmem<T> *>(smem_raw);
    const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    pdl_trigger();
    nms_stamp(a, 0);                                   // CTA resident
    pdl_wait();                                        // the candidate scan has completed
    nms_stamp(a, 1);                                   // candidates complete
    const int C = min(a.cand_count[img], a.key_cap);
    unsigned long long *gkeys = a.cand_keys + (size_t)img * a.key_cap;
    const unsigned long long *keys;
    __syncthreads();                                   // everyone has read the count ...
    if (tid == 0) a.cand_count[img] = 0;               // ... re-arm it for the next call (no reset kernel)

    if (C <= kNmsThreads / 2) {
        // Few keys (the ordinary, uncrowded image): SEVERAL threads per key.  With Cp = C rounded up to a power of two,
        // thread (part, i) = (tid / Cp, tid % Cp) counts the keys above key i in its 1 / parts slice of the list; the
        // slice counts meet in a shared-memory rank (integer additions: order does not matter).  A quarter to a half of
        // the one-thread-per-key scan, which was a quarter of this kernel's time at 230 candidates.
        int Cp = 32;
        while (Cp < C) Cp <<= 1;
        const int parts = kNmsThreads / Cp;
        for (int i = tid; i < C; i += kNmsThreads) sm.keys[i] = gkeys[i];
        if (tid == 0 && (C & 1)) sm.keys[C] = 0ULL;    // zero pad: never greater than a real key (score bits > 0)
        if (tid < Cp) sm.kept_local[tid] = 0;
        __syncthreads();
        const int i = tid & (Cp - 1), part = tid / Cp;
        const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(sm.keys);
        const int c2 = (C + 1) >> 1;
        const int j0 = (int)(((long long)part * c2) / parts), j1 = (int)(((long long)(part + 1) * c2) / parts);
        const unsigned long long k0 = i < C ? sm.keys[i] : 0ULL;
        int r0 = 0;
        if (i < C) {
#pragma unroll 4
            for (int j = j0; j < j1; ++j) {
                const ulonglong2 kj = k2[j];
                r0 += (kj.x > k0) + (kj.y > k0);
            }
            atomicAdd(&sm.kept_local[i], r0);
        }
        __syncthreads();
        if (tid < C) sm.sorted[sm.kept_local[tid]] = sm.keys[tid];
        keys = sm.sorted;
    } else if (C <= kRankCut) {
        for (int i = tid; i < C; i += kNmsThreads) sm.keys[i] = gkeys[i];
        if (tid == 0 && (C & 1)) sm.keys[C] = 0ULL;    // zero pad: never greater than a real key (score bits > 0)
        __syncthreads();
        const unsigned long long k0 = tid < C ? sm.keys[tid] : 0ULL;
        const unsigned long long k1 = tid + kNmsThreads < C ? sm.keys[tid + kNmsThreads] : 0ULL;
        int r0 = 0, r1 = 0;
        if (tid < C) {                                 // warps without a key skip the scan
            // keys are distinct (they carry the anchor index); two keys per 16-byte broadcast load
            const ulonglong2 *k2 = reinterpret_cast<const ulonglong2 *>(sm.keys);
            const int c2 = (C + 1) >> 1;
#pragma unroll 4
            for (int j = 0; j < c2; ++j) {
                const ulonglong2 kj = k2[j];
                r0 += (kj.x > k0) + (kj.y > k0);
                r1 += (kj.x > k1) + (kj.y > k1);
            }
        }
        if (tid < C) sm.sorted[r0] = k0;
        if (tid + kNmsThreads < C) sm.sorted[r1] = k1;
        keys = sm.sorted;
    } else if (C <= kSortSmemCap) {
        // Crowded images (thousands of keys): one pass of bucketing + exact ranks inside the buckets instead of a bitonic
        // network (78 block barriers for 4096 padded keys: 37 us at 2 100 candidates).  A key's bucket is a monotone
        // function of its score bits -- 2^17 ulps per bucket, 512 buckets down from 1.0: scores above 1 / 256 spread over
        // them, anything lower shares the last one -- so buckets are ordered among themselves and only keys of one bucket
        // have to be compared with each other, on the full 64-bit key: the result is the exact descending order, whatever
        // the distribution (a bucket that swallows everything only costs time).  Keys are bucketed into shared memory and
        // leave in final order to the image's global scratch, from where the NMS chunks read them.
        constexpr int kBuckets = 512;
        int *hist = reinterpret_cast<int *>(sm.sorted);          // [kBuckets] counts, then bucket starts
        int *cursor = hist + kBuckets;                           // [kBuckets] fill cursors
        int *warp_tot = cursor + kBuckets;                       // [32]
        auto bucket_of = [](unsigned long long key) {            // 0 = highest scores
            const int v = (int)((unsigned)(key >> 32) >> 17) - ((0x3F800000 >> 17) - (kBuckets - 1));
            return (kBuckets - 1) - min(max(v, 0), kBuckets - 1);
        };
        for (int i = tid; i < 2 * kBuckets; i += kNmsThreads) hist[i] = 0;
        __syncthreads();
        for (int i = tid; i < C; i += kNmsThreads) atomicAdd(&hist[bucket_of(gkeys[i])], 1);
        __syncthreads();
        {   // exclusive scan of the 512 counts (threads 0..511: 16 warps)
            int v = tid < kBuckets ? hist[tid] : 0, incl = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += u;
            }
            if (tid < kBuckets && lane == 31) warp_tot[warp] = incl;
            __syncthreads();
            if (tid < kBuckets) {
                int before = 0;
                for (int w = 0; w < warp; ++w) before += warp_tot[w];
                hist[tid] = before + incl - v;                   // start of bucket tid
            }
        }
        __syncthreads();
        for (int i = tid; i < C; i += kNmsThreads) {
            const unsigned long long key = gkeys[i];
            const int bkt = bucket_of(key);
            sm.keys[hist[bkt] + atomicAdd(&cursor[bkt], 1)] = key;
        }
        __syncthreads();
        for (int i = tid; i < C; i += kNmsThreads) {             // exact rank inside the bucket
            const unsigned long long key = sm.keys[i];
            const int bkt = bucket_of(key);
            const int b0 = hist[bkt], b1 = b0 + cursor[bkt];
            int r = 0;
            for (int j = b0; j < b1; ++j) r += sm.keys[j] > key;
            gkeys[b0 + r] = key;
        }
        keys = gkeys;
    } else {
        int npad = 1;
        while (npad < C) npad <<= 1;
        unsigned long long *work = gkeys;                        // more keys than shared memory holds: the global scratch
        for (int i = C + tid; i < npad; i += kNmsThreads) gkeys[i] = 0ULL;
        __syncthreads();
        for (int k = 2; k <= npad; k <<= 1) {          // bitonic sort, descending
            for (int j = k >> 1; j > 0; j >>= 1) {
                for (int p = tid; p < (npad >> 1); p += kNmsThreads) {
                    const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                    const int ixj = i | j;
                    const unsigned long long x = work[i], y = work[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (x < y) : (x > y)) { work[i] = y; work[ixj] = x; }
                }
                __syncthreads();
            }
        }
        keys = work;
    }
    __syncthreads();
    nms_stamp(a, 2);                                   // keys sorted

    float *boxes_out = a.boxes + (size_t)img * a.max_det * 4;
    float *scores_out = a.scores + (size_t)img * a.max_det;
    int kept = 0;
    // Chunks of the sorted list.  One chunk when the image's candidates fit a CTA; otherwise (crowded image) the chunks
    // start at 128 candidates and double: the boxes that get kept sit near the top of the list, where a chunk is resolved
    // with one barrier per kept box -- among 4 warps that costs a third of what it costs among 32 -- while the long tail
    // mostly dies in the parallel test against the boxes kept so far.
    int step = C <= kChunk ? kChunk : 128;
    int n_chunk = 0;
    for (int base = 0, n = 0; base < C && kept < a.max_det; base += n, step = min(2 * step, kChunk)) {
        n = min(step, C - base);
        // ---- a, b: decode, test against the boxes kept so far
        bool alive = tid < n;
        float4 box = make_float4(0.f, 0.f, 0.f, 0.f);
        NmsBox nb = nms_prepare(box);
        float score = 0.f;
        int anchor = -1;
        if (alive) {
            const unsigned long long key = keys[base + tid];
            score = __uint_as_float((unsigned)(key >> 32));
            anchor = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
            int l, loc, k;
            const Anchor an = anchor_from_index(t, anchor, &l, &loc, &k);
            box = decode_box(an, load_code(t, a, img, anchor, l, loc, k), t.sf);
            sm.box[tid] = box;
            nb = nms_prepare(box);
            sm.nbox[tid] = nb;
        }
        // A survivor has to be tested against EVERY kept box: one thread walking ~88 boxes of a crowded image is 10-18 us
        // per chunk (profiles/r03d_nms_chunks.txt), whatever the chunk's size.  While a chunk has fewer candidates than the
        // CTA has threads, g threads share a candidate's walk (kept boxes p, p + g, ...; combined by shuffles); a full
        // chunk takes four kept boxes per trip, their overlaps computed independently of each other.
        const int share = kept > 0 ? kNmsThreads / n : 1;
        if (share >= 2) {
            int g = 2;
            while (g * 2 <= share && g < 32) g *= 2;
            __syncthreads();                           // sm.nbox of the chunk complete
            const int cnd = tid / g, p = tid & (g - 1);
            bool hit = false;
            if (cnd < n) {
                const NmsBox cb = sm.nbox[cnd];
                for (int j = p; j < kept; j += g)
                    if (nms_iou(cb, sm.kept_nbox[j]) > a.iou_thr) { hit = true; break; }
            }
            for (int o = g >> 1; o > 0; o >>= 1) hit |= __shfl_xor_sync(0xffffffffu, hit ? 1 : 0, o) != 0;
            if (cnd < n && p == 0) sm.kept_local[cnd] = hit ? 0 : 1;
            __syncthreads();
            if (tid < n) alive = sm.kept_local[tid] != 0;
        } else if (alive) {
            int j = 0;
            for (; j + 4 <= kept; j += 4) {
                const bool o0 = nms_iou(nb, sm.kept_nbox[j]) > a.iou_thr, o1 = nms_iou(nb, sm.kept_nbox[j + 1]) > a.iou_thr;
                const bool o2 = nms_iou(nb, sm.kept_nbox[j + 2]) > a.iou_thr, o3 = nms_iou(nb, sm.kept_nbox[j + 3]) > a.iou_thr;
                if (o0 | o1 | o2 | o3) { alive = false; break; }
            }
            for (; alive && j < kept; ++j)
                if (nms_iou(nb, sm.kept_nbox[j]) > a.iou_thr) alive = false;
        }
        if (base == 0) nms_stamp(a, 3);                // first chunk decoded and tested against the kept boxes
        if (n_chunk < 5) nms_stamp(a, 6 + 2 * n_chunk); // chunk i: tested against the kept boxes (slot 6 + 2 i) ...
        const int n_warps = (n + 31) >> 5;             // warps that hold candidates of this chunk
        if (warp >= n_warps && lane == 0) { sm.alive[0][warp] = 0u; sm.alive[1][warp] = 0u; }
        __syncthreads();                               // sm.box complete, idle warps' ballots zeroed
        // ---- c: resolve the chunk in score order, ONE barrier per kept box: every warp publishes its ballot of live
        //      candidates (double buffered), every thread finds the first live candidate from the 16 words, reads that
        //      candidate's box from sm.box and drops itself if it overlaps it too much; the winner records itself.
        int nk = 0;
        const int room = a.max_det - kept;
        unsigned my_word = __ballot_sync(0xffffffffu, alive);
        // only the warps that hold candidates take part (named barrier 1): fewer warps, cheaper barrier
        for (int iter = 0; warp < n_warps && nk < room; ++iter) {
            unsigned *pub = sm.alive[iter & 1];
            if (lane == 0) pub[warp] = my_word;
            asm volatile("bar.sync 1, %0;" ::"r"(n_warps * 32) : "memory");
            const unsigned wv = lane < kMaskWords ? pub[lane] : 0u;
            const unsigned nz = __ballot_sync(0xffffffffu, wv != 0u);
            if (nz == 0u) break;                       // uniform over the CTA: nothing left alive in this chunk
            const int fw = __ffs(nz) - 1;
            const int first = (fw << 5) + __ffs(__shfl_sync(0xffffffffu, wv, fw)) - 1;
            if (tid == first) {
                sm.kept_local[nk] = first;
                alive = false;
            } else if (alive && nms_iou(nb, sm.nbox[first]) > a.iou_thr) {
                alive = false;
            }
            ++nk;
            my_word = __ballot_sync(0xffffffffu, alive);
        }
        if (tid == 0) sm.n_kept = nk;                  // warp 0 always takes part
        __syncthreads();                               // sm.kept_local, sm.n_kept complete
        nk = sm.n_kept;
        // ---- e: the kept candidates write themselves out
        for (int k = tid; k < nk; k += kNmsThreads) {
            const int i = sm.kept_local[k];
            const float4 bx = sm.box[i];
            const unsigned long long key = keys[base + i];
            sm.kept_nbox[kept + k] = sm.nbox[i];
            reinterpret_cast<float4 *>(boxes_out)[kept + k] = bx;
            scores_out[kept + k] = __uint_as_float((unsigned)(key >> 32));
            if (a.sel_anchor)
                a.sel_anchor[(size_t)img * a.max_det + kept + k] = (int)(0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull));
        }
        kept += nk;
        __syncthreads();
        if (n_chunk < 5) nms_stamp(a, 7 + 2 * n_chunk); // ... and resolved (slot 7 + 2 i)
        ++n_chunk;
    }
    nms_stamp(a, 4);                                   // all chunks resolved
    // zero padding (detector/utils/nms.py:47-52)
    for (int k = kept + tid; k < a.max_det; k += kNmsThreads) {
        reinterpret_cast<float4 *>(boxes_out)[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        scores_out[k] = 0.f;
        if (a.sel_anchor) a.sel_anchor[(size_t)img * a.max_det + k] = -1;
    }
    if (tid == 0) {
        a.num_boxes[img] = kept;
        if (a.n_candidates) a.n_candidates[img] = C;
        nms_stamp(a, 5);                               // results published
    }
}

__global__ void set_pointer_kernel(const float **slot, const float *value) { *slot = value; }

}  // namespace

int launch_set_pointer(const float **slot, const float *value, cudaStream_t s)
{
    set_pointer_kernel<<<1, 1, 0, s>>>(slot, value);
    return 1;
}

int launch_anchors(const AnchorTable &t, float *out, cudaStream_t s)
{
    anchors_kernel<<<(t.num_anchors + 255) / 256, 256, 0, s>>>(t, reinterpret_cast<float4 *>(out));
    return 1;
}

// per device, at handle creation: the sort / NMS kernel needs more than the default 48 KB of dynamic shared memory
int detect_prepare()
{
    cudaError_t e = cudaFuncSetAttribute(sort_nms_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NmsSmem<512>));
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(sort_nms_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(NmsSmem<1024>));
    return e == cudaSuccess ? 0 : -(int)e;
}

int launch_detect(const AnchorTable &t, const DetectArgs &a, cudaStream_t s, cudaEvent_t after_candidates)
{
    int launches = 0;
    if (a.cls) {
        const long long total = (long long)a.B * t.num_anchors;
        const long long per_cta = 256LL * 4 * kCandVec;
        const int vec_ok = (reinterpret_cast<uintptr_t>(a.cls) % 16 == 0) ? 1 : 0;
        prof_mark(s, "candidates_flat");
        candidates_flat_kernel<<<(unsigned)((total + per_cta - 1) / per_cta), 256, 0, s>>>(a, t.num_anchors, total, vec_ok);
    } else {
        dim3 grid((t.num_anchors + 255) / 256, a.B);
        prof_mark(s, "candidates_nchw");
        candidates_nchw_kernel<<<grid, 256, 0, s>>>(t, a);
    }
    ++launches;
    if (after_candidates) cudaEventRecord(after_candidates, s);
    prof_mark(s, "sort_nms");
    if (a.max_det > kWideNmsAbove)
        launch_k(sort_nms_kernel<1024>, dim3(a.B), dim3(1024), sizeof(NmsSmem<1024>), s, true, t, a);
    else
        launch_k(sort_nms_kernel<512>, dim3(a.B), dim3(512), sizeof(NmsSmem<512>), s, true, t, a);
    ++launches;
    return launches;
}

}  // namespace mpn
