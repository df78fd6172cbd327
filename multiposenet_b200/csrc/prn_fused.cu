// Pose Residual Network, bf16 mode, SMALL-BATCH regime (<= 256 persons per call): ONE persistent, cooperative kernel.
//
// Replaces detector/prn.py:15-25 (slim.fully_connected x2 + residual):
//   y1 = relu(x W1 + b1),  y2 = relu(y1 W2 + b2),  logits = x + y2
//
// Below ~210 persons the PRN is bound by the HBM stream of its 140 MB of bf16 weights, and a
// 70 MB stream timed alone cannot beat ~16 us on a B200 (7 us of launch + ramp + tail around 9.6 us of transfer,
// tools/stream_bench.cu).  So the two layers are NOT two kernels: one CTA per SM stays resident and the weight stream
// never stops --
//   phase 1  fc1: CTA (hq, z) owns hidden quarter hq (256 units) and K split z of the 34272-long reduction;
//            TMA ring -> tcgen05.mma, D[weight row, person] = W tile x activations: M = 128 weight rows (two MMAs per
//            256-row tile), N = persons padded to 16, accumulators in TMEM -> fp32 partial sums to L2
//   barrier  grid-wide (all CTAs are co-resident: cooperative launch)
//   phase 2  split-K reduce in fixed order + bias + ReLU + bf16 -> y1 (deterministic, no atomics on data)
//   barrier
//   phase 3  fc2: CTA t owns 240 output columns (143 tiles); A = y1 (L2-resident), B = W2 tile;
//            epilogue logits = x + relu(acc + b2): in place (logits == x, how mpn_run calls it) as a TMA reduce-add of
//            relu(acc + b2) into x, performed in L2; out of place by loading x
// The producer warp runs ahead: while the CTA sits in the barriers / the reduce it is already pulling its first W2
// tiles into the ring, so HBM stays busy across the phase boundaries.  Under programmatic dependent launch the prologue
// and the first three W1 boxes are under way before the crop kernel has finished.
//
// The weights are the A operand and the persons the N dimension of the MMA, so everything that follows the weight
// stream scales with the person count: MMA time (128 N / 256 cycles per instruction: 40 instead of 128 cycles at 80
// persons), accumulator size (2 x N TMEM columns), TMEM drain (64 B / cycle / SM).  A TMEM lane then is a weight row,
// i.e. a hidden unit / output column: 32 lanes of a warp are one 128-byte row segment of a person's row in global
// memory, and the epilogues store (or TMA-reduce-add) straight from the accumulator layout, no transpose.
//
// Roles: warp 0 = TMA producer (one thread), warp 1 = MMA issuer (one thread) + TMEM allocation, warps 2..17 =
// epilogue / reduce (TMEM lane quadrant = warp % 4; the four warps of a quadrant share its (half, 16-person chunk) items).
// Activations are fetched as 16-row TMA boxes, only as many as there are persons.
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "handle.cuh"
#include "mpn_math.cuh"
#include "tcgen05_utils.cuh"

namespace mpn {

namespace {

using namespace tc;

// fc1 runs in kWaves passes over the hidden dimension (MPN_FC1_WAVES, build.py).  One wave: CTA = hidden quarter x 1 of
// 37 K splits, and the weight stream stops while the grid stores its partial sums, meets, reduces and meets again.  Two
// waves: wave w covers hidden half w (CTA = one of its two quarters x 1 of 74 K splits), and only the epilogue warps
// take part in the store / barrier / reduce / barrier sequence of a wave -- the producer and the MMA warp go straight
// on to the next wave's weights (second accumulator stage in TMEM), and the first half of fc2's reduction only needs the
// first half of y1, so the weight stream continues through both sequences.  Price: twice the partial sums.
#ifndef MPN_FC1_WAVES
#define MPN_FC1_WAVES 1
#endif
constexpr int kWaves = MPN_FC1_WAVES;
static_assert(kWaves == 1 || kWaves == 2, "fc1 waves");
constexpr int kHq = 4;                                   // hidden quarters of fc1
constexpr int kTilesPerWave = kHq / kWaves;              // CTA tiles (quarters) side by side in one wave
constexpr int kFc1N = 256, kFc2N = 240;                  // weight rows per CTA tile of the two layers
constexpr int kRedThreads = 3 * kWaves;                  // threads that share one float4 of the split-K reduce
constexpr int kRedLoads = 14;                            //   each with <= 14 partial loads in flight
constexpr int kRedGroups = 32 / kRedThreads;             // float4 outputs per warp and pass
constexpr int kStageCols = 128;                          // TMEM column offset of the second accumulator stage (<= 128 persons)
constexpr int kBarLine = 16;                             // grid barrier counters sit on separate 128-byte lines
constexpr int kXBox = 16;                                // rows per activation TMA box (persons are padded to 16)
constexpr int kXBoxBytes = kXBox * 128;
constexpr int kXTileBytes = 128 * 128;                   // one 128-row activation tile (16 KB)
constexpr int kWTileBytes = 256 * 128;                   // one weight tile (32 KB; fc2 uses 240 of the 256 rows)
constexpr int kWHalfBytes = 128 * 128;                   // one UMMA A operand: 128 weight rows
constexpr int kAccCols = 256;                            // TMEM column of the second accumulator half
constexpr int kRingBytes = 192 * 1024;                   // stage = [W tile | X tile 0 | X tile 1 (only for > 128 persons)]
constexpr int kMaxStages = 4;                            //   <= 128 persons: 4 stages of 48 KB, else 3 stages of 64 KB
constexpr int kEarly = 3;                                // W1 boxes requested before the person count is known
constexpr int kEpiWarps = 16;                           // 4 per TMEM lane quadrant: the epilogues are issue-latency bound
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kStgOffset = kRingBytes;                   // 16 x 2 KB: per-warp 16 x 32 fp32 boxes of the in-place epilogue
constexpr int kBarOffset = kStgOffset + kEpiWarps * 2048;
constexpr int kOffOffset = kBarOffset + (2 * kMaxStages + 4) * 8 + 16;   // fused crop: person offsets of <= 256 images, u16
constexpr int kSmemBytes = kOffOffset + 528 + 1024;
constexpr uint32_t kTmemCols = 512;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

// crop_and_resize (create_pb.py:106-109) inside this kernel, for calls of mpn_run that fit it (padded normalised map,
// capacity <= 256 persons, in place): the crop kernel, its launch and its 16 MB round trip through L2 leave the call's
// critical chain.  The epilogue warps are idle while a layer streams, and that is when they sample --
//   during fc1: the bf16 operand.  The four CTAs of a K split (one per hidden quarter) consume the same 928 crop columns of
//     every person; each of them samples a quarter of the persons (in chunks of kCropChunk k blocks, staged in shared memory,
//     written to crops_bf16 as 16-byte vectors) and publishes the chunk on a counter of its split; the TMA producer
//     fetches the activation boxes of a k block when the four parts of its chunk are there.  The weight boxes do not wait.
//   during fc2: the fp32 residual.  The CTA samples the 240 columns of its own output tile for every person and stores
//     them where the TMA reduce-add of its epilogue then adds relu(acc + b2): nobody else touches those columns.
// The arithmetic is crop_padded_kernel's (heatmap.cu), operation for operation: same bits.
struct CropFuse {
    const float *nh;          // padded normalised map [B, hh, ww, 20]; NULL: the crops are in memory already
    int hh, ww;
    PersonList pl;            // derived mode: num_boxes, det_boxes, B (<= 256), max_det + where the flat list goes
    __nv_bfloat16 *x_bf16;    // [rows, D]: written here, fetched by TMA (tmap_x)
    unsigned long long *xready;   // [splits, 4] never-reset counters: 4 arrivals per launch and chunk in use
    int only;                 // development aid: sample (both forms) and stop
};
constexpr int kCropH = 56, kCropW = 36, kCropCh = 17, kCropPad = 20, kCropGroups = 5;
constexpr int kCropChunk = 5;                            // k blocks per published chunk of the bf16 operand (<= 4 chunks)
constexpr int kStgRows = kEpiWarps * 2048 / 128;         // 128-byte rows of the staging area (256)

struct FusedArgs {
    CropFuse crop;
    const int *n_dev;
    int n_host;
    int early_boxes;         // W1 boxes requested before the person count is known: kEarly, or 0 when the call's capacity
                             // exceeds this kernel's regime (crowded calls: it mostly exits, and 14 MB of W1 would be
                             // fetched and drained for nothing -- 10 us of a c3 / c4 call)
    int D, hidden;
    int nkb1;                // k blocks of fc1 (ceil(D / 64))
    int nkb2;                // k blocks of fc2 (hidden / 64)
    int splits;              // K splits of fc1
    int tiles2;              // output tiles of fc2 (ceil(D / 240))
    float *partial;          // [splits, rows_cap, hidden]
    size_t split_stride;
    const float *b1;
    __nv_bfloat16 *y1;       // [rows, hidden]
    const float *b2;
    const float *x;          // fp32 crops [N, D] (residual)
    float *logits;           // [N, D]
    unsigned long long *arrivals;   // grid barriers: 2 * kWaves never-reset arrival counters (grid arrivals per launch each)
    unsigned long long *trace;      // optional [grid, 16] globaltimer stamps (mpn_debug_fused_trace)
};

struct FusedMaps {
    CUtensorMap x, w1, y1, w2;
    CUtensorMap out, out16;    // fp32 boxes (16 persons x 32 / 16 columns) over the buffer of the in-place call
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// Grid barriers over counters that are never reset: every launch that gets past the person-count check adds exactly
// grid arrivals to each of the 2 * kWaves counters (one per meeting point; a CTA may arrive at a later meeting point
// before a slow CTA has arrived at an earlier one, so they cannot share a counter), so counter 0 at kernel entry, rounded
// down to a multiple of grid, is this launch's base for all of them (no CTA can see more than grid - 1 arrivals at the
// first meeting point before it has arrived itself).  All CTAs are co-resident (cooperative launch).  One thread per CTA
// arrives; one round trip to L2 to arrive, one to observe.
// (A two-level version -- arrivals on per-group lines, last arrival of a group on a top counter, waiters polling a
// separate epoch word -- was measured: its three dependent fence + atomic hops cost 3.2 and 4.0 us per barrier against
// 1.4 and 2.1 us for this one.)
__device__ __forceinline__ void grid_arrive(unsigned long long *arrivals)
{
    // Called by ONE thread after a CTA-wide bar.sync: the barrier orders the other threads' stores before this
    // fence, whose cumulativity publishes them at GPU scope (the cooperative-groups grid.sync pattern).
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    atomicAdd(arrivals, 1ULL);
}

__device__ __forceinline__ bool grid_poll(const unsigned long long *arrivals, unsigned long long target)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(arrivals) : "memory");
    if (v < target) return false;
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    return true;
}

__device__ __forceinline__ void grid_wait(const unsigned long long *arrivals, unsigned long long target)
{
    // Poll with RELAXED loads and fence once at the end: an acquire load is compiled to LDG + CCTL.IVALL, and an L1
    // invalidation every few hundred nanoseconds disturbs the memory pipeline of the warps that are still working.
    for (unsigned spin = 0; spin < (1u << 24); ++spin) {
        if (grid_poll(arrivals, target)) return;
        __nanosleep(64);
    }
    __trap();
}

__device__ __forceinline__ void stamp(const FusedArgs &a, int slot)
{
    if (a.trace) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        a.trace[(size_t)blockIdx.x * 16 + slot] = t;
    }
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)     // no arrival: the arrive comes later
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

// ---- crop_and_resize sampling (fused crop; crop_padded_kernel's arithmetic) --------------------------------------------------
struct CropGeo {
    float ay, hs, ax, ws;             // in_y = ay + cy * hs, in_x = ax + cx * ws
    const float4 *img;                // the person's image in the padded normalised map
};

// Person `row` of the flat list (create_pb.py:96-103) is box k of image b, where s_off (exclusive scan of num_boxes, B + 1
// entries) brackets row.
__device__ __forceinline__ CropGeo crop_geo(const CropFuse &cf, const unsigned short *s_off, int row)
{
    int lo = 0, hi = cf.pl.B;                          // s_off[lo] <= row < s_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int)s_off[mid] <= row) lo = mid; else hi = mid;
    }
    const int k = row - (int)s_off[lo];
    const float4 box = __ldcg(reinterpret_cast<const float4 *>(cf.pl.det_boxes) + (lo * cf.pl.max_det + k));
    const float hm1 = (float)(cf.hh - 1), wm1 = (float)(cf.ww - 1);
    CropGeo g;
    g.hs = fdiv(fmul(fsub(box.z, box.x), hm1), (float)(kCropH - 1));
    g.ay = fmul(box.x, hm1);
    g.ws = fdiv(fmul(fsub(box.w, box.y), wm1), (float)(kCropW - 1));
    g.ax = fmul(box.y, wm1);
    g.img = reinterpret_cast<const float4 *>(cf.nh + (size_t)lo * cf.hh * cf.ww * kCropPad);
    return g;
}

struct CropTaps {
    float4 tl, tr, bl, br;
    float lx, ly;
    bool inside;
};

// the four taps of channels 4 grp .. 4 grp + 3 at crop pixel p = cy * 36 + cx (loads only)
__device__ __forceinline__ CropTaps crop_taps(const CropGeo &g, int hh, int ww, int p, int grp)
{
    const int cy = p / kCropW, cx = p - cy * kCropW;
    const float in_y = fadd(g.ay, fmul((float)cy, g.hs)), in_x = fadd(g.ax, fmul((float)cx, g.ws));
    CropTaps t;
    t.inside = !(in_y < 0.0f || in_y > (float)(hh - 1)) && !(in_x < 0.0f || in_x > (float)(ww - 1));
    const int top = (int)floorf(in_y), bot = (int)ceilf(in_y), left = (int)floorf(in_x), right = (int)ceilf(in_x);
    t.ly = fsub(in_y, (float)top);
    t.lx = fsub(in_x, (float)left);
    const unsigned o_tl = t.inside ? (unsigned)((top * ww + left) * kCropGroups + grp) : 0u;
    const unsigned o_tr = t.inside ? (unsigned)((top * ww + right) * kCropGroups + grp) : 0u;
    const unsigned o_bl = t.inside ? (unsigned)((bot * ww + left) * kCropGroups + grp) : 0u;
    const unsigned o_br = t.inside ? (unsigned)((bot * ww + right) * kCropGroups + grp) : 0u;
    t.tl = __ldg(g.img + o_tl); t.tr = __ldg(g.img + o_tr); t.bl = __ldg(g.img + o_bl); t.br = __ldg(g.img + o_br);
    return t;
}

__device__ __forceinline__ float crop_lerp(float tl, float tr, float bl, float br, float lx, float ly, bool inside)
{
    const float tpv = fadd(tl, fmul(fsub(tr, tl), lx));
    const float btv = fadd(bl, fmul(fsub(br, bl), lx));
    return inside ? fadd(tpv, fmul(fsub(btv, tpv), ly)) : 0.0f;
}

__device__ __forceinline__ void crop_values(const CropTaps &t, float o[4])
{
    o[0] = crop_lerp(t.tl.x, t.tr.x, t.bl.x, t.br.x, t.lx, t.ly, t.inside);
    o[1] = crop_lerp(t.tl.y, t.tr.y, t.bl.y, t.br.y, t.lx, t.ly, t.inside);
    o[2] = crop_lerp(t.tl.z, t.tr.z, t.bl.z, t.br.z, t.lx, t.ly, t.inside);
    o[3] = crop_lerp(t.tl.w, t.tr.w, t.bl.w, t.br.w, t.lx, t.ly, t.inside);
}

// Samples columns [kA, kB) of person rows first, first + step, ... (n_rows of them, rows >= N skipped) into a staging image
// of `rowlen` T per row: threads-per-row x rows-at-once is chosen from the row count; two items (8 tap loads) in flight
// per thread.  Called by all epilogue warps.
template <typename T>
__device__ __forceinline__ void crop_sample_rows(const CropFuse &cf, const unsigned short *s_off, int N, int first, int step,
                                                 int n_rows, int kA, int kB, T *stg, int rowlen, int tid_e)
{
    const int tp = (kEpiWarps * 32) / n_rows;          // threads per row (n_rows <= 512)
    const int pi = tid_e / tp, sub = tid_e - pi * tp;
    const int row = first + pi * step;
    if (pi >= n_rows || row >= N) return;
    const CropGeo g = crop_geo(cf, s_off, row);
    const int p_lo = kA / kCropCh, p_hi = (kB - 1) / kCropCh;
    const int n_items = (p_hi - p_lo + 1) * kCropGroups;
    T *dst = stg + (size_t)pi * rowlen - kA;
    auto put = [&](int p, int grp, const float o[4]) {
        const int k0 = p * kCropCh + 4 * grp;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int k = k0 + t;
            if ((t == 0 || grp < kCropGroups - 1) && k >= kA && k < kB) {     // group 4 holds channel 16 only
                if (sizeof(T) == 2) reinterpret_cast<__nv_bfloat16 *>(dst)[k] = __float2bfloat16_rn(o[t]);
                else reinterpret_cast<float *>(dst)[k] = o[t];
            }
        }
    };
    for (int j = sub; j < n_items; j += 2 * tp) {
        const int j2 = j + tp;
        const int pa = p_lo + j / kCropGroups, ga = j - (j / kCropGroups) * kCropGroups;
        const bool two = j2 < n_items;
        const int pb = two ? p_lo + j2 / kCropGroups : pa, gb = two ? j2 - (j2 / kCropGroups) * kCropGroups : ga;
        const CropTaps ta = crop_taps(g, cf.hh, cf.ww, pa, ga);
        const CropTaps tb = crop_taps(g, cf.hh, cf.ww, pb, gb);
        float o[4];
        crop_values(ta, o);
        put(pa, ga, o);
        if (two) {
            crop_values(tb, o);
            put(pb, gb, o);
        }
    }
}

// number of persons of the call = sum of num_boxes (every warp for itself: one L2 round trip)
__device__ __forceinline__ int crop_count_persons(const PersonList &pl, int lane)
{
    int n = 0;
    for (int b0 = 0; b0 < pl.B; b0 += 32) n += (b0 + lane < pl.B) ? __ldcg(pl.num_boxes + b0 + lane) : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
    return n;
}

// kInPlace: logits == x (how mpn_run calls it).  The fc2 epilogue then never loads the residual: bias and ReLU are applied
// in registers, the 16-person x 32-column item goes to the warp's staging box (dense rows, no swizzle) and a TMA reduce-add performs
// x += y2 in L2 -- the same single fp32 rounding, no residual registers, 4 B / element less L2 -> SM traffic.
// kCrop: the instantiation that can sample the crops itself (CropFuse); the others carry none of that code.
// kRed: fc1's K splits are added up by the L2 (red.global.add.f32 into ONE [persons, hidden] accumulator that the reduce
// step reads, clears and converts) instead of being stored as 37 partial sums and read back -- no 12 MB read-back, but the
// order of the additions, and with it the last bits of the logits, changes from run to run (MPN_PRN_RED_ADD=1; measured,
// off by default: every bit-identity guarantee of the library rests on the fixed-order reduce).
template <bool kInPlace, bool kCrop, bool kRed = false>
__global__ void __launch_bounds__(kThreads, 1)
prn_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                 const __grid_constant__ CUtensorMap tmap_y1, const __grid_constant__ CUtensorMap tmap_w2,
                 const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_out16,
                 const FusedArgs args)
{
    extern __shared__ uint8_t smem_raw[];
    pdl_trigger();
    if (args.early_boxes == 0 && !(kCrop && args.crop.nh != nullptr)) {
        // the call's capacity exceeds this kernel's regime: it will most likely exit, so look at the count first, before
        // tensor memory, barriers and weight boxes are set up for nothing
        pdl_wait();
        const int n = args.n_dev ? *args.n_dev : args.n_host;
        if (!(n > 0 && n <= kPrnFusedMaxRows)) return;
    }
    const int G = gridDim.x, c = blockIdx.x;
    // this launch's barrier base; read before this CTA can possibly have arrived anywhere
    const unsigned long long bar_base = (ld_acquire_u64(args.arrivals) / (unsigned long long)G) * (unsigned long long)G;
    const unsigned long long bar_target = bar_base + (unsigned long long)G;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kBarOffset);
    uint64_t *empty_bar = full_bar + kMaxStages;
    uint64_t *tmem_full_bar = empty_bar + kMaxStages;      // [2]: one per accumulator stage
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;          // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_y1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w2)) : "memory");
        for (int i = 0; i < kMaxStages; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tmem_full_bar + i, 1); mbar_init(tmem_empty_bar + i, kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (threadIdx.x == 0) stamp(args, 0);                                   // prologue done

    const bool has_fc1 = c < kTilesPerWave * args.splits;
    const int tq = c % kTilesPerWave, z = c / kTilesPerWave;               // quarter inside a wave, K split
    const int kb0 = has_fc1 ? (int)(((long long)z * args.nkb1) / args.splits) : 0;
    const int kb1 = has_fc1 ? (int)(((long long)(z + 1) * args.nkb1) / args.splits) : 0;

    // Everything above, and the first weight tiles, do not depend on the crops: under programmatic dependent launch
    // they overlap the tail of the crop kernel.  Weight slot st sits at st * 32 KB in both ring layouts, so the first
    // kEarly W1 boxes can be requested before the person count (and with it the layout) is known; their barriers get the
    // transaction bytes now and the producer's arrival (with the activation bytes) after the wait.
    int early = 0;
    if (threadIdx.x == 0 && has_fc1) {
        early = min(args.early_boxes, kb1 - kb0);
        for (int i = 0; i < early; ++i) {
            mbar_expect_tx(full_bar + i, kFc1N * 128);
            tma_load_2d(smem + i * kWTileBytes, &tmap_w1, full_bar + i, (kb0 + i) * BLOCK_K, tq * kFc1N, kEvictFirst);
        }
    }
    // fused crop: this launch's base of the chunk counters of split z (read before any of the four CTAs can have arrived)
    const bool crop = kCrop && kInPlace && kWaves == 1 && args.crop.nh != nullptr;
    unsigned long long x_target[4] = {0, 0, 0, 0};
    if (crop && threadIdx.x == 0 && has_fc1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) x_target[i] = (ld_acquire_u64(args.crop.xready + z * 4 + i) / 4ull) * 4ull + 4ull;
    }
    pdl_wait();                                        // crops and person count (fused crop: map and detections) are complete
    if (threadIdx.x == 0) stamp(args, 7);
    const int N = crop ? crop_count_persons(args.crop.pl, lane) : (args.n_dev ? *args.n_dev : args.n_host);
    unsigned short *s_off = reinterpret_cast<unsigned short *>(smem + kOffOffset);
    if (crop && warp >= 2) {
        // exclusive scan of num_boxes -> shared memory (every CTA); the last CTA also writes the flat person list and the
        // offsets for whoever wants them afterwards (decode kernel, person_offsets output, mpn_debug_fetch)
        const PersonList &pl = args.crop.pl;
        const int te = threadIdx.x - 64;
        if (warp == 2) {
            int running = 0;
            for (int b0 = 0; b0 < pl.B; b0 += 32) {
                const int b = b0 + lane;
                const int nbx = (b < pl.B) ? __ldcg(pl.num_boxes + b) : 0;
                int incl = nbx;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += v;
                }
                if (b < pl.B) s_off[b] = (unsigned short)(running + incl - nbx);
                running += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (lane == 0) s_off[pl.B] = (unsigned short)running;
        }
        epi_bar_sync();
        if (c == G - 1) {
            for (int b = te; b <= pl.B; b += kEpiWarps * 32) {
                pl.person_offsets[b] = (int)s_off[b];
                if (pl.person_offsets_out) pl.person_offsets_out[b] = (int)s_off[b];
            }
            if (pl.person_box)
                for (int idx = te; idx < pl.B * pl.max_det; idx += kEpiWarps * 32) {
                    const int b = idx / pl.max_det, k = idx - b * pl.max_det;
                    if (k < (int)s_off[b + 1] - (int)s_off[b]) {
                        const int row = (int)s_off[b] + k;
                        reinterpret_cast<float4 *>(pl.person_box)[row] = __ldcg(reinterpret_cast<const float4 *>(pl.det_boxes) + idx);
                        pl.person_img[row] = b;
                    }
                }
        }
    }
    // bf16 operand of the fused crop: chunks of `cb` k blocks, the same for the four CTAs of a split
    const int nkz = kb1 - kb0;
    const int cb = crop ? max(1, min(kCropChunk, kStgRows / max(1, (N + 3) >> 2))) : 1 << 20;
    const bool run = N > 0 && N <= kPrnFusedMaxRows;   // uniform over the grid; the general kernels take N > 256
    const int nm = (N + 127) >> 7;                   // 128-row M tiles
    const int nb = (N + kXBox - 1) / kXBox;          // 16-row activation boxes
    // ring: <= 128 persons: 4 stages, W slots [0, 128 KB), X slots of 16 KB behind them; else 3 stages, W slots
    // [0, 96 KB), X slots of 32 KB behind them
    const int n_stages = nm == 1 ? 4 : 3;
    const int x_base = n_stages * kWTileBytes, x_stride = nm * kXTileBytes;
    const uint32_t x_bytes = (uint32_t)nb * kXBoxBytes;
    // accumulator stages: round r (fc1 waves, then fc2 tiles) accumulates in stage r % n_acc, at TMEM columns
    // stage * 128 + half * 256; two stages need the persons to fit 128 columns
    const int n_acc = (kWaves > 1 && nm == 1) ? 2 : 1;
    const bool only = crop && args.crop.only != 0;     // development aid: sample and stop
    if (!run || (only && warp < 2)) {
        if (threadIdx.x == 0)
            for (int i = 0; i < early; ++i) { mbar_arrive(full_bar + i); mbar_wait(full_bar + i, 0); }   // drain
    } else

    if (warp == 0) {
        if (lane == 0) {   // ================= TMA producer =================
            int it = 0;
            // fused crop: the activation boxes of a k block are requested when the four CTAs of the split have published
            // its chunk; the weight boxes do not wait for them (xfl = fc1 iterations whose activation boxes are issued,
            // kept with its chunk / position incrementally; xrdy = chunks known to be complete)
            int xfl = 0, xch = 0, xin = 0, xrdy = 0;
            auto x_ready = [&](bool block) -> bool {
                while (xrdy <= xch) {
                    const unsigned long long *cnt = args.crop.xready + z * 4 + xrdy;
                    const unsigned long long target = xrdy == 0 ? x_target[0] : xrdy == 1 ? x_target[1] : xrdy == 2 ? x_target[2] : x_target[3];
                    if (block) grid_wait(cnt, target);
                    else if (!grid_poll(cnt, target)) return false;
                    asm volatile("fence.proxy.async;" ::: "memory");
                    ++xrdy;
                }
                return true;
            };
            auto x_issue = [&]() {
                const int st = xfl % n_stages;
                uint8_t *xs = smem + x_base + st * x_stride;
                for (int b = 0; b < nb; ++b)
                    tma_load_2d(xs + b * kXBoxBytes, &tmap_x, full_bar + st, (kb0 + xfl) * BLOCK_K, b * kXBox, kEvictLast);
                ++xfl;
                if (++xin == cb) { xin = 0; ++xch; }
            };
            for (int w = 0; w < kWaves; ++w) {
                const int hq = w * kTilesPerWave + tq;
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int st = it % n_stages;
                    uint8_t *xs = smem + x_base + st * x_stride;
                    if (crop && xfl <= it - n_stages) {   // the slot's previous user needs its activation boxes first
                        while (xfl <= it - n_stages) { x_ready(true); x_issue(); }
                        while (xfl < it && x_ready(false)) x_issue();
                    }
                    if (it < early) {       // W box already in flight
                        mbar_arrive_expect_tx(full_bar + st, x_bytes);
                    } else {
                        mbar_wait(empty_bar + st, (((uint32_t)(it / n_stages)) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(full_bar + st, x_bytes + kFc1N * 128);
                        tma_load_2d(smem + st * kWTileBytes, &tmap_w1, full_bar + st, kb * BLOCK_K, hq * kFc1N, kEvictFirst);
                    }
                    if (!crop) {
                        for (int b = 0; b < nb; ++b)
                            tma_load_2d(xs + b * kXBoxBytes, &tmap_x, full_bar + st, kb * BLOCK_K, b * kXBox, kEvictLast);
                    } else {
                        while (xfl <= it && x_ready(false)) x_issue();
                    }
                }
            }
            if (crop) while (xfl < it) { x_ready(true); x_issue(); }
            stamp(args, 1);                                                 // all fc1 loads issued
            // fc2: the W2 tiles do not depend on y1 -- run ahead by up to n_stages stages while the grid reduces; k
            // block kb of fc2 needs the y1 columns of wave kb * kWaves / nkb2 only.
            // Measured and not kept (profiles/r01f_summary.md): TMA L2 prefetches of the rest of the W2 tile issued here
            // or after barrier 1, a sliding L2 prefetch window ahead of the ring in both layers, an L2 prefetch of W1
            // from a kernel in the front half of the call, and a per-CTA rotation of the k order.  With W2 L2-resident
            // phase 3 got only 1.1 us shorter, so HBM alone is not what bounds the streaming phases; neither are the
            // activation boxes every CTA re-reads (multicasting them inside CTA pairs: no gain), nor the ring depth, nor
            // the number of W2 boxes on the SM before barrier 2 (profiles/r01g_*.txt).
            int flushed = it;          // fc2 iterations up to here have had their y1 boxes issued; fkb = its k block
            int fkb = 0;               //   (kept incrementally: no divisions on this thread's critical path)
            int ready = 0;             // waves of y1 known to be complete
            auto y1_ready = [&](bool block) -> bool {   // is the y1 k block of iteration `flushed` there?
                const int want = (kWaves > 1 && fkb * kWaves >= args.nkb2) ? 2 : 1;
                while (ready < want) {
                    const unsigned long long *cnt = args.arrivals + (2 * ready + 1) * kBarLine;
                    if (block) grid_wait(cnt, bar_target);
                    else if (ready == 0 || !grid_poll(cnt, bar_target)) return false;   // no polling while running ahead
                    asm volatile("fence.proxy.async;" ::: "memory");
                    ++ready;
                    if (ready == kWaves) stamp(args, 6);                    // producer saw the last y1 barrier
                }
                return true;
            };
            auto y1_issue = [&]() {    // the y1 boxes of iteration `flushed`
                const int st = flushed % n_stages;
                uint8_t *stg = smem + x_base + st * x_stride;
                for (int b = 0; b < nb; ++b)
                    tma_load_2d(stg + b * kXBoxBytes, &tmap_y1, full_bar + st, fkb * BLOCK_K, b * kXBox, kEvictLast);
                ++flushed;
                if (++fkb == args.nkb2) fkb = 0;
            };
            for (int tile = c; tile < args.tiles2; tile += G) {
                for (int kb = 0; kb < args.nkb2; ++kb, ++it) {
                    const int st = it % n_stages;
                    // the slot is released by the MMAs of iteration it - n_stages, which need that iteration's y1 boxes
                    // (and once the wait is over, every W box in the ring gets its y1 boxes at once: the MMAs must not
                    // find the later stages empty-handed when the first one completes)
                    if (flushed <= it - n_stages) {
                        while (flushed <= it - n_stages) { y1_ready(true); y1_issue(); }
                        while (flushed < it && y1_ready(false)) y1_issue();
                    }
                    mbar_wait(empty_bar + st, (((uint32_t)(it / n_stages)) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(full_bar + st, x_bytes + kFc2N * 128);
                    tma_load_2d(smem + st * kWTileBytes, &tmap_w2, full_bar + st, kb * BLOCK_K, tile * kFc2N, kEvictFirst);
                    while (flushed <= it && y1_ready(false)) y1_issue();
                }
            }
            while (flushed < it) { y1_ready(true); y1_issue(); }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {   // ================= MMA issuer =================
            // D[weight row, person] = W tile (A, M = 128, two halves of the 256-row tile) x activations (B, N = 16 nb):
            // the accumulator holds one weight row per TMEM lane and one person per column, so its size, the MMA time
            // (128 N / 256 cycles per instruction) and the TMEM drain all scale with the number of persons.
            const uint32_t idesc = make_idesc_bf16(128, nb * kXBox);
            int it = 0, round = 0;
            auto acc_stage = [&]() -> uint32_t {   // this round's accumulator columns, drained by the epilogue of round - n_acc
                const int stage = round % n_acc;
                if (round >= n_acc) {
                    mbar_wait(tmem_empty_bar + stage, ((uint32_t)(round / n_acc - 1)) & 1u);
                    tc_fence_after();
                }
                return tmem_base + (uint32_t)(stage * kStageCols);
            };
            if (has_fc1) {
                for (int w = 0; w < kWaves; ++w) {
                    const uint32_t acc = acc_stage();
                    for (int kb = kb0; kb < kb1; ++kb, ++it) {
                        const int st = it % n_stages;
                        mbar_wait(full_bar + st, ((uint32_t)(it / n_stages)) & 1u);
                        tc_fence_after();
                        const uint32_t w_addr = smem_u32(smem + st * kWTileBytes);
                        const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smem + x_base + st * x_stride));
#pragma unroll
                        for (int m = 0; m < kFc1N / 128; ++m) {
                            const uint64_t adesc = make_kmajor_sw128_desc(w_addr + m * kWHalfBytes);
#pragma unroll
                            for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                                umma_bf16(acc + (uint32_t)(m * kAccCols), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k),
                                          idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                        umma_commit(empty_bar + st);
                    }
                    umma_commit(tmem_full_bar + round % n_acc);
                    ++round;
                    stamp(args, w == 0 ? 2 : 14);                           // all MMAs of the wave issued
                }
            }
            for (int tile = c; tile < args.tiles2; tile += G) {
                const uint32_t acc = acc_stage();
                for (int kb = 0; kb < args.nkb2; ++kb, ++it) {
                    const int st = it % n_stages;
                    mbar_wait(full_bar + st, ((uint32_t)(it / n_stages)) & 1u);
                    tc_fence_after();
                    // rows 240..255 of the weight slot are not part of this tile (whatever the slot held before):
                    // they only reach accumulator lanes 112..127 of the second half, which nobody reads
                    const uint32_t w_addr = smem_u32(smem + st * kWTileBytes);
                    const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smem + x_base + st * x_stride));
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        const uint64_t adesc = make_kmajor_sw128_desc(w_addr + m * kWHalfBytes);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_bf16(acc + (uint32_t)(m * kAccCols), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k),
                                      idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar + st);
                }
                umma_commit(tmem_full_bar + round % n_acc);
                ++round;
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue / reduce warps =================
        // 16 warps: TMEM lane quadrant q = warp % 4 (hardware rule); the four warps of a quadrant share its work items.
        // A work item is (accumulator half m, 16-person chunk ch): 32 lanes = 32 consecutive weight rows (= 32 consecutive
        // hidden units / output columns, i.e. one 128-byte row segment per person in global memory), 16 columns = 16
        // persons.  No transpose: a warp-wide store of r[j] IS the coalesced row segment of person j.
        const int ew = warp - 2, q = warp & 3, cg = ew >> 2;
        const int tid_e = threadIdx.x - 64;
        float *stg = reinterpret_cast<float *>(smem + kStgOffset + ew * 2048);
        const int n_items = 2 * nb;
        int round = 0;
        if (crop && has_fc1) {
            // ---- fused crop, bf16 operand: this CTA's quarter of the persons (rows tq, tq + 4, ...) over the columns of
            // its K split, a chunk of cb k blocks at a time: sample into the staging area, write out as 16-byte vectors,
            // publish on the split's chunk counter
            const int nq = (N + kTilesPerWave - 1) / kTilesPerWave;
            __nv_bfloat16 *stg16 = reinterpret_cast<__nv_bfloat16 *>(smem + kStgOffset);
            const int rowlen = cb * BLOCK_K;
            for (int ch = 0, r0 = 0; r0 < nkz; ++ch, r0 += cb) {
                const int kA = (kb0 + r0) * BLOCK_K, kB = min((kb0 + min(nkz, r0 + cb)) * BLOCK_K, args.D);
                crop_sample_rows<__nv_bfloat16>(args.crop, s_off, N, tq, kTilesPerWave, nq, kA, kB, stg16, rowlen, tid_e);
                epi_bar_sync();
                const int vec_row = (kB - kA) >> 3;                         // 16-byte vectors per row
                for (int v = tid_e; v < nq * vec_row; v += kEpiWarps * 32) {
                    const int pi = v / vec_row, off = v - pi * vec_row, row = tq + pi * kTilesPerWave;
                    if (row < N)
                        __stcg(reinterpret_cast<uint4 *>(args.crop.x_bf16 + (size_t)row * args.D + kA) + off,
                               reinterpret_cast<const uint4 *>(stg16 + (size_t)pi * rowlen)[off]);
                }
                epi_bar_sync();
                if (tid_e == 0) {
                    grid_arrive(args.crop.xready + z * 4 + ch);
                    stamp(args, ch == 0 ? 11 : 12);                         // first / latest chunk published
                }
            }
        }
        for (int w = 0; w < (only ? 0 : kWaves); ++w) {
            const int col0 = w * (args.hidden / kWaves);                    // first hidden unit of the wave
            if (has_fc1) {   // ---- fc1 partial sums: partial[z][person][hq*256 + m*128 + q*32 + lane]
                const int stage = round % n_acc, hq = w * kTilesPerWave + tq;
                const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(stage * kStageCols);
                mbar_wait(tmem_full_bar + stage, ((uint32_t)(round / n_acc)) & 1u);
                tc_fence_after();
                if (tid_e == 0) stamp(args, w == 0 ? 3 : 13);               // the wave's accumulators are complete
                for (int item = cg; item < (kFc1N / 128) * nb; item += 4) {
                    const int m = item >= nb ? 1 : 0, ch = item - m * nb;
                    uint32_t r[16];
                    tmem_ld16(t_lane + (uint32_t)(m * kAccCols + ch * 16), r);
                    float *dst = args.partial + (kRed ? (size_t)0 : (size_t)z * args.split_stride) +
                                 (size_t)(ch * 16) * args.hidden + hq * kFc1N + m * 128 + q * 32 + lane;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (ch * 16 + j < N) {
                            if (kRed) atomicAdd(dst + (size_t)j * args.hidden, __uint_as_float(r[j]));   // result unused: RED
                            else __stcg(dst + (size_t)j * args.hidden, __uint_as_float(r[j]));
                        }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar + stage);
                ++round;
            }
            epi_bar_sync();
            if (tid_e == 0) {
                stamp(args, w == 0 ? 4 : 11);                               // partial sums stored
                grid_arrive(args.arrivals + (2 * w) * kBarLine);
                grid_wait(args.arrivals + (2 * w) * kBarLine, bar_target);
                stamp(args, w == 0 ? 5 : 12);                               // every CTA's partial sums of the wave are there
            }
            epi_bar_sync();
            {   // ---- y1 = relu(sum_z partial + b1) in bf16: this CTA's slice of the wave's N x (hidden / kWaves) outputs,
                // float4 at a time.  kRedThreads threads per output vector (the last lanes of a warp idle), each with its
                // <= 14 partial loads in flight at once (one L2 round trip), then a fixed-order combine ((s0 + s1) + s2)
                // ...: deterministic, no atomics on data, and the same association whatever the person count.
                const int vec_per_row = (args.hidden / kWaves) >> 2;
                const int total = N * vec_per_row;
                const int per = (total + G - 1) / G;
                const int v_end = min(total, (c + 1) * per);
                if (kRed) {
                    // the sums are complete in the accumulator: read, clear (for the next launch), bias, ReLU, bf16
                    for (int v = c * per + tid_e; v < v_end; v += kEpiWarps * 32) {
                        const int row = v / vec_per_row, c4 = v - row * vec_per_row;
                        float4 *src = reinterpret_cast<float4 *>(args.partial + (size_t)row * args.hidden + col0 + c4 * 4);
                        float4 tot = __ldcg(src);
                        __stcg(src, make_float4(0.f, 0.f, 0.f, 0.f));
                        const float4 bb = __ldg(reinterpret_cast<const float4 *>(args.b1 + col0 + c4 * 4));
                        tot.x = fmaxf(__fadd_rn(tot.x, bb.x), 0.0f);
                        tot.y = fmaxf(__fadd_rn(tot.y, bb.y), 0.0f);
                        tot.z = fmaxf(__fadd_rn(tot.z, bb.z), 0.0f);
                        tot.w = fmaxf(__fadd_rn(tot.w, bb.w), 0.0f);
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(tot.x, tot.y), hi = __floats2bfloat162_rn(tot.z, tot.w);
                        uint2 o;
                        o.x = *reinterpret_cast<const unsigned *>(&lo);
                        o.y = *reinterpret_cast<const unsigned *>(&hi);
                        __stcg(reinterpret_cast<uint2 *>(args.y1 + (size_t)row * args.hidden + col0 + c4 * 4), o);
                    }
                } else {
                const int s_part = (args.splits + kRedThreads - 1) / kRedThreads;   // <= kRedLoads
                const int grp = lane / kRedThreads, part = lane - kRedThreads * grp;
                const int s_lo = part * s_part, s_hi = min(args.splits, s_lo + s_part);
                for (int v0 = c * per; v0 < v_end; v0 += kEpiWarps * kRedGroups) {
                    const int v = v0 + ew * kRedGroups + grp;
                    const bool live = lane < kRedGroups * kRedThreads && v < v_end;
                    const int row = live ? v / vec_per_row : 0, c4 = live ? v - row * vec_per_row : 0;
                    const float *src = args.partial + (size_t)s_lo * args.split_stride + (size_t)row * args.hidden + col0 + c4 * 4;
                    float4 pv[kRedLoads];
#pragma unroll
                    for (int i = 0; i < kRedLoads; ++i)
                        pv[i] = (live && s_lo + i < s_hi) ? __ldcg(reinterpret_cast<const float4 *>(src + (size_t)i * args.split_stride))
                                                          : make_float4(0.f, 0.f, 0.f, 0.f);
                    float4 acc = pv[0];
#pragma unroll
                    for (int i = 1; i < kRedLoads; ++i) {
                        if (s_lo + i < s_hi) {
                            acc.x = __fadd_rn(acc.x, pv[i].x); acc.y = __fadd_rn(acc.y, pv[i].y);
                            acc.z = __fadd_rn(acc.z, pv[i].z); acc.w = __fadd_rn(acc.w, pv[i].w);
                        }
                    }
                    float4 tot = acc;                                       // + the sums of lanes g*kRedThreads + 1, + 2, ...
#pragma unroll
                    for (int k = 1; k < kRedThreads; ++k) {
                        tot.x = __fadd_rn(tot.x, __shfl_down_sync(0xffffffffu, acc.x, k));
                        tot.y = __fadd_rn(tot.y, __shfl_down_sync(0xffffffffu, acc.y, k));
                        tot.z = __fadd_rn(tot.z, __shfl_down_sync(0xffffffffu, acc.z, k));
                        tot.w = __fadd_rn(tot.w, __shfl_down_sync(0xffffffffu, acc.w, k));
                    }
                    if (live && part == 0) {
                        const float4 bb = __ldg(reinterpret_cast<const float4 *>(args.b1 + col0 + c4 * 4));
                        tot.x = fmaxf(__fadd_rn(tot.x, bb.x), 0.0f);
                        tot.y = fmaxf(__fadd_rn(tot.y, bb.y), 0.0f);
                        tot.z = fmaxf(__fadd_rn(tot.z, bb.z), 0.0f);
                        tot.w = fmaxf(__fadd_rn(tot.w, bb.w), 0.0f);
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(tot.x, tot.y), hi = __floats2bfloat162_rn(tot.z, tot.w);
                        uint2 o;
                        o.x = *reinterpret_cast<const unsigned *>(&lo);
                        o.y = *reinterpret_cast<const unsigned *>(&hi);
                        __stcg(reinterpret_cast<uint2 *>(args.y1 + (size_t)row * args.hidden + col0 + c4 * 4), o);
                    }
                }
                }
            }
            epi_bar_sync();
            if (tid_e == 0) {
                stamp(args, w == 0 ? 8 : 15);                               // y1 slice of the wave stored
                grid_arrive(args.arrivals + (2 * w + 1) * kBarLine);
            }
        }
        // ---- fc2 epilogue: logits = x + relu(acc + b2)   (detector/prn.py:22,24)
        // Lane = output column n0 + m*128 + q*32 + lane (its bias sits in a register), columns of the item = 16 persons.
        // In place: relu(acc + b2) goes to the warp's 16 x 32 staging box and one thread issues the TMA reduce-add into
        // x; the last 32-lane group of a 240-column tile has 16 live columns and uses the 16-wide box.
        for (int tile = c; tile < args.tiles2; tile += G) {
            const int n0 = tile * kFc2N;
            float bias[2];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int tc = m * 128 + q * 32 + lane;
                bias[m] = (tc < kFc2N && n0 + tc < args.D) ? __ldg(args.b2 + n0 + tc) : 0.0f;
            }
            if (crop) {
                // ---- fused crop, fp32 residual: the tile's columns of every person, sampled while the tile's weights
                // stream, staged and stored as 16-byte vectors where the reduce-adds below will add relu(acc + b2)
                float *stgf = reinterpret_cast<float *>(smem + kStgOffset);
                const int kA = n0, kB = min(n0 + kFc2N, args.D);
                constexpr int kPass = (kEpiWarps * 2048) / (kFc2N * 4);     // 34 persons fit the staging area
                if (tile != c) {                                            // an earlier tile's boxes may still be read
                    if (lane == 0) tma_store_wait_all();
                    __syncwarp();
                    epi_bar_sync();
                }
                for (int r0 = 0; r0 < N; r0 += kPass) {
                    const int rows = min(kPass, N - r0);
                    crop_sample_rows<float>(args.crop, s_off, N, r0, 1, rows, kA, kB, stgf, kFc2N, tid_e);
                    epi_bar_sync();
                    const int vec_row = (kB - kA) >> 2;
                    for (int v = tid_e; v < rows * vec_row; v += kEpiWarps * 32) {
                        const int pi = v / vec_row, off = v - pi * vec_row;
                        __stcg(reinterpret_cast<float4 *>(args.logits + (size_t)(r0 + pi) * args.D + kA) + off,
                               reinterpret_cast<const float4 *>(stgf + (size_t)pi * kFc2N)[off]);
                    }
                    asm volatile("fence.proxy.async;" ::: "memory");        // generic stores before the async-proxy reduce-adds
                    epi_bar_sync();
                }
                if (tid_e == 0) stamp(args, 13);                            // residual of the tile stored
                if (only) continue;
            }
            const int stage = round % n_acc;
            const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(stage * kStageCols);
            mbar_wait(tmem_full_bar + stage, ((uint32_t)(round / n_acc)) & 1u);
            tc_fence_after();
            if (tid_e == 0) stamp(args, 9);                                 // fc2 accumulators complete
            for (int item = cg; item < n_items; item += 4) {
                const int m = item >= nb ? 1 : 0, ch = item - m * nb;
                const int tc0 = m * 128 + q * 32;                           // first tile column of this warp's lanes
                if (tc0 >= kFc2N || n0 + tc0 >= args.D) continue;           // warp-uniform
                const bool half = tc0 + 32 > kFc2N;                         // 16 live columns
                const float b = m ? bias[1] : bias[0];
                uint32_t r[16];
                tmem_ld16(t_lane + (uint32_t)(m * kAccCols + ch * 16), r);
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = fmaxf(__fadd_rn(__uint_as_float(r[j]), b), 0.0f);
                if (kInPlace) {
                    if (lane == 0) tma_store_wait_read();                   // the previous item has left the buffer
                    __syncwarp();
                    if (!half) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) stg[j * 32 + lane] = v[j];
                    } else if (lane < 16) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) stg[j * 16 + lane] = v[j];
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    // persons N .. of the last 16-row box receive meaningless sums; nothing ever reads them (rows past
                    // the end of the buffer and columns past D are clipped by the tensor map)
                    if (lane == 0) tma_reduce_add_2d(half ? &tmap_out16 : &tmap_out, stg, n0 + tc0, ch * 16);
                } else if (tc0 + lane < kFc2N && n0 + tc0 + lane < args.D) {
                    const size_t off = (size_t)(ch * 16) * args.D + n0 + tc0 + lane;
                    float xr[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) xr[j] = ch * 16 + j < N ? __ldcg(args.x + off + (size_t)j * args.D) : 0.0f;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (ch * 16 + j < N) __stcs(args.logits + off + (size_t)j * args.D, __fadd_rn(xr[j], v[j]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar + stage);
            ++round;
        }
        if (kInPlace && lane == 0) tma_store_wait_all();                    // this thread's reduce-adds have completed
        if (tid_e == 0) stamp(args, 10);                                    // logits stored
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

struct FusedState {
    FusedMaps maps;
    float *partial;
    size_t split_stride;
    unsigned long long *bar;   // grid barrier arrival counters, then the chunk counters of the fused crop ([splits, 4])
    unsigned long long *trace;
    int grid, splits, rows_cap;
    bool red_add;              // MPN_PRN_RED_ADD=1: in-place launches add fc1's K splits up in L2 (kRed)
    float *red_acc;            //   ... in this [rows_cap, hidden] accumulator: zero at creation, cleared by every reader
    const float *out_ptr;      // buffer maps.out describes (handle-owned buffers only)
};

int prn_fused_prepare(mpn_handle *h)
{
    const int D = h->D, Hd = h->cfg.prn_hidden;
    if (Hd != kHq * kFc1N || D % 16 != 0) return MPN_OK;   // shape not covered: the general kernels are used alone
    int dev = h->cfg.device, sms = 0, coop = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop || sms < kTilesPerWave) return MPN_OK;
    if (cudaFuncSetAttribute(prn_fused_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(prn_fused_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(prn_fused_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess ||
        cudaFuncSetAttribute(prn_fused_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, prn_fused_kernel<true, true>, kThreads, kSmemBytes) != cudaSuccess ||
        per_sm < 1) {
        cudaGetLastError();
        snprintf(h->err, sizeof(h->err), "prn_fused_kernel cannot be made resident (smem %d)", kSmemBytes);
        return MPN_ERR_CUDA;
    }
    FusedState *st = new FusedState;
    memset(st, 0, sizeof(*st));
    st->grid = sms;
    st->splits = sms / kTilesPerWave;
    const int nkb1 = (D + BLOCK_K - 1) / BLOCK_K;
    if (st->splits > nkb1) st->splits = nkb1;
    if (st->splits > kRedThreads * kRedLoads) st->splits = kRedThreads * kRedLoads;   // partial loads one reduce thread keeps in flight
    st->rows_cap = h->prn_ws.n_max < kPrnFusedMaxRows ? h->prn_ws.n_max : kPrnFusedMaxRows;
    {
        const char *ra = getenv("MPN_PRN_RED_ADD");
        st->red_add = ra && ra[0] == '1' && kWaves == 1;
    }
    st->split_stride = (size_t)st->rows_cap * Hd;
    const uint64_t rows = (uint64_t)h->prn_ws.n_max;
    bool ok = cudaMalloc(&st->partial, (size_t)st->splits * st->split_stride * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&st->red_acc, st->split_stride * sizeof(float)) == cudaSuccess &&
              cudaMemset(st->red_acc, 0, st->split_stride * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&st->bar, (2 * kWaves * kBarLine + 4 * st->splits) * sizeof(unsigned long long)) == cudaSuccess &&
              cudaMemset(st->bar, 0, (2 * kWaves * kBarLine + 4 * st->splits) * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && encode_2d(&st->maps.x, h->crops_bf16, rows, (uint64_t)D, kXBox) &&
         encode_2d(&st->maps.w1, h->W1t, (uint64_t)Hd, (uint64_t)D, kFc1N) &&
         encode_2d(&st->maps.y1, h->prn_ws.y1_bf16, rows, (uint64_t)Hd, kXBox) &&
         encode_2d(&st->maps.w2, h->W2t, (uint64_t)D, (uint64_t)Hd, kFc2N);
    if (!ok) {
        cudaGetLastError();
        if (st->partial) cudaFree(st->partial);
        if (st->red_acc) cudaFree(st->red_acc);
        if (st->bar) cudaFree(st->bar);
        delete st;
        snprintf(h->err, sizeof(h->err), "fused PRN setup failed (allocation or cuTensorMapEncodeTiled)");
        return MPN_ERR_CUDA;
    }
    h->fused = st;
    return MPN_OK;
}

void prn_fused_release(mpn_handle *h)
{
    FusedState *st = static_cast<FusedState *>(h->fused);
    if (!st) return;
    cudaFree(st->partial);
    cudaFree(st->red_acc);
    cudaFree(st->bar);
    if (st->trace) cudaFree(st->trace);
    delete st;
    h->fused = nullptr;
}

// Development aid: per-CTA globaltimer stamps of the phases of the most recent fused launch (16 slots per CTA).
int prn_fused_trace(mpn_handle *h, int enable, unsigned long long *host_out, int capacity, int *grid_out)
{
    FusedState *st = static_cast<FusedState *>(h->fused);
    if (!st) return MPN_ERR_UNSUPPORTED;
    if (enable && !st->trace) {
        if (cudaMalloc(&st->trace, (size_t)st->grid * 16 * sizeof(unsigned long long)) != cudaSuccess) return MPN_ERR_CUDA;
        cudaMemset(st->trace, 0, (size_t)st->grid * 16 * sizeof(unsigned long long));
    }
    if (host_out && st->trace) {
        const int n = capacity < st->grid * 16 ? capacity : st->grid * 16;
        if (cudaMemcpy(host_out, st->trace, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
            return MPN_ERR_CUDA;
    }
    if (grid_out) *grid_out = st->grid;
    if (!enable && st->trace) { cudaFree(st->trace); st->trace = nullptr; }
    return MPN_OK;
}

// Can this call's crop_and_resize run inside the kernel (CropFuse)?  One fc1 wave, the 56 x 36 x 17 crop, at most 256 images.
bool prn_fused_can_crop(const mpn_handle *h, int batch)
{
    return h->fused != nullptr && kWaves == 1 && h->cfg.crop_height == kCropH && h->cfg.crop_width == kCropW &&
           h->cfg.num_keypoints == kCropCh && batch <= 256 && h->crops_bf16 != nullptr;
}

int launch_prn_fused(mpn_handle *h, const float *x_f32, const int *n_dev, int n_host, int n_max, float *logits, cudaStream_t s,
                     const FusedCropCall *fc)
{
    FusedState *st = static_cast<FusedState *>(h->fused);
    if (!st) return -(int)cudaErrorInvalidValue;
    FusedArgs a;
    memset(&a, 0, sizeof(a));
    if (fc) {
        if (x_f32 != logits || !prn_fused_can_crop(h, fc->pl.B)) return -(int)cudaErrorInvalidValue;
        a.crop.nh = fc->nh; a.crop.hh = fc->hh; a.crop.ww = fc->ww; a.crop.pl = fc->pl;
        a.crop.x_bf16 = h->crops_bf16;
        a.crop.xready = st->bar + 2 * kWaves * kBarLine;
        a.crop.only = fc->only;
    }
    a.n_dev = n_dev; a.n_host = n_host;
    a.early_boxes = n_max <= kPrnFusedMaxRows ? kEarly : 0;
    a.D = h->D; a.hidden = h->cfg.prn_hidden;
    a.nkb1 = (h->D + BLOCK_K - 1) / BLOCK_K;
    a.nkb2 = h->cfg.prn_hidden / BLOCK_K;
    a.splits = st->splits;
    a.tiles2 = (h->D + kFc2N - 1) / kFc2N;
    a.partial = st->partial; a.split_stride = st->split_stride;
    a.b1 = h->b1; a.y1 = h->prn_ws.y1_bf16; a.b2 = h->b2; a.x = x_f32; a.logits = logits;
    a.arrivals = st->bar;
    a.trace = st->trace;
    prof_mark(s, "prn_fused");
    // Cooperative launch (co-residency of the grid is checked by the driver).  With programmatic dependent launch the
    // CTAs become resident one by one while the crop kernel drains, which is safe for the grid barrier because the crop
    // kernel never waits for this one and nothing else runs at that point of the call.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(st->grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = (size_t)kSmemBytes; cfg.stream = s;
    cudaLaunchAttribute attr[2];
    int na = 0;
    attr[na].id = cudaLaunchAttributeCooperative; attr[na].val.cooperative = 1; ++na;
    if (g_pdl) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    const bool in_place = x_f32 == logits;
    if (in_place && st->red_add && !fc) a.partial = st->red_acc;
    if (in_place && st->out_ptr != logits) {
        const uint64_t rows = n_dev ? (uint64_t)h->prn_ws.n_max : (uint64_t)n_host;    // mpn_prn: the caller's buffer has n_host rows
        if (!encode_2d_f32_box(&st->maps.out, logits, rows, (uint64_t)h->D, 32, kXBox) ||
            !encode_2d_f32_box(&st->maps.out16, logits, rows, (uint64_t)h->D, 16, kXBox))
            return -(int)cudaErrorInvalidValue;
        st->out_ptr = n_dev ? logits : nullptr;        // a caller's buffer may change size between calls: never cached
    }
    cudaError_t e = fc ? cudaLaunchKernelEx(&cfg, prn_fused_kernel<true, true>, st->maps.x, st->maps.w1, st->maps.y1, st->maps.w2,
                                            st->maps.out, st->maps.out16, a)
                  : (in_place && st->red_add)
                      ? cudaLaunchKernelEx(&cfg, prn_fused_kernel<true, false, true>, st->maps.x, st->maps.w1, st->maps.y1,
                                           st->maps.w2, st->maps.out, st->maps.out16, a)
                  : in_place ? cudaLaunchKernelEx(&cfg, prn_fused_kernel<true, false>, st->maps.x, st->maps.w1, st->maps.y1,
                                                  st->maps.w2, st->maps.out, st->maps.out16, a)
                             : cudaLaunchKernelEx(&cfg, prn_fused_kernel<false, false>, st->maps.x, st->maps.w1, st->maps.y1,
                                                  st->maps.w2, st->maps.out, st->maps.out16, a);
    if (e != cudaSuccess) return -(int)e;
    return 1;
}

}  // namespace mpn
