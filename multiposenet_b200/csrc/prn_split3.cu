// Pose Residual Network, fp32 mode, on the tensor cores: the "split" form, 80 persons per launch, <= 240 per call.
//
// Replaces detector/prn.py:15-25 with fp32-ACCURATE arithmetic (north_star: 1e-4 in fp32 mode; measured ~1e-6) at a small
// multiple of the bf16 kernel's time instead of the SIMT kernels' 9x (prn_simt.cu stays as the path for more persons).
//
// Every fp32 number is the exact sum of three bf16 numbers: v = h + m + l with h = bf16(v), m = bf16(v - h),
// l = bf16(v - h - m) (8 + 8 + 8 mantissa bits; both subtractions are exact in fp32).  A product of two such numbers is
//     x w = xh (wh + wm + wl) + xm (wh + wm) + xl wh   (+ three terms below 2^-32 |x w|, dropped)
// and every bf16 x bf16 product is exact in the tensor core's fp32 accumulator.  So a layer becomes ONE pass over the
// weights, stored as their three parts side by side along the reduction ([wh | wm | wl], 421 MB for the two layers), with
// the persons' three parts as three groups of accumulator columns ("variants"):
//     k blocks of the wh third:  B = [xh; xm; xl]  (N = 3 x persons)
//     k blocks of the wm third:  B = [xh; xm]      (N = 2 x persons)
//     k blocks of the wl third:  B = [xh]          (N = 1 x persons)
// -- the same activation boxes are fetched for each third, only the instruction's N and the weight box change -- and the
// epilogues add the variants ((v0 + v1) + v2: the large term first) before bias and ReLU.  Three weight streams of the
// bf16 kernel's size instead of six, and the small terms accumulate apart from the large one.
//
// Structure: prn_fused.cu's (one persistent cooperative kernel, fc1 on hidden quarters x 37 K splits, fp32 partial sums,
// grid barrier, fixed-order reduce, grid barrier, fc2 on 240-column tiles); the reduce writes y1's three parts; out of
// place (logits = x + relu(...), x loaded in fp32).  3 x 80 = 240 of a half's 256 TMEM columns bound the persons of a
// launch; a call launches the kernel once per group of 80 persons (each streams the weights again; still 2.5x faster
// than the SIMT kernels at 240).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "handle.cuh"
#include "tcgen05_utils.cuh"

namespace mpn {

namespace {

using namespace tc;

constexpr int kRows3 = kPrnSplit3GroupRows;              // persons per launch; variant v of person n is row v * kRows3 + n
constexpr int kGroups3 = kPrnSplit3MaxRows / kRows3;     // launches per call: persons [80 g, 80 g + 80), each streams the weights
constexpr int kFc1N = 256, kFc2N = 240;                  // weight rows per CTA tile of the two layers
constexpr int kXBox = 16, kXBoxBytes = kXBox * 128;      // activation boxes: 16 rows of one k block
constexpr int kWTileBytes = 256 * 128, kWHalfBytes = 128 * 128, kXSlotBytes = 256 * 128;
constexpr int kStages = 3;                               // stage = [W tile 32 KB | up to 15 activation boxes, 32 KB]
constexpr int kAccCols = 256;                            // TMEM column of the second accumulator half
constexpr int kEarly = 3;                                // W1 boxes requested before the person count is known
constexpr int kEpiWarps = 16;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kXBase = kStages * kWTileBytes;
constexpr int kBarOffset = kStages * (kWTileBytes + kXSlotBytes);
constexpr int kSmemBytes = kBarOffset + (2 * kStages + 2) * 8 + 16 + 1024;
constexpr int kBarLine = 16;                             // the two grid barrier counters sit on separate 128-byte lines
constexpr uint32_t kTmemCols = 512;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
static_assert(3 * kRows3 <= kAccCols && kRows3 % kXBox == 0, "three variants of every person in one accumulator half");

struct Split3Args {
    const int *n_dev;
    int n_host;
    int row0;                // first person of this launch's group (x and logits already point at it)
    int D, hidden;
    int third1;              // k blocks per third of fc1's reduction (ceil(D / 64); the thirds are padded to that)
    int third2;              // k blocks per third of fc2's reduction (hidden / 64)
    int splits;              // K splits of fc1 (over all three thirds)
    int tiles2;              // output tiles of fc2
    float *partial;          // [splits, kRows3, hidden]
    size_t split_stride;
    const float *b1;
    __nv_bfloat16 *y1;       // [3 * kRows3, hidden]: the three parts of y1
    const float *b2;
    const float *x;          // fp32 crops [N, D] (residual)
    float *logits;           // [N, D]
    unsigned long long *arrivals;   // two never-reset grid barrier counters (grid arrivals per running launch each)
};

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// grid barrier (prn_fused.cu): one thread per CTA arrives after a CTA-wide barrier, waiters poll with relaxed loads
__device__ __forceinline__ void grid_arrive(unsigned long long *arrivals)
{
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    atomicAdd(arrivals, 1ULL);
}
__device__ __forceinline__ void grid_wait(const unsigned long long *arrivals, unsigned long long target)
{
    for (unsigned spin = 0; spin < (1u << 24); ++spin) {
        unsigned long long v;
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(arrivals) : "memory");
        if (v >= target) {
            asm volatile("fence.acq_rel.gpu;" ::: "memory");
            return;
        }
        __nanosleep(64);
    }
    __trap();
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

// v = h + m + l, exactly (for finite v whose low parts do not underflow)
__device__ __forceinline__ void split3(float v, __nv_bfloat16 &h, __nv_bfloat16 &m, __nv_bfloat16 &l)
{
    h = __float2bfloat16_rn(v);
    const float r1 = __fsub_rn(v, __bfloat162float(h));
    m = __float2bfloat16_rn(r1);
    l = __float2bfloat16_rn(__fsub_rn(r1, __bfloat162float(m)));
}

__global__ void __launch_bounds__(kThreads, 1)
prn_split3_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                  const __grid_constant__ CUtensorMap tmap_y1, const __grid_constant__ CUtensorMap tmap_w2,
                  const Split3Args args)
{
    extern __shared__ uint8_t smem_raw[];
    pdl_trigger();
    if (args.row0 > 0) {   // a later group of the call: mostly empty, so look at the count before anything is set up
        pdl_wait();
        const int n = args.n_dev ? *args.n_dev : args.n_host;
        if (!(n > args.row0 && n <= kRows3 * kGroups3)) return;
    }
    const int G = gridDim.x, c = blockIdx.x;
    const unsigned long long bar_target =
        (ld_acquire_u64(args.arrivals) / (unsigned long long)G) * (unsigned long long)G + (unsigned long long)G;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kBarOffset);
    uint64_t *empty_bar = full_bar + kStages;
    uint64_t *tmem_full_bar = empty_bar + kStages;
    uint64_t *tmem_empty_bar = tmem_full_bar + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_y1)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_w2)) : "memory");
        for (int i = 0; i < kStages; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        mbar_init(tmem_full_bar, 1);
        mbar_init(tmem_empty_bar, kEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // fc1: CTA (hq, z) = hidden quarter hq x K split z of the 3 * third1 k blocks of [wh | wm | wl]
    const bool has_fc1 = c < 4 * args.splits;
    const int hq = c & 3, z = c >> 2;
    const int nk1 = 3 * args.third1;
    const int kb0 = has_fc1 ? (int)(((long long)z * nk1) / args.splits) : 0;
    const int kb1 = has_fc1 ? (int)(((long long)(z + 1) * nk1) / args.splits) : 0;
    const int t0 = kb0 / args.third1;                    // the third the split starts in: 3 - t0 variants are live in it

    int early = 0;
    if (threadIdx.x == 0 && has_fc1 && args.row0 == 0) { // weight boxes that do not depend on the crops (first group only:
                                                         // the later groups of a call are mostly empty and exit)
        early = min(kEarly, kb1 - kb0);
        for (int i = 0; i < early; ++i) {
            mbar_expect_tx(full_bar + i, kFc1N * 128);
            tma_load_2d(smem + i * kWTileBytes, &tmap_w1, full_bar + i, (kb0 + i) * BLOCK_K, hq * kFc1N, kEvictFirst);
        }
    }
    pdl_wait();                                          // the parts of the crops and the person count are complete
    const int n_all = args.n_dev ? *args.n_dev : args.n_host;
    const int N = min(kRows3, n_all - args.row0);        // this group's persons
    const bool run = N > 0 && n_all <= kRows3 * kGroups3;   // uniform over the grid; more persons: the SIMT kernels
    const int nb = (N + kXBox - 1) / kXBox;              // 16-row activation boxes per variant
    if (!run) {
        if (threadIdx.x == 0)
            for (int i = 0; i < early; ++i) { mbar_arrive(full_bar + i); mbar_wait(full_bar + i, 0); }   // drain
    } else

    if (warp == 0) {
        if (lane == 0) {   // ================= TMA producer =================
            int it = 0;
            {
                int t = t0, kin = kb0 - t0 * args.third1;
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int st = it % kStages, nact = 3 - t;
                    const uint32_t x_bytes = (uint32_t)(nact * nb) * kXBoxBytes;
                    if (it < early) {
                        mbar_arrive_expect_tx(full_bar + st, x_bytes);
                    } else {
                        mbar_wait(empty_bar + st, (((uint32_t)(it / kStages)) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(full_bar + st, x_bytes + kFc1N * 128);
                        tma_load_2d(smem + st * kWTileBytes, &tmap_w1, full_bar + st, kb * BLOCK_K, hq * kFc1N, kEvictFirst);
                    }
                    uint8_t *xs = smem + kXBase + st * kXSlotBytes;
                    for (int v = 0; v < nact; ++v)
                        for (int b = 0; b < nb; ++b)
                            tma_load_2d(xs + (v * nb + b) * kXBoxBytes, &tmap_x, full_bar + st, kin * BLOCK_K,
                                        v * kRows3 + b * kXBox, kEvictLast);
                    if (++kin == args.third1) { kin = 0; ++t; }
                }
            }
            // fc2: the W2 boxes do not depend on y1 -- run ahead by up to kStages stages while the grid reduces
            int flushed = it, ft = 0, fk = 0;             // fc2 iterations before `flushed` have their y1 boxes; its (third, k)
            bool ready = false;
            auto y1_issue = [&]() {
                const int st = flushed % kStages, nact = 3 - ft;
                uint8_t *xs = smem + kXBase + st * kXSlotBytes;
                for (int v = 0; v < nact; ++v)
                    for (int b = 0; b < nb; ++b)
                        tma_load_2d(xs + (v * nb + b) * kXBoxBytes, &tmap_y1, full_bar + st, fk * BLOCK_K, v * kRows3 + b * kXBox,
                                    kEvictLast);
                ++flushed;
                if (++fk == args.third2) { fk = 0; if (++ft == 3) ft = 0; }
            };
            auto wait_y1 = [&]() {
                if (!ready) {
                    grid_wait(args.arrivals + kBarLine, bar_target);
                    asm volatile("fence.proxy.async;" ::: "memory");
                    ready = true;
                }
            };
            for (int tile = c; tile < args.tiles2; tile += G) {
                int t = 0, kin = 0;
                for (int kb = 0; kb < 3 * args.third2; ++kb, ++it) {
                    const int st = it % kStages;
                    if (flushed <= it - kStages) {
                        while (flushed <= it - kStages) { wait_y1(); y1_issue(); }
                        while (flushed < it) y1_issue();
                    }
                    mbar_wait(empty_bar + st, (((uint32_t)(it / kStages)) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(full_bar + st, (uint32_t)((3 - t) * nb) * kXBoxBytes + kFc2N * 128);
                    tma_load_2d(smem + st * kWTileBytes, &tmap_w2, full_bar + st, kb * BLOCK_K, tile * kFc2N, kEvictFirst);
                    while (ready && flushed <= it) y1_issue();
                    if (++kin == args.third2) { kin = 0; ++t; }
                }
            }
            while (flushed < it) { wait_y1(); y1_issue(); }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {   // ================= MMA issuer =================
            // D[weight row, (variant, person)] = W tile (A, M = 128 twice) x activation parts (B, N = live variants x 16 nb)
            const uint32_t idesc3 = make_idesc_bf16(128, 3 * nb * kXBox), idesc2 = make_idesc_bf16(128, 2 * nb * kXBox),
                           idesc1 = make_idesc_bf16(128, nb * kXBox);
            int it = 0;
            if (has_fc1) {
                int t = t0, kin = kb0 - t0 * args.third1;
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int st = it % kStages;
                    const uint32_t idesc = t == 0 ? idesc3 : t == 1 ? idesc2 : idesc1;
                    mbar_wait(full_bar + st, ((uint32_t)(it / kStages)) & 1u);
                    tc_fence_after();
                    const uint32_t w_addr = smem_u32(smem + st * kWTileBytes);
                    const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smem + kXBase + st * kXSlotBytes));
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        const uint64_t adesc = make_kmajor_sw128_desc(w_addr + m * kWHalfBytes);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_bf16(tmem_base + (uint32_t)(m * kAccCols), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k),
                                      idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar + st);
                    if (++kin == args.third1) { kin = 0; ++t; }
                }
                umma_commit(tmem_full_bar);
            }
            int round = has_fc1 ? 1 : 0;
            for (int tile = c; tile < args.tiles2; tile += G) {
                if (round > 0) {   // the epilogue has drained the previous accumulators
                    mbar_wait(tmem_empty_bar, ((uint32_t)(round - 1)) & 1u);
                    tc_fence_after();
                }
                int t = 0, kin = 0;
                for (int kb = 0; kb < 3 * args.third2; ++kb, ++it) {
                    const int st = it % kStages;
                    const uint32_t idesc = t == 0 ? idesc3 : t == 1 ? idesc2 : idesc1;
                    mbar_wait(full_bar + st, ((uint32_t)(it / kStages)) & 1u);
                    tc_fence_after();
                    const uint32_t w_addr = smem_u32(smem + st * kWTileBytes);
                    const uint64_t bdesc = make_kmajor_sw128_desc(smem_u32(smem + kXBase + st * kXSlotBytes));
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        const uint64_t adesc = make_kmajor_sw128_desc(w_addr + m * kWHalfBytes);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_bf16(tmem_base + (uint32_t)(m * kAccCols), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k),
                                      idesc, (kb > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar + st);
                    if (++kin == args.third2) { kin = 0; ++t; }
                }
                umma_commit(tmem_full_bar);
                ++round;
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue / reduce warps =================
        // TMEM lane quadrant q = warp % 4; the four warps of a quadrant share its (accumulator half, 16-person chunk) items.
        // A lane is a weight row (hidden unit / output column), the 16 columns of an item are 16 persons; variant v of the
        // chunk sits nb chunks further on.
        const int ew = warp - 2, q = warp & 3, cg = ew >> 2;
        const int tid_e = threadIdx.x - 64;
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        int round = 0;
        if (has_fc1) {   // ---- fc1 partial sums: partial[z][person][hq*256 + m*128 + q*32 + lane] = (v0 + v1) + v2
            const int nact0 = 3 - t0;
            mbar_wait(tmem_full_bar, 0);
            tc_fence_after();
            for (int item = cg; item < 2 * nb; item += 4) {
                const int m = item >= nb ? 1 : 0, ch = item - m * nb;
                uint32_t r[16], r2[16];
                tmem_ld16(t_lane + (uint32_t)(m * kAccCols + ch * 16), r);
                for (int v = 1; v < nact0; ++v) {
                    tmem_ld16(t_lane + (uint32_t)(m * kAccCols + (v * nb + ch) * 16), r2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__fadd_rn(__uint_as_float(r[j]), __uint_as_float(r2[j])));
                }
                float *dst = args.partial + (size_t)z * args.split_stride + (size_t)(ch * 16) * args.hidden + hq * kFc1N +
                             m * 128 + q * 32 + lane;
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (ch * 16 + j < N) __stcg(dst + (size_t)j * args.hidden, __uint_as_float(r[j]));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar);
            ++round;
        }
        epi_bar_sync();
        if (tid_e == 0) {
            grid_arrive(args.arrivals);
            grid_wait(args.arrivals, bar_target);
        }
        epi_bar_sync();
        {   // ---- y1 = relu(sum_z partial + b1), fp32, written as its three bf16 parts.  Three threads per float4 of
            // outputs, each with its <= 14 partial loads in flight, fixed-order combine (prn_fused.cu).
            const int vec_per_row = args.hidden >> 2;
            const int total = N * vec_per_row;
            const int per = (total + G - 1) / G;
            const int v_end = min(total, (c + 1) * per);
            const int s_third = (args.splits + 2) / 3;                     // <= 14 (splits <= 42)
            const int grp = lane / 3, part = lane - 3 * grp;
            const int s_lo = part * s_third, s_hi = min(args.splits, s_lo + s_third);
            for (int v0 = c * per; v0 < v_end; v0 += kEpiWarps * 10) {
                const int v = v0 + ew * 10 + grp;
                const bool live = lane < 30 && v < v_end;
                const int row = live ? v / vec_per_row : 0, c4 = live ? v - row * vec_per_row : 0;
                const float *src = args.partial + (size_t)s_lo * args.split_stride + (size_t)row * args.hidden + c4 * 4;
                float4 pv[14];
#pragma unroll
                for (int i = 0; i < 14; ++i)
                    pv[i] = (live && s_lo + i < s_hi) ? __ldcg(reinterpret_cast<const float4 *>(src + (size_t)i * args.split_stride))
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
                float4 acc = pv[0];
#pragma unroll
                for (int i = 1; i < 14; ++i) {
                    if (s_lo + i < s_hi) {
                        acc.x = __fadd_rn(acc.x, pv[i].x); acc.y = __fadd_rn(acc.y, pv[i].y);
                        acc.z = __fadd_rn(acc.z, pv[i].z); acc.w = __fadd_rn(acc.w, pv[i].w);
                    }
                }
                float4 tot = acc;
#pragma unroll
                for (int k = 1; k < 3; ++k) {
                    tot.x = __fadd_rn(tot.x, __shfl_down_sync(0xffffffffu, acc.x, k));
                    tot.y = __fadd_rn(tot.y, __shfl_down_sync(0xffffffffu, acc.y, k));
                    tot.z = __fadd_rn(tot.z, __shfl_down_sync(0xffffffffu, acc.z, k));
                    tot.w = __fadd_rn(tot.w, __shfl_down_sync(0xffffffffu, acc.w, k));
                }
                if (live && part == 0) {
                    const float4 bb = __ldg(reinterpret_cast<const float4 *>(args.b1 + c4 * 4));
                    const float y[4] = {fmaxf(__fadd_rn(tot.x, bb.x), 0.0f), fmaxf(__fadd_rn(tot.y, bb.y), 0.0f),
                                        fmaxf(__fadd_rn(tot.z, bb.z), 0.0f), fmaxf(__fadd_rn(tot.w, bb.w), 0.0f)};
                    __nv_bfloat16 p[3][4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) split3(y[e], p[0][e], p[1][e], p[2][e]);
#pragma unroll
                    for (int pt = 0; pt < 3; ++pt) {
                        uint2 o;
                        o.x = (unsigned)__bfloat16_as_ushort(p[pt][0]) | ((unsigned)__bfloat16_as_ushort(p[pt][1]) << 16);
                        o.y = (unsigned)__bfloat16_as_ushort(p[pt][2]) | ((unsigned)__bfloat16_as_ushort(p[pt][3]) << 16);
                        __stcg(reinterpret_cast<uint2 *>(args.y1 + (size_t)(pt * kRows3 + row) * args.hidden + c4 * 4), o);
                    }
                }
            }
        }
        epi_bar_sync();
        if (tid_e == 0) grid_arrive(args.arrivals + kBarLine);
        // ---- fc2 epilogue: logits = x + relu(((v0 + v1) + v2) + b2)   (detector/prn.py:22,24)
        for (int tile = c; tile < args.tiles2; tile += G) {
            const int n0 = tile * kFc2N;
            float bias[2];
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const int tc = m * 128 + q * 32 + lane;
                bias[m] = (tc < kFc2N && n0 + tc < args.D) ? __ldg(args.b2 + n0 + tc) : 0.0f;
            }
            mbar_wait(tmem_full_bar, ((uint32_t)round) & 1u);
            tc_fence_after();
            for (int item = cg; item < 2 * nb; item += 4) {
                const int m = item >= nb ? 1 : 0, ch = item - m * nb;
                const int tc0 = m * 128 + q * 32;
                if (tc0 >= kFc2N || n0 + tc0 >= args.D) continue;           // warp-uniform
                const float b = m ? bias[1] : bias[0];
                uint32_t r[16], r2[16];
                tmem_ld16(t_lane + (uint32_t)(m * kAccCols + ch * 16), r);
#pragma unroll
                for (int v = 1; v < 3; ++v) {
                    tmem_ld16(t_lane + (uint32_t)(m * kAccCols + (v * nb + ch) * 16), r2);
#pragma unroll
                    for (int j = 0; j < 16; ++j) r[j] = __float_as_uint(__fadd_rn(__uint_as_float(r[j]), __uint_as_float(r2[j])));
                }
                if (tc0 + lane < kFc2N && n0 + tc0 + lane < args.D) {
                    const size_t off = (size_t)(ch * 16) * args.D + n0 + tc0 + lane;
                    float xr[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) xr[j] = ch * 16 + j < N ? __ldcg(args.x + off + (size_t)j * args.D) : 0.0f;
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (ch * 16 + j < N)
                            __stcs(args.logits + off + (size_t)j * args.D,
                                   __fadd_rn(xr[j], fmaxf(__fadd_rn(__uint_as_float(r[j]), b), 0.0f)));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar);
            ++round;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- data preparation --------------------------------------------------------------------------------------------------
// w [rows, cols] fp32 row-major -> wt3 [cols, 3 * third] bf16: wt3[c][t * third + r] = part t of w[r][c] (one-time)
__global__ void __launch_bounds__(256) split3_weights_kernel(const float *__restrict__ w, const int rows, const int cols,
                                                             __nv_bfloat16 *__restrict__ wt3, const int third)
{
    __shared__ float tile[32][33];
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;   // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int r = r0 + i, c = c0 + tx;
        tile[i][tx] = (r < rows && c < cols) ? w[(size_t)r * cols + c] : 0.0f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, r = r0 + tx;
        if (c < cols && r < rows) {
            __nv_bfloat16 h, m, l;
            split3(tile[tx][i], h, m, l);
            __nv_bfloat16 *dst = wt3 + (size_t)c * (3 * third) + r;
            dst[0] = h; dst[third] = m; dst[2 * (size_t)third] = l;
        }
    }
}

// x [N, D] fp32 (already at the group's first person) -> x3 [3 * kRows3, Dpad] bf16 (row v * kRows3 + n = part v of
// person row0 + n); nothing when the call has more persons than the groups cover
__global__ void __launch_bounds__(256) split3_rows_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ x3,
                                                          const int *__restrict__ n_dev, const int n_host, const int row0,
                                                          const int D, const int Dpad)
{
    pdl_trigger();
    pdl_wait();
    const int n_all = n_dev ? *n_dev : n_host;
    if (n_all > kRows3 * kGroups3) return;
    const int row = blockIdx.y;
    if (row >= n_all - row0) return;
    const int d = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (d >= D) return;
    const float4 v = __ldg(reinterpret_cast<const float4 *>(x + (size_t)row * D + d));
    const float f[4] = {v.x, v.y, v.z, v.w};
    __nv_bfloat16 p[3][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split3(f[e], p[0][e], p[1][e], p[2][e]);
#pragma unroll
    for (int pt = 0; pt < 3; ++pt) {
        uint2 o;
        o.x = (unsigned)__bfloat16_as_ushort(p[pt][0]) | ((unsigned)__bfloat16_as_ushort(p[pt][1]) << 16);
        o.y = (unsigned)__bfloat16_as_ushort(p[pt][2]) | ((unsigned)__bfloat16_as_ushort(p[pt][3]) << 16);
        *reinterpret_cast<uint2 *>(x3 + (size_t)(pt * kRows3 + row) * Dpad + d) = o;
    }
}

}  // namespace

struct Split3State {
    CUtensorMap x, w1, y1, w2;
    __nv_bfloat16 *W1t3, *W2t3, *x3, *y1_3;
    float *partial;
    size_t split_stride;
    unsigned long long *bar;
    int grid, splits, third1, third2, Dpad;
};

int prn_split3_prepare(mpn_handle *h)
{
    const int D = h->D, Hd = h->cfg.prn_hidden;
    if (!(h->cfg.prn_modes & 1) || Hd != 4 * kFc1N || D % 16 != 0 || getenv("MPN_FP32_SIMT") != nullptr) return MPN_OK;
    int dev = h->cfg.device, sms = 0, coop = 0, per_sm = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    if (!coop || sms < 4) return MPN_OK;
    if (cudaFuncSetAttribute(prn_split3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess ||
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, prn_split3_kernel, kThreads, kSmemBytes) != cudaSuccess ||
        per_sm < 1) {
        cudaGetLastError();
        snprintf(h->err, sizeof(h->err), "prn_split3_kernel cannot be made resident (smem %d)", kSmemBytes);
        return MPN_ERR_CUDA;
    }
    Split3State *st = new Split3State;
    memset(st, 0, sizeof(*st));
    st->grid = sms;
    st->third1 = (D + BLOCK_K - 1) / BLOCK_K;
    st->third2 = Hd / BLOCK_K;
    st->Dpad = st->third1 * BLOCK_K;
    st->splits = sms / 4;
    if (st->splits > 42) st->splits = 42;                // the reduce keeps <= 14 partial loads per thread in flight
    st->split_stride = (size_t)kRows3 * Hd;
    const size_t w1_elems = (size_t)Hd * 3 * st->Dpad, w2_elems = (size_t)D * 3 * Hd, x3_elems = (size_t)3 * kRows3 * st->Dpad,
                 y1_elems = (size_t)3 * kRows3 * Hd;
    const size_t bar_bytes = 2 * kBarLine * sizeof(unsigned long long);
    bool ok = cudaMalloc(&st->W1t3, w1_elems * 2) == cudaSuccess && cudaMalloc(&st->W2t3, w2_elems * 2) == cudaSuccess &&
              cudaMalloc(&st->x3, x3_elems * 2) == cudaSuccess && cudaMalloc(&st->y1_3, y1_elems * 2) == cudaSuccess &&
              cudaMalloc(&st->partial, (size_t)st->splits * st->split_stride * sizeof(float)) == cudaSuccess &&
              cudaMalloc(&st->bar, bar_bytes) == cudaSuccess &&
              // the padding columns of the thirds, the rows of absent persons: zeros, written once
              cudaMemset(st->W1t3, 0, w1_elems * 2) == cudaSuccess && cudaMemset(st->x3, 0, x3_elems * 2) == cudaSuccess &&
              cudaMemset(st->y1_3, 0, y1_elems * 2) == cudaSuccess && cudaMemset(st->bar, 0, bar_bytes) == cudaSuccess;
    ok = ok && encode_2d(&st->x, st->x3, (uint64_t)3 * kRows3, (uint64_t)st->Dpad, kXBox) &&
         encode_2d(&st->w1, st->W1t3, (uint64_t)Hd, (uint64_t)3 * st->Dpad, kFc1N) &&
         encode_2d(&st->y1, st->y1_3, (uint64_t)3 * kRows3, (uint64_t)Hd, kXBox) &&
         encode_2d(&st->w2, st->W2t3, (uint64_t)D, (uint64_t)3 * Hd, kFc2N);
    if (!ok) {
        cudaGetLastError();
        cudaFree(st->W1t3); cudaFree(st->W2t3); cudaFree(st->x3); cudaFree(st->y1_3); cudaFree(st->partial); cudaFree(st->bar);
        delete st;
        snprintf(h->err, sizeof(h->err), "split-bf16 PRN setup failed (allocation or cuTensorMapEncodeTiled)");
        return MPN_ERR_CUDA;
    }
    h->split3 = st;
    return MPN_OK;
}

void prn_split3_release(mpn_handle *h)
{
    Split3State *st = static_cast<Split3State *>(h->split3);
    if (!st) return;
    cudaFree(st->W1t3); cudaFree(st->W2t3); cudaFree(st->x3); cudaFree(st->y1_3); cudaFree(st->partial); cudaFree(st->bar);
    delete st;
    h->split3 = nullptr;
}

// W1 [D, hidden], W2 [hidden, D] fp32 on the device -> the three-part K-major operands
int prn_split3_set_weights(mpn_handle *h, const float *dW1, const float *dW2, cudaStream_t s)
{
    Split3State *st = static_cast<Split3State *>(h->split3);
    if (!st) return 0;
    const int D = h->D, Hd = h->cfg.prn_hidden;
    split3_weights_kernel<<<dim3((Hd + 31) / 32, (D + 31) / 32), 256, 0, s>>>(dW1, D, Hd, st->W1t3, st->Dpad);
    split3_weights_kernel<<<dim3((D + 31) / 32, (Hd + 31) / 32), 256, 0, s>>>(dW2, Hd, D, st->W2t3, Hd);
    return 2;
}

int launch_prn_split3(mpn_handle *h, const float *x_f32, const int *n_dev, int n_host, int n_max, float *logits, cudaStream_t s)
{
    Split3State *st = static_cast<Split3State *>(h->split3);
    if (!st || n_max <= 0) return 0;
    const int D = h->D;
    const int groups = n_max >= kRows3 * kGroups3 ? kGroups3 : (n_max + kRows3 - 1) / kRows3;
    int launches = 0;
    for (int g = 0; g < groups; ++g) {
        const int row0 = g * kRows3;
        const int rows = n_max - row0 < kRows3 ? n_max - row0 : kRows3;
        prof_mark(s, "prn_split3_parts");
        launch_k(split3_rows_kernel, dim3((D / 4 + 255) / 256, rows), dim3(256), 0, s, true, x_f32 + (size_t)row0 * D, st->x3, n_dev,
                 n_host, row0, D, st->Dpad);
        Split3Args a;
        memset(&a, 0, sizeof(a));
        a.n_dev = n_dev; a.n_host = n_host; a.row0 = row0;
        a.D = D; a.hidden = h->cfg.prn_hidden;
        a.third1 = st->third1; a.third2 = st->third2;
        a.splits = st->splits;
        a.tiles2 = (D + kFc2N - 1) / kFc2N;
        a.partial = st->partial; a.split_stride = st->split_stride;
        a.b1 = h->b1; a.y1 = st->y1_3; a.b2 = h->b2;
        a.x = x_f32 + (size_t)row0 * D; a.logits = logits + (size_t)row0 * D;
        a.arrivals = st->bar;
        prof_mark(s, "prn_split3");
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
        cudaError_t e = cudaLaunchKernelEx(&cfg, prn_split3_kernel, st->x, st->w1, st->y1, st->w2, a);
        if (e != cudaSuccess) return -(int)e;
        launches += 2;
    }
    return launches;
}

}  // namespace mpn
