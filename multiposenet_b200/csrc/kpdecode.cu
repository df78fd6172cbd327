// Spatial softmax + argmax keypoint decode.
//
// Replaces create_pb.py:115-142: logits [N, 56*36, 17] -> softmax over the 2016 positions of every (person, channel)
// (tf.nn.softmax(axis=1): p = exp(l - max) * (1 / sum)), keypoint_scores = max p, keypoint_positions =
// (argmax // 36 / 56, argmax % 36 / 36) with tf.argmax's first-index tie rule.
//
// One thread-block CLUSTER of 4 CTAs per person, each CTA of 17 x 32 threads owning a quarter of the 2016 positions (a
// contiguous 34 KB slab of the person's logit row, read exactly once, fully coalesced; thread (channel c, lane q) sees
// positions q, q + 32, ...: 16 values).
//
// argmax rule, bit-matched to the oracle without depending on the summation order of the denominator:
// p[i] = e[i] * r with r = 1/sum and e[i] = exp(l[i] - G) <= 1, G the channel's maximum.  e[i] * r == r iff e[i] == 1.0f
// (for e[i] <= 1 - 2^-24 the product is at least half an ulp below r and rounds to a smaller float), so the first index
// with maximal probability is the first index whose e[i] is exactly 1.0f.  The exp recipe is monotone near 0, so
// "exp(d) == 1.0f" is the comparison d >= x0 with x0 the most negative float whose recipe value is 1.0f (found once per
// device by bisection with the recipe itself, kpdecode_prepare).
//
// ONE pass, ONE block barrier and ONE cluster barrier per person (an earlier version took the maximum first -- block
// reduce, cluster barrier -- then the sum -- block reduce, cluster barrier -- then a third cluster barrier to keep the
// peers alive: a serial chain of ~5 us per person that left HBM idle, 1.8 TB/s at 2801 persons).  Every thread folds its
// 16 values into a partial (m, s = sum exp(l - m), f = first position with l - m >= x0); partials merge like an online
// softmax: M = max m, S = sum s * exp(m - M), F = min f over the partials with m == M.  A partial with m < M can still
// hold positions that satisfy the GLOBAL rule l - M >= x0 when m - M >= x0 (two different floats within half an ulp of
// 1.0 of each other: only possible for |M| < 0.5); its own f was taken against m, not M, so such a "near tie" is only
// flagged and the flagged channel is rescanned exactly against M by CTA 0 (never seen on real logits, covered by a
// test).  keypoint_scores = 1 / S is held to 1e-4 only, so the terms use the hardware ex2 approximation.
//
// The peers write their per-channel partials straight into CTA 0's shared memory (distributed shared memory), then the
// cluster barrier, then CTA 0 merges and stores.  keypoint_decode_stream_kernel: PERSISTENT clusters that walk the persons
// with every CTA's slab arriving by one bulk copy (TMA) into a two-slot shared-memory ring, two CTAs per SM, so that the
// next person's bytes are in flight while the current one is reduced.  (A second kernel with one cluster per person and
// the logits loaded straight into registers was measured slower at every person count and is gone; every crop size the
// handle accepts -- 17 * positions a multiple of 32 -- satisfies the bulk copy's 16-byte granularity.)
#include <cstdlib>

#include "bulk_copy.cuh"
#include "common.cuh"
#include "mpn_math.cuh"

namespace mpn {

namespace {

constexpr int kNK = 17;
constexpr int kLanes = 32;
constexpr int kThreads = kNK * kLanes;   // 544
constexpr int kCluster = 4;
constexpr int kMaxPerThread = 16;        // positions per thread: ceil(2048 / 4 / 32)
constexpr int kSlots = 2;                // slabs per CTA of the streaming kernel (two CTAs per SM: four slabs per SM)
constexpr int kStreamSlabBytes = kSlots * (kMaxPerThread * kLanes) * kNK * 4;   // upper bound: slabs of 512 positions
constexpr int kNone = 0x7fffffff;

__device__ float g_exp_one_x0;     // most negative x with exact_expf(x) == 1.0f

__global__ void exp_one_threshold_kernel()
{
    // negative floats are ordered by their magnitude bits: bisect the largest magnitude that still gives exactly 1.0f
    unsigned lo = 0x80000000u, hi = 0xBF800000u;      // -0.0f (gives 1) .. -1.0f (does not)
    while (hi - lo > 1u) {
        const unsigned mid = lo + (hi - lo) / 2u;
        if (exact_expf(__uint_as_float(mid)) == 1.0f) lo = mid; else hi = mid;
    }
    g_exp_one_x0 = __uint_as_float(lo);
}

__device__ __forceinline__ unsigned cluster_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// address of `p` (a shared-memory object of this CTA) in CTA `rank` of the cluster
template <typename T>
__device__ __forceinline__ T *peer_shared(T *p, unsigned rank)
{
    unsigned long long out;
    asm volatile("mapa.u64 %0, %1, %2;" : "=l"(out) : "l"(reinterpret_cast<unsigned long long>(p)), "r"(rank));
    return reinterpret_cast<T *>(out);
}

// Partials of the four CTAs of a cluster, held by CTA 0 (double buffered for the persistent kernel).
// Residency: cudaOccupancyMaxActiveClusters reports 71 clusters = 2 CTAs per SM whatever the ring size (shared memory
// and the 40-register cap would allow three: 3 x 76 KB, 3 x 21760 registers -- but 51 warps of 1280 registers do not
// split over the four 16 K-register files of an SM).  A three-slot ring at that residency was measured: no change (the
// inner loop is bound by instruction issue), so the ring stays at two slots and the scratch stays packed.
struct ClusterStats {
    float m[2][kCluster][kNK];
    float s[2][kCluster][kNK];
    int f[2][kCluster][kNK];              // first qualifying position, kNone if none; NEGATIVE (~f) when the CTA saw a near tie
};

struct BlockScratch {
    float m[kThreads];                    // indexed by thread (q * 17 + c): conflict-free both ways (17 is odd)
    float s[kThreads];
    unsigned char first[kThreads];        // index (0..15) of the thread's first value that reaches its maximum (exp == 1)
    float gmax[kNK];
    int found;
    unsigned char rescan[kNK + 3];        // CTA 0: channel needs the exact rescan
};

__device__ __forceinline__ float exp2f_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// a partial's sum moved onto the common maximum M (s == 0 covers the empty partial, whose m is -inf)
__device__ __forceinline__ float rescaled(float s, float m, float M) { return s == 0.0f ? 0.0f : fmul(s, __expf(fsub(m, M))); }

// Folds this thread's values (v[i] = logit at position p0 + q + 32 i, -inf where there is none) and merges the CTA's 32
// partials per channel: warp w reduces channel w and writes its (M, S, F, near) to slot `rank` of buffer `buf` in CTA 0.
// The kernel is bound by instruction issue, so the per-value work is kept to: max; subtract, compare, mask bit; one fma
// into the base-2 exponent, ex2, add.
__device__ __forceinline__ void slab_partials(const float (&v)[kMaxPerThread], int p0, int q, int c, int warp, int lane,
                                              float x0, BlockScratch &sc, ClusterStats *stats0, int buf, unsigned rank)
{
    float m = v[0];
#pragma unroll
    for (int i = 1; i < kMaxPerThread; ++i) m = fmaxf(m, v[i]);
    const float kLog2e = 1.4426950408889634f;
    const float m2 = m == -__int_as_float(0x7f800000) ? 0.0f : -m * kLog2e;    // empty partial: keep the exponents at -inf
    // Walked from the last value to the first: `first` ends up as the lowest index whose logit reaches the thread's
    // maximum under the recipe (one compare + one select per value), the sum is kept as a packed pair of running sums
    // (even / odd values), added once at the end.
    int first = kMaxPerThread;
    f32x2 ss = f2_pack(0.0f, 0.0f);
    const f32x2 mm = f2_bcast(m), l2e = f2_bcast(kLog2e), mm2 = f2_bcast(m2);
#pragma unroll
    for (int i = kMaxPerThread - 2; i >= 0; i -= 2) {  // two logits per FADD2 / FFMA2, same operations per lane
        const f32x2 vv = f2_pack(v[i], v[i + 1]);
        float d0, d1, t0, t1;
        f2_unpack(f2_sub(vv, mm), d0, d1);
        f2_unpack(f2_fma(vv, l2e, mm2), t0, t1);
        first = (d1 >= x0) ? i + 1 : first;
        first = (d0 >= x0) ? i : first;
        ss = f2_add(ss, f2_pack(exp2f_approx(t0), exp2f_approx(t1)));
    }
    float s0, s1;
    f2_unpack(ss, s0, s1);
    const float s = fadd(s0, s1);
    const int me = q * kNK + c;                        // == threadIdx.x
    sc.m[me] = m; sc.s[me] = s; sc.first[me] = (unsigned char)first;
    __syncthreads();
    {   // warp w = channel w: the 32 partials (q = lane) -> one
        const int src = lane * kNK + warp;
        const float pm = sc.m[src];
        float M = pm;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
        float S = rescaled(sc.s[src], pm, M);
        const int h = sc.first[src];
        int F = (pm == M && h < kMaxPerThread) ? p0 + lane + kLanes * h : kNone;
        const bool near = pm < M && fsub(pm, M) >= x0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {             // butterfly: a fixed summation tree
            S = fadd(S, __shfl_xor_sync(0xffffffffu, S, o));
            F = min(F, __shfl_xor_sync(0xffffffffu, F, o));
        }
        const unsigned any_near = __ballot_sync(0xffffffffu, near);
        if (lane == 0) {
            stats0->m[buf][rank][warp] = M;            // CTA 0's shared memory (local for rank 0, DSMEM otherwise)
            stats0->s[buf][rank][warp] = S;
            stats0->f[buf][rank][warp] = any_near ? ~F : F;
        }
    }
}

// CTA 0, after the cluster barrier: merge the four CTAs' partials in fixed order, rescan flagged channels exactly, store.
__device__ __forceinline__ void cluster_finish(const ClusterStats &st, int buf, BlockScratch &sc, const float *person, int P,
                                               int n, int crop_h, int crop_w, float x0, float *scores, float *positions,
                                               int *argmax_out)
{
    const int tid = threadIdx.x;
    float S = 0.0f;
    int F = kNone;
    int near = 0;
    if (tid < kNK) {
        float M = st.m[buf][0][tid];
        for (int r = 1; r < kCluster; ++r) M = fmaxf(M, st.m[buf][r][tid]);
        for (int r = 0; r < kCluster; ++r) {           // fixed order: CTA 0, 1, 2, 3
            const float m = st.m[buf][r][tid];
            const int fr = st.f[buf][r][tid];
            S = fadd(S, rescaled(st.s[buf][r][tid], m, M));
            if (m == M) F = min(F, fr < 0 ? ~fr : fr);
            // a near tie inside a CTA matters only if that CTA's maximum is itself within reach of M
            near |= (fr < 0 || m < M) && fsub(m, M) >= x0;
        }
        sc.rescan[tid] = (unsigned char)near;
        sc.gmax[tid] = M;
    }
    // exact rescan (block-uniform decision; practically never taken): first position with l - M >= x0
    if (__syncthreads_or(near)) {
        for (int ch = 0; ch < kNK; ++ch) {
            if (!sc.rescan[ch]) continue;              // uniform
            __syncthreads();
            if (tid == 0) sc.found = kNone;
            __syncthreads();
            const float M = sc.gmax[ch];
            int best = kNone;
            for (int p = tid; p < P; p += kThreads)
                if (fsub(__ldcg(person + (size_t)p * kNK + ch), M) >= x0) { best = p; break; }
            if (best != kNone) atomicMin(&sc.found, best);
            __syncthreads();
            if (tid == ch) F = sc.found;
        }
    }
    if (tid < kNK) {
        if (F == kNone) F = 0;                         // only with NaN logits
        const float rcp = fdiv(1.0f, S);
        const size_t o = (size_t)n * kNK + tid;
        scores[o] = fmul(1.0f, rcp);
        positions[o * 2 + 0] = fdiv((float)(F / crop_w), (float)crop_h);
        positions[o * 2 + 1] = fdiv((float)(F % crop_w), (float)crop_w);
        if (argmax_out) argmax_out[o] = F;
    }
}

// ---- streaming variant ------------------------------------------------------------------------------------------------
// PER: positions per CTA when known at compile time (504 for the reference's 56 x 36 crop: only the last of a thread's 16
// loads then needs a bounds test, on the lane alone; the per-value tests of the general form were a quarter of the
// kernel's instructions), 0 = derived from the runtime crop size.
template <int PER>
__global__ void __cluster_dims__(kCluster, 1, 1) __maxnreg__(40)
keypoint_decode_stream_kernel(const float *__restrict__ logits, const int *__restrict__ n_dev, const int n_host,
                              const int crop_h, const int crop_w, float *__restrict__ scores,
                              float *__restrict__ positions, int *__restrict__ argmax_out)
{
    extern __shared__ __align__(128) float s_slab[];      // kSlots x slab_floats
    __shared__ unsigned long long s_bar[kSlots];
    __shared__ BlockScratch sc;
    __shared__ ClusterStats stats;        // used in CTA 0 only
    const unsigned rank = cluster_rank();
    const int n_clusters = gridDim.x / kCluster;
    const int tid = threadIdx.x, c = tid % kNK, q = tid / kNK;
    const int P = PER > 0 ? PER * kCluster : crop_h * crop_w;
    const int per = PER > 0 ? PER : (P + kCluster - 1) / kCluster;   // per * 17 * 4 bytes is a multiple of 16 (checked by the host)
    const int p0 = (int)rank * per, p1 = min(P, p0 + per);
    const int slab_floats = per * kNK;
    const unsigned bytes = (unsigned)max(p1 - p0, 0) * kNK * 4u;
    pdl_trigger();
    if (tid == 0) bulk_barrier_init(s_bar, kSlots);
    pdl_wait();                           // the PRN has completed
    __syncthreads();
    const int N = n_dev ? *n_dev : n_host;
    const float x0 = g_exp_one_x0;
    ClusterStats *stats0 = peer_shared(&stats, 0);
    int n = blockIdx.x / kCluster;
    if (tid == 0 && bytes)
        for (int j = 0; j < kSlots - 1; ++j) {
            const long long nj = (long long)n + (long long)j * n_clusters;
            if (nj < N) bulk_fetch(s_slab + j * slab_floats, logits + (size_t)nj * P * kNK + (size_t)p0 * kNK, bytes, &s_bar[j]);
        }
    for (int it = 0; n < N; ++it, n += n_clusters) {       // uniform over the cluster
        const int slot = it % kSlots;
        const float *src = s_slab + slot * slab_floats;
        // slot (it - 1) % kSlots was last read in iteration it - 1, which every thread left through a cluster barrier
        const long long n_ahead = (long long)n + (long long)(kSlots - 1) * n_clusters;
        if (tid == 0 && n_ahead < N && bytes) {
            const int sa = (it + kSlots - 1) % kSlots;
            bulk_fetch(s_slab + sa * slab_floats, logits + (size_t)n_ahead * P * kNK + (size_t)p0 * kNK, bytes, &s_bar[sa]);
        }
        if (bytes) bulk_wait(&s_bar[slot], (unsigned)(it / kSlots) & 1u);
        float v[kMaxPerThread];
#pragma unroll
        for (int i = 0; i < kMaxPerThread; ++i) {
            const bool in = PER > 0 ? (kLanes * i + kLanes <= PER || q + kLanes * i < PER) : (p0 + q + kLanes * i < p1);
            v[i] = in ? src[tid + kThreads * i] : -__int_as_float(0x7f800000);
        }
        // partials of iteration it go to buffer it & 1 of CTA 0: CTA 0 read it last in iteration it - 2, i.e. before it
        // arrived at the barrier of iteration it - 1, which this CTA has already passed
        slab_partials(v, p0, q, c, tid >> 5, tid & 31, x0, sc, stats0, it & 1, rank);
        cluster_sync_all();               // all partials of this person have landed in CTA 0
        if (rank == 0)
            cluster_finish(stats, it & 1, sc, logits + (size_t)n * P * kNK, P, n, crop_h, crop_w, x0, scores, positions,
                           argmax_out);
    }
    // (no barrier on the way out: after the last in-loop barrier nobody touches another CTA's shared memory)
}

}  // namespace

constexpr int kPerDefault = 56 * 36 / kCluster;        // 504: the reference's crop (create_pb.py:19)

int kpdecode_prepare(cudaStream_t s, int crop_h, int crop_w, int *resident_clusters)
{
    exp_one_threshold_kernel<<<1, 1, 0, s>>>();
    if (cudaGetLastError() != cudaSuccess) return -1;
    // the streaming kernel's slab ring: sized for the default 56 x 36 crop and anything smaller
    if (cudaFuncSetAttribute(keypoint_decode_stream_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, kStreamSlabBytes) !=
            cudaSuccess ||
        cudaFuncSetAttribute(keypoint_decode_stream_kernel<kPerDefault>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             kStreamSlabBytes) != cudaSuccess)
        return -1;
    const int P = crop_h * crop_w, per = (P + kCluster - 1) / kCluster;
    if (P > kMaxPerThread * kLanes * kCluster || per % 4 != 0 || P % 4 != 0 || kSlots * per * kNK * 4 > kStreamSlabBytes) return -2;
    // clusters that fit the device at once (per handle: the launch grid is min(persons, this))
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kCluster * 1024); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSlots * per * kNK * 4;
    int n = 0;
    const cudaError_t e = P == kPerDefault * kCluster
                              ? cudaOccupancyMaxActiveClusters(&n, keypoint_decode_stream_kernel<kPerDefault>, &cfg)
                              : cudaOccupancyMaxActiveClusters(&n, keypoint_decode_stream_kernel<0>, &cfg);
    if (e != cudaSuccess || n < 1) n = 64;
    cudaGetLastError();
    *resident_clusters = n;
    return 0;
}

int launch_keypoint_decode(const float *logits, const int *n_dev, int n_host, int n_max, int crop_h, int crop_w,
                           int resident_clusters, float *scores, float *positions, int *argmax, cudaStream_t s)
{
    if (n_max <= 0) return 0;
    const int P = crop_h * crop_w, per = (P + kCluster - 1) / kCluster;
    const int clusters = n_max < resident_clusters ? n_max : resident_clusters;
    prof_mark(s, "keypoint_decode");
    if (P == kPerDefault * kCluster)
        launch_k(keypoint_decode_stream_kernel<kPerDefault>, dim3(clusters * kCluster), dim3(kThreads),
                 (size_t)(kSlots * per * kNK * 4), s, true, logits, n_dev, n_host, crop_h, crop_w, scores, positions, argmax);
    else
        launch_k(keypoint_decode_stream_kernel<0>, dim3(clusters * kCluster), dim3(kThreads), (size_t)(kSlots * per * kNK * 4), s,
                 true, logits, n_dev, n_host, crop_h, crop_w, scores, positions, argmax);
    return 1;
}

}  // namespace mpn
