// Pose Residual Network, bf16 mode, LARGE-BATCH regime (> 256 persons per call: crowded scenes, the PRN-only sweep of
// BASELINE configs[4]): two persistent tcgen05 GEMM kernels, one CTA per SM, warp-specialised like prn_fused.cu
// (warp 0 = TMA producer, warp 1 = MMA issuer, 16 epilogue warps).
//
// Replaces detector/prn.py:15-25:  y1 = relu(x W1 + b1),  y2 = relu(y1 W2 + b2),  logits = x + y2.
//
// Above ~210 persons the layers are bound by the tensor pipe and by the L2 -> SM operand traffic
//   bytes = M N K * 2 B * (1 / BM + 1 / BN), so the tiles are as large as tensor memory allows:
//   fc1  (K = 34272, N = 1024): 256 x 256 tiles (two 128-row accumulators share every weight box), split along K so
//        that ~148 work items exist whatever the person count; fp32 partial sums + a fixed-order reduce (bias, ReLU,
//        bf16).  The K loop of one item is hundreds of k blocks long, the epilogue is negligible.
//   fc2  (K = 1024, N = 34272): 128 x 240 tiles, TWO accumulator stages in tensor memory: the epilogue of tile i
//        (x + relu(acc + b2), 245 KB of global traffic -- this layer is HBM bound on its own output) overlaps the
//        16-k-block main loop of tile i + 1.
// The person count M exists only on the device: every CTA derives its work list from it; the kernels exit at once for
// M <= skip_le (those calls are served by prn_fused.cu).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "handle.cuh"
#include "tcgen05_utils.cuh"

namespace mpn {

namespace {

using namespace tc;

constexpr int kEpiWarps = 16;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kATileBytes = 128 * 128;                   // one 128-row activation tile (16 KB)
constexpr int kBTileBytes = 256 * 128;                   // one weight box (32 KB; fc2 uses 240 of the 256 rows)
constexpr int kRingBytes = 192 * 1024;
constexpr int kStgOffset = kRingBytes;                   // 16 x 2 KB per-warp transpose buffers
constexpr int kB2Offset = kStgOffset + kEpiWarps * 2048;
constexpr int kBarOffset = kB2Offset + 1024;
constexpr int kMaxStages = 6;
constexpr int kSmemBytes = kBarOffset + (2 * kMaxStages + 4) * 8 + 16 + 1024;
constexpr uint32_t kTmemCols = 512;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

enum { EPI_PARTIAL = 0, EPI_RESIDUAL = 1, EPI_REDUCE = 2 };   // fc1 partial sums; fc2 out of place; fc2 in place

struct BigArgs {
    const int *m_dev;
    int m_host;
    int skip_le;
    int n_total;             // output columns (1024 or 34272)
    int nkb;                 // k blocks of the layer
    int n_tiles;             // output column tiles
    int sms;                 // work items the split-K heuristic aims at
    float *out;              // fc1: partial sums; fc2: logits [M, n_total]
    size_t out_floats;       // fc1: capacity of the partial buffer
    const float *bias;       // fc2
    const float *residual;   // fc2: x [M, n_total]
};

// K splits of fc1 for M persons: enough work items for every SM, bounded by the partial-sum buffer.  Pure function of
// (M, capacity) so that the GEMM and the reduce kernel agree without talking to the host.
__host__ __device__ inline int big_splits(int M, int nkb, int sms, size_t cap_floats, int hidden)
{
    const int tiles = ((M + 255) / 256) * (hidden / 256);
    int s = (sms + tiles / 2) / tiles;
    if (s < 1) s = 1;
    if (s > 32) s = 32;
    if (s > nkb) s = nkb;
    const size_t per_split = (size_t)((M + 255) / 256) * 256 * hidden;
    while (s > 1 && (size_t)s * per_split > cap_floats) --s;
    return s;
}

// fc2 work order: BANDS of kBand weight tiles (16 x 480 KB = 7.9 MB of W2t), each band swept over ALL row blocks before
// the next one starts.  The band and y1 stay in L2 while the 768 MB of residual / logit traffic of the layer streams
// through it, so W2t is read from HBM once.  (Row-block-major order touched all 70 MB of W2t between two uses of any
// tile, with 36 MB of streaming traffic in between: ncu showed 1.17 GB of DRAM reads against 0.46 GB compulsory.)
constexpr int kBand = 16;
__device__ __forceinline__ void fc2_item(int item, int m_tiles, int n_tiles, int &mt, int &nt)
{
    const int per_band = m_tiles * kBand;
    const int band = item / per_band;
    const int n0 = band * kBand;
    const int bw = min(kBand, n_tiles - n0);
    const int rem = item - band * per_band;
    mt = rem / bw;
    nt = n0 + rem - mt * bw;
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

// fc2 epilogue of one warp, OUT OF PLACE: logits = x + relu(acc + b2) (detector/prn.py:22,24) for its 32 accumulator rows
// (TMEM lanes of t_acc) and its column group cg (64 columns as chunks of 16; the last group of a 240-column tile has 3).
// Chunks go through the per-warp transpose so that the residual loads and the logit stores are 64-byte row segments.
// Every chunk waits for its own residual loads (HBM: the crops were written hundreds of MB ago): ~15 us per tile against
// 4 us for the tile's MMAs, which is what holds this variant at 0.6 PFLOP/s.  Requesting the residual earlier needs
// 64 registers per thread that 576-thread CTAs do not have (measured: the spills made it 25 % slower); the in-place
// variant below avoids the loads altogether.
template <int BN>
__device__ __forceinline__ void residual_epilogue(const BigArgs &args, int M, int row0, int n0, uint32_t t_acc, int cg, int lane,
                                                  int sub_row, int sub_col, float *stg, const float *s_b2)
{
    const int col_base = cg * 64 + sub_col;
    const size_t off = (size_t)(row0 + sub_row) * args.n_total + n0 + col_base;
    const size_t row_step = (size_t)8 * args.n_total;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        if (cg * 64 + ch * 16 >= BN) break;                 // warp-uniform (last column group: 3 chunks)
        const bool col_ok = n0 + col_base + ch * 16 < args.n_total;
        float4 xr[4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
            xr[i] = (col_ok && row0 + sub_row + 8 * i < M)
                        ? __ldcs(reinterpret_cast<const float4 *>(args.residual + off + i * row_step + ch * 16))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        uint32_t r[16];
        tmem_ld16(t_acc + (uint32_t)(cg * 64 + ch * 16), r);
        stage_write(stg, lane, r);
        __syncwarp();
        if (col_ok) {
            const float4 b = *reinterpret_cast<const float4 *>(s_b2 + col_base + ch * 16);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (row0 + sub_row + 8 * i < M) {
                    const float4 a = stage_read(stg, lane, i);
                    float4 o;
                    o.x = __fadd_rn(xr[i].x, fmaxf(__fadd_rn(a.x, b.x), 0.0f));
                    o.y = __fadd_rn(xr[i].y, fmaxf(__fadd_rn(a.y, b.y), 0.0f));
                    o.z = __fadd_rn(xr[i].z, fmaxf(__fadd_rn(a.z, b.z), 0.0f));
                    o.w = __fadd_rn(xr[i].w, fmaxf(__fadd_rn(a.w, b.w), 0.0f));
                    __stcs(reinterpret_cast<float4 *>(args.out + off + i * row_step + ch * 16), o);
                }
            }
        }
        __syncwarp();
    }
}

// fc2 epilogue IN PLACE (the logits replace the residual: x += relu(acc + b2)): the warp never loads x.  Bias and ReLU are
// applied in registers, the 32 x 16 chunk goes to the warp's transpose buffer, whose layout is exactly a 64-byte-swizzled
// TMA box, and one thread hands it to a TMA reduce-add: the fp32 addition x + y2 is performed in L2 (the same single
// rounding as the load / add / store version, bit for bit).  The epilogue is then bounded by the TMEM drain instead of
// four dependent HBM round trips per tile, and 4 B / element of L2 -> SM traffic disappear.
// Rows past the person count inside the last row block receive meaningless sums; nothing ever reads them.
template <int BN>
__device__ __forceinline__ void reduce_epilogue(const CUtensorMap *tmap_out, int row0, int n0, uint32_t t_acc, int cg, int lane,
                                                float *stg, const float *s_b2)
{
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        if (cg * 64 + ch * 16 >= BN) break;                 // warp-uniform (last column group: 3 chunks)
        uint32_t r[16];
        tmem_ld16(t_acc + (uint32_t)(cg * 64 + ch * 16), r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 b = *reinterpret_cast<const float4 *>(s_b2 + cg * 64 + ch * 16 + 4 * j);   // broadcast
            r[4 * j + 0] = __float_as_uint(fmaxf(__fadd_rn(__uint_as_float(r[4 * j + 0]), b.x), 0.0f));
            r[4 * j + 1] = __float_as_uint(fmaxf(__fadd_rn(__uint_as_float(r[4 * j + 1]), b.y), 0.0f));
            r[4 * j + 2] = __float_as_uint(fmaxf(__fadd_rn(__uint_as_float(r[4 * j + 2]), b.z), 0.0f));
            r[4 * j + 3] = __float_as_uint(fmaxf(__fadd_rn(__uint_as_float(r[4 * j + 3]), b.w), 0.0f));
        }
        if (lane == 0) tma_store_wait_read();               // the previous chunk has left the buffer
        __syncwarp();
        stage_write(stg, lane, r);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) tma_reduce_add_2d(tmap_out, stg, n0 + cg * 64 + ch * 16, row0);
    }
}

template <int BN, int MSUB, int ACC, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
big_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_out, const BigArgs args)
{
    static_assert(MSUB * ACC * 256 <= 512, "tensor memory columns");
    constexpr int kStageBytes = MSUB * kATileBytes + kBTileBytes;
    constexpr int kStages = kRingBytes / kStageBytes;
    constexpr uint32_t kTxBytes = MSUB * kATileBytes + BN * 128;
    extern __shared__ uint8_t smem_raw[];
    const int M = args.m_dev ? *args.m_dev : args.m_host;
    if (M <= 0 || M <= args.skip_le) return;            // uniform over the grid
    const int G = gridDim.x, c = blockIdx.x;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kBarOffset);
    uint64_t *empty_bar = full_bar + kMaxStages;
    uint64_t *tmem_full_bar = empty_bar + kMaxStages;    // [2]
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;        // [2]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
        for (int i = 0; i < kStages; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tmem_full_bar + i, 1); mbar_init(tmem_empty_bar + i, kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- work list (identical in every role): item -> (K split z, M tile, N tile)
    const int m_tiles = (M + 128 * MSUB - 1) / (128 * MSUB);
    const int splits = EPI == EPI_PARTIAL ? big_splits(M, args.nkb, args.sms, args.out_floats, args.n_total) : 1;
    const int per_split = m_tiles * args.n_tiles;
    const int items = per_split * splits;
    const size_t split_stride = (size_t)((M + 255) / 256) * 256 * args.n_total;     // fc1 partial layout

    if (warp == 0) {
        if (lane == 0) {   // ================= TMA producer =================
            int it = 0;
            for (int item = c; item < items; item += G) {
                const int z = item / per_split, rem = item - z * per_split;
                int mt = rem / args.n_tiles, nt = rem - mt * args.n_tiles;
                if (EPI != EPI_PARTIAL) fc2_item(item, m_tiles, args.n_tiles, mt, nt);
                const int kb0 = (int)(((long long)z * args.nkb) / splits), kb1 = (int)(((long long)(z + 1) * args.nkb) / splits);
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int st = it % kStages;
                    mbar_wait(empty_bar + st, (((uint32_t)(it / kStages)) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(full_bar + st, kTxBytes);
                    uint8_t *stage = smem + st * kStageBytes;
                    tma_load_2d(stage, &tmap_b, full_bar + st, kb * BLOCK_K, nt * BN, kEvictLast);
#pragma unroll
                    for (int sub = 0; sub < MSUB; ++sub)
                        tma_load_2d(stage + kBTileBytes + sub * kATileBytes, &tmap_a, full_bar + st, kb * BLOCK_K,
                                    (mt * MSUB + sub) * 128, kEvictLast);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {   // ================= MMA issuer =================
            constexpr uint32_t idesc = make_idesc_bf16(128, BN);
            int it = 0, n_item = 0;
            for (int item = c; item < items; item += G, ++n_item) {
                const int z = item / per_split;
                const int kb0 = (int)(((long long)z * args.nkb) / splits), kb1 = (int)(((long long)(z + 1) * args.nkb) / splits);
                const int as = n_item % ACC;
                if (n_item >= ACC) {                     // the epilogue has drained this accumulator stage
                    mbar_wait(tmem_empty_bar + as, ((uint32_t)(n_item / ACC - 1)) & 1u);
                    tc_fence_after();
                }
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int st = it % kStages;
                    mbar_wait(full_bar + st, ((uint32_t)(it / kStages)) & 1u);
                    tc_fence_after();
                    const uint32_t s_addr = smem_u32(smem + st * kStageBytes);
                    const uint64_t bdesc = make_kmajor_sw128_desc(s_addr);
#pragma unroll
                    for (int sub = 0; sub < MSUB; ++sub) {
                        const uint64_t adesc = make_kmajor_sw128_desc(s_addr + kBTileBytes + sub * kATileBytes);
#pragma unroll
                        for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                            umma_bf16(tmem_base + (uint32_t)((as * MSUB + sub) * 256), adesc + (uint64_t)(2 * k),
                                      bdesc + (uint64_t)(2 * k), idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar + st);
                }
                umma_commit(tmem_full_bar + as);
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue warps: TMEM lane quadrant q = warp % 4, column group cg = (warp - 2) / 4 =========
        const int ew = warp - 2, q = warp & 3, cg = ew >> 2;
        const int tid_e = threadIdx.x - 64;
        float *stg = reinterpret_cast<float *>(smem + kStgOffset + ew * 2048);
        float *s_b2 = reinterpret_cast<float *>(smem + kB2Offset);
        const int sub_row = lane >> 2, sub_col = (lane & 3) << 2;
        int n_item = 0;
        for (int item = c; item < items; item += G, ++n_item) {
            const int z = item / per_split, rem = item - z * per_split;
            int mt = rem / args.n_tiles, nt = rem - mt * args.n_tiles;
            if (EPI != EPI_PARTIAL) fc2_item(item, m_tiles, args.n_tiles, mt, nt);
            const int as = n_item % ACC;
            const int n0 = nt * BN;
            if (EPI != EPI_PARTIAL) {
                if (n_item > 0) epi_bar_sync();                             // previous tile is done with s_b2
                if (tid_e < BN / 4) {
                    const int n = n0 + tid_e * 4;
                    reinterpret_cast<float4 *>(s_b2)[tid_e] =
                        n < args.n_total ? __ldg(reinterpret_cast<const float4 *>(args.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                epi_bar_sync();
            }
            mbar_wait(tmem_full_bar + as, ((uint32_t)(n_item / ACC)) & 1u);
            tc_fence_after();
#pragma unroll 1
            for (int sub = 0; sub < MSUB; ++sub) {
                const int row0 = (mt * MSUB + sub) * 128 + q * 32;
                if (row0 >= M) break;                                       // warp-uniform
                const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((as * MSUB + sub) * 256);
                const int col_base = cg * 64 + sub_col;
                if (EPI == EPI_PARTIAL) {
                    float *dst = args.out + (size_t)z * split_stride + (size_t)(row0 + sub_row) * args.n_total + n0 + col_base;
                    const size_t row_step = (size_t)8 * args.n_total;
#pragma unroll
                    for (int ch = 0; ch < 4; ++ch) {
                        uint32_t r[16];
                        tmem_ld16(t_acc + (uint32_t)(cg * 64 + ch * 16), r);
                        stage_write(stg, lane, r);
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                            if (row0 + sub_row + 8 * i < M)
                                __stcg(reinterpret_cast<float4 *>(dst + i * row_step + ch * 16), stage_read(stg, lane, i));
                        __syncwarp();
                    }
                } else if (EPI == EPI_RESIDUAL) {
                    residual_epilogue<BN>(args, M, row0, n0, t_acc, cg, lane, sub_row, sub_col, stg, s_b2);
                } else {
                    reduce_epilogue<BN>(&tmap_out, row0, n0, t_acc, cg, lane, stg, s_b2);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(tmem_empty_bar + as)) : "memory");
        }
    }
    if (EPI == EPI_REDUCE && warp >= 2 && lane == 0) tma_store_wait_all();   // this thread's reduce-adds have completed
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
}

// ---- fc2 on CTA PAIRS (tcgen05 cta_group::2) --------------------------------------------------------------------------
// fc2 is bound by L2 -> SM operand traffic (a 128 x 240 tile pulls (128 + 240) x 1024 x 2 B = 736 KB for 31 M MACs).  Two
// CTAs of a cluster work on ONE 256 x 240 tile: each loads its own 128 rows of y1 and only HALF of the weight box (120
// rows of W2t); the leader CTA issues tcgen05.mma.cta_group::2 (M = 256), which reads both halves from the two CTAs'
// shared memory and writes rows 0..127 of the accumulator to the leader's tensor memory and rows 128..255 to the peer's.
// Per CTA and tile that is (128 + 120) x 1024 x 2 B = 496 KB: a third less operand traffic, and half the MMA
// instructions.  Protocol (as in CUTLASS's 2-SM kernels):
//   * both producers wait on their OWN empty barrier and issue cp.async.bulk.tensor ... cta_group::2 with the LEADER's
//     full barrier as the completion barrier; the leader's producer arms it with the bytes of both CTAs;
//   * the leader's MMA thread waits on its full barrier, issues the MMAs, and tcgen05.commit ... multicast arrives on the
//     empty barrier (stage free) resp. the accumulator-full barrier of BOTH CTAs;
//   * every CTA's epilogue warps drain their own 128 accumulator rows and arrive on the LEADER's accumulator-empty
//     barrier (count 2 x 16 warps), which the MMA thread waits on before reusing that accumulator stage.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;          // clears the CTA-rank bit of a shared::cluster address: CTA 0 of the pair
constexpr int kPairBHalf = 120;                          // weight rows per CTA
constexpr int kPairBBytes = 16 * 1024;                   // 120 x 128 B = 15 KB, padded so that every tile stays 1024-aligned
constexpr int kPairStageBytes = kATileBytes + kPairBBytes;
constexpr int kPairStages = kRingBytes / kPairStageBytes;        // 6
constexpr uint32_t kPairTxBytes = kATileBytes + kPairBHalf * 128;   // per CTA
static_assert(kPairStages <= kMaxStages, "barrier slots");

__device__ __forceinline__ unsigned pair_rank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %cluster_ctarank;" : "=r"(r));
    return r;
}

__device__ __forceinline__ void pair_sync()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void tma_load_2d_pair(void *smem_dst, const CUtensorMap *tmap, uint64_t *leader_bar, int c0, int c1,
                                                 uint64_t hint)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0),
          "r"(c1), "l"(hint)
        : "memory");
}

__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

// arrives (once the MMAs issued so far have completed) on the barrier at this shared-memory offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((unsigned short)3)
                 : "memory");
}

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
big_fc2_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_out, const BigArgs args)
{
    constexpr int BN = 240, ACC = 2;
    extern __shared__ uint8_t smem_raw[];
    const int M = args.m_dev ? *args.m_dev : args.m_host;
    if (M <= 0 || M <= args.skip_le) return;            // uniform over the grid
    const int P = gridDim.x >> 1, pair = blockIdx.x >> 1;
    const unsigned rank = pair_rank();
    const bool leader = rank == 0;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + kBarOffset);
    uint64_t *empty_bar = full_bar + kMaxStages;
    uint64_t *tmem_full_bar = empty_bar + kMaxStages;    // [2]
    uint64_t *tmem_empty_bar = tmem_full_bar + 2;        // [2]  (the leader's are the ones in use)
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
        for (int i = 0; i < kPairStages; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(tmem_full_bar + i, 1); mbar_init(tmem_empty_bar + i, 2 * kEpiWarps); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    pair_sync();                                         // barriers and tensor memory of both CTAs are ready
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int m_pairs = (M + 255) / 256;
    const int items = m_pairs * args.n_tiles;

    if (warp == 0) {
        if (lane == 0) {   // ================= TMA producer (both CTAs) =================
            int it = 0;
            for (int item = pair; item < items; item += P) {
                int mp, nt;
                fc2_item(item, m_pairs, args.n_tiles, mp, nt);
                for (int kb = 0; kb < args.nkb; ++kb, ++it) {
                    const int st = it % kPairStages;
                    mbar_wait(empty_bar + st, (((uint32_t)(it / kPairStages)) & 1u) ^ 1u);
                    if (leader) mbar_arrive_expect_tx(full_bar + st, 2 * kPairTxBytes);
                    uint8_t *stage = smem + st * kPairStageBytes;
                    tma_load_2d_pair(stage, &tmap_b, full_bar + st, kb * BLOCK_K, nt * BN + (int)rank * kPairBHalf, kEvictLast);
                    tma_load_2d_pair(stage + kPairBBytes, &tmap_a, full_bar + st, kb * BLOCK_K, (mp * 2 + (int)rank) * 128,
                                     kEvictLast);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && leader) {   // ================= MMA issuer (leader CTA only) =================
            constexpr uint32_t idesc = make_idesc_bf16(256, BN);
            int it = 0, n_item = 0;
            for (int item = pair; item < items; item += P, ++n_item) {
                const int as = n_item % ACC;
                if (n_item >= ACC) {                     // both CTAs' epilogues have drained this accumulator stage
                    mbar_wait(tmem_empty_bar + as, ((uint32_t)(n_item / ACC - 1)) & 1u);
                    tc_fence_after();
                }
                for (int kb = 0; kb < args.nkb; ++kb, ++it) {
                    const int st = it % kPairStages;
                    mbar_wait(full_bar + st, ((uint32_t)(it / kPairStages)) & 1u);
                    tc_fence_after();
                    const uint32_t s_addr = smem_u32(smem + st * kPairStageBytes);
                    const uint64_t bdesc = make_kmajor_sw128_desc(s_addr);
                    const uint64_t adesc = make_kmajor_sw128_desc(s_addr + kPairBBytes);
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k)
                        umma_bf16_pair(tmem_base + (uint32_t)(as * 256), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                                       (kb > 0 || k > 0) ? 1u : 0u);
                    umma_commit_pair(empty_bar + st);
                }
                umma_commit_pair(tmem_full_bar + as);
            }
        }
        __syncwarp();
    } else {
        // ================= epilogue warps (both CTAs): this CTA's 128 rows of the pair's tile =================
        const int ew = warp - 2, q = warp & 3, cg = ew >> 2;
        const int tid_e = threadIdx.x - 64;
        float *stg = reinterpret_cast<float *>(smem + kStgOffset + ew * 2048);
        float *s_b2 = reinterpret_cast<float *>(smem + kB2Offset);
        const int sub_row = lane >> 2, sub_col = (lane & 3) << 2;
        const uint32_t leader_empty = smem_u32(tmem_empty_bar) & kPeerBitMask;
        int n_item = 0;
        for (int item = pair; item < items; item += P, ++n_item) {
            int mp, nt;
            fc2_item(item, m_pairs, args.n_tiles, mp, nt);
            const int as = n_item % ACC;
            const int n0 = nt * BN;
            if (n_item > 0) epi_bar_sync();                                 // previous tile is done with s_b2
            if (tid_e < BN / 4) {
                const int n = n0 + tid_e * 4;
                reinterpret_cast<float4 *>(s_b2)[tid_e] =
                    n < args.n_total ? __ldg(reinterpret_cast<const float4 *>(args.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            epi_bar_sync();
            const int row0 = (mp * 2 + (int)rank) * 128 + q * 32;
            mbar_wait(tmem_full_bar + as, ((uint32_t)(n_item / ACC)) & 1u);
            tc_fence_after();
            if (row0 < M) {                                                 // warp-uniform
                const uint32_t t_acc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * 256);
                if (EPI == EPI_RESIDUAL) residual_epilogue<BN>(args, M, row0, n0, t_acc, cg, lane, sub_row, sub_col, stg, s_b2);
                else reduce_epilogue<BN>(&tmap_out, row0, n0, t_acc, cg, lane, stg, s_b2);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(leader_empty + (uint32_t)(as * 8)) : "memory");
        }
    }
    if (EPI == EPI_REDUCE && warp >= 2 && lane == 0) tma_store_wait_all();   // this thread's reduce-adds have completed
    tc_fence_before();
    pair_sync();                                         // nobody's shared / tensor memory is in use by the peer any more
    if (warp == 1)
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
}

// y1[m, j] = relu(sum_z partial[z, m, j] + b1[j]) in bf16, z ascending (detector/prn.py:20); splits derived from M
__global__ void __launch_bounds__(256) big_fc1_reduce_kernel(const float *__restrict__ partial, const size_t cap_floats,
                                                             const float *__restrict__ bias, const int hidden,
                                                             const int nkb, const int sms, const int *__restrict__ m_dev,
                                                             const int m_host, const int skip_le,
                                                             __nv_bfloat16 *__restrict__ y1)
{
    const int M = m_dev ? *m_dev : m_host;
    if (M <= 0 || M <= skip_le) return;
    const int splits = big_splits(M, nkb, sms, cap_floats, hidden);
    const size_t split_stride = (size_t)((M + 255) / 256) * 256 * hidden;
    const size_t total4 = (size_t)M * hidden / 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
        float4 acc = __ldcs(reinterpret_cast<const float4 *>(partial) + i);
        for (int z = 1; z < splits; ++z) {
            const float4 v = __ldcs(reinterpret_cast<const float4 *>(partial + (size_t)z * split_stride) + i);
            acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y);
            acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
        }
        const float4 b = __ldg(reinterpret_cast<const float4 *>(bias + (i * 4) % hidden));
        acc.x = fmaxf(__fadd_rn(acc.x, b.x), 0.0f); acc.y = fmaxf(__fadd_rn(acc.y, b.y), 0.0f);
        acc.z = fmaxf(__fadd_rn(acc.z, b.z), 0.0f); acc.w = fmaxf(__fadd_rn(acc.w, b.w), 0.0f);
        const __nv_bfloat162 lo = __floats2bfloat162_rn(acc.x, acc.y), hi = __floats2bfloat162_rn(acc.z, acc.w);
        uint2 o;
        o.x = *reinterpret_cast<const unsigned *>(&lo);
        o.y = *reinterpret_cast<const unsigned *>(&hi);
        reinterpret_cast<uint2 *>(y1)[i] = o;
    }
}

}  // namespace

struct BigState {
    CUtensorMap a1, b1, a2, b2, b2_half;
    CUtensorMap out_map;     // fp32 [rows, D] chunk boxes over the buffer of the most recent in-place call
    const float *out_ptr;
    uint64_t rows;
    bool fc2_pairs;          // fc2 on CTA pairs (cta_group::2); MPN_FC2_PAIRS=0 keeps the single-CTA kernel
    float *partial;
    size_t partial_floats;
    int sms;
};

int prn_big_prepare(mpn_handle *h)
{
    const int D = h->D, Hd = h->cfg.prn_hidden;
    if (Hd % 256 != 0 || D % 16 != 0) return MPN_OK;       // shape not covered: prn_tcgen05.cu is used
    if (h->prn_ws.n_max <= kPrnFusedMaxRows && h->fused) return MPN_OK;   // the handle can never exceed the fused regime
    BigState *st = new BigState;
    memset(st, 0, sizeof(*st));
    cudaDeviceGetAttribute(&st->sms, cudaDevAttrMultiProcessorCount, h->cfg.device);
    const uint64_t rows = (uint64_t)h->prn_ws.n_max;
    // one 256-row-padded copy of the partial sums per K split; splits * persons is bounded by ~sms * 64 + padding
    const size_t rows_pad = (size_t)(h->prn_ws.n_max + 255) / 256 * 256;
    st->partial_floats = (rows_pad + (size_t)st->sms * 128) * Hd;
    auto k1 = big_gemm_kernel<256, 2, 1, EPI_PARTIAL>;
    auto k2 = big_gemm_kernel<240, 1, 2, EPI_RESIDUAL>;
    auto k2r = big_gemm_kernel<240, 1, 2, EPI_REDUCE>;
    bool ok = cudaMalloc(&st->partial, st->partial_floats * sizeof(float)) == cudaSuccess &&
              cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) == cudaSuccess &&
              cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) == cudaSuccess &&
              cudaFuncSetAttribute(k2r, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) == cudaSuccess;
    ok = ok && encode_2d(&st->a1, h->crops_bf16, rows, (uint64_t)D, 128) && encode_2d(&st->b1, h->W1t, (uint64_t)Hd, (uint64_t)D, 256) &&
         encode_2d(&st->a2, h->prn_ws.y1_bf16, rows, (uint64_t)Hd, 128) && encode_2d(&st->b2, h->W2t, (uint64_t)D, (uint64_t)Hd, 240);
    const char *pairs_env = getenv("MPN_FC2_PAIRS");
    st->fc2_pairs = !(pairs_env && pairs_env[0] == '0') && st->sms % 2 == 0;
    if (ok && st->fc2_pairs)       // (clusters need not be co-resident: there is no grid-wide synchronisation in this kernel)
        ok = cudaFuncSetAttribute(big_fc2_pair_kernel<EPI_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) == cudaSuccess &&
             cudaFuncSetAttribute(big_fc2_pair_kernel<EPI_REDUCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) == cudaSuccess &&
             encode_2d(&st->b2_half, h->W2t, (uint64_t)D, (uint64_t)Hd, kPairBHalf);
    st->rows = rows;
    if (!ok) {
        cudaGetLastError();
        if (st->partial) cudaFree(st->partial);
        delete st;
        snprintf(h->err, sizeof(h->err), "large-batch PRN setup failed (allocation, shared memory or cuTensorMapEncodeTiled)");
        return MPN_ERR_CUDA;
    }
    h->big = st;
    return MPN_OK;
}

void prn_big_release(mpn_handle *h)
{
    BigState *st = static_cast<BigState *>(h->big);
    if (!st) return;
    cudaFree(st->partial);
    delete st;
    h->big = nullptr;
}

int launch_prn_big(mpn_handle *h, const float *x_f32, const int *n_dev, int n_host, int n_max, float *logits, int skip_le,
                   cudaStream_t s)
{
    BigState *st = static_cast<BigState *>(h->big);
    if (!st) return -(int)cudaErrorInvalidValue;
    if (n_max <= 0) return 0;
    const int D = h->D, Hd = h->cfg.prn_hidden;
    BigArgs a;
    a.m_dev = n_dev; a.m_host = n_host; a.skip_le = skip_le; a.sms = st->sms;
    // fc1: [M, D] x [D, hidden] -> partial sums
    a.n_total = Hd; a.nkb = (D + BLOCK_K - 1) / BLOCK_K; a.n_tiles = Hd / 256;
    a.out = st->partial; a.out_floats = st->partial_floats; a.bias = nullptr; a.residual = nullptr;
    prof_mark(s, "prn_big_fc1");
    big_gemm_kernel<256, 2, 1, EPI_PARTIAL><<<st->sms, kThreads, kSmemBytes, s>>>(st->a1, st->b1, st->out_map, a);
    prof_mark(s, "prn_big_reduce");
    big_fc1_reduce_kernel<<<st->sms * 4, 256, 0, s>>>(st->partial, st->partial_floats, h->b1, Hd, a.nkb, st->sms, n_dev, n_host,
                                                      skip_le, h->prn_ws.y1_bf16);
    // fc2: [M, hidden] x [hidden, D] + bias + ReLU + residual -> logits
    a.n_total = D; a.nkb = Hd / BLOCK_K; a.n_tiles = (D + 239) / 240;
    a.out = logits; a.out_floats = 0; a.bias = h->b2; a.residual = x_f32;
    // logits == x: in place, the residual addition is a TMA reduce-add performed in L2 (the tensor map describes the
    // caller's buffer; it is kernel-parameter data, so re-encoding it for another buffer does not disturb earlier launches)
    const bool in_place = x_f32 == logits;
    if (in_place && st->out_ptr != logits) {
        const uint64_t out_rows = n_dev ? st->rows : (uint64_t)n_host;     // mpn_prn: the caller's buffer holds n_host rows
        if (!encode_2d_f32_chunk(&st->out_map, logits, out_rows, (uint64_t)D, 32)) return -(int)cudaErrorInvalidValue;
        st->out_ptr = n_dev ? logits : nullptr;        // a caller's buffer may change size between calls: never cached
    }
    prof_mark(s, "prn_big_fc2");
    if (st->fc2_pairs && in_place)
        big_fc2_pair_kernel<EPI_REDUCE><<<st->sms, kThreads, kSmemBytes, s>>>(st->a2, st->b2_half, st->out_map, a);
    else if (st->fc2_pairs)
        big_fc2_pair_kernel<EPI_RESIDUAL><<<st->sms, kThreads, kSmemBytes, s>>>(st->a2, st->b2_half, st->out_map, a);
    else if (in_place)
        big_gemm_kernel<240, 1, 2, EPI_REDUCE><<<st->sms, kThreads, kSmemBytes, s>>>(st->a2, st->b2, st->out_map, a);
    else
        big_gemm_kernel<240, 1, 2, EPI_RESIDUAL><<<st->sms, kThreads, kSmemBytes, s>>>(st->a2, st->b2, st->out_map, a);
    return 3;
}

}  // namespace mpn
