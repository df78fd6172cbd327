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
constexpr int kMaxStages = 4;
constexpr int kSmemBytes = kBarOffset + (2 * kMaxStages + 4) * 8 + 16 + 1024;
constexpr uint32_t kTmemCols = 512;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

enum { EPI_PARTIAL = 0, EPI_RESIDUAL = 1 };

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

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory"); }

template <int BN, int MSUB, int ACC, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
big_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const BigArgs args)
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
                const int mt = rem / args.n_tiles, nt = rem - mt * args.n_tiles;
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
            const int mt = rem / args.n_tiles, nt = rem - mt * args.n_tiles;
            const int as = n_item % ACC;
            const int n0 = nt * BN;
            if (EPI == EPI_RESIDUAL) {
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
                } else {
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
                                    float4 o;   // x + relu(acc + b2)   (detector/prn.py:22,24)
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
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0)
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(tmem_empty_bar + as)) : "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
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
    CUtensorMap a1, b1, a2, b2;
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
    bool ok = cudaMalloc(&st->partial, st->partial_floats * sizeof(float)) == cudaSuccess &&
              cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) == cudaSuccess &&
              cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) == cudaSuccess;
    ok = ok && encode_2d(&st->a1, h->crops_bf16, rows, (uint64_t)D, 128) && encode_2d(&st->b1, h->W1t, (uint64_t)Hd, (uint64_t)D, 256) &&
         encode_2d(&st->a2, h->prn_ws.y1_bf16, rows, (uint64_t)Hd, 128) && encode_2d(&st->b2, h->W2t, (uint64_t)D, (uint64_t)Hd, 240);
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
    big_gemm_kernel<256, 2, 1, EPI_PARTIAL><<<st->sms, kThreads, kSmemBytes, s>>>(st->a1, st->b1, a);
    prof_mark(s, "prn_big_reduce");
    big_fc1_reduce_kernel<<<st->sms * 4, 256, 0, s>>>(st->partial, st->partial_floats, h->b1, Hd, a.nkb, st->sms, n_dev, n_host,
                                                      skip_le, h->prn_ws.y1_bf16);
    // fc2: [M, hidden] x [hidden, D] + bias + ReLU + residual -> logits
    a.n_total = D; a.nkb = Hd / BLOCK_K; a.n_tiles = (D + 239) / 240;
    a.out = logits; a.out_floats = 0; a.bias = h->b2; a.residual = x_f32;
    prof_mark(s, "prn_big_fc2");
    big_gemm_kernel<240, 1, 2, EPI_RESIDUAL><<<st->sms, kThreads, kSmemBytes, s>>>(st->a2, st->b2, a);
    return 3;
}

}  // namespace mpn
