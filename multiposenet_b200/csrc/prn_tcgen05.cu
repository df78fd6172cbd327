// Pose Residual Network, bf16 mode: the two dense layers on the 5th-generation tensor cores.
//
// Replaces detector/prn.py:15-25 (slim.fully_connected x2 + residual) for the throughput path:
//   y1 = relu(x W1 + b1),  y2 = relu(y1 W2 + b2),  logits = x + y2
// with bf16 operands (x, W1, y1, W2 rounded to nearest even) and fp32 accumulation in tensor memory.
//
// One kernel template, D[M, N] = A[M, K] * Bt[N, K]^T with both operands K-major:
//   * TMA (cp.async.bulk.tensor.2d, 128-byte swizzle) stages 128 x 64 A tiles and BLOCK_N x 64 B tiles through a
//     ring of shared-memory buffers guarded by full / empty mbarriers            (warp 0, one thread)
//   * tcgen05.mma.cta_group::1.kind::f16, M = 128, N = BLOCK_N, K = 16 per instruction, accumulator in TMEM; the
//     slot is released with tcgen05.commit                                         (warp 1, one thread)
//   * four epilogue warps read the accumulator with tcgen05.ld (32 lanes x 16 columns per instruction) and apply
//       fc1: store the split-K partial sum (fp32) for the deterministic reduce that follows
//       fc2: logits = x + relu(acc + b2), written straight to HBM                  (warps 2..5)
// The weights are kept transposed ([out, in], K-major) so that both layers use the same K-major/K-major
// instruction form.  The number of persons M is only known on the device: CTAs whose M tile is empty exit at once.
//
// fc1 (K = 34272, N = 1024) is split along K so that ~one wave of CTAs streams W1 exactly once; fc2 (K = 1024,
// N = 34272 = 357 * 96) gets one CTA per 96 output columns, two CTAs resident per SM so that the epilogue of one
// overlaps the weight stream of the other.  For M <= ~200 persons both layers are bound by the HBM stream of the
// 140 MB of bf16 weights, which is what the staging depth is sized for.
#include <cuda.h>

#include <cstdio>

#include "common.cuh"
#include "handle.cuh"
#include "tcgen05_utils.cuh"

namespace mpn {

namespace {

using namespace tc;

constexpr int BLOCK_M = 128;
constexpr int kGemmThreads = 192;    // warp 0 TMA, warp 1 MMA + TMEM allocation, warps 2..5 epilogue
constexpr int kFc1BlockN = 128, kFc1Stages = 6;
constexpr int kFc2BlockN = 96, kFc2Stages = 3;

enum { EPI_FC1_PARTIAL = 0, EPI_FC2_RESIDUAL = 1 };

struct GemmArgs {
    const int *m_dev;
    int m_host;
    int num_k_blocks;        // ceil(K / 64); the K tail is zero-filled by TMA
    float *out;              // fc1: partial [splits, rows, N]; fc2: logits [M, N]
    size_t split_stride;
    int ldo;
    const float *bias;       // fc2
    const float *residual;   // fc2: x [M, N]
    int skip_le;             // exit when M <= skip_le (those calls are served by prn_fused.cu)
};

struct TensorMaps {
    CUtensorMap a1, b1, a2, b2;
};

template <int BLOCK_N, int STAGES>
struct SmemLayout {
    static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
    static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kBarOffset = STAGES * kStageBytes;
    static constexpr int kTotal = kBarOffset + (2 * STAGES + 1) * 8 + 8 + 1024;   // + alignment slack
    static constexpr int kTmemCols = BLOCK_N <= 32 ? 32 : BLOCK_N <= 64 ? 64 : BLOCK_N <= 128 ? 128 : 256;
};

template <int BLOCK_N, int STAGES, int EPI, int MIN_CTAS>
__global__ void __launch_bounds__(kGemmThreads, MIN_CTAS)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmArgs args)
{
    using L = SmemLayout<BLOCK_N, STAGES>;
    static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "UMMA N for M = 128");
    static_assert(L::kABytes % 1024 == 0 && L::kBBytes % 1024 == 0, "tiles must keep 1024-byte alignment");
    extern __shared__ uint8_t smem_raw[];

    const int M = args.m_dev ? *args.m_dev : args.m_host;
    const int m0 = blockIdx.y * BLOCK_M;
    if (m0 >= M || M <= args.skip_le) return;  // uniform per CTA, before any barrier or TMEM allocation
    const int n0 = blockIdx.x * BLOCK_N;
    const int splits = gridDim.z, z = blockIdx.z;
    const int kb_begin = (int)(((long long)z * args.num_k_blocks) / splits);
    const int kb_end = (int)(((long long)(z + 1) * args.num_k_blocks) / splits);
    const int num_kb = kb_end - kb_begin;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem + L::kBarOffset);
    uint64_t *empty_bar = full_bar + STAGES;
    uint64_t *tmem_full_bar = empty_bar + STAGES;
    uint32_t *tmem_base_slot = reinterpret_cast<uint32_t *>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
        for (int i = 0; i < STAGES; ++i) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_base_slot)),
                     "r"((uint32_t)L::kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_base_slot;

    if (warp == 0) {
        if (lane == 0) {   // ---- TMA producer ----
            for (int i = 0; i < num_kb; ++i) {
                const int st = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(empty_bar + st, ph ^ 1u);
                mbar_arrive_expect_tx(full_bar + st, (uint32_t)L::kStageBytes);
                uint8_t *a_dst = smem + st * L::kStageBytes;
                const int kc = (kb_begin + i) * BLOCK_K;
                tma_load_2d(a_dst, &tmap_a, full_bar + st, kc, m0, kEvictLast);
                tma_load_2d(a_dst + L::kABytes, &tmap_b, full_bar + st, kc, n0, kEvictFirst);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {   // ---- MMA issuer ----
            constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N);
            for (int i = 0; i < num_kb; ++i) {
                const int st = i % STAGES;
                const uint32_t ph = (uint32_t)(i / STAGES) & 1u;
                mbar_wait(full_bar + st, ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(smem + st * L::kStageBytes);
                const uint64_t adesc = make_kmajor_sw128_desc(a_addr);
                const uint64_t bdesc = make_kmajor_sw128_desc(a_addr + L::kABytes);
#pragma unroll
                for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                    // advance 16 elements = 32 bytes along K inside the swizzle row: +2 in the (addr >> 4) field
                    umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                              (i > 0 || k > 0) ? 1u : 0u);
                }
                umma_commit(empty_bar + st);      // slot reusable once these MMAs have read it
            }
            umma_commit(tmem_full_bar);           // accumulator complete
        }
        __syncwarp();
    } else {
        // ---- epilogue: warp w may touch TMEM lanes [32 * (w % 4), +32) ----
        const int q = warp & 3;
        const int row = m0 + q * 32 + lane;
        mbar_wait(tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);
        const bool row_ok = row < M;
#pragma unroll 1
        for (int c = 0; c < BLOCK_N; c += 16) {
            uint32_t r[16];
            tmem_ld16(t_row + (uint32_t)c, r);
            if (row_ok) {
                const int n = n0 + c;
                if (EPI == EPI_FC1_PARTIAL) {
                    float4 *dst = reinterpret_cast<float4 *>(args.out + (size_t)z * args.split_stride +
                                                             (size_t)row * args.ldo + n);
#pragma unroll
                    for (int v = 0; v < 4; ++v)
                        dst[v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                             __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
                } else {
                    const float4 *bias = reinterpret_cast<const float4 *>(args.bias + n);
                    const float4 *res = reinterpret_cast<const float4 *>(args.residual + (size_t)row * args.ldo + n);
                    float4 *dst = reinterpret_cast<float4 *>(args.out + (size_t)row * args.ldo + n);
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float4 b = __ldg(bias + v), x = __ldg(res + v);
                        float4 o;   // x + relu(acc + b2)   (detector/prn.py:22,24)
                        o.x = __fadd_rn(x.x, fmaxf(__fadd_rn(__uint_as_float(r[4 * v]), b.x), 0.0f));
                        o.y = __fadd_rn(x.y, fmaxf(__fadd_rn(__uint_as_float(r[4 * v + 1]), b.y), 0.0f));
                        o.z = __fadd_rn(x.z, fmaxf(__fadd_rn(__uint_as_float(r[4 * v + 2]), b.z), 0.0f));
                        o.w = __fadd_rn(x.w, fmaxf(__fadd_rn(__uint_as_float(r[4 * v + 3]), b.w), 0.0f));
                        dst[v] = o;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)L::kTmemCols) : "memory");
    }
}

template <typename K>
cudaError_t set_smem(K kernel, int bytes)
{
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

}  // namespace

int prn_bf16_prepare(mpn_handle *h)
{
    TensorMaps *tm = new TensorMaps;
    const uint64_t D = (uint64_t)h->D, Hd = (uint64_t)h->cfg.prn_hidden, rows = (uint64_t)h->prn_ws.n_max;
    if (h->D % kFc2BlockN != 0 || h->cfg.prn_hidden % kFc1BlockN != 0) {
        delete tm;
        snprintf(h->err, sizeof(h->err), "bf16 PRN needs D %% %d == 0 and hidden %% %d == 0", kFc2BlockN, kFc1BlockN);
        return MPN_ERR_UNSUPPORTED;
    }
    const bool ok = encode_2d(&tm->a1, h->crops_bf16, rows, D, BLOCK_M) && encode_2d(&tm->b1, h->W1t, Hd, D, kFc1BlockN) &&
                    encode_2d(&tm->a2, h->prn_ws.y1_bf16, rows, Hd, BLOCK_M) && encode_2d(&tm->b2, h->W2t, D, Hd, kFc2BlockN);
    if (!ok) {
        delete tm;
        snprintf(h->err, sizeof(h->err), "cuTensorMapEncodeTiled failed");
        return MPN_ERR_CUDA;
    }
    cudaError_t e = set_smem(gemm_bf16_kernel<kFc1BlockN, kFc1Stages, EPI_FC1_PARTIAL, 1>,
                             SmemLayout<kFc1BlockN, kFc1Stages>::kTotal);
    if (e == cudaSuccess)
        e = set_smem(gemm_bf16_kernel<kFc2BlockN, kFc2Stages, EPI_FC2_RESIDUAL, 2>, SmemLayout<kFc2BlockN, kFc2Stages>::kTotal);
    if (e != cudaSuccess) {
        delete tm;
        snprintf(h->err, sizeof(h->err), "cudaFuncSetAttribute(smem) failed: %s", cudaGetErrorString(e));
        return MPN_ERR_CUDA;
    }
    h->tmaps = tm;
    return MPN_OK;
}

void prn_bf16_release(mpn_handle *h)
{
    if (h->tmaps) delete static_cast<TensorMaps *>(h->tmaps);
    h->tmaps = nullptr;
}

int launch_prn_bf16(const PrnWeights &w, const PrnWorkspace &ws, const float *x_f32, const __nv_bfloat16 *x_bf16,
                    const int *n_dev, int n_host, int n_max, float *logits, void *tmaps, int skip_le, cudaStream_t s)
{
    (void)x_bf16;
    if (n_max <= 0) return 0;
    if (!tmaps) return -(int)cudaErrorInvalidValue;
    const TensorMaps *tm = static_cast<const TensorMaps *>(tmaps);
    const int D = w.D, Hd = w.hidden;
    const int m_tiles = (n_max + BLOCK_M - 1) / BLOCK_M;
    int launches = 0;
    {   // fc1: split-K partial sums, then bias + ReLU + bf16 in the reduce
        const int n_tiles = Hd / kFc1BlockN;
        const int nkb = (D + BLOCK_K - 1) / BLOCK_K;
        int splits = 148 / (n_tiles * m_tiles);
        if (splits < 1) splits = 1;
        if (splits > nkb) splits = nkb;
        while (splits > 1 && (size_t)splits * n_max * Hd > ws.partial_floats) --splits;
        GemmArgs a;
        a.m_dev = n_dev; a.m_host = n_host; a.num_k_blocks = nkb;
        a.out = ws.partial; a.split_stride = (size_t)n_max * Hd; a.ldo = Hd; a.bias = nullptr; a.residual = nullptr;
        a.skip_le = skip_le;
        dim3 grid(n_tiles, m_tiles, splits);
        prof_mark(s, "prn_bf16_fc1");
        gemm_bf16_kernel<kFc1BlockN, kFc1Stages, EPI_FC1_PARTIAL, 1>
            <<<grid, kGemmThreads, SmemLayout<kFc1BlockN, kFc1Stages>::kTotal, s>>>(tm->a1, tm->b1, a);
        ++launches;
        launches += launch_fc1_reduce(ws.partial, splits, a.split_stride, w.b1, Hd, n_dev, n_host, n_max, nullptr,
                                      ws.y1_bf16, skip_le, s);
    }
    {   // fc2 + bias + ReLU + residual
        GemmArgs a;
        a.m_dev = n_dev; a.m_host = n_host; a.num_k_blocks = Hd / BLOCK_K;
        a.out = logits; a.split_stride = 0; a.ldo = D; a.bias = w.b2; a.residual = x_f32; a.skip_le = skip_le;
        dim3 grid(D / kFc2BlockN, m_tiles, 1);
        prof_mark(s, "prn_bf16_fc2");
        gemm_bf16_kernel<kFc2BlockN, kFc2Stages, EPI_FC2_RESIDUAL, 2>
            <<<grid, kGemmThreads, SmemLayout<kFc2BlockN, kFc2Stages>::kTotal, s>>>(tm->a2, tm->b2, a);
        ++launches;
    }
    return launches;
}

}  // namespace mpn
