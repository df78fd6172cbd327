// Pose Residual Network, fp32 mode (parity mode): plain FFMA GEMMs with a fixed summation order.
//
// Replaces detector/prn.py:15-25 (slim.fully_connected x2 + residual):
//   y1 = relu(x W1 + b1)   x [N, D=34272], W1 [D, 1024]
//   y2 = relu(y1 W2 + b2)  W2 [1024, D]
//   logits = x + y2
// fc1 has a 34272-long reduction and only 1024 outputs, so it is split along K (deterministic two-pass: partial
// sums to workspace, then a reduce kernel that also applies bias + ReLU); fc2 applies bias, ReLU and the residual
// add in its epilogue.  N (persons) is only known on the device: tiles beyond it exit at once.
//
// This is the 1e-4 path of BASELINE.json ("fp32"); the throughput path is prn_tcgen05.cu (bf16 tensor cores).
#include "common.cuh"

namespace mpn {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, TM = 4, TN = 4;
constexpr int kThreads = (BM / TM) * (BN / TN);   // 256

enum { EPI_PARTIAL = 0, EPI_BIAS_RELU_RESIDUAL = 1 };

// C[M, Nc] (+)= A[M, K-range] * B[K-range, Nc];  A row-major lda, B row-major ldb (= Nc)
template <int EPI>
__global__ void __launch_bounds__(kThreads) sgemm_kernel(const float *__restrict__ A, const int lda,
                                                         const float *__restrict__ Bm, const int Nc,
                                                         const int *__restrict__ m_dev, const int m_host,
                                                         const int k_per_split, const float *__restrict__ bias,
                                                         const float *__restrict__ residual, float *__restrict__ C,
                                                         const size_t split_stride, const int skip_le)
{
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];
    const int M = m_dev ? *m_dev : m_host;
    const int m0 = blockIdx.y * BM;
    if (m0 >= M || M <= skip_le) return;
    const int n0 = blockIdx.x * BN;
    const int k_begin = blockIdx.z * k_per_split, k_end = k_begin + k_per_split;
    const int tid = threadIdx.x, tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int a_row = tid / 4, a_k4 = (tid % 4) * 4;         // A tile: 64 rows x 16 k
    const int b_row = tid / 16, b_c4 = (tid % 16) * 4;       // B tile: 16 k x 64 cols
    const bool a_ok = (m0 + a_row) < M;
    const bool b_ok = (n0 + b_c4) < Nc;                      // Nc % 4 == 0
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;

    const float *a_ptr = A + (size_t)(m0 + a_row) * lda + a_k4;
    const float *b_ptr = Bm + (size_t)b_row * Nc + n0 + b_c4;
    float4 a_reg = make_float4(0.f, 0.f, 0.f, 0.f), b_reg = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a_ok) a_reg = __ldg(reinterpret_cast<const float4 *>(a_ptr + k_begin));
    if (b_ok) b_reg = __ldg(reinterpret_cast<const float4 *>(b_ptr + (size_t)k_begin * Nc));
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        As[a_k4 + 0][a_row] = a_reg.x; As[a_k4 + 1][a_row] = a_reg.y;
        As[a_k4 + 2][a_row] = a_reg.z; As[a_k4 + 3][a_row] = a_reg.w;
        *reinterpret_cast<float4 *>(&Bs[b_row][b_c4]) = b_reg;
        __syncthreads();
        if (k0 + BK < k_end) {   // prefetch the next slab while this one is consumed
            if (a_ok) a_reg = __ldg(reinterpret_cast<const float4 *>(a_ptr + k0 + BK));
            if (b_ok) b_reg = __ldg(reinterpret_cast<const float4 *>(b_ptr + (size_t)(k0 + BK) * Nc));
        }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 av = *reinterpret_cast<const float4 *>(&As[k][ty * TM]);
            const float4 bv = *reinterpret_cast<const float4 *>(&Bs[k][tx * TN]);
            const float a[TM] = {av.x, av.y, av.z, av.w};
            const float b[TN] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    const int col = n0 + tx * TN;
    if (col >= Nc) return;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int row = m0 + ty * TM + i;
        if (row >= M) continue;
        float4 o = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (EPI == EPI_PARTIAL) {
            *reinterpret_cast<float4 *>(C + blockIdx.z * split_stride + (size_t)row * Nc + col) = o;
        } else {
            const float4 bb = __ldg(reinterpret_cast<const float4 *>(bias + col));
            const float4 rr = __ldg(reinterpret_cast<const float4 *>(residual + (size_t)row * Nc + col));
            o.x = __fadd_rn(rr.x, fmaxf(__fadd_rn(o.x, bb.x), 0.0f));   // x + relu(y1 W2 + b2)  (prn.py:22,24)
            o.y = __fadd_rn(rr.y, fmaxf(__fadd_rn(o.y, bb.y), 0.0f));
            o.z = __fadd_rn(rr.z, fmaxf(__fadd_rn(o.z, bb.z), 0.0f));
            o.w = __fadd_rn(rr.w, fmaxf(__fadd_rn(o.w, bb.w), 0.0f));
            *reinterpret_cast<float4 *>(C + (size_t)row * Nc + col) = o;
        }
    }
}

// y1[m, j] = relu(sum_z partial[z, m, j] + b1[j]), z ascending  (prn.py:20)
__global__ void __launch_bounds__(256) fc1_reduce_kernel(const float *__restrict__ partial, const int splits,
                                                         const size_t split_stride, const float *__restrict__ bias,
                                                         const int hidden, const int *__restrict__ m_dev,
                                                         const int m_host, float *__restrict__ y1,
                                                         __nv_bfloat16 *__restrict__ y1_bf16, const int skip_le)
{
    const int M = m_dev ? *m_dev : m_host;
    if (M <= skip_le) return;
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)M * hidden) return;
    float s = partial[i];
    for (int z = 1; z < splits; ++z) s = __fadd_rn(s, partial[(size_t)z * split_stride + i]);
    s = fmaxf(__fadd_rn(s, __ldg(bias + (i % hidden))), 0.0f);
    if (y1) y1[i] = s;
    if (y1_bf16) y1_bf16[i] = __float2bfloat16_rn(s);
}

__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float *__restrict__ x, __nv_bfloat16 *__restrict__ y,
                                                          const int *__restrict__ n_dev, const int n_host,
                                                          const int row_len)
{
    const int N = n_dev ? *n_dev : n_host;
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= (size_t)N * row_len) return;
    const float4 v = __ldg(reinterpret_cast<const float4 *>(x + i));
    __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<unsigned *>(&lo);
    o.y = *reinterpret_cast<unsigned *>(&hi);
    *reinterpret_cast<uint2 *>(y + i) = o;
}

// w [rows, cols] fp32 row-major -> wt [cols, rows] bf16 row-major (one-time weight preparation)
__global__ void __launch_bounds__(256) transpose_to_bf16_kernel(const float *__restrict__ w, const int rows,
                                                                const int cols, __nv_bfloat16 *__restrict__ wt)
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
        if (c < cols && r < rows) wt[(size_t)c * rows + r] = __float2bfloat16_rn(tile[tx][i]);
    }
}

int pick_splits(int k_slabs, int ctas_without_split)
{
    // largest divisor of the slab count that keeps the grid within ~2 waves of 148 SMs
    int best = 1;
    for (int s = 1; s <= k_slabs && s <= 64; ++s)
        if (k_slabs % s == 0 && (long long)ctas_without_split * s <= 2 * 148 * 2) best = s;
    return best;
}

}  // namespace

int launch_prn_fp32(const PrnWeights &w, const PrnWorkspace &ws, const float *x, const int *n_dev, int n_host,
                    int n_max, float *logits, int skip_le, cudaStream_t s)
{
    if (n_max <= 0) return 0;
    const int D = w.D, Hd = w.hidden;
    const int m_tiles = (n_max + BM - 1) / BM;
    int launches = 0;
    // fc1: split-K
    int splits = pick_splits(D / BK, (Hd / BN) * m_tiles);
    while (splits > 1 && (size_t)splits * n_max * Hd > ws.partial_floats) --splits;
    while ((D / BK) % splits != 0) --splits;
    const size_t split_stride = (size_t)n_max * Hd;
    {
        dim3 grid(Hd / BN, m_tiles, splits);
        prof_mark(s, "prn_fp32_fc1");
        sgemm_kernel<EPI_PARTIAL><<<grid, kThreads, 0, s>>>(x, D, w.W1, Hd, n_dev, n_host, D / splits, nullptr,
                                                            nullptr, ws.partial, split_stride, skip_le);
        ++launches;
        launches += launch_fc1_reduce(ws.partial, splits, split_stride, w.b1, Hd, n_dev, n_host, n_max, ws.y1, nullptr, skip_le, s);
    }
    // fc2 + bias + ReLU + residual
    {
        dim3 grid((D + BN - 1) / BN, m_tiles, 1);
        prof_mark(s, "prn_fp32_fc2");
        sgemm_kernel<EPI_BIAS_RELU_RESIDUAL><<<grid, kThreads, 0, s>>>(ws.y1, Hd, w.W2, D, n_dev, n_host, Hd, w.b2, x,
                                                                       logits, 0, skip_le);
        ++launches;
    }
    return launches;
}

int launch_fc1_reduce(const float *partial, int splits, size_t split_stride, const float *bias, int hidden,
                      const int *m_dev, int m_host, int m_max, float *y1, __nv_bfloat16 *y1_bf16, int skip_le,
                      cudaStream_t s)
{
    const size_t total = (size_t)m_max * hidden;
    prof_mark(s, "prn_fc1_reduce");
    fc1_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(partial, splits, split_stride, bias, hidden, m_dev,
                                                                       m_host, y1, y1_bf16, skip_le);
    return 1;
}

int launch_f32_to_bf16(const float *x, __nv_bfloat16 *y, const int *n_rows_dev, int n_rows_host, int row_len,
                       int n_rows_max, cudaStream_t s)
{
    if (n_rows_max <= 0) return 0;
    const size_t total4 = ((size_t)n_rows_max * row_len + 3) / 4;
    prof_mark(s, "f32_to_bf16");
    f32_to_bf16_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, s>>>(x, y, n_rows_dev, n_rows_host, row_len);
    return 1;
}

int launch_transpose_to_bf16(const float *w, int rows, int cols, __nv_bfloat16 *wt, cudaStream_t s)
{
    dim3 grid((cols + 31) / 32, (rows + 31) / 32);
    transpose_to_bf16_kernel<<<grid, 256, 0, s>>>(w, rows, cols, wt);
    return 1;
}

}  // namespace mpn
