// Micro-benchmark (development tool, not part of the library): how fast can one B200 stream a 70 MB bf16 weight
// matrix from HBM into shared memory with (a) plain 16-byte loads, (b) contiguous cp.async.bulk copies of pre-tiled
// data, (c) 2-D TMA boxes of 128 rows x 128 bytes out of a row-major [1024, 34272] matrix (what prn_tcgen05.cu does).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/stream_bench tools/stream_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                      \
    do {                                                                                           \
        cudaError_t e_ = (x);                                                                      \
        if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); }       \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        uint32_t done;
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}

// (a) plain loads
__global__ void __launch_bounds__(256) ldg_kernel(const uint4 *__restrict__ src, size_t n16, unsigned *sink)
{
    unsigned acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n16; i += 8 * stride) {
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldg(src + i + j * stride);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc ^= v[j].x ^ v[j].y ^ v[j].z ^ v[j].w;
    }
    for (; i < n16; i += stride) { uint4 v = __ldg(src + i); acc ^= v.x ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}

// (b) contiguous bulk copies: CTA c owns chunks c, c + grid, ...
__global__ void __launch_bounds__(64) bulk_kernel(const uint8_t *__restrict__ src, size_t total, int chunk, int stages,
                                                  unsigned *sink)
{
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)stages * chunk);
    uint64_t *empty = full + stages;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t nchunks = total / chunk;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        int it = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
            const int st = it % stages;
            mbar_wait(empty + st, ((it / stages) & 1) ^ 1);
            mbar_expect(full + st, chunk);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(smem_u32(smem + (size_t)st * chunk)), "l"(src + c * chunk), "r"(chunk), "r"(smem_u32(full + st))
                         : "memory");
        }
    } else if (warp == 1 && lane == 0) {
        int it = 0;
        unsigned acc = 0;
        for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x, ++it) {
            const int st = it % stages;
            mbar_wait(full + st, (it / stages) & 1);
            acc ^= *reinterpret_cast<volatile unsigned *>(smem + (size_t)st * chunk);
            mbar_arrive(empty + st);
        }
        if (acc == 0x12345678u) *sink = acc;
    }
}

// (c) 2-D TMA boxes [box_rows x 64 bf16] of a row-major [rows, cols] matrix; CTA = (row tile, k split)
__global__ void __launch_bounds__(64) tma_kernel(const __grid_constant__ CUtensorMap tmap, int box_rows, int nkb,
                                                 int stages, unsigned *sink)
{
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int bytes = box_rows * 128;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + (size_t)stages * bytes);
    uint64_t *empty = full + stages;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int splits = gridDim.y, z = blockIdx.y;
    const int kb0 = (int)((long long)z * nkb / splits), kb1 = (int)((long long)(z + 1) * nkb / splits);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        for (int i = 0; i < kb1 - kb0; ++i) {
            const int st = i % stages;
            mbar_wait(empty + st, ((i / stages) & 1) ^ 1);
            mbar_expect(full + st, bytes);
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(smem_u32(smem + (size_t)st * bytes)), "l"(reinterpret_cast<uint64_t>(&tmap)),
                           "r"(smem_u32(full + st)), "r"((kb0 + i) * 64), "r"((int)blockIdx.x * box_rows)
                         : "memory");
        }
    } else if (warp == 1 && lane == 0) {
        unsigned acc = 0;
        for (int i = 0; i < kb1 - kb0; ++i) {
            const int st = i % stages;
            mbar_wait(full + st, (i / stages) & 1);
            acc ^= *reinterpret_cast<volatile unsigned *>(smem + (size_t)st * bytes);
            mbar_arrive(empty + st);
        }
        if (acc == 0x12345678u) *sink = acc;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <typename F>
float time_it(F f, int reps, void *flush, size_t flush_bytes)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        // evict the 126 MB L2 with CLEAN lines (a memset would leave it full of dirty lines whose write-back then
        // competes with the measured reads)
        ldg_kernel<<<148 * 8, 256>>>(reinterpret_cast<const uint4 *>(flush), flush_bytes / 16, nullptr);
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main()
{
    const size_t cols = 34272;
    const int copies = getenv("COPIES") ? atoi(getenv("COPIES")) : 1;
    const size_t rows = 1024 * (size_t)copies;
    const size_t bytes = rows * cols * 2;   // 70.2 MB per copy
    uint8_t *w; unsigned *sink; void *flush;
    const size_t flush_bytes = 256u << 20;
    CK(cudaMalloc(&w, bytes + (1 << 20)));
    CK(cudaMalloc(&sink, 4));
    CK(cudaMalloc(&flush, flush_bytes));
    CK(cudaMemset(w, 1, bytes));
    CK(cudaMemset(flush, 2, flush_bytes));
    CK(cudaDeviceSynchronize());
    printf("weights %.1f MB\n", bytes / 1e6);

    for (int per_sm : {2, 4, 8}) {
        float ms = time_it([&] { ldg_kernel<<<148 * per_sm, 256>>>(reinterpret_cast<const uint4 *>(w), bytes / 16, sink); }, 5, flush, flush_bytes);
        printf("ldg   grid=148x%d                       %7.2f us  %7.1f GB/s\n", per_sm, ms * 1e3, bytes / ms / 1e6);
    }
    CK(cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int chunk : {8192, 16384, 32768})
        for (int stages : {2, 4, 6})
            for (int per_sm : {1, 2}) {
                const size_t smem = (size_t)chunk * stages + 16 * stages + 64;
                if (smem * per_sm > 220 * 1024) continue;
                float ms = time_it([&] { bulk_kernel<<<148 * per_sm, 64, smem>>>(w, bytes, chunk, stages, sink); }, 5, flush, flush_bytes);
                printf("bulk  chunk=%5d stages=%d ctas/sm=%d      %7.2f us  %7.1f GB/s\n", chunk, stages, per_sm, ms * 1e3, bytes / ms / 1e6);
            }
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(p);
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    for (int box_rows : {128, 256})
        for (int stages : {3, 6, 10})
            for (int sw : {0, 1}) {
                CUtensorMap tm;
                const cuuint64_t dims[2] = {cols, rows};
                const cuuint64_t strides[1] = {cols * 2};
                const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
                const cuuint32_t el[2] = {1, 1};
                if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, dims, strides, box, el, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); continue; }
                const int tiles = rows / box_rows, nkb = (cols + 63) / 64;
                const size_t smem = (size_t)box_rows * 128 * stages + 16 * stages + 64 + 1024;
                if (smem > 220 * 1024) continue;
                for (int ctas : {148, 296}) {
                    const int splits = ctas * copies / tiles > 0 ? (ctas / (int)(1024 / box_rows)) : 1;
                    dim3 grid(tiles, splits);
                    if ((long long)tiles * splits > 148 * 16) continue;
                    float ms = time_it([&] { tma_kernel<<<grid, 64, smem>>>(tm, box_rows, nkb, stages, sink); }, 5, flush, flush_bytes);
                    printf("tma2d box=%3dx64 stages=%2d swz=%d grid=%dx%d  %7.2f us  %7.1f GB/s\n", box_rows, stages, sw, tiles, splits,
                           ms * 1e3, bytes / ms / 1e6);
                }
            }
    return 0;
}
