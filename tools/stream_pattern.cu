// stream_pattern.cu -- how close can the step kernel's DATA MOVEMENT get to the copy peak?
//
// Movement-only model of step_tma_kernel (no arithmetic): persistent CTAs, a ring of shared-memory stages, one
// producer lane issuing bulk-async copies, the "compute" warps only hand the stage back.  Per 128-env tile the
// real kernel reads ~27 KB and writes ~17 KB spread over ~60 row streams of the SoA state ([field, ld] arrays,
// 512-byte segments at a 4 MB pitch).  This tool moves the same bytes in three ways:
//   soa2d  : one 2-D tensor-map box {TILE, RD_ROWS} in, one {TILE, WR_ROWS} out            (what the kernel does)
//   soa1d  : one 1-D bulk copy per row segment                                             (op-count sensitivity)
//   tile   : tile-major layout, ONE contiguous run per tile in and out                     (AoSoA alternative)
// and prints achieved GB/s for each, to decide whether a tile-major state layout is worth a redesign.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/stream_pattern tools/stream_pattern.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int RD_ROWS = 54;      // 4-byte rows read per env   (216 B; the kernel reads 236 B incl. 1-byte rows)
constexpr int WR_ROWS = 34;      // 4-byte rows written per env (136 B; the kernel writes ~140-174 B)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(src) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct Maps { CUtensorMap in, out; };

// MODE 0 = soa2d, 1 = soa1d, 2 = tile-major.  The stage holds RD_ROWS x TILE floats; the first WR_ROWS rows go back.
template <int MODE, int TILE, int STAGES>
__global__ void __launch_bounds__(160) move_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t ld,
                                                   const __grid_constant__ Maps maps, int num_tiles) {
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int STAGE_BYTES = RD_ROWS * TILE * 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&bars[s]), 1); mbar_init(smem_u32(&bars[STAGES + s]), 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int my_tiles = (num_tiles > (int)blockIdx.x) ? (num_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    if (tid >= 128) {
        if (tid == 128) {
            auto issue_stores = [&](int k) {
                const int64_t t = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
                const uint32_t sb = smem_u32(smem + (size_t)(k % STAGES) * STAGE_BYTES);
                if (MODE == 0) tma_store_2d(&maps.out, (int)(t * TILE), 0, sb);
                else if (MODE == 1) { for (int r = 0; r < WR_ROWS; ++r) bulk_store(dst + r * ld + t * TILE, sb + r * TILE * 4, TILE * 4); }
                else bulk_store(dst + t * (int64_t)WR_ROWS * TILE, sb, WR_ROWS * TILE * 4);
                bulk_commit();
            };
            for (int k = 0; k < my_tiles; ++k) {
                const int s = k % STAGES, use = k / STAGES;
                const uint32_t sb = smem_u32(smem + (size_t)s * STAGE_BYTES), full = smem_u32(&bars[s]);
                if (use > 0) {
                    mbar_wait(smem_u32(&bars[STAGES + s]), (uint32_t)((use - 1) & 1));
                    issue_stores(k - STAGES);
                    bulk_wait_read0();
                }
                const int64_t t = (int64_t)blockIdx.x + (int64_t)k * gridDim.x;
                mbar_expect_tx(full, STAGE_BYTES);
                if (MODE == 0) tma_load_2d(sb, &maps.in, (int)(t * TILE), 0, full);
                else if (MODE == 1) { for (int r = 0; r < RD_ROWS; ++r) bulk_load(sb + r * TILE * 4, src + r * ld + t * TILE, TILE * 4, full); }
                else bulk_load(sb, src + t * (int64_t)RD_ROWS * TILE, STAGE_BYTES, full);
            }
            const int first = my_tiles > STAGES ? my_tiles - STAGES : 0;
            for (int k = first; k < my_tiles; ++k) {
                mbar_wait(smem_u32(&bars[STAGES + k % STAGES]), (uint32_t)((k / STAGES) & 1));
                issue_stores(k);
            }
            bulk_wait0();
        }
    } else {
        for (int k = 0; k < my_tiles; ++k) {
            const int s = k % STAGES, use = k / STAGES;
            float* sp = reinterpret_cast<float*>(smem + (size_t)s * STAGE_BYTES);
            mbar_wait(smem_u32(&bars[s]), (uint32_t)(use & 1));
            for (int c = tid; c < TILE; c += 128) sp[c] = sp[c] + 1.0f;            // touch the tile
            fence_async_smem();
            mbar_arrive(smem_u32(&bars[STAGES + s]));
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static void make_map(EncodeTiledFn enc, CUtensorMap* m, void* base, int64_t n, int64_t ld, int rows, int tile) {
    const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)tile, (cuuint32_t)rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("tensor map encode failed: %d\n", (int)r); exit(1); }
}

template <int MODE, int TILE, int STAGES>
static void run(const char* name, EncodeTiledFn enc, float* src, float* dst, int64_t n, int sm_count) {
    Maps maps;
    memset(&maps, 0, sizeof(maps));
    make_map(enc, &maps.in, src, n, n, RD_ROWS, TILE);
    make_map(enc, &maps.out, dst, n, n, WR_ROWS, TILE);
    auto kern = move_kernel<MODE, TILE, STAGES>;
    const size_t smem = (size_t)STAGES * RD_ROWS * TILE * 4 + 2 * STAGES * 8;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 160, smem));
    const int num_tiles = (int)(n / TILE);
    const int grid = sm_count * per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int w = 0; w < 5; ++w) kern<<<grid, 160, smem>>>(src, dst, n, maps, num_tiles);
    CK(cudaDeviceSynchronize());
    const int reps = 50;
    CK(cudaEventRecord(e0));
    for (int r = 0; r < reps; ++r) kern<<<grid, 160, smem>>>(src, dst, n, maps, num_tiles);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double us = ms * 1e3 / reps;
    const double bytes = (double)n * (RD_ROWS + WR_ROWS) * 4;
    printf("%-6s tile %3d stages %d ctas/sm %d : %7.1f us  %7.0f GB/s\n", name, TILE, STAGES, per_sm, us, bytes / us * 1e-3);
}

__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, int64_t n4) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) b[i] = a[i];
}

int main(int argc, char** argv) {
    const int64_t n = argc > 1 ? atoll(argv[1]) : (1 << 20);
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(sym);
    float *src, *dst;
    CK(cudaMalloc(&src, (size_t)n * RD_ROWS * 4));
    CK(cudaMalloc(&dst, (size_t)n * RD_ROWS * 4));
    CK(cudaMemset(src, 0, (size_t)n * RD_ROWS * 4));
    CK(cudaMemset(dst, 0, (size_t)n * RD_ROWS * 4));
    printf("%s, %d SMs, %lld envs, %d B read + %d B written per env\n", prop.name, prop.multiProcessorCount, (long long)n,
           RD_ROWS * 4, WR_ROWS * 4);
    {   // plain copy of the same number of bytes split 50/50, for the peak on this box
        const int64_t n4 = n * (RD_ROWS + WR_ROWS) / 2 / 4;
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        for (int w = 0; w < 3; ++w) copy_kernel<<<prop.multiProcessorCount * 8, 512>>>((const float4*)src, (float4*)dst, n4);
        CK(cudaEventRecord(e0));
        for (int r = 0; r < 20; ++r) copy_kernel<<<prop.multiProcessorCount * 8, 512>>>((const float4*)src, (float4*)dst, n4);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("copy   (float4 grid-stride, read == write)      : %7.1f us  %7.0f GB/s\n", ms * 1e3 / 20, (double)n4 * 32 / (ms * 1e3 / 20) * 1e-3);
    }
    for (int rep = 0; rep < 2; ++rep) {
        run<0, 128, 2>("soa2d", enc, src, dst, n, prop.multiProcessorCount);
        run<0, 128, 3>("soa2d", enc, src, dst, n, prop.multiProcessorCount);
        run<0, 256, 2>("soa2d", enc, src, dst, n, prop.multiProcessorCount);
        run<0, 64, 4>("soa2d", enc, src, dst, n, prop.multiProcessorCount);
        run<1, 128, 2>("soa1d", enc, src, dst, n, prop.multiProcessorCount);
        run<2, 128, 2>("tile", enc, src, dst, n, prop.multiProcessorCount);
        run<2, 128, 3>("tile", enc, src, dst, n, prop.multiProcessorCount);
        run<2, 256, 2>("tile", enc, src, dst, n, prop.multiProcessorCount);
        run<2, 64, 4>("tile", enc, src, dst, n, prop.multiProcessorCount);
    }
    return 0;
}
