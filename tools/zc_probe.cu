// Probe (experiment): how fast can SMs write / read mapped page-locked host memory on this box, compared with the copy
// engine?  Decides whether a step kernel that stores its results straight into host memory can beat chunked DMA.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/zc_probe tools/zc_probe.cu && tools/zc_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void st_v4(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// one elected thread per CTA: global -> shared (bulk load), shared -> global (bulk store), CHUNK bytes at a time
template <int CHUNK>
__global__ void bulk_copy(const char* __restrict__ src, char* __restrict__ dst, size_t bytes) {
    extern __shared__ __align__(128) char sm[];
    __shared__ uint64_t bar;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm), bb = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        uint32_t phase = 0;
        for (size_t off = (size_t)blockIdx.x * CHUNK; off + CHUNK <= bytes; off += (size_t)gridDim.x * CHUNK) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sb), "l"(src + off), "r"(CHUNK), "r"(bb) : "memory");
            uint32_t ok = 0;
            while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                     : "=r"(ok) : "r"(bb), "r"(phase) : "memory");
            phase ^= 1;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(sb), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

// the step kernel's pattern: per 128-env tile, ROWS row segments of 512 B, `ld4` bytes apart (SoA rows of the observation)
template <int ROWS>
__global__ void bulk_rows(const char* __restrict__ src, char* __restrict__ dst, int tiles, size_t ld4) {
    extern __shared__ __align__(128) char sm[];
    __shared__ uint64_t bar;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(sm), bb = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bb));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        uint32_t phase = 0;
        for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(ROWS * 512) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(sb), "l"(src + (size_t)t * ROWS * 512), "r"(ROWS * 512), "r"(bb) : "memory");
            uint32_t ok = 0;
            while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                                     : "=r"(ok) : "r"(bb), "r"(phase) : "memory");
            phase ^= 1;
            for (int r = 0; r < ROWS; ++r)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(dst + (size_t)r * ld4 + (size_t)t * 512), "r"(sb + r * 512), "r"(512) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main() {
    const size_t bytes = 128u << 20;
    char *d, *d2, *h, *hd;
    CK(cudaMalloc(&d, bytes)); CK(cudaMalloc(&d2, bytes));
    CK(cudaHostAlloc(&h, bytes, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&hd, h, 0));
    CK(cudaMemset(d, 1, bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto report = [&](const char* what, float ms) { printf("%-44s %7.3f ms  %6.1f GB/s\n", what, ms, bytes / ms / 1e6); };
    float ms;
    for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep) report("copy engine D2H", ms);
        CK(cudaEventRecord(e0)); CK(cudaMemcpyAsync(d2, h, bytes, cudaMemcpyHostToDevice)); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep) report("copy engine H2D", ms);
    }
    for (int grid : {148, 444, 1184}) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0)); st_v4<<<grid, 256>>>((const float4*)d, (float4*)hd, bytes / 16); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            char w[80]; snprintf(w, 80, "SM st.v4 device->host, grid %d x 256", grid); if (rep) report(w, ms);
        }
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0)); st_v4<<<grid, 256>>>((const float4*)hd, (float4*)d2, bytes / 16); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            char w[80]; snprintf(w, 80, "SM ld.v4 host->device, grid %d x 256", grid); if (rep) report(w, ms);
        }
    }
    for (int grid : {148, 444, 888}) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0)); bulk_copy<16384><<<grid, 32, 16384>>>(d, hd, bytes); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            char w[80]; snprintf(w, 80, "bulk store 16 KB device->host, grid %d", grid); if (rep) report(w, ms);
        }
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0)); bulk_copy<16384><<<grid, 32, 16384>>>(hd, d2, bytes); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            char w[80]; snprintf(w, 80, "bulk load 16 KB host->device, grid %d", grid); if (rep) report(w, ms);
        }
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0)); bulk_copy<512><<<grid, 32, 16384>>>(d, hd, bytes); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaEventElapsedTime(&ms, e0, e1));
            char w[80]; snprintf(w, 80, "bulk store 512 B device->host, grid %d", grid); if (rep) report(w, ms);
        }
    }
    {
        // 32 rows x 1M envs x 4 B = 128 MB, row segments 4 MB apart: the SoA observation as the step kernel would write it
        const int tiles = (1 << 20) / 128;
        for (int grid : {148, 444}) {
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0)); bulk_rows<32><<<grid, 32, 32 * 512>>>(d, hd, tiles, (size_t)4 << 20); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1));
                char w[80]; snprintf(w, 80, "SoA rows: 32 x 512 B per tile -> host, grid %d", grid); if (rep) report(w, ms);
            }
            for (int rep = 0; rep < 2; ++rep) {
                CK(cudaEventRecord(e0)); bulk_rows<32><<<grid, 32, 32 * 512>>>(d, d2, tiles, (size_t)4 << 20); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                CK(cudaEventElapsedTime(&ms, e0, e1));
                char w[80]; snprintf(w, 80, "SoA rows: 32 x 512 B per tile -> HBM, grid %d", grid); if (rep) report(w, ms);
            }
        }
    }
    // both directions at once: copy engine H2D while SMs store to the host
    {
        cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
        CK(cudaDeviceSynchronize());
        CK(cudaEventRecord(e0, s1));
        CK(cudaMemcpyAsync(d2, h + (64u << 20), 64u << 20, cudaMemcpyHostToDevice, s2));
        bulk_copy<16384><<<444, 32, 16384, s1>>>(d, hd, 64u << 20);
        CK(cudaStreamSynchronize(s2));
        CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("duplex: CE H2D 64 MB + bulk store 64 MB       %7.3f ms\n", ms);
        CK(cudaEventRecord(e0, s1));
        CK(cudaMemcpyAsync(d2, h + (64u << 20), 64u << 20, cudaMemcpyHostToDevice, s2));
        CK(cudaMemcpyAsync(h, d, 64u << 20, cudaMemcpyDeviceToHost, s1));
        CK(cudaStreamSynchronize(s2));
        CK(cudaEventRecord(e1, s1)); CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("duplex: CE H2D 64 MB + CE D2H 64 MB           %7.3f ms\n", ms);
    }
    return 0;
}
