// f32x2_probe.cu -- can Blackwell's packed FP32 (add/mul.rn.f32x2, SASS FADD2 / FMUL2) carry the joint update of env_step
// without changing a bit?  The reference rounds every product and sum separately (dexsim_core.cuh), so a packed form is
// only usable if ptxas keeps mul and add apart.  This probe computes
//     jv' = 0.9 jv + 0.1 a ;  jp' = jp + 0.01 jv'
// three ways -- scalar __fmul_rn / __fadd_rn (the product's form), packed mul + packed add, packed mul + scalar add --
// and counts the lanes whose bits differ from the scalar form.
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o tools/f32x2_probe tools/f32x2_probe.cu && ./tools/f32x2_probe
//     cuobjdump -sass tools/f32x2_probe | grep -E "FFMA2|FMUL2|FADD2"
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack2(float x, float y) {
    return (unsigned long long)__float_as_uint(x) | ((unsigned long long)__float_as_uint(y) << 32);
}
__device__ __forceinline__ float lo(unsigned long long v) { return __uint_as_float((unsigned)v); }
__device__ __forceinline__ float hi(unsigned long long v) { return __uint_as_float((unsigned)(v >> 32)); }
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

__global__ void probe(const float* __restrict__ jv, const float* __restrict__ a, const float* __restrict__ jp, int n,
                      unsigned long long* diff_packed, unsigned long long* diff_mixed) {
    const int i = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    if (i + 1 >= n) return;
    // scalar reference form (what env_step does)
    float r_jv[2], r_jp[2];
    for (int k = 0; k < 2; ++k) {
        r_jv[k] = __fadd_rn(__fmul_rn(0.9f, jv[i + k]), __fmul_rn(0.1f, a[i + k]));
        r_jp[k] = __fadd_rn(jp[i + k], __fmul_rn(r_jv[k], 0.01f));
    }
    // packed mul + packed add
    const unsigned long long pjv = pack2(jv[i], jv[i + 1]), pa = pack2(a[i], a[i + 1]), pjp = pack2(jp[i], jp[i + 1]);
    const unsigned long long v2 = add2(mul2(pack2(0.9f, 0.9f), pjv), mul2(pack2(0.1f, 0.1f), pa));
    const unsigned long long p2 = add2(pjp, mul2(v2, pack2(0.01f, 0.01f)));
    // packed mul, scalar adds
    const unsigned long long m1 = mul2(pack2(0.9f, 0.9f), pjv), m2 = mul2(pack2(0.1f, 0.1f), pa);
    const float x_jv0 = __fadd_rn(lo(m1), lo(m2)), x_jv1 = __fadd_rn(hi(m1), hi(m2));
    const unsigned long long m3 = mul2(pack2(x_jv0, x_jv1), pack2(0.01f, 0.01f));
    const float x_jp0 = __fadd_rn(jp[i], lo(m3)), x_jp1 = __fadd_rn(jp[i + 1], hi(m3));
    unsigned long long dp = 0, dm = 0;
    dp += __float_as_uint(lo(v2)) != __float_as_uint(r_jv[0]);
    dp += __float_as_uint(hi(v2)) != __float_as_uint(r_jv[1]);
    dp += __float_as_uint(lo(p2)) != __float_as_uint(r_jp[0]);
    dp += __float_as_uint(hi(p2)) != __float_as_uint(r_jp[1]);
    dm += __float_as_uint(x_jv0) != __float_as_uint(r_jv[0]);
    dm += __float_as_uint(x_jv1) != __float_as_uint(r_jv[1]);
    dm += __float_as_uint(x_jp0) != __float_as_uint(r_jp[0]);
    dm += __float_as_uint(x_jp1) != __float_as_uint(r_jp[1]);
    if (dp) atomicAdd(diff_packed, dp);
    if (dm) atomicAdd(diff_mixed, dm);
}

int main() {
    const int n = 1 << 22;
    float *h = (float*)malloc(3 * n * sizeof(float)), *d;
    srand(7);
    for (int i = 0; i < 3 * n; ++i) h[i] = 2.0f * rand() / RAND_MAX - 1.0f;
    cudaMalloc(&d, 3 * n * sizeof(float));
    cudaMemcpy(d, h, 3 * n * sizeof(float), cudaMemcpyHostToDevice);
    unsigned long long* cnt;
    cudaMalloc(&cnt, 16);
    cudaMemset(cnt, 0, 16);
    probe<<<n / 2 / 256, 256>>>(d, d + n, d + 2 * n, n, cnt, cnt + 1);
    unsigned long long out[2];
    cudaMemcpy(out, cnt, 16, cudaMemcpyDeviceToHost);
    printf("values compared: %d x 2 per form\npacked mul + packed add : %llu differ from the separately rounded form\n"
           "packed mul + scalar add : %llu differ\n", 2 * n, out[0], out[1]);
    return cudaGetLastError() != cudaSuccess;
}
