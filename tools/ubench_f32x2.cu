// Micro-benchmark: throughput of scalar FMUL/FADD vs packed f32x2 (FMUL2/FADD2) on sm_100a, no FMA contraction.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b)
{
    float2 v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = make_float2(a + i + threadIdx.x, b - i);
    float2 m = make_float2(a, b), s = make_float2(b, a);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {  // scalar: 2 FMUL + 2 FADD per element pair
                v[i].x = __fadd_rn(__fmul_rn(v[i].x, m.x), s.x);
                v[i].y = __fadd_rn(__fmul_rn(v[i].y, m.y), s.y);
            } else {          // packed: 1 FMUL2 + 1 FADD2
                v[i] = __fadd2_rn(__fmul2_rn(v[i], m), s);
            }
        }
    }
    float acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
int main()
{
    float* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 20000;
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.9999f); else k<1><<<148 * 8, 256>>>(d, iters, 1.0001f, 0.9999f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double flops = 148.0 * 8 * 256 * (double)iters * 8 * 4;  // lane-ops (mul+add on 2 floats)
            if (rep) printf("mode %d: %.3f ms, %.2f T lane-ops/s (%s)\n", mode, ms, flops / ms / 1e9, mode ? "f32x2" : "scalar");
        }
    }
    return 0;
}
