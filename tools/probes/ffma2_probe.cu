// Issue-rate probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.  Each thread runs 8 independent chains.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float a, float b)
{
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = threadIdx.x * 1e-3f + i;
    if (MODE == 0) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 16; i++) x[i] = fmaf(x[i], a, b);
        }
    } else {
        unsigned long long p[8], pa, pb;
        float2 fa = make_float2(a, a), fb = make_float2(b, b);
        pa = *reinterpret_cast<unsigned long long*>(&fa); pb = *reinterpret_cast<unsigned long long*>(&fb);
#pragma unroll
        for (int i = 0; i < 8; i++) { float2 t = make_float2(x[2 * i], x[2 * i + 1]); p[i] = *reinterpret_cast<unsigned long long*>(&t); }
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(pa), "l"(pb));
        }
#pragma unroll
        for (int i = 0; i < 8; i++) { float2 t = *reinterpret_cast<float2*>(&p[i]); x[2 * i] = t.x; x[2 * i + 1] = t.y; }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int mode = 0; mode < 2; mode++) {
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) probe<0><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f); else probe<1><<<148 * 8, 256>>>(out, iters, 0.999f, 0.001f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double fma = 148.0 * 8 * 256 * 16.0 * iters;
            printf("mode %s: %.3f ms, %.2f Tfma/s (%.1f TFLOP/s)\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms * 1e-9, 2 * fma / ms * 1e-9);
        }
    }
    return 0;
}
