// dmma_peak.cu -- measures the FP64 tensor-pipe (DMMA, mma.sync.m8n8k4.f64) and the plain FP64 FMA peak of this GPU.
// The roofline denominators for the reduced-camera solve of lba.cu (SURVEY.md section 8d asks for a MEASURED FP64 peak).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_peak tools/dmma_peak.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dmma_kernel(double* out, int iters)
{
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[2 * i]), "+d"(c[2 * i + 1]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    if (s == 123.456) out[0] = s;
}

__global__ void dfma_kernel(double* out, int iters)
{
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; i++) c[i] = i;
    const double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += c[i];
    if (s == 123.456) out[0] = s;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d;
    cudaMalloc(&d, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int threads = 128; threads <= 1024; threads *= 2) {
        const int ctas = sms * (2048 / threads);
        float best_mma = 1e30f, best_fma = 1e30f;
        for (int rep = 0; rep < 5; rep++) {
            float ms;
            cudaEventRecord(e0); dmma_kernel<<<ctas, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best_mma) best_mma = ms;
            cudaEventRecord(e0); dfma_kernel<<<ctas, threads>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
            cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best_fma) best_fma = ms;
        }
        const double warps = (double)ctas * threads / 32;
        const double mma_flop = warps * iters * 8.0 * 512.0;                 // m8n8k4: 8*8*4 FMA = 512 flop per warp instruction
        const double fma_flop = (double)ctas * threads * iters * 16.0 * 2.0;
        printf("{\"threads\": %d, \"ctas\": %d, \"dmma_tflops\": %.3f, \"dfma_tflops\": %.3f}\n", threads, ctas,
               mma_flop / best_mma * 1e-9, fma_flop / best_fma * 1e-9);
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
