// dmma_probe.cu -- does the FP64 tensor-core MMA (mma.sync.m8n8k4.f64) on B200 run concurrently with the FP64 vector
// pipe (DFMA), and at what rate?  Decides whether the network product V = Z I belongs on DMMA.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NF, int NM>   // per iteration: NF independent DFMA chains x 4, NM independent DMMA accumulators
__global__ void __launch_bounds__(256) probe(double* out, int iters, double a, double b) {
    double f[NF > 0 ? NF : 1];
    double m0[NM > 0 ? NM : 1], m1[NM > 0 ? NM : 1];
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) f[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) { m0[i] = threadIdx.x * 1e-3; m1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < NF; ++i) f[i] = fma(f[i], a, b);
        }
#pragma unroll
        for (int i = 0; i < NM; ++i) dmma(m0[i], m1[i], a, b);
    }
    double s = 0;
    for (int i = 0; i < (NF > 0 ? NF : 1); ++i) s += f[i];
    for (int i = 0; i < (NM > 0 ? NM : 1); ++i) s += m0[i] + m1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NF, int NM>
static void run(const char* name, double* out, int blocks) {
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<NF, NM><<<blocks, 256>>>(out, 100, 0.999, 1e-3);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<NF, NM><<<blocks, 256>>>(out, iters, 0.999, 1e-3);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double threads = (double)blocks * 256;
    const double fma_flops = threads * iters * 4.0 * NF * 2;
    const double mma_flops = threads / 32 * iters * (double)NM * 8 * 8 * 4 * 2;
    printf("%-28s %8.3f ms   DFMA %7.2f TF   DMMA %7.2f TF   sum %7.2f TF\n", name, ms, fma_flops / ms / 1e9,
           mma_flops / ms / 1e9, (fma_flops + mma_flops) / ms / 1e9);
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double* out;
    const int blocks = sms * 4;
    cudaMalloc(&out, sizeof(double) * blocks * 256);
    run<8, 0>("DFMA only (8 chains x4)", out, blocks);
    run<0, 8>("DMMA only (8 accumulators)", out, blocks);
    run<0, 4>("DMMA only (4 accumulators)", out, blocks);
    run<8, 8>("DFMA 32 + DMMA 8 per iter", out, blocks);
    run<8, 4>("DFMA 32 + DMMA 4 per iter", out, blocks);
    run<8, 2>("DFMA 32 + DMMA 2 per iter", out, blocks);
    run<8, 1>("DFMA 32 + DMMA 1 per iter", out, blocks);
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
