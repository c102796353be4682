// latency_probe.cu -- FP64 pipe on B200: dependent-issue latency of DFMA and the pipe utilisation reached with W warps
// per SM partition each carrying C independent DFMA chains (what ILP x occupancy the stepper needs to fill the pipe).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o latency_probe latency_probe.cu && ./latency_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void probe(double* out, int iters, double a, double b, long long* cycles) {
    double f[C];
    for (int i = 0; i < C; ++i) f[i] = threadIdx.x * 1e-3 + i;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
#pragma unroll
            for (int i = 0; i < C; ++i) f[i] = fma(f[i], a, b);
        }
    }
    const long long t1 = clock64();
    double s = 0;
    for (int i = 0; i < C; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int C>
static void run(double* out, long long* cyc, int sms, int warps_per_sm) {
    const int iters = 2000;
    probe<C><<<sms, 32 * warps_per_sm>>>(out, 10, 0.999, 1e-3, cyc);
    cudaDeviceSynchronize();
    probe<C><<<sms, 32 * warps_per_sm>>>(out, iters, 0.999, 1e-3, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
    const double per_round = (double)c / (iters * 16.0);            // cycles per "round" of C DFMAs per warp
    const double wps = warps_per_sm / 4.0;                          // warps per SM partition
    printf("chains %d  warps/SMSP %.0f   %6.2f cycles per dependent step   pipe utilisation %5.1f %%\n", C, wps, per_round,
           100.0 * (C * wps * 2.0) / per_round);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out; long long* cyc;
    cudaMalloc(&out, sizeof(double) * sms * 1024);
    cudaMalloc(&cyc, 8);
    for (int w : {4, 8, 12, 16, 24}) {
        run<1>(out, cyc, sms, w); run<2>(out, cyc, sms, w); run<3>(out, cyc, sms, w); run<4>(out, cyc, sms, w); run<8>(out, cyc, sms, w);
    }
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
