// issue_probe.cu -- does a DFMA block the warp scheduler's issue slot for both of its two pipe cycles on B200?
// Loop body: 32 independent DFMAs + NI independent integer (or shared-memory) instructions.  If the FP64 pipe only needs
// every other issue slot, time stays flat until NI ~ 32; if time grows from NI = 1, every non-FP64 instruction costs.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o issue_probe issue_probe.cu && ./issue_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int NI, int KIND>   // KIND 0: IMAD (FMA pipe), 1: LOP3 (ALU pipe), 2: LDS.128 + 2 LOP3
__global__ void __launch_bounds__(128) probe(double* out, int iters, double a, double b, int m) {
    __shared__ double2 sm[256];
    sm[threadIdx.x] = make_double2(a, b); sm[threadIdx.x + 128] = make_double2(b, a);
    __syncthreads();
    double f[8];
    unsigned u[8];
    double2 acc = make_double2(0, 0);
    for (int i = 0; i < 8; ++i) { f[i] = threadIdx.x * 1e-3 + i; u[i] = threadIdx.x * 2654435761u + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
#pragma unroll
            for (int i = 0; i < NI / 4; ++i) {
                if (KIND == 0) u[i % 8] = u[i % 8] * (unsigned)m + u[(i + 1) % 8];             // one IMAD
                else if (KIND == 1) u[i % 8] = (u[i % 8] ^ u[(i + 1) % 8]) & (unsigned)m;      // one LOP3
                else { const double2 v = sm[(u[0] + i * 7 + r * 13 + it) & 255]; acc.x += 0; u[i % 8] ^= __double2loint(v.x) ^ __double2hiint(v.y); }
            }
        }
    }
    double s = acc.x;
    for (int i = 0; i < 8; ++i) s += f[i] + u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NI, int KIND>
static void run(double* out, int blocks) {
    const int iters = 4000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<NI, KIND><<<blocks, 128>>>(out, 100, 0.999, 1e-3, 12345);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<NI, KIND><<<blocks, 128>>>(out, iters, 0.999, 1e-3, 12345);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("kind %d  extra/32 DFMA = %2d   %8.3f ms   DFMA %6.2f TF\n", KIND, NI, ms, (double)blocks * 128 * iters * 64.0 / ms / 1e9);
}

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* out;
    const int blocks = sms * 3;    // 12 warps per SM like the stepper
    cudaMalloc(&out, sizeof(double) * blocks * 128);
    run<0, 0>(out, blocks); run<8, 0>(out, blocks); run<16, 0>(out, blocks); run<32, 0>(out, blocks); run<64, 0>(out, blocks);
    run<8, 1>(out, blocks); run<16, 1>(out, blocks); run<32, 1>(out, blocks);
    run<4, 2>(out, blocks); run<8, 2>(out, blocks); run<16, 2>(out, blocks);
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
