// Micro-benchmark: fp64 pipe of one SM on sm_100a - dependent-issue latency and throughput of DFMA
// as a function of resident warps.  Decides how the EKF's serial factorizations are laid out.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_pipe fp64_pipe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CHAINS>
__global__ void dfma_kernel(double *out, long long *cyc, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) x[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) x[i] = fma(x[i], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void lds_bar_kernel(double *out, long long *cyc, int iters) {
    __shared__ double sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    double v = 0;
    for (int it = 0; it < iters; ++it) {
        if (threadIdx.x == (it & 31)) sh[it & 63] = v + 1.0;
        __syncthreads();
        v += sh[it & 63];
    }
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *out; long long *cyc;
    cudaMalloc(&out, sizeof(double) * 4096); cudaMalloc(&cyc, sizeof(long long) * 8);
    long long h;
    const int iters = 4096;
    for (int threads : {32, 64, 128, 192, 256, 512, 1024}) {
        dfma_kernel<1><<<1, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double lat = (double)h / iters;
        dfma_kernel<8><<<1, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double thr = (double)h / (iters * 8.0);
        printf("threads %4d: dependent DFMA %.1f cyc/op; 8 chains: %.2f cyc per warp-instr per warp -> %.1f lanes/clk/SM\n", threads, lat, thr,
               threads / thr);
    }
    for (int threads : {32, 192, 256}) {
        lds_bar_kernel<<<1, threads>>>(out, cyc, iters);
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads %4d: STS + BAR + LDS round trip %.1f cyc\n", threads, (double)h / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
