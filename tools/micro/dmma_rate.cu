// Micro-benchmark: fp64 tensor pipe (DMMA) of sm_100a - warp-instructions per clock per SM for the
// mma.sync f64 shapes, as a function of resident warps and independent accumulator chains.  Decides the
// warp tile of the EKF update GEMMs (backend.cu upd_gemm).
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_rate dmma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__device__ __forceinline__ void mma(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    if (SHAPE == 0) {  // m8n8k4
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a[0]), "d"(b[0]));
    } else if (SHAPE == 1) {  // m16n8k4
        asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
    } else if (SHAPE == 2) {  // m16n8k8
        asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
    } else {  // m16n8k16
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                     : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                     : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
    }
}

template <int SHAPE, int CHAINS>
__global__ void k(double *out, long long *cyc, int iters) {
    double c[CHAINS][4], a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = 1e-3 * (threadIdx.x + i);
    for (int i = 0; i < 4; ++i) b[i] = 1e-3 * (threadIdx.x - i);
    for (int n = 0; n < CHAINS; ++n)
        for (int i = 0; i < 4; ++i) c[n][i] = n + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < CHAINS; ++n) mma<SHAPE>(c[n], a, b);
    }
    long long t1 = clock64();
    double s = 0;
    for (int n = 0; n < CHAINS; ++n)
        for (int i = 0; i < 4; ++i) s += c[n][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int SHAPE, int CHAINS>
void run(const char *name, double flops_per_instr, double *out, long long *cyc) {
    const int iters = 2048;
    for (int threads : {32, 128, 256, 512}) {
        k<SHAPE, CHAINS><<<1, threads>>>(out, cyc, iters);
        long long h;
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double per_sm = (double)iters * CHAINS * (threads / 32) * flops_per_instr / (double)h;
        printf("%-9s chains %d warps %2d: %7.1f cyc per instr per warp, %7.1f flop/clk/SM -> %.1f TFLOP/s at 148 SMs x 1.965 GHz\n", name, CHAINS,
               threads / 32, (double)h / (iters * CHAINS), per_sm, per_sm * 148 * 1.965e9 / 1e12);
    }
}

int main() {
    double *out;
    long long *cyc;
    cudaMalloc(&out, sizeof(double) * 4096);
    cudaMalloc(&cyc, sizeof(long long) * 8);
    run<0, 1>("m8n8k4", 512, out, cyc);
    run<0, 8>("m8n8k4", 512, out, cyc);
    run<1, 8>("m16n8k4", 1024, out, cyc);
    run<2, 8>("m16n8k8", 2048, out, cyc);
    run<3, 8>("m16n8k16", 4096, out, cyc);
    run<3, 2>("m16n8k16", 4096, out, cyc);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
