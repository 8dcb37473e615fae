#include <cstdio>
__host__ __device__ int score(const int *din) {
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = din[k];
    int best = -255;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
#pragma unroll
        for (int j = 1; j < 9; ++j) {
            int e = d[(k + j) & 15];
            mn = min(mn, e);
            mx = max(mx, e);
        }
        best = max(best, max(mn, -mx));
    }
    return best;
}
__global__ void k(const int *d, int *o) { *o = score(d); }
int main() {
    int d[16] = {30, 68, 72, 16, 4, 10, 10, 13, 9, 9, 10, 22, 34, 73, 64, 23};
    int *dd, *dout, out = -1;
    cudaMalloc(&dd, 64); cudaMalloc(&dout, 4);
    cudaMemcpy(dd, d, 64, cudaMemcpyHostToDevice);
    k<<<1, 1>>>(dd, dout);
    cudaMemcpy(&out, dout, 4, cudaMemcpyDeviceToHost);
    printf("host %d device %d\n", score(d), out);
}
