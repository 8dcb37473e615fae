#include <cstdio>
#include <cstdlib>
__host__ __device__ int score_ref(const int *d) {
    int best = -255;
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; ++j) { int e = d[(k + j) & 15]; mn = e < mn ? e : mn; mx = e > mx ? e : mx; }
        int a = mn > -mx ? mn : -mx;
        best = a > best ? a : best;
    }
    return best;
}
// (a) doubling trick, mins of d and mins of -d
__device__ int score_a(const int *din) {
    int d[16], n[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { d[k] = din[k]; n[k] = -din[k]; }
    int best = -255;
    int a2[16], b2[16], a4[16], b4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) { a2[k] = min(d[k], d[(k + 1) & 15]); b2[k] = min(n[k], n[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; ++k) { a4[k] = min(a2[k], a2[(k + 2) & 15]); b4[k] = min(b2[k], b2[(k + 2) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int a9 = min(min(a4[k], a4[(k + 4) & 15]), d[(k + 8) & 15]);
        int b9 = min(min(b4[k], b4[(k + 4) & 15]), n[(k + 8) & 15]);
        best = max(best, max(a9, b9));
    }
    return best;
}
__device__ __forceinline__ int imin(int a, int b) { int r; asm("min.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ int imax(int a, int b) { int r; asm("max.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// (c) original structure with asm min/max
__device__ int score_c(const int *din) {
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = din[k];
    int best = -255;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int mn = d[k], mx = d[k];
#pragma unroll
        for (int j = 1; j < 9; ++j) { int e = d[(k + j) & 15]; mn = imin(mn, e); mx = imax(mx, e); }
        best = imax(best, imax(mn, -mx));
    }
    return best;
}
// (d) ternaries
__device__ int score_d(const int *din) {
    int d[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = din[k];
    return score_ref(d);
}
__global__ void k(const int *d, int *o, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    o[i * 4 + 0] = score_a(d + i * 16); o[i * 4 + 1] = score_c(d + i * 16); o[i * 4 + 2] = score_d(d + i * 16); o[i*4+3] = 0;
}
int main() {
    const int N = 4096;
    int *h = (int *)malloc(N * 64), *ho = (int *)malloc(N * 16);
    srand(1);
    for (int i = 0; i < N * 16; ++i) h[i] = rand() % 511 - 255;
    int *dd, *dout;
    cudaMalloc(&dd, N * 64); cudaMalloc(&dout, N * 16);
    cudaMemcpy(dd, h, N * 64, cudaMemcpyHostToDevice);
    k<<<N / 128, 128>>>(dd, dout, N);
    cudaMemcpy(ho, dout, N * 16, cudaMemcpyDeviceToHost);
    int bad[3] = {0, 0, 0};
    for (int i = 0; i < N; ++i) { int r = score_ref(h + i * 16); for (int v = 0; v < 3; ++v) bad[v] += ho[i * 4 + v] != r; }
    printf("mismatches a=%d c=%d d=%d of %d\n", bad[0], bad[1], bad[2], N);
}
