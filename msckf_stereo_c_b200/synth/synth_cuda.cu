// synth_cuda.cu — device build of the synthetic stereo generator (scene.h), in its own library
// (libmskf_synth_cuda.so): input generation for bench.py and the tests, not part of the engine.
#include <cuda_runtime.h>

#include "scene.h"

__global__ void synth_render_kernel(const SynthTraj *traj, const SynthCam *cams, const float *rays0, const float *rays1,
                                    const double *times, uint8_t *out, int n_streams) {
    // out layout: [stream][cam][rows*cols]
    const int s = blockIdx.z >> 1, cam = blockIdx.z & 1;
    const SynthCam *sc = &cams[cam];
    const int col = blockIdx.x * blockDim.x + threadIdx.x, row = blockIdx.y * blockDim.y + threadIdx.y;
    __shared__ float Rwc[9], o[3];
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        double R[9], p[3];
        synth_pose(&traj[s], times[s], R, p);
        synth_cam_pose(R, p, sc, Rwc, o);
    }
    __syncthreads();
    if (row >= sc->rows || col >= sc->cols) return;
    const float *rays = cam == 0 ? rays0 : rays1;
    out[((size_t)s * 2 + cam) * sc->rows * sc->cols + (size_t)row * sc->cols + col] =
        synth_pixel(&traj[s], sc, rays, Rwc, o, row, col);
}

extern "C" int mskf_synth_render_device(const void *d_traj, const void *d_cams, const float *d_rays0,
                                        const float *d_rays1, const double *d_times, uint8_t *d_out, int n_streams,
                                        int rows, int cols, void *cuda_stream) {
    dim3 b(32, 8), g((cols + 31) / 32, (rows + 7) / 8, n_streams * 2);
    synth_render_kernel<<<g, b, 0, (cudaStream_t)cuda_stream>>>((const SynthTraj *)d_traj, (const SynthCam *)d_cams, d_rays0,
                                                                d_rays1, d_times, d_out, n_streams);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
