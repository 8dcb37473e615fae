// frontend.cu — ImageProcessor hot path on sm_100a, batched over streams.
//
// Compiled with -fmad=false: every floating-point expression below is evaluated with the
// same IEEE operations, in the same order, as the CPU oracle built with -ffp-contract=off,
// and all pixel sums are integers, so pyramids, corners, KLT tracks and the published
// measurements are bit-identical to the oracle (SPEC.md).
//
// Reference call sites replaced (msckf_core/src/image_processor.cpp):
//   pyr_down_kernel      createImagePyramids :213-245  (cg::pyr_down :239,242)
//   klt_kernel           cg::optical_flow_multi_level :410 (temporal), :569 (stereo)
//   detect_kernel        CornerDetector::detect_features :259,657, set_grid_position :647
//   fe_prep_track        trackFeatures :362-389, predictFeatureTracking :321-350
//   fe_after_track       trackFeatures :415-440, stereoMatch :542-548
//   fe_after_stereo      stereoMatch :574-617, trackFeatures :465-513, addNewFeatures :632-649
//   fe_sieve             addNewFeatures :659-688 / initializeFirstFrame :261-268
//   fe_finish            addNewFeatures :690-750, initializeFirstFrame :270-316,
//                        pruneGridFeatures :758-768, publish :1137-1182, stereoCallback :192-200
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace mskf {

// ======================================================================================
// Pyramid: 5x5 Gaussian [1 4 6 4 1]^2 / 256, round half up, BORDER_REFLECT_101, decimate.
// One CTA produces a 64x16 output tile from a 132x36 input tile staged in shared memory;
// the level-1 launch also lands level 0 in the stream's pyramid (the copy the reference
// makes at image_processor.cpp:144-145,234-235).
// ======================================================================================
#define PD_TW 64
#define PD_TH 16
#define PD_IW (2 * PD_TW + 4)
#define PD_IH (2 * PD_TH + 4)

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

template <bool COPY_SRC>
__global__ void __launch_bounds__(256) pyr_down_kernel(FeConst fc, FeBuffers fb, int level) {
    const int s = blockIdx.z >> 1, cam = blockIdx.z & 1;
    const FeStep st = fb.step[s];
    if (!st.active) return;
    uint8_t *pyr = (cam == 0 ? fb.pyr[st.slot] : fb.pyr[2]) + (size_t)s * fc.pyr_bytes;
    const int irows = fc.lvl_rows[level - 1], icols = fc.lvl_cols[level - 1];
    const int orows = fc.lvl_rows[level], ocols = fc.lvl_cols[level];
    const uint8_t *src = COPY_SRC ? (cam == 0 ? fb.src0[s] : fb.src1[s]) : pyr + fc.lvl_off[level - 1];
    uint8_t *dst = pyr + fc.lvl_off[level];
    const int P = fc.pitch;  // row pitch of every level (the caller's level-0 image is tightly packed: the same value)

    __shared__ uint8_t tin[PD_IH][PD_IW + 4];
    __shared__ unsigned short hrow[PD_IH][PD_TW];

    const int ox0 = blockIdx.x * PD_TW, oy0 = blockIdx.y * PD_TH;
    const int ix0 = 2 * ox0 - 2, iy0 = 2 * oy0 - 2;
    for (int idx = threadIdx.x; idx < PD_IH * PD_IW; idx += 256) {
        int r = idx / PD_IW, c = idx - r * PD_IW;
        int gy = reflect101(iy0 + r, irows), gx = reflect101(ix0 + c, icols);
        uint8_t v = src[(size_t)gy * P + gx];
        tin[r][c] = v;
        if (COPY_SRC) {
            // interior of the tile = this CTA's share of the level-0 landing copy
            int yy = iy0 + r, xx = ix0 + c;
            if (r >= 2 && r < PD_IH - 2 && c >= 2 && c < PD_IW - 2 && yy < irows && xx < icols)
                pyr[(size_t)yy * P + xx] = v;
        }
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < PD_IH * PD_TW; idx += 256) {
        int r = idx / PD_TW, c = idx - r * PD_TW;
        const uint8_t *p = &tin[r][2 * c];
        hrow[r][c] = (unsigned short)(p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4]);
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < PD_TH * PD_TW; idx += 256) {
        int r = idx / PD_TW, c = idx - r * PD_TW;
        int oy = oy0 + r, ox = ox0 + c;
        if (oy < orows && ox < ocols) {
            int v = hrow[2 * r][c] + 4 * hrow[2 * r + 1][c] + 6 * hrow[2 * r + 2][c] + 4 * hrow[2 * r + 3][c] +
                    hrow[2 * r + 4][c];
            dst[(size_t)oy * P + ox] = (uint8_t)((v + 128) >> 8);
        }
    }
}

// Bandwidth-oriented version used whenever the level width is a multiple of VEC bytes and the
// row fits the strip buffer: one CTA makes a full-width strip of PS_ROWS output rows.
//   stage   36 input rows -> shared memory with VEC-byte loads (16 B at level 0, 4 B below); the
//           level-1 launch writes the same registers back as the level-0 landing copy
//   filter  a thread owns two adjacent output columns and walks down the strip with a 5-row
//           register window; the 5 taps run on two packed 16-bit lanes (255*16*16 < 2^16, so
//           neither pass can carry between lanes): ~25 instructions per output pixel
#define PS_ROWS 16
#define PS_IN (2 * PS_ROWS + 4)
#define PS_PAD 16  // pixel 0 of a staged row sits at byte 16 (16-byte aligned stores); halo at 14, 15

__device__ __forceinline__ unsigned hfilt2(const uint8_t *row, int q) {
    // two horizontal 5-tap sums centred on pixels 4q and 4q + 2, packed (lo, hi)
    const unsigned *w = (const unsigned *)(row + PS_PAD - 4) + q;
    const unsigned W0 = w[0], W1 = w[1], W2 = w[2];
    const unsigned V = __funnelshift_r(W0, W1, 16);  // p-2 p-1 p0 p1
    const unsigned X = __funnelshift_r(W1, W2, 16);  // p2 p3 p4 p5
    const unsigned A = V & 0x00FF00FFu, B = (V >> 8) & 0x00FF00FFu;
    const unsigned C = W1 & 0x00FF00FFu, D = (W1 >> 8) & 0x00FF00FFu;
    const unsigned E = X & 0x00FF00FFu;
    return A + E + 6u * C + 4u * (B + D);
}

template <bool COPY_SRC, int VEC>
__global__ void __launch_bounds__(256) pyr_down_strip_kernel(FeConst fc, FeBuffers fb, int level, int row_stride) {
    const int s = blockIdx.z >> 1, cam = blockIdx.z & 1;
    const FeStep st = fb.step[s];
    if (!st.active) return;
    uint8_t *pyr = (cam == 0 ? fb.pyr[st.slot] : fb.pyr[2]) + (size_t)s * fc.pyr_bytes;
    const int irows = fc.lvl_rows[level - 1], icols = fc.lvl_cols[level - 1];
    const int orows = fc.lvl_rows[level], ocols = fc.lvl_cols[level];
    const uint8_t *src = COPY_SRC ? (cam == 0 ? fb.src0[s] : fb.src1[s]) : pyr + fc.lvl_off[level - 1];
    uint8_t *dst = pyr + fc.lvl_off[level];
    const int P = fc.pitch;  // row pitch of every level (the caller's level-0 image is tightly packed: the same value)
    extern __shared__ __align__(16) uint8_t ps_smem[];  // [PS_IN][row_stride]
    const int oy0 = blockIdx.x * PS_ROWS;
    const int iy0 = 2 * oy0 - 2;
    typedef typename std::conditional<VEC == 16, uint4, unsigned>::type vec_t;
    const int vpr = icols / VEC;  // vectors per row
    for (int e = threadIdx.x; e < PS_IN * vpr; e += 256) {
        const int r = e / vpr, v = e - r * vpr;
        const int gy = reflect101(iy0 + r, irows);
        const vec_t val = *(const vec_t *)(src + (size_t)gy * P + (size_t)v * VEC);
        *(vec_t *)(ps_smem + (size_t)r * row_stride + PS_PAD + v * VEC) = val;
        if (COPY_SRC) {
            const int yy = iy0 + r;  // interior rows of the strip = this CTA's share of the level-0 landing copy
            if (r >= 2 && r < PS_IN - 2 && yy < irows) *(vec_t *)(pyr + (size_t)yy * P + (size_t)v * VEC) = val;
        }
    }
    __syncthreads();
    // BORDER_REFLECT_101 columns: -2, -1 and icols, icols + 1 (+2 zero bytes read by the last pair)
    for (int r = threadIdx.x; r < PS_IN; r += 256) {
        uint8_t *row = ps_smem + (size_t)r * row_stride + PS_PAD;
        row[-2] = row[reflect101(-2, icols)];
        row[-1] = row[reflect101(-1, icols)];
        row[icols] = row[reflect101(icols, icols)];
        row[icols + 1] = row[reflect101(icols + 1, icols)];
        row[icols + 2] = 0;
        row[icols + 3] = 0;
        row[icols + 4] = 0;
        row[icols + 5] = 0;
    }
    __syncthreads();
    const int ncp = (ocols + 1) >> 1;               // column pairs
    const int cp = min(256, (ncp + 31) & ~31);      // threads per row group
    const int groups = 256 / cp;                    // row groups working on disjoint output rows
    const int g = threadIdx.x / cp, t = threadIdx.x - g * cp;
    if (g >= groups) return;
    const int rpg = (PS_ROWS + groups - 1) / groups;
    const int r_begin = g * rpg, r_end = min(min(PS_ROWS, r_begin + rpg), orows - oy0);
    if (r_begin >= r_end) return;
    const bool odd_w = (ocols & 1) != 0;
    for (int q = t; q < ncp; q += cp) {
        const uint8_t *base = ps_smem + (size_t)(2 * r_begin) * row_stride;
        unsigned h0 = hfilt2(base, q), h1 = hfilt2(base + row_stride, q), h2 = hfilt2(base + 2 * row_stride, q);
        for (int r = r_begin; r < r_end; ++r) {
            const uint8_t *rp = ps_smem + (size_t)(2 * r + 3) * row_stride;
            const unsigned h3 = hfilt2(rp, q), h4 = hfilt2(rp + row_stride, q);
            const unsigned v = h0 + h4 + 6u * h2 + 4u * (h1 + h3);
            const unsigned o = ((v + 0x00800080u) >> 8) & 0x00FF00FFu;  // (sum + 128) >> 8 on both lanes
            uint8_t *d = dst + (size_t)(oy0 + r) * P + 2 * q;
            if (!odd_w) {
                *(unsigned short *)d = (unsigned short)((o & 0xFFu) | ((o >> 8) & 0xFF00u));
            } else {
                d[0] = (uint8_t)(o & 0xFFu);
                if (2 * q + 1 < ocols) d[1] = (uint8_t)(o >> 16);
            }
            h0 = h2;
            h1 = h3;
            h2 = h4;
        }
    }
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// Levels l0 .. L-1 of one image in ONE launch, one CTA per image (level 1 of a 752 x 480 frame is 90 KB: the
// rest of the pyramid fits in shared memory).  Level l0 - 1 comes from global memory in chunks of rows (with
// their two halo rows on either side), level l0 is filtered from the chunks into a resident shared-memory copy
// and into the pyramid, and every further level is filtered from shared memory into shared memory and the
// pyramid, alternating between the chunk area and the resident area.  Same arithmetic as the strip kernel
// (hfilt2, a 5-row register window per column pair walking down); rows are mirrored (BORDER_REFLECT_101) at
// access.  Replaces one launch per level >= 2 of pyr_down_strip_kernel<false, 4>, which at these sizes keeps
// 94 resp. 47 of 256 threads busy per 16-row strip (0.9 TB/s: 14 % of the HBM peak, 75 us for two levels).
#define PT_THREADS 256
__device__ __forceinline__ void pt_patch_borders(uint8_t *buf, int rows, int cols, int rs) {
    for (int r = threadIdx.x; r < rows; r += PT_THREADS) {
        uint8_t *row = buf + (size_t)r * rs + PS_PAD;
        row[-2] = row[reflect101(-2, cols)];
        row[-1] = row[reflect101(-1, cols)];
        row[cols] = row[reflect101(cols, cols)];
        row[cols + 1] = row[reflect101(cols + 1, cols)];
        row[cols + 2] = 0;
        row[cols + 3] = 0;
        row[cols + 4] = 0;
        row[cols + 5] = 0;
    }
}
// Output rows [o_begin, o_end) of a level (ocols wide) from the source rows held in `in`: source row y of the
// level sits at in + (y - in_y0) * rs_in for y in [in_y0, in_y0 + in_rows); rows outside [0, irows) are mirrored
// first.  Results go to the pyramid (dst, pitch P) and, when out != nullptr, to out + oy * rs_out + PS_PAD.
__device__ __forceinline__ void pt_filter_rows(const uint8_t *in, int in_y0, int rs_in, int irows, int o_begin, int o_end, int ocols,
                                               uint8_t *dst, int P, uint8_t *out, int rs_out) {
    const int ncp = (ocols + 1) >> 1;                      // column pairs
    const int cp = min(PT_THREADS, (ncp + 31) & ~31);      // threads per row group
    const int groups = PT_THREADS / cp;                    // row groups working on disjoint output rows
    const int g = threadIdx.x / cp, t = threadIdx.x - g * cp;
    if (g >= groups) return;
    const int n_rows = o_end - o_begin;
    const int rpg = (n_rows + groups - 1) / groups;
    const int r_begin = o_begin + g * rpg, r_end = min(o_end, r_begin + rpg);
    if (r_begin >= r_end) return;
    const bool odd_w = (ocols & 1) != 0;
    auto rowp = [&](int y) { return in + (size_t)(reflect101(y, irows) - in_y0) * rs_in; };
    for (int q = t; q < ncp; q += cp) {
        unsigned h0 = hfilt2(rowp(2 * r_begin - 2), q), h1 = hfilt2(rowp(2 * r_begin - 1), q), h2 = hfilt2(rowp(2 * r_begin), q);
        for (int r = r_begin; r < r_end; ++r) {
            const unsigned h3 = hfilt2(rowp(2 * r + 1), q), h4 = hfilt2(rowp(2 * r + 2), q);
            const unsigned v = h0 + h4 + 6u * h2 + 4u * (h1 + h3);
            const unsigned o = ((v + 0x00800080u) >> 8) & 0x00FF00FFu;  // (sum + 128) >> 8 on both lanes
            uint8_t *d = dst + (size_t)r * P + 2 * q;
            const uint8_t b0 = (uint8_t)(o & 0xFFu), b1 = (uint8_t)(o >> 16);
            if (!odd_w) {
                *(unsigned short *)d = (unsigned short)(b0 | (b1 << 8));
            } else {
                d[0] = b0;
                if (2 * q + 1 < ocols) d[1] = b1;
            }
            if (out) {
                uint8_t *so = out + (size_t)r * rs_out + PS_PAD + 2 * q;
                so[0] = b0;
                if (2 * q + 1 < ocols) so[1] = b1;
            }
            h0 = h2;
            h1 = h3;
            h2 = h4;
        }
    }
}
__device__ __forceinline__ int pt_stride(int cols) { return (PS_PAD + cols + 8 + 15) & ~15; }

__global__ void __launch_bounds__(PT_THREADS) pyr_tail_kernel(FeConst fc, FeBuffers fb, int l0, int chunk_out_rows, unsigned area_b_off) {
    const int s = blockIdx.x >> 1, cam = blockIdx.x & 1;
    const FeStep st = fb.step[s];
    if (!st.active) return;
    uint8_t *pyr = (cam == 0 ? fb.pyr[st.slot] : fb.pyr[2]) + (size_t)s * fc.pyr_bytes;
    const int P = fc.pitch;
    extern __shared__ __align__(16) uint8_t ps_smem[];
    __shared__ __align__(8) unsigned long long bar;
    uint8_t *area_a = ps_smem, *area_b = ps_smem + area_b_off;  // chunk of level l0 - 1 | resident level l0
    const unsigned bar_a = smem_u32(&bar);
    const bool bulk = (P % 16) == 0;
    unsigned phase = 0;
    if (bulk) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }
    // ---- level l0 from global memory, chunk by chunk
    {
        const int irows = fc.lvl_rows[l0 - 1], icols = fc.lvl_cols[l0 - 1];
        const int orows = fc.lvl_rows[l0], ocols = fc.lvl_cols[l0];
        const uint8_t *src = pyr + fc.lvl_off[l0 - 1];
        uint8_t *dst = pyr + fc.lvl_off[l0];
        const int rs_in = pt_stride(icols), rs_out = pt_stride(ocols);
        const bool resident = l0 + 1 < fc.levels;
        const int wpr = icols / 4;  // launch condition: icols % 4 == 0, pitch % 4 == 0
        for (int o_begin = 0; o_begin < orows; o_begin += chunk_out_rows) {
            const int o_end = min(orows, o_begin + chunk_out_rows);
            // source rows the chunk reads, after mirroring: [y_lo, y_hi]
            int y_lo = max(0, 2 * o_begin - 2), y_hi = min(irows - 1, 2 * (o_end - 1) + 2);
            if (2 * o_begin - 2 < 0) y_hi = max(y_hi, min(irows - 1, 2));        // rows -2, -1 mirror to 2, 1
            if (2 * (o_end - 1) + 2 > irows - 1) y_lo = min(y_lo, max(0, irows - 3));  // rows past the end mirror back
            const int n_in = y_hi - y_lo + 1;
            __syncthreads();  // the previous chunk has been consumed
            if (bulk) {
                // one bulk copy (TMA engine) per source row, issued by the lanes of warp 0, all signalling one mbarrier;
                // a row is copied in whole 16-byte units (the bytes past its end lie inside the pitch and are
                // overwritten by the border patch or never read)
                if (threadIdx.x < 32) {
                    // the border bytes of the previous chunk were written through the generic proxy
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    const unsigned row_bytes = (unsigned)((icols + 15) & ~15);
                    if (threadIdx.x == 0)
                        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(row_bytes * (unsigned)n_in) : "memory");
                    __syncwarp();
                    for (int r = threadIdx.x; r < n_in; r += 32)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                         smem_u32(area_a + (size_t)r * rs_in + PS_PAD)),
                                     "l"(src + (size_t)(y_lo + r) * P), "r"(row_bytes), "r"(bar_a)
                                     : "memory");
                }
                unsigned done = 0;
                while (!done) {
                    asm volatile(
                        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                        : "=r"(done)
                        : "r"(bar_a), "r"(phase)
                        : "memory");
                }
                phase ^= 1u;
            } else {
                for (int e = threadIdx.x; e < n_in * wpr; e += PT_THREADS) {
                    const int r = e / wpr, w = e - r * wpr;
                    *(unsigned *)(area_a + (size_t)r * rs_in + PS_PAD + 4 * w) = *(const unsigned *)(src + (size_t)(y_lo + r) * P + 4 * w);
                }
                __syncthreads();
            }
            pt_patch_borders(area_a, n_in, icols, rs_in);
            __syncthreads();
            pt_filter_rows(area_a, y_lo, rs_in, irows, o_begin, o_end, ocols, dst, P, resident ? area_b : nullptr, rs_out);
        }
    }
    // ---- further levels from shared memory
    uint8_t *in = area_b, *out = area_a;
    for (int l = l0 + 1; l < fc.levels; ++l) {
        const int irows = fc.lvl_rows[l - 1], icols = fc.lvl_cols[l - 1];
        const int orows = fc.lvl_rows[l], ocols = fc.lvl_cols[l];
        const int rs_in = pt_stride(icols), rs_out = pt_stride(ocols);
        __syncthreads();
        pt_patch_borders(in, irows, icols, rs_in);
        __syncthreads();
        pt_filter_rows(in, 0, rs_in, irows, 0, orows, ocols, pyr + fc.lvl_off[l], P, l + 1 < fc.levels ? out : nullptr, rs_out);
        uint8_t *t = in;
        in = out;
        out = t;
    }
}

// Level 0 -> 1 with the strip staged by the TMA engine: when a strip's 36 input rows are one
// contiguous span of the image (every strip except the top and bottom ones, whose reflected rows are
// fetched row by row) ONE cp.async.bulk moves 27 KB global -> shared and signals an mbarrier, and the
// level-0 landing copy is ONE cp.async.bulk shared -> global of the 32 interior rows: no thread touches
// the staging.  Rows are packed (stride = icols, no halo columns); the two edge column pairs patch
// their BORDER_REFLECT_101 bytes in registers.  Needs icols % 16 == 0.

__device__ __forceinline__ unsigned hfilt2_packed(const uint8_t *row, int q, bool first, bool last) {
    const unsigned *w = (const unsigned *)(row - 4) + q;
    unsigned W0 = w[0], W2 = w[2];
    const unsigned W1 = w[1];
    if (first) W0 = __byte_perm(W1, 0u, 0x1200);  // p[-2] = p[2], p[-1] = p[1]
    if (last) W2 = (W1 >> 16) & 0xFFu;             // p[icols] = p[icols - 2]
    const unsigned V = __funnelshift_r(W0, W1, 16);
    const unsigned X = __funnelshift_r(W1, W2, 16);
    const unsigned A = V & 0x00FF00FFu, B = (V >> 8) & 0x00FF00FFu;
    const unsigned C = W1 & 0x00FF00FFu, D = (W1 >> 8) & 0x00FF00FFu;
    const unsigned E = X & 0x00FF00FFu;
    return A + E + 6u * C + 4u * (B + D);
}

template <bool COPY_SRC>
__global__ void __launch_bounds__(256) pyr_down_bulk_kernel(FeConst fc, FeBuffers fb, int level) {
    const int s = blockIdx.z >> 1, cam = blockIdx.z & 1;
    const FeStep st = fb.step[s];
    if (!st.active) return;
    uint8_t *pyr = (cam == 0 ? fb.pyr[st.slot] : fb.pyr[2]) + (size_t)s * fc.pyr_bytes;
    const int irows = fc.lvl_rows[level - 1], icols = fc.lvl_cols[level - 1];
    const int orows = fc.lvl_rows[level], ocols = fc.lvl_cols[level];
    const uint8_t *src = COPY_SRC ? (cam == 0 ? fb.src0[s] : fb.src1[s]) : pyr + fc.lvl_off[level - 1];
    uint8_t *dst = pyr + fc.lvl_off[level];
    const int P = fc.pitch;  // row pitch of every level (the caller's level-0 image is tightly packed: the same value)
    extern __shared__ __align__(128) uint8_t ps_smem[];  // 16 B pad | PS_IN rows of icols bytes | 16 B pad
    __shared__ __align__(8) unsigned long long bar;
    uint8_t *tile = ps_smem + 16;
    const int oy0 = blockIdx.x * PS_ROWS;
    const int iy0 = 2 * oy0 - 2;
    const unsigned bar_a = smem_u32(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_a), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = (unsigned)(PS_IN * icols);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(total) : "memory");
        if (iy0 >= 0 && iy0 + PS_IN <= irows) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(tile)),
                         "l"(src + (size_t)iy0 * icols), "r"(total), "r"(bar_a)
                         : "memory");
        } else {
            for (int r = 0; r < PS_IN; ++r) {
                const int gy = reflect101(iy0 + r, irows);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 smem_u32(tile + (size_t)r * icols)),
                             "l"(src + (size_t)gy * icols), "r"((unsigned)icols), "r"(bar_a)
                             : "memory");
            }
        }
    }
    {
        // every thread waits for the bytes to land (phase 0 of the barrier)
        unsigned done = 0;
        while (!done) {
            asm volatile(
                "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(done)
                : "r"(bar_a), "r"(0)
                : "memory");
        }
    }
    if (COPY_SRC && threadIdx.x == 0) {
        // level-0 landing copy: the interior rows of the strip are image rows [2 oy0, 2 oy0 + 2 PS_ROWS)
        const int y_begin = 2 * oy0, n_rows = min(2 * PS_ROWS, irows - y_begin);
        if (n_rows > 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(pyr + (size_t)y_begin * icols),
                         "r"(smem_u32(tile + 2 * icols)), "r"((unsigned)(n_rows * icols))
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    const int ncp = ocols >> 1;  // icols % 16 == 0 => ocols even
    const int cp = min(256, (ncp + 31) & ~31);
    const int groups = 256 / cp;
    const int g = threadIdx.x / cp, t = threadIdx.x - g * cp;
    const int rpg = (PS_ROWS + groups - 1) / groups;
    const int r_begin = g * rpg, r_end = min(min(PS_ROWS, r_begin + rpg), orows - oy0);
    if (g < groups && r_begin < r_end) {
        for (int q = t; q < ncp; q += cp) {
            const bool first = q == 0, last = q == ncp - 1;
            const uint8_t *base = tile + (size_t)(2 * r_begin) * icols;
            unsigned h0 = hfilt2_packed(base, q, first, last), h1 = hfilt2_packed(base + icols, q, first, last),
                     h2 = hfilt2_packed(base + 2 * icols, q, first, last);
            for (int r = r_begin; r < r_end; ++r) {
                const uint8_t *rp = tile + (size_t)(2 * r + 3) * icols;
                const unsigned h3 = hfilt2_packed(rp, q, first, last), h4 = hfilt2_packed(rp + icols, q, first, last);
                const unsigned v = h0 + h4 + 6u * h2 + 4u * (h1 + h3);
                const unsigned o = ((v + 0x00800080u) >> 8) & 0x00FF00FFu;
                *(unsigned short *)(dst + (size_t)(oy0 + r) * P + 2 * q) = (unsigned short)((o & 0xFFu) | ((o >> 8) & 0xFF00u));
                h0 = h2;
                h1 = h3;
                h2 = h4;
            }
        }
    }
    // the bulk store reads shared memory asynchronously: it must have finished before the CTA retires
    if (COPY_SRC && threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ======================================================================================
// Pyramidal Lucas-Kanade, one warp per feature, all levels in one launch.
// Integer fixed point (14-bit bilinear weights, 5 fractional sample bits), warp-shuffle
// reductions of the 2x2 structure tensor and the mismatch vector, fp64 2x2 solve.
// ======================================================================================
struct BilinW { int ix, iy, w00, w01, w10, w11; };

__device__ __forceinline__ BilinW bilin_weights(float x, float y) {
    BilinW b;
    float fx = floorf(x), fy = floorf(y);
    b.ix = (int)fx; b.iy = (int)fy;
    float a = x - fx, c = y - fy;
    b.w00 = __float2int_rn((1.f - a) * (1.f - c) * 16384.f);
    b.w01 = __float2int_rn(a * (1.f - c) * 16384.f);
    b.w10 = __float2int_rn((1.f - a) * c * 16384.f);
    b.w11 = 16384 - b.w00 - b.w01 - b.w10;
    return b;
}
__device__ __forceinline__ long long warp_sum(long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

#define KLT_WARPS 4
#define KLT_REG_CTAS 6  // register version: <= 85 registers per thread -> 24 warps per SM (80 used, no spills)
// mode 0: temporal (A = previous cam0 pyramid, B = current cam0 pyramid)
// mode 1: stereo   (A = current cam0 pyramid,  B = current cam1 pyramid)
// mode 2: stereo of new candidates: entries flagged in k_skip are not matched (their cell has no
//         vacancy, so addNewFeatures (image_processor.cpp:735-750) never looks at them)
//
// Lane l owns image column x0 + l of the patch: per patch row it loads ONE byte (the row below),
// keeps the row above in a register and takes the right-hand neighbours from lane l + 1 by
// shuffle, so a bilinear sample costs one load and one shuffle instead of four clamped loads.
// Needs win + 3 <= 32 lanes.
template <int WIN>
__global__ void __launch_bounds__(KLT_WARPS * 32) klt_kernel(FeConst fc, FeBuffers fb, int mode) {
    const int s = blockIdx.y;
    const FeStep st = fb.step[s];
    if (!st.active) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * KLT_WARPS + warp;
    if (f >= fb.k_n[s]) return;
    if (mode == 2 && fb.k_skip[(size_t)s * fc.cap_k + f]) {
        if (lane == 0) fb.k_status[(size_t)s * fc.cap_k + f] = 0;
        return;
    }
    const uint8_t *pa = (mode == 0 ? fb.pyr[st.slot ^ 1] : fb.pyr[st.slot]) + (size_t)s * fc.pyr_bytes;
    const uint8_t *pb = (mode == 0 ? fb.pyr[st.slot] : fb.pyr[2]) + (size_t)s * fc.pyr_bytes;

    const int win = WIN ? WIN : fc.klt_win, half = win >> 1, tw = win + 2;
    extern __shared__ short klt_smem[];
    short *T = klt_smem + (size_t)warp * ((tw * tw + 2 * win * win + 3) & ~3);
    short2 *Gr = (short2 *)(T + tw * tw + ((tw * tw) & 1));  // (Ix, Iy) per window pixel, 4-byte aligned

    const float2 p0 = fb.k_a[(size_t)s * fc.cap_k + f];
    float2 q0 = fb.k_b[(size_t)s * fc.cap_k + f];
    int status = 1;
    const int L = fc.levels;
    const float top_scale = 1.0f / (float)(1 << (L - 1));
    float qx = q0.x * top_scale, qy = q0.y * top_scale;
    for (int l = L - 1; l >= 0; --l) {
        const int rows = fc.lvl_rows[l], cols = fc.lvl_cols[l];
        const uint8_t *A = pa + fc.lvl_off[l], *B = pb + fc.lvl_off[l];
        const float sc = 1.0f / (float)(1 << l);
        const float px = p0.x * sc, py = p0.y * sc;
        const BilinW wa = bilin_weights(px, py);
        __syncwarp();
        {
            // template patch T[j][i] = A sampled at (px + i - half - 1, py + j - half - 1), i, j in [0, tw)
            const int xc = min(max(wa.ix - half - 1 + lane, 0), cols - 1);
            const int y0 = wa.iy - half - 1;
            if (WIN) {
                int px[(WIN ? WIN : 1) + 3], nx[(WIN ? WIN : 1) + 3];
#pragma unroll
                for (int j = 0; j <= WIN + 2; ++j) px[j] = __ldg(A + (size_t)min(max(y0 + j, 0), rows - 1) * fc.pitch + xc);
#pragma unroll
                for (int j = 0; j <= WIN + 2; ++j) nx[j] = __shfl_down_sync(0xffffffffu, px[j], 1);
                if (lane < tw) {
#pragma unroll
                    for (int j = 0; j < WIN + 2; ++j)
                        T[j * tw + lane] = (short)((wa.w00 * px[j] + wa.w01 * nx[j] + wa.w10 * px[j + 1] + wa.w11 * nx[j + 1] + 256) >> 9);
                }
            } else {
                int top = __ldg(A + (size_t)min(max(y0, 0), rows - 1) * fc.pitch + xc);
                int rt = __shfl_down_sync(0xffffffffu, top, 1);
                for (int j = 0; j < tw; ++j) {
                    const int bot = __ldg(A + (size_t)min(max(y0 + j + 1, 0), rows - 1) * fc.pitch + xc);
                    const int rb = __shfl_down_sync(0xffffffffu, bot, 1);
                    if (lane < tw) T[j * tw + lane] = (short)((wa.w00 * top + wa.w01 * rt + wa.w10 * bot + wa.w11 * rb + 256) >> 9);
                    top = bot;
                    rt = rb;
                }
            }
        }
        __syncwarp();
        // A lane's partial sums fit 32 bits: |gx|, |gy|, |diff| <= 255 * 32 = 8160, so a column of win <= 29
        // products stays below 29 * 8160^2 = 1.93e9 < 2^31 (one IMAD per term instead of a multiply and a
        // 64-bit add); only the warp reduction needs 64 bits.
        int a11i = 0, a12i = 0, a22i = 0;
        if (lane < win) {
#pragma unroll
            for (int j = 0; j < win; ++j) {
                const int gx = (int)T[(j + 1) * tw + lane + 2] - (int)T[(j + 1) * tw + lane];
                const int gy = (int)T[(j + 2) * tw + lane + 1] - (int)T[j * tw + lane + 1];
                Gr[j * win + lane] = make_short2((short)gx, (short)gy);
                a11i += gx * gx;
                a12i += gx * gy;
                a22i += gy * gy;
            }
        }
        long long A11 = warp_sum((long long)a11i), A12 = warp_sum((long long)a12i), A22 = warp_sum((long long)a22i);
        __syncwarp();
        const double a11 = (double)A11, a12 = (double)A12, a22 = (double)A22;
        const double m1 = a11 * a22, m2 = a12 * a12;
        const double D = m1 - m2;
        const double df = a11 - a22;
        const double disc = df * df + 4.0 * m2;
        const double lam = (a11 + a22 - sqrt(disc)) * 0.5;
        const double min_eig = lam / (4194304.0 * (double)(win * win));
        const bool ok = !(min_eig < fc.klt_min_eig || D < 1.1920929e-07);
        if (!ok) {
            if (l == 0) status = 0;
        } else {
            const double Dinv = 1.0 / D;
            double pdx = 0, pdy = 0;
            for (int it = 0; it < fc.klt_max_iters; ++it) {
                if (qx < 0.f || qy < 0.f || qx > (float)(cols - 1) || qy > (float)(rows - 1)) {
                    if (l == 0) status = 0;
                    break;
                }
                const BilinW wb = bilin_weights(qx, qy);
                int b1i = 0, b2i = 0;
                {
                    const int xc = min(max(wb.ix - half + lane, 0), cols - 1);
                    const int y0 = wb.iy - half;
                    const short *Trow = T + tw + lane + 1;
                    const short2 *Grow = Gr + lane;
                    if (WIN) {
                        // all rows of my column first (independent loads in flight together), then the
                        // neighbours by shuffle, then the arithmetic
                        int px[(WIN ? WIN : 1) + 1], nx[(WIN ? WIN : 1) + 1];
#pragma unroll
                        for (int j = 0; j <= WIN; ++j) px[j] = __ldg(B + (size_t)min(max(y0 + j, 0), rows - 1) * fc.pitch + xc);
#pragma unroll
                        for (int j = 0; j <= WIN; ++j) nx[j] = __shfl_down_sync(0xffffffffu, px[j], 1);
                        if (lane < win) {
#pragma unroll
                            for (int j = 0; j < WIN; ++j) {
                                const int val = (wb.w00 * px[j] + wb.w01 * nx[j] + wb.w10 * px[j + 1] + wb.w11 * nx[j + 1] + 256) >> 9;
                                const int diff = val - (int)Trow[j * tw];
                                const short2 g = Grow[j * win];
                                b1i += diff * (int)g.x;
                                b2i += diff * (int)g.y;
                            }
                        }
                    } else {
                        int top = __ldg(B + (size_t)min(max(y0, 0), rows - 1) * fc.pitch + xc);
                        int rt = __shfl_down_sync(0xffffffffu, top, 1);
                        for (int j = 0; j < win; ++j) {
                            const int bot = __ldg(B + (size_t)min(max(y0 + j + 1, 0), rows - 1) * fc.pitch + xc);
                            const int rb = __shfl_down_sync(0xffffffffu, bot, 1);
                            if (lane < win) {
                                const int val = (wb.w00 * top + wb.w01 * rt + wb.w10 * bot + wb.w11 * rb + 256) >> 9;
                                const int diff = val - (int)Trow[j * tw];
                                const short2 g = Grow[j * win];
                                b1i += diff * (int)g.x;
                                b2i += diff * (int)g.y;
                            }
                            top = bot;
                            rt = rb;
                        }
                    }
                }
                const long long b1 = warp_sum((long long)b1i), b2 = warp_sum((long long)b2i);
                const double fb1 = (double)b1, fb2 = (double)b2;
                const double dx = (a12 * fb2 - a22 * fb1) * Dinv * 2.0;
                const double dy = (a12 * fb1 - a11 * fb2) * Dinv * 2.0;
                const float fdx = (float)dx, fdy = (float)dy;
                qx += fdx;
                qy += fdy;
                if (dx * dx + dy * dy <= fc.klt_eps2) break;
                if (it > 0 && fabs(dx + pdx) < 0.01 && fabs(dy + pdy) < 0.01) {
                    qx -= fdx * 0.5f;
                    qy -= fdy * 0.5f;
                    break;
                }
                pdx = dx;
                pdy = dy;
            }
        }
        if (l > 0) { qx *= 2.0f; qy *= 2.0f; }
    }
    if (lane == 0) {
        fb.k_b[(size_t)s * fc.cap_k + f] = make_float2(qx, qy);
        fb.k_status[(size_t)s * fc.cap_k + f] = (uint8_t)status;
    }
}

// --------------------------------------------------------------------------------------
// Register version for compile-time windows (15, 21): no shared memory at all.
//   * Lane l owns column l of the (WIN + 2)-wide template patch: it loads the WIN + 3 raw bytes of its
//     column (independent loads, all in flight together), takes the right-hand neighbours by shuffle and
//     keeps its column of T in registers.  The gradients of window column i = l - 1 need T of lanes
//     l - 1 / l + 1 (two shuffles per row) and the lane's own column for the vertical difference; they stay
//     in registers as one packed (gx, gy) word per row for all iterations of the level.
//   * The mismatch vector is accumulated as sum(val * g) - sum(T * g): the second term is a per-level
//     constant, so an iteration touches neither T nor shared memory (integers: the sums are exact, the
//     result is identical to the oracle's sum((val - T) * g)).
//   * Every warp sum is two REDUX.SUM (the 32-bit partial split into its low 16 bits and the rest: both
//     halves of the 64-bit total fit 32 bits) instead of a 5-step 64-bit shuffle tree.
// --------------------------------------------------------------------------------------
// base + y * cols as ONE 32 x 32 + 64-bit multiply-add (the compiler otherwise folds the column offset into the
// product and adds the 64-bit level pointer separately: three instructions per row instead of one)
__device__ __forceinline__ const uint8_t *row_ptr(const uint8_t *base, int y, int cols) {
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"((unsigned)y), "r"((unsigned)cols), "l"((unsigned long long)base));
    return (const uint8_t *)r;
}
__device__ __forceinline__ long long warp_sum_i32(int v) {
    const int lo = v & 0xffff, hi = v >> 16;  // v = hi * 65536 + lo, 0 <= lo < 65536
    return ((long long)__reduce_add_sync(0xffffffffu, hi) << 16) + (long long)__reduce_add_sync(0xffffffffu, lo);
}

//   * PITCH > 0: the row pitch of the pyramid is a compile-time constant (every level keeps the pitch of level
//     0), so the rows of a patch are IMMEDIATE offsets of one 64-bit base address: one LDG per row instead of a
//     multiply, a 64-bit add and the LDG (3 of the 12 instructions a window row costs per iteration).
template <int WIN, int PITCH>
__global__ void __launch_bounds__(KLT_WARPS * 32, KLT_REG_CTAS) klt_reg_kernel(FeConst fc, FeBuffers fb, int mode) {
    static_assert(WIN >= 3 && WIN + 3 <= 32 && (WIN & 1), "window");
    const int pitch = PITCH ? PITCH : fc.pitch;
    // packed gradients of the level: [row][lane] per warp (a lane only ever reads its own words: no bank
    // conflicts, no synchronisation); in registers they cost 2 x WIN live values across the iteration loop
    // and capped the kernel at 16 warps per SM
    __shared__ int2 s_G[KLT_WARPS][WIN][32];
    const int s = blockIdx.y;
    const FeStep st = fb.step[s];
    if (!st.active) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // modes 0 / 1: the grid covers the list, one feature per warp.  Mode 2 (stereo match of the new candidates): a
    // small grid strides over the candidates fe_sieve compacted (the ones whose cell has a vacancy: ~15 % of the
    // list; a warp per list entry spent half of the launch on CTAs that had nothing to match)
    const int n_items = mode == 2 ? fb.k_nm[s] : fb.k_n[s];
    for (int item = blockIdx.x * KLT_WARPS + warp; item < n_items; item += gridDim.x * KLT_WARPS) {
    const int f = mode == 2 ? fb.k_idx[(size_t)s * fc.cap_k + item] : item;
    __syncwarp();
    const uint8_t *pa = (mode == 0 ? fb.pyr[st.slot ^ 1] : fb.pyr[st.slot]) + (size_t)s * fc.pyr_bytes;
    const uint8_t *pb = (mode == 0 ? fb.pyr[st.slot] : fb.pyr[2]) + (size_t)s * fc.pyr_bytes;
    constexpr int half = WIN >> 1, tw = WIN + 2;
    const bool act = lane >= 1 && lane <= WIN;  // lanes that own a window column (i = lane - 1)

    const float2 p0 = fb.k_a[(size_t)s * fc.cap_k + f];
    const float2 q0 = fb.k_b[(size_t)s * fc.cap_k + f];
    int status = 1;
    const int L = fc.levels;
    const float top_scale = 1.0f / (float)(1 << (L - 1));
    float qx = q0.x * top_scale, qy = q0.y * top_scale;
    for (int l = L - 1; l >= 0; --l) {
        const int rows = fc.lvl_rows[l], cols = fc.lvl_cols[l];
        const uint8_t *A = pa + fc.lvl_off[l], *B = pb + fc.lvl_off[l];
        const float sc = 1.0f / (float)(1 << l);
        const float px = p0.x * sc, py = p0.y * sc;
        const BilinW wa = bilin_weights(px, py);
        int2 *G = &s_G[warp][0][lane];  // G[32 j] = (gx, gy) of window column lane - 1, row j: one LDS.64, no unpacking
        int a11i = 0, a12i = 0, a22i = 0, c1i = 0, c2i = 0;
        {
            // template column T[j], j in [0, tw): A sampled at (px + lane - half - 1, py + j - half - 1)
            const int xc = min(max(wa.ix - half - 1 + lane, 0), cols - 1);
            const int y0 = wa.iy - half - 1;
            const uint8_t *Ax = A + xc;  // one 32 x 32 + 64-bit multiply-add per row address
            int raw[WIN + 3];
#pragma unroll
            if (y0 >= 0 && y0 + WIN + 2 <= rows - 1) {  // warp-uniform: no row of the patch needs clamping
                const uint8_t *Ay = row_ptr(Ax, y0, pitch);
#pragma unroll
                for (int j = 0; j <= WIN + 2; ++j) raw[j] = __ldg(PITCH ? Ay + j * PITCH : row_ptr(Ay, j, pitch));
            } else {
#pragma unroll
                for (int j = 0; j <= WIN + 2; ++j) raw[j] = __ldg(row_ptr(Ax, min(max(y0 + j, 0), rows - 1), pitch));
            }
            int T[WIN + 2];
            int n0 = __shfl_down_sync(0xffffffffu, raw[0], 1);
#pragma unroll
            for (int j = 0; j < WIN + 2; ++j) {
                const int n1 = __shfl_down_sync(0xffffffffu, raw[j + 1], 1);
                T[j] = (((wa.w00 * raw[j] + 256) + wa.w01 * n0) + wa.w10 * raw[j + 1] + wa.w11 * n1) >> 9;
                n0 = n1;
            }
#pragma unroll
            for (int j = 0; j < WIN; ++j) {
                const int tl = __shfl_up_sync(0xffffffffu, T[j + 1], 1), tr = __shfl_down_sync(0xffffffffu, T[j + 1], 1);
                const int gx = tr - tl, gy = T[j + 2] - T[j];  // garbage on the lanes without a window column: masked below
                G[32 * j] = make_int2(gx, gy);
                a11i += gx * gx;
                a12i += gx * gy;
                a22i += gy * gy;
                c1i += T[j + 1] * gx;
                c2i += T[j + 1] * gy;
            }
        }
        // A lane's partial sums fit 32 bits: |gx|, |gy|, T, val <= 255 * 32 = 8160, so a column of WIN <= 29
        // products stays below 29 * 8160^2 = 1.93e9 < 2^31; the warp totals need 64 bits.
        if (!act) a11i = a12i = a22i = c1i = c2i = 0;
        const long long A11 = warp_sum_i32(a11i), A12 = warp_sum_i32(a12i), A22 = warp_sum_i32(a22i);
        const double a11 = (double)A11, a12 = (double)A12, a22 = (double)A22;
        const double m1 = a11 * a22, m2 = a12 * a12;
        const double D = m1 - m2;
        const double df = a11 - a22;
        const double disc = df * df + 4.0 * m2;
        const double lam = (a11 + a22 - sqrt(disc)) * 0.5;
        const double min_eig = lam / (4194304.0 * (double)(WIN * WIN));
        const bool ok = !(min_eig < fc.klt_min_eig || D < 1.1920929e-07);
        if (!ok) {
            if (l == 0) status = 0;
        } else {
            const long long C1 = warp_sum_i32(c1i), C2 = warp_sum_i32(c2i);
            const double Dinv = 1.0 / D;
            double pdx = 0, pdy = 0;
            for (int it = 0; it < fc.klt_max_iters; ++it) {
                if (qx < 0.f || qy < 0.f || qx > (float)(cols - 1) || qy > (float)(rows - 1)) {
                    if (l == 0) status = 0;
                    break;
                }
                const BilinW wb = bilin_weights(qx, qy);
                int b1i = 0, b2i = 0;
                {
                    const int xc = min(max(wb.ix - half - 1 + lane, 0), cols - 1);
                    const int y0 = wb.iy - half;
                    const uint8_t *Bx = B + xc;
                    int pxv[WIN + 1];
#pragma unroll
                    if (y0 >= 0 && y0 + WIN <= rows - 1) {
                        const uint8_t *By = row_ptr(Bx, y0, pitch);
#pragma unroll
                        for (int j = 0; j <= WIN; ++j) pxv[j] = __ldg(PITCH ? By + j * PITCH : row_ptr(By, j, pitch));
                    } else {
#pragma unroll
                        for (int j = 0; j <= WIN; ++j) pxv[j] = __ldg(row_ptr(Bx, min(max(y0 + j, 0), rows - 1), pitch));
                    }
                    int n0 = __shfl_down_sync(0xffffffffu, pxv[0], 1);
#pragma unroll
                    for (int j = 0; j < WIN; ++j) {
                        const int n1 = __shfl_down_sync(0xffffffffu, pxv[j + 1], 1);
                        const int val = (((wb.w00 * pxv[j] + 256) + wb.w01 * n0) + wb.w10 * pxv[j + 1] + wb.w11 * n1) >> 9;
                        const int2 gj = G[32 * j];
                        b1i += val * gj.x;
                        b2i += val * gj.y;
                        n0 = n1;
                    }
                }
                if (!act) b1i = b2i = 0;
                const long long b1 = warp_sum_i32(b1i) - C1, b2 = warp_sum_i32(b2i) - C2;
                const double fb1 = (double)b1, fb2 = (double)b2;
                const double dx = (a12 * fb2 - a22 * fb1) * Dinv * 2.0;
                const double dy = (a12 * fb1 - a11 * fb2) * Dinv * 2.0;
                const float fdx = (float)dx, fdy = (float)dy;
                qx += fdx;
                qy += fdy;
                if (dx * dx + dy * dy <= fc.klt_eps2) break;
                if (it > 0 && fabs(dx + pdx) < 0.01 && fabs(dy + pdy) < 0.01) {
                    qx -= fdx * 0.5f;
                    qy -= fdy * 0.5f;
                    break;
                }
                pdx = dx;
                pdy = dy;
            }
        }
        if (l > 0) { qx *= 2.0f; qy *= 2.0f; }
    }
    if (lane == 0) {
        fb.k_b[(size_t)s * fc.cap_k + f] = make_float2(qx, qy);
        fb.k_status[(size_t)s * fc.cap_k + f] = (uint8_t)status;
    }
    }  // item loop
}

// ======================================================================================
// FAST-9/16 + score + 3x3 NMS + Shi-Tomasi response + best-per-fine-cell (atomicMax).
// Shared-memory tile with a 5-pixel halo; the 16-bit brighter/darker ring masks decide
// cornerness, the arc score is the sliding minimum over 9 contiguous ring differences.
// ======================================================================================
#define DT_W 64
#define DT_H 16
#define DT_HALO 5
__constant__ int c_fdx[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
__constant__ int c_fdy[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

__device__ __forceinline__ bool has_arc9(unsigned m) {  // 9 contiguous set bits in a 16-bit ring
    m |= m << 16;
    unsigned a = m & (m >> 1);
    a &= a >> 2;
    a &= a >> 4;          // 8 contiguous
    a &= m >> 8;          // 9 contiguous
    return (a & 0xffffu) != 0;
}

// Stages of one 64x16 tile (ncu source page of the first version: tile load 18 %, ring test 22 %, arc score
// 22 % at 8 of 32 lanes, NMS 11 %, Shi-Tomasi 24 % of the issue slots):
//  1. the tile with its halo comes in as aligned 32-bit words (the tile starts 8 columns left of the
//     block, so image words map to tile words; images whose width is not a multiple of 4 take bytes);
//  2. ring test for every pixel of the (64+2)x(16+2) score window; pixels that pass are COMPACTED into a
//     list so that
//  3. the arc score runs on full warps (on a corner-dense texture a quarter of the pixels pass, which
//     kept every warp in the score code at a quarter of its lanes);
//  4. strict 3x3 NMS, maxima compacted;
//  5. Shi-Tomasi response, 8 lanes per maximum (one row of the 8x8 box each, integer sums: order-free).
#define DT_X0 8
#define DT_STRIDE 84  // bytes per tile row: 64 + 8 left + 8 right, + 4 so that rows start in different banks
#define DT_SW (DT_W + 2)
#define DT_SH (DT_H + 2)
__global__ void __launch_bounds__(256, 6) detect_kernel(FeConst fc, FeBuffers fb) {
    const int s = blockIdx.z;
    const FeStep st = fb.step[s];
    if (!st.active) return;
    const uint8_t *img = fb.pyr[st.slot] + (size_t)s * fc.pyr_bytes;  // level 0 of current cam0
    const int rows = fc.rows, cols = fc.cols;
    __shared__ __align__(16) uint8_t tile[DT_H + 2 * DT_HALO][DT_STRIDE];
    __shared__ __align__(16) uint8_t score[DT_SH][DT_SW + 2];
    __shared__ unsigned short s_list[DT_SH * DT_SW];  // pixels that passed the ring test: row << 7 | column | side << 15
    __shared__ int s_nlist, s_nmax;
    __shared__ unsigned short s_max[DT_H * DT_W / 4];  // strict 3x3 maxima: at most one per 2x2 block
    // fine-cell column / row of every pixel column / row of the tile (two runtime divisions per candidate otherwise:
    // 7 % of the kernel's instructions at 4 of 32 lanes, ncu source page)
    __shared__ unsigned short s_cx[DT_W], s_cy[DT_H];
    const int x0 = blockIdx.x * DT_W, y0 = blockIdx.y * DT_H;
    if (threadIdx.x == 0) {
        s_nlist = 0;
        s_nmax = 0;
    }
    if (threadIdx.x < DT_W) s_cx[threadIdx.x] = (unsigned short)((x0 + threadIdx.x) / fc.det_cell_w);
    else if (threadIdx.x < DT_W + DT_H) s_cy[threadIdx.x - DT_W] = (unsigned short)((y0 + threadIdx.x - DT_W) / fc.det_cell_h);
    if ((cols & 3) == 0 && (((size_t)img) & 3) == 0) {
        constexpr int WPR = (DT_W + 2 * DT_X0) / 4;  // 20 words per tile row
        for (int idx = threadIdx.x; idx < (DT_H + 2 * DT_HALO) * WPR; idx += 256) {
            const int r = idx / WPR, w = idx - r * WPR;
            const int gy = y0 - DT_HALO + r, gx = x0 - DT_X0 + 4 * w;
            unsigned v = 0;
            if (gy >= 0 && gy < rows && gx >= 0 && gx < cols) v = *reinterpret_cast<const unsigned *>(img + (size_t)gy * cols + gx);
            *reinterpret_cast<unsigned *>(&tile[r][4 * w]) = v;
        }
    } else {
        for (int idx = threadIdx.x; idx < (DT_H + 2 * DT_HALO) * (DT_W + 2 * DT_X0); idx += 256) {
            const int r = idx / (DT_W + 2 * DT_X0), c = idx - r * (DT_W + 2 * DT_X0);
            const int gy = y0 - DT_HALO + r, gx = x0 - DT_X0 + c;
            uint8_t v = 0;
            if (gy >= 0 && gy < rows && gx >= 0 && gx < cols) v = img[(size_t)gy * cols + gx];
            tile[r][c] = v;
        }
    }
    __syncthreads();
    const int t = fc.fast_threshold;
    const int lane = threadIdx.x & 31;
    // Ring offsets as immediates: one LDS per ring pixel.  Bresenham circle of radius 3, k = 0 at (0, +3).
#define RING_PX(pc, dx, dy) ((int)(pc)[(dy) * DT_STRIDE + (dx)])
#define RING_LIST(X) X(0, 0, 3) X(1, 1, 3) X(2, 2, 2) X(3, 3, 1) X(4, 3, 0) X(5, 3, -1) X(6, 2, -2) X(7, 1, -3) \
    X(8, 0, -3) X(9, -1, -3) X(10, -2, -2) X(11, -3, -1) X(12, -3, 0) X(13, -3, 1) X(14, -2, 2) X(15, -1, 3)
    // The score window is cleared with word stores, the ring test runs over it in DT_RT_ITERS passes of 256
    // pixels, and the pixels that pass are compacted WITHOUT atomics: every pass keeps its warp ballot in a
    // register, the eight warp totals are prefix-summed once, and each warp then writes its entries (list
    // order is irrelevant: the score, the NMS and the per-cell arg-max below are order-free).  The first
    // version did one shared-memory atomicAdd per warp and pass, which the compiler wrapped in a second
    // warp-aggregation sequence: 40 of the 164 instructions per pixel (ncu source page).
    constexpr int DT_RT_ITERS = (DT_SH * DT_SW + 255) / 256;
    for (int w = threadIdx.x; w < (int)(sizeof(score) / 4); w += 256) reinterpret_cast<unsigned *>(&score[0][0])[w] = 0u;
    const bool tile_inside = y0 - 1 >= 3 && y0 + DT_H < rows - 3 && x0 - 1 >= 3 && x0 + DT_W < cols - 3;  // no pixel of the window needs the border test
    unsigned bal_it[DT_RT_ITERS];
    int pass_it[DT_RT_ITERS], code_it[DT_RT_ITERS];  // list entry: window row << 7 | window column | side << 15
    static_assert(DT_W == 64 && DT_H == 16 && DT_RT_ITERS == 5, "window-to-thread mapping below");
#pragma unroll
    for (int it = 0; it < DT_RT_ITERS; ++it) {
        // Window pixel (r, c) of this thread in pass `it`, without a division: the 64 inner columns go four rows
        // per pass (18 rows: four and a half passes), the two edge columns fill the idle half of the last pass.
        int r = it * 4 + (threadIdx.x >> 6), c = 1 + (threadIdx.x & 63);
        bool valid = true;
        if (it == DT_RT_ITERS - 1 && threadIdx.x >= 128) {
            const int e = threadIdx.x - 128;
            r = e >> 1;
            c = (e & 1) * (DT_SW - 1);
            valid = e < 2 * DT_SH;
        }
        int pass = 0;  // 1: darker arc, 2: brighter arc
        if (valid) {
            const int gy = y0 - 1 + r, gx = x0 - 1 + c;
            if (tile_inside || (gy >= 3 && gy < rows - 3 && gx >= 3 && gx < cols - 3)) {
                const uint8_t *pc = &tile[r + DT_HALO - 1][c + DT_X0 - 1];
                const int v = pc[0];
                const int lo = v - t, hi = v + t;  // darker ring pixel: p < lo (d = v - p > t); brighter: p > hi
                // Both tests of a ring pixel in one multiply-add: y = (p - lo) + 65536 (hi - p) has the sign of
                // p - lo in bit 15 and the sign of hi - p in bit 31 (a borrow out of the low half only happens
                // when p < lo, where hi - p > 2t >= 1 keeps the high half non-negative).  Bit 15 moves to bit k,
                // bit 31 to bit 16 + k: the darker mask ends up in the low half of z, the brighter one in the high.
                const unsigned cst = ((unsigned)hi << 16) - (unsigned)lo;
                unsigned z = 0;
#define RING_TEST(k, dx, dy)                                                                   \
    {                                                                                          \
        const unsigned y = (unsigned)RING_PX(pc, dx, dy) * 0xffff0001u + cst;                  \
        z |= (y >> (15 - k)) & (0x00010001u << k);                                             \
    }
                RING_LIST(RING_TEST)
#undef RING_TEST
                // an arc of 9 needs 9 set bits, and only one side can have them: one arc test instead of two
                const unsigned dark = z & 0xffffu, bright = z >> 16;
                const bool dside = __popc(dark) >= 9;
                if (has_arc9(dside ? dark : bright)) pass = dside ? 1 : 2;
            }
        }
        bal_it[it] = __ballot_sync(0xffffffffu, pass != 0);
        pass_it[it] = pass;
        code_it[it] = (r << 7) | c;
    }
    {
        __shared__ int s_wcnt[8];
        int wtot = 0;
#pragma unroll
        for (int it = 0; it < DT_RT_ITERS; ++it) wtot += __popc(bal_it[it]);
        if (lane == 0) s_wcnt[threadIdx.x >> 5] = wtot;
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int cnt = s_wcnt[w];
            if (w < (int)(threadIdx.x >> 5)) wbase += cnt;
            total += cnt;
        }
        if (threadIdx.x == 0) s_nlist = total;
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int it = 0; it < DT_RT_ITERS; ++it) {
            if (pass_it[it]) s_list[wbase + __popc(bal_it[it] & lt)] = (unsigned short)(code_it[it] | ((pass_it[it] - 1) << 15));
            wbase += __popc(bal_it[it]);
        }
    }
    __syncthreads();
    // S = max over the 16 arcs of 9 contiguous ring pixels of min(d) (darker ring) or min(-d) (brighter ring),
    // d = centre - ring.  A ring cannot hold a darker and a brighter 9-arc at once, and the side without an arc
    // scores <= t, so only the side that passed is evaluated.  Sliding minimum of width 9 by doubling
    // (2, 4, 8, +1); only `min` chains on purpose: ptxas 12.9 for sm_100a miscompiles interleaved min/max
    // chains fused into VIMNMX3 (tools/scratch/t2.cu).
    const int nlist = s_nlist;
    for (int e = threadIdx.x; e < nlist; e += 256) {
        const int code = s_list[e];
        const int sg = (code >> 15) ? -1 : 1;
        const int r = (code >> 7) & 31, c = code & 127;
        const uint8_t *pc = &tile[r + DT_HALO - 1][c + DT_X0 - 1];
        const int v = pc[0];
        int n[16], a2[16], a4[16];
#define RING_DIFF(k, dx, dy) n[k] = (v - RING_PX(pc, dx, dy)) * sg;
        RING_LIST(RING_DIFF)
#undef RING_DIFF
#pragma unroll
        for (int k = 0; k < 16; ++k) a2[k] = min(n[k], n[(k + 1) & 15]);
#pragma unroll
        for (int k = 0; k < 16; ++k) a4[k] = min(a2[k], a2[(k + 2) & 15]);
        int best = -255;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            int a9 = min(min(a4[k], a4[(k + 4) & 15]), n[(k + 8) & 15]);
            best = max(best, a9);
        }
        score[r][c] = (uint8_t)(best - 1);
    }
#undef RING_LIST
#undef RING_PX
    __syncthreads();
    if (fb.dbg_score && s == 0) {
        for (int idx = threadIdx.x; idx < DT_H * DT_W; idx += 256) {
            const int r = idx / DT_W, c = idx - r * DT_W;
            if (y0 + r < rows && x0 + c < cols) fb.dbg_score[(size_t)(y0 + r) * cols + x0 + c] = score[r + 1][c + 1];
        }
    }
    // strict 3x3 maxima among the pixels that have a score (the compacted list again: full warps); the maxima
    // are compacted like the list above (ballots in registers, one prefix sum, no atomics)
    {
        unsigned bal_it[DT_RT_ITERS];
        int idx_it[DT_RT_ITERS];
#pragma unroll
        for (int it = 0; it < DT_RT_ITERS; ++it) {
            const int e = it * 256 + threadIdx.x;
            bool is_max = false;
            int idx = 0;
            if (e < nlist) {
                const int li = s_list[e];
                const int r = (li >> 7) & 31, c = li & 127;  // score-window coordinates
                if (r >= 1 && r <= DT_H && c >= 1 && c <= DT_W) {
                    const int sc = score[r][c];
                    const int gy = y0 + r - 1, gx = x0 + c - 1;
                    idx = (r - 1) * DT_W + (c - 1);
                    if (sc != 0) {
                        is_max = sc > score[r - 1][c - 1] && sc > score[r - 1][c] && sc > score[r - 1][c + 1] && sc > score[r][c - 1] &&
                                 sc > score[r][c + 1] && sc > score[r + 1][c - 1] && sc > score[r + 1][c] && sc > score[r + 1][c + 1];
                        if (is_max) {
                            const int k = s_cy[r - 1] * fc.det_cols + s_cx[c - 1];
                            if (fb.det_occ[(size_t)s * fc.det_cells + k]) is_max = false;
                            if (gx < 5 || gy < 5 || gx > cols - 6 || gy > rows - 6) is_max = false;  // response 0: never a candidate
                        }
                    }
                }
            }
            bal_it[it] = __ballot_sync(0xffffffffu, is_max);
            idx_it[it] = is_max ? idx : -1;
        }
        __shared__ int s_wcnt2[8];
        int wtot = 0;
#pragma unroll
        for (int it = 0; it < DT_RT_ITERS; ++it) wtot += __popc(bal_it[it]);
        if (lane == 0) s_wcnt2[threadIdx.x >> 5] = wtot;
        __syncthreads();
        int wbase = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const int cnt = s_wcnt2[w];
            if (w < (int)(threadIdx.x >> 5)) wbase += cnt;
            total += cnt;
        }
        if (threadIdx.x == 0) s_nmax = total;
        const unsigned lt = (1u << lane) - 1u;
#pragma unroll
        for (int it = 0; it < DT_RT_ITERS; ++it) {
            if (idx_it[it] >= 0) s_max[wbase + __popc(bal_it[it] & lt)] = (unsigned short)idx_it[it];
            wbase += __popc(bal_it[it]);
        }
    }
    __syncthreads();
    // Shi-Tomasi response over the 8x8 box: 8 lanes per maximum, lane q sums row q - 4 of the box
    const int nmax = s_nmax;
    const int q = threadIdx.x & 7;
    for (int base = 0; base < nmax; base += 32) {
        const int e = base + (threadIdx.x >> 3);
        int dXX = 0, dYY = 0, dXY = 0, gy = 0, gx = 0;
        if (e < nmax) {
            const int idx = s_max[e];
            const int r = idx / DT_W, c = idx - r * DT_W;
            gy = y0 + r;
            gx = x0 + c;
            const uint8_t *row = &tile[r + DT_HALO + q - 4][c + DT_X0 - 4];
            int pm = row[-1], p0 = row[0];
#pragma unroll
            for (int xx = 0; xx < 8; ++xx) {
                const int pp = row[xx + 1];
                const int dx = pp - pm;
                const int dy = (int)row[xx + DT_STRIDE] - (int)row[xx - DT_STRIDE];
                dXX += dx * dx;
                dYY += dy * dy;
                dXY += dx * dy;
                pm = p0;
                p0 = pp;
            }
        }
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            dXX += __shfl_xor_sync(0xffffffffu, dXX, o);
            dYY += __shfl_xor_sync(0xffffffffu, dYY, o);
            dXY += __shfl_xor_sync(0xffffffffu, dXY, o);
        }
        if (q == 0 && e < nmax) {
            float fXX = (float)dXX / 128.0f, fYY = (float)dYY / 128.0f, fXY = (float)dXY / 128.0f;
            float trc = fXX + fYY;
            float d1 = fXX - fYY;
            float xy2 = fXY * fXY;
            float disc = d1 * d1 + 4.0f * xy2;
            float resp = 0.5f * (trc - sqrtf(disc));
            if (resp > 0.0f) {
                const int k = s_cy[gy - y0] * fc.det_cols + s_cx[gx - x0];
                unsigned long long key = ((unsigned long long)__float_as_uint(resp) << 32) |
                                         (unsigned long long)(0xffffffffu - (unsigned)(gy * cols + gx));
                atomicMax(&fb.det_best[(size_t)s * fc.det_cells + k], key);
            }
        }
    }
}

// ======================================================================================
// Point maps (cg::undistort_points / project_points, SPEC = cv::undistortPoints 5 iters)
// ======================================================================================
__device__ __forceinline__ float2 undistort_pt(const FeConst &fc, int cam, float2 p, const double *R) {
    const double *K = fc.K[cam], *D = fc.D[cam];
    double x = ((double)p.x - K[2]) / K[0], y = ((double)p.y - K[3]) / K[1];
    if (fc.cam_model[cam] == 0) {
        double x0 = x, y0 = y;
        for (int it = 0; it < 5; ++it) {
            double r2 = x * x + y * y;
            double icd = 1.0 / (1.0 + (D[1] * r2 + D[0]) * r2);
            double dx = 2.0 * D[2] * x * y + D[3] * (r2 + 2.0 * x * x);
            double dy = D[2] * (r2 + 2.0 * y * y) + 2.0 * D[3] * x * y;
            x = (x0 - dx) * icd;
            y = (y0 - dy) * icd;
        }
    } else {
        double thd = sqrt(x * x + y * y);
        thd = fmin(fmax(-1.5707963267948966, thd), 1.5707963267948966);
        double scale = 1.0;
        if (thd > 1e-8) {
            double th = thd;
            for (int it = 0; it < 10; ++it) {
                double t2 = th * th, t4 = t2 * t2, t6 = t4 * t2, t8 = t6 * t2;
                double k0t2 = D[0] * t2, k1t4 = D[1] * t4, k2t6 = D[2] * t6, k3t8 = D[3] * t8;
                double fix = (th * (1 + k0t2 + k1t4 + k2t6 + k3t8) - thd) / (1 + 3 * k0t2 + 5 * k1t4 + 7 * k2t6 + 9 * k3t8);
                th = th - fix;
                if (fabs(fix) < 1e-10) break;
            }
            scale = tan(th) / thd;
        }
        x *= scale;
        y *= scale;
    }
    if (R) {
        double X = R[0] * x + R[1] * y + R[2];
        double Y = R[3] * x + R[4] * y + R[5];
        double W = R[6] * x + R[7] * y + R[8];
        x = X / W;
        y = Y / W;
    } else {
        // identity rectification, evaluated with the same operations as the general case
        double X = 1.0 * x + 0.0 * y + 0.0;
        double Y = 0.0 * x + 1.0 * y + 0.0;
        double W = 0.0 * x + 0.0 * y + 1.0;
        x = X / W;
        y = Y / W;
    }
    return make_float2((float)(x * 1.0 + 0.0), (float)(y * 1.0 + 0.0));
}
__device__ __forceinline__ float2 distort_pt(const FeConst &fc, int cam, float2 p) {
    const double *K = fc.K[cam], *D = fc.D[cam];
    double x = (double)p.x, y = (double)p.y, xd, yd;
    if (fc.cam_model[cam] == 0) {
        double r2 = x * x + y * y;
        double cd = 1.0 + (D[1] * r2 + D[0]) * r2;
        xd = x * cd + 2.0 * D[2] * x * y + D[3] * (r2 + 2.0 * x * x);
        yd = y * cd + D[2] * (r2 + 2.0 * y * y) + 2.0 * D[3] * x * y;
    } else {
        double r = sqrt(x * x + y * y);
        double th = atan(r);
        double t2 = th * th, t4 = t2 * t2, t6 = t4 * t2, t8 = t4 * t4;
        double thd = th * (1 + D[0] * t2 + D[1] * t4 + D[2] * t6 + D[3] * t8);
        double sc = r > 1e-8 ? thd / r : 1.0;
        xd = x * sc;
        yd = y * sc;
    }
    return make_float2((float)(xd * K[0] + K[2]), (float)(yd * K[1] + K[3]));
}
// stereoMatch epipolar gate, image_processor.cpp:587-617
__device__ __forceinline__ bool epipolar_ok(const FeConst &fc, float2 c0, float2 c1) {
    float2 u0 = undistort_pt(fc, 0, c0, nullptr), u1 = undistort_pt(fc, 1, c1, nullptr);
    double p0x = (double)u0.x, p0y = (double)u0.y, p1x = (double)u1.x, p1y = (double)u1.y;
    const double *E = fc.E;
    double l0 = E[0] * p0x + E[1] * p0y + E[2] * 1.0;
    double l1 = E[3] * p0x + E[4] * p0y + E[5] * 1.0;
    double l2 = E[6] * p0x + E[7] * p0y + E[8] * 1.0;
    double error = fabs(p1x * l0 + p1y * l1 + 1.0 * l2) / sqrt(l0 * l0 + l1 * l1);
    return !(error > fc.stereo_gate);
}
__device__ __forceinline__ bool in_image(const FeConst &fc, float2 p) {
    return !(p.y < 0 || p.y > (float)(fc.rows - 1) || p.x < 0 || p.x > (float)(fc.cols - 1));
}
__device__ __forceinline__ int grid_code(const FeConst &fc, float2 p) {
    int row = (int)(p.y / (float)fc.grid_h);
    int col = (int)(p.x / (float)fc.grid_w);
    return row * fc.grid_col + col;
}

// compulsory window traffic of one feature through all levels: an (w+2)^2 template patch and a
// (w+1)^2 search patch per level
__host__ __device__ __forceinline__ double klt_bytes_per_feature(const FeConst &fc) {
    return (double)fc.levels * (double)((fc.klt_win + 2) * (fc.klt_win + 2) + (fc.klt_win + 1) * (fc.klt_win + 1));
}

#define FE_THREADS 128
#define FE_MAX_CELLS 128  // coarse grid cells incl. the overflow row (image_processor.cpp:663-665 can index past grid_row)

// Order-preserving block compaction: returns the number kept; dst index of element i
// (when flag) is written to pos[i].  Executed by all FE_THREADS threads.
__device__ int block_compact_positions(const uint8_t *flag, int n, int *pos, int *s_warp) {
    __shared__ int s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < n; start += FE_THREADS) {
        int i = start + threadIdx.x;
        int f = (i < n) ? (flag[i] != 0) : 0;
        unsigned bal = __ballot_sync(0xffffffffu, f);
        int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (f) pos[i] = off + __popc(bal & ((1u << lane) - 1));
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < FE_THREADS / 32; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    return s_base;
}

// --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(FE_THREADS) fe_prep_track(FeConst fc, FeBuffers fb) {
    const int s = blockIdx.x;
    const FeStep st = fb.step[s];
    if (!st.active) return;
    if (st.is_first) {
        if (threadIdx.x == 0) fb.k_n[s] = 0;
        return;
    }
    const int gp = fb.gslot[s];  // prev grid buffer
    const int n = fb.g_n[gp][s];
    const size_t go = (size_t)s * fc.max_f, ko = (size_t)s * fc.cap_k;
    for (int i = threadIdx.x; i < n; i += FE_THREADS) {
        float2 p = fb.g_cam0[gp][go + i];
        fb.k_a[ko + i] = p;
        double x = (double)p.x, y = (double)p.y;
        const double *H = st.H0;
        double p20 = H[0] * x + H[1] * y + H[2] * 1.0;
        double p21 = H[3] * x + H[4] * y + H[5] * 1.0;
        double p22 = H[6] * x + H[7] * y + H[8] * 1.0;
        fb.k_b[ko + i] = make_float2((float)(p20 / p22), (float)(p21 / p22));
        fb.t_id[go + i] = fb.g_id[gp][go + i];
        fb.t_life[go + i] = fb.g_life[gp][go + i];
        fb.t_p0[go + i] = p;
        fb.t_p1[go + i] = fb.g_cam1[gp][go + i];
    }
    if (threadIdx.x == 0) {
        fb.k_n[s] = n;
        fb.info[s].before_tracking = n;
        fb.work[(size_t)s * MSKF_PROF_TAGS + PK_KLT_TEMPORAL] += (double)n * klt_bytes_per_feature(fc);
    }
}

// after the temporal track: bounds check, compaction, stereo initial guess
__global__ void __launch_bounds__(FE_THREADS) fe_after_track(FeConst fc, FeBuffers fb) {
    const int s = blockIdx.x;
    const FeStep st = fb.step[s];
    if (!st.active || st.is_first) return;
    __shared__ int s_warp[FE_THREADS / 32];
    extern __shared__ int s_pos[];  // [max_f]
    const int n = fb.k_n[s];
    if (n == 0) return;  // trackFeatures returns early (:383); counters keep their old values
    const size_t go = (size_t)s * fc.max_f, ko = (size_t)s * fc.cap_k;
    uint8_t *flag = fb.k_status + ko;
    for (int i = threadIdx.x; i < n; i += FE_THREADS)
        if (flag[i] && !in_image(fc, fb.k_b[ko + i])) flag[i] = 0;
    __syncthreads();
    int m = block_compact_positions(flag, n, s_pos, s_warp);
    // gather into registers, then scatter (positions are <= source index; two-phase keeps it race free)
    for (int start = 0; start < n; start += FE_THREADS) {
        int i = start + threadIdx.x;
        bool keep = i < n && flag[i];
        float2 c0 = make_float2(0, 0), q0 = c0, q1 = c0;
        unsigned long long id = 0;
        int life = 0;
        if (keep) {
            c0 = fb.k_b[ko + i];
            id = fb.t_id[go + i];
            life = fb.t_life[go + i];
            q0 = fb.t_p0[go + i];
            q1 = fb.t_p1[go + i];
        }
        __syncthreads();
        if (keep) {
            int d = s_pos[i];
            fb.k_a[ko + d] = c0;  // current cam0 point becomes the stereo template point
            fb.t_id[go + d] = id;
            fb.t_life[go + d] = life;
            fb.t_p0[go + d] = q0;
            fb.t_p1[go + d] = q1;
            float2 u = undistort_pt(fc, 0, c0, fc.R01);
            fb.k_b[ko + d] = distort_pt(fc, 1, u);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        fb.k_n[s] = m;
        fb.info[s].after_tracking = m;
        fb.work[(size_t)s * MSKF_PROF_TAGS + PK_KLT_STEREO] += (double)m * klt_bytes_per_feature(fc);
    }
}

// ======================================================================================
// twoPointRansac (image_processor.cpp:911-1135) for one camera, executed by the whole CTA of a
// stream.  Dead code in the reference (both calls are commented out, :482-493): runs only with
// use_ransac.  The sequential float / double sums of the reference are kept sequential (one thread)
// so that the result is bit-identical to the oracle; the seven hypotheses run on seven threads.
// cg::uniform_integer is unseeded in the reference: SPEC = counter-based hash shared with the oracle.
// ======================================================================================
__device__ __forceinline__ unsigned ransac_hash(unsigned a, unsigned b) {
    unsigned h = a * 0x9E3779B1u ^ (b + 0x7F4A7C15u) * 0x85EBCA77u;
    h ^= h >> 15; h *= 0x2C1B3C6Du;
    h ^= h >> 12; h *= 0x297A2D39u;
    h ^= h >> 15;
    return h;
}
__device__ __forceinline__ int ransac_uniform(unsigned call, int cam, int iter, int draw, int lo, int hi) {
    return lo + (int)(ransac_hash(call * 2u + (unsigned)cam, (unsigned)(iter * 2 + draw)) % (unsigned)(hi - lo + 1));
}
#define RANSAC_ITERS 7  // ceil(log(1 - 0.99) / log(1 - 0.7 * 0.7)), image_processor.cpp:927

struct RansacSmem {       // carved from dynamic shared memory, n = matched features
    float2 *p1, *p2;      // [n]
    double *coeff;        // [n][3]
    double *dist;         // [n]
    float *sq;            // [2n]
    int *raw;             // [n]
    unsigned *sets;       // [RANSAC_ITERS][words]
    uint8_t *marker;      // [n]
};

__device__ void ransac_one_cam(const FeConst &fc, int cam, const double *R, int n, const int *mlist, const float2 *prev, const float2 *curr,
                               unsigned call, RansacSmem sm, uint8_t *out_marker) {
    __shared__ float s_sf;
    __shared__ double s_npu, s_mean;
    __shared__ int s_cnt, s_mode, s_nraw, s_count[RANSAC_ITERS], s_best;
    const int words = (n + 31) / 32;
    for (int j = threadIdx.x; j < n; j += FE_THREADS) {
        const int i = mlist[j];
        float2 u1 = undistort_pt(fc, cam, prev[i], nullptr), u2 = undistort_pt(fc, cam, curr[i], nullptr);
        const double x = (double)u1.x, y = (double)u1.y;
        u1.x = (float)(R[0] * x + R[1] * y + R[2] * 1.0);
        u1.y = (float)(R[3] * x + R[4] * y + R[5] * 1.0);
        sm.p1[j] = u1;
        sm.p2[j] = u2;
        sm.sq[2 * j] = sqrtf(u1.x * u1.x + u1.y * u1.y);
        sm.sq[2 * j + 1] = sqrtf(u2.x * u2.x + u2.y * u2.y);
    }
    __syncthreads();
    if (threadIdx.x == 0) {  // rescalePoints, image_processor.cpp:891-909 (sequential float sum)
        float sf = 0.0f;
        for (int j = 0; j < n; ++j) {
            sf += sm.sq[2 * j];
            sf += sm.sq[2 * j + 1];
        }
        sf = (float)(2 * n) / sf * sqrtf(2.0f);
        s_sf = sf;
        s_npu = 2.0 / (fc.K[cam][0] + fc.K[cam][1]) * (double)sf;
    }
    __syncthreads();
    const float sf = s_sf;
    const double npu = s_npu;
    for (int j = threadIdx.x; j < n; j += FE_THREADS) {
        float2 a = sm.p1[j], b = sm.p2[j];
        a.x *= sf; a.y *= sf; b.x *= sf; b.y *= sf;
        const float dx = a.x - b.x, dy = a.y - b.y;
        const double d = (double)sqrtf(dx * dx + dy * dy);
        sm.dist[j] = d;
        sm.marker[j] = d > 50.0 * npu ? 0 : 1;
        sm.coeff[3 * j + 0] = (double)dy;
        sm.coeff[3 * j + 1] = (double)(-dx);
        sm.coeff[3 * j + 2] = (double)(a.x * b.y - a.y * b.x);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double mean = 0.0;
        int cnt = 0, nraw = 0;
        for (int j = 0; j < n; ++j)
            if (sm.marker[j]) {
                mean += sm.dist[j];
                ++cnt;
                sm.raw[nraw++] = j;
            }
        mean /= cnt;
        s_mean = mean;
        s_cnt = cnt;
        s_nraw = nraw;
        s_mode = cnt < 3 ? 0 : (mean < npu ? 1 : 2);
    }
    __syncthreads();
    if (s_mode == 0) {
        for (int j = threadIdx.x; j < n; j += FE_THREADS) out_marker[j] = 0;
        __syncthreads();
        return;
    }
    if (s_mode == 1) {  // degenerate motion: threshold the point distance
        for (int j = threadIdx.x; j < n; j += FE_THREADS) out_marker[j] = (sm.marker[j] && !(sm.dist[j] > fc.ransac_threshold * npu)) ? 1 : 0;
        __syncthreads();
        return;
    }
    for (int e = threadIdx.x; e < RANSAC_ITERS * words; e += FE_THREADS) sm.sets[e] = 0u;
    __syncthreads();
    if (threadIdx.x < RANSAC_ITERS) {
        const int it = threadIdx.x, nraw = s_nraw;
        const int i1 = ransac_uniform(call, cam, it, 0, 0, nraw - 1);
        const int idf = ransac_uniform(call, cam, it, 1, 1, nraw - 1);
        const int i2 = i1 + idf < nraw ? i1 + idf : i1 + idf - nraw;
        const int pa = sm.raw[i1], pb = sm.raw[i2];
        double c[3][2];
        for (int k = 0; k < 3; ++k) {
            c[k][0] = sm.coeff[3 * pa + k];
            c[k][1] = sm.coeff[3 * pb + k];
        }
        double l1[3];
        for (int k = 0; k < 3; ++k) l1[k] = fabs(c[k][0]) + fabs(c[k][1]);
        int base = 0;
        for (int k = 1; k < 3; ++k)
            if (l1[k] < l1[base]) base = k;
        const int ia = base == 0 ? 1 : 0, ib = base == 2 ? 1 : 2;
        double mdl[3];
        {
            const double a00 = c[ia][0], a01 = c[ib][0], a10 = c[ia][1], a11 = c[ib][1];
            const double det = a00 * a11 - a01 * a10, id = 1.0 / det;
            const double i00 = a11 * id, i01 = -a01 * id, i10 = -a10 * id, i11 = a00 * id;
            const double b0 = -c[base][0], b1 = -c[base][1];
            mdl[base] = 1.0;
            mdl[ia] = i00 * b0 + i01 * b1;
            mdl[ib] = i10 * b0 + i11 * b1;
        }
        int count = 0;
        unsigned *set = sm.sets + it * words;
        for (int j = 0; j < n; ++j) {
            if (!sm.marker[j]) continue;
            double e = sm.coeff[3 * j] * mdl[0];
            e += sm.coeff[3 * j + 1] * mdl[1];
            e += sm.coeff[3 * j + 2] * mdl[2];
            if (fabs(e) < fc.ransac_threshold * npu) {
                set[j >> 5] |= 1u << (j & 31);
                ++count;
            }
        }
        // hypotheses with too few inliers are skipped (:1064); the refit (:1067-1112) only feeds
        // best_error, which never influences the result
        s_count[it] = ((double)count < 0.2 * (double)n) ? -1 : count;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int best = -1, best_count = 0;
        for (int it = 0; it < RANSAC_ITERS; ++it)
            if (s_count[it] > best_count) {
                best_count = s_count[it];
                best = it;
            }
        s_best = best;
    }
    __syncthreads();
    const int best = s_best;
    for (int j = threadIdx.x; j < n; j += FE_THREADS)
        out_marker[j] = (best >= 0 && ((sm.sets[best * words + (j >> 5)] >> (j & 31)) & 1u)) ? 1 : 0;
    __syncthreads();
}

// after the stereo match of tracked features: gates, compaction, re-bucket into the grid
__global__ void __launch_bounds__(FE_THREADS) fe_after_stereo(FeConst fc, FeBuffers fb) {
    const int s = blockIdx.x;
    const FeStep st = fb.step[s];
    if (!st.active) return;
    __shared__ int s_cnt[FE_MAX_CELLS], s_start[FE_MAX_CELLS + 1];
    extern __shared__ int s_dyn[];
    uint8_t *s_code = (uint8_t *)s_dyn;  // [max_f] grid cell of each survivor, 255 = dropped
    const size_t go = (size_t)s * fc.max_f, ko = (size_t)s * fc.cap_k;
    const int gc = fb.gslot[s] ^ 1;  // curr grid buffer
    // reset the detector tables of this stream
    for (int k = threadIdx.x; k < fc.det_cells; k += FE_THREADS) {
        fb.det_best[(size_t)s * fc.det_cells + k] = 0ull;
        fb.det_occ[(size_t)s * fc.det_cells + k] = 0;
    }
    const int n = st.is_first ? 0 : fb.k_n[s];
    const bool tracked_any = !st.is_first && fb.info[s].before_tracking > 0;
    const uint8_t *flag = fb.k_status + ko;
    for (int i = threadIdx.x; i < n; i += FE_THREADS) {
        uint8_t code = 255;
        if (flag[i]) {
            float2 c0 = fb.k_a[ko + i], c1 = fb.k_b[ko + i];
            if (in_image(fc, c1) && epipolar_ok(fc, c0, c1)) code = (uint8_t)grid_code(fc, c0);
        }
        s_code[i] = code;
    }
    for (int c = threadIdx.x; c < fc.n_cells_all; c += FE_THREADS) s_cnt[c] = 0;
    __syncthreads();
    if (fc.use_ransac && tracked_any) {
        // matched list in index order, then twoPointRansac on the temporal pairs of cam0 and cam1
        // (the commented-out calls at image_processor.cpp:482-493)
        __shared__ int s_nm;
        uint8_t *base = (uint8_t *)s_dyn + ((fc.max_f + 15) & ~15);
        int *mlist = (int *)base;                                   // [max_f]
        RansacSmem sm;
        sm.p1 = (float2 *)(mlist + fc.max_f);
        sm.p2 = sm.p1 + fc.max_f;
        sm.coeff = (double *)(sm.p2 + fc.max_f);
        sm.dist = sm.coeff + 3 * fc.max_f;
        sm.sq = (float *)(sm.dist + fc.max_f);
        sm.raw = (int *)(sm.sq + 2 * fc.max_f);
        sm.sets = (unsigned *)(sm.raw + fc.max_f);
        sm.marker = (uint8_t *)(sm.sets + RANSAC_ITERS * ((fc.max_f + 31) / 32));
        uint8_t *mk0 = sm.marker + fc.max_f, *mk1 = mk0 + fc.max_f;
        if (threadIdx.x == 0) {
            int m = 0;
            for (int i = 0; i < n; ++i)
                if (s_code[i] != 255) mlist[m++] = i;
            s_nm = m;
            fb.info[s].after_matching = m;
        }
        __syncthreads();
        const int nm = s_nm;
        const unsigned call = fb.track_calls[s];
        if (nm > 0) {
            ransac_one_cam(fc, 0, st.R0, nm, mlist, fb.t_p0 + go, fb.k_a + ko, call, sm, mk0);
            ransac_one_cam(fc, 1, st.R1, nm, mlist, fb.t_p1 + go, fb.k_b + ko, call, sm, mk1);
            for (int j = threadIdx.x; j < nm; j += FE_THREADS)
                if (mk0[j] == 0 || mk1[j] == 0) s_code[mlist[j]] = 255;
        }
        __syncthreads();
        if (threadIdx.x == 0) fb.track_calls[s] = call + 1;
    }
    // stable counting sort by grid cell = publish order of the std::map; one warp per cell
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = warp; c < fc.n_cells_all; c += FE_THREADS / 32) {
        int cnt = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            int i = i0 + lane;
            cnt += __popc(__ballot_sync(0xffffffffu, i < n && s_code[i] == c));
        }
        if (lane == 0) s_cnt[c] = cnt;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int c = 0; c < fc.n_cells_all; ++c) {
            s_start[c] = acc;
            acc += s_cnt[c];
        }
        s_start[fc.n_cells_all] = acc;
        fb.g_n[gc][s] = acc;
        if (tracked_any) {
            if (!fc.use_ransac) fb.info[s].after_matching = acc;
            fb.info[s].after_ransac = acc;
        }
    }
    __syncthreads();
    for (int c = warp; c < fc.n_cells_all; c += FE_THREADS / 32) {
        int off = s_start[c];
        for (int i0 = 0; i0 < n; i0 += 32) {
            int i = i0 + lane;
            bool mine = i < n && s_code[i] == c;
            unsigned bal = __ballot_sync(0xffffffffu, mine);
            if (mine) {
                int d = off + __popc(bal & ((1u << lane) - 1));
                float2 c0 = fb.k_a[ko + i];
                fb.g_id[gc][go + d] = fb.t_id[go + i];
                fb.g_life[gc][go + d] = fb.t_life[go + i] + 1;
                fb.g_resp[gc][go + d] = 0.0f;
                fb.g_cam0[gc][go + d] = c0;
                fb.g_cam1[gc][go + d] = fb.k_b[ko + i];
                fb.g_cell[gc][go + d] = c;
                // CornerDetector::set_grid_position on the truncated position (:634-647)
                int x = (int)c0.x, y = (int)c0.y;
                if (x >= 0 && y >= 0 && x < fc.cols && y < fc.rows)
                    fb.det_occ[(size_t)s * fc.det_cells + (y / fc.det_cell_h) * fc.det_cols + (x / fc.det_cell_w)] = 1;
            }
            off += __popc(bal);
        }
    }
}

// detector output -> candidate list for the stereo match of new features
__global__ void __launch_bounds__(FE_THREADS) fe_sieve(FeConst fc, FeBuffers fb, int members_cap) {
    const int s = blockIdx.x;
    const FeStep st = fb.step[s];
    if (!st.active) return;
    __shared__ int s_warp[FE_THREADS / 32];
    __shared__ int s_cell_n[FE_MAX_CELLS], s_cell_off[FE_MAX_CELLS + 1], s_vac[FE_MAX_CELLS];
    __shared__ int s_base;
    extern __shared__ int s_dyn[];
    // dynamic: resp[det_cells] | xy[det_cells] (packed y << 16 | x) | sel[n_cells][grid_max] | mem[4][members_cap] | code[det_cells]
    float *s_resp = (float *)s_dyn;
    unsigned *s_xy = (unsigned *)(s_resp + fc.det_cells);
    int *s_sel = (int *)(s_xy + fc.det_cells);
    int *s_mem = s_sel + fc.n_cells * fc.grid_max;
    uint8_t *s_code = (uint8_t *)(s_mem + (FE_THREADS / 32) * members_cap);
    const size_t ko = (size_t)s * fc.cap_k, dofs = (size_t)s * fc.det_cells;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // compact the per-fine-cell winners (cell-major = detect order) into shared memory
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int start = 0; start < fc.det_cells; start += FE_THREADS) {
        int k = start + threadIdx.x;
        unsigned long long key = k < fc.det_cells ? fb.det_best[dofs + k] : 0ull;
        float resp = __uint_as_float((unsigned)(key >> 32));
        int f = (key != 0ull && (double)resp > fc.detection_threshold) ? 1 : 0;
        unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (f) {
            int d = off + __popc(bal & ((1u << lane) - 1));
            unsigned ridx = 0xffffffffu - (unsigned)(key & 0xffffffffull);
            unsigned x = ridx % (unsigned)fc.cols, y = ridx / (unsigned)fc.cols;
            s_resp[d] = resp;
            s_xy[d] = (y << 16) | x;
            s_code[d] = (uint8_t)grid_code(fc, make_float2((float)x, (float)y));
            fb.nf_resp[dofs + d] = resp;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int tot = 0;
            for (int w = 0; w < FE_THREADS / 32; ++w) tot += s_warp[w];
            s_base += tot;
        }
        __syncthreads();
    }
    const int n = s_base;
    if (threadIdx.x == 0) fb.nf_n[s] = n;
    auto det_pt = [&](int i) { return make_float2((float)(s_xy[i] & 0xffffu), (float)(s_xy[i] >> 16)); };
    if (st.is_first) {
        for (int i = threadIdx.x; i < n; i += FE_THREADS) {
            float2 p = det_pt(i);
            fb.k_a[ko + i] = p;
            fb.k_b[ko + i] = distort_pt(fc, 1, undistort_pt(fc, 0, p, fc.R01));
            fb.k_skip[ko + i] = 0;
            fb.k_idx[ko + i] = i;
        }
        if (threadIdx.x == 0) {
            fb.k_n[s] = n;
            fb.k_nm[s] = n;
            fb.work[(size_t)s * MSKF_PROF_TAGS + PK_KLT_NEW] += (double)n * klt_bytes_per_feature(fc);
        }
        return;
    }
    // one warp per coarse cell: members in detect order; more than grid_max -> the grid_max best
    // responses (stable: ties keep detect order), image_processor.cpp:668-677
    int *mem = s_mem + warp * members_cap;
    for (int c = warp; c < fc.n_cells; c += FE_THREADS / 32) {
        int cnt = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            int i = i0 + lane;
            bool mine = i < n && s_code[i] == c;
            unsigned bal = __ballot_sync(0xffffffffu, mine);
            if (mine) {
                int d = cnt + __popc(bal & ((1u << lane) - 1));
                if (d < members_cap) mem[d] = i;
            }
            cnt += __popc(bal);
        }
        cnt = min(cnt, members_cap);
        __syncwarp();
        int *sel = s_sel + c * fc.grid_max;
        int kept = 0;
        if (cnt <= fc.grid_max) {
            for (int d = lane; d < cnt; d += 32) sel[d] = mem[d];
            kept = cnt;
        } else {
            float last_r = 3.0e38f;
            int last_i = -1;
            for (int k = 0; k < fc.grid_max; ++k) {
                // best not yet taken: max response, ties -> smallest index
                float br = -1.0f;
                int bi = 0x7fffffff;
                for (int d = lane; d < cnt; d += 32) {
                    int i = mem[d];
                    float r = s_resp[i];
                    bool after_last = (r < last_r) || (r == last_r && i > last_i);
                    if (after_last && (r > br || (r == br && i < bi))) { br = r; bi = i; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    float r2 = __shfl_xor_sync(0xffffffffu, br, o);
                    int i2 = __shfl_xor_sync(0xffffffffu, bi, o);
                    if (r2 > br || (r2 == br && i2 < bi)) { br = r2; bi = i2; }
                }
                if (lane == 0) sel[k] = bi;
                last_r = br;
                last_i = bi;
            }
            kept = fc.grid_max;
        }
        if (lane == 0) s_cell_n[c] = kept;
        __syncwarp();
    }
    // cells already holding grid_min tracked features take no new ones: their candidates keep their
    // place in the list (the indices matter, image_processor.cpp:698) but are not stereo-matched
    {
        const int gc = fb.gslot[s] ^ 1;
        const int ncur = fb.g_n[gc][s];
        const size_t go = (size_t)s * fc.max_f;
        for (int c = warp; c < fc.n_cells; c += FE_THREADS / 32) {
            int cur = 0;
            for (int i0 = 0; i0 < ncur; i0 += 32) {
                int i = i0 + lane;
                cur += __popc(__ballot_sync(0xffffffffu, i < ncur && fb.g_cell[gc][go + i] == c));
            }
            if (lane == 0) s_vac[c] = cur < fc.grid_min ? 1 : 0;
        }
    }
    __syncthreads();
    __shared__ int s_nm;
    if (threadIdx.x == 0) {
        s_nm = 0;
        int acc = 0, matched = 0;
        for (int c = 0; c < fc.n_cells; ++c) {
            s_cell_off[c] = acc;
            acc += s_cell_n[c];
            if (s_vac[c]) matched += s_cell_n[c];
        }
        s_cell_off[fc.n_cells] = acc;
        fb.k_n[s] = acc;
        fb.work[(size_t)s * MSKF_PROF_TAGS + PK_KLT_NEW] += (double)matched * klt_bytes_per_feature(fc);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < fc.n_cells * fc.grid_max; e += FE_THREADS) {
        int c = e / fc.grid_max, k = e - c * fc.grid_max;
        if (k >= s_cell_n[c]) continue;
        float2 p = det_pt(s_sel[e]);
        int d = s_cell_off[c] + k;
        fb.k_a[ko + d] = p;
        fb.k_b[ko + d] = distort_pt(fc, 1, undistort_pt(fc, 0, p, fc.R01));
        fb.k_skip[ko + d] = s_vac[c] ? 0 : 1;
        // the stereo match of the new candidates runs over the compacted list of the ones in cells with a
        // vacancy (order irrelevant: results are written by list index); the others are settled here
        if (s_vac[c]) fb.k_idx[ko + atomicAdd(&s_nm, 1)] = d;
        else fb.k_status[ko + d] = 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) fb.k_nm[s] = s_nm;
}

// stereo-matched new features -> grid, prune, publish, rotate prev/curr
__global__ void __launch_bounds__(FE_THREADS) fe_finish(FeConst fc, FeBuffers fb) {
    const int s = blockIdx.x;
    const FeStep st = fb.step[s];
    if (!st.active) return;
    __shared__ int s_warp[FE_THREADS / 32];
    __shared__ int s_m, s_add[FE_MAX_CELLS], s_keep[FE_MAX_CELLS], s_idbase[FE_MAX_CELLS + 1], s_outoff[FE_MAX_CELLS + 1];
    __shared__ unsigned long long s_base_id;
    extern __shared__ int s_dyn[];
    const size_t go = (size_t)s * fc.max_f, ko = (size_t)s * fc.cap_k, dofs = (size_t)s * fc.det_cells;
    const int gc = fb.gslot[s] ^ 1, gp = fb.gslot[s];
    const int n = fb.k_n[s];
    int *s_pos = s_dyn;                                  // [cap_k]
    int *s_addsel = s_dyn + fc.cap_k;                    // [n_cells][grid_min] inlier indices to add
    int *s_final = s_addsel + fc.n_cells * fc.grid_min;  // [n_cells_all][grid_max] encoded final members
    float *s_iresp = (float *)(s_final + fc.n_cells_all * fc.grid_max);  // [cap_k] inlier responses
    int *s_glife = (int *)(s_iresp + fc.cap_k);          // [max_f] lifetimes of the tracked grid entries
    uint8_t *s_icode = (uint8_t *)(s_glife + fc.max_f);  // [cap_k] inlier cells
    uint8_t *s_gcell = s_icode + fc.cap_k;               // [max_f] cells of the tracked grid entries
    uint8_t *flag = fb.k_status + ko;
    for (int i = threadIdx.x; i < n; i += FE_THREADS) {
        if (flag[i]) {
            float2 c1 = fb.k_b[ko + i];
            if (!in_image(fc, c1) || !epipolar_ok(fc, fb.k_a[ko + i], c1)) flag[i] = 0;
        }
    }
    __syncthreads();
    int m = block_compact_positions(flag, n, s_pos, s_warp);
    // compact inliers in place (dst <= src): cam0 -> k_a, cam1 -> k_b
    // response_inliers[j] = new_features_responses[i] with i the POST-sieve index (:698)
    float *resp_in = fb.in_resp + ko;
    for (int start = 0; start < n; start += FE_THREADS) {
        int i = start + threadIdx.x;
        bool keep = i < n && flag[i];
        float2 a, b;
        float r = 0;
        if (keep) {
            a = fb.k_a[ko + i];
            b = fb.k_b[ko + i];
            r = fb.nf_resp[dofs + i];
        }
        __syncthreads();
        if (keep) {
            int d = s_pos[i];
            fb.k_a[ko + d] = a;
            fb.k_b[ko + d] = b;
            resp_in[d] = r;
            s_iresp[d] = r;
            s_icode[d] = (uint8_t)grid_code(fc, a);
        }
        __syncthreads();
    }
    const int ncur = st.is_first ? 0 : fb.g_n[gc][s];
    for (int i = threadIdx.x; i < ncur; i += FE_THREADS) {
        s_gcell[i] = (uint8_t)fb.g_cell[gc][go + i];
        s_glife[i] = fb.g_life[gc][go + i];
    }
    __syncthreads();
    // one WARP per cell: choose the new features to add (best responses, stable), then the members that
    // survive pruneGridFeatures.  Lanes scan the candidate / member lists 32 at a time; a selection round
    // is a warp arg-max over a 64-bit key (value, reversed position), so ties keep list order.
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const unsigned lt_mask = (1u << lane) - 1u;
        for (int c = warp; c < fc.n_cells_all; c += FE_THREADS / 32) {
            int cur = 0;
            for (int i0 = 0; i0 < ncur; i0 += 32) {
                const int i = i0 + lane;
                cur += __popc(__ballot_sync(0xffffffffu, i < ncur && s_gcell[i] == c));
            }
            int add = 0;
            if (c < fc.n_cells && cur < fc.grid_min) {
                const int vacancy = fc.grid_min - cur;
                float last_r = 3.0e38f;
                int last_i = -1;
                for (int k = 0; k < vacancy; ++k) {
                    // best (largest response, then smallest index) candidate of this cell after (last_r, last_i)
                    unsigned long long best = 0ull;  // (response bits | 1 << 63 never set: responses are >= 0) << 32 | ~index
                    bool have = false;
                    for (int i = lane; i < m; i += 32) {
                        if (s_icode[i] != c) continue;
                        const float r = s_iresp[i];
                        const bool after_last = (r < last_r) || (r == last_r && i > last_i);
                        if (!after_last) continue;
                        const unsigned long long key = ((unsigned long long)__float_as_uint(r) << 32) | (unsigned)(0x7fffffff - i);
                        if (!have || key > best) { best = key; have = true; }
                    }
                    unsigned long long kb = have ? (best | (1ull << 63)) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const unsigned long long other = __shfl_xor_sync(0xffffffffu, kb, o);
                        kb = other > kb ? other : kb;
                    }
                    if (!(kb >> 63)) break;
                    const int bi = 0x7fffffff - (int)(unsigned)(kb & 0xffffffffu);
                    const float br = __uint_as_float((unsigned)((kb >> 32) & 0x7fffffffu));
                    if (lane == 0) s_addsel[c * fc.grid_min + add] = bi;
                    ++add;
                    last_r = br;
                    last_i = bi;
                }
            }
            // members after the additions: tracked entries (index i, lifetime from the grid) then
            // new ones (lifetime 1); pruneGridFeatures keeps the grid_max longest-lived (stable)
            const int total = cur + add;
            int keep = 0;
            int *fin = s_final + c * fc.grid_max;
            if (total <= fc.grid_max) {
                for (int i0 = 0; i0 < ncur; i0 += 32) {
                    const int i = i0 + lane;
                    const bool mine = i < ncur && s_gcell[i] == c;
                    const unsigned bal = __ballot_sync(0xffffffffu, mine);
                    if (mine) fin[keep + __popc(bal & lt_mask)] = i;
                    keep += __popc(bal);
                }
                for (int a = lane; a < add; a += 32) fin[keep + a] = -1 - a;
                keep += add;
            } else {
                int last_l = 0x7fffffff, last_o = -1;
                for (int k = 0; k < fc.grid_max; ++k) {
                    // key: (lifetime << 32 | ~order) with the member's encoding as payload
                    long long best = -1;
                    int benc = 0, o_base = 0;
                    for (int i0 = 0; i0 < ncur; i0 += 32) {
                        const int i = i0 + lane;
                        const bool mine = i < ncur && s_gcell[i] == c;
                        const unsigned bal = __ballot_sync(0xffffffffu, mine);
                        if (mine) {
                            const int o = o_base + __popc(bal & lt_mask), l = s_glife[i];
                            const bool after_last = (l < last_l) || (l == last_l && o > last_o);
                            const long long key = ((long long)l << 32) | (unsigned)(0x7fffffff - o);
                            if (after_last && key > best) { best = key; benc = i; }
                        }
                        o_base += __popc(bal);
                    }
                    for (int a = lane; a < add; a += 32) {
                        const int o = o_base + a, l = 1;
                        const bool after_last = (l < last_l) || (l == last_l && o > last_o);
                        const long long key = ((long long)l << 32) | (unsigned)(0x7fffffff - o);
                        if (after_last && key > best) { best = key; benc = -1 - a; }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        const long long ob = __shfl_xor_sync(0xffffffffu, best, o);
                        const int oe = __shfl_xor_sync(0xffffffffu, benc, o);
                        if (ob > best) { best = ob; benc = oe; }
                    }
                    if (lane == 0) fin[keep] = benc;
                    ++keep;
                    last_l = (int)(best >> 32);
                    last_o = 0x7fffffff - (int)(unsigned)(best & 0xffffffffu);
                }
            }
            if (lane == 0) {
                s_add[c] = add;
                s_keep[c] = keep;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long base = fb.next_id[s];
        int acc_id = 0, acc_out = 0;
        for (int c = 0; c < fc.n_cells_all; ++c) {
            s_idbase[c] = acc_id;
            s_outoff[c] = acc_out;
            acc_id += s_add[c];
            acc_out += s_keep[c];
        }
        s_outoff[fc.n_cells_all] = acc_out;
        s_base_id = base;
        fb.next_id[s] = base + (unsigned long long)acc_id;
        fb.g_n[gp][s] = acc_out;  // the finished grid goes to the buffer that was "prev"
        fb.msg_n[s] = acc_out;
        s_m = acc_out;
    }
    __syncthreads();
    const unsigned long long base = s_base_id;
    // write the finished grid (publish order) into buffer gp and the measurement message
    for (int e = threadIdx.x; e < fc.n_cells_all * fc.grid_max; e += FE_THREADS) {
        const int c = e / fc.grid_max, k = e - c * fc.grid_max;
        if (k >= s_keep[c]) continue;
        int enc = s_final[e];
        int d = s_outoff[c] + k;
        unsigned long long id;
        float resp;
        int life;
        float2 c0, c1;
        if (enc >= 0) {
            id = fb.g_id[gc][go + enc];
            resp = fb.g_resp[gc][go + enc];
            life = s_glife[enc];
            c0 = fb.g_cam0[gc][go + enc];
            c1 = fb.g_cam1[gc][go + enc];
        } else {
            int a = -1 - enc;
            int src = s_addsel[c * fc.grid_min + a];
            id = base + (unsigned long long)(s_idbase[c] + a);
            resp = s_iresp[src];
            life = 1;
            c0 = fb.k_a[ko + src];
            c1 = fb.k_b[ko + src];
        }
        fb.g_id[gp][go + d] = id;
        fb.g_resp[gp][go + d] = resp;
        fb.g_life[gp][go + d] = life;
        fb.g_cam0[gp][go + d] = c0;
        fb.g_cam1[gp][go + d] = c1;
        fb.g_cell[gp][go + d] = c;
        float2 u0 = undistort_pt(fc, 0, c0, nullptr), u1 = undistort_pt(fc, 1, c1, nullptr);
        mskf_feature f;
        f.id = (uint32_t)id;
        f.pad = 0;
        f.u0 = (double)u0.x; f.v0 = (double)u0.y; f.u1 = (double)u1.x; f.v1 = (double)u1.y;
        fb.msg[go + d] = f;
        fb.stale[go + d] = f;
    }
    if (threadIdx.x == 0) {
        int nout = s_m;
        if (fc.compat_stale) {
            if (nout > fb.stale_hw[s]) fb.stale_hw[s] = nout;
            fb.msg_total[s] += nout;
        } else {
            fb.stale_hw[s] = nout;
            fb.msg_total[s] = nout;
        }
        fb.info[s].time_stamp = st.t;
        // rotate prev/curr: the finished grid already sits in buffer gp, so "prev" stays gp
    }
}

}  // namespace mskf

using namespace mskf;

// Plan of pyr_tail_kernel for a geometry: first level it produces (0: not applicable), output rows per chunk of that
// level, offset of the resident area, dynamic shared memory.
struct PyrTailPlan {
    int l0, chunk_out_rows;
    unsigned area_b_off;
    size_t smem;
};
static PyrTailPlan pyr_tail_plan(const FeConst &fc) {
    PyrTailPlan pl = {0, 0, 0, 0};
    auto stride = [](int cols) { return (PS_PAD + cols + 8 + 15) & ~15; };
    if (fc.levels < 3 || (fc.pitch % 4) != 0) return pl;
    if (fc.lvl_rows[fc.levels - 1] < 4 || fc.lvl_cols[fc.levels - 1] < 4) return pl;
    for (int l0 = 2; l0 < fc.levels; ++l0) {
        if (fc.levels - l0 < 2 && l0 > 2) break;  // a single remaining level gains nothing over the strip kernel
        const int icols = fc.lvl_cols[l0 - 1];
        if ((icols % 4) != 0) continue;
        const size_t resident = (size_t)fc.lvl_rows[l0] * stride(fc.lvl_cols[l0]);
        if (resident > 64 * 1024) continue;
        const int rs_in = stride(icols), orows = fc.lvl_rows[l0];
        int nchunks = 1;
        while (nchunks < orows && (size_t)(2 * ((orows + nchunks - 1) / nchunks) + 6) * rs_in > 48 * 1024) ++nchunks;
        const int c = (orows + nchunks - 1) / nchunks;
        size_t area_a = (size_t)(2 * c + 6) * rs_in;
        for (int l = l0 + 1; l < fc.levels; l += 2)  // levels l0 + 1, l0 + 3, ... live in the chunk area
            area_a = std::max(area_a, (size_t)fc.lvl_rows[l] * stride(fc.lvl_cols[l]));
        size_t area_b = resident;
        for (int l = l0 + 2; l < fc.levels; l += 2) area_b = std::max(area_b, (size_t)fc.lvl_rows[l] * stride(fc.lvl_cols[l]));
        area_a = (area_a + 15) & ~(size_t)15;
        pl.l0 = l0;
        pl.chunk_out_rows = c;
        pl.area_b_off = (unsigned)area_a;
        pl.smem = area_a + area_b;
        return pl;
    }
    return pl;
}

static int fe_members_cap(const FeConst &fc) { return (fc.grid_w / fc.det_cell_w + 2) * (fc.grid_h / fc.det_cell_h + 2); }
static size_t fe_sieve_smem(const FeConst &fc) {
    return (size_t)fc.det_cells * 9 + 16 + sizeof(int) * ((size_t)fc.n_cells * fc.grid_max + (FE_THREADS / 32) * fe_members_cap(fc));
}
static size_t fe_after_stereo_smem(const FeConst &fc) {
    size_t b = ((size_t)fc.max_f + 15) & ~(size_t)15;
    if (fc.use_ransac)
        b += (size_t)fc.max_f * (4 + 8 + 8 + 24 + 8 + 8 + 4 + 3) + 4 * RANSAC_ITERS * (((size_t)fc.max_f + 31) / 32) + 64;
    return b + 16;
}
static size_t fe_finish_smem(const FeConst &fc) {
    return ((size_t)fc.cap_k + (size_t)fc.n_cells * fc.grid_min + (size_t)fc.n_cells_all * fc.grid_max) * sizeof(int) +
           (size_t)fc.cap_k * 5 + (size_t)fc.max_f * 5 + 16;
}

int fe_create(mskf_handle *h) {
    const mskf_config &c = h->cfg;
    FeConst &fc = h->fc;
    memset(&fc, 0, sizeof(fc));
    if (c.pyramid_levels < 1 || c.pyramid_levels > MSKF_MAX_LEVELS || (c.klt_win & 1) == 0 || c.klt_win < 3 ||
        c.klt_win > 29) {
        h->err = "bad pyramid_levels / klt_win (odd, 3..29)";
        return MSKF_ERR_ARG;
    }
    fc.rows = c.img_rows; fc.cols = c.img_cols; fc.levels = c.pyramid_levels;
    unsigned off = 0;
    int r = c.img_rows, q = c.img_cols;
    for (int l = 0; l < fc.levels; ++l) {
        fc.lvl_rows[l] = r; fc.lvl_cols[l] = q; fc.lvl_off[l] = off;
        off += (unsigned)(r * c.img_cols);  // every level keeps the pitch of level 0
        r = (r + 1) / 2; q = (q + 1) / 2;
    }
    fc.pitch = c.img_cols;
    fc.pyr_bytes = (off + 255u) & ~255u;
    fc.klt_win = c.klt_win; fc.klt_max_iters = c.klt_max_iters;
    fc.klt_eps2 = c.klt_eps * c.klt_eps; fc.klt_min_eig = c.klt_min_eig;
    fc.grid_row = c.grid_row; fc.grid_col = c.grid_col; fc.grid_min = c.grid_min_feature_num;
    fc.grid_max = c.grid_max_feature_num;
    fc.grid_h = c.img_rows / c.grid_row; fc.grid_w = c.img_cols / c.grid_col;
    fc.n_cells = c.grid_row * c.grid_col;
    fc.n_cells_all = ((c.img_rows - 1) / fc.grid_h) * c.grid_col + (c.img_cols - 1) / fc.grid_w + 1;
    if (fc.n_cells_all < fc.n_cells) fc.n_cells_all = fc.n_cells;
    if (fc.n_cells_all > FE_MAX_CELLS || fc.grid_min > fc.grid_max) {
        h->err = "grid too large (max 128 cells incl. overflow) or grid_min > grid_max";
        return MSKF_ERR_ARG;
    }
    fc.det_rows = c.det_rows; fc.det_cols = c.det_cols;
    fc.det_cell_h = c.img_rows / c.det_rows + 1; fc.det_cell_w = c.img_cols / c.det_cols + 1;
    fc.det_cells = c.det_rows * c.det_cols;
    fc.fast_threshold = c.fast_threshold;
    fc.detection_threshold = c.detection_threshold;
    fc.max_f = ((fc.n_cells_all * (fc.grid_max + fc.grid_min) + 31) / 32) * 32;
    fc.cap_k = fc.det_cells > fc.max_f ? fc.det_cells : fc.max_f;
    fc.cam_model[0] = c.cam0_model; fc.cam_model[1] = c.cam1_model;
    for (int i = 0; i < 4; ++i) {
        fc.K[0][i] = c.cam0_intrinsics[i]; fc.D[0][i] = c.cam0_distortion[i];
        fc.K[1][i] = c.cam1_intrinsics[i]; fc.D[1][i] = c.cam1_distortion[i];
    }
    fc.compat_stale = c.compat_stale_features;
    fc.use_ransac = c.use_ransac;  // the reference never calls twoPointRansac (image_processor.cpp:482-493): default 0
    fc.ransac_threshold = c.ransac_threshold;
    // extrinsics as ImageProcessor::loadParameters derives them (image_processor.cpp:63-72)
    auto R = [](const double *T, int i, int j) { return T[i * 4 + j]; };
    double R_c0_imu[9], t_c0_imu[3], T1[16], R_c1_imu[9], t_c1_imu[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R_c0_imu[i * 3 + j] = R(c.T_cam0_imu, j, i);
    for (int i = 0; i < 3; ++i) {
        double sacc = R_c0_imu[i * 3 + 0] * c.T_cam0_imu[3] + R_c0_imu[i * 3 + 1] * c.T_cam0_imu[7] + R_c0_imu[i * 3 + 2] * c.T_cam0_imu[11];
        t_c0_imu[i] = -sacc;
    }
    // T_cam1_imu = T_cn_cnm1 * T_cam0_imu  (R = Ra Rb, t = Ra tb + ta)
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            double sacc = 0;
            for (int k = 0; k < 3; ++k) sacc += R(c.T_cn_cnm1, i, k) * R(c.T_cam0_imu, k, j);
            T1[i * 4 + j] = sacc;
        }
        double sacc = R(c.T_cn_cnm1, i, 0) * c.T_cam0_imu[3] + R(c.T_cn_cnm1, i, 1) * c.T_cam0_imu[7] + R(c.T_cn_cnm1, i, 2) * c.T_cam0_imu[11];
        T1[i * 4 + 3] = sacc + c.T_cn_cnm1[i * 4 + 3];
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R_c1_imu[i * 3 + j] = T1[j * 4 + i];
    for (int i = 0; i < 3; ++i) {
        double sacc = R_c1_imu[i * 3 + 0] * T1[3] + R_c1_imu[i * 3 + 1] * T1[7] + R_c1_imu[i * 3 + 2] * T1[11];
        t_c1_imu[i] = -sacc;
    }
    // R_cam0_cam1 = R_cam1_imu^T * R_cam0_imu ; t_cam0_cam1 = R_cam1_imu^T (t_cam0_imu - t_cam1_imu)
    double tdiff[3] = {t_c0_imu[0] - t_c1_imu[0], t_c0_imu[1] - t_c1_imu[1], t_c0_imu[2] - t_c1_imu[2]};
    double t01[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            double sacc = 0;
            for (int k = 0; k < 3; ++k) sacc += R_c1_imu[k * 3 + i] * R_c0_imu[k * 3 + j];
            fc.R01[i * 3 + j] = sacc;
        }
        t01[i] = R_c1_imu[0 * 3 + i] * tdiff[0] + R_c1_imu[1 * 3 + i] * tdiff[1] + R_c1_imu[2 * 3 + i] * tdiff[2];
    }
    double sk[9] = {0, -t01[2], t01[1], t01[2], 0, -t01[0], -t01[1], t01[0], 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double sacc = 0;
            for (int k = 0; k < 3; ++k) sacc += sk[i * 3 + k] * fc.R01[k * 3 + j];
            fc.E[i * 3 + j] = sacc;
        }
    double npu = 4.0 / (c.cam0_intrinsics[0] + c.cam0_intrinsics[1] + c.cam1_intrinsics[0] + c.cam1_intrinsics[1]);
    fc.stereo_gate = c.stereo_threshold * npu;

    FeBuffers &fb = h->fb;
    memset(&fb, 0, sizeof(fb));
    const size_t S = h->S;
    int rc;
#define A(p, n) if ((rc = dev_alloc(h, &(p), (n))) != MSKF_OK) return rc
    for (int i = 0; i < 3; ++i) A(fb.pyr[i], S * fc.pyr_bytes);
    A(fb.staging, 2 * S * 2 * (size_t)fc.rows * fc.cols);
    A(fb.src0, S); A(fb.src1, S);
    A(fb.step, S);
    for (int g = 0; g < 2; ++g) {
        A(fb.g_id[g], S * fc.max_f); A(fb.g_resp[g], S * fc.max_f); A(fb.g_life[g], S * fc.max_f);
        A(fb.g_cam0[g], S * fc.max_f); A(fb.g_cam1[g], S * fc.max_f); A(fb.g_cell[g], S * fc.max_f);
        A(fb.g_n[g], S);
    }
    A(fb.gslot, S); A(fb.next_id, S);
    A(fb.k_a, S * fc.cap_k); A(fb.k_b, S * fc.cap_k); A(fb.k_status, S * fc.cap_k); A(fb.k_skip, S * fc.cap_k); A(fb.k_n, S); A(fb.k_idx, S * fc.cap_k); A(fb.k_nm, S);
    A(fb.t_id, S * fc.max_f); A(fb.t_life, S * fc.max_f); A(fb.t_p0, S * fc.max_f); A(fb.t_p1, S * fc.max_f); A(fb.track_calls, S);
    A(fb.det_best, S * fc.det_cells); A(fb.det_occ, S * fc.det_cells);
    A(fb.nf_resp, S * fc.det_cells); A(fb.nf_n, S); A(fb.in_resp, S * fc.cap_k);
    A(fb.msg, S * fc.max_f); A(fb.msg_n, S); A(fb.stale, S * fc.max_f); A(fb.stale_hw, S); A(fb.msg_total, S);
    A(fb.info, S);
#undef A
    fb.work = h->d_work;
    if (fe_sieve_smem(fc) > 200 * 1024 || fe_finish_smem(fc) > 200 * 1024) {
        h->err = "detector grid / feature capacity too large for the bookkeeping kernels' shared memory";
        return MSKF_ERR_ARG;
    }
    if (fe_after_stereo_smem(fc) > 200 * 1024) {
        h->err = "feature capacity too large for twoPointRansac's shared memory";
        return MSKF_ERR_ARG;
    }
    if ((rc = smem_optin(h, fe_after_stereo, fe_after_stereo_smem(fc))) != MSKF_OK) return rc;
    if ((rc = smem_optin(h, fe_sieve, fe_sieve_smem(fc))) != MSKF_OK) return rc;
    if ((rc = smem_optin(h, fe_finish, fe_finish_smem(fc))) != MSKF_OK) return rc;
    // pyramid strip kernels (launch_pyr_level): the widest strip is level 0's
    if ((rc = smem_optin(h, pyr_down_bulk_kernel<true>, (size_t)PS_IN * fc.lvl_cols[0] + 32)) != MSKF_OK) return rc;
    if ((rc = smem_optin(h, pyr_down_strip_kernel<true, 4>, 0)) != MSKF_OK) return rc;
    if ((rc = smem_optin(h, pyr_down_strip_kernel<false, 4>, 0)) != MSKF_OK) return rc;
    if ((rc = smem_optin(h, pyr_tail_kernel, pyr_tail_plan(fc).smem)) != MSKF_OK) return rc;
    return MSKF_OK;
}

static void launch_klt(mskf_handle *h, int tag, dim3 g, size_t smem, int mode) {
    const FeConst &fc = h->fc;
    const FeBuffers &fb = h->fb;
    cudaStream_t q = h->stream;
    // standard geometries (EuRoC 752 wide, the stress preset's 1280) get their pitch as a compile-time constant
    if (fc.klt_win == 21 && fc.pitch == 752) MSKF_LAUNCH(h, tag, (klt_reg_kernel<21, 752><<<g, KLT_WARPS * 32, 0, q>>>(fc, fb, mode)));
    else if (fc.klt_win == 15 && fc.pitch == 752) MSKF_LAUNCH(h, tag, (klt_reg_kernel<15, 752><<<g, KLT_WARPS * 32, 0, q>>>(fc, fb, mode)));
    else if (fc.klt_win == 21 && fc.pitch == 1280) MSKF_LAUNCH(h, tag, (klt_reg_kernel<21, 1280><<<g, KLT_WARPS * 32, 0, q>>>(fc, fb, mode)));
    else if (fc.klt_win == 21) MSKF_LAUNCH(h, tag, (klt_reg_kernel<21, 0><<<g, KLT_WARPS * 32, 0, q>>>(fc, fb, mode)));
    else if (fc.klt_win == 15) MSKF_LAUNCH(h, tag, (klt_reg_kernel<15, 0><<<g, KLT_WARPS * 32, 0, q>>>(fc, fb, mode)));
    else MSKF_LAUNCH(h, tag, (klt_kernel<0><<<g, KLT_WARPS * 32, smem, q>>>(fc, fb, mode)));
}

// pyramid level l for `images` (= 2 S) images: vectorised strip kernel when the geometry allows it
static void launch_pyr_level(mskf_handle *h, int l, int images) {
    const FeConst &fc = h->fc;
    const FeBuffers &fb = h->fb;
    cudaStream_t q = h->stream;
    const int icols = fc.lvl_cols[l - 1];
    const int row_stride = (PS_PAD + icols + 8 + 15) & ~15;
    const size_t smem = (size_t)PS_IN * row_stride;
    const int tag = l == 1 ? PK_PYR_L1 : PK_PYR_LN;
    if (smem <= 200 * 1024 && (icols % 4) == 0 && icols >= 8 && (fc.pitch % 4) == 0) {
        dim3 g((fc.lvl_rows[l] + PS_ROWS - 1) / PS_ROWS, 1, images);
        if (l == 1 && (icols % 16) == 0) {
            const size_t bsmem = (size_t)PS_IN * icols + 32;
            MSKF_LAUNCH(h, tag, (pyr_down_bulk_kernel<true><<<g, 256, bsmem, q>>>(fc, fb, l)));
        } else if (l == 1) {
            MSKF_LAUNCH(h, tag, (pyr_down_strip_kernel<true, 4><<<g, 256, smem, q>>>(fc, fb, l, row_stride)));
        } else {
            MSKF_LAUNCH(h, tag, (pyr_down_strip_kernel<false, 4><<<g, 256, smem, q>>>(fc, fb, l, row_stride)));
        }
    } else {
        dim3 g((fc.lvl_cols[l] + PD_TW - 1) / PD_TW, (fc.lvl_rows[l] + PD_TH - 1) / PD_TH, images);
        if (l == 1) MSKF_LAUNCH(h, tag, (pyr_down_kernel<true><<<g, 256, 0, q>>>(fc, fb, l)));
        else MSKF_LAUNCH(h, tag, (pyr_down_kernel<false><<<g, 256, 0, q>>>(fc, fb, l)));
    }
}

int fe_step(mskf_handle *h, bool any_first, int max_prev, int n_active) {
    h->cur = h->stream;
    const FeConst &fc = h->fc;
    const FeBuffers &fb = h->fb;
    cudaStream_t q = h->stream;
    const int S = h->S;
    if (fc.levels == 1) {
        h->err = "pyramid_levels must be >= 2";
        return MSKF_ERR_ARG;
    }
    // pyramids
    const PyrTailPlan tail = pyr_tail_plan(fc);
    for (int l = 1; l < fc.levels; ++l) {
        if (tail.l0 && l == tail.l0) {
            MSKF_LAUNCH(h, PK_PYR_LN, (pyr_tail_kernel<<<S * 2, PT_THREADS, tail.smem, q>>>(fc, fb, tail.l0, tail.chunk_out_rows, tail.area_b_off)));
        } else if (!tail.l0 || l < tail.l0) {
            launch_pyr_level(h, l, S * 2);
        }
        if (l == 1) {
            int rc = stage_end_consume(h);
            if (rc != MSKF_OK) return rc;
        }
        // algorithmic bytes: read level l-1, write level l (+ the level-0 landing copy at l == 1)
        double in = (double)fc.lvl_rows[l - 1] * fc.lvl_cols[l - 1], out = (double)fc.lvl_rows[l] * fc.lvl_cols[l];
        h->work_host[l == 1 ? PK_PYR_L1 : PK_PYR_LN] += 2.0 * n_active * (in + out + (l == 1 ? in : 0.0));
    }
    h->work_host[PK_DETECT] += (double)n_active * fc.rows * fc.cols;
    const size_t klt_smem = (size_t)KLT_WARPS * ((((fc.klt_win + 2) * (fc.klt_win + 2) + 2 * fc.klt_win * fc.klt_win + 3) & ~3)) * sizeof(short) + 16;
    const size_t pos_smem = (size_t)fc.cap_k * sizeof(int);
    MSKF_LAUNCH(h, PK_FE_BOOK, (fe_prep_track<<<S, FE_THREADS, 0, q>>>(fc, fb)));
    if (max_prev > 0) {
        dim3 g((max_prev + KLT_WARPS - 1) / KLT_WARPS, S);
        launch_klt(h, PK_KLT_TEMPORAL, g, klt_smem, 0);
        MSKF_LAUNCH(h, PK_FE_BOOK, (fe_after_track<<<S, FE_THREADS, pos_smem, q>>>(fc, fb)));
        launch_klt(h, PK_KLT_STEREO, g, klt_smem, 1);
    }
    MSKF_LAUNCH(h, PK_FE_BOOK, (fe_after_stereo<<<S, FE_THREADS, fe_after_stereo_smem(fc), q>>>(fc, fb)));
    {
        dim3 g((fc.cols + DT_W - 1) / DT_W, (fc.rows + DT_H - 1) / DT_H, S);
        MSKF_LAUNCH(h, PK_DETECT, (detect_kernel<<<g, 256, 0, q>>>(fc, fb)));
    }
    MSKF_LAUNCH(h, PK_FE_BOOK, (fe_sieve<<<S, FE_THREADS, fe_sieve_smem(fc), q>>>(fc, fb, fe_members_cap(fc))));
    {
        int cap = any_first ? fc.det_cells : fc.n_cells * fc.grid_max;
        dim3 g((cap + KLT_WARPS - 1) / KLT_WARPS, S);
        // the register kernels stride over the compacted candidates: a quarter of the list's warps is enough in
        // steady state (the first frame matches every detection and keeps the full grid)
        if (!any_first && (fc.klt_win == 21 || fc.klt_win == 15)) g.x = (g.x + 3) / 4;
        launch_klt(h, PK_KLT_NEW, g, klt_smem, 2);
    }
    const size_t fin_smem = fe_finish_smem(fc);
    // fe_finish rewrites the CameraMeasurement: the back end of the previous frame must have read it
    if (h->msg_consumed_valid) MSKF_CUDA_CHECK(h, cudaStreamWaitEvent(q, h->ev_msg_consumed, 0));
    MSKF_LAUNCH(h, PK_FE_BOOK, (fe_finish<<<S, FE_THREADS, fin_smem, q>>>(fc, fb)));
    MSKF_CUDA_CHECK(h, cudaEventRecord(h->ev_msg_ready, q));
    MSKF_CUDA_CHECK(h, cudaGetLastError());
    return MSKF_OK;
}

// ---- stand-alone operator helpers (mskf_op_detect / mskf_op_klt), one-stream handle ----
static int op_prepare(mskf_handle *t) {
    FeStep st;
    memset(&st, 0, sizeof(st));
    st.active = 1;
    st.is_first = 1;
    st.slot = 0;
    HostStream &hs = t->hs[0];
    {
        int rc = stage_begin_consume(t);
        if (rc != MSKF_OK) return rc;
    }
    MSKF_CUDA_CHECK(t, cudaMemcpyAsync(t->fb.step, &st, sizeof(st), cudaMemcpyHostToDevice, t->stream));
    MSKF_CUDA_CHECK(t, cudaMemcpyAsync(t->fb.src0, &hs.src0, sizeof(uint8_t *), cudaMemcpyHostToDevice, t->stream));
    MSKF_CUDA_CHECK(t, cudaMemcpyAsync(t->fb.src1, &hs.src1, sizeof(uint8_t *), cudaMemcpyHostToDevice, t->stream));
    MSKF_CUDA_CHECK(t, cudaStreamSynchronize(t->stream));
    const FeConst &fc = t->fc;
    for (int l = 1; l < fc.levels; ++l) {
        launch_pyr_level(t, l, 2);
        if (l == 1) {
            int rc = stage_end_consume(t);
            if (rc != MSKF_OK) return rc;
        }
    }
    MSKF_CUDA_CHECK(t, cudaGetLastError());
    hs.slot = 0;
    hs.pending = false;
    return MSKF_OK;
}

int fe_op_detect(mskf_handle *t, const float *occ, int n_occ, float *out_xy, double *out_resp, int cap, int *n,
                 uint8_t *score_map) {
    int rc = op_prepare(t);
    if (rc != MSKF_OK) return rc;
    const FeConst &fc = t->fc;
    if (score_map) {
        rc = dev_alloc(t, &t->fb.dbg_score, (size_t)fc.rows * fc.cols);
        if (rc != MSKF_OK) return rc;
    }
    std::vector<uint8_t> occv(fc.det_cells, 0);
    for (int i = 0; i < n_occ; ++i) {
        int x = (int)occ[2 * i], y = (int)occ[2 * i + 1];
        if (x < 0 || y < 0 || x >= fc.cols || y >= fc.rows) continue;
        occv[(y / fc.det_cell_h) * fc.det_cols + (x / fc.det_cell_w)] = 1;
    }
    MSKF_CUDA_CHECK(t, cudaMemcpyAsync(t->fb.det_occ, occv.data(), occv.size(), cudaMemcpyHostToDevice, t->stream));
    MSKF_CUDA_CHECK(t, cudaMemsetAsync(t->fb.det_best, 0, sizeof(unsigned long long) * fc.det_cells, t->stream));
    dim3 g((fc.cols + DT_W - 1) / DT_W, (fc.rows + DT_H - 1) / DT_H, 1);
    detect_kernel<<<g, 256, 0, t->stream>>>(fc, t->fb);
    MSKF_CUDA_CHECK(t, cudaGetLastError());
    std::vector<unsigned long long> keys(fc.det_cells);
    MSKF_CUDA_CHECK(t, cudaMemcpyAsync(keys.data(), t->fb.det_best, sizeof(unsigned long long) * fc.det_cells, cudaMemcpyDeviceToHost, t->stream));
    if (score_map)
        MSKF_CUDA_CHECK(t, cudaMemcpyAsync(score_map, t->fb.dbg_score, (size_t)fc.rows * fc.cols, cudaMemcpyDeviceToHost, t->stream));
    MSKF_CUDA_CHECK(t, cudaStreamSynchronize(t->stream));
    int cnt = 0;
    for (int k = 0; k < fc.det_cells; ++k) {
        if (!keys[k]) continue;
        unsigned bits = (unsigned)(keys[k] >> 32);
        float resp;
        memcpy(&resp, &bits, 4);
        if (!((double)resp > fc.detection_threshold)) continue;
        unsigned ridx = 0xffffffffu - (unsigned)(keys[k] & 0xffffffffull);
        if (cnt < cap) {
            out_xy[2 * cnt] = (float)(ridx % (unsigned)fc.cols);
            out_xy[2 * cnt + 1] = (float)(ridx / (unsigned)fc.cols);
            out_resp[cnt] = (double)resp;
        }
        ++cnt;
    }
    *n = cnt;
    return MSKF_OK;
}

int fe_op_klt(mskf_handle *t, const float *pts_a, float *pts_b, uint8_t *status, int n) {
    int rc = op_prepare(t);
    if (rc != MSKF_OK) return rc;
    const FeConst &fc = t->fc;
    MSKF_CUDA_CHECK(t, cudaMemcpyAsync(t->fb.k_a, pts_a, sizeof(float2) * n, cudaMemcpyHostToDevice, t->stream));
    MSKF_CUDA_CHECK(t, cudaMemcpyAsync(t->fb.k_b, pts_b, sizeof(float2) * n, cudaMemcpyHostToDevice, t->stream));
    MSKF_CUDA_CHECK(t, cudaMemcpyAsync(t->fb.k_n, &n, sizeof(int), cudaMemcpyHostToDevice, t->stream));
    const size_t klt_smem = (size_t)KLT_WARPS * ((((fc.klt_win + 2) * (fc.klt_win + 2) + 2 * fc.klt_win * fc.klt_win + 3) & ~3)) * sizeof(short) + 16;
    dim3 g((n + KLT_WARPS - 1) / KLT_WARPS, 1);
    launch_klt(t, PK_KLT_STEREO, g, klt_smem, 1);
    MSKF_CUDA_CHECK(t, cudaGetLastError());
    MSKF_CUDA_CHECK(t, cudaMemcpyAsync(pts_b, t->fb.k_b, sizeof(float2) * n, cudaMemcpyDeviceToHost, t->stream));
    MSKF_CUDA_CHECK(t, cudaMemcpyAsync(status, t->fb.k_status, n, cudaMemcpyDeviceToHost, t->stream));
    MSKF_CUDA_CHECK(t, cudaStreamSynchronize(t->stream));
    return MSKF_OK;
}
