// cubeio.cu -- the data formats either side of the hot path: spectro-polarimetric cubes
// arrive frequency-first, (freq, Y, X) Stokes Q and U planes (README.md:288-297, io/io_functions.py:58-84), and
// Faraday-depth cubes leave as (2, n_phi, Y, X) real / imaginary planes (io/io_functions.py:151-203), while the
// kernels work on pixel-major planar rows [P, 2m] / [P, 2n].  These are the HBM-bound transposes between the two,
// plus the per-channel image statistics the reference computes before building a Dataset:
//   csr_pack_slab      Q, U [m, npix] -> data [P, ld] = [Re | Im] of pixels p0 .. p0+P-1; NaN / Inf samples become 0
//                      and are reported in valid[P, m] (what the per-pixel weights must be multiplied with)
//   csr_unpack_slab    F [P, ld] planar -> re, im [n, npix] columns p0 .. p0+P-1
//   csr_channel_stats  per channel of an image cube [m, Y, X] over the window [y0, yn) x [x0, xn): nansum and nanstd
//                      (utils/utilities.py:68-110 `calculate_noise` without sigma clipping; io_functions.py:15-33
//                      `filter_cubes` drops channels whose nansum is zero)
// All three move each byte once: 32 x 32 shared-memory tiles for the transposes (coalesced on both sides), one CTA
// per channel with fp64 accumulation for the statistics.
#include <math.h>

#include "plan.cuh"

namespace csr {

constexpr int TP = 32;  // transpose tile

// grid: (ceil(P / 32), ceil(m / 32)), block (32, 8)
__global__ void pack_slab_kernel(const float* __restrict__ q, const float* __restrict__ u, long long npix, int m,
                                 long long p0, int P, float* __restrict__ out, int ld, unsigned char* __restrict__ valid) {
    __shared__ float tq[TP][TP + 1], tu[TP][TP + 1];
    const int pb = blockIdx.x * TP, cb = blockIdx.y * TP;
    for (int r = threadIdx.y; r < TP; r += 8) {  // r: channel within the tile, threadIdx.x: pixel (contiguous in q / u)
        const int c = cb + r, p = pb + threadIdx.x;
        float a = 0.f, b = 0.f;
        if (c < m && p < P) {
            a = q[(size_t)c * npix + p0 + p];
            b = u[(size_t)c * npix + p0 + p];
        }
        tq[r][threadIdx.x] = a;
        tu[r][threadIdx.x] = b;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < TP; r += 8) {  // r: pixel within the tile, threadIdx.x: channel (contiguous in out)
        const int p = pb + r, c = cb + threadIdx.x;
        if (p < P && c < m) {
            float a = tq[threadIdx.x][r], b = tu[threadIdx.x][r];
            const bool ok = isfinite(a) && isfinite(b);
            if (!ok) a = b = 0.f;
            out[(size_t)p * ld + c] = a;
            out[(size_t)p * ld + m + c] = b;
            if (valid) valid[(size_t)p * m + c] = ok ? 1 : 0;
        }
    }
}

// The same transpose for any frequency-first array the Python layer holds: `a` / `b` are the two planes (Re / Im of a
// complex (L, P) array: b = a + 1 with pix_stride = 2 and chan_stride = 2 P; the two halves of a real (2L, P) array:
// b = a + L * P), element (c, p) at plane[c * chan_stride + p * pix_stride].  No NaN handling: values pass unchanged.
__global__ void pack_planar_kernel(const float* __restrict__ a, const float* __restrict__ b, long long chan_stride,
                                   int pix_stride, int L, int P, float* __restrict__ out, int ld) {
    __shared__ float ta[TP][TP + 1], tb[TP][TP + 1];
    const int pb = blockIdx.x * TP, cb = blockIdx.y * TP;
    for (int r = threadIdx.y; r < TP; r += 8) {
        const int c = cb + r, p = pb + threadIdx.x;
        float va = 0.f, vb = 0.f;
        if (c < L && p < P) {
            va = a[(size_t)c * chan_stride + (size_t)p * pix_stride];
            if (b) vb = b[(size_t)c * chan_stride + (size_t)p * pix_stride];
        }
        ta[r][threadIdx.x] = va;
        tb[r][threadIdx.x] = vb;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < TP; r += 8) {
        const int p = pb + r, c = cb + threadIdx.x;
        if (p < P && c < L) {
            out[(size_t)p * ld + c] = ta[threadIdx.x][r];
            if (b) out[(size_t)p * ld + L + c] = tb[threadIdx.x][r];
        }
    }
}
// ... and back: planar rows [P, ld] -> the two planes
__global__ void unpack_planar_kernel(const float* __restrict__ in, int ld, int L, int P, float* __restrict__ a,
                                     float* __restrict__ b, long long chan_stride, int pix_stride) {
    __shared__ float ta[TP][TP + 1], tb[TP][TP + 1];
    const int pb = blockIdx.x * TP, jb = blockIdx.y * TP;
    for (int r = threadIdx.y; r < TP; r += 8) {
        const int p = pb + r, j = jb + threadIdx.x;
        float va = 0.f, vb = 0.f;
        if (p < P && j < L) {
            va = in[(size_t)p * ld + j];
            if (b) vb = in[(size_t)p * ld + L + j];
        }
        ta[r][threadIdx.x] = va;
        tb[r][threadIdx.x] = vb;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < TP; r += 8) {
        const int j = jb + r, p = pb + threadIdx.x;
        if (j < L && p < P) {
            a[(size_t)j * chan_stride + (size_t)p * pix_stride] = ta[threadIdx.x][r];
            if (b) b[(size_t)j * chan_stride + (size_t)p * pix_stride] = tb[threadIdx.x][r];
        }
    }
}

// grid: (ceil(P / 32), ceil(n / 32)), block (32, 8)
__global__ void unpack_slab_kernel(const float* __restrict__ in, int ld, int n, long long npix, long long p0, int P,
                                   float* __restrict__ re, float* __restrict__ im) {
    __shared__ float tr[TP][TP + 1], ti[TP][TP + 1];
    const int pb = blockIdx.x * TP, jb = blockIdx.y * TP;
    for (int r = threadIdx.y; r < TP; r += 8) {  // r: pixel, threadIdx.x: phi cell (contiguous in `in`)
        const int p = pb + r, j = jb + threadIdx.x;
        float a = 0.f, b = 0.f;
        if (p < P && j < n) {
            a = in[(size_t)p * ld + j];
            b = in[(size_t)p * ld + n + j];
        }
        tr[r][threadIdx.x] = a;
        ti[r][threadIdx.x] = b;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < TP; r += 8) {  // r: phi cell, threadIdx.x: pixel (contiguous in re / im)
        const int j = jb + r, p = pb + threadIdx.x;
        if (j < n && p < P) {
            re[(size_t)j * npix + p0 + p] = tr[threadIdx.x][r];
            im[(size_t)j * npix + p0 + p] = ti[threadIdx.x][r];
        }
    }
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += red[i];
    return r;
}

// one CTA per channel: nansum, count, then the mean-centred second pass numpy.nanstd does
__global__ void __launch_bounds__(1024)
channel_stats_kernel(const float* __restrict__ img, int Y, int X, int x0, int xn, int y0, int yn,
                     float* __restrict__ stats) {
    __shared__ double red[32];
    const float* plane = img + (size_t)blockIdx.x * Y * X;
    // a warp per image row, lanes along x: coalesced and free of index divisions
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    double sum = 0.0, cnt = 0.0;
    for (int y = y0 + warp; y < yn; y += nwarps) {
        const float* row = plane + (size_t)y * X;
        for (int x = x0 + lane; x < xn; x += 32) {
            const float v = row[x];
            if (!isnan(v)) {
                sum += v;
                cnt += 1.0;
            }
        }
    }
    sum = block_sum(sum, red);
    cnt = block_sum(cnt, red);
    const double mean = cnt > 0.0 ? sum / cnt : 0.0;
    double ss = 0.0;
    for (int y = y0 + warp; y < yn; y += nwarps) {
        const float* row = plane + (size_t)y * X;
        for (int x = x0 + lane; x < xn; x += 32) {
            const float v = row[x];
            if (!isnan(v)) {
                const double d = (double)v - mean;
                ss += d * d;
            }
        }
    }
    ss = block_sum(ss, red);
    if (threadIdx.x == 0) {
        stats[2 * blockIdx.x] = (float)sum;
        stats[2 * blockIdx.x + 1] = cnt > 0.0 ? (float)sqrt(ss / cnt) : __int_as_float(0x7fc00000);  // nanstd of all-NaN = NaN
    }
}

}  // namespace csr

using namespace csr;

extern "C" int csr_pack_slab(const float* q, const float* u, long long npix, int m, long long p0, int P, float* out,
                             int ld_out, unsigned char* valid, void* stream) {
    CSR_REQUIRE(q && u && out && m > 0 && P > 0 && p0 >= 0 && p0 + P <= npix && ld_out >= 2 * m, "csr_pack_slab: bad argument");
    dim3 grid((P + TP - 1) / TP, (m + TP - 1) / TP), block(TP, 8);
    pack_slab_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(q, u, npix, m, p0, P, out, ld_out, valid);
    CSR_LAUNCHED();
    return CSR_OK;
}

extern "C" int csr_unpack_slab(const float* in, int ld, int n, long long npix, long long p0, int P, float* re, float* im,
                               void* stream) {
    CSR_REQUIRE(in && re && im && n > 0 && P > 0 && p0 >= 0 && p0 + P <= npix && ld >= 2 * n, "csr_unpack_slab: bad argument");
    dim3 grid((P + TP - 1) / TP, (n + TP - 1) / TP), block(TP, 8);
    unpack_slab_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, ld, n, npix, p0, P, re, im);
    CSR_LAUNCHED();
    return CSR_OK;
}

extern "C" int csr_channel_stats(const float* img, int m, int Y, int X, int x0, int xn, int y0, int yn, float* stats,
                                 void* stream) {
    CSR_REQUIRE(img && stats && m > 0 && Y > 0 && X > 0, "csr_channel_stats: bad argument");
    CSR_REQUIRE(0 <= x0 && x0 < xn && xn <= X && 0 <= y0 && y0 < yn && yn <= Y, "csr_channel_stats: bad window");
    channel_stats_kernel<<<m, 1024, 0, (cudaStream_t)stream>>>(img, Y, X, x0, xn, y0, yn, stats);
    CSR_LAUNCHED();
    return CSR_OK;
}

extern "C" int csr_pack_planar(const float* a, const float* b, long long chan_stride, int pix_stride, int L, int P, float* out,
                               int ld_out, void* stream) {
    CSR_REQUIRE(a && out && L > 0 && P > 0 && pix_stride > 0 && ld_out >= (b ? 2 * L : L), "csr_pack_planar: bad argument");
    dim3 grid((P + TP - 1) / TP, (L + TP - 1) / TP), block(TP, 8);
    pack_planar_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(a, b, chan_stride, pix_stride, L, P, out, ld_out);
    CSR_LAUNCHED();
    return CSR_OK;
}

extern "C" int csr_unpack_planar(const float* in, int ld, int L, int P, float* a, float* b, long long chan_stride, int pix_stride,
                                 void* stream) {
    CSR_REQUIRE(in && a && L > 0 && P > 0 && pix_stride > 0 && ld >= (b ? 2 * L : L), "csr_unpack_planar: bad argument");
    dim3 grid((P + TP - 1) / TP, (L + TP - 1) / TP), block(TP, 8);
    unpack_planar_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(in, ld, L, P, a, b, chan_stride, pix_stride);
    CSR_LAUNCHED();
    return CSR_OK;
}
