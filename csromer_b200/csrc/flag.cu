// flag.cu -- per-pixel channel flagging on the device.
//
// The reference's flaggers look at ONE line of sight's sigma array on the host (transformers/flaggers/mean_flagger.py:10-46,
// hampel_flagger.py:9-60) and zero the weights of the outlier channels; a cube needs that decision for every pixel (the
// README recipe builds one Dataset per pixel, README.md:327-332).  Here sigma is a [P, m] device matrix and the weights
// [P, m] are zeroed in place -- exactly the per-pixel masks the hot path consumes (GemmEpilogue::chan_scale with
// chan_per_pixel, csr_fista_args::w_per_pixel).  One CTA per line of sight; statistics in fp64 like NumPy's.
//   csr_flag_mean    outlier <=> sigma > center + nsigma * std(sigma) / sqrt(m), center = mean(sigma) or a given value
//                    (numpy.std: population standard deviation)
//   csr_flag_hampel  smooth = moving average over `window` channels ("valid" part, numpy.convolve), center = median(smooth),
//                    spread = 1.4826 * median |smooth - center|; outlier <=> |sigma - center| > nsigma * spread.
//                    The medians come from a bitonic sort of the row in shared memory (m <= 8192).
#include <math.h>

#include "plan.cuh"

namespace csr {

constexpr int FLAG_THREADS = 256;

__device__ __forceinline__ double flag_block_sum(double v, double* red) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double r = 0.0;
    for (int i = 0; i < FLAG_THREADS / 32; ++i) r += red[i];
    return r;
}

__global__ void __launch_bounds__(FLAG_THREADS)
flag_mean_kernel(const float* __restrict__ sigma, int m, const float* __restrict__ center_in, float nsigma, float* __restrict__ w,
                 int* __restrict__ flagged) {
    __shared__ double red[FLAG_THREADS / 32];
    const int p = blockIdx.x;
    const float* row = sigma + (size_t)p * m;
    double sum = 0.0;
    for (int i = threadIdx.x; i < m; i += FLAG_THREADS) sum += (double)row[i];
    const double mean = flag_block_sum(sum, red) / (double)m;
    double ss = 0.0;
    for (int i = threadIdx.x; i < m; i += FLAG_THREADS) {
        const double d = (double)row[i] - mean;
        ss += d * d;
    }
    const double sd = sqrt(flag_block_sum(ss, red) / (double)m);
    const double thr = (center_in ? (double)center_in[p] : mean) + (double)nsigma * sd / sqrt((double)m);
    int n = 0;
    for (int i = threadIdx.x; i < m; i += FLAG_THREADS)
        if ((double)row[i] > thr) {
            w[(size_t)p * m + i] = 0.f;
            ++n;
        }
    if (flagged) {
        const double tot = flag_block_sum((double)n, red);
        if (threadIdx.x == 0) flagged[p] = (int)tot;
    }
}

// ascending bitonic sort of v[0 .. n2) in shared memory (n2 a power of two; padding = +inf)
__device__ __forceinline__ void bitonic_sort(float* v, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += FLAG_THREADS) {
                const int q = i ^ j;
                if (q > i) {
                    const bool up = (i & k) == 0;
                    const float a = v[i], b = v[q];
                    if ((a > b) == up) {
                        v[i] = b;
                        v[q] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}
__device__ __forceinline__ double median_sorted(const float* v, int n) {  // numpy.median of the n smallest entries
    return (n & 1) ? (double)v[n >> 1] : 0.5 * ((double)v[(n >> 1) - 1] + (double)v[n >> 1]);
}

__global__ void __launch_bounds__(FLAG_THREADS)
flag_hampel_kernel(const float* __restrict__ sigma, int m, int window, float nsigma, int n2, float* __restrict__ w,
                   int* __restrict__ flagged) {
    extern __shared__ float hs[];  // [n2] sort buffer
    __shared__ double red[FLAG_THREADS / 32];
    const int p = blockIdx.x;
    const float* row = sigma + (size_t)p * m;
    const int ns = m - window + 1;  // length of the "valid" moving average
    const float inf = __int_as_float(0x7f800000);
    for (int i = threadIdx.x; i < n2; i += FLAG_THREADS) {
        float v = inf;
        if (i < ns) {
            double acc = 0.0;
            for (int t = 0; t < window; ++t) acc += (double)row[i + t];
            v = (float)(acc / (double)window);
        }
        hs[i] = v;
    }
    __syncthreads();
    bitonic_sort(hs, n2);
    const double center = median_sorted(hs, ns);
    __syncthreads();
    for (int i = threadIdx.x; i < n2; i += FLAG_THREADS)
        if (i < ns) hs[i] = (float)fabs((double)hs[i] - center);
    __syncthreads();
    bitonic_sort(hs, n2);
    const double spread = 1.4826 * median_sorted(hs, ns);
    int n = 0;
    for (int i = threadIdx.x; i < m; i += FLAG_THREADS)
        if (fabs((double)row[i] - center) > (double)nsigma * spread) {
            w[(size_t)p * m + i] = 0.f;
            ++n;
        }
    if (flagged) {
        const double tot = flag_block_sum((double)n, red);
        if (threadIdx.x == 0) flagged[p] = (int)tot;
    }
}

}  // namespace csr

using namespace csr;

extern "C" int csr_flag_mean(const float* sigma, int P, int m, const float* mean_sigma, float nsigma, float* w, int* flagged,
                             void* stream) {
    CSR_REQUIRE(sigma && w && P > 0 && m > 0, "csr_flag_mean: bad argument");
    flag_mean_kernel<<<P, FLAG_THREADS, 0, (cudaStream_t)stream>>>(sigma, m, mean_sigma, nsigma, w, flagged);
    CSR_LAUNCHED();
    return CSR_OK;
}

extern "C" int csr_flag_hampel(const float* sigma, int P, int m, int window, float nsigma, float* w, int* flagged, void* stream) {
    CSR_REQUIRE(sigma && w && P > 0 && m > 0 && window >= 1 && window <= m, "csr_flag_hampel: bad argument");
    int n2 = 1;
    while (n2 < m - window + 1) n2 <<= 1;
    CSR_REQUIRE(n2 <= 8192, "csr_flag_hampel: at most 8192 channels per line of sight (m=%d)", m);
    flag_hampel_kernel<<<P, FLAG_THREADS, n2 * sizeof(float), (cudaStream_t)stream>>>(sigma, m, window, nsigma, n2, w, flagged);
    CSR_LAUNCHED();
    return CSR_OK;
}
