// restore.cu -- the steps that follow FISTA in the cube recipe, all HBM-bound,
// one CTA per line of sight:
//   * FDF-edge noise estimate            README.md:336-338        noise = eta/2 (std Re F[edges] + std Im F[edges])
//   * restore beam                       reconstruction/parameter.py:139-169   Gaussian "same" convolution of Re and Im
//   * restored = beam * model + residual wrappers/reconstructors/csromer_reconstructor.py:213, README.md:375
//   * peak of |F| with 3-point parabola  csromer_reconstructor.py:59-84 (RM at peak, quadratic interpolation)
#include <string.h>

#include "plan.cuh"

namespace csr {

constexpr int RS_THREADS = 256;
constexpr int MAX_TAPS = 255;

struct BeamTaps {
    float t[MAX_TAPS + 1];
    int count;
};

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double red[RS_THREADS / 32];
    __shared__ double total;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < RS_THREADS / 32; ++i) t += red[i];
        total = t;
    }
    __syncthreads();
    return total;
}

// numpy.std (ddof = 0) of the real and of the imaginary part over the masked phi cells, two passes in fp64
__global__ void __launch_bounds__(RS_THREADS)
edge_noise_kernel(const float* __restrict__ F, int ld, int n, const unsigned char* __restrict__ mask, float eta,
                  float* __restrict__ noise) {
    const int p = blockIdx.x;
    const float* re = F + (size_t)p * ld;
    const float* im = re + n;
    double sr = 0.0, si = 0.0, cnt = 0.0;
    for (int j = threadIdx.x; j < n; j += RS_THREADS)
        if (mask[j]) {
            sr += re[j];
            si += im[j];
            cnt += 1.0;
        }
    sr = block_sum(sr);
    si = block_sum(si);
    cnt = block_sum(cnt);
    const double mr = cnt > 0 ? sr / cnt : 0.0, mi = cnt > 0 ? si / cnt : 0.0;
    double vr = 0.0, vi = 0.0;
    for (int j = threadIdx.x; j < n; j += RS_THREADS)
        if (mask[j]) {
            const double dr = re[j] - mr, di = im[j] - mi;
            vr += dr * dr;
            vi += di * di;
        }
    vr = block_sum(vr);
    vi = block_sum(vi);
    if (threadIdx.x == 0) noise[p] = cnt > 0 ? (float)(eta * 0.5 * (sqrt(vr / cnt) + sqrt(vi / cnt))) : 0.f;
}

// out = conv_same(x, taps) (+ add): scipy.signal.convolve(x, k, mode="same") of the real and the imaginary half
// separately -- zero beyond the ends, output sample j = sum_t k[t] x[j + c - t] with c = (len(k) - 1) / 2
__global__ void __launch_bounds__(RS_THREADS)
convolve_phi_kernel(const float* __restrict__ x, int ld, int n, const BeamTaps taps, const float* __restrict__ add,
                    float* __restrict__ out) {
    extern __shared__ float row[];  // [2n] the line of sight
    const int p = blockIdx.x;
    const float* src = x + (size_t)p * ld;
    for (int j = threadIdx.x; j < 2 * n; j += RS_THREADS) row[j] = src[j];
    __syncthreads();
    const int c = (taps.count - 1) / 2;
    for (int j = threadIdx.x; j < 2 * n; j += RS_THREADS) {
        const int half = j >= n ? n : 0;  // stay inside the Re or the Im half
        const int jj = j - half;
        float acc = 0.f;
        for (int t = 0; t < taps.count; ++t) {
            const int s = jj + c - t;
            if (s >= 0 && s < n) acc = fmaf(taps.t[t], row[half + s], acc);
        }
        if (add) acc += add[(size_t)p * ld + j];
        out[(size_t)p * ld + j] = acc;
    }
}

// per line of sight: index and value of max |F|, and the 3-point parabola through |F| around it
// (csromer_reconstructor.py:59-84: offset = 0.5 (a - c) / (a - 2 b + c), peak = b - 0.25 (a - c) offset)
__global__ void __launch_bounds__(RS_THREADS)
peak_kernel(const float* __restrict__ F, int ld, int n, int* __restrict__ argmax, float* __restrict__ stats) {
    __shared__ float bestv[RS_THREADS];
    __shared__ int besti[RS_THREADS];
    const int p = blockIdx.x;
    const float* re = F + (size_t)p * ld;
    const float* im = re + n;
    float bv = -1.f;
    int bi = 0;
    for (int j = threadIdx.x; j < n; j += RS_THREADS) {
        const float a = hypotf(re[j], im[j]);
        if (a > bv) {  // first maximum wins within a thread (j ascending)
            bv = a;
            bi = j;
        }
    }
    bestv[threadIdx.x] = bv;
    besti[threadIdx.x] = bi;
    __syncthreads();
    for (int s = RS_THREADS / 2; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const float ov = bestv[threadIdx.x + s];
            const int oi = besti[threadIdx.x + s];
            if (ov > bestv[threadIdx.x] || (ov == bestv[threadIdx.x] && oi < besti[threadIdx.x])) {
                bestv[threadIdx.x] = ov;
                besti[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const int j = besti[0];
        const float b = bestv[0];
        float off = 0.f, pk = b;
        if (j > 0 && j < n - 1) {
            const float a = hypotf(re[j - 1], im[j - 1]), c = hypotf(re[j + 1], im[j + 1]);
            const float den = a - 2.f * b + c;
            if (den != 0.f) {
                off = 0.5f * (a - c) / den;
                pk = b - 0.25f * (a - c) * off;
            }
        }
        argmax[p] = j;
        stats[3 * p] = b;
        stats[3 * p + 1] = off;
        stats[3 * p + 2] = pk;
    }
}

}  // namespace csr

using namespace csr;

extern "C" int csr_edge_noise(const float* F, int ld, int n, int P, const unsigned char* edge_mask, float eta,
                              float* noise, void* stream) {
    CSR_REQUIRE(F && edge_mask && noise && P > 0 && n > 0 && ld >= 2 * n, "csr_edge_noise: bad argument");
    edge_noise_kernel<<<P, RS_THREADS, 0, (cudaStream_t)stream>>>(F, ld, n, edge_mask, eta, noise);
    CSR_LAUNCHED();
    return CSR_OK;
}

extern "C" int csr_convolve_phi(const float* x, int ld, int n, int P, const float* taps_host, int ntaps, const float* add,
                                float* out, void* stream) {
    CSR_REQUIRE(x && out && taps_host && P > 0 && n > 0 && ld >= 2 * n, "csr_convolve_phi: bad argument");
    CSR_REQUIRE(ntaps >= 1 && ntaps <= MAX_TAPS && (ntaps & 1), "csr_convolve_phi: the beam needs an odd tap count <= %d (got %d)",
                MAX_TAPS, ntaps);
    CSR_REQUIRE((size_t)2 * n * sizeof(float) <= 200 * 1024, "csr_convolve_phi: line of sight too long for shared memory (n=%d)", n);
    BeamTaps taps;
    memset(&taps, 0, sizeof(taps));
    for (int t = 0; t < ntaps; ++t) taps.t[t] = taps_host[t];
    taps.count = ntaps;
    const int smem = 2 * n * (int)sizeof(float);
    static int configured = 0;
    if (smem > configured) {
        CSR_CUDA(cudaFuncSetAttribute(convolve_phi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = smem;
    }
    convolve_phi_kernel<<<P, RS_THREADS, smem, (cudaStream_t)stream>>>(x, ld, n, taps, add, out);
    CSR_LAUNCHED();
    return CSR_OK;
}

extern "C" int csr_peak(const float* F, int ld, int n, int P, int* argmax, float* stats, void* stream) {
    CSR_REQUIRE(F && argmax && stats && P > 0 && n > 0 && ld >= 2 * n, "csr_peak: bad argument");
    peak_kernel<<<P, RS_THREADS, 0, (cudaStream_t)stream>>>(F, ld, n, argmax, stats);
    CSR_LAUNCHED();
    return CSR_OK;
}
