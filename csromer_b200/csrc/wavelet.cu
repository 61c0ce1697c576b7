// wavelet.cu -- PyWavelets-compatible undecimated (swt/iswt) and decimated (wavedec/waverec, periodization)
// transforms of the real planar vector [Re|Im] of every line of sight (K5 building blocks).
//
// Replaces dictionaries/undecimated.py:45-86,141-171 (pywt.swt / pywt.iswt / ravel_coeffs) and
// dictionaries/discrete.py:28-44,70-91 (pywt.wavedec / pywt.waverec, mode="periodization").  PyWavelets
// 1.4.1 is not part of the reference tree; its published algorithm is restated:
//   swt level l (step = 2^(l-1)):  c[o] = sum_t f[t] a[(o + (F*step)/2 - t*step) mod N]
//   iswt level l: pywt averages the two interleaved periodization IDWTs; summed over both parities this is
//       a'[o] = 1/2 sum_t ( rec_lo[t] a[(o + step*(F/2-1-t)) mod N] + rec_hi[t] d[(o + step*(F/2-1-t)) mod N] )
//   dwt (periodization): c[o] = sum_t f[t] x'[(2o + F/2 - t) mod N'],  x' = x with its last sample repeated
//       once when N is odd;   idwt: x[(2o + t - F/2 + 1) mod N] += cA[o] rec_lo[t] + cD[o] rec_hi[t]
// Every level is one gather kernel over [P, band] (grid.y = line of sight), memory-bound and coalesced.
#include <string.h>

#include "plan.cuh"

namespace csr {

struct FilterPair {
    float lo[64];
    float hi[64];
    int F;
};

// ---- undecimated ------------------------------------------------------------------------------------------
__global__ void swt_level_kernel(const float* __restrict__ a_in, int ld_in, float* __restrict__ a_out, int ld_a,
                                 float* __restrict__ d_out, int ld_d, int N, int step, const FilterPair f) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= N) return;
    const int p = blockIdx.y;
    const float* src = a_in + (size_t)p * ld_in;
    int idx = (o + (f.F * step) / 2) % N;
    float ca = 0.f, cd = 0.f;
    const int dec = step % N;
    for (int t = 0; t < f.F; ++t) {
        const float v = src[idx];
        ca = fmaf(f.lo[t], v, ca);
        cd = fmaf(f.hi[t], v, cd);
        idx -= dec;
        if (idx < 0) idx += N;
    }
    a_out[(size_t)p * ld_a + o] = ca;
    d_out[(size_t)p * ld_d + o] = cd;
}

// one inverse level; `add` (optional, pitch ld_add) is added to the result (append_signal on the last level)
__global__ void iswt_level_kernel(const float* __restrict__ a_in, int ld_a, const float* __restrict__ d_in, int ld_d,
                                  float* __restrict__ out, int ld_out, const float* __restrict__ add, int ld_add, int N,
                                  int step, const FilterPair f) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= N) return;
    const int p = blockIdx.y;
    const float* a = a_in + (size_t)p * ld_a;
    const float* d = d_in + (size_t)p * ld_d;
    // idx(t) = (o + step*(F/2 - 1 - t)) mod N, walked downwards in t
    long long start = (long long)o + (long long)step * (f.F / 2 - 1);
    int idx = (int)(((start % N) + N) % N);
    const int dec = step % N;
    float acc = 0.f;
    for (int t = 0; t < f.F; ++t) {
        acc = fmaf(f.lo[t], a[idx], acc);
        acc = fmaf(f.hi[t], d[idx], acc);
        idx -= dec;
        if (idx < 0) idx += N;
    }
    acc *= 0.5f;
    if (add) acc += add[(size_t)p * ld_add + o];
    out[(size_t)p * ld_out + o] = acc;
}

// ---- decimated, periodization ----------------------------------------------------------------------------
__global__ void dwt_level_kernel(const float* __restrict__ x_in, int ld_in, int N, float* __restrict__ a_out, int ld_a,
                                 float* __restrict__ d_out, int ld_d, const FilterPair f) {
    const int Np = N + (N & 1);
    const int half = Np / 2;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= half) return;
    const int p = blockIdx.y;
    const float* src = x_in + (size_t)p * ld_in;
    int idx = (2 * o + f.F / 2) % Np;
    float ca = 0.f, cd = 0.f;
    for (int t = 0; t < f.F; ++t) {
        const float v = src[min(idx, N - 1)];  // index N (only when N is odd) repeats the last sample
        ca = fmaf(f.lo[t], v, ca);
        cd = fmaf(f.hi[t], v, cd);
        idx -= 1;
        if (idx < 0) idx += Np;
    }
    a_out[(size_t)p * ld_a + o] = ca;
    d_out[(size_t)p * ld_d + o] = cd;
}

// x[q], q in [0, 2*half): gather form of the periodization IDWT.  `keep` = number of leading samples of the
// result the next level needs (waverec drops the last one when it is one too long); `add` as in iswt.
__global__ void idwt_level_kernel(const float* __restrict__ a_in, int ld_a, const float* __restrict__ d_in, int ld_d,
                                  int half, float* __restrict__ out, int ld_out, int keep,
                                  const float* __restrict__ add, int ld_add, const FilterPair f) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= keep) return;
    const int p = blockIdx.y;
    const int N = 2 * half;
    const float* a = a_in + (size_t)p * ld_a;
    const float* d = d_in + (size_t)p * ld_d;
    float acc = 0.f;
    const int base = q + f.F / 2 - 1;
    for (int t = (base & 1); t < f.F; t += 2) {  // (base - t) must be even
        int e = (base - t) % N;
        if (e < 0) e += N;
        const int o = e >> 1;
        acc = fmaf(f.lo[t], a[o], acc);
        acc = fmaf(f.hi[t], d[o], acc);
    }
    if (add) acc += add[(size_t)p * ld_add + q];
    out[(size_t)p * ld_out + q] = acc;
}

__global__ void copy_rows_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst,
                                 int cols) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < cols) dst[(size_t)blockIdx.y * ld_dst + c] = src[(size_t)blockIdx.y * ld_src + c];
}

static FilterPair make_pair(const float* lo, const float* hi, int F) {
    FilterPair f;
    memset(&f, 0, sizeof(f));
    for (int t = 0; t < F; ++t) {
        f.lo[t] = lo[t];
        f.hi[t] = hi[t];
    }
    f.F = F;
    return f;
}

static inline dim3 grid_for(int cols, int P) { return dim3((cols + 255) / 256, P); }

int wavelet_ld(const csr_wavelet* w) { return round_up(w->N, 8); }

// coefficient layout: [signal (N) if append_signal][cA_L][cD_L]...[cD_1]
int wavelet_decompose(const csr_wavelet* w, const float* signal, int ld_sig, float* coef, int ld_coef, int P,
                      float* ws, cudaStream_t stream) {
    const int N = w->N, L = w->levels;
    const int ldw = wavelet_ld(w);
    float* ping = ws;
    float* pong = ws + (size_t)P * ldw;
    const int base = w->append_signal ? N : 0;
    if (w->append_signal) {
        copy_rows_kernel<<<grid_for(N, P), 256, 0, stream>>>(signal, ld_sig, coef, ld_coef, N);
        CSR_LAUNCHED();
    }
    const FilterPair f = make_pair(w->dec_lo, w->dec_hi, w->F);
    const float* cur = signal;
    int ld_cur = ld_sig;
    int len = N;
    for (int lev = 1; lev <= L; ++lev) {
        // band index of cD_lev in [cA_L, cD_L, ..., cD_1] is L - lev + 1
        float* d_out = coef + base + w->offsets[L - lev + 1];
        float* a_out;
        int ld_a;
        if (lev == L) {
            a_out = coef + base + w->offsets[0];
            ld_a = ld_coef;
        } else {
            a_out = (lev & 1) ? ping : pong;
            ld_a = ldw;
        }
        if (w->kind == 0) {
            swt_level_kernel<<<grid_for(N, P), 256, 0, stream>>>(cur, ld_cur, a_out, ld_a, d_out, ld_coef, N,
                                                                 1 << (lev - 1), f);
        } else {
            dwt_level_kernel<<<grid_for((len + 1) / 2, P), 256, 0, stream>>>(cur, ld_cur, len, a_out, ld_a, d_out,
                                                                             ld_coef, f);
            len = (len + 1) / 2;
        }
        CSR_LAUNCHED();
        cur = a_out;
        ld_cur = ld_a;
    }
    if (L == 0) {  // degenerate: coefficients = signal
        copy_rows_kernel<<<grid_for(N, P), 256, 0, stream>>>(signal, ld_sig, coef + base, ld_coef, N);
        CSR_LAUNCHED();
    }
    return CSR_OK;
}

int wavelet_reconstruct(const csr_wavelet* w, const float* coef, int ld_coef, float* signal, int ld_sig, int P,
                        float* ws, cudaStream_t stream) {
    const int N = w->N, L = w->levels;
    const int ldw = wavelet_ld(w);
    float* ping = ws;
    float* pong = ws + (size_t)P * ldw;
    const int base = w->append_signal ? N : 0;
    const FilterPair f = make_pair(w->rec_lo, w->rec_hi, w->F);
    const float* cur = coef + base + w->offsets[0];
    int ld_cur = ld_coef;
    if (L == 0) {
        copy_rows_kernel<<<grid_for(N, P), 256, 0, stream>>>(cur, ld_cur, signal, ld_sig, N);
        CSR_LAUNCHED();
        return CSR_OK;
    }
    for (int lev = L; lev >= 1; --lev) {
        const int band = L - lev + 1;
        const float* d_in = coef + base + w->offsets[band];
        float* out;
        int ld_out;
        const float* add = nullptr;
        if (lev == 1) {
            out = signal;
            ld_out = ld_sig;
            if (w->append_signal) add = coef;
        } else {
            out = (lev & 1) ? ping : pong;
            ld_out = ldw;
        }
        if (w->kind == 0) {
            iswt_level_kernel<<<grid_for(N, P), 256, 0, stream>>>(cur, ld_cur, d_in, ld_coef, out, ld_out, add, ld_coef,
                                                                  N, 1 << (lev - 1), f);
        } else {
            const int half = w->sizes[band];  // cA (after trimming) and cD have this many samples
            const int keep = (lev == 1) ? N : w->sizes[band + 1];
            idwt_level_kernel<<<grid_for(keep, P), 256, 0, stream>>>(cur, ld_cur, d_in, ld_coef, half, out, ld_out,
                                                                     keep, add, ld_coef, f);
        }
        CSR_LAUNCHED();
        cur = out;
        ld_cur = ld_out;
    }
    return CSR_OK;
}

}  // namespace csr

using namespace csr;

extern "C" int csr_wavelet_create(csr_wavelet** out, int kind, const double* dec_lo, const double* dec_hi,
                                  const double* rec_lo, const double* rec_hi, int F, int level, int N,
                                  int append_signal, int norm, int device) {
    CSR_REQUIRE(out && dec_lo && dec_hi && rec_lo && rec_hi, "csr_wavelet_create: null pointer");
    CSR_REQUIRE(kind == 0 || kind == 1, "csr_wavelet_create: kind must be 0 (undecimated) or 1 (decimated)");
    CSR_REQUIRE(F >= 2 && F <= 64 && F % 2 == 0, "csr_wavelet_create: filter length must be even and in [2, 64] (F=%d)", F);
    CSR_REQUIRE(N > 0 && level >= 0 && level <= 32, "csr_wavelet_create: bad N=%d or level=%d", N, level);
    if (kind == 0) CSR_REQUIRE(level == 0 || N % (1 << level) == 0, "csr_wavelet_create: N=%d is not a multiple of 2^%d", N, level);
    *out = nullptr;
    csr_wavelet* w = new csr_wavelet();
    memset(w, 0, sizeof(*w));
    w->device = device;
    w->kind = kind;
    w->F = F;
    w->levels = level;
    w->N = N;
    w->append_signal = append_signal ? 1 : 0;
    w->norm = norm ? 1 : 0;
    // pywt: swt(norm=True) scales the filter bank by 1/sqrt(2), iswt(norm=True) by sqrt(2)
    const double sa = (kind == 0 && norm) ? 0.70710678118654752440 : 1.0;
    const double ss = (kind == 0 && norm) ? 1.41421356237309504880 : 1.0;
    for (int t = 0; t < F; ++t) {
        w->dec_lo[t] = (float)(dec_lo[t] * sa);
        w->dec_hi[t] = (float)(dec_hi[t] * sa);
        w->rec_lo[t] = (float)(rec_lo[t] * ss);
        w->rec_hi[t] = (float)(rec_hi[t] * ss);
    }
    // band sizes [cA_L, cD_L, ..., cD_1]
    if (kind == 0) {
        for (int b = 0; b <= level; ++b) w->sizes[b] = N;
    } else {
        int len = N;
        for (int lev = 1; lev <= level; ++lev) {
            len = (len + 1) / 2;
            w->sizes[level - lev + 1] = len;
        }
        w->sizes[0] = len;
        if (level == 0) w->sizes[0] = N;
    }
    int off = 0;
    for (int b = 0; b <= level; ++b) {
        w->offsets[b] = off;
        off += w->sizes[b];
    }
    w->ncoef = off + (w->append_signal ? N : 0);
    *out = w;
    return CSR_OK;
}

extern "C" void csr_wavelet_destroy(csr_wavelet* wav) { delete wav; }
extern "C" int csr_wavelet_ncoef(const csr_wavelet* wav) { return wav ? wav->ncoef : 0; }
extern "C" int csr_wavelet_levels(const csr_wavelet* wav) { return wav ? wav->levels : 0; }
extern "C" int csr_wavelet_level_sizes(const csr_wavelet* wav, int* sizes, int cap) {
    CSR_REQUIRE(wav && sizes && cap > wav->levels, "csr_wavelet_level_sizes: bad argument");
    for (int b = 0; b <= wav->levels; ++b) sizes[b] = wav->sizes[b];
    return wav->levels + 1;
}
extern "C" size_t csr_wavelet_ws_bytes(const csr_wavelet* wav, int P) {
    if (!wav || P <= 0) return 0;
    return align256((size_t)2 * P * wavelet_ld(wav) * sizeof(float));
}

extern "C" int csr_wavelet_decompose(csr_wavelet* wav, const float* signal, int ld_sig, float* coef, int ld_coef,
                                     int P, void* ws, size_t ws_bytes, void* stream) {
    CSR_REQUIRE(wav && signal && coef && P > 0, "csr_wavelet_decompose: bad argument");
    CSR_REQUIRE(ld_sig >= wav->N && ld_coef >= wav->ncoef, "csr_wavelet_decompose: pitch too small");
    if (ws_bytes < csr_wavelet_ws_bytes(wav, P) || !ws) {
        set_error("csr_wavelet_decompose: workspace too small (%zu < %zu)", ws_bytes, csr_wavelet_ws_bytes(wav, P));
        return CSR_ERR_WORKSPACE;
    }
    CSR_CUDA(cudaSetDevice(wav->device));
    return wavelet_decompose(wav, signal, ld_sig, coef, ld_coef, P, (float*)ws, (cudaStream_t)stream);
}

extern "C" int csr_wavelet_reconstruct(csr_wavelet* wav, const float* coef, int ld_coef, float* signal, int ld_sig,
                                       int P, void* ws, size_t ws_bytes, void* stream) {
    CSR_REQUIRE(wav && signal && coef && P > 0, "csr_wavelet_reconstruct: bad argument");
    CSR_REQUIRE(ld_sig >= wav->N && ld_coef >= wav->ncoef, "csr_wavelet_reconstruct: pitch too small");
    if (ws_bytes < csr_wavelet_ws_bytes(wav, P) || !ws) {
        set_error("csr_wavelet_reconstruct: workspace too small (%zu < %zu)", ws_bytes, csr_wavelet_ws_bytes(wav, P));
        return CSR_ERR_WORKSPACE;
    }
    CSR_CUDA(cudaSetDevice(wav->device));
    return wavelet_reconstruct(wav, coef, ld_coef, signal, ld_sig, P, (float*)ws, (cudaStream_t)stream);
}
