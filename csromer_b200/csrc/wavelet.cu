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
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

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
    const int o = blockIdx.y * blockDim.x + threadIdx.x;
    if (o >= N) return;
    const int p = blockIdx.x;
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
    const int o = blockIdx.y * blockDim.x + threadIdx.x;
    if (o >= N) return;
    const int p = blockIdx.x;
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
// one output pair of a level; `src` may be global or shared memory (the per-level kernels and the fused kernel share these
// functions: same multiply-adds in the same order, same bits)
__device__ __forceinline__ void dwt_per_sample(const float* src, int N, int o, const FilterPair& f, float& ca, float& cd) {
    const int Np = N + (N & 1);
    int idx = (2 * o + f.F / 2) % Np;
    ca = 0.f;
    cd = 0.f;
    for (int t = 0; t < f.F; ++t) {
        const float v = src[min(idx, N - 1)];  // index N (only when N is odd) repeats the last sample
        ca = fmaf(f.lo[t], v, ca);
        cd = fmaf(f.hi[t], v, cd);
        idx -= 1;
        if (idx < 0) idx += Np;
    }
}
__global__ void dwt_level_kernel(const float* __restrict__ x_in, int ld_in, int N, float* __restrict__ a_out, int ld_a,
                                 float* __restrict__ d_out, int ld_d, const FilterPair f) {
    const int half = (N + (N & 1)) / 2;
    const int o = blockIdx.y * blockDim.x + threadIdx.x;
    if (o >= half) return;
    const int p = blockIdx.x;
    float ca, cd;
    dwt_per_sample(x_in + (size_t)p * ld_in, N, o, f, ca, cd);
    a_out[(size_t)p * ld_a + o] = ca;
    d_out[(size_t)p * ld_d + o] = cd;
}

// x[q], q in [0, 2*half): gather form of the periodization IDWT.  `keep` = number of leading samples of the
// result the next level needs (waverec drops the last one when it is one too long); `add` as in iswt.
// ---- decimated transform with the other PyWavelets signal-extension modes (dictionaries/discrete.py:37-39 hands
// ``mode`` to pywt.wavedec): full convolution of the extended signal sampled at the odd positions,
// cA[k] = sum_j lo[j] x_ext[2k + 1 - j], k < (N + F - 1)/2; the inverse is the extension-independent ("valid") part of the
// up-sampled convolution, out[q] = sum_k cA[k] rlo[q + F - 2 - 2k] + cD[k] rhi[..], q < 2 Nc - F + 2.
enum ExtMode { EXT_PER = 0, EXT_ZERO, EXT_CONSTANT, EXT_SYMMETRIC, EXT_REFLECT, EXT_PERIODIC, EXT_SMOOTH, EXT_ANTISYMMETRIC,
               EXT_ANTIREFLECT, EXT_COUNT };

__device__ __forceinline__ int floor_mod(int i, int p) {
    int j = i % p;
    return j < 0 ? j + p : j;
}

__device__ __forceinline__ float ext_sample(const float* __restrict__ x, int i, int N, int ext) {
    if (i >= 0 && i < N) return x[i];
    switch (ext) {
        case EXT_ZERO: return 0.f;
        case EXT_CONSTANT: return x[i < 0 ? 0 : N - 1];
        case EXT_PERIODIC: return x[floor_mod(i, N)];
        case EXT_SYMMETRIC: {
            const int j = floor_mod(i, 2 * N);
            return x[j < N ? j : 2 * N - 1 - j];
        }
        case EXT_ANTISYMMETRIC: {
            const int j = floor_mod(i, 2 * N);
            return j < N ? x[j] : -x[2 * N - 1 - j];
        }
        case EXT_REFLECT: {
            if (N == 1) return x[0];
            const int p = 2 * N - 2, j = floor_mod(i, p);
            return x[j < N ? j : p - j];
        }
        case EXT_ANTIREFLECT: {
            if (N == 1) return x[0];
            const int p = 2 * N - 2, j = floor_mod(i, p), cyc = (i - j) / p;
            const float base = j < N ? x[j] : 2.f * x[N - 1] - x[p - j];
            return fmaf(2.f * (float)cyc, x[N - 1] - x[0], base);
        }
        case EXT_SMOOTH: {
            if (N == 1) return x[0];
            return i < 0 ? fmaf((float)i, x[1] - x[0], x[0]) : fmaf((float)(i - (N - 1)), x[N - 1] - x[N - 2], x[N - 1]);
        }
    }
    return 0.f;
}

__device__ __forceinline__ void dwt_ext_sample(const float* src, int N, int ext, int k, const FilterPair& f, float& ca, float& cd) {
    ca = 0.f;
    cd = 0.f;
    const int top = 2 * k + 1;
    if (top - (f.F - 1) >= 0 && top < N) {  // interior: no extension involved
        for (int t = 0; t < f.F; ++t) {
            const float v = src[top - t];
            ca = fmaf(f.lo[t], v, ca);
            cd = fmaf(f.hi[t], v, cd);
        }
    } else {
        for (int t = 0; t < f.F; ++t) {
            const float v = ext_sample(src, top - t, N, ext);
            ca = fmaf(f.lo[t], v, ca);
            cd = fmaf(f.hi[t], v, cd);
        }
    }
}
__global__ void dwt_ext_level_kernel(const float* __restrict__ x_in, int ld_in, int N, int ext, float* __restrict__ a_out,
                                     int ld_a, float* __restrict__ d_out, int ld_d, const FilterPair f) {
    const int nc = (N + f.F - 1) / 2;
    const int k = blockIdx.y * blockDim.x + threadIdx.x;
    if (k >= nc) return;
    const int p = blockIdx.x;
    float ca, cd;
    dwt_ext_sample(x_in + (size_t)p * ld_in, N, ext, k, f, ca, cd);
    a_out[(size_t)p * ld_a + k] = ca;
    d_out[(size_t)p * ld_d + k] = cd;
}

// out has ``keep`` <= 2 nc - F + 2 samples (the running approximation is cut to the next detail band's length, like waverec)
__device__ __forceinline__ float idwt_ext_sample(const float* a, const float* d, int nc, int q, const FilterPair& f) {
    float acc = 0.f;
    const int base = q + f.F - 2;  // tap t = base - 2k
    for (int t = (base & 1); t < f.F; t += 2) {
        const int k = (base - t) >> 1;
        if (k >= 0 && k < nc) {
            acc = fmaf(f.lo[t], a[k], acc);
            acc = fmaf(f.hi[t], d[k], acc);
        }
    }
    return acc;
}
__global__ void idwt_ext_level_kernel(const float* __restrict__ a_in, int ld_a, const float* __restrict__ d_in, int ld_d,
                                      int nc, float* __restrict__ out, int ld_out, int keep,
                                      const float* __restrict__ add, int ld_add, const FilterPair f) {
    const int q = blockIdx.y * blockDim.x + threadIdx.x;
    if (q >= keep) return;
    const int p = blockIdx.x;
    float acc = idwt_ext_sample(a_in + (size_t)p * ld_a, d_in + (size_t)p * ld_d, nc, q, f);
    if (add) acc += add[(size_t)p * ld_add + q];
    out[(size_t)p * ld_out + q] = acc;
}

__device__ __forceinline__ float idwt_per_sample(const float* a, const float* d, int half, int q, const FilterPair& f) {
    const int N = 2 * half;
    float acc = 0.f;
    const int base = q + f.F / 2 - 1;
    for (int t = (base & 1); t < f.F; t += 2) {  // (base - t) must be even
        int e = (base - t) % N;
        if (e < 0) e += N;
        const int o = e >> 1;
        acc = fmaf(f.lo[t], a[o], acc);
        acc = fmaf(f.hi[t], d[o], acc);
    }
    return acc;
}
__global__ void idwt_level_kernel(const float* __restrict__ a_in, int ld_a, const float* __restrict__ d_in, int ld_d,
                                  int half, float* __restrict__ out, int ld_out, int keep,
                                  const float* __restrict__ add, int ld_add, const FilterPair f) {
    const int q = blockIdx.y * blockDim.x + threadIdx.x;
    if (q >= keep) return;
    const int p = blockIdx.x;
    float acc = idwt_per_sample(a_in + (size_t)p * ld_a, d_in + (size_t)p * ld_d, half, q, f);
    if (add) acc += add[(size_t)p * ld_add + q];
    out[(size_t)p * ld_out + q] = acc;
}

// ---- decimated transform, all levels of one line of sight in ONE kernel ------------------------------------------------
// The level-by-level path costs one launch and one round trip through HBM per level (12 levels for db2 at N = 15936).  Here a
// CTA keeps the running approximation of its line of sight in shared memory (N + N/2 floats): the signal is read once, every
// detail band is written once, the approximations never leave the SM.  Same sample functions as the per-level kernels: same bits.
struct DwtBands {
    int L, N, ext, base;     // levels, signal length, extension mode (0 = periodization), offset of the bands in a coefficient row
    int sizes[34], offsets[34];
};
constexpr int DWT_FUSED_THREADS = 512;

__global__ void __launch_bounds__(DWT_FUSED_THREADS)
dwt_fused_kernel(const float* __restrict__ signal, int ld_sig, float* __restrict__ coef, int ld_coef, const DwtBands b, const FilterPair f,
                 int append_signal) {
    extern __shared__ float dwt_smem[];
    const int p = blockIdx.x;
    const int n1 = b.sizes[b.L];  // length of cD_1 = of the first approximation
    float* cur = dwt_smem;        // [N]
    float* nxt = dwt_smem + ((b.N + 3) & ~3);  // [n1]
    const float* src = signal + (size_t)p * ld_sig;
    float* row = coef + (size_t)p * ld_coef;
    for (int i = threadIdx.x; i < b.N; i += blockDim.x) {
        const float v = src[i];
        cur[i] = v;
        if (append_signal) row[i] = v;
    }
    __syncthreads();
    int len = b.N;
    (void)n1;
    for (int lev = 1; lev <= b.L; ++lev) {
        const int band = b.L - lev + 1;
        const int nc = b.sizes[band];
        float* d_out = row + b.base + b.offsets[band];
        float* a_out = (lev == b.L) ? row + b.base + b.offsets[0] : nxt;
        for (int k = threadIdx.x; k < nc; k += blockDim.x) {
            float ca, cd;
            if (b.ext == 0) dwt_per_sample(cur, len, k, f, ca, cd);
            else dwt_ext_sample(cur, len, b.ext, k, f, ca, cd);
            a_out[k] = ca;
            d_out[k] = cd;
        }
        __syncthreads();
        float* t = cur;
        cur = nxt;
        nxt = t;
        len = nc;
    }
}

__global__ void __launch_bounds__(DWT_FUSED_THREADS)
idwt_fused_kernel(const float* __restrict__ coef, int ld_coef, float* __restrict__ signal, int ld_sig, const DwtBands b, const FilterPair f,
                  int add_signal) {
    extern __shared__ float dwt_smem[];
    const int p = blockIdx.x;
    const int n1 = b.sizes[b.L];
    float* cur = dwt_smem;                      // [n1]
    float* nxt = dwt_smem + ((n1 + 3) & ~3);    // [n1]
    const float* row = coef + (size_t)p * ld_coef;
    for (int i = threadIdx.x; i < b.sizes[0]; i += blockDim.x) cur[i] = row[b.base + b.offsets[0] + i];
    __syncthreads();
    for (int lev = b.L; lev >= 1; --lev) {
        const int band = b.L - lev + 1;
        const int half = b.sizes[band];
        const int keep = (lev == 1) ? b.N : b.sizes[band + 1];
        const float* d_in = row + b.base + b.offsets[band];
        for (int q = threadIdx.x; q < keep; q += blockDim.x) {
            float v = b.ext == 0 ? idwt_per_sample(cur, d_in, half, q, f) : idwt_ext_sample(cur, d_in, half, q, f);
            if (lev == 1) {
                if (add_signal) v += row[q];
                signal[(size_t)p * ld_sig + q] = v;
            } else {
                nxt[q] = v;
            }
        }
        __syncthreads();
        float* t = cur;
        cur = nxt;
        nxt = t;
    }
}

__global__ void copy_rows_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int ld_dst,
                                 int cols) {
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c < cols) dst[(size_t)blockIdx.x * ld_dst + c] = src[(size_t)blockIdx.x * ld_src + c];
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

// lines of sight on grid.x (limit 2^31 - 1), column blocks on grid.y
static inline dim3 grid_for(int cols, int P) { return dim3(P, (cols + 255) / 256); }

// pitch of the two scratch slabs: the longest intermediate band (extension modes: (N + F - 1)/2 may exceed N for tiny N)
int wavelet_ld(const csr_wavelet* w) { return round_up(w->ext ? max(w->N, (w->N + w->F) / 2 + 1) : w->N, 8); }

// ---- fused undecimated transforms: all levels of one tile of one line of sight in shared memory ---------------
// A tile of T samples plus a halo of H = (F/2)(2^L - 1) samples on each side (the cumulative support of L a-trous
// levels) is loaded once, every level is computed buffer-to-buffer in shared memory, and only final results touch
// HBM.  Buffer index i <-> global sample (t0 - H + i) mod N (periodic signal).  At every level only the entries
// whose taps stay inside the buffer are computed; the uncomputed rim lies inside the halo, never in the T valid outputs.
constexpr int WAV_THREADS = 512;
constexpr int WAV_SMEM_BUDGET = 72 * 1024;  // per CTA: three CTAs per SM overlap each other's load / compute phases

struct FusedTiling {
    int H, T, B, ntiles;
};
static FusedTiling fused_tiling(const csr_wavelet* w, int nbuf) {
    FusedTiling t;
    t.H = (w->F / 2) * ((1 << w->levels) - 1);
    const int bmax = WAV_SMEM_BUDGET / (4 * nbuf);
    const int tmax = bmax - 2 * t.H;
    if (tmax < 256) {
        t.T = t.B = t.ntiles = 0;  // filter/level combination too wide for the budget: caller falls back
        return t;
    }
    t.ntiles = (w->N + tmax - 1) / tmax;
    t.T = round_up((w->N + t.ntiles - 1) / t.ntiles, 32);
    t.B = t.T + 2 * t.H;
    return t;
}
bool wavelet_fusable(const csr_wavelet* w) {
    return w->kind == 0 && !w->append_signal && w->levels >= 1 && fused_tiling(w, 3).T > 0;
}

__device__ __forceinline__ int wrap_index(int j, int N) {
    j %= N;
    return j < 0 ? j + N : j;
}

struct BandOffsets {
    int off[34];
};

// Inner loops with the filter length known at compile time: the taps unroll, the coefficients become constant-bank
// operands of the FMAs and the loads of one output are all issued before its FMAs.  FLEN = 0 is the generic loop.
// USE_A / USE_D: an operand whose whole tile is identically zero is not read (its contribution is exactly 0)
template <int FLEN, bool USE_A, bool USE_D>
__device__ __forceinline__ void synth_level(const float* __restrict__ a, const float* __restrict__ d,
                                            float* __restrict__ o, int B, int step, const FilterPair& f) {
    const int F = FLEN ? FLEN : f.F;
    const int rim = step * (F / 2);
    for (int i = rim + threadIdx.x; i < B - rim; i += WAV_THREADS) {
        const float* ap = a + i + step * (F / 2 - 1);
        const float* dp = d + i + step * (F / 2 - 1);
        float acc0 = 0.f, acc1 = 0.f;
        if (FLEN) {
#pragma unroll
            for (int t = 0; t < FLEN; ++t) {
                if (USE_A) acc0 = fmaf(f.lo[t], ap[-t * step], acc0);
                if (USE_D) acc1 = fmaf(f.hi[t], dp[-t * step], acc1);
            }
        } else {
            for (int t = 0; t < F; ++t) {
                if (USE_A) acc0 = fmaf(f.lo[t], ap[-t * step], acc0);
                if (USE_D) acc1 = fmaf(f.hi[t], dp[-t * step], acc1);
            }
        }
        o[i] = 0.5f * (acc0 + acc1);
    }
}

template <int FLEN>
__device__ __forceinline__ void analysis_taps(const float* __restrict__ ap, int step, const FilterPair& f, float& ca,
                                              float& cd) {
    ca = 0.f;
    cd = 0.f;
    if (FLEN) {
#pragma unroll
        for (int t = 0; t < FLEN; ++t) {
            const float v = ap[-t * step];
            ca = fmaf(f.lo[t], v, ca);
            cd = fmaf(f.hi[t], v, cd);
        }
    } else {
        for (int t = 0; t < f.F; ++t) {
            const float v = ap[-t * step];
            ca = fmaf(f.lo[t], v, ca);
            cd = fmaf(f.hi[t], v, cd);
        }
    }
}

#define CSR_DISPATCH_FLEN(FVAL, CALL) \
    switch (FVAL) {                   \
        case 2: { constexpr int FL = 2; CALL; } break;   \
        case 4: { constexpr int FL = 4; CALL; } break;   \
        case 6: { constexpr int FL = 6; CALL; } break;   \
        case 8: { constexpr int FL = 8; CALL; } break;   \
        case 12: { constexpr int FL = 12; CALL; } break; \
        default: { constexpr int FL = 0; CALL; } break;  \
    }

// coefficients [cA_L | cD_L | ... | cD_1] -> signal; writes fp32 (sig) and/or the scaled fp16 hi/lo GEMM operand
__global__ void __launch_bounds__(WAV_THREADS)
swt_synth_fused_kernel(const float* __restrict__ coef, int ld_coef, int N, int L, int H, int T, int B,
                       const BandOffsets bands, const FilterPair f, float* __restrict__ sig, int ld_sig,
                       __half* __restrict__ hi, __half* __restrict__ lo, int ld_h, const float* __restrict__ scale,
                       unsigned char* __restrict__ tile_flags, int flag_ld) {
    extern __shared__ float wsm[];
    float* a = wsm;
    float* o = wsm + B;
    float* d = wsm + 2 * B;
    const int p = blockIdx.x;
    const int t0 = blockIdx.y * T;
    const int base = t0 - H;
    const float* crow = coef + (size_t)p * ld_coef;
    // Sparsity shortcut (exact): thresholded coefficient vectors are mostly zero.  A band tile that is identically
    // zero contributes exactly 0, so it is not staged; while the running approximation is itself identically zero
    // the whole level is skipped.  Which band tiles hold anything is found in ONE pass: every thread tests its
    // samples of all L + 1 bands with independent loads (one memory latency instead of L + 1 dependent load / vote
    // phases -- most tiles are zero in every band, and they are done after this pass), then the non-zero bands
    // are staged (their second read hits L2).
    __shared__ unsigned int band_nz_s;
    if (threadIdx.x == 0) band_nz_s = 0u;
    __syncthreads();
    {
        unsigned int mine = 0u;
        for (int i = threadIdx.x; i < B; i += WAV_THREADS) {
            const int gi = wrap_index(base + i, N);
            for (int b = 0; b <= L; ++b) mine |= (crow[bands.off[b] + gi] != 0.f) ? (1u << b) : 0u;
        }
        mine = __reduce_or_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0 && mine) atomicOr(&band_nz_s, mine);
    }
    __syncthreads();
    const unsigned int band_nz = band_nz_s;
    bool a_nz = (band_nz & 1u) != 0;
    if (a_nz) {
        for (int i = threadIdx.x; i < B; i += WAV_THREADS) a[i] = crow[bands.off[0] + wrap_index(base + i, N)];
    }
    for (int lev = L; lev >= 1; --lev) {
        const int step = 1 << (lev - 1);
        const float* dband = crow + bands.off[L - lev + 1];
        const bool d_nz = ((band_nz >> (L - lev + 1)) & 1u) != 0;
        if (d_nz) {
            for (int i = threadIdx.x; i < B; i += WAV_THREADS) d[i] = dband[wrap_index(base + i, N)];
        }
        if (a_nz || d_nz) __syncthreads();  // staged operands visible (uniform branch)
        if (!a_nz && !d_nz) continue;  // result identically zero: nothing to compute, `a` stays "zero"
        // only entries whose taps stay inside the buffer are computed (no per-tap clamping); the uncomputed rim
        // grows by step*F/2 per level and is covered by the halo
        if (a_nz && d_nz) {
            CSR_DISPATCH_FLEN(f.F, (synth_level<FL, true, true>(a, d, o, B, step, f)))
        } else if (a_nz) {
            CSR_DISPATCH_FLEN(f.F, (synth_level<FL, true, false>(a, d, o, B, step, f)))
        } else {
            CSR_DISPATCH_FLEN(f.F, (synth_level<FL, false, true>(a, d, o, B, step, f)))
        }
        a_nz = true;
        __syncthreads();
        float* tmp = a;
        a = o;
        o = tmp;
    }
    const float sc = scale ? scale[p] : 1.f;
    for (int i = H + threadIdx.x; i < H + T; i += WAV_THREADS) {
        const int g = t0 + (i - H);
        if (g >= N) break;
        const float v = a_nz ? a[i] : 0.f;
        if (sig) sig[(size_t)p * ld_sig + g] = v;
        if (hi) {
            __half h, l;
            split_half(v * sc, h, l);
            hi[(size_t)p * ld_h + g] = h;
            lo[(size_t)p * ld_h + g] = l;
            // forward GEMM K-block skipping (dft_gemm.cu): mark the (pixel tile, 128-column block) this non-zero
            // operand value falls in; the caller zeroed the flags, untouched blocks are identically zero
            if (tile_flags && v != 0.f) tile_flags[((size_t)(p >> 7) * flag_ld + (g >> 7)) * 4 + ((p >> 5) & 3)] = 1;
        }
    }
}

// u = z - step g ; x = soft(u, step lambda) ; z = x + beta (x - x_old)   (fista.py:80-86, l1.py:30-32).  The roundings
// are spelled out so that every kernel of this file produces the same bits whatever the compiler would contract.
__device__ __forceinline__ float soft_update(float zo, float xo, float g, float ustep, float thr, float beta, float& znew,
                                             float& dsum) {
    const float u = __fmaf_rn(-ustep, g, zo);
    const float xn = copysignf(fmaxf(fabsf(u) - thr, 0.f), u);
    const float dx = __fsub_rn(xn, xo);
    znew = __fmaf_rn(beta, dx, xn);
    dsum += fabsf(dx);
    return xn;
}

// one analysis level over the whole buffer, filter length fixed at compile time (the dispatch happens once per level,
// not per element).  d-band results are written (UPDATE = false) or consumed by the FISTA update (UPDATE = true).
template <int FLEN, bool UPDATE>
__device__ __forceinline__ void analysis_level(const float* __restrict__ a, float* __restrict__ o, int B, int step, int H,
                                               int T, int t0, int N, const FilterPair& f, float* __restrict__ cband,
                                               float* __restrict__ xband, float* __restrict__ zband, float ustep,
                                               float thr, float beta, float& dsum) {
    const int F = FLEN ? FLEN : f.F;
    const int rim = step * (F / 2);
    const int vlo = H, vhi = min(H + T, H + (N - t0));  // buffer indices whose global sample exists and belongs to the tile
    for (int i = rim + threadIdx.x; i < B - rim; i += WAV_THREADS) {
        const bool valid = i >= vlo && i < vhi;
        const int g = t0 + (i - H);
        float zo = 0.f, xo = 0.f;
        if (UPDATE && valid) {  // issue the state loads before the filter loop
            zo = __ldcg(zband + g);
            xo = __ldcg(xband + g);
        }
        float ca, cd;
        analysis_taps<FLEN>(a + i + (F * step) / 2, step, f, ca, cd);
        o[i] = ca;
        if (valid) {
            if (UPDATE) {
                float zn;
                const float xn = soft_update(zo, xo, cd, ustep, thr, beta, zn, dsum);
                __stcs(xband + g, xn);
                __stcs(zband + g, zn);
            } else {
                cband[g] = cd;
            }
        }
    }
}

// signal -> coefficients.  UPDATE = false: write the coefficients (decompose).  UPDATE = true: the signal is the
// gradient in phi space; every coefficient g of its analysis is consumed on the spot by
// u = z - step g ; x = soft(u, step lambda) ; z = x + beta (x - x_old)   (fista.py:80-86, l1.py:30-32).
template <bool UPDATE>
__global__ void __launch_bounds__(WAV_THREADS)
swt_analysis_fused_kernel(const float* __restrict__ signal, int ld_sig, int N, int L, int H, int T, int B,
                          const BandOffsets bands, const FilterPair f, float* __restrict__ coef, int ld_coef,
                          const CoefUpdate up) {
    extern __shared__ float wsm[];
    float* a = wsm;
    float* o = wsm + B;
    const int p = blockIdx.x;
    const int t0 = blockIdx.y * T;
    const int base = t0 - H;
    float thr = 0.f;
    bool active = true;
    if (UPDATE) {
        const float l0 = up.lambda0[p], nz = up.noise[p];
        const float lam = __fmaf_rn(-(float)up.iter, nz, l0);
        active = (nz < l0) && (up.iter == 0 || lam > 0.f) && (!up.iter_cap || up.iter < up.iter_cap[p]);
        thr = __fmul_rn(up.step, lam);  // (explicit rounding: the compiler would contract this product into |u| - thr)
        if (!active) return;  // whole CTA: this line of sight is frozen
    }
    const float* srow = signal + (size_t)p * ld_sig;
    for (int i = threadIdx.x; i < B; i += WAV_THREADS) a[i] = srow[wrap_index(base + i, N)];
    __syncthreads();
    const size_t crow = (size_t)p * ld_coef;
    float dsum = 0.f;
    for (int lev = 1; lev <= L; ++lev) {
        const int step = 1 << (lev - 1);
        const int band = bands.off[L - lev + 1];
        float* cband = UPDATE ? nullptr : coef + crow + band;
        float* xband = UPDATE ? up.x + crow + band : nullptr;
        float* zband = UPDATE ? up.z + crow + band : nullptr;
        CSR_DISPATCH_FLEN(f.F, (analysis_level<FL, UPDATE>(a, o, B, step, H, T, t0, N, f, cband, xband, zband, up.step, thr,
                                                         up.beta, dsum)))
        __syncthreads();
        float* tmp = a;
        a = o;
        o = tmp;
    }
    for (int i = H + threadIdx.x; i < H + T; i += WAV_THREADS) {  // approximation band cA_L
        const int g = t0 + (i - H);
        if (g >= N) break;
        const float ca = a[i];
        if (UPDATE) {
            const float zo = up.z[crow + bands.off[0] + g], xo = up.x[crow + bands.off[0] + g];
            float zn;
            const float xn = soft_update(zo, xo, ca, up.step, thr, up.beta, zn, dsum);
            up.x[crow + bands.off[0] + g] = xn;
            up.z[crow + bands.off[0] + g] = zn;
        } else {
            coef[crow + bands.off[0] + g] = ca;
        }
    }
    if (UPDATE && up.delta != nullptr) {
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, sft);
        if ((threadIdx.x & 31) == 0 && dsum != 0.f) atomicAdd(up.delta + p, dsum);
    }
}

// ---- the same two kernels, four consecutive samples per thread ------------------------------------------------
// The one-sample-per-thread kernels above are bound by shared-memory load issue (F loads per output and level) and
// by memory-level parallelism (two 4-byte state loads in flight per thread).  Here a thread owns four consecutive
// samples: an a-trous level with step % 4 == 0 needs one 16-byte shared load per tap for all four outputs, steps 1
// and 2 read one aligned register window (5 / 7 vectors for F = 12) that covers every tap of the four outputs, and the
// coefficient state moves as 16-byte global accesses.  Every output is still accumulated tap by tap in ascending t,
// so results are bit-identical to the scalar kernels.  Requirements: N % 4 == 0, pitches % 4 == 0, 16-byte aligned
// bases, F in the unrolled set; anything else runs the scalar kernels.  All rims are rounded up to multiples of four
// (halo H4 = sum over levels of round_up(step F/2, 4)).
constexpr int WAV4_THREADS = 256;

__host__ __device__ constexpr int round_up4(int v) { return (v + 3) & ~3; }

static int fused_halo4(const csr_wavelet* w) {
    int h = 0;
    for (int lev = 1; lev <= w->levels; ++lev) h += round_up4((w->F / 2) << (lev - 1));
    return h;
}
static FusedTiling fused_tiling4(const csr_wavelet* w, int nbuf) {
    FusedTiling t;
    t.H = fused_halo4(w);
    const int bmax = WAV_SMEM_BUDGET / (4 * nbuf);
    const int tmax = bmax - 2 * t.H;
    if (tmax < 256) {
        t.T = t.B = t.ntiles = 0;
        return t;
    }
    t.ntiles = (w->N + tmax - 1) / tmax;
    t.T = round_up((w->N + t.ntiles - 1) / t.ntiles, 32);
    t.B = t.T + 2 * t.H;
    return t;
}
static bool vec4_filter(int F) { return F == 2 || F == 4 || F == 6 || F == 8 || F == 12; }
static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// taps of four consecutive outputs i .. i+3.  OFF = (tap 0 position - output position) / step: F/2 for the analysis,
// F/2 - 1 for the synthesis; tap t sits at  i + j + (OFF - t) step.  SMALL = step (1 or 2) or 0 for step % 4 == 0.
template <int FLEN, int OFF, int SMALL, bool TWO>
__device__ __forceinline__ void taps4(const float* __restrict__ a, int i, int step, const float (&f0)[64],
                                      const float (&f1)[64], float (&c0)[4], float (&c1)[4]) {
    // packed fp32 FMAs (FFMA2, sm_100): two independent round-to-nearest FMAs per issue slot -- same bits as two FFMAs
    float2 p0[2] = {make_float2(c0[0], c0[1]), make_float2(c0[2], c0[3])};
    float2 p1[2] = {make_float2(c1[0], c1[1]), make_float2(c1[2], c1[3])};
    if (SMALL == 0) {
        const float* ap = a + i + OFF * step;
#pragma unroll
        for (int t = 0; t < FLEN; ++t) {
            const float4 v = *reinterpret_cast<const float4*>(ap - t * step);
            const float2 g0 = make_float2(f0[t], f0[t]);
            p0[0] = __ffma2_rn(g0, make_float2(v.x, v.y), p0[0]);
            p0[1] = __ffma2_rn(g0, make_float2(v.z, v.w), p0[1]);
            if (TWO) {
                const float2 g1 = make_float2(f1[t], f1[t]);
                p1[0] = __ffma2_rn(g1, make_float2(v.x, v.y), p1[0]);
                p1[1] = __ffma2_rn(g1, make_float2(v.z, v.w), p1[1]);
            }
        }
    } else {
        constexpr int LEFT = round_up4((FLEN - 1 - OFF) * SMALL);  // lowest tap: i + (OFF - (FLEN-1)) step
        constexpr int RIGHT = round_up4(OFF * SMALL);              // highest tap: i + 3 + OFF step
        constexpr int CNT = (LEFT + 4 + RIGHT) / 4;
        float w[4 * CNT];
#pragma unroll
        for (int c = 0; c < CNT; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(a + i - LEFT + 4 * c);
            w[4 * c] = v.x;
            w[4 * c + 1] = v.y;
            w[4 * c + 2] = v.z;
            w[4 * c + 3] = v.w;
        }
#pragma unroll
        for (int t = 0; t < FLEN; ++t) {
            const float2 g0 = make_float2(f0[t], f0[t]), g1 = make_float2(f1[t], f1[t]);
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float2 v = make_float2(w[2 * j + (OFF - t) * SMALL + LEFT], w[2 * j + 1 + (OFF - t) * SMALL + LEFT]);
                p0[j] = __ffma2_rn(g0, v, p0[j]);
                if (TWO) p1[j] = __ffma2_rn(g1, v, p1[j]);
            }
        }
    }
    c0[0] = p0[0].x;
    c0[1] = p0[0].y;
    c0[2] = p0[1].x;
    c0[3] = p0[1].y;
    if (TWO) {
        c1[0] = p1[0].x;
        c1[1] = p1[0].y;
        c1[2] = p1[1].x;
        c1[3] = p1[1].y;
    }
}

__device__ __forceinline__ bool nonzero4(const float4& v) { return v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f; }

// sfl: this band's state flags of the tile's blocks in shared memory (index (g >> kCoefBlockShift) - blk0), or null;
// fout: this band's row of the flags being written, or null.  mode / el1 / rfl / rout: wavelet-domain screening
// (CoefUpdate::screen_mode; el1 = band_l1 * eps[p]; rfl = redo marks in shared memory; rout = redo marks being written)
template <int FLEN, bool UPDATE, int SMALL>
__device__ __forceinline__ void analysis_level4(const float* __restrict__ a, float* __restrict__ o, int B, int step, int H,
                                                int T, int t0, int N, const FilterPair& f, float* __restrict__ cband,
                                                float* __restrict__ xband, float* __restrict__ zband, float ustep,
                                                float thr, float beta, float& dsum, const unsigned char* sfl, int blk0,
                                                unsigned char* __restrict__ fout, int mode, float el1,
                                                const unsigned char* rfl, unsigned char* __restrict__ rout) {
    const int rim = round_up4(step * (FLEN / 2));
    const int vlo = H, vhi = min(H + T, H + (N - t0));
    for (int i = rim + 4 * threadIdx.x; i < B - rim; i += 4 * WAV4_THREADS) {
        const bool valid = i >= vlo && i < vhi;
        const int g = t0 + (i - H);
        float4 zo = make_float4(0.f, 0.f, 0.f, 0.f), xo = zo;
        bool st = true;  // "the state may hold non-zeros here"
        bool update = UPDATE && valid;
        if (UPDATE && valid) {  // state loads in flight while the filters run
            if (sfl) st = sfl[(g >> kCoefBlockShift) - blk0] != 0;
            if (mode == 2) {  // redo pass: only zero-state blocks that failed the screening test
                update = rfl[(g >> kCoefBlockShift) - blk0] != 0;
                st = false;
            }
            if (st) {
                zo = __ldcg(reinterpret_cast<const float4*>(zband + g));
                xo = __ldcg(reinterpret_cast<const float4*>(xband + g));
            }
        }
        float ca[4] = {0.f, 0.f, 0.f, 0.f}, cd[4] = {0.f, 0.f, 0.f, 0.f};
        taps4<FLEN, FLEN / 2, SMALL, true>(a, i, step, f.lo, f.hi, ca, cd);
        *reinterpret_cast<float4*>(o + i) = make_float4(ca[0], ca[1], ca[2], ca[3]);
        if (valid) {
            if (UPDATE) {
                if (mode == 1 && !st) {  // zero state, gradient possibly inexact: provably below the threshold?
                    const float mx = fmaxf(fmaxf(fabsf(cd[0]), fabsf(cd[1])), fmaxf(fabsf(cd[2]), fabsf(cd[3])));
                    if (!(ustep * (mx + el1) * 1.000001f < thr)) rout[g >> kCoefBlockShift] = 1;
                    update = false;
                }
                if (update) {
                    float4 xn, zn;
                    xn.x = soft_update(zo.x, xo.x, cd[0], ustep, thr, beta, zn.x, dsum);
                    xn.y = soft_update(zo.y, xo.y, cd[1], ustep, thr, beta, zn.y, dsum);
                    xn.z = soft_update(zo.z, xo.z, cd[2], ustep, thr, beta, zn.z, dsum);
                    xn.w = soft_update(zo.w, xo.w, cd[3], ustep, thr, beta, zn.w, dsum);
                    const bool nzr = nonzero4(xn) || nonzero4(zn);
                    if (st || nzr) {  // a zero block whose new values are zero again keeps its (zero) memory
                        __stcs(reinterpret_cast<float4*>(xband + g), xn);
                        __stcs(reinterpret_cast<float4*>(zband + g), zn);
                    }
                    if (nzr && fout) fout[g >> kCoefBlockShift] = 1;
                }
            } else {
                *reinterpret_cast<float4*>(cband + g) = make_float4(cd[0], cd[1], cd[2], cd[3]);
            }
        }
    }
}

template <int FLEN, bool UPDATE>
__global__ void __launch_bounds__(WAV4_THREADS, 3)
swt_analysis_fused4_kernel(const float* __restrict__ signal, int ld_sig, int N, int L, int H, int T, int B,
                           const BandOffsets bands, const FilterPair f, float* __restrict__ coef, int ld_coef,
                           const CoefUpdate up) {
    extern __shared__ __align__(16) float wsm4[];
    constexpr int MAXBLK = WAV_SMEM_BUDGET / 8 / (1 << kCoefBlockShift) + 3;  // flag blocks a tile can touch
    __shared__ unsigned char sflags[UPDATE ? kCoefFlagBands * MAXBLK : 1];
    __shared__ unsigned char rflags[UPDATE ? kCoefFlagBands * MAXBLK : 1];
    float* a = wsm4;
    float* o = wsm4 + B;
    const int p = blockIdx.x;
    const int t0 = blockIdx.y * T;
    const int base = t0 - H;
    const int NB = (N + (1 << kCoefBlockShift) - 1) >> kCoefBlockShift;
    const int blk0 = t0 >> kCoefBlockShift;
    const int nblk = min(NB - 1, (t0 + T - 1) >> kCoefBlockShift) - blk0 + 1;
    const bool use_flags = UPDATE && up.flags_in != nullptr;
    const int mode = UPDATE ? up.screen_mode : 0;
    float thr = 0.f, eps = 0.f;
    if (UPDATE) {
        const float l0 = up.lambda0[p], nz = up.noise[p];
        const float lam = __fmaf_rn(-(float)up.iter, nz, l0);
        const bool active = (nz < l0) && (up.iter == 0 || lam > 0.f) && (!up.iter_cap || up.iter < up.iter_cap[p]);
        thr = __fmul_rn(up.step, lam);  // (explicit rounding: the compiler would contract this product into |u| - thr)
        if (!active) {  // whole CTA: this line of sight is frozen; its state flags carry over
            if (use_flags && mode != 2) {
                for (int idx = threadIdx.x; idx < (L + 1) * nblk; idx += WAV4_THREADS) {
                    const size_t at = (size_t)p * up.flag_ld + (idx / nblk) * NB + blk0 + idx % nblk;
                    if (up.flags_in[at]) up.flags_out[at] = 1;
                }
            }
            return;
        }
        if (use_flags) {
            int any = 0;
            for (int idx = threadIdx.x; idx < (L + 1) * nblk; idx += WAV4_THREADS) {
                const size_t at = (size_t)p * up.flag_ld + (idx / nblk) * NB + blk0 + idx % nblk;
                sflags[(idx / nblk) * MAXBLK + idx % nblk] = up.flags_in[at];
                if (mode == 2) any |= (rflags[(idx / nblk) * MAXBLK + idx % nblk] = up.redo[at]);
            }
            if (mode == 2 && !__syncthreads_or(any)) return;  // nothing of this tile failed the screening test
        }
        if (mode == 1) eps = up.eps[p];
    }
    const float* srow = signal + (size_t)p * ld_sig;
    for (int i = 4 * threadIdx.x; i < B; i += 4 * WAV4_THREADS)
        *reinterpret_cast<float4*>(a + i) = *reinterpret_cast<const float4*>(srow + wrap_index(base + i, N));
    __syncthreads();
    const size_t crow = (size_t)p * ld_coef;
    float dsum = 0.f;
    for (int lev = 1; lev <= L; ++lev) {
        const int step = 1 << (lev - 1);
        const int bi = L - lev + 1;
        const int band = bands.off[bi];
        float* cband = UPDATE ? nullptr : coef + crow + band;
        float* xband = UPDATE ? up.x + crow + band : nullptr;
        float* zband = UPDATE ? up.z + crow + band : nullptr;
        const unsigned char* sfl = use_flags ? sflags + bi * MAXBLK : nullptr;
        const unsigned char* rfl = use_flags ? rflags + bi * MAXBLK : nullptr;
        unsigned char* fout = use_flags ? up.flags_out + (size_t)p * up.flag_ld + bi * NB : nullptr;
        unsigned char* rout = (use_flags && mode == 1) ? up.redo + (size_t)p * up.flag_ld + bi * NB : nullptr;
        const float el1 = mode == 1 ? up.band_l1[bi] * eps : 0.f;
        if (lev == 1)
            analysis_level4<FLEN, UPDATE, 1>(a, o, B, step, H, T, t0, N, f, cband, xband, zband, up.step, thr, up.beta, dsum, sfl, blk0, fout, mode, el1, rfl, rout);
        else if (lev == 2)
            analysis_level4<FLEN, UPDATE, 2>(a, o, B, step, H, T, t0, N, f, cband, xband, zband, up.step, thr, up.beta, dsum, sfl, blk0, fout, mode, el1, rfl, rout);
        else
            analysis_level4<FLEN, UPDATE, 0>(a, o, B, step, H, T, t0, N, f, cband, xband, zband, up.step, thr, up.beta, dsum, sfl, blk0, fout, mode, el1, rfl, rout);
        __syncthreads();
        float* tmp = a;
        a = o;
        o = tmp;
    }
    const int vhi = min(H + T, H + (N - t0));
    for (int i = H + 4 * threadIdx.x; i < vhi; i += 4 * WAV4_THREADS) {  // approximation band cA_L
        const int g = t0 + (i - H);
        const float4 ca = *reinterpret_cast<const float4*>(a + i);
        if (UPDATE) {
            float* zb = up.z + crow + bands.off[0] + g;
            float* xb = up.x + crow + bands.off[0] + g;
            const int blk = (g >> kCoefBlockShift) - blk0;
            bool st = use_flags ? sflags[blk] != 0 : true;
            bool update = true;
            if (mode == 2) {
                update = rflags[blk] != 0;
                st = false;
            }
            if (mode == 1 && !st) {
                const float mx = fmaxf(fmaxf(fabsf(ca.x), fabsf(ca.y)), fmaxf(fabsf(ca.z), fabsf(ca.w)));
                if (!(up.step * (mx + up.band_l1[0] * eps) * 1.000001f < thr)) up.redo[(size_t)p * up.flag_ld + (g >> kCoefBlockShift)] = 1;
                update = false;
            }
            if (!update) continue;
            float4 zo = make_float4(0.f, 0.f, 0.f, 0.f), xo = zo;
            if (st) {
                zo = __ldcg(reinterpret_cast<const float4*>(zb));
                xo = __ldcg(reinterpret_cast<const float4*>(xb));
            }
            float4 xn, zn;
            xn.x = soft_update(zo.x, xo.x, ca.x, up.step, thr, up.beta, zn.x, dsum);
            xn.y = soft_update(zo.y, xo.y, ca.y, up.step, thr, up.beta, zn.y, dsum);
            xn.z = soft_update(zo.z, xo.z, ca.z, up.step, thr, up.beta, zn.z, dsum);
            xn.w = soft_update(zo.w, xo.w, ca.w, up.step, thr, up.beta, zn.w, dsum);
            const bool nzr = nonzero4(xn) || nonzero4(zn);
            if (st || nzr) {
                __stcs(reinterpret_cast<float4*>(xb), xn);
                __stcs(reinterpret_cast<float4*>(zb), zn);
            }
            if (nzr && use_flags) up.flags_out[(size_t)p * up.flag_ld + (g >> kCoefBlockShift)] = 1;
        } else {
            *reinterpret_cast<float4*>(coef + crow + bands.off[0] + g) = ca;
        }
    }
    if (UPDATE && up.delta != nullptr) {
#pragma unroll
        for (int sft = 16; sft > 0; sft >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, sft);
        if ((threadIdx.x & 31) == 0 && dsum != 0.f) atomicAdd(up.delta + p, dsum);
    }
}

// adjoint tiles (128 lines of sight x 256 phi columns) whose gradient a set coefficient flag can see: every coefficient
// of block [bs, be) of the level-l band depends on the signal inside [bs - H_l, be + H_l) only, H_l = (F/2)(2^l - 1)
__global__ void need_from_flags_kernel(const unsigned char* __restrict__ flags, int flag_ld, int N, int NB, int L, int halfF,
                                       unsigned char* __restrict__ need, int need_ld) {
    const int p = blockIdx.x;
    unsigned char* row = need + (size_t)(p >> 7) * need_ld;
    for (int idx = threadIdx.x; idx < (L + 1) * NB; idx += blockDim.x) {
        if (!flags[(size_t)p * flag_ld + idx]) continue;
        const int k = idx % NB, band = idx / NB;
        const int lev = band == 0 ? L : L - band + 1;   // bands [cA_L, cD_L, ..., cD_1]
        const int H = halfF * ((1 << lev) - 1);         // a level-l coefficient sees the signal within (F/2)(2^l - 1) of it
        const int bs = k << kCoefBlockShift, be = min(N, bs + (1 << kCoefBlockShift));
        const int len = be - bs + 2 * H;
        if (len >= N) {
            for (int t = 0; t <= (N - 1) >> 8; ++t) row[t] = 1;
        } else {
            const int lo = wrap_index(bs - H, N), hi = lo + len;
            for (int t = lo >> 8; t <= (min(hi, N) - 1) >> 8; ++t) row[t] = 1;
            if (hi > N)
                for (int t = 0; t <= (hi - N - 1) >> 8; ++t) row[t] = 1;
        }
    }
}

// state flags of a coefficient matrix: byte [p, band * NB + k] = "block k of band `band` holds a non-zero"
__global__ void coef_flags_kernel(const float* __restrict__ z, int ld_coef, int N, int NB, const BandOffsets bands,
                                  unsigned char* __restrict__ flags, int flag_ld) {
    const int p = blockIdx.x, band = blockIdx.y / NB, k = blockIdx.y % NB;
    const int s0 = k << kCoefBlockShift, s1 = min(N, s0 + (1 << kCoefBlockShift));
    const float* row = z + (size_t)p * ld_coef + bands.off[band];
    int nz = 0;
    for (int i = s0 + 4 * threadIdx.x; i < s1; i += 4 * blockDim.x) nz |= nonzero4(*reinterpret_cast<const float4*>(row + i));
    nz = __syncthreads_or(nz);
    if (threadIdx.x == 0) flags[(size_t)p * flag_ld + band * NB + k] = nz ? 1 : 0;
}

template <int FLEN, int SMALL, bool USE_A, bool USE_D>
__device__ __forceinline__ void synth_level4(const float* __restrict__ a, const float* __restrict__ d,
                                             float* __restrict__ o, int B, int step, const FilterPair& f) {
    const int rim = round_up4(step * (FLEN / 2));
    for (int i = rim + 4 * threadIdx.x; i < B - rim; i += 4 * WAV4_THREADS) {
        float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
        if (USE_A) taps4<FLEN, FLEN / 2 - 1, SMALL, false>(a, i, step, f.lo, f.lo, acc0, acc0);
        if (USE_D) taps4<FLEN, FLEN / 2 - 1, SMALL, false>(d, i, step, f.hi, f.hi, acc1, acc1);
        *reinterpret_cast<float4*>(o + i) = make_float4(0.5f * (acc0[0] + acc1[0]), 0.5f * (acc0[1] + acc1[1]),
                                                        0.5f * (acc0[2] + acc1[2]), 0.5f * (acc0[3] + acc1[3]));
    }
}

template <int FLEN, int SMALL>
__device__ __forceinline__ void synth_level4_any(const float* a, const float* d, float* o, int B, int step,
                                                 const FilterPair& f, bool a_nz, bool d_nz) {
    if (a_nz && d_nz) synth_level4<FLEN, SMALL, true, true>(a, d, o, B, step, f);
    else if (a_nz) synth_level4<FLEN, SMALL, true, false>(a, d, o, B, step, f);
    else synth_level4<FLEN, SMALL, false, true>(a, d, o, B, step, f);
}

template <int FLEN>
__global__ void __launch_bounds__(WAV4_THREADS, 3)
swt_synth_fused4_kernel(const float* __restrict__ coef, int ld_coef, int N, int L, int H, int T, int B,
                        const BandOffsets bands, const FilterPair f, float* __restrict__ sig, int ld_sig,
                        __half* __restrict__ hi, __half* __restrict__ lo, int ld_h, const float* __restrict__ scale,
                        unsigned char* __restrict__ tile_flags, int flag_ld, const unsigned char* __restrict__ cflags,
                        int cflag_ld) {
    extern __shared__ __align__(16) float wsm4[];
    float* a = wsm4;
    float* o = wsm4 + B;
    float* d = wsm4 + 2 * B;
    const int p = blockIdx.x;
    const int t0 = blockIdx.y * T;
    const int base = t0 - H;
    const float* crow = coef + (size_t)p * ld_coef;
    // which band tiles hold anything: from the coefficient-state flags when the caller maintains them (a zero flag
    // guarantees zeros; a set flag over zeros only costs work), else by one pass of independent 16-byte loads
    __shared__ unsigned int band_nz_s;
    if (threadIdx.x == 0) band_nz_s = 0u;
    __syncthreads();
    if (cflags != nullptr) {
        const int NB = (N + (1 << kCoefBlockShift) - 1) >> kCoefBlockShift;
        const int ws = wrap_index(base, N);  // the buffer covers [ws, ws + B) modulo N
        unsigned int mine = 0u;
        for (int idx = threadIdx.x; idx < (L + 1) * NB; idx += WAV4_THREADS) {
            const int b = idx / NB, k = idx - b * NB;
            const int bs = k << kCoefBlockShift, be = min(N, bs + (1 << kCoefBlockShift));
            const bool ov = B >= N || (ws + B <= N ? (bs < ws + B && be > ws) : (be > ws || bs < ws + B - N));
            if (ov && cflags[(size_t)p * cflag_ld + idx]) mine |= 1u << b;
        }
        mine = __reduce_or_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0 && mine) atomicOr(&band_nz_s, mine);
    } else {
        unsigned int mine = 0u;
        for (int i = 4 * threadIdx.x; i < B; i += 4 * WAV4_THREADS) {
            const int gi = wrap_index(base + i, N);
            for (int b = 0; b <= L; ++b) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(crow + bands.off[b] + gi));
                mine |= nonzero4(v) ? (1u << b) : 0u;
            }
        }
        mine = __reduce_or_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31) == 0 && mine) atomicOr(&band_nz_s, mine);
    }
    __syncthreads();
    const unsigned int band_nz = band_nz_s;
    bool a_nz = (band_nz & 1u) != 0;
    if (a_nz) {
        for (int i = 4 * threadIdx.x; i < B; i += 4 * WAV4_THREADS)
            *reinterpret_cast<float4*>(a + i) = __ldg(reinterpret_cast<const float4*>(crow + bands.off[0] + wrap_index(base + i, N)));
    }
    for (int lev = L; lev >= 1; --lev) {
        const int step = 1 << (lev - 1);
        const float* dband = crow + bands.off[L - lev + 1];
        const bool d_nz = ((band_nz >> (L - lev + 1)) & 1u) != 0;
        if (d_nz) {
            for (int i = 4 * threadIdx.x; i < B; i += 4 * WAV4_THREADS)
                *reinterpret_cast<float4*>(d + i) = __ldg(reinterpret_cast<const float4*>(dband + wrap_index(base + i, N)));
        }
        if (a_nz || d_nz) __syncthreads();
        if (!a_nz && !d_nz) continue;
        if (lev == 1) synth_level4_any<FLEN, 1>(a, d, o, B, step, f, a_nz, d_nz);
        else if (lev == 2) synth_level4_any<FLEN, 2>(a, d, o, B, step, f, a_nz, d_nz);
        else synth_level4_any<FLEN, 0>(a, d, o, B, step, f, a_nz, d_nz);
        a_nz = true;
        __syncthreads();
        float* tmp = a;
        a = o;
        o = tmp;
    }
    const float sc = scale ? scale[p] : 1.f;
    const int vhi = min(H + T, H + (N - t0));
    for (int i = H + 4 * threadIdx.x; i < vhi; i += 4 * WAV4_THREADS) {
        const int g = t0 + (i - H);
        const float4 v = a_nz ? *reinterpret_cast<const float4*>(a + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (sig) *reinterpret_cast<float4*>(sig + (size_t)p * ld_sig + g) = v;
        if (hi) {
            __align__(8) __half h[4];
            __align__(8) __half l[4];
            split_half(v.x * sc, h[0], l[0]);
            split_half(v.y * sc, h[1], l[1]);
            split_half(v.z * sc, h[2], l[2]);
            split_half(v.w * sc, h[3], l[3]);
            *reinterpret_cast<uint2*>(hi + (size_t)p * ld_h + g) = *reinterpret_cast<const uint2*>(h);
            *reinterpret_cast<uint2*>(lo + (size_t)p * ld_h + g) = *reinterpret_cast<const uint2*>(l);
            if (tile_flags && (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f))
                tile_flags[((size_t)(p >> 7) * flag_ld + (g >> 7)) * 4 + ((p >> 5) & 3)] = 1;
        }
    }
}

static BandOffsets band_offsets(const csr_wavelet* w) {
    BandOffsets b;
    memset(&b, 0, sizeof(b));
    for (int i = 0; i <= w->levels; ++i) b.off[i] = w->offsets[i];
    return b;
}

template <typename K>
static int set_wavelet_smem(K kernel, int bytes) {
    CSR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return CSR_OK;
}

bool profile_start(cudaStream_t stream, cudaEvent_t* start, cudaEvent_t* stop);
void profile_stop(int kind, cudaStream_t stream, cudaEvent_t start, cudaEvent_t stop);

int coef_flag_ld(const csr_wavelet* w) {
    const int NB = (w->N + (1 << kCoefBlockShift) - 1) >> kCoefBlockShift;
    return round_up((w->levels + 1) * NB, 16);
}
// the fused loop may keep coefficient-state flags iff both fused kernels will run their four-sample versions
bool coef_flags_usable(const csr_wavelet* w, const float* x, const float* z, int ld_coef) {
    return options().wavelet_vec4 != 0 && wavelet_fusable(w) && vec4_filter(w->F) && fused_tiling4(w, 3).T > 0 &&
           fused_tiling4(w, 2).T > 0 && w->N % 4 == 0 && ld_coef % 4 == 0 && aligned16(x) && aligned16(z) &&
           w->levels + 1 <= kCoefFlagBands;
}
int launch_coef_flags(const csr_wavelet* w, const float* z, int ld_coef, int P, unsigned char* flags, int flag_ld,
                      cudaStream_t stream) {
    const int NB = (w->N + (1 << kCoefBlockShift) - 1) >> kCoefBlockShift;
    coef_flags_kernel<<<dim3(P, (w->levels + 1) * NB), 128, 0, stream>>>(z, ld_coef, w->N, NB, band_offsets(w), flags, flag_ld);
    CSR_LAUNCHED();
    return CSR_OK;
}

// tiles of the adjoint GEMM that the set flags depend on (need: [ceil(P/128), need_ld] bytes, zeroed by the caller)
int launch_need_from_flags(const csr_wavelet* w, const unsigned char* flags, int flag_ld, int P, unsigned char* need,
                           int need_ld, cudaStream_t stream) {
    const int NB = (w->N + (1 << kCoefBlockShift) - 1) >> kCoefBlockShift;
    need_from_flags_kernel<<<P, 128, 0, stream>>>(flags, flag_ld, w->N, NB, w->levels, w->F / 2, need, need_ld);
    CSR_LAUNCHED();
    return CSR_OK;
}
// l1 norms of the cascaded analysis filters, band order [cA_L, cD_L, ..., cD_1], rounded up
void wavelet_band_l1(const csr_wavelet* w, float* l1) {
    const int F = w->F, L = w->levels;
    std::vector<double> lo(1, 1.0);
    for (int lev = 1; lev <= L; ++lev) {
        const int step = 1 << (lev - 1);
        std::vector<double> a(lo.size() + (size_t)(F - 1) * step, 0.0), d(a.size(), 0.0);
        for (size_t i = 0; i < lo.size(); ++i)
            for (int t = 0; t < F; ++t) {
                a[i + (size_t)t * step] += lo[i] * (double)w->dec_lo[t];
                d[i + (size_t)t * step] += lo[i] * (double)w->dec_hi[t];
            }
        double sd = 0.0, sa = 0.0;
        for (double v : d) sd += fabs(v);
        for (double v : a) sa += fabs(v);
        l1[L - lev + 1] = (float)(sd * 1.0001);
        if (lev == L) l1[0] = (float)(sa * 1.0001);
        lo.swap(a);
    }
}

#define CSR_DISPATCH_FLEN4(FVAL, CALL) \
    switch (FVAL) {                    \
        case 2: { constexpr int FL = 2; CALL; } break;   \
        case 4: { constexpr int FL = 4; CALL; } break;   \
        case 6: { constexpr int FL = 6; CALL; } break;   \
        case 8: { constexpr int FL = 8; CALL; } break;   \
        default: { constexpr int FL = 12; CALL; } break; \
    }

int swt_synth_fused(const csr_wavelet* w, const float* coef, int ld_coef, float* sig, int ld_sig, __half* hi,
                    __half* lo, int ld_h, const float* scale, int P, unsigned int* tile_flags, int flag_ld,
                    const unsigned char* coef_flags, int coef_flag_ld, cudaStream_t stream) {
    const FusedTiling t4 = fused_tiling4(w, 3);
    const bool vec4 = options().wavelet_vec4 != 0 && vec4_filter(w->F) && t4.T > 0 && w->N % 4 == 0 && ld_coef % 4 == 0 &&
                      aligned16(coef) && (!sig || (ld_sig % 4 == 0 && aligned16(sig))) &&
                      (!hi || (ld_h % 4 == 0 && ((uintptr_t)hi & 7) == 0 && ((uintptr_t)lo & 7) == 0));
    const FusedTiling t = vec4 ? t4 : fused_tiling(w, 3);
    const int smem = 3 * t.B * 4;
    CSR_REQUIRE(vec4 || coef_flags == nullptr, "swt_synth_fused: coefficient flags need the four-sample kernels");
    static bool configured = false;
    if (!configured) {
        int rc = set_wavelet_smem(swt_synth_fused_kernel, WAV_SMEM_BUDGET + 1024);
        CSR_DISPATCH_FLEN4(2, if (!rc) rc = set_wavelet_smem(swt_synth_fused4_kernel<FL>, WAV_SMEM_BUDGET + 1024))
        CSR_DISPATCH_FLEN4(4, if (!rc) rc = set_wavelet_smem(swt_synth_fused4_kernel<FL>, WAV_SMEM_BUDGET + 1024))
        CSR_DISPATCH_FLEN4(6, if (!rc) rc = set_wavelet_smem(swt_synth_fused4_kernel<FL>, WAV_SMEM_BUDGET + 1024))
        CSR_DISPATCH_FLEN4(8, if (!rc) rc = set_wavelet_smem(swt_synth_fused4_kernel<FL>, WAV_SMEM_BUDGET + 1024))
        CSR_DISPATCH_FLEN4(12, if (!rc) rc = set_wavelet_smem(swt_synth_fused4_kernel<FL>, WAV_SMEM_BUDGET + 1024))
        if (rc) return rc;
        configured = true;
    }
    cudaEvent_t ev0, ev1;
    const bool prof = profile_start(stream, &ev0, &ev1);
    if (vec4) {
        CSR_DISPATCH_FLEN4(w->F, (swt_synth_fused4_kernel<FL><<<dim3(P, t.ntiles), WAV4_THREADS, smem, stream>>>(
            coef, ld_coef, w->N, w->levels, t.H, t.T, t.B, band_offsets(w), make_pair(w->rec_lo, w->rec_hi, w->F), sig, ld_sig,
            hi, lo, ld_h, scale, reinterpret_cast<unsigned char*>(tile_flags), flag_ld, coef_flags, coef_flag_ld)))
    } else {
        swt_synth_fused_kernel<<<dim3(P, t.ntiles), WAV_THREADS, smem, stream>>>(
            coef, ld_coef, w->N, w->levels, t.H, t.T, t.B, band_offsets(w), make_pair(w->rec_lo, w->rec_hi, w->F), sig, ld_sig,
            hi, lo, ld_h, scale, reinterpret_cast<unsigned char*>(tile_flags), flag_ld);
    }
    CSR_LAUNCHED();
    if (prof) profile_stop(6, stream, ev0, ev1);
    return CSR_OK;
}

int swt_analysis_fused(const csr_wavelet* w, const float* signal, int ld_sig, float* coef, int ld_coef, int P,
                       const CoefUpdate* update, cudaStream_t stream) {
    const FusedTiling t4 = fused_tiling4(w, 2);
    const bool vec4 = options().wavelet_vec4 != 0 && vec4_filter(w->F) && t4.T > 0 && w->N % 4 == 0 && ld_coef % 4 == 0 &&
                      ld_sig % 4 == 0 && aligned16(signal) &&
                      (update ? (aligned16(update->x) && aligned16(update->z)) : aligned16(coef));
    const FusedTiling t = vec4 ? t4 : fused_tiling(w, 2);
    const int smem = 2 * t.B * 4;
    CSR_REQUIRE(vec4 || !update || update->flags_in == nullptr, "swt_analysis_fused: coefficient flags need the four-sample kernels");
    static bool configured = false;
    if (!configured) {
        int rc = set_wavelet_smem(swt_analysis_fused_kernel<true>, WAV_SMEM_BUDGET + 1024);
        if (!rc) rc = set_wavelet_smem(swt_analysis_fused_kernel<false>, WAV_SMEM_BUDGET + 1024);
#define CSR_SET4(FV)                                                                                                   \
    CSR_DISPATCH_FLEN4(FV, if (!rc) rc = set_wavelet_smem(swt_analysis_fused4_kernel<FL, true>, WAV_SMEM_BUDGET + 1024);  \
                       if (!rc) rc = set_wavelet_smem(swt_analysis_fused4_kernel<FL, false>, WAV_SMEM_BUDGET + 1024))
        CSR_SET4(2) CSR_SET4(4) CSR_SET4(6) CSR_SET4(8) CSR_SET4(12)
#undef CSR_SET4
        if (rc) return rc;
        configured = true;
    }
    const dim3 grid(P, t.ntiles);
    const FilterPair f = make_pair(w->dec_lo, w->dec_hi, w->F);
    cudaEvent_t ev0, ev1;
    const bool prof = update && profile_start(stream, &ev0, &ev1);
    CoefUpdate none;
    memset(&none, 0, sizeof(none));
    if (vec4) {
        if (update) {
            CSR_DISPATCH_FLEN4(w->F, (swt_analysis_fused4_kernel<FL, true><<<grid, WAV4_THREADS, smem, stream>>>(
                signal, ld_sig, w->N, w->levels, t.H, t.T, t.B, band_offsets(w), f, coef, ld_coef, *update)))
        } else {
            CSR_DISPATCH_FLEN4(w->F, (swt_analysis_fused4_kernel<FL, false><<<grid, WAV4_THREADS, smem, stream>>>(
                signal, ld_sig, w->N, w->levels, t.H, t.T, t.B, band_offsets(w), f, coef, ld_coef, none)))
        }
    } else if (update) {
        swt_analysis_fused_kernel<true><<<grid, WAV_THREADS, smem, stream>>>(signal, ld_sig, w->N, w->levels, t.H, t.T, t.B,
                                                                            band_offsets(w), f, coef, ld_coef, *update);
    } else {
        swt_analysis_fused_kernel<false><<<grid, WAV_THREADS, smem, stream>>>(signal, ld_sig, w->N, w->levels, t.H, t.T, t.B,
                                                                             band_offsets(w), f, coef, ld_coef, none);
    }
    CSR_LAUNCHED();
    if (prof) profile_stop(update->screen_mode == 2 ? 11 : 7, stream, ev0, ev1);
    return CSR_OK;
}

// coefficient layout: [signal (N) if append_signal][cA_L][cD_L]...[cD_1]
// the fused decimated kernels: all levels of a line of sight in shared memory (N + N/2 floats must fit)
static DwtBands dwt_bands(const csr_wavelet* w) {
    DwtBands b = {};
    b.L = w->levels;
    b.N = w->N;
    b.ext = w->ext;
    b.base = w->append_signal ? w->N : 0;
    for (int i = 0; i <= w->levels; ++i) {
        b.sizes[i] = w->sizes[i];
        b.offsets[i] = w->offsets[i];
    }
    return b;
}
static size_t dwt_fused_smem(const csr_wavelet* w, bool analysis) {
    const size_t n1 = (size_t)((w->sizes[w->levels] + 3) & ~3);
    return sizeof(float) * (analysis ? (size_t)((w->N + 3) & ~3) + n1 : 2 * n1);
}
static bool dwt_fusable(const csr_wavelet* w) {
    return w->kind == 1 && w->levels >= 1 && options().fused_wavelet != 0 && dwt_fused_smem(w, true) <= 200 * 1024;
}

int wavelet_decompose(const csr_wavelet* w, const float* signal, int ld_sig, float* coef, int ld_coef, int P,
                      float* ws, cudaStream_t stream) {
    if (wavelet_fusable(w) && options().fused_wavelet != 0)
        return swt_analysis_fused(w, signal, ld_sig, coef, ld_coef, P, nullptr, stream);
    if (dwt_fusable(w)) {
        const size_t smem = dwt_fused_smem(w, true);
        CSR_CUDA(cudaFuncSetAttribute(dwt_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dwt_fused_kernel<<<P, DWT_FUSED_THREADS, smem, stream>>>(signal, ld_sig, coef, ld_coef, dwt_bands(w),
                                                                 make_pair(w->dec_lo, w->dec_hi, w->F), w->append_signal);
        CSR_LAUNCHED();
        return CSR_OK;
    }
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
        } else if (w->ext == 0) {
            dwt_level_kernel<<<grid_for((len + 1) / 2, P), 256, 0, stream>>>(cur, ld_cur, len, a_out, ld_a, d_out,
                                                                             ld_coef, f);
            len = (len + 1) / 2;
        } else {
            dwt_ext_level_kernel<<<grid_for((len + w->F - 1) / 2, P), 256, 0, stream>>>(cur, ld_cur, len, w->ext, a_out,
                                                                                        ld_a, d_out, ld_coef, f);
            len = (len + w->F - 1) / 2;
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
    if (wavelet_fusable(w) && options().fused_wavelet != 0)
        return swt_synth_fused(w, coef, ld_coef, signal, ld_sig, nullptr, nullptr, 0, nullptr, P, nullptr, 0, nullptr, 0, stream);
    if (dwt_fusable(w)) {
        const size_t smem = dwt_fused_smem(w, false);
        CSR_CUDA(cudaFuncSetAttribute(idwt_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        idwt_fused_kernel<<<P, DWT_FUSED_THREADS, smem, stream>>>(coef, ld_coef, signal, ld_sig, dwt_bands(w),
                                                                  make_pair(w->rec_lo, w->rec_hi, w->F), w->append_signal);
        CSR_LAUNCHED();
        return CSR_OK;
    }
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
            if (w->ext == 0)
                idwt_level_kernel<<<grid_for(keep, P), 256, 0, stream>>>(cur, ld_cur, d_in, ld_coef, half, out, ld_out,
                                                                         keep, add, ld_coef, f);
            else
                idwt_ext_level_kernel<<<grid_for(keep, P), 256, 0, stream>>>(cur, ld_cur, d_in, ld_coef, half, out, ld_out,
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
    CSR_REQUIRE(kind >= 0 && kind <= EXT_COUNT,
                "csr_wavelet_create: kind must be 0 (undecimated), 1 (decimated, periodization) or 2..%d (decimated with "
                "signal extension zero, constant, symmetric, reflect, periodic, smooth, antisymmetric, antireflect)", (int)EXT_COUNT);
    const int ext = kind >= 1 ? kind - 1 : 0;
    if (kind > 1) kind = 1;
    CSR_REQUIRE(F >= 2 && F <= 64 && F % 2 == 0, "csr_wavelet_create: filter length must be even and in [2, 64] (F=%d)", F);
    CSR_REQUIRE(N > 0 && level >= 0 && level <= 32, "csr_wavelet_create: bad N=%d or level=%d", N, level);
    if (kind == 0) CSR_REQUIRE(level == 0 || N % (1 << level) == 0, "csr_wavelet_create: N=%d is not a multiple of 2^%d", N, level);
    *out = nullptr;
    csr_wavelet* w = new csr_wavelet();
    memset(w, 0, sizeof(*w));
    w->uid = next_uid();
    w->device = device;
    w->kind = kind;
    w->ext = ext;
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
            len = ext ? (len + F - 1) / 2 : (len + 1) / 2;
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
