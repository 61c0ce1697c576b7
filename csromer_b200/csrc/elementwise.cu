// elementwise.cu -- the HBM-bound helpers around the GEMMs: operand splitting, the per-line-of-sight diagonal
// scalings (K6), Galactic-RM de-rotation, the final chi2 / L1 evaluation, and the coefficient-space FISTA
// update used by the wavelet bases.  One CTA per line of sight (row), threads stride over the row, so every
// global access is a coalesced 128-byte line.
#include "plan.cuh"

namespace csr {

constexpr int EW_THREADS = 256;

__device__ __forceinline__ float block_reduce(float v, bool is_max) {
    __shared__ float red[EW_THREADS / 32];
    __shared__ float result;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const float o = __shfl_xor_sync(0xffffffffu, v, s);
        v = is_max ? fmaxf(v, o) : v + o;
    }
    __syncthreads();  // protects `red` / `result` against the previous call
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        float r = threadIdx.x < EW_THREADS / 32 ? red[threadIdx.x] : (is_max ? 0.f : 0.f);
#pragma unroll
        for (int s = 16; s > 0; s >>= 1) {
            const float o = __shfl_xor_sync(0xffffffffu, r, s);
            r = is_max ? fmaxf(r, o) : r + o;
        }
        if (threadIdx.x == 0) result = r;
    }
    __syncthreads();
    return result;
}

// ---- fp32 row -> scaled fp16 hi/lo operand -------------------------------------------------------------
// value = in[row, c] * chan_scale[(per_pixel ? row*m : 0) + (c mod m)]   (chan_scale optional).
// fixed_scale = 0: scale[row] <- power of two from max(|value|, |ref2[row, :]|) over the row
// fixed_scale = 1: scale[row] is read.
__global__ void __launch_bounds__(EW_THREADS)
split_rows_kernel(const float* __restrict__ in, int ld_in, int cols, const float* __restrict__ chan_scale,
                  int chan_pp, int m, const float* __restrict__ ref2, int ld_ref2, __half* __restrict__ hi,
                  __half* __restrict__ lo, int ld_out, float* __restrict__ scale, int fixed_scale) {
    const int row = blockIdx.x;
    const float* src = in + (size_t)row * ld_in;
    const float* cs = chan_scale ? chan_scale + (chan_pp ? (size_t)row * m : 0) : nullptr;
    float s;
    if (!fixed_scale) {
        float mx = 0.f;
        for (int c = threadIdx.x; c < cols; c += EW_THREADS) {
            float v = src[c];
            if (cs) v *= cs[c >= m ? c - m : c];
            mx = fmaxf(mx, fabsf(v));
        }
        if (ref2) {
            const float* r2 = ref2 + (size_t)row * ld_ref2;
            for (int c = threadIdx.x; c < cols; c += EW_THREADS) mx = fmaxf(mx, fabsf(r2[c]));
        }
        mx = block_reduce(mx, true);
        s = operand_scale(mx);
        if (threadIdx.x == 0) scale[row] = s;
    } else {
        s = scale[row];
    }
    __half* dh = hi + (size_t)row * ld_out;
    __half* dl = lo + (size_t)row * ld_out;
    for (int c = threadIdx.x; c < cols; c += EW_THREADS) {
        float v = src[c];
        if (cs) v *= cs[c >= m ? c - m : c];
        __half h, l;
        split_half(v * s, h, l);
        dh[c] = h;
        dl[c] = l;
    }
}

int launch_split_rows(const float* in, int ld_in, int rows, int cols, const float* chan_scale, int chan_pp, int m,
                      const float* ref2, int ld_ref2, __half* hi, __half* lo, int ld_out, float* scale,
                      int fixed_scale, cudaStream_t stream) {
    split_rows_kernel<<<rows, EW_THREADS, 0, stream>>>(in, ld_in, cols, chan_scale, chan_pp, m, ref2, ld_ref2, hi, lo,
                                                       ld_out, scale, fixed_scale);
    CSR_LAUNCHED();
    return CSR_OK;
}

// ---- K6: per-line-of-sight diagonals ---------------------------------------------------------------------
//   b[p, c]        = (w/(s k)) d                 the data term of backward(.) (ndft.py:32-36)
//   amask[p?, i]   = norm_factor/n if w > 0      backward o forward_normalized collapses to (1/n) E^H 1[w>0] E
//   fwd_chan[p?,i] = s norm_factor/(n w) if w>0  forward_normalized's channel factor without k (ndft.py:24-28)
__global__ void __launch_bounds__(EW_THREADS)
prepare_kernel(const float* __restrict__ data, int ld2m, int m, int n, const float* __restrict__ w, int w_pp,
               const float* __restrict__ s_in, int s_pp, const float* __restrict__ k, float norm_factor,
               float* __restrict__ b, float* __restrict__ amask, float* __restrict__ fwd_chan, float2* __restrict__ crow) {
    const int p = blockIdx.x;
    const float kp = k[p];
    const float* wrow = w + (w_pp ? (size_t)p * m : 0);
    const float* s = s_in + (s_pp ? (size_t)p * m : 0);
    const float* d = data + (size_t)p * ld2m;
    float* brow = b + (size_t)p * ld2m;
    const float inv_n = norm_factor / (float)n;
    float live = 0.f;  // channels with w > 0
    for (int i = threadIdx.x; i < m; i += EW_THREADS) {
        const float wi = wrow[i], si = s[i];
        live += wi > 0.f ? 1.f : 0.f;
        const float sk = wi / (si * kp);
        brow[i] = sk * d[i];
        brow[m + i] = sk * d[m + i];
        if (w_pp || p == 0) {
            const size_t o = (w_pp ? (size_t)p * m : 0) + i;
            amask[o] = wi > 0.f ? inv_n : 0.f;
            fwd_chan[o] = wi > 0.f ? si * inv_n / wi : 0.f;
        }
    }
    for (int c = 2 * m + threadIdx.x; c < ld2m; c += EW_THREADS) brow[c] = 0.f;
    if (crow != nullptr) {  // (kernel-uniform) coefficients of the incremental-screening bound (plan.cuh, IncrState)
        live = block_reduce(live, false);
        if (threadIdx.x == 0)
            crow[p] = w_pp ? make_float2(fabsf(inv_n) * (float)m * 1.0001f, fabsf(inv_n) * ((float)m - live) * 1.0001f)
                           : make_float2(fabsf(inv_n) * live * 1.0001f, 0.f);
    }
}

int launch_prepare(const csr_plan* plan, const csr_fista_args& a, float* b, float* amask, float* fwd_chan, float2* crow,
                   cudaStream_t stream) {
    prepare_kernel<<<a.P, EW_THREADS, 0, stream>>>(a.data, plan->ld2m, plan->m, plan->n, a.w, a.w_per_pixel, a.s,
                                                   a.s_per_pixel, a.k, a.norm_factor, b, amask, fwd_chan, crow);
    CSR_LAUNCHED();
    return CSR_OK;
}
// same for the per-line-of-sight sampling path (nudft.cu), which has no plan object
int launch_prepare_nu(const csr_fista_args& a, int m, int n, int ld2m, float* b, float* amask, float* fwd_chan,
                      cudaStream_t stream) {
    prepare_kernel<<<a.P, EW_THREADS, 0, stream>>>(a.data, ld2m, m, n, a.w, a.w_per_pixel, a.s, a.s_per_pixel, a.k,
                                                   a.norm_factor, b, amask, fwd_chan, nullptr);
    CSR_LAUNCHED();
    return CSR_OK;
}

// ---- final evaluate: residual, chi2 = 0.5 sum w |d - model|^2 (chi2.py:21-32), l1 = sum sqrt(x^2 + tiny) ----
__global__ void __launch_bounds__(EW_THREADS)
finalize_kernel(const float* __restrict__ data, const float* __restrict__ model, int ld2m, int m,
                const float* __restrict__ w, int w_pp, float* __restrict__ residual, const float* __restrict__ x,
                int ldx, int ncoef, float* __restrict__ cost) {
    const int p = blockIdx.x;
    const float* wrow = w + (w_pp ? (size_t)p * m : 0);
    const float* d = data + (size_t)p * ld2m;
    const float* md = model + (size_t)p * ld2m;
    float chi = 0.f;
    for (int i = threadIdx.x; i < m; i += EW_THREADS) {
        const float rr = d[i] - md[i], ri = d[m + i] - md[m + i];
        if (residual) {
            residual[(size_t)p * ld2m + i] = rr;
            residual[(size_t)p * ld2m + m + i] = ri;
        }
        chi += wrow[i] * (rr * rr + ri * ri);
    }
    chi = block_reduce(chi, false);
    float l1 = 0.f;
    const float* xr = x + (size_t)p * ldx;
    const float tiny = 1.17549435e-38f;  // np.finfo(np.float32).tiny (l1.py:18)
    for (int c = threadIdx.x; c < ncoef; c += EW_THREADS) l1 += sqrtf(xr[c] * xr[c] + tiny);
    l1 = block_reduce(l1, false);
    if (threadIdx.x == 0 && cost) {
        cost[2 * p] = 0.5f * chi;
        cost[2 * p + 1] = l1;
    }
}

int launch_finalize_dims(const csr_fista_args& a, int m, int ld2m, const float* model, float* residual, const float* x,
                         int ldx, int ncoef, cudaStream_t stream) {
    finalize_kernel<<<a.P, EW_THREADS, 0, stream>>>(a.data, model, ld2m, m, a.w, a.w_per_pixel, residual, x, ldx, ncoef,
                                                    a.cost);
    CSR_LAUNCHED();
    return CSR_OK;
}
int launch_finalize(const csr_plan* plan, const csr_fista_args& a, const float* model, float* residual,
                    const float* x, int ldx, int ncoef, cudaStream_t stream) {
    return launch_finalize_dims(a, plan->m, plan->ld2m, model, residual, x, ldx, ncoef, stream);
}

// ---- coefficient-space FISTA update (wavelet bases): fista.py:80-86 + l1.py:30-32 on [P, ncoef] ------------
__global__ void __launch_bounds__(EW_THREADS)
coef_update_kernel(const float* __restrict__ grad, float* __restrict__ x, float* __restrict__ z, int ld, int ncoef,
                   const float* __restrict__ lambda0, const float* __restrict__ noise, const int* __restrict__ iter_cap,
                   int iter, float step, float beta, float* __restrict__ delta) {
    const int p = blockIdx.x;
    const float l0 = lambda0[p], nz = noise[p];
    const float lam = l0 - (float)iter * nz;
    const bool active = (nz < l0) && (iter == 0 || lam > 0.f) && (!iter_cap || iter < iter_cap[p]);
    if (!active) return;
    const float thr = step * lam;
    const size_t base = (size_t)p * ld;
    float d = 0.f;
    for (int c = threadIdx.x; c < ncoef; c += EW_THREADS) {
        const float zo = z[base + c], xo = x[base + c];
        const float u = zo - step * grad[base + c];
        const float xn = copysignf(fmaxf(fabsf(u) - thr, 0.f), u);
        x[base + c] = xn;
        z[base + c] = xn + beta * (xn - xo);
        d += fabsf(xn - xo);
    }
    if (delta) {
        d = block_reduce(d, false);
        if (threadIdx.x == 0) delta[p] = d;
    }
}

int launch_coef_update(const float* grad, float* x, float* z, int ld, int ncoef, int P, const float* lambda0,
                       const float* noise, const int* iter_cap, int iter, float step, float beta, float* delta,
                       cudaStream_t stream) {
    coef_update_kernel<<<P, EW_THREADS, 0, stream>>>(grad, x, z, ld, ncoef, lambda0, noise, iter_cap, iter, step, beta,
                                                     delta);
    CSR_LAUNCHED();
    return CSR_OK;
}

// ---- Galactic RM de-rotation (base/dataset.py:377-381) ----------------------------------------------------
__global__ void __launch_bounds__(EW_THREADS)
derotate_kernel(float* __restrict__ data, int ld2m, int m, const double* __restrict__ lambda2,
                const float* __restrict__ phi_gal) {
    const int p = blockIdx.x;
    const double pg = (double)phi_gal[p];
    float* d = data + (size_t)p * ld2m;
    for (int i = threadIdx.x; i < m; i += EW_THREADS) {
        double sn, cs;
        sincos(-2.0 * pg * lambda2[i], &sn, &cs);
        const float re = d[i], im = d[m + i];
        d[i] = (float)(re * cs - im * sn);
        d[m + i] = (float)(re * sn + im * cs);
    }
}

// flags[p, blk] = 1 if any of the 128 values z[p, 128*blk .. 128*blk+127] is non-zero (the FISTA epilogue's
// "all-zero block" shortcut starts from here; x == z at that point)
// tile_flags (optional, zeroed by the caller): byte (p/32)%4 of word [p/128, blk] is set when the block is non-zero --
// the forward GEMM skips K blocks whose whole 128-row A tile is zero (dft_gemm.cu, build_kmask)
__global__ void block_flags_kernel(const float* __restrict__ z, int ld, int cols, unsigned char* __restrict__ flags,
                                   int flag_ld, unsigned char* __restrict__ tile_flags) {
    const int p = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int blk = warp; blk < flag_ld; blk += EW_THREADS / 32) {
        int nz = 0;
        for (int c = blk * 128 + lane; c < min(cols, blk * 128 + 128); c += 32) nz |= (z[(size_t)p * ld + c] != 0.f);
        nz = __any_sync(0xffffffffu, nz);
        if (lane == 0) {
            flags[(size_t)p * flag_ld + blk] = (unsigned char)nz;
            if (nz && tile_flags) tile_flags[((size_t)(p >> 7) * flag_ld + blk) * 4 + ((p >> 5) & 3)] = 1;
        }
    }
}
int launch_block_flags(const float* z, int ld, int cols, int P, unsigned char* flags, int flag_ld,
                       unsigned int* tile_flags, cudaStream_t stream) {
    if (tile_flags) CSR_CUDA(cudaMemsetAsync(tile_flags, 0, sizeof(unsigned int) * (size_t)((P + 127) / 128) * flag_ld, stream));
    block_flags_kernel<<<P, EW_THREADS, 0, stream>>>(z, ld, cols, flags, flag_ld, reinterpret_cast<unsigned char*>(tile_flags));
    CSR_LAUNCHED();
    return CSR_OK;
}

// tile_flags of the STATE (x or z non-zero), as the FISTA epilogue maintains them under zero_skip -- for runs without that
// shortcut, so that the forward transform groups its K blocks identically in both (GemmEpilogue::group_sparse)
__global__ void state_tile_flags_kernel(const float* __restrict__ x, const float* __restrict__ z, int ld, int cols, int flag_ld,
                                        unsigned char* __restrict__ tile_flags) {
    const int p = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int blk = warp; blk < flag_ld; blk += EW_THREADS / 32) {
        int nz = 0;
        for (int c = blk * 128 + lane; c < min(cols, blk * 128 + 128); c += 32)
            nz |= (x[(size_t)p * ld + c] != 0.f) | (z[(size_t)p * ld + c] != 0.f);
        nz = __any_sync(0xffffffffu, nz);
        if (lane == 0 && nz) tile_flags[((size_t)(p >> 7) * flag_ld + blk) * 4 + ((p >> 5) & 3)] = 1;
    }
}
int launch_state_tile_flags(const float* x, const float* z, int ld, int cols, int P, int flag_ld, unsigned int* tile_flags,
                            cudaStream_t stream) {
    CSR_CUDA(cudaMemsetAsync(tile_flags, 0, sizeof(unsigned int) * (size_t)((P + 127) / 128) * flag_ld, stream));
    state_tile_flags_kernel<<<P, EW_THREADS, 0, stream>>>(x, z, ld, cols, flag_ld, reinterpret_cast<unsigned char*>(tile_flags));
    CSR_LAUNCHED();
    return CSR_OK;
}

// scale a [P, ld] fp32 matrix in place by a scalar (used for the delta / len(x) normalisation)
__global__ void scale_vec_kernel(float* v, int n, float f) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] *= f;
}
int launch_scale_vec(float* v, int n, float f, cudaStream_t stream) {
    scale_vec_kernel<<<(n + 255) / 256, 256, 0, stream>>>(v, n, f);
    CSR_LAUNCHED();
    return CSR_OK;
}

}  // namespace csr

using namespace csr;

extern "C" int csr_derotate(csr_plan* plan, float* data, const float* phi_gal, int P, void* stream) {
    CSR_REQUIRE(plan && data && phi_gal && P > 0, "csr_derotate: bad argument");
    CSR_CUDA(cudaSetDevice(plan->device));
    derotate_kernel<<<P, EW_THREADS, 0, (cudaStream_t)stream>>>(data, plan->ld2m, plan->m, plan->d_l2, phi_gal);
    CSR_LAUNCHED();
    return CSR_OK;
}
