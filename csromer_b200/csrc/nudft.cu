// nudft.cu -- K4: the lambda^2 <-> phi transform when every line of sight has its OWN lambda^2 sampling (channels
// deleted per pixel by a flagger, transformers/flaggers/mean_flagger.py:36-40; per-pixel Dataset objects in the
// README cube recipe, README.md:327-332), where no twiddle matrix can be shared and the tensor-core GEMM of
// dft_gemm.cu does not apply.  Same operator as transformers/dfts/ndft.py:17-43 / nufft.py:50-77, evaluated per
// line of sight on the uniform Faraday-depth grid phi_j = phi0 + j dphi (reconstruction/parameter.py:105-116):
//
//   forward   P_i = sum_j x_j exp(+2i a_i phi_j)          adjoint   F_j = sum_i b_i exp(-2i a_i phi_j)
//   a_i = lambda2_i - l2_ref of that line of sight (fp64, per line of sight)
//
// No table is stored: twiddles are generated on the fly from a two-level factorisation of the uniform grid,
//   exp(i th (jb + j0)) = U(jb) V(j0),   V(j0), j0 < B, held per channel (registers / shared memory),
// with U(jb) evaluated exactly for every block (fp64 phase in turns, range-reduced, sincospif) -- no recurrences,
// so the error does not grow with n.  Per (channel, phi cell) pair the inner loop is one complex multiply-add
// (4 FFMA) plus ~1.5 amortised instructions, i.e. the kernels are FP32-pipe bound, not HBM bound: one line of
// sight moves 8(m + n) bytes per transform and does 8 m n flops.  The forward transform skips blocks of B phi cells
// that are identically zero (exact; FISTA iterates are sparse).
#include <math.h>

#include "plan.cuh"

namespace csr {

constexpr int NU_B = 16;             // phi cells per block of the forward transform (V in registers)
constexpr int NU_FWD_THREADS = 128;  // channels per CTA
constexpr int NU_FWD_TILE = 2048;    // phi cells staged in shared memory at a time
constexpr int NU_JB = 32;            // phi cells per thread of the adjoint transform
constexpr int NU_ADJ_THREADS = 128;
constexpr int NU_IT = 64;            // channels staged per shared-memory tile of the adjoint

struct NuGrid {
    const double* a;     // [P, ldl]  lambda2 - l2_ref
    int ldl;
    const int* m_count;  // [P] valid channels of each line of sight, or null (= m)
    int m, n;
    double phi0, dphi;
};

__device__ __forceinline__ float2 cis_turns(double t) {  // exp(2 pi i t), t in turns
    t -= rint(t);
    float s, c;
    sincospif(2.0f * (float)t, &s, &c);
    return make_float2(c, s);
}
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ void cfma(float2& acc, float2 a, float2 b) {
    acc.x = fmaf(a.x, b.x, acc.x);
    acc.x = fmaf(-a.y, b.y, acc.x);
    acc.y = fmaf(a.x, b.y, acc.y);
    acc.y = fmaf(a.y, b.x, acc.y);
}

// out[p, i] = row_scale[p] chan_scale[p, i] sum_j x[p, j] exp(+2i a_i phi_j) - sub[p, i]      (planar [Re | Im])
__global__ void __launch_bounds__(NU_FWD_THREADS)
nudft_forward_kernel(NuGrid g, const float* __restrict__ x, int ldx, const float* __restrict__ chan_scale,
                     const float* __restrict__ row_scale, const float* __restrict__ sub, float* __restrict__ out,
                     int ld_out) {
    __shared__ __align__(16) float xs_re[NU_FWD_TILE];
    __shared__ __align__(16) float xs_im[NU_FWD_TILE];
    __shared__ unsigned int blk_nz[NU_FWD_TILE / NU_B / 32];
    const int p = blockIdx.y;
    const int i = blockIdx.x * NU_FWD_THREADS + threadIdx.x;
    const int mp = g.m_count ? g.m_count[p] : g.m;
    const bool valid = i < mp;
    const double a = valid ? g.a[(size_t)p * g.ldl + i] : 0.0;
    const double turn_step = a * g.dphi * (1.0 / M_PI);  // 2 a dphi / (2 pi)
    const double turn0 = a * g.phi0 * (1.0 / M_PI);
    float2 V[NU_B];
#pragma unroll
    for (int j0 = 0; j0 < NU_B; ++j0) V[j0] = cis_turns(turn_step * j0);
    float2 acc = make_float2(0.f, 0.f);
    const float* xr = x + (size_t)p * ldx;
    for (int t0 = 0; t0 < g.n; t0 += NU_FWD_TILE) {
        __syncthreads();
        if (threadIdx.x < NU_FWD_TILE / NU_B / 32) blk_nz[threadIdx.x] = 0u;
        __syncthreads();
        for (int c = threadIdx.x; c < NU_FWD_TILE; c += NU_FWD_THREADS) {
            const int j = t0 + c;
            const float re = j < g.n ? xr[j] : 0.f, im = j < g.n ? xr[g.n + j] : 0.f;
            xs_re[c] = re;
            xs_im[c] = im;
            // 16 consecutive lanes hold one block of NU_B cells
            const unsigned int nzmask = __ballot_sync(0xffffffffu, re != 0.f || im != 0.f);
            if ((threadIdx.x & 31) == 0) {
                const int b0 = c / NU_B;
                if (nzmask & 0xffffu) atomicOr(&blk_nz[b0 >> 5], 1u << (b0 & 31));
                if (nzmask >> 16) atomicOr(&blk_nz[(b0 + 1) >> 5], 1u << ((b0 + 1) & 31));
            }
        }
        __syncthreads();
        const int nblk = min(NU_FWD_TILE, g.n - t0 + NU_B - 1) / NU_B;
        for (int w = 0; w * 32 < nblk; ++w) {
            unsigned int bits = blk_nz[w];
            while (bits) {  // uniform over the CTA
                const int b = w * 32 + __ffs(bits) - 1;
                bits &= bits - 1;
                float2 S = make_float2(0.f, 0.f);
#pragma unroll
                for (int q = 0; q < NU_B / 4; ++q) {
                    const float4 r4 = *reinterpret_cast<const float4*>(&xs_re[b * NU_B + 4 * q]);
                    const float4 i4 = *reinterpret_cast<const float4*>(&xs_im[b * NU_B + 4 * q]);
                    cfma(S, make_float2(r4.x, i4.x), V[4 * q + 0]);
                    cfma(S, make_float2(r4.y, i4.y), V[4 * q + 1]);
                    cfma(S, make_float2(r4.z, i4.z), V[4 * q + 2]);
                    cfma(S, make_float2(r4.w, i4.w), V[4 * q + 3]);
                }
                const float2 U = cis_turns(turn0 + turn_step * (double)(t0 + b * NU_B));
                cfma(acc, S, U);
            }
        }
    }
    if (i < g.m) {
        float sc = valid ? 1.f : 0.f;
        if (valid && chan_scale) sc *= chan_scale[(size_t)p * g.m + i];
        if (valid && row_scale) sc *= row_scale[p];
        float re = acc.x * sc, im = acc.y * sc;
        if (sub) {
            re -= sub[(size_t)p * ld_out + i];
            im -= sub[(size_t)p * ld_out + g.m + i];
        }
        out[(size_t)p * ld_out + i] = re;
        out[(size_t)p * ld_out + g.m + i] = im;
    }
}

// out[p, j] = row_scale[p] sum_i chan_scale[p, i] b[p, i] exp(-2i a_i phi_j)
__global__ void __launch_bounds__(NU_ADJ_THREADS)
nudft_adjoint_kernel(NuGrid g, const float* __restrict__ b, int ldb, const float* __restrict__ chan_scale,
                     const float* __restrict__ row_scale, float* __restrict__ out, int ld_out) {
    __shared__ __align__(16) float2 Vs[NU_IT][NU_JB];  // exp(-2 pi i turn_step_i j0)
    __shared__ double ts_s[NU_IT], t0_s[NU_IT];
    __shared__ float2 bv_s[NU_IT];
    const int p = blockIdx.y;
    const int jb = (blockIdx.x * NU_ADJ_THREADS + threadIdx.x) * NU_JB;
    const bool valid = jb < g.n;
    const int mp = g.m_count ? g.m_count[p] : g.m;
    float2 acc[NU_JB];
#pragma unroll
    for (int j0 = 0; j0 < NU_JB; ++j0) acc[j0] = make_float2(0.f, 0.f);
    for (int i0 = 0; i0 < mp; i0 += NU_IT) {
        __syncthreads();
        // fill the channel tile: NU_ADJ_THREADS / NU_IT threads per channel share its NU_JB twiddles
        {
            constexpr int TPC = NU_ADJ_THREADS / NU_IT;
            const int ii = threadIdx.x / TPC, part = threadIdx.x % TPC;
            const int i = i0 + ii;
            double ts = 0.0, t0 = 0.0;
            float2 bv = make_float2(0.f, 0.f);
            if (i < mp) {
                const double a = g.a[(size_t)p * g.ldl + i];
                ts = -a * g.dphi * (1.0 / M_PI);
                t0 = -a * g.phi0 * (1.0 / M_PI);
                const float cs = chan_scale ? chan_scale[(size_t)p * g.m + i] : 1.f;
                bv = make_float2(cs * b[(size_t)p * ldb + i], cs * b[(size_t)p * ldb + g.m + i]);
            }
            for (int j0 = part; j0 < NU_JB; j0 += TPC) Vs[ii][j0] = cis_turns(ts * j0);
            if (part == 0) {
                ts_s[ii] = ts;
                t0_s[ii] = t0;
                bv_s[ii] = bv;
            }
        }
        __syncthreads();
        if (valid) {
            const int cnt = min(NU_IT, mp - i0);
            for (int ii = 0; ii < cnt; ++ii) {
                const float2 c = cmul(bv_s[ii], cis_turns(t0_s[ii] + ts_s[ii] * (double)jb));
#pragma unroll
                for (int q = 0; q < NU_JB / 2; ++q) {
                    const float4 v = *reinterpret_cast<const float4*>(&Vs[ii][2 * q]);
                    cfma(acc[2 * q], c, make_float2(v.x, v.y));
                    cfma(acc[2 * q + 1], c, make_float2(v.z, v.w));
                }
            }
        }
    }
    if (valid) {
        const float rs = row_scale ? row_scale[p] : 1.f;
        float* o = out + (size_t)p * ld_out;
#pragma unroll
        for (int j0 = 0; j0 < NU_JB; ++j0) {
            if (jb + j0 < g.n) {
                o[jb + j0] = rs * acc[j0].x;
                o[g.n + jb + j0] = rs * acc[j0].y;
            }
        }
    }
}

static int check_grid(const csr_nu_sampling* s, int P, NuGrid* g) {
    CSR_REQUIRE(s && s->lambda2 && s->m > 0 && s->n > 0 && s->ldl >= s->m && P > 0, "non-uniform sampling: bad argument");
    CSR_REQUIRE(s->dphi != 0.0 && isfinite(s->dphi) && isfinite(s->phi0), "non-uniform sampling: bad phi grid");
    g->a = s->lambda2;
    g->ldl = s->ldl;
    g->m_count = s->m_count;
    g->m = s->m;
    g->n = s->n;
    g->phi0 = s->phi0;
    g->dphi = s->dphi;
    return CSR_OK;
}

// the NUFFT path (nufft.cu): O(Nf log Nf) per line of sight in shared memory, used when the grid fits (n <= 8192) and the
// option `nufft` is on; ~1e-6 of the result's scale away from these exact sums
int nufft_grid_size(int n, int* logNf);
int launch_nufft_forward(const double* a, int ldl, const int* m_count, int m, int n, double phi0, double dphi, const float* x, int ldx,
                         const float* chan_scale, const float* row_scale, const float* sub, float* out, int ld_out, int P,
                         cudaStream_t stream);
int launch_nufft_adjoint_fista(const double* a, int ldl, const int* m_count, int m, int n, double phi0, double dphi, const float* b, int ldb,
                               float* x, float* z, int ld, const float* lambda0, const float* noise, const int* iter_cap, int iter,
                               float step, float beta, float* delta, int P, cudaStream_t stream);
int launch_nufft_adjoint(const double* a, int ldl, const int* m_count, int m, int n, double phi0, double dphi, const float* b, int ldb,
                         const float* chan_scale, const float* row_scale, float* out, int ld_out, int P, cudaStream_t stream);
static bool use_nufft(const NuGrid& g) { return options().nufft != 0 && nufft_grid_size(g.n, nullptr) > 0; }

int launch_nudft_forward(const NuGrid& g, const float* x, int ldx, const float* chan_scale, const float* row_scale,
                         const float* sub, float* out, int ld_out, int P, cudaStream_t stream) {
    if (use_nufft(g))
        return launch_nufft_forward(g.a, g.ldl, g.m_count, g.m, g.n, g.phi0, g.dphi, x, ldx, chan_scale, row_scale, sub, out, ld_out, P,
                                    stream);
    dim3 grid((g.m + NU_FWD_THREADS - 1) / NU_FWD_THREADS, P);
    nudft_forward_kernel<<<grid, NU_FWD_THREADS, 0, stream>>>(g, x, ldx, chan_scale, row_scale, sub, out, ld_out);
    CSR_LAUNCHED();
    return CSR_OK;
}
int launch_nudft_adjoint(const NuGrid& g, const float* b, int ldb, const float* chan_scale, const float* row_scale,
                         float* out, int ld_out, int P, cudaStream_t stream) {
    if (use_nufft(g))
        return launch_nufft_adjoint(g.a, g.ldl, g.m_count, g.m, g.n, g.phi0, g.dphi, b, ldb, chan_scale, row_scale, out, ld_out, P, stream);
    dim3 grid((g.n + NU_ADJ_THREADS * NU_JB - 1) / (NU_ADJ_THREADS * NU_JB), P);
    nudft_adjoint_kernel<<<grid, NU_ADJ_THREADS, 0, stream>>>(g, b, ldb, chan_scale, row_scale, out, ld_out);
    CSR_LAUNCHED();
    return CSR_OK;
}

// helpers shared with fista.cu / elementwise.cu / wavelet.cu
int launch_prepare_nu(const csr_fista_args& a, int m, int n, int ld2m, float* b, float* amask, float* fwd_chan,
                      cudaStream_t stream);
int launch_finalize_dims(const csr_fista_args& a, int m, int ld2m, const float* model, float* residual, const float* x,
                         int ldx, int ncoef, cudaStream_t stream);
int launch_coef_update(const float* grad, float* x, float* z, int ld, int ncoef, int P, const float* lambda0,
                       const float* noise, const int* iter_cap, int iter, float step, float beta, float* delta,
                       cudaStream_t stream);
int launch_scale_vec(float* v, int n, float f, cudaStream_t stream);
int wavelet_ld(const csr_wavelet* w);
int wavelet_decompose(const csr_wavelet* w, const float* signal, int ld_sig, float* coef, int ld_coef, int P,
                      float* ws, cudaStream_t stream);
int wavelet_reconstruct(const csr_wavelet* w, const float* coef, int ld_coef, float* signal, int ld_sig, int P,
                        float* ws, cudaStream_t stream);

struct NuWs {
    float *b, *amask, *fwd_chan, *z, *f_dirty, *model, *r, *g, *sig, *gcoef, *wav_ws;
    size_t bytes;
};
static NuWs carve_nu(int m, int n, const csr_wavelet* wav, int P, void* ws) {
    char* base = (char*)ws;
    size_t off = 0;
    auto take = [&](size_t count) -> float* {
        float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
        off += align256(count * sizeof(float));
        return p;
    };
    const int ld2m = round_up(2 * m, 8), ld2n = round_up(2 * n, 8);
    const int ldc = wav ? round_up(wav->ncoef, 8) : ld2n;
    NuWs w;
    w.b = take((size_t)P * ld2m);
    w.amask = take((size_t)P * m);
    w.fwd_chan = take((size_t)P * m);
    w.z = take((size_t)P * ldc);
    w.f_dirty = take((size_t)P * ld2n);
    w.model = take((size_t)P * ld2m);
    w.r = take((size_t)P * ld2m);
    w.g = take((size_t)P * ld2n);
    w.sig = w.gcoef = w.wav_ws = nullptr;
    if (wav) {
        w.sig = take((size_t)P * ld2n);
        w.gcoef = take((size_t)P * ldc);
        w.wav_ws = take((size_t)2 * P * wavelet_ld(wav));
    }
    w.bytes = off;
    return w;
}

}  // namespace csr

using namespace csr;

extern "C" int csr_nudft_forward(const csr_nu_sampling* s, const float* x, int ldx, const float* chan_scale,
                                 const float* row_scale, float* out, int ld_out, int P, void* stream) {
    NuGrid g;
    int rc = check_grid(s, P, &g);
    if (rc) return rc;
    CSR_REQUIRE(x && out && ldx >= 2 * s->n && ld_out >= 2 * s->m, "csr_nudft_forward: bad argument");
    return launch_nudft_forward(g, x, ldx, chan_scale, row_scale, nullptr, out, ld_out, P, (cudaStream_t)stream);
}

extern "C" int csr_nudft_adjoint(const csr_nu_sampling* s, const float* b, int ldb, const float* chan_scale,
                                 const float* row_scale, float* out, int ld_out, int P, void* stream) {
    NuGrid g;
    int rc = check_grid(s, P, &g);
    if (rc) return rc;
    CSR_REQUIRE(b && out && ldb >= 2 * s->m && ld_out >= 2 * s->n, "csr_nudft_adjoint: bad argument");
    return launch_nudft_adjoint(g, b, ldb, chan_scale, row_scale, out, ld_out, P, (cudaStream_t)stream);
}

extern "C" size_t csr_fista_nu_ws_bytes(const csr_nu_sampling* s, const csr_wavelet* wav, int P) {
    if (!s || P <= 0 || s->m <= 0 || s->n <= 0) return 0;
    return carve_nu(s->m, s->n, wav, P, nullptr).bytes;
}

// The FISTA loop of csr_fista for per-line-of-sight samplings: same arguments and semantics, `w` and `s` are [P, m].
// Per iteration: forward (residual fused) -> adjoint -> threshold / momentum update; with a wavelet basis the
// synthesis / analysis kernels of wavelet.cu sit around the two transforms.
extern "C" int csr_fista_nu(const csr_nu_sampling* s, csr_wavelet* wav, const csr_fista_args* args, void* ws,
                            size_t ws_bytes, void* stream_) {
    CSR_REQUIRE(args, "csr_fista_nu: null args");
    const csr_fista_args& a = *args;
    NuGrid g;
    int rc = check_grid(s, a.P, &g);
    if (rc) return rc;
    CSR_REQUIRE(a.iters >= 0 && a.data && a.w && a.s && a.k && a.x && a.lambda0 && a.noise, "csr_fista_nu: null input array");
    CSR_REQUIRE(a.w_per_pixel == 1 && a.s_per_pixel == 1, "csr_fista_nu: w and s must be per line of sight ([P, m])");
    CSR_REQUIRE(a.f_dirty_in == nullptr, "csr_fista_nu: f_dirty_in is not supported");
    const int P = a.P, m = s->m, n = s->n;
    const int ld2m = round_up(2 * m, 8), ld2n = round_up(2 * n, 8);
    const int ncoef = wav ? wav->ncoef : 2 * n;
    if (wav) CSR_REQUIRE(wav->N == 2 * n, "csr_fista_nu: wavelet length %d does not match 2n = %d", wav->N, 2 * n);
    const int ldc = wav ? round_up(wav->ncoef, 8) : ld2n;
    CSR_REQUIRE(a.ldx == ldc, "csr_fista_nu: ldx=%d, expected %d", a.ldx, ldc);
    NuWs w = carve_nu(m, n, wav, P, ws);
    if (!ws || ws_bytes < w.bytes) {
        set_error("csr_fista_nu: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
        return CSR_ERR_WORKSPACE;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    float* f_dirty = a.f_dirty ? a.f_dirty : w.f_dirty;
    float* model = a.model ? a.model : w.model;
    float* x = a.x;

    if ((rc = launch_prepare_nu(a, m, n, ld2m, w.b, w.amask, w.fwd_chan, stream))) return rc;
    if ((rc = launch_nudft_adjoint(g, w.b, ld2m, nullptr, nullptr, f_dirty, ld2n, P, stream))) return rc;
    if (a.use_dirty_as_guess) {
        if (wav) {
            if ((rc = wavelet_decompose(wav, f_dirty, ld2n, x, ldc, P, w.wav_ws, stream))) return rc;
        } else {
            CSR_CUDA(cudaMemcpyAsync(x, f_dirty, sizeof(float) * (size_t)P * ld2n, cudaMemcpyDeviceToDevice, stream));
        }
    }
    CSR_CUDA(cudaMemcpyAsync(w.z, x, sizeof(float) * (size_t)P * ldc, cudaMemcpyDeviceToDevice, stream));
    if (a.delta) CSR_CUDA(cudaMemsetAsync(a.delta, 0, sizeof(float) * P, stream));

    double t = 1.0;
    for (int it = 0; it < a.iters; ++it) {
        const double t0 = t;
        t = 0.5 * (1.0 + sqrt(1.0 + 4.0 * t * t));
        const float beta = (float)((t0 - 1.0) / t);
        const bool last = (it == a.iters - 1);
        const float* zs = w.z;
        if (wav) {
            if ((rc = wavelet_reconstruct(wav, w.z, ldc, w.sig, ld2n, P, w.wav_ws, stream))) return rc;
            zs = w.sig;
        }
        // r = 1[w>0]/n (E z) - (w/(s k)) d ; g = E^H r          (chi2.py:49-52, dataset.py:374)
        if ((rc = launch_nudft_forward(g, zs, ld2n, w.amask, nullptr, w.b, w.r, ld2m, P, stream))) return rc;
        if (!wav && use_nufft(g)) {   // the threshold / momentum update sits in the output stage of the adjoint NUFFT
            if ((rc = launch_nufft_adjoint_fista(g.a, g.ldl, g.m_count, g.m, g.n, g.phi0, g.dphi, w.r, ld2m, x, w.z, ldc, a.lambda0, a.noise,
                                                 a.iter_cap, it, a.step, beta, (last && a.delta) ? a.delta : nullptr, P, stream)))
                return rc;
            continue;
        }
        if ((rc = launch_nudft_adjoint(g, w.r, ld2m, nullptr, nullptr, w.g, ld2n, P, stream))) return rc;
        const float* grad = w.g;
        if (wav) {
            if ((rc = wavelet_decompose(wav, w.g, ld2n, w.gcoef, ldc, P, w.wav_ws, stream))) return rc;
            grad = w.gcoef;
        }
        if ((rc = launch_coef_update(grad, x, w.z, ldc, ncoef, P, a.lambda0, a.noise, a.iter_cap, it, a.step, beta,
                                     (last && a.delta) ? a.delta : nullptr, stream)))
            return rc;
    }
    if (a.delta && a.iters > 0) {
        if ((rc = launch_scale_vec(a.delta, P, 1.0f / (float)ncoef, stream))) return rc;
    }
    if (a.model || a.residual || a.cost) {
        const float* xs = x;
        if (wav) {
            if ((rc = wavelet_reconstruct(wav, x, ldc, w.sig, ld2n, P, w.wav_ws, stream))) return rc;
            xs = w.sig;
        }
        if ((rc = launch_nudft_forward(g, xs, ld2n, w.fwd_chan, a.k, nullptr, model, ld2m, P, stream))) return rc;
        if ((rc = launch_finalize_dims(a, m, ld2m, model, a.residual, x, ldc, ncoef, stream))) return rc;
    }
    return CSR_OK;
}
