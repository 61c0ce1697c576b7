// nufft.cu -- K4n: the per-line-of-sight lambda^2 <-> phi transform as a non-uniform FFT, one CTA per line of sight,
// everything between the global loads and stores in shared memory.
//
// Replaces the role pynufft plays in transformers/dfts/nufft.py:36-75 (plan, FFT, interpolate) for samplings that differ
// from pixel to pixel, where no twiddle matrix can be shared (nudft.cu evaluates the same sums directly, O(m n), FP32-pipe
// bound).  On the uniform grid phi_j = phi_c + j' dphi (j' = j - n/2) the operator is a trigonometric polynomial at the
// non-uniform frequencies omega_i = 2 a_i dphi, a_i = lambda2_i - l2_ref:
//   forward (type 2)  y_i = c_i sum_j' x_j' exp(+i omega_i j'),   adjoint (type 1)  F_j' = sum_i conj(c_i) b_i exp(-i omega_i j'),
//   c_i = exp(2i a_i phi_c).
// NUFFT with the "exponential of semicircle" kernel psi(t) = exp(beta (sqrt(1 - (2t/W)^2) - 1)), W taps, on a grid of
// Nf = 2^q >= 2n points (Barnett, Magland & af Klinteberg 2019):
//   type 2: u_j' = x_j' / psihat(j') -> FFT (+) of the zero-padded Nf-vector -> y_i = c_i sum_l U_{k0+l} psi(kappa_i - k0 - l)
//   type 1: G_k = sum_i conj(c_i) b_i psi(kappa_i - k) -> FFT (-) -> F_j' = g_j' / psihat(j'),     kappa_i = omega_i Nf / 2 pi,
// the exact adjoint of type 2 (same kernel, same grid), so FISTA sees a consistent operator pair.  W = 8, beta = 2.30 W
// gives ~1e-7 aliasing error; the fp32 FFT (q <= 14 stages) leaves ~1e-6 of the result's scale -- inside the stated
// 1e-5 transform tolerance, and checked against the exact kernels (tests/test_gpu_parity.py).
//
// Data movement: one line of sight reads 8 n + 8 m and writes 8 m (or 8 n) bytes of HBM per transform as 16-byte / fully
// coalesced row accesses; the Nf-point complex FFT (128 KiB at n = 7968) never leaves shared memory: radix-2 butterflies
// in place, decimation in frequency for type 2 (natural-order input, outputs addressed bit-reversed by the
// interpolation) and decimation in time for type 1 (the spreading writes bit-reversed, the cropped output is natural), so
// no separate permutation pass exists; twiddles come from a 64 KiB table built once per CTA.  Cost per line of sight:
// 5 Nf log2 Nf = 1.1 MFLOP instead of 8 m n = 64 MFLOP.  Needs Nf <= 16384 (n <= 8192); longer grids stay on nudft.cu.
#include <math.h>

#include "plan.cuh"

namespace csr {

bool profile_start(cudaStream_t stream, cudaEvent_t* start, cudaEvent_t* stop);
void profile_stop(int kind, cudaStream_t stream, cudaEvent_t start, cudaEvent_t stop);

constexpr int NF_THREADS = 1024;
constexpr int NF_MAX = 16384;
constexpr int NF_SPARSE_MAX = 96;   // non-zero cells up to which the forward transform sums over the cells directly

struct NufftGrid {
    const double* a;     // [P, ldl]  lambda2 - l2_ref
    int ldl;
    const int* m_count;  // [P] or null
    int m, n, Nf, logNf, W;
    float beta;
    double phic, dphi;    // phi of the centre cell j = n/2, cell size
    const float* deconv;  // [n] 1 / psihat(j - n/2)
};

// optional FISTA update fused into the output stage of the adjoint transform (delta basis, csr_fista_nu): the value that would be
// written as the gradient goes straight into u = z - step g, x = soft(u, step lambda), z = x + beta (x - x_old)
// (fista.py:80-86, l1.py:30-32; the arithmetic of coef_update_kernel); x / z entries that do not change are not written
struct NufftUpdate {
    float* x;
    float* z;
    int ld;
    const float* lambda0;
    const float* noise;
    const int* iter_cap;
    int iter;
    float step, beta;
    float* delta;   // optional [P]: sum |x_new - x_old| of the line of sight
};

__device__ __forceinline__ float es_kernel(float t, float inv_half_w, float beta) {
    const float u = t * inv_half_w, s = 1.f - u * u;
    return s > 0.f ? expf(beta * (sqrtf(s) - 1.f)) : 0.f;
}
__device__ __forceinline__ float2 cmulf(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
__device__ __forceinline__ int brev(int i, int logn) { return (int)(__brev((unsigned)i) >> (32 - logn)); }
// Skewed slot of twiddle k: a stage reads the table with a power-of-two stride (2^(logn - s) entries), which would put a
// warp's 32 reads into one bank; one pad entry per 16, 256 and 4096 entries spreads every such stride over the banks.
__host__ __device__ __forceinline__ int tw_slot(int k) { return k + (k >> 4) + (k >> 8) + (k >> 12); }

// in-place FFT over d[0 .. N), N = 2^logn, tw[k] = exp(+2 pi i k / N) for k < N / 2 (skewed slots).  Two radix-2 stages
// are fused into one radix-4 pass over registers (four loads, four stores and two twiddle loads per four points and two
// stages: half the shared-memory traffic of radix-2, which is what bounds the transform); an odd logn starts / ends with
// one radix-2 stage.
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mul_i(float2 a) { return make_float2(-a.y, a.x); }    // a * i
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }   // a * (-i)

// DIF: natural-order input -> bit-reversed output, exponent sign +
__device__ __forceinline__ void fft_dif_plus(float2* d, const float2* tw, int logn) {
    const int n = 1 << logn;
    int s = logn;
    if (logn & 1) {  // one radix-2 stage first
        const int half = 1 << (s - 1);
        for (int b = threadIdx.x; b < (n >> 1); b += NF_THREADS) {
            const int pos = b & (half - 1), i0 = ((b >> (s - 1)) << s) + pos, i1 = i0 + half;
            const float2 u = d[i0], v = d[i1], w = tw[tw_slot(pos << (logn - s))];
            d[i0] = cadd(u, v);
            d[i1] = cmulf(csub(u, v), w);
        }
        __syncthreads();
        --s;
    }
    for (; s >= 2; s -= 2) {  // stages s and s - 1: blocks of 4 q points, q = 2^(s-2)
        const int q = 1 << (s - 2);
        for (int b = threadIdx.x; b < (n >> 2); b += NF_THREADS) {
            const int pos = b & (q - 1), i0 = ((b >> (s - 2)) << s) + pos;
            const float2 a0 = d[i0], a1 = d[i0 + q], a2 = d[i0 + 2 * q], a3 = d[i0 + 3 * q];
            const float2 ws = tw[tw_slot(pos << (logn - s))];          // W_s^pos; W_s^(pos+q) = i W_s^pos
            const float2 wt = tw[tw_slot(pos << (logn - s + 1))];      // W_(s-1)^pos
            const float2 b0 = cadd(a0, a2), b2 = cmulf(csub(a0, a2), ws);
            const float2 b1 = cadd(a1, a3), b3 = mul_i(cmulf(csub(a1, a3), ws));
            d[i0] = cadd(b0, b1);
            d[i0 + q] = cmulf(csub(b0, b1), wt);
            d[i0 + 2 * q] = cadd(b2, b3);
            d[i0 + 3 * q] = cmulf(csub(b2, b3), wt);
        }
        __syncthreads();
    }
}
// DIT: bit-reversed input -> natural-order output, exponent sign -
__device__ __forceinline__ void fft_dit_minus(float2* d, const float2* tw, int logn) {
    const int n = 1 << logn;
    int s = 1;
    for (; s + 1 <= logn; s += 2) {  // stages s and s + 1: blocks of 4 q points, q = 2^(s-1)
        const int q = 1 << (s - 1);
        for (int b = threadIdx.x; b < (n >> 2); b += NF_THREADS) {
            const int pos = b & (q - 1), i0 = ((b >> (s - 1)) << (s + 1)) + pos;
            const float2 a0 = d[i0], a1 = d[i0 + q], a2 = d[i0 + 2 * q], a3 = d[i0 + 3 * q];
            float2 w1 = tw[tw_slot(pos << (logn - s))];                // conj(W_s^pos)
            float2 w2 = tw[tw_slot(pos << (logn - s - 1))];            // conj(W_(s+1)^pos); conj(W_(s+1)^(pos+q)) = -i times it
            w1.y = -w1.y;
            w2.y = -w2.y;
            const float2 t1 = cmulf(a1, w1), t3 = cmulf(a3, w1);
            const float2 b0 = cadd(a0, t1), b1 = csub(a0, t1), b2 = cadd(a2, t3), b3 = csub(a2, t3);
            const float2 u2 = cmulf(b2, w2), u3 = mul_mi(cmulf(b3, w2));
            d[i0] = cadd(b0, u2);
            d[i0 + 2 * q] = csub(b0, u2);
            d[i0 + q] = cadd(b1, u3);
            d[i0 + 3 * q] = csub(b1, u3);
        }
        __syncthreads();
    }
    if (s == logn) {  // odd logn: the last radix-2 stage
        const int half = 1 << (s - 1);
        for (int b = threadIdx.x; b < (n >> 1); b += NF_THREADS) {
            const int pos = b & (half - 1), i0 = ((b >> (s - 1)) << s) + pos, i1 = i0 + half;
            float2 w = tw[tw_slot(pos << (logn - s))];
            w.y = -w.y;
            const float2 u = d[i0], v = cmulf(d[i1], w);
            d[i0] = cadd(u, v);
            d[i1] = csub(u, v);
        }
        __syncthreads();
    }
}

__device__ __forceinline__ void build_twiddles(float2* tw, int Nf) {
    for (int k = threadIdx.x; k < Nf / 2; k += NF_THREADS) {
        float s, c;
        sincospif(2.0f * (float)k / (float)Nf, &s, &c);  // (k / Nf is exact in fp32: both are small powers of two apart)
        tw[tw_slot(k)] = make_float2(c, s);
    }
}

// fractional grid position of channel i and its phase factor c_i
__device__ __forceinline__ void channel_geometry(const NufftGrid& g, double a, double& kappa, float2& c) {
    double t = a * g.dphi / 3.14159265358979323846;  // omega / 2 pi, in turns
    t -= floor(t);
    kappa = t * (double)g.Nf;
    double pc = a * g.phic / 3.14159265358979323846;  // 2 a phi_c / 2 pi
    pc -= rint(pc);
    float s, cc;
    sincospif(2.0f * (float)pc, &s, &cc);
    c = make_float2(cc, s);
}

// out[p, i] = row_scale[p] chan_scale[p, i] y_i - sub[p, i]      (planar [Re | Im])
__global__ void __launch_bounds__(NF_THREADS, 1)
nufft_forward_kernel(NufftGrid g, const float* __restrict__ x, int ldx, const float* __restrict__ chan_scale,
                     const float* __restrict__ row_scale, const float* __restrict__ sub, float* __restrict__ out, int ld_out, int P,
                     int sparse_max) {
    extern __shared__ __align__(16) float2 nsm[];
    float2* d = nsm;
    float2* tw = nsm + g.Nf;
    // sparse input: the non-zero cells of the line of sight, (j', Re, Im); with at most NF_SPARSE_MAX of them (an L1 iterate has
    // a few dozen) the sum over THOSE cells is cheaper than any FFT and exact
    float* const nz_base = reinterpret_cast<float*>(nsm + g.Nf + tw_slot(g.Nf / 2) + 1);   // [2][3 * NF_SPARSE_MAX]
    float* nz_list = nz_base;
    __shared__ int nz_count;
    build_twiddles(tw, g.Nf);
    const float inv_half_w = 2.f / (float)g.W;
    const int half_n = g.n >> 1;
    for (int p = blockIdx.x; p < P; p += gridDim.x) {
        const float* xr = x + (size_t)p * ldx;
        nz_list = nz_base;
        __syncthreads();  // (the previous line of sight has been read out; the twiddle table is complete)
        if (threadIdx.x == 0) nz_count = 0;
        __syncthreads();
        // natural order: grid index k' <-> j' = k' (k' < Nf/2) or k' - Nf; phi cells j' in [-n/2, n - n/2)
        for (int idx = threadIdx.x; idx < g.Nf; idx += NF_THREADS) {
            const int jp = idx < (g.Nf >> 1) ? idx : idx - g.Nf;
            const int j = jp + half_n;
            float2 v = make_float2(0.f, 0.f);
            if (j >= 0 && j < g.n) {
                const float dc = __ldg(g.deconv + j);
                const float re = xr[j], im = xr[g.n + j];
                v = make_float2(re * dc, im * dc);
                if (sparse_max > 0 && (re != 0.f || im != 0.f)) {
                    const int slot = atomicAdd(&nz_count, 1);
                    if (slot < sparse_max) {
                        nz_list[3 * slot] = (float)jp;
                        nz_list[3 * slot + 1] = re;
                        nz_list[3 * slot + 2] = im;
                    }
                }
            }
            d[idx] = v;
        }
        __syncthreads();
        const int nnz = nz_count;
        const bool direct = sparse_max > 0 && nnz <= sparse_max;   // (CTA-uniform)
        if (!direct) fft_dif_plus(d, tw, g.logNf);
        if (direct) {   // the slots were handed out in arrival order: sort by cell, so that the sums below run in a fixed order
            float* sorted = nz_list + 3 * NF_SPARSE_MAX;
            if ((int)threadIdx.x < nnz) {
                const float key = nz_list[3 * threadIdx.x];
                int rank = 0;
                for (int q = 0; q < nnz; ++q) rank += nz_list[3 * q] < key;
                sorted[3 * rank] = key;
                sorted[3 * rank + 1] = nz_list[3 * threadIdx.x + 1];
                sorted[3 * rank + 2] = nz_list[3 * threadIdx.x + 2];
            }
            __syncthreads();
            nz_list = sorted;
        }
        const int mc = g.m_count ? min(g.m_count[p], g.m) : g.m;
        const float rs = row_scale ? row_scale[p] : 1.f;
        const double* ap = g.a + (size_t)p * g.ldl;
        float* o = out + (size_t)p * ld_out;
        for (int i = threadIdx.x; i < g.m; i += NF_THREADS) {
            float2 y = make_float2(0.f, 0.f);
            if (i < mc) {
                if (direct) {   // y_i = sum over the non-zero cells of x_j exp(2i a_i phi_j), phase reduced in fp64 turns
                    const double a = ap[i];
                    const double t0 = a * g.phic / 3.14159265358979323846, dt = a * g.dphi / 3.14159265358979323846;
                    float2 acc = make_float2(0.f, 0.f);
                    for (int q = 0; q < nnz; ++q) {
                        double t = fma(dt, (double)nz_list[3 * q], t0);
                        t -= rint(t);
                        float sn, cs_;
                        sincospif(2.0f * (float)t, &sn, &cs_);
                        const float re = nz_list[3 * q + 1], im = nz_list[3 * q + 2];
                        acc.x += re * cs_ - im * sn;
                        acc.y += re * sn + im * cs_;
                    }
                    y = acc;
                } else {
                double kappa;
                float2 c;
                channel_geometry(g, ap[i], kappa, c);
                const int k0 = (int)ceil(kappa - 0.5 * (double)g.W);
                float2 acc = make_float2(0.f, 0.f);
                for (int l = 0; l < g.W; ++l) {
                    const int k = k0 + l;
                    const float wgt = es_kernel((float)(kappa - (double)k), inv_half_w, g.beta);
                    const float2 u = d[brev(k & (g.Nf - 1), g.logNf)];
                    acc.x = fmaf(wgt, u.x, acc.x);
                    acc.y = fmaf(wgt, u.y, acc.y);
                }
                y = cmulf(acc, c);
                }
                const float cs = chan_scale ? chan_scale[(size_t)p * g.m + i] : 1.f;
                y.x *= rs * cs;
                y.y *= rs * cs;
                if (sub) {
                    y.x -= sub[(size_t)p * ld_out + i];
                    y.y -= sub[(size_t)p * ld_out + g.m + i];
                }
            }
            o[i] = y.x;
            o[g.m + i] = y.y;
        }
    }
}

// out[p, j] = row_scale[p] sum_i chan_scale[p, i] b[p, i] exp(-2i a_i phi_j)
__global__ void __launch_bounds__(NF_THREADS, 1)
nufft_adjoint_kernel(NufftGrid g, const float* __restrict__ b, int ldb, const float* __restrict__ chan_scale,
                     const float* __restrict__ row_scale, float* __restrict__ out, int ld_out, int P, const NufftUpdate upd) {
    extern __shared__ __align__(16) float2 nsm[];
    float2* d = nsm;
    float2* tw = nsm + g.Nf;
    build_twiddles(tw, g.Nf);
    const float inv_half_w = 2.f / (float)g.W;
    const int half_n = g.n >> 1;
    for (int p = blockIdx.x; p < P; p += gridDim.x) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < g.Nf; idx += NF_THREADS) d[idx] = make_float2(0.f, 0.f);
        __syncthreads();
        const int mc = g.m_count ? min(g.m_count[p], g.m) : g.m;
        const double* ap = g.a + (size_t)p * g.ldl;
        const float* br = b + (size_t)p * ldb;
        for (int i = threadIdx.x; i < mc; i += NF_THREADS) {
            double kappa;
            float2 c;
            channel_geometry(g, ap[i], kappa, c);
            const float cs = chan_scale ? chan_scale[(size_t)p * g.m + i] : 1.f;
            const float2 v = cmulf(make_float2(br[i] * cs, br[g.m + i] * cs), make_float2(c.x, -c.y));
            const int k0 = (int)ceil(kappa - 0.5 * (double)g.W);
            for (int l = 0; l < g.W; ++l) {
                const int k = k0 + l;
                const float wgt = es_kernel((float)(kappa - (double)k), inv_half_w, g.beta);
                // (an fp32 atomicAdd on shared memory is a compare-and-swap loop: one 64-bit loop per cell instead of two 32-bit ones)
                unsigned long long* cell = reinterpret_cast<unsigned long long*>(d + brev(k & (g.Nf - 1), g.logNf));
                const float ax = wgt * v.x, ay = wgt * v.y;
                unsigned long long seen = *cell, assumed;
                do {
                    assumed = seen;
                    const float2 cur = make_float2(__uint_as_float((unsigned)(assumed & 0xffffffffull)), __uint_as_float((unsigned)(assumed >> 32)));
                    const unsigned long long want = (unsigned long long)__float_as_uint(cur.x + ax) |
                                                    ((unsigned long long)__float_as_uint(cur.y + ay) << 32);
                    seen = atomicCAS(cell, assumed, want);
                } while (seen != assumed);
            }
        }
        __syncthreads();
        fft_dit_minus(d, tw, g.logNf);
        const float rs = row_scale ? row_scale[p] : 1.f;
        if (upd.x == nullptr) {
            float* o = out + (size_t)p * ld_out;
            for (int j = threadIdx.x; j < g.n; j += NF_THREADS) {
                const int jp = j - half_n;
                const float2 v = d[jp & (g.Nf - 1)];
                const float dc = __ldg(g.deconv + j) * rs;
                o[j] = v.x * dc;
                o[g.n + j] = v.y * dc;
            }
        } else {
            const float l0 = upd.lambda0[p], nz = upd.noise[p];
            const float lam = l0 - (float)upd.iter * nz;
            const bool active = (nz < l0) && (upd.iter == 0 || lam > 0.f) && (!upd.iter_cap || upd.iter < upd.iter_cap[p]);
            const float thr = upd.step * lam;
            float* xr = upd.x + (size_t)p * upd.ld;
            float* zr = upd.z + (size_t)p * upd.ld;
            float dsum = 0.f;
            if (active) {
                // four cells per thread and pass: all eight (x, z) pairs are requested before the first store (x and z may not be
                // declared restrict against each other's stores, so a one-cell loop would serialise on its loads)
                constexpr int NB = 4;
                for (int j0 = threadIdx.x; j0 < g.n; j0 += NB * NF_THREADS) {
                    float zo[NB][2], xo[NB][2], dc[NB];
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        const int j = j0 + q * NF_THREADS;
                        const bool ok = j < g.n;
                        dc[q] = ok ? __ldg(g.deconv + j) * rs : 0.f;
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            zo[q][c] = ok ? __ldcg(zr + (c ? g.n + j : j)) : 0.f;
                            xo[q][c] = ok ? __ldcg(xr + (c ? g.n + j : j)) : 0.f;
                        }
                    }
#pragma unroll
                    for (int q = 0; q < NB; ++q) {
                        const int j = j0 + q * NF_THREADS;
                        if (j >= g.n) continue;
                        const float2 v = d[(j - half_n) & (g.Nf - 1)];
#pragma unroll
                        for (int c = 0; c < 2; ++c) {
                            const int col = c ? g.n + j : j;
                            const float grad = (c ? v.y : v.x) * dc[q];
                            const float u = zo[q][c] - upd.step * grad;
                            const float xn = copysignf(fmaxf(fabsf(u) - thr, 0.f), u);
                            const float zn = xn + upd.beta * (xn - xo[q][c]);
                            if (__float_as_uint(xn) != __float_as_uint(xo[q][c])) xr[col] = xn;
                            if (__float_as_uint(zn) != __float_as_uint(zo[q][c])) zr[col] = zn;
                            dsum += fabsf(xn - xo[q][c]);
                        }
                    }
                }
            }
            if (upd.delta != nullptr) {   // (CTA-uniform; a frozen line of sight keeps its value, like coef_update_kernel)
                __shared__ float red[NF_THREADS / 32];
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, sh);
                if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dsum;
                __syncthreads();
                if (threadIdx.x == 0 && active) {
                    float t = 0.f;
                    for (int q = 0; q < NF_THREADS / 32; ++q) t += red[q];
                    upd.delta[p] = t;
                }
            }
        }
    }
}

// 1 / psihat(j') at xi = 2 pi j' / Nf:  psihat(xi) = int psi(t) cos(xi t) dt over |t| <= W/2; with t = (W/2) sin(theta) the
// integrand is smooth, composite Simpson on 512 intervals is far below fp32 resolution
__global__ void nufft_deconv_kernel(int n, int Nf, int W, double beta, float* __restrict__ deconv) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double kPi = 3.14159265358979323846;
    const double xi = 2.0 * kPi * (double)(j - (n >> 1)) / (double)Nf, hw = 0.5 * (double)W;
    const int S = 512;
    const double h = kPi / S;
    double acc = 0.0;
    for (int q = 0; q <= S; ++q) {
        const double th = -0.5 * kPi + q * h;
        double st, ct;
        sincos(th, &st, &ct);
        const double f = exp(beta * (ct - 1.0)) * cos(xi * hw * st) * hw * ct;
        acc += f * ((q == 0 || q == S) ? 1.0 : ((q & 1) ? 4.0 : 2.0));
    }
    deconv[j] = (float)(1.0 / (acc * h / 3.0));
}

// ---- host side ---------------------------------------------------------------------------------------------------------
struct NufftPlanCache {
    int n, Nf, W, device;
    float* deconv;
};
static thread_local NufftPlanCache g_nufft_cache[4] = {};

// Nf, log2 Nf for a grid of n cells, or 0 when the transform does not fit shared memory
int nufft_grid_size(int n, int* logNf) {
    int q = 1;
    while ((1 << q) < 2 * n) ++q;
    if ((1 << q) > NF_MAX || q < 6) return 0;
    if (logNf) *logNf = q;
    return 1 << q;
}

static int nufft_setup(int n, int W, cudaStream_t stream, NufftGrid* g) {
    int device = 0;
    CSR_CUDA(cudaGetDevice(&device));
    int q = 0;
    const int Nf = nufft_grid_size(n, &q);
    CSR_REQUIRE(Nf > 0, "nufft: a grid of %d cells does not fit shared memory", n);
    NufftPlanCache* slot = nullptr;
    for (auto& c : g_nufft_cache)
        if (c.deconv && c.n == n && c.Nf == Nf && c.W == W && c.device == device) slot = &c;
    if (!slot) {
        static thread_local int next = 0;
        slot = &g_nufft_cache[next];
        next = (next + 1) % 4;
        if (slot->deconv) cudaFree(slot->deconv);
        slot->deconv = nullptr;
        CSR_CUDA(cudaMalloc(&slot->deconv, sizeof(float) * n));
        slot->n = n;
        slot->Nf = Nf;
        slot->W = W;
        slot->device = device;
        nufft_deconv_kernel<<<(n + 127) / 128, 128, 0, stream>>>(n, Nf, W, 2.30 * W, slot->deconv);
        CSR_LAUNCHED();
    }
    g->n = n;
    g->Nf = Nf;
    g->logNf = q;
    g->W = W;
    g->beta = 2.30f * (float)W;
    g->deconv = slot->deconv;
    return CSR_OK;
}

static int nufft_width() {
    static const int w = getenv("CSR_NUFFT_W") ? atoi(getenv("CSR_NUFFT_W")) : 8;
    return w < 2 ? 2 : (w > 16 ? 16 : w);
}

template <typename Kernel>
static int nufft_configure(Kernel kernel, int Nf, int* grid, int P) {
    const int smem = (int)sizeof(float2) * (Nf + tw_slot(Nf / 2) + 1) + 2 * 3 * NF_SPARSE_MAX * (int)sizeof(float);
    CSR_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    int device = 0, sms = 0;
    CSR_CUDA(cudaGetDevice(&device));
    CSR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, NF_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    *grid = min(P, sms * per_sm);
    return smem;
}

int launch_nufft_forward(const double* a, int ldl, const int* m_count, int m, int n, double phi0, double dphi, const float* x, int ldx,
                         const float* chan_scale, const float* row_scale, const float* sub, float* out, int ld_out, int P,
                         cudaStream_t stream) {
    NufftGrid g = {};
    int rc = nufft_setup(n, nufft_width(), stream, &g);
    if (rc) return rc;
    g.a = a;
    g.ldl = ldl;
    g.m_count = m_count;
    g.m = m;
    g.dphi = dphi;
    g.phic = phi0 + (double)(n >> 1) * dphi;
    int grid = 1;
    const int smem = nufft_configure(nufft_forward_kernel, g.Nf, &grid, P);
    if (smem < 0) return smem;
    cudaEvent_t ev0, ev1;
    const bool prof = profile_start(stream, &ev0, &ev1);
    static const int sparse_max = getenv("CSR_NUFFT_SPARSE") ? min(atoi(getenv("CSR_NUFFT_SPARSE")), NF_SPARSE_MAX) : NF_SPARSE_MAX;
    nufft_forward_kernel<<<grid, NF_THREADS, smem, stream>>>(g, x, ldx, chan_scale, row_scale, sub, out, ld_out, P, sparse_max);
    CSR_LAUNCHED();
    if (prof) profile_stop(14, stream, ev0, ev1);
    return CSR_OK;
}

int launch_nufft_adjoint_update(const double* a, int ldl, const int* m_count, int m, int n, double phi0, double dphi, const float* b, int ldb,
                                const float* chan_scale, const float* row_scale, float* out, int ld_out, int P, const NufftUpdate& upd,
                                cudaStream_t stream);
int launch_nufft_adjoint(const double* a, int ldl, const int* m_count, int m, int n, double phi0, double dphi, const float* b, int ldb,
                         const float* chan_scale, const float* row_scale, float* out, int ld_out, int P, cudaStream_t stream) {
    NufftUpdate none = {};
    return launch_nufft_adjoint_update(a, ldl, m_count, m, n, phi0, dphi, b, ldb, chan_scale, row_scale, out, ld_out, P, none, stream);
}
// ... with the FISTA update of (x, z) in its output stage (out is not written then)
int launch_nufft_adjoint_fista(const double* a, int ldl, const int* m_count, int m, int n, double phi0, double dphi, const float* b, int ldb,
                               float* x, float* z, int ld, const float* lambda0, const float* noise, const int* iter_cap, int iter,
                               float step, float beta, float* delta, int P, cudaStream_t stream) {
    NufftUpdate upd = {x, z, ld, lambda0, noise, iter_cap, iter, step, beta, delta};
    return launch_nufft_adjoint_update(a, ldl, m_count, m, n, phi0, dphi, b, ldb, nullptr, nullptr, nullptr, 0, P, upd, stream);
}
int launch_nufft_adjoint_update(const double* a, int ldl, const int* m_count, int m, int n, double phi0, double dphi, const float* b, int ldb,
                                const float* chan_scale, const float* row_scale, float* out, int ld_out, int P, const NufftUpdate& upd,
                                cudaStream_t stream) {
    NufftGrid g = {};
    int rc = nufft_setup(n, nufft_width(), stream, &g);
    if (rc) return rc;
    g.a = a;
    g.ldl = ldl;
    g.m_count = m_count;
    g.m = m;
    g.dphi = dphi;
    g.phic = phi0 + (double)(n >> 1) * dphi;
    int grid = 1;
    const int smem = nufft_configure(nufft_adjoint_kernel, g.Nf, &grid, P);
    if (smem < 0) return smem;
    cudaEvent_t ev0, ev1;
    const bool prof = profile_start(stream, &ev0, &ev1);
    nufft_adjoint_kernel<<<grid, NF_THREADS, smem, stream>>>(g, b, ldb, chan_scale, row_scale, out, ld_out, P, upd);
    CSR_LAUNCHED();
    if (prof) profile_stop(15, stream, ev0, ev1);
    return CSR_OK;
}

}  // namespace csr
