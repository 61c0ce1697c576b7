// plan.cu -- plan lifetime and the twiddle-table builder (K1).
//
// Replaces the np.exp(2j * l2 * phi) that transformers/dfts/ndft.py:19,25,34,42 rebuilds on every call
// (~99 % of the reference's CPU time) and the pynufft plan of
// transformers/dfts/nufft.py:36-48: the matrix is built ONCE per (lambda2, phi) sampling, in fp64 like the
// reference (phase = fl(fl(2*l2_i) * phi_j), then sincos), and stored as fp16 hi/lo pairs of T * 2^12 in the
// two K-major orientations the tensor-core GEMMs read through TMA.
#include <stdarg.h>
#include <string.h>

#include <cuda_fp8.h>

#include <math.h>

#include "plan.cuh"

namespace csr {

static thread_local char g_error[512] = "";
static thread_local long long g_launches = 0;

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}
long long& launch_counter() { return g_launches; }
unsigned long long next_uid() {
    static unsigned long long uid = 0;
    return ++uid;
}

// T[c, jj] (2m x 2n): rows = [Re chan | Im chan], cols = [Re phi | Im phi]:   [[cos, -sin], [sin, cos]]
// one thread per (i, j); j fastest -> coalesced writes of the forward-orientation table.
__global__ void twiddle_fwd_kernel(const double* __restrict__ lambda2, double l2_ref, const double* __restrict__ phi,
                                   int m, int n, int ld2n, __half* __restrict__ T_hi, __half* __restrict__ T_lo) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= n) return;
    const double l2 = lambda2[i] - l2_ref;
    const double theta = (2.0 * l2) * phi[j];
    double sn, cs;
    sincos(theta, &sn, &cs);
    const float c = (float)(cs * (double)kTwiddleScale), s = (float)(sn * (double)kTwiddleScale);
    // hi/lo split from the fp64 value so that hi + lo carries ~22 bits of the exact twiddle
    __half ch = __float2half_rn(c), sh = __float2half_rn(s);
    __half cl = __float2half_rn((float)(cs * (double)kTwiddleScale - (double)__half2float(ch)));
    __half sl = __float2half_rn((float)(sn * (double)kTwiddleScale - (double)__half2float(sh)));
    const size_t r_re = (size_t)i * ld2n, r_im = (size_t)(m + i) * ld2n;
    T_hi[r_re + j] = ch;
    T_lo[r_re + j] = cl;
    T_hi[r_re + n + j] = __hneg(sh);
    T_lo[r_re + n + j] = __hneg(sl);
    T_hi[r_im + j] = sh;
    T_lo[r_im + j] = sl;
    T_hi[r_im + n + j] = ch;
    T_lo[r_im + n + j] = cl;
}

// transposed orientation Tt[jj, c] (2n x 2m); i fastest -> coalesced writes.
__global__ void twiddle_adj_kernel(const double* __restrict__ lambda2, double l2_ref, const double* __restrict__ phi,
                                   int m, int n, int ld2m, __half* __restrict__ Tt_hi, __half* __restrict__ Tt_lo,
                                   unsigned char* __restrict__ Tt8, int ld8) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= m) return;
    const double l2 = lambda2[i] - l2_ref;
    const double theta = (2.0 * l2) * phi[j];
    double sn, cs;
    sincos(theta, &sn, &cs);
    const float c = (float)(cs * (double)kTwiddleScale), s = (float)(sn * (double)kTwiddleScale);
    __half ch = __float2half_rn(c), sh = __float2half_rn(s);
    __half cl = __float2half_rn((float)(cs * (double)kTwiddleScale - (double)__half2float(ch)));
    __half sl = __float2half_rn((float)(sn * (double)kTwiddleScale - (double)__half2float(sh)));
    const size_t r_re = (size_t)j * ld2m, r_im = (size_t)(n + j) * ld2m;
    Tt_hi[r_re + i] = ch;           // Tt[j, i]       = T[i, j]       =  cos
    Tt_lo[r_re + i] = cl;
    Tt_hi[r_re + m + i] = sh;       // Tt[j, m+i]     = T[m+i, j]     =  sin
    Tt_lo[r_re + m + i] = sl;
    Tt_hi[r_im + i] = __hneg(sh);   // Tt[n+j, i]     = T[i, n+j]     = -sin
    Tt_lo[r_im + i] = __hneg(sl);
    Tt_hi[r_im + m + i] = ch;       // Tt[n+j, m+i]   = T[m+i, n+j]   =  cos
    Tt_lo[r_im + m + i] = cl;
    // e4m3 copy for the first screening pass (round to nearest, |value| <= 256)
    const unsigned char c8 = __nv_cvt_float_to_fp8((float)(cs * (double)kTwiddleScale8), __NV_SATFINITE, __NV_E4M3);
    const unsigned char s8 = __nv_cvt_float_to_fp8((float)(sn * (double)kTwiddleScale8), __NV_SATFINITE, __NV_E4M3);
    const unsigned char ns8 = __nv_cvt_float_to_fp8((float)(-sn * (double)kTwiddleScale8), __NV_SATFINITE, __NV_E4M3);
    const size_t q_re = (size_t)j * ld8, q_im = (size_t)(n + j) * ld8;
    const int h8 = ld8 / 2;  // the second half of an fp8 row starts at ld8/2 (a multiple of 16): half-row views exist
    Tt8[q_re + i] = c8;
    Tt8[q_re + h8 + i] = s8;
    Tt8[q_im + i] = ns8;
    Tt8[q_im + h8 + i] = c8;
}

}  // namespace csr

namespace csr {
Options& options() {
    static Options o = {getenv("CSR_NO_ZERO_SKIP") == nullptr, getenv("CSR_NO_K_SKIP") == nullptr,
                        getenv("CSR_NO_FUSED_WAVELET") == nullptr, getenv("CSR_NO_SCREEN") == nullptr,
                        getenv("CSR_NO_SCREEN8") == nullptr, getenv("CSR_NO_PAIR") == nullptr,
                        getenv("CSR_NO_WAVELET_VEC4") == nullptr, getenv("CSR_NO_RESID_FAST") == nullptr,
                        getenv("CSR_NO_SYM") == nullptr,
                        getenv("CSR_NO_SCREEN16_PAIR") == nullptr,
                        getenv("CSR_NO_WAVELET_SCREEN") == nullptr, getenv("CSR_NO_SCREEN_INCR") == nullptr,
                        getenv("CSR_NO_RESCALE") == nullptr,
                        getenv("CSR_NO_PAIR3") != nullptr ? 0 : (getenv("CSR_PAIR3") != nullptr ? atoi(getenv("CSR_PAIR3")) : 1),
                        getenv("CSR_NO_TMA_EPI") == nullptr, getenv("CSR_NO_NUFFT") == nullptr,
                        getenv("CSR_FWD128") != nullptr,    // (measured: no faster than the 256-column kernel; off by default)
                        getenv("CSR_FWD_GROUP") != nullptr ? atoi(getenv("CSR_FWD_GROUP")) : 16,
                        getenv("CSR_NO_GRAPH") == nullptr, getenv("CSR_NO_FWD16") == nullptr};
    return o;
}
}  // namespace csr

using namespace csr;

extern "C" int csr_set_option(const char* name, int value) {
    CSR_REQUIRE(name != nullptr, "csr_set_option: null name");
    Options& o = options();
    if (!strcmp(name, "zero_skip")) o.zero_skip = value != 0;
    else if (!strcmp(name, "k_skip")) o.k_skip = value != 0;
    else if (!strcmp(name, "fused_wavelet")) o.fused_wavelet = value != 0;
    else if (!strcmp(name, "screen")) o.screen = value != 0;
    else if (!strcmp(name, "screen8")) o.screen8 = value != 0;
    else if (!strcmp(name, "pair")) o.pair = value != 0;
    else if (!strcmp(name, "wavelet_vec4")) o.wavelet_vec4 = value != 0;
    else if (!strcmp(name, "wavelet_screen")) o.wavelet_screen = value != 0;
    else if (!strcmp(name, "screen16_pair")) o.screen16_pair = value != 0;
    else if (!strcmp(name, "sym")) o.sym = value != 0;
    else if (!strcmp(name, "resid_fast")) o.resid_fast = value != 0;
    else if (!strcmp(name, "screen_incr")) o.screen_incr = value != 0;
    else if (!strcmp(name, "rescale")) o.rescale = value != 0;
    else if (!strcmp(name, "pair3")) o.pair3 = value;   // 0 off, 1 dense transforms only (default), 2 every whole-matrix transform
    else if (!strcmp(name, "tma_epi")) o.tma_epi = value != 0;
    else if (!strcmp(name, "nufft")) o.nufft = value != 0;
    else if (!strcmp(name, "fwd128")) o.fwd128 = value != 0;
    else if (!strcmp(name, "graph")) o.graph = value != 0;
    else if (!strcmp(name, "fwd16")) o.fwd16 = value != 0;
    else if (!strcmp(name, "fwd_group")) o.fwd_group = value < 0 ? 0 : value;   // threshold in K blocks; 0 = off
    else {
        set_error("csr_set_option: unknown option '%s'", name);
        return CSR_ERR_ARG;
    }
    return CSR_OK;
}

extern "C" const char* csr_last_error(void) { return g_error; }
extern "C" int csr_version(void) { return 100; }
extern "C" long long csr_launch_count(int reset) {
    long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

extern "C" void csr_plan_destroy(csr_plan* plan) {
    if (!plan) return;
    cudaSetDevice(plan->device);
    cudaFree(plan->d_l2);
    cudaFree(plan->d_phi);
    cudaFree(plan->T_hi);
    cudaFree(plan->T_lo);
    cudaFree(plan->Tt_hi);
    cudaFree(plan->Tt_lo);
    cudaFree(plan->Tt8);
    delete plan;
}

static int plan_build(csr_plan* p, const double* lambda2, const double* phi) {
    const int m = p->m, n = p->n;
    cudaDeviceProp prop;
    CSR_CUDA(cudaGetDeviceProperties(&prop, p->device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; this library contains sm_100a code only", p->device, prop.major, prop.minor);
        return CSR_ERR_CUDA;
    }
    p->num_sms = prop.multiProcessorCount;
    CSR_CUDA(cudaMalloc(&p->d_l2, sizeof(double) * m));
    CSR_CUDA(cudaMalloc(&p->d_phi, sizeof(double) * n));
    CSR_CUDA(cudaMemcpy(p->d_l2, lambda2, sizeof(double) * m, cudaMemcpyHostToDevice));
    CSR_CUDA(cudaMemcpy(p->d_phi, phi, sizeof(double) * n, cudaMemcpyHostToDevice));
    const size_t fwd_elems = (size_t)2 * m * p->ld2n, adj_elems = (size_t)2 * n * p->ld2m;
    CSR_CUDA(cudaMalloc(&p->T_hi, fwd_elems * 2));
    CSR_CUDA(cudaMalloc(&p->T_lo, fwd_elems * 2));
    CSR_CUDA(cudaMalloc(&p->Tt_hi, adj_elems * 2));
    CSR_CUDA(cudaMalloc(&p->Tt_lo, adj_elems * 2));
    CSR_CUDA(cudaMalloc(&p->Tt8, (size_t)2 * n * p->ld8));
    CSR_CUDA(cudaMemset(p->Tt8, 0, (size_t)2 * n * p->ld8));
    p->table_bytes = (fwd_elems + adj_elems) * 4 + (size_t)2 * n * p->ld8;
    CSR_CUDA(cudaMemset(p->T_hi, 0, fwd_elems * 2));
    CSR_CUDA(cudaMemset(p->T_lo, 0, fwd_elems * 2));
    CSR_CUDA(cudaMemset(p->Tt_hi, 0, adj_elems * 2));
    CSR_CUDA(cudaMemset(p->Tt_lo, 0, adj_elems * 2));
    {
        dim3 grid((n + 255) / 256, m);
        twiddle_fwd_kernel<<<grid, 256>>>(p->d_l2, p->l2_ref, p->d_phi, m, n, p->ld2n, p->T_hi, p->T_lo);
        CSR_LAUNCHED();
    }
    {
        dim3 grid((m + 255) / 256, n);
        twiddle_adj_kernel<<<grid, 256>>>(p->d_l2, p->l2_ref, p->d_phi, m, n, p->ld2m, p->Tt_hi, p->Tt_lo, p->Tt8, p->ld8);
        CSR_LAUNCHED();
    }
    CSR_CUDA(cudaDeviceSynchronize());
    int rc;
    if ((rc = make_operand_map(&p->map_fwd_hi, p->T_hi, 2 * m, 2 * n, p->ld2n, BN))) return rc;
    if ((rc = make_operand_map(&p->map_fwd_lo, p->T_lo, 2 * m, 2 * n, p->ld2n, BN))) return rc;
    if ((rc = make_operand_map(&p->map_adj_hi, p->Tt_hi, 2 * n, 2 * m, p->ld2m, BN))) return rc;
    if ((rc = make_operand_map(&p->map_adj_lo, p->Tt_lo, 2 * n, 2 * m, p->ld2m, BN))) return rc;
    if ((rc = make_operand_map(&p->map_adj8, p->Tt8, 2 * n, p->ld8, p->ld8, BN, 1))) return rc;
    if ((rc = make_operand_map(&p->map_adj8_half, p->Tt8, 2 * n, p->ld8, p->ld8, BN / 2, 1))) return rc;
    // mirror symmetry of the phi grid (exact in fp64: sincos is odd / even, so the tables mirror exactly)
    p->sym_ok = (n % 2 == 0 && n >= 4) ? 1 : 0;
    for (int j = 1; j < n && p->sym_ok; ++j)
        if (phi[j] != -phi[n - j]) p->sym_ok = 0;
    p->sym16_ok = p->sym_ok && p->ld2m == 2 * m && m % 8 == 0;
    p->cell = n > 1 ? (phi[n - 1] - phi[0]) / (double)(n - 1) : 0.0;
    p->uniform_ok = (n > 1 && p->cell > 0.0) ? 1 : 0;
    for (int j = 0; j < n && p->uniform_ok; ++j)
        if (fabs(phi[j] - (phi[0] + j * p->cell)) > 1e-9 * p->cell) p->uniform_ok = 0;
    if (p->sym_ok) {
        if ((rc = make_operand_map(&p->map_sym8, p->Tt8, 2 * n, m, p->ld8 / 2, BN / 2, 1))) return rc;
        if (p->sym16_ok && (rc = make_operand_map(&p->map_sym16, p->Tt_hi, 2 * n, m, m, BN / 2))) return rc;
    }
    if ((rc = make_operand_map(&p->map_adj_hi_half, p->Tt_hi, 2 * n, 2 * m, p->ld2m, BN / 2))) return rc;
    if ((rc = make_operand_map(&p->map_adj_lo_half, p->Tt_lo, 2 * n, 2 * m, p->ld2m, BN / 2))) return rc;
    if ((rc = make_operand_map(&p->map_fwd_hi_half, p->T_hi, 2 * m, 2 * n, p->ld2n, BN / 2))) return rc;
    if ((rc = make_operand_map(&p->map_fwd_lo_half, p->T_lo, 2 * m, 2 * n, p->ld2n, BN / 2))) return rc;
    return CSR_OK;
}

extern "C" int csr_plan_create(csr_plan** out, const double* lambda2, int m, double l2_ref, const double* phi, int n,
                               int device) {
    CSR_REQUIRE(out && lambda2 && phi, "csr_plan_create: null pointer");
    CSR_REQUIRE(m > 0 && n > 0, "csr_plan_create: m and n must be positive (m=%d n=%d)", m, n);
    CSR_REQUIRE((long long)2 * n * 2 * m < (1LL << 40), "csr_plan_create: twiddle table too large");
    *out = nullptr;
    CSR_CUDA(cudaSetDevice(device));
    csr_plan* p = new csr_plan();
    memset(p, 0, sizeof(*p));
    p->uid = next_uid();
    p->device = device;
    p->m = m;
    p->n = n;
    p->ld2m = round_up(2 * m, 8);
    p->ld2n = round_up(2 * n, 8);
    p->ld8 = 2 * round_up(m, 16);  // [first half | zero padding | second half | zero padding]
    p->l2_ref = l2_ref;
    int rc = plan_build(p, lambda2, phi);
    if (rc != CSR_OK) {
        csr_plan_destroy(p);
        return rc;
    }
    *out = p;
    return CSR_OK;
}

extern "C" int csr_plan_dims(const csr_plan* plan, int* m, int* n, int* ld2m, int* ld2n, size_t* table_bytes) {
    CSR_REQUIRE(plan, "csr_plan_dims: null plan");
    if (m) *m = plan->m;
    if (n) *n = plan->n;
    if (ld2m) *ld2m = plan->ld2m;
    if (ld2n) *ld2n = plan->ld2n;
    if (table_bytes) *table_bytes = plan->table_bytes;
    return CSR_OK;
}
