// fista.cu -- host-side orchestration behind the C ABI: single transforms (NDFT1D / NUFFT1D methods) and the
// fused fixed-iteration FISTA loop over all lines of sight (optimization/methods/fista.py:41-108 driving
// objectivefunction/priors/chi2.py:43-56 and priors/l1.py:30-32).
//
// Delta basis, per iteration (2 kernel launches):
//   forward GEMM  (A = z as fp16 hi/lo)  epilogue: r = 1[w>0]/n * (E z) - (w/(s k)) d      -> fp16 hi/lo
//   adjoint GEMM  (A = r)                epilogue: z - step*g, soft threshold, momentum     -> x, z, z hi/lo
// Wavelet basis, per iteration: synthesis -> split -> forward GEMM -> adjoint GEMM (plain) -> analysis ->
// coefficient update.  No CPU fallback exists anywhere in this file.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "plan.cuh"

namespace csr {

int launch_split_rows(const float* in, int ld_in, int rows, int cols, const float* chan_scale, int chan_pp, int m,
                      const float* ref2, int ld_ref2, __half* hi, __half* lo, int ld_out, float* scale,
                      int fixed_scale, cudaStream_t stream);
int launch_prepare(const csr_plan* plan, const csr_fista_args& a, float* b, float* amask, float* fwd_chan, float2* crow,
                   cudaStream_t stream);
int launch_finalize(const csr_plan* plan, const csr_fista_args& a, const float* model, float* residual,
                    const float* x, int ldx, int ncoef, cudaStream_t stream);
int launch_coef_update(const float* grad, float* x, float* z, int ld, int ncoef, int P, const float* lambda0,
                       const float* noise, const int* iter_cap, int iter, float step, float beta, float* delta,
                       cudaStream_t stream);
int launch_scale_vec(float* v, int n, float f, cudaStream_t stream);
int launch_state_tile_flags(const float* x, const float* z, int ld, int cols, int P, int flag_ld, unsigned int* tile_flags,
                            cudaStream_t stream);
int launch_block_flags(const float* z, int ld, int cols, int P, unsigned char* flags, int flag_ld,
                       unsigned int* tile_flags, cudaStream_t stream);
int wavelet_ld(const csr_wavelet* w);
bool wavelet_fusable(const csr_wavelet* w);
int swt_synth_fused(const csr_wavelet* w, const float* coef, int ld_coef, float* sig, int ld_sig, __half* hi,
                    __half* lo, int ld_h, const float* scale, int P, unsigned int* tile_flags, int flag_ld,
                    const unsigned char* coef_flags, int coef_flag_ld, cudaStream_t stream);
int swt_analysis_fused(const csr_wavelet* w, const float* signal, int ld_sig, float* coef, int ld_coef, int P,
                       const CoefUpdate* update, cudaStream_t stream);
int coef_flag_ld(const csr_wavelet* w);
bool coef_flags_usable(const csr_wavelet* w, const float* x, const float* z, int ld_coef);
int launch_need_from_flags(const csr_wavelet* w, const unsigned char* flags, int flag_ld, int P, unsigned char* need,
                           int need_ld, cudaStream_t stream);
void wavelet_band_l1(const csr_wavelet* w, float* l1);
int launch_coef_flags(const csr_wavelet* w, const float* z, int ld_coef, int P, unsigned char* flags, int flag_ld,
                      cudaStream_t stream);
int wavelet_decompose(const csr_wavelet* w, const float* signal, int ld_sig, float* coef, int ld_coef, int P,
                      float* ws, cudaStream_t stream);
int wavelet_reconstruct(const csr_wavelet* w, const float* coef, int ld_coef, float* signal, int ld_sig, int P,
                        float* ws, cudaStream_t stream);

struct Bump {
    char* base;
    size_t off;
    template <typename T>
    T* take(size_t count) {
        T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
        off += align256(count * sizeof(T));
        return p;
    }
};

struct TransformWs {
    __half *a_hi, *a_lo;
    float* a_scale;
    size_t bytes;
};
static TransformWs carve_transform(const csr_plan* plan, int P, void* ws) {
    Bump b{(char*)ws, 0};
    TransformWs t;
    const int ldk = plan->ld2n > plan->ld2m ? plan->ld2n : plan->ld2m;
    t.a_hi = b.take<__half>((size_t)P * ldk);
    t.a_lo = b.take<__half>((size_t)P * ldk);
    t.a_scale = b.take<float>(P);
    t.bytes = b.off;
    return t;
}

struct FistaWs {
    float *b, *amask, *fwd_chan, *rscale, *zscale, *z, *f_dirty, *model, *sig, *gcoef, *wav_ws;
    __half *r_hi, *r_lo, *z_hi, *z_lo;
    unsigned char* zero_flags;
    unsigned int* tile_flags;
    unsigned char* single_rt;   // [MT] + one int: which pixel tiles the sixteen-warp forward kernel takes (dft_gemm.cu)
    int flag_ld;
    // safe screening of the adjoint (delta basis): [0] = number of listed tiles, [8 ...) = S_p per line of sight
    float* screen_hdr;
    int* tile_list;
    int* tile_list2;
    unsigned char* r8;  // fp8 copy of the residual operand for the first screening pass
    unsigned long long* mma_count;  // [CSR_PROFILE_KINDS]: tensor-core blocks issued by the in-loop GEMMs, per profile kind
    unsigned char* coef_flags[2];   // wavelet bases: coefficient-state flags (ping-pong), pitch coef_flag_ld(wav)
    unsigned char* redo_flags;      // wavelet-domain screening: blocks that failed the test in the speculative pass
    unsigned char* need[2];         // adjoint tiles [MT, NT] that need the exact gradient (predicted / after the test)
    int* tile_list3;
    float* eps;                     // [P] error bound of the one-product gradient
    // row statistics, operand-scale tracking and the clocks of the incremental screening (dft_gemm.cu, row_update_kernel)
    float *row_stat, *pathlen, *qclock, *rhat, *zscale_w, *env, *env_scratch;
    float2 *crow, *expiry;
    int2* ext;
    int* sym_list;                  // pair tiles the mirror-symmetric screening kernel has to process, [0] = their number
    size_t expiry_count;
    size_t bytes;
};
static FistaWs carve_fista(const csr_plan* plan, const csr_wavelet* wav, int P, void* ws) {
    Bump b{(char*)ws, 0};
    FistaWs f;
    const int ldc = wav ? round_up(wav->ncoef, 8) : plan->ld2n;
    f.b = b.take<float>((size_t)P * plan->ld2m);
    f.amask = b.take<float>((size_t)P * plan->m);
    f.fwd_chan = b.take<float>((size_t)P * plan->m);
    f.rscale = b.take<float>(P);
    f.zscale = b.take<float>(P);
    f.z = b.take<float>((size_t)P * ldc);
    f.f_dirty = b.take<float>((size_t)P * plan->ld2n);
    f.model = b.take<float>((size_t)P * plan->ld2m);
    f.r_hi = b.take<__half>((size_t)round_up(P, 8) * plan->ld2m);  // (rows rounded up: the 4-D operand view reads groups of 8)
    f.r_lo = b.take<__half>((size_t)round_up(P, 8) * plan->ld2m);
    f.z_hi = b.take<__half>((size_t)P * plan->ld2n);
    f.z_lo = b.take<__half>((size_t)P * plan->ld2n);
    f.flag_ld = (2 * plan->n + 127) / 128;
    f.zero_flags = b.take<unsigned char>((size_t)P * f.flag_ld);
    f.tile_flags = b.take<unsigned int>((size_t)((P + BM - 1) / BM) * f.flag_ld);
    f.single_rt = b.take<unsigned char>((size_t)((P + BM - 1) / BM) + 16);
    f.screen_hdr = b.take<float>((size_t)P + 8);
    f.tile_list = b.take<int>((size_t)((P + BM - 1) / BM) * ((2 * plan->n + BN - 1) / BN));
    f.tile_list2 = b.take<int>((size_t)((P + BM - 1) / BM) * ((2 * plan->n + BN - 1) / BN));
    f.r8 = b.take<unsigned char>((size_t)round_up(P, 8) * plan->ld8);
    f.mma_count = b.take<unsigned long long>(CSR_PROFILE_KINDS);
    f.need[0] = b.take<unsigned char>(sym_scratch_bytes(plan, P));
    f.row_stat = b.take<float>((size_t)4 * P);
    f.crow = b.take<float2>(P);
    f.pathlen = b.take<float>(P);
    f.qclock = b.take<float>(P);
    f.rhat = b.take<float>(P);
    f.ext = b.take<int2>(P);
    f.zscale_w = b.take<float>(P);
    f.env = b.take<float>(plan->n);
    f.env_scratch = b.take<float>(plan->n);
    f.expiry_count = wav ? 0 : (size_t)((P + BM / 2 - 1) / (BM / 2)) * ((plan->n + BN - 1) / BN + 1) * 64;
    f.expiry = b.take<float2>(f.expiry_count);
    f.sym_list = b.take<int>((size_t)(((P + BM / 2 - 1) / (BM / 2) + 1) / 2) * ((plan->n + BN - 1) / BN + 1) + 8);
    f.sig = f.gcoef = f.wav_ws = nullptr;
    f.coef_flags[0] = f.coef_flags[1] = f.redo_flags = f.need[1] = nullptr;
    f.tile_list3 = nullptr;
    f.eps = nullptr;
    if (wav) {
        f.sig = b.take<float>((size_t)P * plan->ld2n);
        f.gcoef = b.take<float>((size_t)P * ldc);
        f.wav_ws = b.take<float>((size_t)2 * P * wavelet_ld(wav));
        f.coef_flags[0] = b.take<unsigned char>((size_t)P * coef_flag_ld(wav));
        f.coef_flags[1] = b.take<unsigned char>((size_t)P * coef_flag_ld(wav));
        f.redo_flags = b.take<unsigned char>((size_t)P * coef_flag_ld(wav));
        const size_t tiles = (size_t)((P + BM - 1) / BM) * ((2 * plan->n + BN - 1) / BN);
        f.need[1] = b.take<unsigned char>(tiles);
        f.tile_list3 = b.take<int>(tiles);
        f.eps = b.take<float>(P);
    }
    f.bytes = b.off;
    return f;
}

static int transform(csr_plan* plan, bool adjoint, const float* in, const float* chan_scale, int chan_pp,
                     const float* row_scale, float* out, int P, void* ws, size_t ws_bytes, cudaStream_t stream) {
    CSR_REQUIRE(plan && in && out && P > 0, "transform: bad argument");
    TransformWs t = carve_transform(plan, P, ws);
    if (!ws || ws_bytes < t.bytes) {
        set_error("transform: workspace too small (%zu < %zu)", ws_bytes, t.bytes);
        return CSR_ERR_WORKSPACE;
    }
    CSR_CUDA(cudaSetDevice(plan->device));
    const int cols = adjoint ? 2 * plan->m : 2 * plan->n;
    const int ld = adjoint ? plan->ld2m : plan->ld2n;
    int rc;
    // adjoint: the channel factor (w/s) multiplies the INPUT (ndft.py:32-34); forward: the OUTPUT (ndft.py:26-28)
    rc = launch_split_rows(in, ld, P, cols, adjoint ? chan_scale : nullptr, chan_pp, plan->m, nullptr, 0, t.a_hi, t.a_lo,
                           ld, t.a_scale, 0, stream);
    if (rc) return rc;
    GemmEpilogue e = {};
    e.mode = EPI_PLAIN;
    e.a_scale = t.a_scale;
    e.out = out;
    e.row_scale = row_scale;
    e.chan_scale = adjoint ? nullptr : chan_scale;
    e.chan_per_pixel = chan_pp;
    return launch_dft_gemm(plan, adjoint, t.a_hi, t.a_lo, P, e, stream);
}

struct FistaWs;
static int fista_dispatch(csr_plan* plan, csr_wavelet* wav, const csr_fista_args& a, FistaWs& w, void* ws, size_t ws_bytes,
                          cudaStream_t stream);
}  // namespace csr

using namespace csr;

extern "C" size_t csr_transform_ws_bytes(const csr_plan* plan, int P) {
    if (!plan || P <= 0) return 0;
    return carve_transform(plan, P, nullptr).bytes;
}

extern "C" int csr_forward(csr_plan* plan, const float* x, const float* chan_scale, int chan_per_pixel,
                           const float* row_scale, float* out, int P, void* ws, size_t ws_bytes, void* stream) {
    return transform(plan, false, x, chan_scale, chan_per_pixel, row_scale, out, P, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int csr_adjoint(csr_plan* plan, const float* b, const float* chan_scale, int chan_per_pixel,
                           const float* row_scale, float* out, int P, void* ws, size_t ws_bytes, void* stream) {
    return transform(plan, true, b, chan_scale, chan_per_pixel, row_scale, out, P, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" size_t csr_fista_ws_bytes(const csr_plan* plan, const csr_wavelet* wav, int P) {
    if (!plan || P <= 0) return 0;
    return carve_fista(plan, wav, P, nullptr).bytes;
}

extern "C" int csr_fista(csr_plan* plan, csr_wavelet* wav, const csr_fista_args* args, void* ws, size_t ws_bytes,
                         void* stream_) {
    CSR_REQUIRE(plan && args, "csr_fista: null plan or args");
    const csr_fista_args& a = *args;
    CSR_REQUIRE(a.P > 0 && a.iters >= 0, "csr_fista: P=%d iters=%d", a.P, a.iters);
    CSR_REQUIRE(a.data && a.w && a.s && a.k && a.x && a.lambda0 && a.noise, "csr_fista: null input array");
    const int P = a.P, m = plan->m, n = plan->n;
    const int ncoef = wav ? wav->ncoef : 2 * n;
    if (wav) CSR_REQUIRE(wav->N == 2 * n, "csr_fista: wavelet length %d does not match 2n = %d", wav->N, 2 * n);
    CSR_REQUIRE(a.ldx >= ncoef && a.ldx % 8 == 0, "csr_fista: ldx=%d must be >= %d and a multiple of 8", a.ldx, ncoef);
    if (!wav) CSR_REQUIRE(a.ldx == plan->ld2n, "csr_fista: delta basis needs ldx == ld2n (%d)", plan->ld2n);
    FistaWs w = carve_fista(plan, wav, P, ws);
    if (!ws || ws_bytes < w.bytes) {
        set_error("csr_fista: workspace too small (%zu < %zu)", ws_bytes, w.bytes);
        return CSR_ERR_WORKSPACE;
    }
    if (wav) CSR_REQUIRE(round_up(wav->ncoef, 8) == a.ldx, "csr_fista: wavelet basis needs ldx == round_up(ncoef, 8)");
    cudaStream_t stream = (cudaStream_t)stream_;
    CSR_CUDA(cudaSetDevice(plan->device));
    return fista_dispatch(plan, wav, a, w, ws, ws_bytes, stream);
}

namespace csr {
static int fista_run(csr_plan* plan, csr_wavelet* wav, const csr_fista_args& a, FistaWs& w, cudaStream_t stream) {
    const int P = a.P, m = plan->m, n = plan->n;
    const int ncoef = wav ? wav->ncoef : 2 * n;
    int rc;
    const int ld2m = plan->ld2m, ld2n = plan->ld2n, ldc = a.ldx;
    float* f_dirty = a.f_dirty ? a.f_dirty : w.f_dirty;
    float* model = a.model ? a.model : w.model;

    CSR_CUDA(cudaMemsetAsync(w.mma_count, 0, CSR_PROFILE_KINDS * sizeof(unsigned long long), stream));
    // 1. diagonals and the data term b = (w/(s k)) d; its scale drives every residual operand
    if ((rc = launch_prepare(plan, a, w.b, w.amask, w.fwd_chan, w.crow, stream))) return rc;
    if ((rc = launch_split_rows(w.b, ld2m, P, 2 * m, nullptr, 0, m, nullptr, 0, w.r_hi, w.r_lo, ld2m, w.rscale, 0, stream)))
        return rc;

    // 2. dirty spectrum F_dirty = backward(data) = E^H b   (chi2.py:18-19) -- unless the caller already has it
    if (a.f_dirty_in != nullptr) {
        if (a.f_dirty != nullptr && a.f_dirty != a.f_dirty_in)
            CSR_CUDA(cudaMemcpyAsync(a.f_dirty, a.f_dirty_in, sizeof(float) * (size_t)P * ld2n, cudaMemcpyDeviceToDevice, stream));
        else
            f_dirty = const_cast<float*>(a.f_dirty_in);
    } else {
        GemmEpilogue e = {};
        e.mode = EPI_PLAIN;
        e.a_scale = w.rscale;
        e.out = f_dirty;
        e.mma_count = w.mma_count;
        if ((rc = launch_dft_gemm(plan, true, w.r_hi, w.r_lo, P, e, stream))) return rc;
    }

    // 3. initial state: x0 (given, or F_dirty / its coefficients), z = x0
    float* x = a.x;
    if (a.use_dirty_as_guess) {
        if (wav) {
            if ((rc = wavelet_decompose(wav, f_dirty, ld2n, x, ldc, P, w.wav_ws, stream))) return rc;
        } else {
            CSR_CUDA(cudaMemcpyAsync(x, f_dirty, sizeof(float) * (size_t)P * ld2n, cudaMemcpyDeviceToDevice, stream));
        }
    }
    CSR_CUDA(cudaMemcpyAsync(w.z, x, sizeof(float) * (size_t)P * ldc, cudaMemcpyDeviceToDevice, stream));
    // operand scale of the phi-space vector: from the dirty spectrum and the starting point
    if (wav) {
        if ((rc = wavelet_reconstruct(wav, w.z, ldc, w.sig, ld2n, P, w.wav_ws, stream))) return rc;
        if ((rc = launch_split_rows(w.sig, ld2n, P, 2 * n, nullptr, 0, m, f_dirty, ld2n, w.z_hi, w.z_lo, ld2n, w.zscale, 0, stream)))
            return rc;
    } else {
        if ((rc = launch_split_rows(w.z, ld2n, P, 2 * n, nullptr, 0, m, f_dirty, ld2n, w.z_hi, w.z_lo, ld2n, w.zscale, 0, stream)))
            return rc;
    }
    const bool zero_skip = options().zero_skip != 0;
    // forward GEMM skips K blocks whose A tile is identically zero (exact); option k_skip = 0 keeps the dense K loop
    const bool k_skip = zero_skip && options().k_skip != 0;
    const bool use_kskip = !wav && k_skip && 2 * n <= kMaxSkipK;
    // sparse tiles of the forward transform accumulate their few active K blocks as one chunk (dft_gemm.cu, group_sparse).
    // The grouping is read off the state flags, so runs WITHOUT the shortcuts that maintain them compute the same flags with a
    // kernel of their own: every combination of the exact shortcuts still produces the same bits.
    const int fwd_group = (!wav && 2 * n <= kMaxSkipK && options().fwd128 == 0 && options().pair3 < 2) ? options().fwd_group : 0;
    // adjoint tiles whose state is zero are screened with one product first (exact); option screen = 0 turns it off
    const bool use_screen = !wav && zero_skip && options().screen != 0;
    // Per-pixel weights: the weighted residual (1/n) 1[w>0] E z - (w/(s k)) d stays signal-sized (the reference's
    // normalised model cannot follow channel-to-channel weight changes), its l1 norm makes the fp8 bound useless (every
    // tile fails it, measured on the C4 workload) -- screen with one fp16 pass on CTA pairs instead of the fp8 -> fp16 cascade
    const bool screen16_pair = use_screen && a.w_per_pixel != 0 && options().pair != 0 && options().screen16_pair != 0 && P > BM;
    const bool use_screen8 = use_screen && !screen16_pair && options().screen8 != 0;  // fp8 pass in front of the fp16 one
    if (!wav && zero_skip) {
        if ((rc = launch_block_flags(w.z, ld2n, 2 * n, P, w.zero_flags, w.flag_ld, w.tile_flags, stream))) return rc;
    }
    // mirror-symmetric screening on CTA pairs (+phi and -phi decided together: half the K)
    const bool sym = use_screen && options().pair != 0 && options().sym != 0 && plan->sym_ok && (!screen16_pair || plan->sym16_ok);
    // incremental screening (exact): a tile that passed is tested again only when its margin may have been used up
    const bool incr = sym && (screen16_pair || use_screen8) && options().screen_incr != 0 && w.expiry_count > 0 && plan->uniform_ok;
    // the operand scales of z and r follow the row maxima (nothing changes while a run stays within the fp16 range)
    const bool rescale = options().rescale != 0;
    const bool row_stats = incr || rescale;
    if (row_stats) {
        CSR_CUDA(cudaMemsetAsync(w.row_stat, 0, sizeof(float) * 4 * (size_t)P, stream));
        CSR_CUDA(cudaMemcpyAsync(w.zscale_w, w.zscale, sizeof(float) * P, cudaMemcpyDeviceToDevice, stream));
    }
    IncrState incr_state = {};
    if (incr) {
        CSR_CUDA(cudaMemsetAsync(w.pathlen, 0, sizeof(float) * P, stream));
        CSR_CUDA(cudaMemsetAsync(w.qclock, 0, sizeof(float) * P, stream));
        CSR_CUDA(cudaMemsetAsync(w.rhat, 0, sizeof(float) * P, stream));
        CSR_CUDA(cudaMemsetAsync(w.ext, 0, sizeof(int2) * P, stream));
        CSR_CUDA(cudaMemsetAsync(w.expiry, 0, sizeof(float2) * w.expiry_count, stream));
        // |RMTF| envelope of the base channel set: the shared weights' live channels, or all channels when pixels flag their own
        if ((rc = launch_rmtf_envelope(plan, a.w_per_pixel ? nullptr : a.w, w.env, w.env_scratch, stream))) return rc;
        incr_state.pathlen = w.pathlen;
        incr_state.qclock = w.qclock;
        incr_state.rhat = w.rhat;
        incr_state.ext = w.ext;
        incr_state.crow = w.crow;
        incr_state.env = w.env;
    }
    if (use_screen8) CSR_CUDA(cudaMemsetAsync(w.r8, 0, (size_t)round_up(P, 8) * plan->ld8, stream));  // (the padding of the fp8 rows stays zero)
    if (a.delta) CSR_CUDA(cudaMemsetAsync(a.delta, 0, sizeof(float) * P, stream));

    const bool fused_wavelet = wav && wavelet_fusable(wav) && options().fused_wavelet != 0;
    const bool wav_kskip = fused_wavelet && k_skip && 2 * n <= kMaxSkipK;
    // coefficient-state flags: blocks of 512 coefficients whose x and z are identically zero need no state traffic in
    // the update kernel and no scan in the synthesis kernel (exact)
    const bool wav_flags = fused_wavelet && zero_skip && coef_flags_usable(wav, x, w.z, ldc);
    const int cfl_ld = wav ? coef_flag_ld(wav) : 0;
    int cfl_cur = 0;
    if (wav_flags && (rc = launch_coef_flags(wav, w.z, ldc, P, w.coef_flags[0], cfl_ld, stream))) return rc;
    // wavelet-domain screening of the adjoint (exact): the three-product gradient is computed only on the tiles that
    // coefficients with non-zero state can see; elsewhere ONE product and the bound |c| + ||h_band||_1 eps < lambda
    // prove that a coefficient stays zero, and the few blocks that fail are redone with the exact gradient
    const bool wav_screen = wav_flags && options().screen != 0 && options().wavelet_screen != 0;
    float band_l1[kCoefFlagBands] = {};
    if (wav_screen) wavelet_band_l1(wav, band_l1);
    const size_t n_tiles = (size_t)((P + BM - 1) / BM) * ((2 * n + BN - 1) / BN);
    const int NTc = (2 * n + BN - 1) / BN;

    // 4. iterations
    double t = 1.0;
    for (int it = 0; it < a.iters; ++it) {
        const double t0 = t;
        t = 0.5 * (1.0 + sqrt(1.0 + 4.0 * t * t));
        const float beta = (float)((t0 - 1.0) / t);
        const bool last = (it == a.iters - 1);
        if (wav && it > 0) {  // (iteration 0 reuses the split made above)
            if (fused_wavelet) {  // synthesis of all levels + operand split in one shared-memory-staged kernel
                // the synthesis kernel records which (pixel tile, 128-column block)s of the operand hold non-zeros,
                // so that the forward GEMM can skip the all-zero K blocks (exact)
                if (wav_kskip)
                    CSR_CUDA(cudaMemsetAsync(w.tile_flags, 0, sizeof(unsigned int) * (size_t)((P + BM - 1) / BM) * w.flag_ld, stream));
                if ((rc = swt_synth_fused(wav, w.z, ldc, nullptr, 0, w.z_hi, w.z_lo, ld2n, w.zscale, P,
                                          wav_kskip ? w.tile_flags : nullptr, w.flag_ld,
                                          wav_flags ? w.coef_flags[cfl_cur] : nullptr, cfl_ld, stream)))
                    return rc;
            } else {
                if ((rc = wavelet_reconstruct(wav, w.z, ldc, w.sig, ld2n, P, w.wav_ws, stream))) return rc;
                if ((rc = launch_split_rows(w.sig, ld2n, P, 2 * n, nullptr, 0, m, nullptr, 0, w.z_hi, w.z_lo, ld2n, w.zscale, 1, stream)))
                    return rc;
            }
        }
        {
            GemmEpilogue e = {};
            e.mode = EPI_RESID;
            e.a_scale = w.zscale;
            e.chan_scale = w.amask;
            e.chan_per_pixel = a.w_per_pixel;
            e.b = w.b;
            e.out_hi = w.r_hi;
            e.out_lo = w.r_lo;
            e.out_scale = w.rscale;
            if (use_kskip || (wav_kskip && it > 0)) {
                e.a_kflags = w.tile_flags;
                e.a_kflag_ld = w.flag_ld;
            }
            if (fwd_group > 0) {
                if (!zero_skip && (rc = launch_state_tile_flags(x, w.z, ld2n, 2 * n, P, w.flag_ld, w.tile_flags, stream))) return rc;
                e.a_kflags = w.tile_flags;
                e.a_kflag_ld = w.flag_ld;
                e.no_kskip = use_kskip ? 0 : 1;
                e.group_sparse = fwd_group;
                e.single_rt = w.single_rt;
            }
            if (use_screen || wav_screen) {
                CSR_CUDA(cudaMemsetAsync(w.screen_hdr, 0, sizeof(float) * ((size_t)P + 8), stream));
                e.abs_sum_out = w.screen_hdr + 8;
                if (use_screen8) {
                    e.out8 = w.r8;
                    e.ld8 = plan->ld8;
                }
            }
            e.mma_count = w.mma_count;
            if (row_stats) e.row_stat = w.row_stat;
            if ((rc = launch_dft_gemm(plan, false, w.z_hi, w.z_lo, P, e, stream))) return rc;
        }
        if (!wav && use_screen) {
            // cascade: fp8 screen over all tiles -> list 1 -> fp16 screen -> list 2 -> three-product FISTA kernel
            GemmEpilogue e = {};
            e.a_scale = w.rscale;
            e.lambda0 = a.lambda0;
            e.noise = a.noise;
            e.iter_cap = a.iter_cap;
            e.iter = it;
            e.abs_sum = w.screen_hdr + 8;
            e.tile_state = w.tile_flags;
            e.flag_ld = w.flag_ld;
            e.mma_count = w.mma_count;
            int* counts = reinterpret_cast<int*>(w.screen_hdr);
            if (incr) {
                e.qclock = w.qclock;
                e.rhat = w.rhat;
                e.expiry = w.expiry;
            }
            if (screen16_pair) {  // one fp16 pass on CTA pairs over every quiet tile
                e.mode = EPI_SCREEN;
                e.out_list = w.tile_list;
                e.out_count = counts;
                if (sym) {
                    if ((rc = launch_screen_sym_pair(plan, w.r_hi, false, P, e, w.need[0], w.sym_list + 8, w.sym_list, incr ? &incr_state : nullptr, it > 0, stream))) return rc;
                } else {
                    if ((rc = launch_screen16_pair(plan, w.r_hi, P, e, stream))) return rc;
                }
            } else {
                if (use_screen8) {
                    e.mode = EPI_SCREEN8;
                    e.out_list = w.tile_list2;
                    e.out_count = counts + 1;
                    if (sym) {
                        if ((rc = launch_screen_sym_pair(plan, w.r8, true, P, e, w.need[0], w.sym_list + 8, w.sym_list, incr ? &incr_state : nullptr, it > 0, stream))) return rc;
                    } else if (options().pair != 0 && P > BM) {
                        if ((rc = launch_screen8_pair(plan, w.r8, P, e, stream))) return rc;
                    } else {
                        if ((rc = launch_dft_gemm(plan, true, reinterpret_cast<const __half*>(w.r8), nullptr, P, e, stream))) return rc;
                    }
                    e.tile_list = w.tile_list2;
                    e.tile_count = counts + 1;
                }
                e.expiry = nullptr;  // (the second, fp16 pass of the cascade tests every tile it is handed)
                e.mode = EPI_SCREEN;
                e.out_list = w.tile_list;
                e.out_count = counts;
                if ((rc = launch_dft_gemm(plan, true, w.r_hi, w.r_lo, P, e, stream))) return rc;
            }
        }
        if (!wav) {
            GemmEpilogue e = {};
            e.mode = EPI_FISTA;
            e.a_scale = w.rscale;
            e.x = x;
            e.z = w.z;
            e.out_hi = w.z_hi;
            e.out_lo = w.z_lo;
            e.out_scale = row_stats ? w.zscale_w : w.zscale;
            if (row_stats) e.row_stat = w.row_stat;
            if (incr && it > 0) e.ext = w.ext;   // (iteration 0's end state is scanned by row_update_kernel)
            e.lambda0 = a.lambda0;
            e.noise = a.noise;
            e.iter_cap = a.iter_cap;
            e.iter = it;
            e.step = a.step;
            e.beta = beta;
            e.delta = (last && a.delta) ? a.delta : nullptr;
            e.zero_flags = zero_skip ? w.zero_flags : nullptr;
            e.flag_ld = w.flag_ld;
            e.tile_flags = zero_skip ? w.tile_flags : nullptr;
            if (use_screen) {
                e.tile_list = w.tile_list;
                e.tile_count = reinterpret_cast<const int*>(w.screen_hdr);
            }
            e.mma_count = w.mma_count;
            if ((rc = launch_dft_gemm(plan, true, w.r_hi, w.r_lo, P, e, stream))) return rc;
        } else {
            GemmEpilogue e = {};
            e.mode = EPI_PLAIN;
            e.a_scale = w.rscale;
            e.out = w.sig;  // gradient in phi space
            e.mma_count = w.mma_count;
            int* counts = reinterpret_cast<int*>(w.screen_hdr);  // [0] exact tiles, [1] one-product tiles, [2] redo tiles
            if (wav_screen) {
                if ((rc = launch_screen_eps(w.screen_hdr + 8, w.rscale, 2 * m, w.eps, P, stream))) return rc;
                CSR_CUDA(cudaMemsetAsync(w.need[0], 0, n_tiles, stream));
                if ((rc = launch_need_from_flags(wav, w.coef_flags[cfl_cur], cfl_ld, P, w.need[0], NTc, stream))) return rc;
                if ((rc = launch_split_tile_lists(w.need[0], nullptr, P, 2 * n, w.tile_list, counts, w.tile_list2, counts + 1, stream)))
                    return rc;
                GemmEpilogue e1 = e;
                e1.mode = EPI_PLAIN1;
                e1.tile_list = w.tile_list2;
                e1.tile_count = counts + 1;
                if (options().pair != 0 && options().sym != 0 && plain1_sym_available(plan, w.r_hi, P)) {
                    // mirror-symmetric on CTA pairs: +phi and -phi from one product with half the K
                    if ((rc = launch_sym_pair_list(plan, w.need[0], P, w.tile_list2, counts + 1, stream))) return rc;
                    if ((rc = launch_plain1_sym_pair(plan, w.r_hi, P, e1, stream))) return rc;
                } else if (options().pair != 0 && P > BM) {  // on CTA pairs: every pair with a tile outside the exact list
                    if ((rc = launch_pair_list(w.need[0], P, 2 * n, w.tile_list2, counts + 1, stream))) return rc;
                    if ((rc = launch_plain1_pair(plan, w.r_hi, P, e1, stream))) return rc;
                } else {
                    if ((rc = launch_dft_gemm(plan, true, w.r_hi, w.r_lo, P, e1, stream))) return rc;
                }
                e.tile_list = w.tile_list;
                e.tile_count = counts;
            }
            if ((rc = launch_dft_gemm(plan, true, w.r_hi, w.r_lo, P, e, stream))) return rc;
            if (fused_wavelet) {  // analysis of all levels + threshold + momentum in one kernel
                CoefUpdate up = {};
                up.x = x;
                up.z = w.z;
                up.lambda0 = a.lambda0;
                up.noise = a.noise;
                up.iter_cap = a.iter_cap;
                up.iter = it;
                up.step = a.step;
                up.beta = beta;
                up.delta = (last && a.delta) ? a.delta : nullptr;
                if (wav_flags) {
                    CSR_CUDA(cudaMemsetAsync(w.coef_flags[cfl_cur ^ 1], 0, (size_t)P * cfl_ld, stream));
                    up.flags_in = w.coef_flags[cfl_cur];
                    up.flags_out = w.coef_flags[cfl_cur ^ 1];
                    up.flag_ld = cfl_ld;
                    cfl_cur ^= 1;
                }
                if (wav_screen) {
                    CSR_CUDA(cudaMemsetAsync(w.redo_flags, 0, (size_t)P * cfl_ld, stream));
                    up.screen_mode = 1;
                    up.eps = w.eps;
                    for (int b = 0; b < kCoefFlagBands; ++b) up.band_l1[b] = band_l1[b];
                    up.redo = w.redo_flags;
                }
                if ((rc = swt_analysis_fused(wav, w.sig, ld2n, nullptr, ldc, P, &up, stream))) return rc;
                if (wav_screen) {  // exact gradient where a block failed the test, then those blocks only
                    CSR_CUDA(cudaMemsetAsync(w.need[1], 0, n_tiles, stream));
                    if ((rc = launch_need_from_flags(wav, w.redo_flags, cfl_ld, P, w.need[1], NTc, stream))) return rc;
                    if ((rc = launch_split_tile_lists(w.need[1], w.need[0], P, 2 * n, w.tile_list3, counts + 2, nullptr, nullptr, stream)))
                        return rc;
                    e.tile_list = w.tile_list3;
                    e.tile_count = counts + 2;
                    if ((rc = launch_dft_gemm(plan, true, w.r_hi, w.r_lo, P, e, stream))) return rc;
                    up.screen_mode = 2;
                    if ((rc = swt_analysis_fused(wav, w.sig, ld2n, nullptr, ldc, P, &up, stream))) return rc;
                }
            } else {
                if ((rc = wavelet_decompose(wav, w.sig, ld2n, w.gcoef, ldc, P, w.wav_ws, stream))) return rc;
                if ((rc = launch_coef_update(w.gcoef, x, w.z, ldc, ncoef, P, a.lambda0, a.noise, a.iter_cap, it, a.step, beta,
                                             (last && a.delta) ? a.delta : nullptr, stream)))
                    return rc;
            }
        }
        if (row_stats && !last) {  // row clocks of the incremental screening; operand scales follow the row maxima
            RowUpdate u = {};
            u.row_stat = w.row_stat;
            u.noise = a.noise;
            u.abs_sum = (use_screen || wav_screen) ? w.screen_hdr + 8 : nullptr;
            u.rscale = w.rscale;
            u.zscale_read = wav ? w.zscale_w : w.zscale;   // (wavelet bases: the synthesis kernel keeps its fixed scale)
            u.zscale_write = w.zscale_w;
            if (incr) u.incr = incr_state;
            u.zero_flags = (!wav && zero_skip) ? w.zero_flags : nullptr;
            u.flag_ld = w.flag_ld;
            u.n = n;
            u.iter = it;
            u.rescale = rescale ? 1 : 0;
            if ((rc = launch_row_update(u, P, stream))) return rc;
        }
    }
    if (a.delta && a.iters > 0) {
        if ((rc = launch_scale_vec(a.delta, P, 1.0f / (float)ncoef, stream))) return rc;
    }

    // 5. final evaluate F(x): model_data, residual, chi2 and L1 values (fista.py:107; chi2.py:21-32)
    if (a.model || a.residual || a.cost) {
        const float* xs = x;
        if (wav) {
            if ((rc = wavelet_reconstruct(wav, x, ldc, w.sig, ld2n, P, w.wav_ws, stream))) return rc;
            xs = w.sig;
        }
        if ((rc = launch_split_rows(xs, ld2n, P, 2 * n, nullptr, 0, m, nullptr, 0, w.z_hi, w.z_lo, ld2n, w.zscale, 0, stream)))
            return rc;
        GemmEpilogue e = {};
        e.mode = EPI_PLAIN;
        e.a_scale = w.zscale;
        e.out = model;
        e.row_scale = a.k;
        e.chan_scale = w.fwd_chan;
        e.chan_per_pixel = a.w_per_pixel;
        if (use_kskip) {  // x is zero wherever the loop's tile flags say "x and z are zero"
            e.a_kflags = w.tile_flags;
            e.a_kflag_ld = w.flag_ld;
        }
        e.mma_count = w.mma_count;
        if ((rc = launch_dft_gemm(plan, false, w.z_hi, w.z_lo, P, e, stream))) return rc;
        if ((rc = launch_finalize(plan, a, model, a.residual, x, ldc, ncoef, stream))) return rc;
    }
    if (a.mma_blocks) {  // executed-work accounting of every GEMM launch of this call (a measurement aid)
        CSR_CUDA(cudaMemcpyAsync(a.mma_blocks, w.mma_count, CSR_PROFILE_KINDS * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, stream));
    }
    return CSR_OK;
}

// ---- CUDA-graph replay of a repeated call -----------------------------------------------------------------------------
// A K = 100 call is ~1700 launches and memsets; between two of them the device idles for a few microseconds, about a tenth of a
// C3 step.  Nothing the host does inside the loop depends on device results (tile lists, counts and clocks all live on the
// device), so the SECOND call with identical arguments -- same plan, wavelet, pointers, sizes, iteration count and option
// values: what a cube loop does with its slabs -- is captured into a graph, and that and every later one is a single graph
// launch.  Keys hold the uid of the plan / wavelet, not just their addresses.  Profiling, the time-line aid and option graph = 0
// keep the plain launches.
bool profile_enabled();
struct FistaGraph {
    std::vector<unsigned char> key;
    cudaGraphExec_t exec;
    long long launches;
    unsigned long long used;
};
static std::vector<FistaGraph> g_graphs;
static std::vector<std::vector<unsigned char>> g_seen;
static unsigned long long g_graph_clock = 0;

static int fista_dispatch(csr_plan* plan, csr_wavelet* wav, const csr_fista_args& a, FistaWs& w, void* ws, size_t ws_bytes,
                          cudaStream_t stream) {
    static const bool timeline = getenv("CSR_TIMELINE") != nullptr;
    static const bool verbose = getenv("CSR_DEBUG_GRAPH") != nullptr;
    cudaStreamCaptureStatus capturing = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &capturing) != cudaSuccess) cudaGetLastError();
    if (options().graph == 0 || a.iters < 2 || profile_enabled() || timeline || capturing != cudaStreamCaptureStatusNone)
        return fista_run(plan, wav, a, w, stream);
    std::vector<unsigned char> key(sizeof(a) + sizeof(Options) + 6 * sizeof(unsigned long long));
    {
        unsigned char* k = key.data();
        memcpy(k, &a, sizeof(a));
        k += sizeof(a);
        memcpy(k, &options(), sizeof(Options));
        k += sizeof(Options);
        const unsigned long long extra[6] = {plan->uid, wav ? wav->uid : 0ull, (unsigned long long)(uintptr_t)ws, (unsigned long long)ws_bytes,
                                             (unsigned long long)(uintptr_t)plan, (unsigned long long)(uintptr_t)wav};
        memcpy(k, extra, sizeof(extra));
    }
    for (FistaGraph& g : g_graphs) {
        if (g.key == key) {
            g.used = ++g_graph_clock;
            CSR_CUDA(cudaGraphLaunch(g.exec, stream));
            launch_counter() += g.launches;
            if (verbose) fprintf(stderr, "csr_fista: graph replay (%lld launches)\n", g.launches);
            return CSR_OK;
        }
    }
    bool seen = false;
    for (const auto& k : g_seen) seen = seen || k == key;
    if (!seen) {  // first call with these arguments: plain launches (a one-off call never pays for a capture)
        if (g_seen.size() >= 16) g_seen.erase(g_seen.begin());
        g_seen.push_back(key);
        return fista_run(plan, wav, a, w, stream);
    }
    const long long before = launch_counter();
    // (the legacy default stream cannot be captured: record on a stream of the library's own -- a graph is not tied to the stream
    //  it was captured on -- and launch the result into the caller's)
    static cudaStream_t cap_streams[64] = {};
    cudaStream_t& cap = cap_streams[plan->device & 63];
    if (cap == nullptr && cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) != cudaSuccess) {
        cap = nullptr;
        cudaGetLastError();
        return fista_run(plan, wav, a, w, stream);
    }
    if (cudaStreamBeginCapture(cap, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
        cudaGetLastError();
        if (verbose) fprintf(stderr, "csr_fista: cudaStreamBeginCapture failed\n");
        return fista_run(plan, wav, a, w, stream);
    }
    const int rc = fista_run(plan, wav, a, w, cap);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(cap, &graph);
    const long long launches = launch_counter() - before;
    cudaGraphExec_t exec = nullptr;
    if (rc != CSR_OK || ce != cudaSuccess || graph == nullptr || cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        launch_counter() = before;
        if (verbose) fprintf(stderr, "csr_fista: capture failed (rc %d, cuda %d)\n", rc, (int)ce);
        if (rc != CSR_OK) return rc;
        return fista_run(plan, wav, a, w, stream);  // (nothing has run yet: a capture only records)
    }
    cudaGraphDestroy(graph);
    if (g_graphs.size() >= 8) {  // evict the least recently used
        size_t lru = 0;
        for (size_t i = 1; i < g_graphs.size(); ++i)
            if (g_graphs[i].used < g_graphs[lru].used) lru = i;
        cudaGraphExecDestroy(g_graphs[lru].exec);
        g_graphs.erase(g_graphs.begin() + lru);
    }
    g_graphs.push_back({key, exec, launches, ++g_graph_clock});
    if (verbose) fprintf(stderr, "csr_fista: captured a graph of %lld launches\n", launches);
    CSR_CUDA(cudaGraphLaunch(exec, stream));
    return CSR_OK;
}
}  // namespace csr
