/* csromer_b200.h -- C ABI of the B200 (sm_100a) CS-ROMER reconstruction hot path.
 *
 * The reference (miguelcarcamov/csromer) is pure Python and has no FFI; its boundary for this path is the
 * Python class API.  This header is the C-ABI a host binding (ctypes here, cgo/JNI elsewhere) talks to; each
 * entry point names the reference method it replaces (paths relative to src/csromer/ in the reference).
 *
 * Conventions
 *   - every function returns 0 on success or a negative csr_status; csr_last_error() gives the text
 *     (thread-local).  Nothing throws.  There is NO CPU fallback: without a CUDA device every compute call
 *     fails with CSR_ERR_CUDA.
 *   - all array arguments are DEVICE pointers owned by the caller; calls are asynchronous on `stream`
 *     (a cudaStream_t passed as void*).  Plans own the twiddle tables only.
 *   - real "planar" packing everywhere: a complex vector of length L is 2L floats [Re(0..L) | Im(0..L)]
 *     (utils/utilities.py:27-32).  Pixel-major: row p of a [P, ld] matrix is line of sight p; `ld` is the
 *     row pitch in elements and must be a multiple of 8.
 *   - "ws" is caller-provided scratch of at least the size the matching *_ws_bytes function returns,
 *     256-byte aligned.
 */
#ifndef CSROMER_B200_H
#define CSROMER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct csr_plan csr_plan;
typedef struct csr_wavelet csr_wavelet;

typedef enum {
    CSR_OK = 0,
    CSR_ERR_ARG = -1,       /* bad argument (null pointer, size, pitch, alignment) */
    CSR_ERR_CUDA = -2,      /* CUDA runtime / driver error, or no sm_100 device */
    CSR_ERR_WORKSPACE = -3, /* workspace too small */
    CSR_ERR_UNSUPPORTED = -4
} csr_status;

const char* csr_last_error(void);
int csr_version(void);

/* ---- plan: the shared twiddle matrix E[i,j] = exp(+2i (lambda2_i - l2_ref) phi_j) ----------------------
 * Replaces the per-call np.exp(...) of transformers/dfts/ndft.py:19,25,34,42 and the pynufft plan of
 * transformers/dfts/nufft.py:36-48.  lambda2/phi are HOST pointers (double).  Tables are built on the device
 * in fp64 and stored as scaled fp16 hi/lo pairs in both operand orientations. */
int csr_plan_create(csr_plan** out, const double* lambda2, int m, double l2_ref, const double* phi, int n,
                    int device);
void csr_plan_destroy(csr_plan* plan);
int csr_plan_dims(const csr_plan* plan, int* m, int* n, int* ld2m, int* ld2n, size_t* table_bytes);

/* ---- single transforms -------------------------------------------------------------------------------
 * csr_forward : out[p, c] = row_scale[p] * chan_scale[p?, c mod m] * sum_j E x      [P, ld2n] -> [P, ld2m]
 *               NDFT1D.forward / forward_normalized (ndft.py:17-28), NUFFT1D.forward* (nufft.py:50-59).
 * csr_adjoint : out[p, j] = row_scale[p] * sum_i conj(E) (chan_scale[p?, i] * b)    [P, ld2m] -> [P, ld2n]
 *               NDFT1D.backward / RMTF (ndft.py:30-43), NUFFT1D.backward / RMTF (nufft.py:61-77).
 * chan_scale: [m] (chan_per_pixel = 0) or [P, m] (1) or NULL (= 1); row_scale: [P] or NULL (= 1). */
size_t csr_transform_ws_bytes(const csr_plan* plan, int P);
int csr_forward(csr_plan* plan, const float* x, const float* chan_scale, int chan_per_pixel,
                const float* row_scale, float* out, int P, void* ws, size_t ws_bytes, void* stream);
int csr_adjoint(csr_plan* plan, const float* b, const float* chan_scale, int chan_per_pixel,
                const float* row_scale, float* out, int P, void* ws, size_t ws_bytes, void* stream);

/* ---- Galactic-RM de-rotation: d_i <- d_i * exp(-2i phi_gal[p] lambda2_i) (base/dataset.py:377-381) ---- */
int csr_derotate(csr_plan* plan, float* data, const float* phi_gal, int P, void* stream);

/* ---- wavelet dictionaries (dictionaries/undecimated.py, dictionaries/discrete.py; PyWavelets semantics) --
 * kind: 0 = undecimated (pywt.swt/iswt, trim_approx=True), 1 = decimated (pywt.wavedec/waverec,
 * mode="periodization"), 2..9 = decimated with the signal-extension mode discrete.py:37-39 passes on: 2 zero,
 * 3 constant, 4 symmetric, 5 reflect, 6 periodic, 7 smooth, 8 antisymmetric, 9 antireflect (bands of
 * (len + F - 1) / 2 samples per level; reconstruct returns the N samples of the signal).  The transform acts on the real planar vector of length N = 2n as ONE signal
 * (objectivefunction/priors/chi2.py:44-55).  Filters are HOST double arrays of length F (even). */
int csr_wavelet_create(csr_wavelet** out, int kind, const double* dec_lo, const double* dec_hi,
                       const double* rec_lo, const double* rec_hi, int F, int level, int N, int append_signal,
                       int norm, int device);
void csr_wavelet_destroy(csr_wavelet* wav);
int csr_wavelet_ncoef(const csr_wavelet* wav);     /* coefficient vector length (incl. appended signal) */
int csr_wavelet_levels(const csr_wavelet* wav);
int csr_wavelet_level_sizes(const csr_wavelet* wav, int* sizes, int cap); /* [cA_L, cD_L, ..., cD_1] */
size_t csr_wavelet_ws_bytes(const csr_wavelet* wav, int P);
/* decompose: signal [P, ld_sig] -> coefficients [P, ld_coef];  reconstruct: the inverse */
int csr_wavelet_decompose(csr_wavelet* wav, const float* signal, int ld_sig, float* coef, int ld_coef, int P,
                          void* ws, size_t ws_bytes, void* stream);
int csr_wavelet_reconstruct(csr_wavelet* wav, const float* coef, int ld_coef, float* signal, int ld_sig, int P,
                            void* ws, size_t ws_bytes, void* stream);

/* ---- fused fixed-iteration FISTA (optimization/methods/fista.py:41-108 with Chi2.calculate_gradient_fista,
 * objectivefunction/priors/chi2.py:43-56, and L1.calculate_prox, priors/l1.py:30-32), all P lines of sight
 * at once.
 *   data      [P, ld2m]  measured polarisation, planar                                (Dataset.data)
 *   w         [m] or [P, m]  weights, 0 = flagged                                      (Dataset.w)
 *   s         [m]            spectral factor                                           (Dataset.s)
 *   k         [P]            normaliser sum(w/s) as held by the dataset object         (Dataset.k)
 *   x         [P, ldx]   in: initial guess; out: solution.  ldx = ld2n (delta basis) or the coefficient
 *                         pitch (wavelet basis; wav != NULL)                           (Parameter.data)
 *   lambda0, noise [P]   regularisation and its per-iteration decrement                (L1.reg, FISTA.noise)
 *   iters                common iteration cap (FISTA.maxiter); a line of sight whose lambda would become
 *                         <= 0 freezes (fista.py:100-106); noise >= lambda0 leaves x untouched (:74-77)
 *   step                 gradient step (1.0 = reference behaviour)
 * outputs (each optional, may be NULL):
 *   f_dirty   [P, ld2n]  Chi2.F_dirty = backward(data)                                 (chi2.py:18-19)
 *   model     [P, ld2m]  Dataset.model_data after the final evaluate                   (chi2.py:28)
 *   residual  [P, ld2m]  Dataset.residual = data - model_data                          (dataset.py:374-375)
 *   cost      [P, 2]     chi2 value, L1 value of the final x (F = chi2 + lambda_final * l1)
 *   delta     [P]        sum |x_new - x_old| of the last iteration / len(x)            (fista.py:89)
 */
typedef struct {
    int P;
    int iters;
    float step;
    float norm_factor;
    int w_per_pixel;
    const float* data;
    const float* w;
    const float* s;
    const float* k;
    float* x;
    int ldx;
    const float* lambda0;
    const float* noise;
    const int* iter_cap;    /* [P] optional per-line-of-sight iteration cap (floor(lambda/noise) when
                               FISTA.maxiter is None, fista.py:59-66); NULL = `iters` for everyone */
    float* f_dirty;
    float* model;
    float* residual;
    float* cost;
    float* delta;
    int use_dirty_as_guess; /* 1: ignore x on input and start from F_dirty (README.md:186) / its coefficients */
    unsigned long long* mma_blocks; /* optional [CSR_PROFILE_KINDS] (device): tensor-core work the in-loop GEMMs issued, in
                               units of one 128 x 256 x 64 multiply-accumulate block, per launch kind in
                               csr_profile_collect's numbering (three-product kinds issue 3 per tile and K block of 64
                               when dense, an fp8 K block of 128 counts 2); the executed-work side of the roofline */
    int s_per_pixel;        /* 1: s is [P, m] (per-line-of-sight samplings, csr_fista_nu); 0: [m] */
    const float* f_dirty_in; /* optional [P, ld2n]: the dirty spectrum of THESE data and weights, already computed by
                               csr_adjoint (what NDFT1D.backward(dataset.data) returned, README.md:183-186): the loop does
                               not compute it again; with f_dirty == NULL it is also what F_dirty refers to afterwards */
} csr_fista_args;

size_t csr_fista_ws_bytes(const csr_plan* plan, const csr_wavelet* wav, int P);
int csr_fista(csr_plan* plan, csr_wavelet* wav, const csr_fista_args* args, void* ws, size_t ws_bytes,
              void* stream);

/* ---- per-line-of-sight samplings (kernel K4) -----------------------------------------------------------------
 * When every line of sight has its OWN lambda^2 list (channels deleted per pixel by a flagger,
 * transformers/flaggers/mean_flagger.py:36-40; one Dataset per pixel in README.md:327-332) no twiddle matrix can be
 * shared.  These entry points evaluate the same operator (transformers/dfts/ndft.py:17-43, nufft.py:50-77) per line
 * of sight on the uniform grid phi_j = phi0 + j dphi (reconstruction/parameter.py:105-116) with twiddles generated
 * on the fly -- no plan, no table.
 *   lambda2  DEVICE fp64 [P, ldl]: lambda2_i - l2_ref of each line of sight; m_count [P] (or NULL = m) is the number
 *            of valid leading channels of each row, the rest is padding (output 0 / ignored on input).
 * csr_nudft_forward: out[p,i] = row_scale[p] chan_scale[p,i] sum_j x[p,j] exp(+2i a_i phi_j)   [P, ldx>=2n] -> [P, ld_out>=2m]
 * csr_nudft_adjoint: out[p,j] = row_scale[p] sum_i chan_scale[p,i] b[p,i] exp(-2i a_i phi_j)   [P, ldb>=2m] -> [P, ld_out>=2n]
 * chan_scale is [P, m] or NULL, row_scale [P] or NULL.
 * csr_fista_nu: csr_fista for such samplings; args->w and args->s must be [P, m] (w_per_pixel = s_per_pixel = 1);
 *            pitches are ld2m = roundup(2m, 8), ld2n = roundup(2n, 8), ldx = ld2n or roundup(ncoef, 8). */
typedef struct {
    const double* lambda2;
    int ldl;
    const int* m_count;
    int m, n;
    double phi0, dphi;
} csr_nu_sampling;
int csr_nudft_forward(const csr_nu_sampling* s, const float* x, int ldx, const float* chan_scale,
                      const float* row_scale, float* out, int ld_out, int P, void* stream);
int csr_nudft_adjoint(const csr_nu_sampling* s, const float* b, int ldb, const float* chan_scale,
                      const float* row_scale, float* out, int ld_out, int P, void* stream);
size_t csr_fista_nu_ws_bytes(const csr_nu_sampling* s, const csr_wavelet* wav, int P);
int csr_fista_nu(const csr_nu_sampling* s, csr_wavelet* wav, const csr_fista_args* args, void* ws, size_t ws_bytes,
                 void* stream);

/* ---- after the loop: the cube recipe's remaining per-line-of-sight steps ("next" rows) ----
 * csr_edge_noise  : noise[p] = eta/2 (std Re F[edges] + std Im F[edges]), numpy.std semantics, over the phi cells whose
 *                   byte in edge_mask[n] is non-zero                                      (README.md:336-338)
 * csr_convolve_phi: out = conv_same(x, taps) (+ add), real and imaginary halves separately: Parameter.convolve with
 *                   the Gaussian restore beam (reconstruction/parameter.py:139-169); taps is a HOST array (odd count
 *                   <= 255); `add` = F_residual gives the restored spectrum (csromer_reconstructor.py:213)
 * csr_peak        : argmax_j |F_j| and stats[p] = {|F| at the peak, parabola offset, parabola peak}
 *                   (csromer_reconstructor.py:59-84, 134-136) */
int csr_edge_noise(const float* F, int ld, int n, int P, const unsigned char* edge_mask, float eta, float* noise,
                   void* stream);
int csr_convolve_phi(const float* x, int ld, int n, int P, const float* taps_host, int ntaps, const float* add,
                     float* out, void* stream);
int csr_peak(const float* F, int ld, int n, int P, int* argmax, float* stats, void* stream);

/* ---- cube layouts either side of the path ----------------------------------------
 * csr_pack_slab    : Stokes Q, U planes [m, npix] (frequency first: README.md:288-297, io/io_functions.py:58-84) ->
 *                    data [P, ld_out] planar rows of pixels p0 .. p0+P-1; non-finite samples become 0 and are reported
 *                    in valid[P, m] (optional; 1 = finite) for the per-pixel weights
 * csr_unpack_slab  : F [P, ld] planar -> re, im [n, npix] columns p0 .. p0+P-1, the (2, n_phi, Y, X) layout of
 *                    io/io_functions.py:151-203
 * csr_channel_stats: stats[m, 2] = {nansum, nanstd} of every channel image [Y, X] over the window [y0,yn) x [x0,xn):
 *                    utils/utilities.py:68-110 (calculate_noise, no sigma clipping); io_functions.py:15-33 drops the
 *                    channels whose nansum is 0 */
int csr_pack_slab(const float* q, const float* u, long long npix, int m, long long p0, int P, float* out, int ld_out,
                  unsigned char* valid, void* stream);
int csr_unpack_slab(const float* in, int ld, int n, long long npix, long long p0, int P, float* re, float* im,
                    void* stream);
int csr_channel_stats(const float* img, int m, int Y, int X, int x0, int xn, int y0, int yn, float* stats, void* stream);

/* The same transposes for the arrays the Python classes hold (frequency / phi FIRST, pixels last: Dataset.data (m, P)
 * complex, Parameter.data (n, P) complex or (2n, P) real -- the batch extension of base/dataset.py and
 * reconstruction/parameter.py:127-137) <-> planar pixel-major rows.  a / b are the two planes, element (c, p) at
 * plane[c * chan_stride + p * pix_stride]: Re / Im of an interleaved complex64 array (b = a + 1, pix_stride = 2,
 * chan_stride = 2 P), or the two halves of a real (2L, P) array (b = a + L P); b = NULL moves a single [L, P] plane. */
int csr_pack_planar(const float* a, const float* b, long long chan_stride, int pix_stride, int L, int P, float* out,
                    int ld_out, void* stream);
int csr_unpack_planar(const float* in, int ld, int L, int P, float* a, float* b, long long chan_stride, int pix_stride,
                      void* stream);

/* ---- per-pixel channel flagging: sigma [P, m] -> the weights w [P, m] of the outlier channels
 * are zeroed in place (the per-pixel masks csr_fista consumes with w_per_pixel = 1); flagged [P] (optional) counts them.
 * csr_flag_mean  : transformers/flaggers/mean_flagger.py:10-46 -- outlier <=> sigma > center + nsigma std / sqrt(m), center =
 *                  mean(sigma) of the line of sight or mean_sigma[p] when given
 * csr_flag_hampel: transformers/flaggers/hampel_flagger.py:9-21 -- |sigma - median(smooth)| > nsigma 1.4826 MAD(smooth),
 *                  smooth = moving average over `window` channels ("valid" part); m - window + 1 <= 8192 */
int csr_flag_mean(const float* sigma, int P, int m, const float* mean_sigma, float nsigma, float* w, int* flagged, void* stream);
int csr_flag_hampel(const float* sigma, int P, int m, int window, float nsigma, float* w, int* flagged, void* stream);

/* ---- Galactic rotation-measure foreground of an image (faraday_sky/faraday_sky.py:62-141):
 * pixel -> sky (zenithal TAN / SIN WCS) -> Galactic -> HEALPix (RING or NESTED) -> Faraday-sky map value (nearest pixel).
 * rm_mean / rm_std [ny, nx] are the per-pixel phi_gal csr_derotate removes.  wcs: degrees, FITS 1-based CRPIX, CD matrix
 * (= CDELT x PC); proj 0 = TAN, 1 = SIN; lonpole normally 180.  swap_axes = 1 reproduces the reference's argument order
 * (array_index_to_world(xx, yy): element [r, c] is evaluated at pixel x = r, y = c); galactic_frame = 1: the WCS is
 * already in Galactic coordinates (GLON / GLAT).
 * csr_healpix_lookup: the same gather for n explicit positions (radians; icrs = 1: right ascension / declination,
 * 0: Galactic l / b) -- FaradaySky.galactic_rm (faraday_sky.py:62-97); out_pix (optional) receives the HEALPix indices. */
typedef struct {
    double crval1, crval2, crpix1, crpix2, cd11, cd12, cd21, cd22, lonpole;
    int proj;
} csr_wcs;
int csr_galactic_rm_image(const float* map_mean, const float* map_std, int nside, int nested, const csr_wcs* wcs, int nx, int ny,
                          int swap_axes, int galactic_frame, float* rm_mean, float* rm_std, void* stream);
int csr_healpix_lookup(const float* map_mean, const float* map_std, int nside, int nested, const double* lon, const double* lat,
                       int icrs, int n, float* out_mean, float* out_std, long long* out_pix, void* stream);

/* ---- compressed-row read-back of the L1 solution (README.md:374-384 writes it into dense planes; > 99 % of it is
 * zero after the soft threshold of objectivefunction/priors/l1.py:30-32) -------------------------------------------
 * csr_sparse_count: counts[p] = number of non-zeros among x[p, 0 .. ncols)
 * csr_sparse_fill : idx / val [offsets[p], offsets[p+1]) = columns and values of the non-zeros of row p, ascending;
 *                   offsets [P + 1] = exclusive prefix sum of counts (64-bit); entries at or beyond `capacity` are not
 *                   written (lists sized before offsets[P] was read back: compare the two afterwards) */
int csr_sparse_count(const float* x, int ld, int ncols, int P, int* counts, void* stream);
int csr_sparse_fill(const float* x, int ld, int ncols, int P, const long long* offsets, int* idx, float* val,
                    long long capacity, void* stream);

/* per-launch timing of the transform GEMMs with CUDA events on the launching stream (measurement aid).
 * csr_profile(1) starts recording; csr_profile_collect sums elapsed ms and launch counts into CSR_PROFILE_KINDS-entry
 * arrays indexed [0 plain-forward, 1 plain-adjoint, 2 resid-forward, 3 resid-adjoint, 4 fista-forward,
 * 5 fista-adjoint, 6 fused wavelet synthesis, 7 fused wavelet analysis + update, 8 fp16 screening adjoint,
 * 9 fp8 screening adjoint, 10 one-product plain adjoint (wavelet-domain screening), 11 wavelet analysis + update redo pass,
 * 12 tile selection / tile-list kernels of the screening, 13 end-of-iteration row update (incremental-screening clocks,
 * operand scales), 14 / 15 shared-memory NUFFT forward / adjoint of the per-line-of-sight samplings] and clears the record. */
#define CSR_PROFILE_KINDS 16
int csr_profile(int enable);
int csr_profile_collect(double* ms_by_kind, long long* count_by_kind);
/* measurement aid: with CSR_TIMELINE=1 in the environment the K-block-skipping residual forward kernel stamps clock64() per
 * role and tile -- [CTA][first 16 tiles][8]: producer start, last load issued, first operand seen by the MMA thread, last MMA
 * commit issued, epilogue start, last accumulator chunk drained, epilogue done, K blocks of the tile; copies up to `cap`
 * values of the last such launch to the HOST array `out`, returns how many (tools/fwd_timeline.py). */
int csr_debug_timeline(long long* out, int cap);

/* run-time switches of the exact shortcuts, all default 1 (measurement / test aid; results are bit-identical either
 * way): "zero_skip" (FISTA epilogue skips blocks that are and stay zero), "k_skip" (forward GEMM skips K blocks whose
 * A tile is identically zero), "screen" (adjoint tiles whose state is zero are first screened with one fp16 product
 * and a rigorous error bound; only tiles that may leave zero run the three-product kernel), "screen8" (an fp8 pass in
 * front of that one), "pair" (that fp8 pass on cta_group::2 CTA pairs), "fused_wavelet" (fused shared-memory undecimated transforms),
 * "wavelet_vec4" (their four-samples-per-thread versions and the coefficient-state flags), "resid_fast" (lean residual epilogue of the forward transform on interior column
 * runs), "sym" (the pair screening kernels decide +phi and -phi together when the phi grid is antisymmetric:
 * half the K), "screen16_pair" (per-pixel weights: one fp16 screening pass on CTA pairs instead of the
 * fp8 -> fp16 cascade), "wavelet_screen" (wavelet
 * bases: one-product gradient plus a rigorous bound away from the active coefficients, exact redo of what fails),
 * "screen_incr" (incremental screening: a screened tile is tested again only when the margin it recorded may be used up --
 * RMTF-envelope bound on the drift of the gradient), "rescale" (the fp16 operand scales of z and r follow the row maxima:
 * nothing changes while a run stays in range), "pair3" (0 / 1 / 2: three-product transforms on cta_group::2 CTA pairs --
 * off / dense transforms only (default) / every whole-matrix transform), "tma_epi" (with pair3 = 2: residual epilogue staged
 * through shared memory, TMA loads of b and TMA stores of r), "fwd128" (K-block-skipping forward transform on 128 x 128
 * tiles with 16 epilogue warps; default 0), "fwd16" (delta-basis forward tiles that accumulate one chunk run on
 * dft_fwd_direct_kernel, sixteen epilogue warps on 64-column pieces; default 1), "graph" (csr_fista: a call repeated with identical arguments -- the same plan,
 * buffers, sizes and option values, what a slab loop does -- is captured once and replayed as one CUDA graph; default 1).
 * Two switches are NOT bit-neutral: "nufft" (default 1) runs the per-line-of-sight transforms (csr_nudft_*, csr_fista_nu) as
 * shared-memory NUFFTs, ~1e-6 of the result's scale away from the direct sums that nufft = 0 evaluates; "fwd_group" (default
 * 16, 0 = off) lets delta-basis forward tiles with at most that many non-zero K blocks accumulate them as ONE chunk instead
 * of window-aligned chunks of two (direct residual epilogue, dft_gemm.cu): the grouping of fp32 partial sums changes, ~1e-7 of
 * the result's scale, identically in every combination of the exact shortcuts. */
int csr_set_option(const char* name, int value);

/* number of kernels launched by this library on the calling thread since the last reset (bench accounting) */
long long csr_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* CSROMER_B200_H */
