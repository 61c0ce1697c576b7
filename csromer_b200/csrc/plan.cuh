// plan.cuh -- internal definitions shared by the translation units of libcsromer_b200.so
#pragma once
#include "common.cuh"

struct csr_plan {
    unsigned long long uid;  // unique per created plan (keys of the CUDA-graph cache: an address can be reused, a uid cannot)
    int device;
    int m, n;        // channels, Faraday-depth samples
    int ld2m, ld2n;  // row pitches (elements) of every [P, 2m] / [P, 2n] matrix, multiples of 8
    int num_sms;
    double l2_ref;
    double* d_l2;    // [m]  lambda2_i as given (device, fp64); l2_ref is subtracted in the kernels
    double* d_phi;   // [n]
    // twiddle tables, real 2m x 2n matrix T = [[cos, -sin], [sin, cos]] scaled by kTwiddleScale, fp16 hi/lo
    __half* T_hi;    // [2m rows, ld2n]  B operand of the forward GEMM  (N = 2m, K = 2n)
    __half* T_lo;
    __half* Tt_hi;   // [2n rows, ld2m]  B operand of the adjoint GEMM  (N = 2n, K = 2m)
    __half* Tt_lo;
    // fp8 (e4m3) copy of Tt scaled by kTwiddleScale8, used only by the first screening pass of the adjoint
    unsigned char* Tt8;  // [2n rows, ld8]
    int ld8;             // row pitch in bytes of every [*, 2m] fp8 matrix, multiple of 16
    CUtensorMap map_fwd_hi, map_fwd_lo, map_adj_hi, map_adj_lo, map_adj8;
    CUtensorMap map_adj8_half;  // Tt8 with a 128-row box: half a B tile per CTA of a cta_group::2 pair
    CUtensorMap map_adj_hi_half;  // Tt_hi with a 128-row box (the one-product plain adjoint on CTA pairs)
    CUtensorMap map_adj_lo_half, map_fwd_hi_half, map_fwd_lo_half;  // the other tables with 128-row boxes (three-product pair kernel)
    // Mirror-symmetric screening (dft_gemm.cu): when phi_j = -phi_{n-j} the tables of +phi and -phi differ only in the
    // sign of the sines, so the rows (cos theta_{.,j}, sin theta_{.,j}), j >= n/2, of Tt -- read as a [4n, m] matrix of
    // half rows -- bound the gradient at both +phi_j and -phi_j with half the K and half the operand bytes
    int sym_ok;                   // the phi grid is exactly antisymmetric about index n/2 (and n is even)
    int sym16_ok;                 // ... and the fp16 table can be read as half rows (ld2m == 2m, m % 8 == 0)
    int uniform_ok;               // phi_j = phi_0 + j cell to rounding (Parameter.calculate_cellsize builds such grids)
    double cell;
    CUtensorMap map_sym8, map_sym16;  // half-row views [2n rows, m], pitch ld8/2 bytes / m halves, 128-row box
    size_t table_bytes;
};

struct csr_wavelet {
    unsigned long long uid;
    int device;
    int kind;  // 0 = undecimated (swt), 1 = decimated (dwt)
    int ext;   // decimated only: 0 = periodization, 1.. = signal-extension mode (csr_wavelet_create: kind - 1)
    int F, levels, N, append_signal, norm;
    int ncoef;             // total coefficient length incl. appended signal
    int sizes[34];         // band lengths  [cA_L, cD_L, ..., cD_1]
    int offsets[34];       // band offsets inside the coefficient vector (after the appended signal)
    float dec_lo[64], dec_hi[64], rec_lo[64], rec_hi[64];  // host copies (norm already applied)
};

namespace csr {

// GEMM geometry (dft_gemm.cu)
constexpr int BM = 128;  // lines of sight per tile (UMMA M)
constexpr int BN = 256;  // output columns per tile (UMMA N)
constexpr int BK = 64;   // fp16 elements per K block = one 128-byte swizzle row
constexpr int kMaxSkipK = 32 * 32 * 128;  // longest K the K-block activity bitmask covers (dft_gemm.cu)

// EPI_PLAIN1: the plain epilogue on ONE fp16 product (A_hi B_hi, one accumulation chain) -- the cheap gradient of the
// wavelet-domain screening; |EPI_PLAIN - EPI_PLAIN1| <= screen_eps per element (dft_gemm.cu)
enum EpilogueMode { EPI_PLAIN = 0, EPI_RESID = 1, EPI_FISTA = 2, EPI_SCREEN = 3, EPI_SCREEN8 = 4, EPI_PLAIN1 = 5 };

struct GemmEpilogue {
    // common
    int mode;
    int P, N, ld_out;          // valid rows / columns, output pitch
    int ld8;                   // pitch of out8 (bytes)
    int m;                     // channels (column -> channel index = c mod m) for chan_scale / a
    const float* a_scale;      // [P] power-of-two scale the A operand rows were multiplied by
    // EPI_PLAIN: out = row_scale[p] * chan_scale[p?, c mod m] * y
    float* out;
    const float* row_scale;    // [P] or null
    const float* chan_scale;   // [m] / [P, m] or null
    int chan_per_pixel;
    // EPI_RESID: r = a[p?, c mod m] * y - b[p, c]  ->  fp16 hi/lo scaled by out_scale[p]
    const float* b;            // [P, ld_out]
    __half* out_hi;
    __half* out_lo;
    const float* out_scale;    // [P]
    // EPI_FISTA: u = z - step*y ; x = soft(u, step*lam) ; z = x + beta (x - x_old)
    float* x;
    float* z;                  // z also written as fp16 hi/lo (out_hi/out_lo, scaled by out_scale)
    const float* lambda0;
    const float* noise;
    const int* iter_cap;       // [P] per-line-of-sight iteration cap or null
    int iter;
    float step, beta;
    float* delta;              // [P] sum |x - x_old| (atomicAdd), may be null
    unsigned char* zero_flags; // [P, flag_ld] 1 = block of 128 columns may hold non-zeros; null = no shortcut
    int flag_ld;
    unsigned int* tile_flags;  // [MT, flag_ld] written by EPI_FISTA: byte q of a word = "rows 32q..32q+31 of pixel tile
                               // mt may hold non-zeros in this 128-column block"; the next forward GEMM reads it as
                               // a_kflags and skips K blocks whose A tile is identically zero (exact)
    // K-block skipping (any mode): word != 0 <=> the A tile (128 rows x 128 K columns) may hold non-zeros; null = dense
    const unsigned int* a_kflags;
    int a_kflag_ld;
    int no_kskip;                   // 1 = the flags are there for group_sparse only: walk every K block
    int skip_single;                // 1 = leave the pixel tiles marked in single_rt to dft_fwd_direct_kernel
    unsigned char* single_rt;       // scratch of the sixteen-warp forward path: [MT] marks + one int (dft_gemm.cu); null = not used
    int group_sparse;               // EPI_RESID forward: tiles with at most this many active K blocks (by a_kflags) accumulate
                                    // them as one chunk and take the direct epilogue (0 = window-aligned chunks everywhere)
    // Safe screening of the FISTA adjoint (exact, see dft_gemm.cu "screening"):
    float* abs_sum_out;             // EPI_RESID: [P] += sum_c |out_hi[p, c]| (atomicAdd; zeroed by the caller)
    unsigned char* out8;            // EPI_RESID: optional fp8 e4m3 copy of out_hi * kOperandScale8, pitch ld8
    const float* abs_sum;           // EPI_SCREEN: [P] that sum for the rows of the A operand
    const unsigned int* tile_state; // EPI_SCREEN: [MT, flag_ld] tile_flags of the current x / z
    int* out_list;                  // EPI_SCREEN: tiles the full three-product kernel still has to process ...
    int* out_count;                 //             ... and their number (zeroed by the caller)
    const int* tile_list;           // any mode: process tiles tile_list[0 .. *tile_count) instead of all MT*NT; null = all
    const int* tile_count;
    int sym_rows8;                  // mirror-symmetric screening: the A tile arrives as [8 groups][alpha rows of 8 LOS | beta rows of
                                    // the same 8 LOS] (4-D tensor map), so one thread holds both rows of a line of sight
    const unsigned char* symq;      // mirror-symmetric screening: per (64-LOS group, symmetric tile) bits of the ranges with state
    unsigned char* need;            // mirror-symmetric screening: [MT, need_ld] bytes, 1 = tile must go on (zeroed by the caller)
    int need_ld;
    // Row statistics of the running iteration (zeroed by row_update_kernel): [p][0] += sum |z_new - z_old|, [1] += sum |z_new|,
    // [2] = max |z_new| (bits, atomicMax) -- written by EPI_FISTA; [3] = max |r * out_scale| (bits) -- written by EPI_RESID.
    float* row_stat;
    // Incremental safe screening (mirror-symmetric pair kernels; see "incremental screening" in dft_gemm.cu):
    const float* qclock;            // [P] additive row clock Q_p = it * noise_p + rhat_p (monotone)
    const float* rhat;              // [P] running maximum of the rounding allowance
    float2* expiry;                 // [MTg * NTs][64] per (tile, row): {X = margin - 2 rhat + Q as of the test, path length L0 then}
    int2* ext;                      // EPI_FISTA: [P] phi cells with state (IncrState::ext), widened when a block turns on
    unsigned long long* mma_count;  // optional: [CSR_PROFILE_KINDS], slot count_slot += K blocks x products issued
    int count_slot;                 // profile kind of this launch (set by the launchers)
    int debug;                 // measurement aid (CSR_DEBUG_EPI): bit 0 skips epilogue loads, bit 1 skips stores
    int timeline;              // measurement aid (CSR_TIMELINE): the residual forward kernel records role time stamps
    int no_rows;               // measurement aid (CSR_NO_RESID_ROWS): direct residual epilogue without the row transposition
    int no_fast;               // 1 = keep the general residual epilogue everywhere (option resid_fast = 0; same bits)
};

constexpr int kCoefBlockShift = 9;  // 512 coefficients per flag
constexpr int kCoefFlagBands = 16;  // fused kernels keep the flags of at most this many bands in shared memory

// FISTA update applied to each wavelet coefficient as its gradient value is produced (wavelet.cu; may be disabled)
struct CoefUpdate {
    float* x;
    float* z;
    const float* lambda0;
    const float* noise;
    const int* iter_cap;
    int iter;
    float step, beta;
    float* delta;
    // coefficient-state flags (exact shortcut): byte [p, band * nb + (sample >> kCoefBlockShift)] = 0 means "x and z are
    // identically zero in this block" -- no state loads, and no stores while the new values are zero again
    const unsigned char* flags_in;
    unsigned char* flags_out;  // zeroed by the caller; set to 1 where non-zero values were written
    int flag_ld;
    // wavelet-domain screening of the adjoint (fista.cu, needs the flags).  screen_mode 0: the gradient is exact
    // everywhere.  1: it is exact only where a set state flag can see it; a zero-state group whose coefficients provably
    // stay below the threshold -- |c| + band_l1 * eps[p] < lambda -- is done, any other marks its block in `redo` and is
    // left untouched.  2: only groups of blocks marked in `redo` are updated (their support is exact by now).
    int screen_mode;
    const float* eps;                 // [P] bound on the error of an inexact gradient element
    float band_l1[kCoefFlagBands];    // l1 norm of the cascaded analysis filter of each band [cA_L, cD_L, ..., cD_1]
    unsigned char* redo;              // [P, flag_ld] like the flags; zeroed by the caller before the mode-1 pass
};

int make_operand_map(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows, int elem_bytes = 2);
int launch_dft_gemm(const csr_plan* plan, bool adjoint, const __half* a_hi, const __half* a_lo, int P,
                    const GemmEpilogue& epi, cudaStream_t stream);
// eps[p] = bound on |three-product - one-product| adjoint output of line of sight p, in output units
// Row clocks of the incremental screening.  The gradient of row p at a phi cell D cells away from every cell whose state
// changed moves by at most  coef(D) * (path length of z_p),  coef(D) = crow.x * env[D] + crow.y  (env = suffix maximum of
// the normalised |RMTF| of the base channel set, crow.x = |norm_factor| m_base / n, crow.y = |norm_factor| dead_p / n for
// the channels a pixel flags on top of the base set).
struct IncrState {
    float* pathlen;         // [P] L_p = sum over iterations of ||z(s+1) - z(s)||_1 (rounded up)
    float* qclock;          // [P] Q_p = it * noise_p + rhat_p
    float* rhat;            // [P]
    int2* ext;              // [P] union since iteration 1 of the phi-cell ranges [lo, hi] that held state (lo > hi: none)
    const float2* crow;     // [P]
    const float* env;       // [n]
};
// end of an iteration: consume row_stat into the row clocks of the incremental screening and follow the operand scales
struct RowUpdate {
    float* row_stat;        // [P, 4]
    const float* noise;     // [P] lambda's decrement per iteration
    const float* abs_sum;   // [P] sum |r_hi| of this iteration (scaled units), may be null
    float* rscale;          // [P] scale of the residual operand (may be lowered for the next forward transform)
    float* zscale_read;     // [P] scale the stored z operand was written with (read by the next forward transform)
    float* zscale_write;    // [P] scale EPI_FISTA writes the next z operand with
    IncrState incr;         // pathlen == nullptr: no incremental screening
    const unsigned char* zero_flags;  // [P, flag_ld] state flags per 128 columns of the [Re | Im] row
    int flag_ld, n, iter;
    int rescale;            // follow the row maxima with the operand scales
};
int launch_row_update(const RowUpdate& u, int P, cudaStream_t stream);
int launch_screen_eps(const float* abs_sum, const float* a_scale, int K, float* eps, int P, cudaStream_t stream);
// need: [MT, NT] bytes.  Tiles with need != 0 (and exclude == 0, if given) -> list_a; with both_lists the others -> list_b
int launch_split_tile_lists(const unsigned char* need, const unsigned char* exclude, int P, int N, int* list_a, int* count_a,
                            int* list_b, int* count_b, cudaStream_t stream);
int launch_screen8_pair(const csr_plan* plan, const unsigned char* a8, int P, const GemmEpilogue& epi, cudaStream_t stream);
// mirror-symmetric screening on CTA pairs (plan->sym_ok): fp8 (a = fp8 residual copy) or fp16 (a = r_hi, plan->sym16_ok);
// tiles that are not quiet or fail the test end up in epi.out_list / out_count; `need` is scratch of sym_scratch_bytes()
size_t sym_scratch_bytes(const csr_plan* plan, int P);
// incr / record: incremental screening (incr->pathlen set; epi.expiry its per-tile store): tiles whose margins still hold
// are skipped; record = this test writes its margins
// pair_list / pair_count: scratch for the pair tiles that have work ((ceil(P/64)+1)/2 * (ceil(n/256)+1) ints, one int)
int launch_screen_sym_pair(const csr_plan* plan, const void* a, bool fp8, int P, const GemmEpilogue& epi, unsigned char* need,
                           int* pair_list, int* pair_count, const IncrState* incr, bool record, cudaStream_t stream);
// env[k] = 1.001 * max over k' >= k of |sum_{i in base} exp(-2i (lambda2_i - l2_ref) k' cell)| / |base|  (k = 0 .. n-1);
// base = channels with w[i] > 0 (w = shared weights [m]) or all channels (w = nullptr); scratch: [n] floats
int launch_rmtf_envelope(const csr_plan* plan, const float* w, float* env, float* scratch, cudaStream_t stream);
int launch_screen16_pair(const csr_plan* plan, const __half* a_hi, int P, const GemmEpilogue& epi, cudaStream_t stream);
// EPI_PLAIN1 on CTA pairs; epi.tile_list / tile_count name PAIRS of pixel tiles (launch_pair_list), null = all
int launch_plain1_pair(const csr_plan* plan, const __half* a_hi, int P, const GemmEpilogue& epi, cudaStream_t stream);
int launch_plain1_sym_pair(const csr_plan* plan, const __half* a_hi, int P, const GemmEpilogue& epi, cudaStream_t stream);
bool plain1_sym_available(const csr_plan* plan, const __half* a_hi, int P);
int launch_sym_pair_list(const csr_plan* plan, const unsigned char* need, int P, int* list, int* count, cudaStream_t stream);
// pairs (two pixel tiles x one column tile) with at least one existing tile whose need byte is 0
int launch_pair_list(const unsigned char* need, int P, int N, int* list, int* count, cudaStream_t stream);

}  // namespace csr
