// dft_gemm.cu -- the lambda^2 <-> phi transform of every line of sight at once, as a dense real GEMM on the
// 5th-generation tensor cores (tcgen05.mma, kind::f16, fp32 accumulators in TMEM, operands staged by TMA).
//
// Replaces: np.dot(x, np.exp(2j*l2*phi)) in transformers/dfts/ndft.py:17-36 (and the pynufft calls of
// transformers/dfts/nufft.py:50-71), with the work that follows each transform inside the FISTA loop fused
// into the epilogue: the residual/weighting of objectivefunction/priors/chi2.py:49-52 + base/dataset.py:374
// (EPI_RESID) and z - grad, the L1 prox of priors/l1.py:30-32 and the momentum step of
// optimization/methods/fista.py:80-86 (EPI_FISTA).
//
// Math.  With the planar packing [Re|Im], forward is  Y[P,2m] = X[P,2n] * T^T  and adjoint is
// G[P,2n] = R[P,2m] * T  for the real matrix T = [[cos,-sin],[sin,cos]] (2m x 2n).  Both operands are
// K-major fp16 "hi + lo" pairs (value * 2^k = hi + lo to ~2^-22), and every K block issues the three
// products  A_hi*B_hi + A_lo*B_hi + A_hi*B_lo  into one fp32 accumulator: fp32-class accuracy at one third
// of the fp16 tensor rate (the dropped A_lo*B_lo term is ~2^-22 relative).
//
// Structure (one CTA per SM, persistent over tiles, 384 threads):
//   warp 0    : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1    : MMA issuer     (one thread issues tcgen05.mma; tcgen05.commit frees smem stages / publishes
//                               the accumulator); also owns the TMEM allocation
//   warps 4-11: epilogue       (tcgen05.ld 16x256b -> register accumulators -> float2 / half2 global I/O)
// TMEM holds two 128x256 fp32 accumulators used ping-pong per K CHUNK: the tensor core's fp32 accumulate
// truncates, so partial sums are moved to registers (round-to-nearest adds) every `chunk_kb` K blocks while the
// next chunk runs; the fused epilogue of tile t overlaps the first chunks of tile t+1 the same way.
//
// Kernels in this file:
//   dft_gemm_kernel<EPI_PLAIN | EPI_RESID | EPI_FISTA>   three-product GEMM with the fused epilogues above; the forward
//                                                        direction can skip K blocks whose A tile is zero (build_kmask)
//   dft_gemm_kernel<EPI_SCREEN | EPI_SCREEN8 | EPI_PLAIN1>  one-product (fp16 / fp8) safe screening of adjoint tiles; one-product
//                                                        plain adjoint (wavelet-domain screening)
//   screen8_pair_kernel<PMODE>                           the one-product kernels on CTA pairs (cta_group::2, M = 256):
//        0  fp8 screening                     3  fp8 screening, mirror-symmetric (+phi and -phi from half the K)
//        1  fp16 plain adjoint (EPI_PLAIN1)   4  fp16 screening, mirror-symmetric
//        2  fp16 screening                    5  fp16 plain adjoint, mirror-symmetric
//   dft_gemm_pair_kernel<MODE, TMA_EPI>                  the three-product transforms on CTA pairs (dense transforms; option pair3)
//   dft_fwd_direct_kernel<CS_PP>                         the sparse tiles of the forward transform (one accumulation chunk, option
//                                                        fwd_group): accumulator read straight from TMEM by SIXTEEN epilogue warps,
//                                                        sub-pieces transposed through shared memory, row-contiguous global accesses
//                                                        (tile_single_kernel marks those tiles; dft_gemm_kernel<EPI_RESID> keeps the rest)
//   dft_fwd128_kernel                                    the forward transform on 128 x 128 tiles (experiment, option fwd128)
//   epilogue_resid_fast, resid_request / resid_compute, resid_direct_rows   residual epilogues of dft_gemm_kernel<EPI_RESID>: lean,
//                                                        software-pipelined over the row groups, direct (sparse tiles, eight warps)
//   sym_select_kernel, row_update_kernel, rmtf_abs_kernel, suffix_max_kernel   incremental safe screening (margins, row clocks, envelope)
//   split_tile_lists_kernel, pair_list_kernel, sym_pair_list_kernel, screen_eps_kernel   small helpers
// and the host-side launchers with their tensor-map cache, per-launch CUDA-event profiling and the role time line
// (CSR_TIMELINE, csr_debug_timeline).
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <cuda_fp8.h>

#include "plan.cuh"

namespace csr {

constexpr int STAGES = 2;      // three-product modes: 2 stages of 96 KiB (A_hi, A_lo, B_hi, B_lo)
constexpr int MAX_STAGES = 4;  // one-product (screening) modes: 4 stages of 48 KiB (A, B) in the same shared memory
constexpr int A_TILE_BYTES = BM * BK * 2;  // 16 KiB
constexpr int B_TILE_BYTES = BN * BK * 2;  // 32 KiB
constexpr int STAGE_BYTES = 2 * A_TILE_BYTES + 2 * B_TILE_BYTES;  // A_hi, A_lo, B_hi, B_lo = 96 KiB
constexpr int EPI_WARPS = 8;               // 4 TMEM lane quarters x 2 column halves
constexpr int EPI_COLS = BN / 2;           // accumulator columns held in registers per epilogue thread
constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;  // warpgroup 0: TMA producer, MMA issuer, 2 idle; warpgroups 1-2: epilogue
constexpr int TMEM_COLS = 512;
constexpr int KMASK_WORDS = 32;             // K-block activity bitmask: 32 x 32 bits, one bit per 128 K columns
static_assert(KMASK_WORDS * 32 * 128 == kMaxSkipK, "bitmask capacity");
constexpr int NUM_WARPS_TOTAL = 4 + EPI_WARPS;
constexpr int CS_SMEM_FLOATS = 2048;  // shared-memory copy of chan_scale for the residual epilogue (m <= 2048, shared weights)
constexpr int XP_ROW_FLOATS = 64 + 8;  // transposing buffer of the direct residual epilogue: 8 rows x 64 columns per warp, rows
                                       // padded by 32 bytes (conflict-free 8-byte fragment writes and 16-byte row reads)
constexpr int XP_WARP_BYTES = 8 * XP_ROW_FLOATS * 4 + 32 * 8;  // ... + (descale, out_scale) of the warp's 32 rows
constexpr int SMEM_MISC_BYTES = 256 + NUM_WARPS_TOTAL * KMASK_WORDS * 4 + 16;  // barriers, K masks, last_listed
constexpr int SMEM_BYTES = 1024 /*alignment slack*/ + STAGES * STAGE_BYTES + SMEM_MISC_BYTES + CS_SMEM_FLOATS * 4 + EPI_WARPS * XP_WARP_BYTES;
constexpr int RASTER_GROUP = 16;  // pixel tiles per rasterisation group (keeps A and B panels L2-resident)
static_assert(SMEM_BYTES <= 232448, "shared memory budget exceeded");

__device__ __forceinline__ void tile_coords(int t, int MT, int NT, int& mt, int& nt) {
    const int per_group = RASTER_GROUP * NT;
    const int g = t / per_group;
    const int first = g * RASTER_GROUP;
    const int gsize = min(RASTER_GROUP, MT - first);
    const int local = t - g * per_group;
    mt = first + local % gsize;
    nt = local / gsize;
}

// ---- K-block skipping (exact) ---------------------------------------------------------------------------------
// FISTA's soft threshold leaves most of phi space identically zero for every line of sight of a pixel tile, so most
// K blocks of the forward transform multiply an all-zero A tile.  The producer of the A operand records, per pixel
// tile and 128 K columns, whether anything non-zero was written (GemmEpilogue::a_kflags); the three roles of this
// kernel each turn that row of flags into a bitmask in their own shared-memory slot and walk the same list of
// active K blocks.  Skipped blocks contribute exact zeros, so results are bit-identical to the dense loop.
// Returns the number of active K blocks of the tile.
__device__ __forceinline__ int build_kmask(const unsigned int* flags_row, int nwords, int KB, uint32_t* mask, int lane) {
    int count = 0;
    for (int w0 = 0; w0 < nwords; w0 += 32) {
        const int idx = w0 + lane;
        const unsigned int v = idx < nwords ? __ldcg(flags_row + idx) : 0u;
        const uint32_t bits = __ballot_sync(0xffffffffu, v != 0u);
        if (lane == 0) mask[w0 >> 5] = bits;
        count += 2 * __popc(bits);
    }
    __syncwarp();
    // the last flag word covers a single K block when KB is odd
    if ((KB & 1) && ((mask[(KB >> 1) >> 5] >> ((KB >> 1) & 31)) & 1u)) count -= 1;
    return count;
}
// smallest active K block >= kb, or KB when there is none (cost proportional to the number of active blocks, not KB)
__device__ __forceinline__ int next_active_kb(const uint32_t* mask, int kb, int KB) {
    while (kb < KB) {
        const int blk = kb >> 1;
        const uint32_t bits = mask[blk >> 5] >> (blk & 31);
        if (bits & 1u) return kb;
        if (bits) return 2 * (blk + __ffs(bits) - 1);
        kb = ((blk >> 5) + 1) << 6;
    }
    return KB;
}

// ---- accumulator fragment --------------------------------------------------------------------------------------
// Each epilogue warp owns 32 rows (its TMEM lane quarter) x 128 columns of the tile and reads them with
// tcgen05.ld.16x256b: thread T then holds, for every 8-column block, the two consecutive columns 2(T%4), 2(T%4)+1
// of rows T/4 and T/4+8 -- the mma-style fragment in which 4 neighbouring lanes cover one 32-byte sector of a row.
// Global I/O of the fused epilogues is done straight from these registers as float2 / half2 (no shared-memory
// transpose, no shuffles).  Register index of (lh, ch, jj, rh, e):
//     row = T/4 + 8*rh + 16*lh,  col = 64*ch + 8*jj + 2*(T%4) + e,  idx = (((lh*2 + ch)*8 + jj)*2 + rh)*2 + e
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// acc += the chunk accumulator in TMEM (taddr: this warp's lane quarter, first of its 128 columns)
__device__ __forceinline__ void drain_chunk(uint32_t taddr, float (&acc)[EPI_COLS]) {
#pragma unroll
    for (int lh = 0; lh < 2; ++lh) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            float v[32];
            tmem_ld_16x256b_x8(taddr + ((uint32_t)(16 * lh) << 16) + (uint32_t)(64 * ch), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[(lh * 2 + ch) * 32 + j] += v[j];
        }
    }
}

// acc = the accumulator in TMEM (single-chain kernels: nothing to add to): the four loads are issued back to back and
// waited for once, instead of four load / wait round trips
__device__ __forceinline__ void tmem_ld_16x256b_x8_nowait(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void load_chunk(uint32_t taddr, float (&acc)[EPI_COLS]) {
    uint32_t r[EPI_COLS];
#pragma unroll
    for (int lh = 0; lh < 2; ++lh)
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
            tmem_ld_16x256b_x8_nowait(taddr + ((uint32_t)(16 * lh) << 16) + (uint32_t)(64 * ch), r + (lh * 2 + ch) * 32);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < EPI_COLS; ++j) acc[j] = __uint_as_float(r[j]);
}

// Fused epilogue on the register fragment of one thread.  row0 / col0: first row of the warp's 32 rows and first
// of its 128 columns.  Every (row, 8-column block) costs one float2 access per array; the loads of 8 blocks of a row
// are issued together before their results are used (x / z alias, the compiler cannot hoist them itself).
template <int MODE>
__device__ __forceinline__ void epilogue_tile(const GemmEpilogue& epi, const float (&acc)[EPI_COLS], int lane, int row0,
                                              int col0) {
    const int quad = lane & 3;
    int tile_nz = 0;  // EPI_FISTA: may any row of this warp hold a non-zero z in these 128 columns afterwards?
#pragma unroll
    for (int lh = 0; lh < 2; ++lh) {
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
            const int row = row0 + (lane >> 2) + 8 * rh + 16 * lh;
            const bool row_ok = row < epi.P;
            float descale = 0.f, oscale = 0.f, rs = 1.f, thr = 0.f;
            bool active = row_ok;
            if (row_ok) {
                descale = kTwiddleDescale / __ldg(epi.a_scale + row);
                if (MODE == EPI_PLAIN) {
                    rs = epi.row_scale ? __ldg(epi.row_scale + row) : 1.f;
                } else {
                    oscale = __ldg(epi.out_scale + row);
                }
                if (MODE == EPI_FISTA) {
                    const float l0 = __ldg(epi.lambda0 + row), nz = __ldg(epi.noise + row);
                    const float lam = l0 - (float)epi.iter * nz;
                    active = (nz < l0) && (epi.iter == 0 || lam > 0.f) && (!epi.iter_cap || epi.iter < __ldg(epi.iter_cap + row));
                    thr = epi.step * lam;
                    descale *= epi.step;
                }
            }
            const size_t row_off = (size_t)row * epi.ld_out;
            const float* cs_row = nullptr;
            if (MODE != EPI_FISTA && epi.chan_scale) cs_row = epi.chan_scale + (epi.chan_per_pixel ? (size_t)row * epi.m : 0);
            const bool cs_even = (epi.m & 1) == 0 && (reinterpret_cast<uintptr_t>(epi.chan_scale) & 7) == 0;
            float dsum = 0.f;
            float st_dz = 0.f, st_zabs = 0.f, st_max = 0.f;  // row statistics (GemmEpilogue::row_stat)
            // Sparsity shortcut (exact): FISTA's soft threshold leaves most of phi space identically zero.  A byte per
            // (row, 128-column block) records "x and z are all zero here"; such a block needs no loads, and if its new
            // values are zero again it needs no stores either.
            unsigned char* flag_ptr = nullptr;
            int flag_in = 1;
            if (MODE == EPI_FISTA && epi.zero_flags != nullptr && row_ok && col0 < epi.N) {
                flag_ptr = epi.zero_flags + (size_t)row * epi.flag_ld + (col0 >> 7);
                flag_in = __ldcg(flag_ptr);
            }
            {
                // one batch = the thread's 16 float2 positions of this row (all 128 columns): 32 x 8-byte loads in
                // flight per thread for the FISTA epilogue
                float2 in0[16], in1[16];
                bool ok[16];
                int nonzero = 0;
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const int col = col0 + 8 * jj + 2 * quad;  // even; N is even, so col + 1 < N too
                    ok[jj] = active && col < epi.N;
                    in0[jj] = make_float2(0.f, 0.f);
                    in1[jj] = make_float2(1.f, 1.f);
                    if (MODE == EPI_FISTA) in1[jj] = make_float2(0.f, 0.f);
                    if (ok[jj] && flag_in && !(epi.debug & 1)) {
                        if (MODE == EPI_FISTA) {
                            in0[jj] = __ldcg(reinterpret_cast<const float2*>(epi.z + row_off + col));
                            in1[jj] = __ldcg(reinterpret_cast<const float2*>(epi.x + row_off + col));
                        } else {
                            if (MODE == EPI_RESID) in0[jj] = __ldcg(reinterpret_cast<const float2*>(epi.b + row_off + col));
                            if (cs_row) {
                                const int c0 = col >= epi.m ? col - epi.m : col;
                                if (cs_even) {  // even m: the pair never straddles the halves and stays 8-byte aligned
                                    in1[jj] = __ldg(reinterpret_cast<const float2*>(cs_row + c0));
                                } else {
                                    const int c1 = col + 1 >= epi.m ? col + 1 - epi.m : col + 1;
                                    in1[jj] = make_float2(__ldg(cs_row + c0), __ldg(cs_row + c1));
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    if (!ok[jj] || (epi.debug & 2)) continue;
                    const int col = col0 + 8 * jj + 2 * quad;
                    const int ai = (((lh * 2 + (jj >> 3)) * 8 + (jj & 7)) * 2 + rh) * 2;
                    const float a0 = acc[ai] * descale, a1 = acc[ai + 1] * descale;
                    if (MODE == EPI_PLAIN) {
                        __stcs(reinterpret_cast<float2*>(epi.out + row_off + col), make_float2(a0 * rs * in1[jj].x, a1 * rs * in1[jj].y));
                    } else if (MODE == EPI_RESID) {
                        const float r0 = __fmaf_rn(in1[jj].x, a0, -in0[jj].x) * oscale, r1 = __fmaf_rn(in1[jj].y, a1, -in0[jj].y) * oscale;  // (explicit: same bits in both residual epilogues)
                        const __half2 h = __floats2half2_rn(r0, r1);
                        const float2 hf = __half22float2(h);
                        const __half2 l = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
                        *reinterpret_cast<__half2*>(epi.out_hi + row_off + col) = h;
                        *reinterpret_cast<__half2*>(epi.out_lo + row_off + col) = l;
                        dsum += fabsf(hf.x) + fabsf(hf.y);  // S_p of the screening bound
                        st_max = fmaxf(st_max, fmaxf(fabsf(r0), fabsf(r1)));
                        if (epi.out8 != nullptr) {
                            const float2 v8 = make_float2(hf.x * kOperandScale8, hf.y * kOperandScale8);
                            if (fmaxf(fabsf(v8.x), fabsf(v8.y)) > 448.f) dsum = __int_as_float(0x7f800000);  // unscreenable row
                            // (fp8 rows are [first half | pad | second half | pad], halves ld8/2 apart: plan.cu)
                            const __nv_fp8x2_storage_t pk = __nv_cvt_float2_to_fp8x2(v8, __NV_SATFINITE, __NV_E4M3);
                            unsigned char* o8 = epi.out8 + (size_t)row * epi.ld8;
                            const int h8 = epi.ld8 >> 1;
                            if (epi.m & 1) {  // odd m: the pair may straddle the halves and the second half sits at odd offsets
                                o8[col < epi.m ? col : col - epi.m + h8] = (unsigned char)(pk & 0xff);
                                o8[col + 1 < epi.m ? col + 1 : col + 1 - epi.m + h8] = (unsigned char)(pk >> 8);
                            } else {
                                *reinterpret_cast<__nv_fp8x2_storage_t*>(o8 + (col < epi.m ? col : col - epi.m + h8)) = pk;
                            }
                        }
                    } else {  // EPI_FISTA: u = z - step*g ; x = soft(u, step*lambda) ; z = x + beta (x - x_old)
                        const float u0 = in0[jj].x - a0, u1 = in0[jj].y - a1;
                        const float x0 = copysignf(fmaxf(fabsf(u0) - thr, 0.f), u0);
                        const float x1 = copysignf(fmaxf(fabsf(u1) - thr, 0.f), u1);
                        const float z0 = x0 + epi.beta * (x0 - in1[jj].x), z1 = x1 + epi.beta * (x1 - in1[jj].y);
                        dsum += fabsf(x0 - in1[jj].x) + fabsf(x1 - in1[jj].y);
                        st_dz += fabsf(z0 - in0[jj].x) + fabsf(z1 - in0[jj].y);
                        st_zabs += fabsf(z0) + fabsf(z1);
                        st_max = fmaxf(st_max, fmaxf(fabsf(z0), fabsf(z1)));
                        in1[jj] = make_float2(x0, x1);  // keep the results; whether they are stored is decided below
                        in0[jj] = make_float2(z0, z1);
                        nonzero |= (x0 != 0.f) | (x1 != 0.f) | (z0 != 0.f) | (z1 != 0.f);
                    }
                }
                if (MODE == EPI_FISTA) {
                    nonzero |= __shfl_xor_sync(0xffffffffu, nonzero, 1);
                    nonzero |= __shfl_xor_sync(0xffffffffu, nonzero, 2);
                    if ((flag_in | nonzero) && !(epi.debug & 2)) {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) {
                            if (!ok[jj]) continue;
                            const int col = col0 + 8 * jj + 2 * quad;
                            const float2 xv = in1[jj], zv = in0[jj];
                            __stcs(reinterpret_cast<float2*>(epi.x + row_off + col), xv);
                            __stcs(reinterpret_cast<float2*>(epi.z + row_off + col), zv);
                            const __half2 h = __floats2half2_rn(zv.x * oscale, zv.y * oscale);
                            const float2 hf = __half22float2(h);
                            const __half2 l = __floats2half2_rn(zv.x * oscale - hf.x, zv.y * oscale - hf.y);
                            __stcs(reinterpret_cast<unsigned int*>(epi.out_hi + row_off + col), *reinterpret_cast<const unsigned int*>(&h));
                            __stcs(reinterpret_cast<unsigned int*>(epi.out_lo + row_off + col), *reinterpret_cast<const unsigned int*>(&l));
                        }
                    }
                    if (flag_ptr != nullptr && quad == 0 && active && nonzero != flag_in) {
                        *flag_ptr = (unsigned char)nonzero;
                        if (nonzero && epi.ext != nullptr) {  // a block turned on: the phi cells with state of this row grow
                            const int nphi = epi.N >> 1, c1 = min(col0 + 127, epi.N - 1);
                            int lo, hi;
                            if (c1 < nphi) lo = col0, hi = c1;
                            else if (col0 >= nphi) lo = col0 - nphi, hi = c1 - nphi;
                            else lo = 0, hi = nphi - 1;  // the block that straddles the two halves
                            atomicMin(&epi.ext[row].x, lo);
                            atomicMax(&epi.ext[row].y, hi);
                        }
                    }
                    if (flag_ptr != nullptr) tile_nz |= active ? nonzero : flag_in;  // frozen rows keep their state
                }
            }
            if (MODE == EPI_FISTA && epi.delta != nullptr) {  // kernel-uniform branch
                dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
                dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
                if (quad == 0 && row_ok && dsum != 0.f) atomicAdd(epi.delta + row, dsum);
            }
            if (MODE == EPI_RESID && epi.abs_sum_out != nullptr) {  // kernel-uniform branch
                dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
                dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
                if (quad == 0 && row_ok && dsum != 0.f) atomicAdd(epi.abs_sum_out + row, dsum);
            }
            if ((MODE == EPI_FISTA || MODE == EPI_RESID) && epi.row_stat != nullptr) {  // kernel-uniform branch
                st_max = fmaxf(st_max, __shfl_xor_sync(0xffffffffu, st_max, 1));
                st_max = fmaxf(st_max, __shfl_xor_sync(0xffffffffu, st_max, 2));
                if (MODE == EPI_FISTA) {
                    st_dz += __shfl_xor_sync(0xffffffffu, st_dz, 1);
                    st_dz += __shfl_xor_sync(0xffffffffu, st_dz, 2);
                    st_zabs += __shfl_xor_sync(0xffffffffu, st_zabs, 1);
                    st_zabs += __shfl_xor_sync(0xffffffffu, st_zabs, 2);
                }
                if (quad == 0 && row_ok) {  // (|.| >= 0 or NaN: as unsigned bit patterns they order like the floats, NaN on top)
                    float* rs = epi.row_stat + 4 * (size_t)row;
                    if (MODE == EPI_FISTA) {
                        if (st_dz != 0.f) atomicAdd(rs, st_dz);
                        if (st_zabs != 0.f) atomicAdd(rs + 1, st_zabs);
                        if (st_max != 0.f) atomicMax(reinterpret_cast<unsigned int*>(rs + 2), __float_as_uint(st_max));
                    } else if (st_max != 0.f) {
                        atomicMax(reinterpret_cast<unsigned int*>(rs + 3), __float_as_uint(st_max));
                    }
                }
            }
        }
    }
    if (MODE == EPI_FISTA && epi.tile_flags != nullptr && epi.zero_flags != nullptr && col0 < epi.N) {  // warp-uniform
        const int any_nz = __any_sync(0xffffffffu, tile_nz);
        if (lane == 0)
            reinterpret_cast<unsigned char*>(epi.tile_flags)[((size_t)(row0 / BM) * epi.flag_ld + (col0 >> 7)) * 4 + ((row0 >> 5) & 3)] =
                (unsigned char)any_nz;
    }
}

// EPI_RESID for a warp whose 128 columns lie inside one half of the row ([0, m) or [m, 2m)) and inside the matrix, m even:
// no per-element bounds, one base pointer per array and row with compile-time offsets.  Same arithmetic, same bits as
// epilogue_tile<EPI_RESID>; the general function keeps the straddling and ragged runs.  (ncu: the general epilogue spends
// more than half of its ~60 instructions per element on predicates, branches and 64-bit address arithmetic.)
__device__ __forceinline__ void epilogue_resid_fast(const GemmEpilogue& epi, const float (&acc)[EPI_COLS], int lane, int row0,
                                                    int col0) {
    const int quad = lane & 3;
    const int cbase = col0 + 2 * quad;                                  // first column pair of this thread
    const int chan0 = cbase >= epi.m ? cbase - epi.m : cbase;           // ... as a channel index
    const int c8 = cbase >= epi.m ? cbase - epi.m + (epi.ld8 >> 1) : cbase;  // ... in the half-row fp8 layout
#pragma unroll
    for (int lh = 0; lh < 2; ++lh) {
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
            const int row = row0 + (lane >> 2) + 8 * rh + 16 * lh;
            const bool row_ok = row < epi.P;
            float dsum = 0.f, rmax = 0.f;
            if (row_ok) {
                const float descale = kTwiddleDescale / __ldg(epi.a_scale + row);
                const float oscale = __ldg(epi.out_scale + row);
                const size_t row_off = (size_t)row * epi.ld_out;
                const float2* bp = reinterpret_cast<const float2*>(epi.b + row_off + cbase);
                const float2* cp = epi.chan_scale
                                       ? reinterpret_cast<const float2*>(epi.chan_scale + (epi.chan_per_pixel ? (size_t)row * epi.m : 0) + chan0)
                                       : nullptr;
                __half2* hp = reinterpret_cast<__half2*>(epi.out_hi + row_off + cbase);
                __half2* lp = reinterpret_cast<__half2*>(epi.out_lo + row_off + cbase);
                __nv_fp8x2_storage_t* p8 = epi.out8 ? reinterpret_cast<__nv_fp8x2_storage_t*>(epi.out8 + (size_t)row * epi.ld8 + c8) : nullptr;
                float2 in0[16], in1[16];
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {  // the row's 32 loads in flight before any result is stored
                    in0[jj] = __ldcg(bp + 4 * jj);
                    in1[jj] = cp ? __ldg(cp + 4 * jj) : make_float2(1.f, 1.f);
                }
#pragma unroll
                for (int jj = 0; jj < 16; ++jj) {
                    const int ai = (((lh * 2 + (jj >> 3)) * 8 + (jj & 7)) * 2 + rh) * 2;
                    const float a0 = acc[ai] * descale, a1 = acc[ai + 1] * descale;
                    const float r0 = __fmaf_rn(in1[jj].x, a0, -in0[jj].x) * oscale, r1 = __fmaf_rn(in1[jj].y, a1, -in0[jj].y) * oscale;  // (explicit: same bits in both residual epilogues)
                    const __half2 h = __floats2half2_rn(r0, r1);
                    const float2 hf = __half22float2(h);
                    const __half2 l = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
                    hp[4 * jj] = h;
                    lp[4 * jj] = l;
                    dsum += fabsf(hf.x) + fabsf(hf.y);
                    rmax = fmaxf(rmax, fmaxf(fabsf(r0), fabsf(r1)));
                    if (p8 != nullptr) {  // kernel-uniform
                        const float2 v8 = make_float2(hf.x * kOperandScale8, hf.y * kOperandScale8);
                        if (fmaxf(fabsf(v8.x), fabsf(v8.y)) > 448.f) dsum = __int_as_float(0x7f800000);  // unscreenable row
                        p8[4 * jj] = __nv_cvt_float2_to_fp8x2(v8, __NV_SATFINITE, __NV_E4M3);
                    }
                }
            }
            if (epi.abs_sum_out != nullptr) {  // kernel-uniform
                dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
                dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
                if (quad == 0 && row_ok && dsum != 0.f) atomicAdd(epi.abs_sum_out + row, dsum);
            }
            if (epi.row_stat != nullptr) {  // kernel-uniform
                rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 1));
                rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 2));
                if (quad == 0 && row_ok && rmax != 0.f)
                    atomicMax(reinterpret_cast<unsigned int*>(epi.row_stat + 4 * (size_t)row + 3), __float_as_uint(rmax));
            }
        }
    }
}

// The same epilogue as a two-deep software pipeline (dft_gemm_kernel<EPI_RESID>; ncu source page of the one-shot version:
// two thirds of the epilogue warps' samples sit on the first use of the b values of each of the four row groups, 8 warps per
// SM cannot hide four serial L2 / HBM round trips per tile).  A row group g = 2 lh + rh is the thread's 16 column pairs of
// one row; resid_request() only issues its 16 loads, resid_compute() consumes them.  The kernel requests group 0
// BEFORE it waits for the tile's accumulator chunks and group g + 1 before group g is computed.  The channel
// factors (shared by all rows unless chan_per_pixel) come from a shared-memory copy made once per CTA.  RAGGED: the warp's 128
// columns may cross m or end at N (m even): per-pair bounds and channel index.  Same arithmetic, same bits as
// epilogue_tile<EPI_RESID>.
template <bool RAGGED>
__device__ __forceinline__ void resid_request(const GemmEpilogue& epi, int lane, int row0, int col0, int g, float2 (&in0)[16]) {
    const int row = row0 + (lane >> 2) + 8 * (g & 1) + 16 * (g >> 1);
    const int cbase = col0 + 2 * (lane & 3);
    const float2* bp = reinterpret_cast<const float2*>(epi.b + (size_t)row * epi.ld_out + cbase);
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
        const bool ok = row < epi.P && (!RAGGED || cbase + 8 * jj < epi.N);
        in0[jj] = ok ? __ldcg(bp + 4 * jj) : make_float2(0.f, 0.f);
    }
}

template <bool RAGGED, bool CS_SMEM>
__device__ __forceinline__ void resid_compute(const GemmEpilogue& epi, const float (&acc)[EPI_COLS], const float2 (&in0)[16], int lane,
                                              int row0, int col0, int g, const float* cs_smem) {
    const int quad = lane & 3;
    const int lh = g >> 1, rh = g & 1;
    const int cbase = col0 + 2 * quad;
    const int h8 = epi.ld8 >> 1;
    const int row = row0 + (lane >> 2) + 8 * rh + 16 * lh;
    const bool row_ok = row < epi.P;
    float dsum = 0.f, rmax = 0.f;
    if (row_ok) {
        const float descale = kTwiddleDescale / __ldg(epi.a_scale + row);
        const float oscale = __ldg(epi.out_scale + row);
        const size_t row_off = (size_t)row * epi.ld_out;
        const float* cs_row = CS_SMEM ? cs_smem : (epi.chan_scale ? epi.chan_scale + (epi.chan_per_pixel ? (size_t)row * epi.m : 0) : nullptr);
        __half2* hp = reinterpret_cast<__half2*>(epi.out_hi + row_off + cbase);
        __half2* lp = reinterpret_cast<__half2*>(epi.out_lo + row_off + cbase);
        unsigned char* o8 = epi.out8 ? epi.out8 + (size_t)row * epi.ld8 : nullptr;
        const int chan_base = (!RAGGED && cbase >= epi.m) ? cbase - epi.m : cbase;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int col = cbase + 8 * jj;
            if (RAGGED && col >= epi.N) continue;
            const int chan = RAGGED ? (col >= epi.m ? col - epi.m : col) : chan_base + 8 * jj;
            float2 cs = make_float2(1.f, 1.f);
            if (CS_SMEM) cs = *reinterpret_cast<const float2*>(cs_smem + chan);
            else if (cs_row != nullptr) cs = __ldg(reinterpret_cast<const float2*>(cs_row + chan));
            const int ai = (((lh * 2 + (jj >> 3)) * 8 + (jj & 7)) * 2 + rh) * 2;
            const float a0 = acc[ai] * descale, a1 = acc[ai + 1] * descale;
            const float r0 = __fmaf_rn(cs.x, a0, -in0[jj].x) * oscale, r1 = __fmaf_rn(cs.y, a1, -in0[jj].y) * oscale;
            const __half2 h = __floats2half2_rn(r0, r1);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
            hp[4 * jj] = h;
            lp[4 * jj] = l;
            dsum += fabsf(hf.x) + fabsf(hf.y);
            rmax = fmaxf(rmax, fmaxf(fabsf(r0), fabsf(r1)));
            if (o8 != nullptr) {  // kernel-uniform
                const float2 v8 = make_float2(hf.x * kOperandScale8, hf.y * kOperandScale8);
                if (fmaxf(fabsf(v8.x), fabsf(v8.y)) > 448.f) dsum = __int_as_float(0x7f800000);  // unscreenable row
                // (fp8 rows are [first half | pad | second half | pad], halves ld8/2 apart: plan.cu)
                *reinterpret_cast<__nv_fp8x2_storage_t*>(o8 + (col < epi.m ? col : col - epi.m + h8)) =
                    __nv_cvt_float2_to_fp8x2(v8, __NV_SATFINITE, __NV_E4M3);
            }
        }
    }
    if (epi.abs_sum_out != nullptr) {  // kernel-uniform
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
        dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
        if (quad == 0 && row_ok && dsum != 0.f) atomicAdd(epi.abs_sum_out + row, dsum);
    }
    if (epi.row_stat != nullptr) {  // kernel-uniform
        rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 1));
        rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 2));
        if (quad == 0 && row_ok && rmax != 0.f)
            atomicMax(reinterpret_cast<unsigned int*>(epi.row_stat + 4 * (size_t)row + 3), __float_as_uint(rmax));
    }
}

// group 0 is in flight in ba when this is called (after the last accumulator chunk has been drained)
template <bool RAGGED, bool CS_SMEM>
__device__ __forceinline__ void resid_pipeline(const GemmEpilogue& epi, const float (&acc)[EPI_COLS], float2 (&ba)[16], float2 (&bb)[16],
                                               int lane, int row0, int col0, const float* cs_smem) {
    resid_request<RAGGED>(epi, lane, row0, col0, 1, bb);
    resid_compute<RAGGED, CS_SMEM>(epi, acc, ba, lane, row0, col0, 0, cs_smem);
    resid_request<RAGGED>(epi, lane, row0, col0, 2, ba);
    resid_compute<RAGGED, CS_SMEM>(epi, acc, bb, lane, row0, col0, 1, cs_smem);
    resid_request<RAGGED>(epi, lane, row0, col0, 3, bb);
    resid_compute<RAGGED, CS_SMEM>(epi, acc, ba, lane, row0, col0, 2, cs_smem);
    resid_compute<RAGGED, CS_SMEM>(epi, acc, bb, lane, row0, col0, 3, cs_smem);
}

// Sparse tiles (GemmEpilogue::group_sparse): the tile's few active K blocks were accumulated as ONE chunk, so there is no
// register accumulator to keep: the thread requests ALL its b values before it waits for the tensor core, then reads the TMEM
// accumulator in four 16-row x 64-column pieces and finishes each from registers (one chunk added to a zero accumulator is the
// chunk itself: same values as the chunked path would give for this grouping).  The caller passes the wait for the
// accumulator and the release of the TMEM buffer (called after the last TMEM read).
// Global accesses are ROW-CONTIGUOUS.  In the accumulator fragment a warp-wide 8-byte access touches eight
// rows (eight 32-byte sectors in eight cache lines): 64 loads and 192 stores of that shape per thread and tile keep the
// load/store unit, not the memory, busy (ncu: the epilogue warps stall on the LSU queue while b arrives at 0.5 TB/s).  Here
// every 8-row x 64-column sub-piece of the accumulator is transposed through a 2.3 KB shared-memory buffer of the warp so that
// a lane owns 4 consecutive columns of a row: b arrives as 16-byte loads (two rows, four cache lines per warp access, all
// requested before the wait for the tensor core), r_hi / r_lo leave as 8-byte and the fp8 copy as 4-byte stores -- a sixth of
// the cache-line touches and half of the memory instructions.  Needs m % 4 == 0.
// Element arithmetic as in resid_compute (same bits); the row sums feed bounds only.
template <bool CS_PP, typename Wait, typename Release>
__device__ __forceinline__ void resid_direct_rows(const GemmEpilogue& epi, uint32_t taddr, int lane, int row0, int col0, const float* cs_smem,
                                                  float* xp, Wait wait_acc, Release release) {
    constexpr bool RAGGED = true;  // (per-group column bound: m % 4 == 0, so a group of 4 columns is inside or outside)
    const bool CS_SMEM = cs_smem != nullptr;
    const int lrow = lane >> 4;          // which of the two rows of an access
    const int lcol = 4 * (lane & 15);    // first of the lane's 4 columns inside a 64-column piece
    float2* rowbuf = reinterpret_cast<float2*>(xp + 8 * XP_ROW_FLOATS);
    // (descale, out_scale) of the warp's 32 rows: requested first, stored after the b requests are out
    float rs_a = 1.f, rs_o = 0.f;
    if (row0 + lane < epi.P) {
        rs_a = __ldg(epi.a_scale + row0 + lane);
        rs_o = __ldg(epi.out_scale + row0 + lane);
    }
    // every b value of the thread: [lh][ch][rh][i] -> row 16 lh + 8 rh + 2 i + lrow, columns 64 ch + lcol .. + 3
    float4 bq[32];
#pragma unroll
    for (int lh = 0; lh < 2; ++lh)
#pragma unroll
        for (int ch = 0; ch < 2; ++ch)
#pragma unroll
            for (int rh = 0; rh < 2; ++rh)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = row0 + 16 * lh + 8 * rh + 2 * i + lrow, col = col0 + 64 * ch + lcol;
                    const bool ok = row < epi.P && (!RAGGED || col < epi.N);
                    bq[((lh * 2 + ch) * 2 + rh) * 4 + i] =
                        ok ? __ldcg(reinterpret_cast<const float4*>(epi.b + (size_t)row * epi.ld_out + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
    __syncwarp();
    rowbuf[lane] = make_float2(row0 + lane < epi.P ? kTwiddleDescale / rs_a : 0.f, rs_o);
    float4 csv[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
        const int col = col0 + 64 * ch + lcol;
        const int chan = col >= epi.m ? col - epi.m : col;
        csv[ch] = make_float4(1.f, 1.f, 1.f, 1.f);
        if (CS_SMEM && (!RAGGED || col < epi.N)) csv[ch] = *reinterpret_cast<const float4*>(cs_smem + chan);
    }
    const int h8 = epi.ld8 >> 1;
    // channel factors of the pixel's own (chan_per_pixel): requested one sub-piece ahead, 4 x 16 bytes per lane
    constexpr bool cs_pp = CS_PP;  // (channel factors from global memory: the pixel's own, or more than CS_SMEM_FLOATS shared ones)
    float4 csn[4];
    auto request_cs = [&](int lh, int ch, int rh) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = row0 + 16 * lh + 8 * rh + 2 * i + lrow, col = col0 + 64 * ch + lcol;
            const int chan = col >= epi.m ? col - epi.m : col;
            csn[i] = (row < epi.P && col < epi.N)
                         ? __ldg(reinterpret_cast<const float4*>(epi.chan_scale + (epi.chan_per_pixel ? (size_t)row * epi.m : 0) + chan))
                         : make_float4(1.f, 1.f, 1.f, 1.f);
        }
    };
    if (cs_pp) request_cs(0, 0, 0);
    __syncwarp();
    wait_acc();
#pragma unroll
    for (int lh = 0; lh < 2; ++lh) {
        float dsum[8], rmax[8];  // per (rh, i): the lane's share of a row
#pragma unroll
        for (int k = 0; k < 8; ++k) dsum[k] = 0.f, rmax[k] = 0.f;
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
            float v[32];
            tmem_ld_16x256b_x8(taddr + ((uint32_t)(16 * lh) << 16) + (uint32_t)(64 * ch), v);
            if (lh == 1 && ch == 1) release();
            const int col = col0 + 64 * ch + lcol;
            const bool col_ok = !RAGGED || col < epi.N;
            const int col8 = col < epi.m ? col : col - epi.m + h8;
#pragma unroll
            for (int rh = 0; rh < 2; ++rh) {
                // fragment -> rows: thread T holds (row T/4, columns 8 j + 2 (T%4), + 1) of this sub-piece
                __syncwarp();
#pragma unroll
                for (int j8 = 0; j8 < 8; ++j8)
                    *reinterpret_cast<float2*>(xp + (lane >> 2) * XP_ROW_FLOATS + 8 * j8 + 2 * (lane & 3)) =
                        make_float2(v[(j8 * 2 + rh) * 2], v[(j8 * 2 + rh) * 2 + 1]);
                __syncwarp();
                float4 csc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) csc[i] = cs_pp ? csn[i] : csv[ch];
                if (cs_pp && !(lh == 1 && ch == 1 && rh == 1)) {  // next sub-piece in processing order (lh, ch, rh)
                    const int nx = (lh * 2 + ch) * 2 + rh + 1;
                    request_cs(nx >> 2, (nx >> 1) & 1, nx & 1);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int lr = 16 * lh + 8 * rh + 2 * i + lrow;  // row inside the warp's 32
                    const int row = row0 + lr;
                    const float4 a = *reinterpret_cast<const float4*>(xp + (2 * i + lrow) * XP_ROW_FLOATS + lcol);
                    if (row >= epi.P || !col_ok) continue;
                    const float2 sc = rowbuf[lr];  // (descale, out_scale)
                    const float4 b = bq[((lh * 2 + ch) * 2 + rh) * 4 + i];
                    const float4 cf = csc[i];
                    const float r0 = __fmaf_rn(cf.x, a.x * sc.x, -b.x) * sc.y, r1 = __fmaf_rn(cf.y, a.y * sc.x, -b.y) * sc.y;
                    const float r2 = __fmaf_rn(cf.z, a.z * sc.x, -b.z) * sc.y, r3 = __fmaf_rn(cf.w, a.w * sc.x, -b.w) * sc.y;
                    const __half2 h01 = __floats2half2_rn(r0, r1), h23 = __floats2half2_rn(r2, r3);
                    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                    const __half2 l01 = __floats2half2_rn(r0 - f01.x, r1 - f01.y), l23 = __floats2half2_rn(r2 - f23.x, r3 - f23.y);
                    const size_t off = (size_t)row * epi.ld_out + col;
                    uint2 hv, lv;
                    hv.x = *reinterpret_cast<const unsigned int*>(&h01);
                    hv.y = *reinterpret_cast<const unsigned int*>(&h23);
                    lv.x = *reinterpret_cast<const unsigned int*>(&l01);
                    lv.y = *reinterpret_cast<const unsigned int*>(&l23);
                    *reinterpret_cast<uint2*>(epi.out_hi + off) = hv;
                    *reinterpret_cast<uint2*>(epi.out_lo + off) = lv;
                    const int k = rh * 4 + i;
                    dsum[k] += (fabsf(f01.x) + fabsf(f01.y)) + (fabsf(f23.x) + fabsf(f23.y));
                    rmax[k] = fmaxf(rmax[k], fmaxf(fmaxf(fabsf(r0), fabsf(r1)), fmaxf(fabsf(r2), fabsf(r3))));
                    if (epi.out8 != nullptr) {  // kernel-uniform
                        const float2 v01 = make_float2(f01.x * kOperandScale8, f01.y * kOperandScale8);
                        const float2 v23 = make_float2(f23.x * kOperandScale8, f23.y * kOperandScale8);
                        if (fmaxf(fmaxf(fabsf(v01.x), fabsf(v01.y)), fmaxf(fabsf(v23.x), fabsf(v23.y))) > 448.f)
                            dsum[k] = __int_as_float(0x7f800000);  // unscreenable row
                        const unsigned int p01 = __nv_cvt_float2_to_fp8x2(v01, __NV_SATFINITE, __NV_E4M3);
                        const unsigned int p23 = __nv_cvt_float2_to_fp8x2(v23, __NV_SATFINITE, __NV_E4M3);
                        *reinterpret_cast<unsigned int*>(epi.out8 + (size_t)row * epi.ld8 + col8) = p01 | (p23 << 16);
                    }
                }
            }
        }
        // row totals: the 16 lanes that share a row
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int row = row0 + 16 * lh + 8 * (k >> 2) + 2 * (k & 3) + lrow;
            float d = dsum[k], r = rmax[k];
#pragma unroll
            for (int sh = 1; sh < 16; sh <<= 1) {
                d += __shfl_xor_sync(0xffffffffu, d, sh);
                r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, sh));
            }
            if ((lane & 15) == 0 && row < epi.P) {
                if (epi.abs_sum_out != nullptr && d != 0.f) atomicAdd(epi.abs_sum_out + row, d);
                if (epi.row_stat != nullptr && r != 0.f)
                    atomicMax(reinterpret_cast<unsigned int*>(epi.row_stat + 4 * (size_t)row + 3), __float_as_uint(r));
            }
        }
    }
}

// ---- screening (exact) ----------------------------------------------------------------------------------------
// A pixel tile x 256-column tile whose x and z are identically zero stays zero through the FISTA update iff
// |step * g| <= step * lambda for every element.  Far from the sources g is a few noise sigmas and lambda is tens of
// them, so that decision does not need fp32-class accuracy: the screening kernel computes g1 = A_hi * B_hi only (one
// third of the tensor work, half of the operand traffic) and bounds what the other two products and the rounding of
// both evaluations could add:  |g3 - g1| <= (|A_lo| |B_hi| + |A_hi| |B_lo| + rounding) summed over K
// <= (2 + 2 + K 2^-13) S_p  in accumulator units, with S_p = sum_k |A_hi[p, k]| (|B_hi| <= 2^12, |B_lo| <= 2,
// |A_lo| <= 2^-11 |A_hi|).  A tile passes when |g1| + kScreenSlack(K) S_p < lambda for all its elements: the full
// kernel would have produced exact zeros there, so nothing is written.  Any other tile is appended to a list the
// three-product FISTA kernel then processes.  Results are bit-identical to running the full kernel on every tile.
//
// A first, cheaper pass does the same with fp8 (e4m3) copies of both operands (A8 = rn(2 A_hi), B8 = rn(B_hi / 16),
// accumulator units 1/8 of the fp16 ones): half the operand bytes and twice the tensor rate.  Its bound, in fp8
// accumulator units:  sum_k |2 A_hi - A8| |B_hi / 16| + |A8| |B_hi / 16 - B8|
//   <= sum_k (2^-4 2 |A_hi| + 2^-10) 256 + (2.125 |A_hi| + 2^-10) (16 + 2^-10)  <=  66.1 S_p + 0.27 K,
// plus the fp16-level terms above (S_p in these units): slack 68 S_p + 0.3 K.  (A row whose 2 A_hi would exceed the
// e4m3 range poisons its S_p with +inf and is never screened.)  Tiles that fail the fp8 test go to the fp16 test, tiles
// that fail that one to the three-product kernel.
__device__ __forceinline__ float screen_slack(int K) { return 8.f + (float)K * (1.f / 2048.f); }

template <bool F8>
__device__ __forceinline__ bool screen_tile_fails(const GemmEpilogue& epi, const float (&acc)[EPI_COLS], int lane, int row0,
                                                  int col0, int K) {
    const int quad = lane & 3;
    bool fail = false;
#pragma unroll
    for (int lh = 0; lh < 2; ++lh) {
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
            const int row = row0 + (lane >> 2) + 8 * rh + 16 * lh;
            if (row >= epi.P) continue;
            const float l0 = __ldg(epi.lambda0 + row), nz = __ldg(epi.noise + row);
            const float lam = l0 - (float)epi.iter * nz;
            const bool active = (nz < l0) && (epi.iter == 0 || lam > 0.f) && (!epi.iter_cap || epi.iter < __ldg(epi.iter_cap + row));
            if (!active) continue;  // a frozen row keeps its (zero) state whatever g is
            // accumulator units: g = acc * kTwiddleDescale / a_scale;  pass iff |acc| + slack * S_p < lambda / that factor
            const float sp = __ldg(epi.abs_sum + row);
            const float lim = F8 ? lam * __ldg(epi.a_scale + row) * (kTwiddleScale8 * kOperandScale8) - (68.f * sp + 0.3f * (float)K)
                                 : lam * __ldg(epi.a_scale + row) * (1.f / kTwiddleDescale) - screen_slack(K) * sp;
            float mx = 0.f;
            bool bad = !(lim > 0.f);
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int col = col0 + 8 * jj + 2 * quad;
                if (col >= epi.N) continue;
                const int ai = (((lh * 2 + (jj >> 3)) * 8 + (jj & 7)) * 2 + rh) * 2;
                mx = fmaxf(mx, fmaxf(fabsf(acc[ai]), fabsf(acc[ai + 1])));
                bad |= (acc[ai] != acc[ai]) | (acc[ai + 1] != acc[ai + 1]);
            }
            fail |= bad | !(mx < lim);
        }
    }
    return fail;
}

__device__ __forceinline__ bool options_rows_ok(const GemmEpilogue& epi) {
    return epi.no_rows == 0 && (reinterpret_cast<uintptr_t>(epi.b) & 15) == 0 &&
           ((reinterpret_cast<uintptr_t>(epi.out_hi) | reinterpret_cast<uintptr_t>(epi.out_lo)) & 7) == 0 &&
           (epi.out8 == nullptr || ((reinterpret_cast<uintptr_t>(epi.out8) | (uintptr_t)epi.ld8 | (uintptr_t)(epi.ld8 >> 1)) & 3) == 0);
}

// GemmEpilogue::single_rt (per call, in the caller's workspace): [0 .. MT) one byte per pixel tile, 1 = its forward tiles go to
// dft_fwd_direct_kernel (written by tile_single_kernel); then, 4-byte aligned, one int: the number of pixel tiles NOT marked
// (zeroed before tile_single_kernel): none -> dft_gemm_kernel<EPI_RESID> returns at once
__host__ __device__ __forceinline__ size_t single_rt_counter_offset(int MT) { return ((size_t)MT + 3) & ~(size_t)3; }

// Measurement aid (CSR_TIMELINE=1, csr_debug_timeline): clock64 stamps of the three roles for the first kTimelineTiles tiles of
// every CTA of the last residual forward launch -- [cta][tile][8]: producer start / last load issued, MMA first operand seen /
// last commit issued, epilogue tile start / last chunk drained / body done, K blocks of the tile.
constexpr int kTimelineTiles = 16;
__device__ long long g_timeline[160 * kTimelineTiles * 8];
#define TL_STAMP(slot)                                                                                          \
    if (((MODE == EPI_RESID && epi.timeline == 1) || (MODE == EPI_FISTA && epi.timeline == 2)) && tl_n < kTimelineTiles) \
        g_timeline[((size_t)blockIdx.x * kTimelineTiles + tl_n) * 8 + (slot)] = clock64();

template <int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1)
dft_gemm_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                const __grid_constant__ CUtensorMap map_pf0, const __grid_constant__ CUtensorMap map_pf1,
                int num_prefetch, int prefetch_lead, int K, int MT, int NT, int chunk_kb, const GemmEpilogue epi) {
    if (MODE == EPI_RESID && epi.skip_single != 0 &&
        *reinterpret_cast<const int*>(epi.single_rt + single_rt_counter_offset(MT)) == 0)
        return;   // (kernel-uniform: every tile is the direct kernel's)
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;                         // [NSTAGES]  TMA -> MMA
    uint64_t* empty = bars + MAX_STAGES;           // [NSTAGES]  MMA -> TMA
    uint64_t* tmem_full = bars + 2 * MAX_STAGES;   // [2]       MMA -> epilogue (one K chunk accumulated)
    uint64_t* tmem_empty = bars + 2 * MAX_STAGES + 2;  // [2]   epilogue -> MMA (chunk drained to registers)
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // tiles: all MT * NT of them, or the list a screening launch left behind
    const int total_tiles = epi.tile_list ? __ldcg(epi.tile_count) : MT * NT;
    constexpr bool IS_SCREEN = (MODE == EPI_SCREEN || MODE == EPI_SCREEN8);
    constexpr bool F8 = (MODE == EPI_SCREEN8);
    constexpr int BKE = F8 ? 2 * BK : BK;  // K elements per block = one 128-byte swizzle row
    const int KB = (K + BKE - 1) / BKE;
    constexpr bool ONE = IS_SCREEN || MODE == EPI_PLAIN1;  // one product per K block
    constexpr int PRODUCTS = ONE ? 1 : 3;
    // a one-product stage is half the size, so the same shared memory holds a ring twice as deep -- which is what
    // keeps enough bytes in flight to hide the L2 -> SM latency when a K block is only 4 MMAs long
    constexpr int NSTAGES = ONE ? MAX_STAGES : STAGES;
    constexpr int STAGE_STRIDE = ONE ? A_TILE_BYTES + B_TILE_BYTES : STAGE_BYTES;
    constexpr int B_OFFSET = ONE ? A_TILE_BYTES : 2 * A_TILE_BYTES;
    int* last_listed = reinterpret_cast<int*>(smem + STAGES * STAGE_BYTES + 256 + NUM_WARPS_TOTAL * KMASK_WORDS * 4);
    uint32_t* my_mask = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + 256) + warp * KMASK_WORDS;
    const bool have_flags = epi.a_kflags != nullptr;
    const bool skip = have_flags && epi.no_kskip == 0;
    // sparse grouping (EPI_RESID): a tile with at most group_sparse active K blocks accumulates them as ONE chunk -- decided
    // from the flags alone, so a run that walks every K block (no_kskip) groups, and rounds, exactly like one that skips
    const int group_max = (MODE == EPI_RESID && have_flags) ? epi.group_sparse : 0;
    const int kwords = (KB + 1) >> 1;  // one flag word per 128 K columns = 2 K blocks

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
        if (num_prefetch > 0) tma_prefetch_desc(&map_pf0);
        if (num_prefetch > 1) tma_prefetch_desc(&map_pf1);
        for (int s = 0; s < NSTAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], EPI_WARPS);
        }
        *last_listed = -1;
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, TMEM_COLS);
        tmem_relinquish();
    }
    // residual epilogue: the channel factors shared by all lines of sight, once per CTA
    float* cs_smem = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + SMEM_MISC_BYTES);
    const bool cs_in_smem = MODE == EPI_RESID && epi.chan_scale != nullptr && epi.chan_per_pixel == 0 && epi.m <= CS_SMEM_FLOATS;
    if (cs_in_smem)
        for (int i = threadIdx.x; i < epi.m; i += NUM_THREADS) cs_smem[i] = __ldg(epi.chan_scale + i);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    auto tile_id = [&](int ti) -> int { return epi.tile_list ? __ldcg(epi.tile_list + ti) : ti; };
    // EPI_SCREEN: is the state (x, z) of this tile identically zero?  (both 128-column blocks of all 128 rows)
    auto tile_quiet = [&](int t) -> bool {
        int mt, nt;
        tile_coords(t, MT, NT, mt, nt);
        const unsigned int* st = epi.tile_state + (size_t)mt * epi.flag_ld + 2 * nt;
        unsigned int wd = __ldcg(st);
        if (2 * nt + 1 < epi.flag_ld) wd |= __ldcg(st + 1);
        return wd == 0u;
    };

    // register budget: the two single-thread roles need few registers, the epilogue warps hold a 128-register
    // accumulator fragment each -- hand the registers over (64512 = 128*40 + 256*232)
    if (warp < 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
      if (warp == 0) {
        // ===================== TMA producer =====================
        {
            int stage = 0;
            uint32_t phase = 0;
            bool quiet_next = (IS_SCREEN && (int)blockIdx.x < total_tiles) ? tile_quiet(tile_id(blockIdx.x)) : true;
            for (int ti = blockIdx.x; ti < total_tiles; ti += gridDim.x) {
                const int t = tile_id(ti);
                int mt, nt;
                tile_coords(t, MT, NT, mt, nt);
                if (MODE == EPI_RESID && epi.skip_single != 0 && epi.single_rt[mt]) continue;   // dft_fwd_direct_kernel's
                const bool quiet = quiet_next;  // (the next tile's state is fetched while this one streams)
                if (IS_SCREEN && ti + (int)gridDim.x < total_tiles) quiet_next = tile_quiet(tile_id(ti + gridDim.x));
                if (IS_SCREEN && !quiet) continue;  // tile with signal: left to the three-product kernel
                if (skip) build_kmask(epi.a_kflags + (size_t)mt * epi.a_kflag_ld, kwords, KB, my_mask, lane);
                const int tl_n = (ti - (int)blockIdx.x) / (int)gridDim.x;
                if (lane == 0) {
                    TL_STAMP(0);
                    // Pull the fp32 tiles the fused epilogue of THIS tile will read (z and x_old, or b) into L2 shortly
                    // before the K loop ends, so the epilogue's loads see L2 latency instead of HBM latency.  Not
                    // earlier: more than an L2's worth of data streams through per tile period and would evict them.
                    // (a K-block-skipping tile is a few blocks long and the producer runs a tile ahead: prefetch at once)
                    const int pf_kb = skip ? 0 : max(0, KB - prefetch_lead);
                    bool pf_due = true;
                    for (int kb = skip ? next_active_kb(my_mask, 0, KB) : 0; kb < KB;
                         kb = skip ? next_active_kb(my_mask, kb + 1, KB) : kb + 1) {
                        if (pf_due && kb >= pf_kb) {
                            if (num_prefetch > 0) tma_prefetch_2d(&map_pf0, nt * BN, mt * BM);
                            if (num_prefetch > 1) tma_prefetch_2d(&map_pf1, nt * BN, mt * BM);
                            pf_due = false;
                        }
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* sb = stage_base + stage * STAGE_STRIDE;
                        mbar_arrive_expect_tx(&full[stage], STAGE_STRIDE);
                        tma_load_2d(sb, &map_a_hi, &full[stage], kb * BKE, mt * BM);
                        if (PRODUCTS > 1) tma_load_2d(sb + A_TILE_BYTES, &map_a_lo, &full[stage], kb * BKE, mt * BM);
                        tma_load_2d(sb + B_OFFSET, &map_b_hi, &full[stage], kb * BKE, nt * BN);
                        if (PRODUCTS > 1) tma_load_2d(sb + B_OFFSET + B_TILE_BYTES, &map_b_lo, &full[stage], kb * BKE, nt * BN);
                        if (++stage == NSTAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                    if (pf_due && skip) {  // (no active K block at all: the epilogue still reads b)
                        if (num_prefetch > 0) tma_prefetch_2d(&map_pf0, nt * BN, mt * BM);
                        if (num_prefetch > 1) tma_prefetch_2d(&map_pf1, nt * BN, mt * BM);
                    }
                    TL_STAMP(1);
                }
                __syncwarp();  // the mask slot is rewritten for the next tile
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The tensor core truncates when it adds into the fp32 accumulator, so the error of a long K loop grows
        // linearly with K.  The accumulation is therefore cut into chunks of `chunk_kb` K blocks: each chunk
        // starts a fresh TMEM accumulator (two of them, ping-pong) which the epilogue warps add into registers
        // with round-to-nearest while the next chunk is being computed.
        {
            constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t chunk = 0;  // running chunk counter -> TMEM buffer = chunk & 1, parity = (chunk >> 1) & 1
            unsigned long long issued = 0;  // K blocks x products (executed-work accounting)
            bool quiet_next = (IS_SCREEN && (int)blockIdx.x < total_tiles) ? tile_quiet(tile_id(blockIdx.x)) : true;
            for (int ti = blockIdx.x; ti < total_tiles; ti += gridDim.x) {
                const bool quiet = quiet_next;
                if (IS_SCREEN && ti + (int)gridDim.x < total_tiles) quiet_next = tile_quiet(tile_id(ti + gridDim.x));
                if (IS_SCREEN && !quiet) continue;
                if (MODE == EPI_RESID && epi.skip_single != 0) {
                    int mt, nt;
                    tile_coords(tile_id(ti), MT, NT, mt, nt);
                    if (epi.single_rt[mt]) continue;
                }
                bool single = false;
                if (skip || group_max > 0) {
                    int mt, nt;
                    tile_coords(tile_id(ti), MT, NT, mt, nt);
                    const int nact = build_kmask(epi.a_kflags + (size_t)mt * epi.a_kflag_ld, kwords, KB, my_mask, lane);
                    single = group_max > 0 && nact <= group_max;
                }
                const int tl_n = (ti - (int)blockIdx.x) / (int)gridDim.x;
                if (lane == 0) {
                    // A chunk is the set of (active) K blocks inside one aligned window of `chunk_kb` blocks, so skipping
                    // zero blocks never changes how the others are grouped: bit-identical to the dense loop.
                    int in_chunk = 0, cur_win = -1;  // K blocks accumulated into the current chunk so far / its window
                    uint32_t d_tmem = 0;
                    int tl_kb = 0;
                    for (int kb = skip ? next_active_kb(my_mask, 0, KB) : 0; kb < KB;
                         kb = skip ? next_active_kb(my_mask, kb + 1, KB) : kb + 1) {
                        const int win = single ? 0 : kb / chunk_kb;
                        if (in_chunk != 0 && win != cur_win) {
                            umma_commit(&tmem_full[chunk & 1]);  // chunk accumulator complete
                            ++chunk;
                            in_chunk = 0;
                        }
                        if (in_chunk == 0) {
                            const int buf = chunk & 1;
                            mbar_wait(&tmem_empty[buf], ((chunk >> 1) & 1) ^ 1);
                            tc_fence_after();
                            d_tmem = tmem_base + (uint32_t)(buf * BN);
                            cur_win = win;
                        }
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        if (tl_kb++ == 0) { TL_STAMP(2); }
                        const uint32_t sa_hi = smem_u32(stage_base + stage * STAGE_STRIDE);
                        const uint32_t sa_lo = sa_hi + A_TILE_BYTES;
                        const uint32_t sb_hi = sa_hi + B_OFFSET;
                        const uint32_t sb_lo = sb_hi + B_TILE_BYTES;
#pragma unroll
                        for (int k4 = 0; k4 < BK / 16; ++k4) {
                            const uint64_t da_hi = umma_desc_sw128(sa_hi + k4 * 32);
                            const uint64_t da_lo = umma_desc_sw128(sa_lo + k4 * 32);
                            const uint64_t db_hi = umma_desc_sw128(sb_hi + k4 * 32);
                            const uint64_t db_lo = umma_desc_sw128(sb_lo + k4 * 32);
                            if (F8) {
                                umma_f8(d_tmem, da_hi, db_hi, idesc, (in_chunk != 0 || k4 != 0) ? 1u : 0u);
                            } else {
                                umma_f16(d_tmem, da_hi, db_hi, idesc, (in_chunk != 0 || k4 != 0) ? 1u : 0u);
                            }
                            if (PRODUCTS > 1) {
                                umma_f16(d_tmem, da_lo, db_hi, idesc, 1u);
                                umma_f16(d_tmem, da_hi, db_lo, idesc, 1u);
                            }
                        }
                        issued += F8 ? 2 : PRODUCTS;  // units of one 128 x 256 x 64 block
                        umma_commit(&empty[stage]);  // smem stage reusable once these MMAs have read it
                        if (++stage == NSTAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                        ++in_chunk;
                    }
                    if (in_chunk != 0) {
                        umma_commit(&tmem_full[chunk & 1]);
                        ++chunk;
                    }
                    TL_STAMP(3);
                    if (((MODE == EPI_RESID && epi.timeline == 1) || (MODE == EPI_FISTA && epi.timeline == 2)) && tl_n < kTimelineTiles)
                        g_timeline[((size_t)blockIdx.x * kTimelineTiles + tl_n) * 8 + 7] = tl_kb;
                }
                chunk = __shfl_sync(0xffffffffu, chunk, 0);
            }
            // one counter per profile kind (csr_profile_collect's numbering), set by the launcher
            if (lane == 0 && epi.mma_count != nullptr && issued != 0) atomicAdd(epi.mma_count + epi.count_slot, issued);
        }
      }
    } else {
        // ===================== epilogue warps =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int ew = warp - 4;
        const int sub = warp & 3;        // TMEM lane quarter this warp may read (hardware rule: warp id mod 4)
        // the lean residual epilogue needs: m even, 8-byte aligned float2 / half2 views, no measurement switches
        const bool resid_fast = MODE == EPI_RESID && (epi.m & 1) == 0 && epi.debug == 0 && epi.no_fast == 0 && (epi.ld_out & 1) == 0 &&
                                ((reinterpret_cast<uintptr_t>(epi.b) | reinterpret_cast<uintptr_t>(epi.chan_scale)) & 7) == 0 &&
                                ((reinterpret_cast<uintptr_t>(epi.out_hi) | reinterpret_cast<uintptr_t>(epi.out_lo)) & 3) == 0 &&
                                (epi.out8 == nullptr || ((reinterpret_cast<uintptr_t>(epi.out8) | (uintptr_t)epi.ld8) & 1) == 0) &&
                                true;
        const int half = ew >> 2;        // which 128 of the tile's 256 columns
        // the direct epilogue with row-contiguous accesses: 16-byte views of b, 8-byte views of r_hi / r_lo, 4-byte views of the
        // fp8 copy, a 4-column group never straddles m, channel factors shared by all rows (or none)
        const bool rows_ok = MODE == EPI_RESID && resid_fast && (epi.m & 3) == 0 && (epi.ld_out & 3) == 0 && options_rows_ok(epi) &&
                             (cs_in_smem || epi.chan_scale == nullptr || (reinterpret_cast<uintptr_t>(epi.chan_scale) & 15) == 0);
        uint32_t chunk = 0;
        for (int ti = blockIdx.x; ti < total_tiles; ti += gridDim.x) {
            const int t = tile_id(ti);
            int mt, nt;
            tile_coords(t, MT, NT, mt, nt);
            if (MODE == EPI_RESID && epi.skip_single != 0 && epi.single_rt[mt]) continue;
            if (IS_SCREEN && !tile_quiet(t)) {  // hand the tile on to the next kernel of the cascade (one warp lists it)
                if (lane == 0 && atomicMax(last_listed, ti) < ti) epi.out_list[atomicAdd(epi.out_count, 1)] = t;
                continue;
            }
            float acc[EPI_COLS];
#pragma unroll
            for (int j = 0; j < EPI_COLS; ++j) acc[j] = 0.f;
            int nchunks = (KB + chunk_kb - 1) / chunk_kb;
            const int tl_n = (ti - (int)blockIdx.x) / (int)gridDim.x;
            if (ew == 0 && lane == 0) { TL_STAMP(4); }
            const int ecol0 = nt * BN + half * EPI_COLS, erow0 = mt * BM + sub * 32;
            const bool ragged = !(ecol0 + EPI_COLS <= epi.N && (ecol0 + EPI_COLS <= epi.m || ecol0 >= epi.m));  // (warp-uniform)
            bool single = false;
            if (skip || group_max > 0) {
                const int nact = build_kmask(epi.a_kflags + (size_t)mt * epi.a_kflag_ld, kwords, KB, my_mask, lane);
                single = group_max > 0 && nact <= group_max;
                if (single) {
                    nchunks = (skip && nact == 0) ? 0 : 1;
                } else if (skip) {  // count the windows that hold an active K block, exactly as the MMA thread walks them
                    int c = 0;
                    if (lane == 0) {
                        int cur_win = -1;
                        for (int kb = next_active_kb(my_mask, 0, KB); kb < KB; kb = next_active_kb(my_mask, kb + 1, KB)) {
                            const int win = kb / chunk_kb;
                            c += (win != cur_win);
                            cur_win = win;
                        }
                    }
                    nchunks = __shfl_sync(0xffffffffu, c, 0);
                }
            }
            if (MODE == EPI_RESID && rows_ok && single && nchunks == 1) {
                // sparse tile: straight from TMEM, every b value requested before the wait (resid_direct_rows)
                if (ecol0 < epi.N) {
                    const int buf = chunk & 1;
                    const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(buf * BN + half * EPI_COLS);
                    auto release = [&]() {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                    };
                    auto wait_acc = [&]() {
                        mbar_wait(&tmem_full[buf], (chunk >> 1) & 1);
                        tc_fence_after();
                        if (ew == 0 && lane == 0) { TL_STAMP(5); }
                    };
                    float* xp = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + SMEM_MISC_BYTES + CS_SMEM_FLOATS * 4 + ew * XP_WARP_BYTES);
                    if (!cs_in_smem && epi.chan_scale != nullptr)
                        resid_direct_rows<true>(epi, taddr, lane, erow0, ecol0, nullptr, xp, wait_acc, release);
                    else
                        resid_direct_rows<false>(epi, taddr, lane, erow0, ecol0, cs_in_smem ? cs_smem : nullptr, xp, wait_acc, release);
                } else {  // columns beyond the matrix: nothing to write, the accumulator is released as soon as it is complete
                    const int buf = chunk & 1;
                    mbar_wait(&tmem_full[buf], (chunk >> 1) & 1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                }
                ++chunk;
                if (ew == 0 && lane == 0) { TL_STAMP(6); }
                continue;
            }
            // residual epilogue, pipelined: the b values of the first row group travel while the K loop runs
            float2 bva[16], bvb[16];
            if (MODE == EPI_RESID && resid_fast) {
                if (ragged) resid_request<true>(epi, lane, erow0, ecol0, 0, bva);
                else resid_request<false>(epi, lane, erow0, ecol0, 0, bva);
            }
            for (int c = 0; c < nchunks; ++c, ++chunk) {
                const int buf = chunk & 1;
                mbar_wait(&tmem_full[buf], (chunk >> 1) & 1);
                tc_fence_after();
                drain_chunk(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(buf * BN + half * EPI_COLS), acc);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            }
            if (ew == 0 && lane == 0) { TL_STAMP(5); }
            if constexpr (IS_SCREEN) {
                const bool fail = screen_tile_fails<F8>(epi, acc, lane, mt * BM + sub * 32, nt * BN + half * EPI_COLS, K);
                if (__any_sync(0xffffffffu, fail)) {
                    if (lane == 0 && atomicMax(last_listed, ti) < ti) epi.out_list[atomicAdd(epi.out_count, 1)] = t;
                }
            } else if (MODE == EPI_RESID && resid_fast) {  // (kernel-uniform; the inner choices are warp-uniform)
                if (ecol0 < epi.N) {
                    if (cs_in_smem) {
                        if (ragged) resid_pipeline<true, true>(epi, acc, bva, bvb, lane, erow0, ecol0, cs_smem);
                        else resid_pipeline<false, true>(epi, acc, bva, bvb, lane, erow0, ecol0, cs_smem);
                    } else {
                        if (ragged) resid_pipeline<true, false>(epi, acc, bva, bvb, lane, erow0, ecol0, nullptr);
                        else resid_pipeline<false, false>(epi, acc, bva, bvb, lane, erow0, ecol0, nullptr);
                    }
                }
            } else {
                epilogue_tile<MODE == EPI_PLAIN1 ? EPI_PLAIN : MODE>(epi, acc, lane, erow0, ecol0);
            }
            if (ew == 0 && lane == 0) { TL_STAMP(6); }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}
#undef TL_STAMP


// ---- fp8 screening on CTA pairs (cta_group::2) ---------------------------------------------------------------------
// The one-CTA fp8 screening kernel is bound by operand delivery into the SM: 48 KB (A 16 + B 32) per 128 x 256 x 128
// block.  Here two CTAs of a cluster (the two SMs of a TPC) screen two pixel tiles of the same column tile with ONE
// M = 256 MMA: each CTA stages its own A tile and only HALF of the B tile (32 KB per block and SM), the tensor core
// reads both halves, and each CTA's TMEM receives its 128 rows.  Six 32 KB stages per CTA; the leader CTA (rank 0)
// issues the MMAs and owns the full / accumulator-empty barriers, MMA commits are multicast to both CTAs.
// Same test, same bound, same lists as dft_gemm_kernel<EPI_SCREEN8>; a pair runs when at least one of its two
// pixel tiles is quiet (the other CTA then just passes its tile on).
// ---- mirror-symmetric screening ------------------------------------------------------------------------------------
// phi_j = -phi_{n-j}: with b = alpha + i beta and theta = 2 (lambda2 - l2_ref) phi_j, j >= n/2,
//   Re F(+-phi_j) = C_alpha +- S_beta,   Im F(+-phi_j) = C_beta -+ S_alpha,   C_x = sum_i x_i cos theta_i, S_x = sum_i x_i sin theta_i
// so four real sums of length m bound the gradient at BOTH +phi_j and -phi_j: the GEMM [2P, m] (rows alpha_p, beta_p: the
// half rows of the residual operand) x [m, n] (the half rows cos / sin of the table rows j >= n/2) has half the K of the
// plain one for the same outputs.  |C| < L/2 and |S| < L/2 for every entry imply |C +- S| < L, with L = the plain test's
// limit (its slack covers both partial sums: S_p = S_alpha + S_beta, K = 2 m).  The one cell without a mirror (j = 0) is an
// extra column tile.  Original column ranges [lo, hi) of the real representation that symmetric column tile nt decides:
__device__ __forceinline__ int sym_ranges(int nt, int NT, int n, int (&lo)[4], int (&hi)[4]) {
    if (nt == NT - 1) {  // phi_0 alone
        lo[0] = 0;
        hi[0] = 1;
        lo[1] = n;
        hi[1] = n + 1;
        return 2;
    }
    const int j0 = n / 2 + (BN / 2) * nt, j1 = min(j0 + BN / 2, n);
    lo[0] = j0;  // Re F(+phi)
    hi[0] = j1;
    lo[1] = n - j1 + 1;  // Re F(-phi): j' = n - j
    hi[1] = n - j0 + 1;
    lo[2] = n + lo[0];   // Im
    hi[2] = n + hi[0];
    lo[3] = n + lo[1];
    hi[3] = n + hi[1];
    return 4;
}

// rows of the half-row view: row = 2 p + h; `ncols` valid columns from col0 (each phi is a (cos, sin) column pair).
// The thread that holds (C_alpha, S_alpha) of a (line of sight, phi) and the one that holds (C_beta, S_beta) are four lanes
// apart (consecutive rows of the fragment): one exchange gives each of them one of the two sums |C_x| + |S_y|.
// Returns which bounds failed, attributed to the ORIGINAL column tiles they belong to: bit 4 part + k, part 0 = Re F (the
// alpha rows test |C_alpha| + |S_beta|), 1 = Im F (beta rows); k = 0 / 1: the first / second original tile of the +phi
// columns of this warp's 64 phi, k = 2 / 3: of its -phi columns (a run of 64 columns meets at most two tiles of 256).
// jw = phi index of the warp's first column pair (regular tiles), -1 for the tile of the cell without a mirror.
template <bool F8>
__device__ __forceinline__ int screen_tile_fails_sym(const GemmEpilogue& epi, const float (&acc)[EPI_COLS], int lane, int row0,
                                                     int col0, int ncols, int Khalf, int jw) {
    const int quad = lane & 3;
    const int part = (lane >> 2) & 1;  // (row0 and the row offsets 8 rh + 16 lh are even)
    const int nphi = epi.N >> 1;
    // columns of this lane's part: + side  base_p + j  (Re: j, Im: n + j),  - side  base_m - j  (Re: n - j, Im: 2n - j);
    // kk = phi offset inside the warp's run; the original tile changes at kk = split
    const int split_p = jw < 0 ? 64 : 256 - (((part ? nphi : 0) + jw) & 255);
    const int split_m = jw < 0 ? 64 : (((part ? 2 * nphi : nphi) - jw) & 255) + 1;
    int fail = 0;
#pragma unroll
    for (int lh = 0; lh < 2; ++lh) {
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
            const int p = (row0 + (lane >> 2) + 8 * rh + 16 * lh) >> 1;  // (both lanes of an exchange share p)
            const bool row_ok = p < epi.P;
            float lim = 0.f;
            bool active = false;
            if (row_ok) {
                const float l0 = __ldg(epi.lambda0 + p), nz = __ldg(epi.noise + p);
                const float lam = l0 - (float)epi.iter * nz;
                active = (nz < l0) && (epi.iter == 0 || lam > 0.f) && (!epi.iter_cap || epi.iter < __ldg(epi.iter_cap + p));
                const float sp = __ldg(epi.abs_sum + p);
                lim = F8 ? lam * __ldg(epi.a_scale + p) * (kTwiddleScale8 * kOperandScale8) - (68.f * sp + 0.6f * (float)Khalf)
                         : lam * __ldg(epi.a_scale + p) * (1.f / kTwiddleDescale) - screen_slack(2 * Khalf) * sp;
            }
            // fast path: one maximum over the thread's 16 column pairs.  |C| + |S'| >= 0 or NaN: as unsigned bit patterns
            // they order like the floats with NaN on top, so an integer max carries a NaN through to the comparison
            unsigned int mxi = 0u;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int col = col0 + 8 * jj + 2 * quad;
                const int ai = (((lh * 2 + (jj >> 3)) * 8 + (jj & 7)) * 2 + rh) * 2;
                // own (C, S); the partner row's S (all lanes take part in the exchange)
                const float s_other = __shfl_xor_sync(0xffffffffu, acc[ai + 1], 4);
                if (col < ncols) mxi = max(mxi, __float_as_uint(fabsf(acc[ai]) + fabsf(s_other)));
            }
            const bool ok = !(row_ok && active) || ((lim > 0.f) && (__uint_as_float(mxi) < lim));
            if (__all_sync(0xffffffffu, ok)) continue;  // (warp-uniform: the slow path exchanges again)
            // slow path (tiles near the sources): which original tile do the failing columns belong to?
            float mp0 = 0.f, mp1 = 0.f, mm0 = 0.f, mm1 = 0.f;  // maxima over the + / - columns in their first / second tile
            bool bad = false;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int col = col0 + 8 * jj + 2 * quad;
                const int ai = (((lh * 2 + (jj >> 3)) * 8 + (jj & 7)) * 2 + rh) * 2;
                const float s_other = __shfl_xor_sync(0xffffffffu, acc[ai + 1], 4);
                if (col >= ncols) continue;
                const float v = fabsf(acc[ai]) + fabsf(s_other);
                bad |= (v != v);
                const int kk = 4 * jj + quad;
                if (kk < split_p) mp0 = fmaxf(mp0, v); else mp1 = fmaxf(mp1, v);
                if (kk < split_m) mm0 = fmaxf(mm0, v); else mm1 = fmaxf(mm1, v);
            }
            if (row_ok && active) {
                if (bad | !(lim > 0.f))
                    fail |= 15 << (4 * part);
                else
                    fail |= ((!(mp0 < lim)) | ((!(mp1 < lim)) << 1) | ((!(mm0 < lim)) << 2) | ((!(mm1 < lim)) << 3)) << (4 * part);
            }
        }
    }
    return fail;
}

// The same test when the A tile was loaded through the 4-D map (GemmEpilogue::sym_rows8): rows 16 G + 8 h + l of the tile
// are row h (alpha / beta) of line of sight 8 G + l, so a thread's rows rh = 0 / 1 are the alpha and the beta row of ONE
// line of sight and no exchange is needed.  row0: first row of the warp (a multiple of 32).
// Incremental screening (exact).  A tile that passes has a MARGIN per row: (lim - max |g1|) in real units of the gradient.
// Between two tests the gradient of row p at a phi cell that lies D cells away from every cell whose state changed moves
// by at most  coef(D) ||path of z_p||_1:  delta g = (nf / n) sum_{i live} exp(-2i (l2_i - l2_ref) d) delta z  cell by cell,
// i.e. |delta g| <= coef(D) |delta z| with coef(D) = (|nf| / n) (m_base env[D] + dead_p), env the suffix maximum of the
// normalised |RMTF| of the base channel set (far from a source ~0.03 instead of the trivial 1) and dead_p the channels
// pixel p flags on top of it.  lambda decays by noise_p per iteration, and the rounding of both evaluations is covered
// by an allowance R_p (forward GEMM error <= 2^-12 ||z||_1 per output, operand storage 2^-19 ||r||_1), kept as a running
// maximum rhat_p.  row_update_kernel maintains per row the path length L_p, the additive clock Q_p = it noise_p + rhat_p
// and the union [lo_p, hi_p] of the phi cells that held state since iteration 1.  The test below records per (tile, row)
//     X = 0.999 margin - 2 rhat_p + Q_p,   L0 = L_p        (as of the test)
// and the tile stays provably zero -- neither screened nor listed -- while for all its rows
//     X - Q_p(now) > coef(D(tile, [lo_p, hi_p] now)) (L_p(now) - L0)          (sym_quiet_kernel).
// [margin(it0) > coef dL + (it - it0) noise + R(it) + R(it0) follows: Q(now) - Q(it0) + 2 rhat(it0) >= the last three
// terms; the union of extents only grows, so D(now) <= every earlier distance and env is non-increasing.]  A row that
// fails or has no positive margin records X = 0: screened again next iteration.  Inactive rows and rows beyond P put no
// constraint (+inf from the initialisation).  exp_tile: the 64 entries of this (group, symmetric tile), or null.
template <bool F8>
__device__ __forceinline__ int screen_tile_fails_sym8(const GemmEpilogue& epi, const float (&acc)[EPI_COLS], int lane, int row0,
                                                      int col0, int ncols, int Khalf, int jw, float2* exp_tile) {
    const int quad = lane & 3;
    const int nphi = epi.N >> 1;
    int fail = 0;
#pragma unroll
    for (int lh = 0; lh < 2; ++lh) {
        const int p = (row0 >> 1) + 8 * lh + (lane >> 2);
        const bool row_ok = p < epi.P;
        float lim = 0.f;
        bool active = false;
        unsigned int exp_bits = 0x7f800000u;  // this thread's bound on the row's expiry (+inf = no constraint)
        if (row_ok) {
            const float l0 = __ldg(epi.lambda0 + p), nz = __ldg(epi.noise + p);
            const float lam = l0 - (float)epi.iter * nz;
            active = (nz < l0) && (epi.iter == 0 || lam > 0.f) && (!epi.iter_cap || epi.iter < __ldg(epi.iter_cap + p));
            const float sp = __ldg(epi.abs_sum + p);
            lim = F8 ? lam * __ldg(epi.a_scale + p) * (kTwiddleScale8 * kOperandScale8) - (68.f * sp + 0.6f * (float)Khalf)
                     : lam * __ldg(epi.a_scale + p) * (1.f / kTwiddleDescale) - screen_slack(2 * Khalf) * sp;
        }
        unsigned int mre = 0u, mim = 0u;  // (integer max of non-negative floats: a NaN ends up on top)
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int col = col0 + 8 * jj + 2 * quad;
            const int aa = (((lh * 2 + (jj >> 3)) * 8 + (jj & 7)) * 2 + 0) * 2, ab = aa + 2;  // alpha row (rh = 0), beta row (rh = 1)
            if (col < ncols) {
                mre = max(mre, __float_as_uint(fabsf(acc[aa]) + fabsf(acc[ab + 1])));      // |C_alpha| + |S_beta|
                mim = max(mim, __float_as_uint(fabsf(acc[ab]) + fabsf(acc[aa + 1])));      // |C_beta| + |S_alpha|
            }
        }
        const bool limok = lim > 0.f;
        const bool pass = !(row_ok && active) || (limok && __uint_as_float(mre) < lim && __uint_as_float(mim) < lim);
        if (exp_tile != nullptr) {  // kernel-uniform; the four lanes of a row agree on one value, one of them records it
            if (row_ok && active) {
                exp_bits = 0u;
                if (pass) {
                    const float unit = F8 ? 1.f / (__ldg(epi.a_scale + p) * (kTwiddleScale8 * kOperandScale8))
                                          : kTwiddleDescale / __ldg(epi.a_scale + p);
                    const float margin = (lim - fmaxf(__uint_as_float(mre), __uint_as_float(mim))) * unit * 0.999f;
                    exp_bits = __float_as_uint(fmaxf(margin - 2.f * __ldg(epi.rhat + p) + __ldg(epi.qclock + p), 0.f));
                }
            }
            exp_bits = min(exp_bits, __shfl_xor_sync(0xffffffffu, exp_bits, 1));
            exp_bits = min(exp_bits, __shfl_xor_sync(0xffffffffu, exp_bits, 2));
            if (quad == 0 && exp_bits != 0x7f800000u) atomicMin(reinterpret_cast<unsigned int*>(exp_tile + (p & 63)), exp_bits);
        }
        if (pass) continue;
        // slow path (tiles near the sources): which original tiles do the failing columns belong to?
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            const int split_p = jw < 0 ? 64 : 256 - (((part ? nphi : 0) + jw) & 255);
            const int split_m = jw < 0 ? 64 : (((part ? 2 * nphi : nphi) - jw) & 255) + 1;
            float mp0 = 0.f, mp1 = 0.f, mm0 = 0.f, mm1 = 0.f;
            bool bad = !limok;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const int col = col0 + 8 * jj + 2 * quad;
                if (col >= ncols) continue;
                const int aa = (((lh * 2 + (jj >> 3)) * 8 + (jj & 7)) * 2 + 0) * 2, ab = aa + 2;
                const float v = part ? fabsf(acc[ab]) + fabsf(acc[aa + 1]) : fabsf(acc[aa]) + fabsf(acc[ab + 1]);
                bad |= (v != v);
                const int kk = 4 * jj + quad;
                if (kk < split_p) mp0 = fmaxf(mp0, v); else mp1 = fmaxf(mp1, v);
                if (kk < split_m) mm0 = fmaxf(mm0, v); else mm1 = fmaxf(mm1, v);
            }
            if (bad)
                fail |= 15 << (4 * part);
            else
                fail |= ((!(mp0 < lim)) | ((!(mp1 < lim)) << 1) | ((!(mm0 < lim)) << 2) | ((!(mm1 < lim)) << 3)) << (4 * part);
        }
    }
    return fail;
}

// The plain adjoint from the four partial sums (rows regrouped by the 4-D map: a thread holds the alpha and the beta row
// of one line of sight): out[p, j] = C_a + S_b, out[p, n - j] = C_a - S_b, out[p, n + j] = C_b - S_a, out[p, 2n - j] =
// C_b + S_a, times the operand descale.  jw: phi index of the warp's first column pair, -1 for the cell without a mirror.
__device__ __forceinline__ void epilogue_plain_sym8(const GemmEpilogue& epi, const float (&acc)[EPI_COLS], int lane, int row0,
                                                    int col0, int ncols, int jw) {
    const int quad = lane & 3;
    const int nphi = epi.N >> 1;
#pragma unroll
    for (int lh = 0; lh < 2; ++lh) {
        const int p = (row0 >> 1) + 8 * lh + (lane >> 2);
        if (p >= epi.P) continue;
        const float descale = kTwiddleDescale / __ldg(epi.a_scale + p);
        float* o = epi.out + (size_t)p * epi.ld_out;
#pragma unroll
        for (int jj = 0; jj < 16; ++jj) {
            const int col = col0 + 8 * jj + 2 * quad;
            if (col >= ncols) continue;
            const int aa = (((lh * 2 + (jj >> 3)) * 8 + (jj & 7)) * 2 + 0) * 2, ab = aa + 2;
            const float ca = acc[aa], sa = acc[aa + 1], cb = acc[ab], sb = acc[ab + 1];
            if (jw < 0) {  // phi_0 itself (its own table row)
                __stcs(o, (ca + sb) * descale);
                __stcs(o + nphi, (cb - sa) * descale);
            } else {
                const int j = jw + 4 * jj + quad;
                __stcs(o + j, (ca + sb) * descale);
                __stcs(o + nphi - j, (ca - sb) * descale);
                __stcs(o + nphi + j, (cb - sa) * descale);
                __stcs(o + 2 * nphi - j, (cb + sa) * descale);
            }
        }
    }
}

// One warp per PAIR tile (two groups of 64 lines of sight x one symmetric column tile) decides what the pair kernel has to
// do.  Per group: symq = 1 when some original tile this symmetric tile overlaps is quiet (its state is identically zero)
// -- there is something to decide -- AND the margins of the last test no longer cover what happened since (incremental
// screening, see above).  A tile that is going to be tested gets its entries reset to {+inf, L_p} (the test takes minima
// of the first), one without a quiet original tile gets {0, .} (nothing is proven for it).  Pair tiles with work are
// appended to `list` (numbering of tile_coords(t, MTP, NTs)); original tiles that hold state are marked in `need` (zeroed
// by the caller): they go on to the next pass whatever the test says.
__global__ void sym_select_kernel(const unsigned int* __restrict__ tile_state, int flag_ld, int MTg, int NTs, int nphi,
                                  unsigned char* __restrict__ symq, unsigned char* __restrict__ need, int need_ld,
                                  int* __restrict__ list, int* __restrict__ count, float2* __restrict__ expiry, IncrState st,
                                  int P, int record) {
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int MTP = (MTg + 1) / 2;
    if (idx >= MTP * NTs) return;
    const int mtp = idx / NTs, nt = idx - mtp * NTs;
    int lo[4], hi[4];
    const int nr = sym_ranges(nt, NTs, nphi, lo, hi);
    // phi cells this symmetric tile decides: [a0, a1] and its mirror [b0, b1] (the cell without a mirror: [0, 0])
    const int a0 = nt == NTs - 1 ? 0 : lo[0], a1 = nt == NTs - 1 ? 0 : hi[0] - 1;
    const int b0 = nt == NTs - 1 ? 0 : lo[1], b1 = nt == NTs - 1 ? 0 : hi[1] - 1;
    int work = 0;
    for (int g = 0; g < 2; ++g) {
        const int mt = 2 * mtp + g;
        if (mt >= MTg) break;
        const unsigned int mask = (mt & 1) ? 0xFFFF0000u : 0x0000FFFFu;
        const unsigned int* row = tile_state + (size_t)(mt >> 1) * flag_ld;
        unsigned char* need_row = need + (size_t)(mt >> 1) * need_ld;
        int any_quiet = 0;
        for (int r = 0; r < nr; ++r)
            for (int t = (lo[r] >> 8) + lane; t <= (hi[r] - 1) >> 8; t += 32) {
                unsigned int wd = row[2 * t];
                if (2 * t + 1 < flag_ld) wd |= row[2 * t + 1];
                if (wd & mask) need_row[t] = 1;  // an original tile with state
                else any_quiet = 1;
            }
        any_quiet = __any_sync(0xffffffffu, any_quiet);
        int screen = any_quiet;
        if (expiry != nullptr) {
            float2* e = expiry + ((size_t)mt * NTs + nt) * 64;
            bool valid = any_quiet != 0;
            float len[2] = {0.f, 0.f};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int p = mt * 64 + lane + 32 * h;
                if (p >= P) continue;
                len[h] = st.pathlen[p];
                if (!valid) continue;
                const float2 rec = e[lane + 32 * h];
                if (!(rec.x > 0.f)) {
                    valid = false;
                } else if (rec.x != __int_as_float(0x7f800000)) {
                    const int2 ex = st.ext[p];
                    int dist = nphi - 1;
                    if (ex.x <= ex.y) {
                        const int da = max(0, max(a0 - ex.y, ex.x - a1)), db = max(0, max(b0 - ex.y, ex.x - b1));
                        dist = min(min(da, db), nphi - 1);
                    }
                    const float2 c = st.crow[p];
                    const float coef = c.x * st.env[dist] + c.y;
                    if (!(rec.x - st.qclock[p] > coef * (len[h] - rec.y) * 1.001f)) valid = false;
                }
            }
            valid = __all_sync(0xffffffffu, valid);
            if (valid) {
                screen = 0;  // still provably zero: neither tested nor listed
            } else {
                const float v = (any_quiet && record) ? __int_as_float(0x7f800000) : 0.f;
                e[lane] = make_float2(v, len[0]);
                e[lane + 32] = make_float2(v, len[1]);
            }
        }
        if (lane == 0) symq[(size_t)mt * NTs + nt] = (unsigned char)screen;
        work |= screen;
    }
    if (work && lane == 0) {  // (tile id in the numbering of tile_coords(t, MTP, NTs): invert it)
        const int g = mtp / RASTER_GROUP, first = g * RASTER_GROUP;
        const int gsize = min(RASTER_GROUP, MTP - first);
        list[atomicAdd(count, 1)] = g * RASTER_GROUP * NTs + nt * gsize + (mtp - first);
    }
}

constexpr int P2_STAGES = 6;
constexpr int P2_STAGE_BYTES = 2 * A_TILE_BYTES;  // A (128 rows x 128 B) + half of B (128 rows x 128 B)
constexpr int P2_SMEM_BYTES = 1024 + P2_STAGES * P2_STAGE_BYTES + 256;

// PMODE 0: the fp8 screening pass.  PMODE 1: the one-product fp16 plain adjoint (EPI_PLAIN1) -- same staging (a 128-byte
// row holds 64 fp16 instead of 128 fp8), kind::f16 MMAs, the plain epilogue, every listed pair computed by both CTAs.
// PMODE 2: the fp16 screening test over all quiet tiles.  PMODE 3 / 4 / 5: modes 0 / 2 / 1 in the mirror-symmetric
// formulation (see "mirror-symmetric screening" above): MT counts groups of 64 lines of sight, NT tiles of 128 phi >= 0.
template <int PMODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
screen8_pair_kernel(const __grid_constant__ CUtensorMap map_a8, const __grid_constant__ CUtensorMap map_b8, int K, int MT,
                    int NT, const GemmEpilogue epi) {
    constexpr bool PLAIN = PMODE == 1 || PMODE == 5;  // PMODE 5: the one-product plain adjoint, mirror-symmetric (rows8 map)
    constexpr bool F16 = PMODE == 1 || PMODE == 2 || PMODE == 4 || PMODE == 5;  // PMODE 2: the fp16 screening test over all quiet tiles
    constexpr bool SYM = PMODE >= 3;          // PMODE 3 / 4: mirror-symmetric fp8 / fp16 screening (MT counts 64-LOS groups)
    constexpr int BKE = F16 ? BK : 2 * BK;    // K elements per 128-byte row
    const int nphi = epi.N >> 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + P2_STAGES * P2_STAGE_BYTES);
    uint64_t* full = bars;                              // [P2_STAGES]  both CTAs' TMA -> leader's MMA thread
    uint64_t* empty = bars + P2_STAGES;                 // [P2_STAGES]  MMA -> each CTA's producer (multicast commit)
    uint64_t* tmem_full = bars + 2 * P2_STAGES;         // [2]  MMA -> each CTA's epilogue warps (multicast commit)
    uint64_t* tmem_empty = bars + 2 * P2_STAGES + 2;    // [2]  epilogue warps of BOTH CTAs -> leader's MMA thread
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * P2_STAGES + 4);
    int* last_listed = reinterpret_cast<int*>(bars + 2 * P2_STAGES + 5);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int MTP = (MT + 1) / 2;
    const bool listed = (PLAIN || SYM) && epi.tile_list != nullptr;  // pair tiles to process (sym_select_kernel / pair lists)
    const int total = listed ? __ldcg(epi.tile_count) : MTP * NT;
    const int KB = (K + BKE - 1) / BKE;
    auto pair_id = [&](int ti) -> int { return listed ? __ldcg(epi.tile_list + ti) : ti; };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a8);
        tma_prefetch_desc(&map_b8);
        for (int s = 0; s < P2_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 2 * EPI_WARPS);
        }
        *last_listed = -1;
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(tmem_ptr, TMEM_COLS);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    auto quiet = [&](int mt, int nt) -> bool {  // does pixel tile mt exist and is its state zero in this column tile?
        if (PLAIN) return mt < MT;  // (plain mode: every existing tile of a listed pair is computed)
        if (mt >= MT) return false;
        if (SYM) return __ldcg(epi.symq + (size_t)mt * NT + nt) != 0;  // is any original tile it overlaps quiet?
        const unsigned int* st = epi.tile_state + (size_t)mt * epi.flag_ld + 2 * nt;
        unsigned int wd = __ldcg(st);
        if (2 * nt + 1 < epi.flag_ld) wd |= __ldcg(st + 1);
        return wd == 0u;
    };
    // symmetric modes: the quiet bytes of both groups of pair tile ti (bits 0-7: CTA 0's group, 8-15: CTA 1's), fetched one
    // tile ahead so that no role waits for an L2 round trip at the top of a tile
    auto qraw = [&](int ti) -> unsigned int {
        if (!SYM || ti >= total) return 0u;
        if (PLAIN) {  // every existing group of a listed pair is computed
            int mtp, nt;
            tile_coords(pair_id(ti), MTP, NT, mtp, nt);
            return (2 * mtp < MT ? 1u : 0u) | (2 * mtp + 1 < MT ? 0x100u : 0u);
        }
        int mtp, nt;
        tile_coords(pair_id(ti), MTP, NT, mtp, nt);
        const unsigned int a = 2 * mtp < MT ? __ldcg(epi.symq + (size_t)(2 * mtp) * NT + nt) : 0u;
        const unsigned int b = 2 * mtp + 1 < MT ? __ldcg(epi.symq + (size_t)(2 * mtp + 1) * NT + nt) : 0u;
        return a | (b << 8);
    };

    if (warp < 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
      if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            unsigned int qn = qraw(cluster_id);
            for (int ti = cluster_id; ti < total; ti += nclusters) {
                int mtp, nt;
                tile_coords(pair_id(ti), MTP, NT, mtp, nt);
                const unsigned int qc = qn;
                qn = qraw(ti + nclusters);
                if (SYM ? qc == 0u : (!quiet(2 * mtp, nt) && !quiet(2 * mtp + 1, nt))) continue;
                const int mt = 2 * mtp + (int)rank;
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sb = stage_base + stage * P2_STAGE_BYTES;
                    if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * P2_STAGE_BYTES);
                    if (SYM && epi.sym_rows8)  // rows regrouped by the 4-D map: 8 groups of (alpha rows | beta rows) of 8 LOS
                        tma_load_4d_pair(sb, &map_a8, &full[stage], kb * BKE, 0, 0, mt * 8);
                    else
                        tma_load_2d_pair(sb, &map_a8, &full[stage], kb * BKE, mt * BM);
                    const int brow = SYM ? (nt == NT - 1 ? 0 : nphi + nt * BN) : nt * BN;  // (half-row view: rows of j >= n/2)
                    tma_load_2d_pair(sb + A_TILE_BYTES, &map_b8, &full[stage], kb * BKE, brow + (int)rank * (BN / 2));
                    if (++stage == P2_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
      } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN);  // M = 256 over the pair; format code 0 = e4m3 here
            int stage = 0;
            uint32_t phase = 0, use = 0;
            unsigned long long issued = 0;
            unsigned int qn = qraw(cluster_id);
            for (int ti = cluster_id; ti < total; ti += nclusters) {
                int mtp, nt;
                tile_coords(pair_id(ti), MTP, NT, mtp, nt);
                const unsigned int qc = qn;
                qn = qraw(ti + nclusters);
                if (SYM ? qc == 0u : (!quiet(2 * mtp, nt) && !quiet(2 * mtp + 1, nt))) continue;
                const int buf = use & 1;
                mbar_wait(&tmem_empty[buf], ((use >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(stage_base + stage * P2_STAGE_BYTES);
                    const uint32_t sb = sa + A_TILE_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < 4; ++k4) {
                        if (F16)
                            umma_f16_pair(d_tmem, umma_desc_sw128(sa + k4 * 32), umma_desc_sw128(sb + k4 * 32), idesc,
                                          (kb != 0 || k4 != 0) ? 1u : 0u);
                        else
                            umma_f8_pair(d_tmem, umma_desc_sw128(sa + k4 * 32), umma_desc_sw128(sb + k4 * 32), idesc,
                                         (kb != 0 || k4 != 0) ? 1u : 0u);
                    }
                    umma_commit_pair(&empty[stage]);
                    if (++stage == P2_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                    issued += F16 ? 2 : 4;  // two CTAs x one 128-element fp8 K block = 4 units of 128 x 256 x 64
                }
                umma_commit_pair(&tmem_full[buf]);
                ++use;
            }
            if (epi.mma_count != nullptr && issued != 0) atomicAdd(epi.mma_count + epi.count_slot, issued);
        }
      }
    } else {
        // ===================== epilogue warps (both CTAs) =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int ew = warp - 4;
        const int sub = warp & 3;
        const int half = ew >> 2;
        uint32_t use = 0;
        int ti = 0;  // running index of this CTA's tiles (for the one-lister-per-tile rule)
        unsigned int qn = qraw(cluster_id);
        for (int tq = cluster_id; tq < total; tq += nclusters, ++ti) {
            int mtp, nt;
            tile_coords(pair_id(tq), MTP, NT, mtp, nt);
            const int mt = 2 * mtp + (int)rank;
            const unsigned int qc = qn;
            qn = qraw(tq + nclusters);
            const bool q_own = SYM ? ((qc >> (8 * rank)) & 0xffu) != 0u : quiet(mt, nt);
            const bool q_other = SYM ? ((qc >> (8 * (1 - rank))) & 0xffu) != 0u : quiet(2 * mtp + 1 - (int)rank, nt);
            const bool exists = mt < MT;
            bool list_it = exists && !q_own;  // a tile with signal goes straight to the next pass of the cascade
            int sym_fail = 0;
            if (q_own || q_other) {
                const int buf = use & 1;
                mbar_wait(&tmem_full[buf], (use >> 1) & 1);
                tc_fence_after();
                if (q_own) {
                    float acc[EPI_COLS];
                    load_chunk(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(buf * BN + half * EPI_COLS), acc);
                    if (PLAIN) {  // the accumulator is in registers: hand the TMEM buffer back before the global stores
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (rank == 0) mbar_arrive(&tmem_empty[buf]);
                            else mbar_arrive_leader(&tmem_empty[buf]);
                        }
                        if (SYM)
                            epilogue_plain_sym8(epi, acc, lane, mt * BM + sub * 32, half * EPI_COLS, nt == NT - 1 ? 2 : nphi - nt * BN,
                                                nt == NT - 1 ? -1 : nphi / 2 + (BN / 2) * nt + (EPI_COLS / 2) * half);
                        else
                            epilogue_tile<EPI_PLAIN>(epi, acc, lane, mt * BM + sub * 32, nt * BN + half * EPI_COLS);
                    } else if (SYM) {
                        const int jw = nt == NT - 1 ? -1 : nphi / 2 + (BN / 2) * nt + (EPI_COLS / 2) * half;
                        const int ncv = nt == NT - 1 ? 2 : nphi - nt * BN;
                        float2* exp_tile = epi.expiry ? epi.expiry + ((size_t)mt * NT + nt) * 64 : nullptr;
                        const int fail = epi.sym_rows8
                                             ? screen_tile_fails_sym8<!F16>(epi, acc, lane, mt * BM + sub * 32, half * EPI_COLS, ncv, K, jw, exp_tile)
                                             : screen_tile_fails_sym<!F16>(epi, acc, lane, mt * BM + sub * 32, half * EPI_COLS, ncv, K, jw);
                        sym_fail = (int)__reduce_or_sync(0xffffffffu, (unsigned)fail);
                        list_it = sym_fail != 0;
                    } else {
                        const bool fail = screen_tile_fails<!F16>(epi, acc, lane, mt * BM + sub * 32, nt * BN + half * EPI_COLS, K);
                        list_it = __any_sync(0xffffffffu, fail) != 0;
                    }
                }
                if (!(PLAIN && q_own)) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (rank == 0) mbar_arrive(&tmem_empty[buf]);
                        else mbar_arrive_leader(&tmem_empty[buf]);
                    }
                }
                ++use;
            }
            if (PLAIN) continue;
            if (SYM) {  // original tiles that go on: those with state (marked by sym_select_kernel) and those a failed bound touches
                if (exists && lane == 0) {
                    unsigned char* row = epi.need + (size_t)(mt >> 1) * epi.need_ld;
                    if (sym_fail) {  // the original tiles the failed bounds of this warp belong to
                        if (nt == NT - 1) {  // the cell without a mirror: columns 0 (Re) and n (Im)
                            if (sym_fail & 0x0f) row[0] = 1;
                            if (sym_fail & 0xf0) row[nphi >> 8] = 1;
                        } else {
                            const int jw = nphi / 2 + (BN / 2) * nt + (EPI_COLS / 2) * half;
                            for (int part = 0; part < 2; ++part) {
                                const int bits = (sym_fail >> (4 * part)) & 15;
                                const int tp = ((part ? nphi : 0) + jw) >> 8;             // + side: tiles tp, tp + 1
                                const int tm = ((part ? 2 * nphi : nphi) - jw) >> 8;      // - side: tiles tm, tm - 1
                                if (bits & 1) row[tp] = 1;
                                if ((bits & 2) && tp + 1 < epi.need_ld) row[tp + 1] = 1;
                                if (bits & 4) row[tm] = 1;
                                if ((bits & 8) && tm > 0) row[tm - 1] = 1;
                            }
                        }
                    }
                }
                continue;
            }
            if (list_it && lane == 0 && atomicMax(last_listed, ti) < ti) {
                // tile id in the numbering of dft_gemm_kernel's tile_coords(t, MT, NT): invert it
                const int g = mt / RASTER_GROUP, first = g * RASTER_GROUP;
                const int gsize = min(RASTER_GROUP, MT - first);
                epi.out_list[atomicAdd(epi.out_count, 1)] = g * RASTER_GROUP * NT + nt * gsize + (mt - first);
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
}

// ---- three-product transforms on CTA pairs (cta_group::2) -----------------------------------------------------------
// dft_gemm_kernel's arithmetic -- the same three fp16 products per K block, the same window-aligned accumulation chunks
// drained into registers, the same K-block skipping -- issued as M = 256 MMAs over the two SMs of a TPC: each CTA stages its
// own A tile (hi, lo) and only HALF of the B tile (hi, lo), 64 KiB per K block and SM instead of 96.  In the sparse regime
// the forward transform is bound by operand delivery (L2 -> SM) and by its epilogue, not by the tensor pipe: the pair moves
// a third less per tile, and the shared memory it frees stages the residual epilogue's global traffic through TMA -- the b
// tile arrives by bulk tensor loads issued before the K loop ends (two 4 KiB passes in flight per warp), the fp16 hi / lo
// operand of the adjoint leaves by bulk tensor stores from a swizzled staging tile -- instead of 8-row x 32-byte accesses
// per warp instruction.  A pair tile skips the K blocks that are zero for BOTH of its pixel tiles; a block that is zero for
// one of them adds exact zeros there, so the results equal the one-CTA kernel's bit for bit.
constexpr int P3_STAGES = 2;
constexpr int P3_STAGE_BYTES = 4 * A_TILE_BYTES;      // A_hi, A_lo, half of B_hi, half of B_lo: 4 x (128 rows x 128 B) = 64 KiB
constexpr int P3_EPI_WARP_BYTES = 12 * 1024;          // per epilogue warp: two 4 KiB passes of b + 4 KiB of output staging
constexpr int P3_BAR_BYTES = 512;
constexpr int P3_SMEM_BYTES = 1024 + P3_STAGES * P3_STAGE_BYTES + EPI_WARPS * P3_EPI_WARP_BYTES + P3_BAR_BYTES + NUM_WARPS_TOTAL * KMASK_WORDS * 4;
static_assert(P3_SMEM_BYTES <= 232448, "shared memory budget exceeded");

// K-block activity of a pair tile: the union of its two pixel tiles' flag rows (row1 may be null)
__device__ __forceinline__ int build_kmask2(const unsigned int* row0, const unsigned int* row1, int nwords, int KB, uint32_t* mask,
                                            int lane) {
    int count = 0;
    for (int w0 = 0; w0 < nwords; w0 += 32) {
        const int idx = w0 + lane;
        unsigned int v = idx < nwords ? __ldcg(row0 + idx) : 0u;
        if (row1 != nullptr && idx < nwords) v |= __ldcg(row1 + idx);
        const uint32_t bits = __ballot_sync(0xffffffffu, v != 0u);
        if (lane == 0) mask[w0 >> 5] = bits;
        count += 2 * __popc(bits);
    }
    __syncwarp();
    if ((KB & 1) && ((mask[(KB >> 1) >> 5] >> ((KB >> 1) & 31)) & 1u)) count -= 1;
    return count;
}

// EPI_RESID with its global traffic staged through shared memory (see above).  One warp, its 32 rows x 128 columns, in four
// passes of 32 columns.  bbuf: two 4 KiB pass buffers (128-byte-swizzled boxes of 32 fp32 x 32 rows), bbar their barriers,
// bphase their parities; passes 0 and 1 of this tile were requested by resid_tma_request before the K loop ended.  obuf:
// 2 KiB hi + 2 KiB lo (64-byte-swizzled boxes of 32 fp16 x 32 rows).  Same arithmetic, same bits as epilogue_tile<EPI_RESID>;
// the fp8 copy (1 byte per element) keeps its direct stores.  Columns / rows beyond the matrix are clipped by the TMA unit.
__device__ __forceinline__ void resid_tma_request(const GemmEpilogue& epi, const CUtensorMap* map_b, uint8_t* bbuf, uint64_t* bbar,
                                                  int pass, int row0, int col0) {
    mbar_arrive_expect_tx(&bbar[pass & 1], 4096);
    tma_load_2d(bbuf + (pass & 1) * 4096, map_b, &bbar[pass & 1], col0 + 32 * pass, row0);
}
__device__ __forceinline__ void epilogue_resid_tma(const GemmEpilogue& epi, const float (&acc)[EPI_COLS], int lane, int row0, int col0,
                                                   const CUtensorMap* map_b, const CUtensorMap* map_hi, const CUtensorMap* map_lo,
                                                   uint8_t* bbuf, uint8_t* obuf, uint64_t* bbar, uint32_t (&bphase)[2]) {
    const int quad = lane & 3, r8 = lane >> 2;
    float descale[4], oscale[4], dsum[4], rmax[4];
    const float* cs_row[4];
    bool row_ok[4];
    const bool cs_even = (epi.m & 1) == 0 && (reinterpret_cast<uintptr_t>(epi.chan_scale) & 7) == 0;
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) {  // ri = 2 lh + rh
        const int row = row0 + r8 + 8 * (ri & 1) + 16 * (ri >> 1);
        row_ok[ri] = row < epi.P;
        descale[ri] = row_ok[ri] ? kTwiddleDescale / __ldg(epi.a_scale + row) : 0.f;
        oscale[ri] = row_ok[ri] ? __ldg(epi.out_scale + row) : 0.f;
        cs_row[ri] = (epi.chan_scale && row_ok[ri]) ? epi.chan_scale + (epi.chan_per_pixel ? (size_t)row * epi.m : 0) : nullptr;
        dsum[ri] = 0.f;
        rmax[ri] = 0.f;
    }
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        const int pc0 = col0 + 32 * pass;
        const uint8_t* bsrc = bbuf + (pass & 1) * 4096;
        mbar_wait(&bbar[pass & 1], bphase[pass & 1]);
        bphase[pass & 1] ^= 1;
        if (lane == 0) tma_store_wait_read<0>();  // the previous pass's stores have read the staging tile
        __syncwarp();
#pragma unroll
        for (int ri = 0; ri < 4; ++ri) {
            const int rin = r8 + 8 * (ri & 1) + 16 * (ri >> 1);
            const int row = row0 + rin;
#pragma unroll
            for (int jl = 0; jl < 4; ++jl) {
                const int col = pc0 + 8 * jl + 2 * quad;
                const int ai = ((((ri >> 1) * 2 + (pass >> 1)) * 8 + (4 * (pass & 1) + jl)) * 2 + (ri & 1)) * 2;
                const bool ok = row_ok[ri] && col < epi.N;
                float2 cs = make_float2(1.f, 1.f);
                if (cs_row[ri] != nullptr && ok) {
                    const int c0 = col >= epi.m ? col - epi.m : col;
                    if (cs_even) {
                        cs = __ldg(reinterpret_cast<const float2*>(cs_row[ri] + c0));
                    } else {
                        const int c1 = col + 1 >= epi.m ? col + 1 - epi.m : col + 1;
                        cs = make_float2(__ldg(cs_row[ri] + c0), __ldg(cs_row[ri] + c1));
                    }
                }
                const int bb = (8 * jl + 2 * quad) * 4;  // byte offset of the column pair within the 128-byte row of b
                const float2 bv = *reinterpret_cast<const float2*>(bsrc + rin * 128 + ((((bb >> 4) ^ (rin & 7)) << 4) | (bb & 15)));
                const float a0 = acc[ai] * descale[ri], a1 = acc[ai + 1] * descale[ri];
                const float r0 = __fmaf_rn(cs.x, a0, -bv.x) * oscale[ri], r1 = __fmaf_rn(cs.y, a1, -bv.y) * oscale[ri];
                const __half2 h = __floats2half2_rn(r0, r1);
                const float2 hf = __half22float2(h);
                const __half2 l = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
                const int ob = (8 * jl + 2 * quad) * 2;  // byte offset within the 64-byte row of the staging boxes
                const int so = rin * 64 + ((((ob >> 4) ^ ((rin >> 1) & 3)) << 4) | (ob & 15));
                *reinterpret_cast<__half2*>(obuf + so) = h;
                *reinterpret_cast<__half2*>(obuf + 2048 + so) = l;
                if (ok) {
                    dsum[ri] += fabsf(hf.x) + fabsf(hf.y);
                    rmax[ri] = fmaxf(rmax[ri], fmaxf(fabsf(r0), fabsf(r1)));
                    if (epi.out8 != nullptr) {
                        const float2 v8 = make_float2(hf.x * kOperandScale8, hf.y * kOperandScale8);
                        if (fmaxf(fabsf(v8.x), fabsf(v8.y)) > 448.f) dsum[ri] = __int_as_float(0x7f800000);  // unscreenable row
                        const __nv_fp8x2_storage_t pk = __nv_cvt_float2_to_fp8x2(v8, __NV_SATFINITE, __NV_E4M3);
                        unsigned char* o8 = epi.out8 + (size_t)row * epi.ld8;
                        const int h8 = epi.ld8 >> 1;
                        if (epi.m & 1) {
                            o8[col < epi.m ? col : col - epi.m + h8] = (unsigned char)(pk & 0xff);
                            o8[col + 1 < epi.m ? col + 1 : col + 1 - epi.m + h8] = (unsigned char)(pk >> 8);
                        } else {
                            *reinterpret_cast<__nv_fp8x2_storage_t*>(o8 + (col < epi.m ? col : col - epi.m + h8)) = pk;
                        }
                    }
                }
            }
        }
        fence_proxy_async();  // the staging tile is complete for the async proxy; the reads of b precede its refill
        __syncwarp();
        if (lane == 0) {
            if (pass + 2 < 4) resid_tma_request(epi, map_b, bbuf, bbar, pass + 2, row0, col0);
            if (pc0 < epi.N) {
                tma_store_2d(map_hi, obuf, pc0, row0);
                tma_store_2d(map_lo, obuf + 2048, pc0, row0);
                tma_store_commit();
            }
        }
    }
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) {
        const int row = row0 + r8 + 8 * (ri & 1) + 16 * (ri >> 1);
        if (epi.abs_sum_out != nullptr) {  // kernel-uniform
            float d = dsum[ri];
            d += __shfl_xor_sync(0xffffffffu, d, 1);
            d += __shfl_xor_sync(0xffffffffu, d, 2);
            if (quad == 0 && row_ok[ri] && d != 0.f) atomicAdd(epi.abs_sum_out + row, d);
        }
        if (epi.row_stat != nullptr) {  // kernel-uniform
            float v = rmax[ri];
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
            v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
            if (quad == 0 && row_ok[ri] && v != 0.f)
                atomicMax(reinterpret_cast<unsigned int*>(epi.row_stat + 4 * (size_t)row + 3), __float_as_uint(v));
        }
    }
}

template <int MODE, bool TMA_EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
dft_gemm_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                     const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                     const __grid_constant__ CUtensorMap map_e_in, const __grid_constant__ CUtensorMap map_e_hi,
                     const __grid_constant__ CUtensorMap map_e_lo, int K, int MT, int NT, int chunk_kb, const GemmEpilogue epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint8_t* epi_base = smem + P3_STAGES * P3_STAGE_BYTES;                       // (1024-byte aligned)
    uint64_t* bars = reinterpret_cast<uint64_t*>(epi_base + EPI_WARPS * P3_EPI_WARP_BYTES);
    uint64_t* full = bars;                              // [P3_STAGES]  both CTAs' TMA -> leader's MMA thread
    uint64_t* empty = bars + P3_STAGES;                 // [P3_STAGES]  MMA -> each CTA's producer (multicast commit)
    uint64_t* tmem_full = bars + 2 * P3_STAGES;         // [2]  MMA -> each CTA's epilogue warps (multicast commit)
    uint64_t* tmem_empty = bars + 2 * P3_STAGES + 2;    // [2]  epilogue warps of BOTH CTAs -> leader's MMA thread
    uint64_t* bbar_all = bars + 2 * P3_STAGES + 4;      // [EPI_WARPS][2]  b passes of the staged epilogue
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * P3_STAGES + 4 + 2 * EPI_WARPS);
    uint32_t* masks = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(bars) + P3_BAR_BYTES);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
    const int MTP = (MT + 1) / 2;
    const int total = MTP * NT;
    const int KB = (K + BK - 1) / BK;
    uint32_t* my_mask = masks + warp * KMASK_WORDS;
    const bool skip = epi.a_kflags != nullptr;
    const int kwords = (KB + 1) >> 1;
    auto pair_mask = [&](int mtp) {  // union of the K-block activity of the pair's two pixel tiles
        const unsigned int* r0 = epi.a_kflags + (size_t)(2 * mtp) * epi.a_kflag_ld;
        build_kmask2(r0, 2 * mtp + 1 < MT ? r0 + epi.a_kflag_ld : nullptr, kwords, KB, my_mask, lane);
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
        if (TMA_EPI) {
            tma_prefetch_desc(&map_e_in);
            tma_prefetch_desc(&map_e_hi);
            tma_prefetch_desc(&map_e_lo);
        }
        for (int s = 0; s < P3_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], 2 * EPI_WARPS);
        }
        for (int b = 0; b < 2 * EPI_WARPS; ++b) mbar_init(&bbar_all[b], 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc_pair(tmem_ptr, TMEM_COLS);
        tmem_relinquish_pair();
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp < 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
      if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int ti = cluster_id; ti < total; ti += nclusters) {
            int mtp, nt;
            tile_coords(ti, MTP, NT, mtp, nt);
            if (skip) pair_mask(mtp);
            if (lane == 0) {
                const int mt = 2 * mtp + (int)rank;
                for (int kb = skip ? next_active_kb(my_mask, 0, KB) : 0; kb < KB; kb = skip ? next_active_kb(my_mask, kb + 1, KB) : kb + 1) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sb = stage_base + stage * P3_STAGE_BYTES;
                    if (rank == 0) mbar_arrive_expect_tx(&full[stage], 2 * P3_STAGE_BYTES);
                    tma_load_2d_pair(sb, &map_a_hi, &full[stage], kb * BK, mt * BM);
                    tma_load_2d_pair(sb + A_TILE_BYTES, &map_a_lo, &full[stage], kb * BK, mt * BM);
                    tma_load_2d_pair(sb + 2 * A_TILE_BYTES, &map_b_hi, &full[stage], kb * BK, nt * BN + (int)rank * (BN / 2));
                    tma_load_2d_pair(sb + 3 * A_TILE_BYTES, &map_b_lo, &full[stage], kb * BK, nt * BN + (int)rank * (BN / 2));
                    if (++stage == P3_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            __syncwarp();  // the mask slot is rewritten for the next tile
        }
      } else if (warp == 1 && rank == 0) {
        // ===================== MMA issuer (leader CTA only) =====================
        constexpr uint32_t idesc = umma_idesc_f16(2 * BM, BN);  // M = 256 over the pair
        int stage = 0;
        uint32_t phase = 0;
        uint32_t chunk = 0;
        unsigned long long issued = 0;
        for (int ti = cluster_id; ti < total; ti += nclusters) {
            int mtp, nt;
            tile_coords(ti, MTP, NT, mtp, nt);
            if (skip) pair_mask(mtp);
            if (lane == 0) {
                int in_chunk = 0, cur_win = -1;
                uint32_t d_tmem = 0;
                for (int kb = skip ? next_active_kb(my_mask, 0, KB) : 0; kb < KB; kb = skip ? next_active_kb(my_mask, kb + 1, KB) : kb + 1) {
                    const int win = kb / chunk_kb;
                    if (in_chunk != 0 && win != cur_win) {
                        umma_commit_pair(&tmem_full[chunk & 1]);
                        ++chunk;
                        in_chunk = 0;
                    }
                    if (in_chunk == 0) {
                        const int buf = chunk & 1;
                        mbar_wait(&tmem_empty[buf], ((chunk >> 1) & 1) ^ 1);
                        tc_fence_after();
                        d_tmem = tmem_base + (uint32_t)(buf * BN);
                        cur_win = win;
                    }
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa_hi = smem_u32(stage_base + stage * P3_STAGE_BYTES);
                    const uint32_t sa_lo = sa_hi + A_TILE_BYTES, sb_hi = sa_hi + 2 * A_TILE_BYTES, sb_lo = sa_hi + 3 * A_TILE_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < BK / 16; ++k4) {
                        const uint64_t da_hi = umma_desc_sw128(sa_hi + k4 * 32), da_lo = umma_desc_sw128(sa_lo + k4 * 32);
                        const uint64_t db_hi = umma_desc_sw128(sb_hi + k4 * 32), db_lo = umma_desc_sw128(sb_lo + k4 * 32);
                        umma_f16_pair(d_tmem, da_hi, db_hi, idesc, (in_chunk != 0 || k4 != 0) ? 1u : 0u);
                        umma_f16_pair(d_tmem, da_lo, db_hi, idesc, 1u);
                        umma_f16_pair(d_tmem, da_hi, db_lo, idesc, 1u);
                    }
                    issued += 6;  // two CTAs x three products, units of one 128 x 256 x 64 block
                    umma_commit_pair(&empty[stage]);
                    if (++stage == P3_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                    ++in_chunk;
                }
                if (in_chunk != 0) {
                    umma_commit_pair(&tmem_full[chunk & 1]);
                    ++chunk;
                }
            }
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
        }
        if (lane == 0 && epi.mma_count != nullptr && issued != 0) atomicAdd(epi.mma_count + epi.count_slot, issued);
      }
    } else {
        // ===================== epilogue warps (both CTAs) =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int ew = warp - 4;
        const int sub = warp & 3;
        const int half = ew >> 2;
        uint8_t* bbuf = epi_base + ew * P3_EPI_WARP_BYTES;
        uint8_t* obuf = bbuf + 8192;
        uint64_t* bbar = bbar_all + 2 * ew;
        uint32_t bphase[2] = {0u, 0u};
        uint32_t chunk = 0;
        const bool resid_fast = MODE == EPI_RESID && (epi.m & 1) == 0 && epi.debug == 0 && epi.no_fast == 0 && (epi.ld_out & 1) == 0 &&
                                ((reinterpret_cast<uintptr_t>(epi.b) | reinterpret_cast<uintptr_t>(epi.chan_scale)) & 7) == 0 &&
                                ((reinterpret_cast<uintptr_t>(epi.out_hi) | reinterpret_cast<uintptr_t>(epi.out_lo)) & 3) == 0 &&
                                (epi.out8 == nullptr || ((reinterpret_cast<uintptr_t>(epi.out8) | (uintptr_t)epi.ld8) & 1) == 0);
        for (int ti = cluster_id; ti < total; ti += nclusters) {
            int mtp, nt;
            tile_coords(ti, MTP, NT, mtp, nt);
            const int mt = 2 * mtp + (int)rank;
            const bool exists = mt < MT;
            const int erow0 = mt * BM + sub * 32, ecol0 = nt * BN + half * EPI_COLS;
            const bool staged = TMA_EPI && MODE == EPI_RESID && exists && ecol0 < epi.N;
            if (staged && lane == 0) {  // the first two passes of b arrive while the K loop runs
                resid_tma_request(epi, &map_e_in, bbuf, bbar, 0, erow0, ecol0);
                resid_tma_request(epi, &map_e_in, bbuf, bbar, 1, erow0, ecol0);
            }
            float acc[EPI_COLS];
#pragma unroll
            for (int j = 0; j < EPI_COLS; ++j) acc[j] = 0.f;
            int nchunks = (KB + chunk_kb - 1) / chunk_kb;
            if (skip) {  // count the windows that hold an active K block, exactly as the MMA thread walks them
                pair_mask(mtp);
                int c = 0;
                if (lane == 0) {
                    int cur_win = -1;
                    for (int kb = next_active_kb(my_mask, 0, KB); kb < KB; kb = next_active_kb(my_mask, kb + 1, KB)) {
                        const int win = kb / chunk_kb;
                        c += (win != cur_win);
                        cur_win = win;
                    }
                }
                nchunks = __shfl_sync(0xffffffffu, c, 0);
            }
            for (int c = 0; c < nchunks; ++c, ++chunk) {
                const int buf = chunk & 1;
                mbar_wait(&tmem_full[buf], (chunk >> 1) & 1);
                tc_fence_after();
                if (exists) drain_chunk(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(buf * BN + half * EPI_COLS), acc);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (rank == 0) mbar_arrive(&tmem_empty[buf]);
                    else mbar_arrive_leader(&tmem_empty[buf]);
                }
            }
            if (!exists) continue;
            if (staged) {
                epilogue_resid_tma(epi, acc, lane, erow0, ecol0, &map_e_in, &map_e_hi, &map_e_lo, bbuf, obuf, bbar, bphase);
            } else if (MODE == EPI_RESID && resid_fast && ecol0 + EPI_COLS <= epi.N && (ecol0 + EPI_COLS <= epi.m || ecol0 >= epi.m)) {
                epilogue_resid_fast(epi, acc, lane, erow0, ecol0);
            } else {
                epilogue_tile<MODE>(epi, acc, lane, erow0, ecol0);
            }
        }
        if (TMA_EPI && lane == 0) tma_store_wait_read<0>();  // the staging tiles stay valid until the last stores have read them
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_pair(tmem_base, TMEM_COLS);
    }
}

// ---- the K-block-skipping forward transform on 128 x 128 tiles -------------------------------------------------------
// In the L1-sparse regime a forward tile multiplies ~9 of 249 K blocks: the kernel is bound by operand delivery and by its
// residual epilogue, not by the tensor pipe, and in dft_gemm_kernel the two do not overlap -- the epilogue warps are the
// ones that drain the 2-block accumulation chunks, TMEM holds two 128 x 256 accumulators, so the MMA side stalls after
// two chunks while the epilogue of the previous tile runs (ncu: the epilogue warps spend 28 % of their time waiting for
// chunks, 48 % in the epilogue).  With 128 output columns per tile TMEM holds FOUR chunk accumulators (the MMA side runs a
// whole tile ahead), a K block stages 64 KiB (three stages instead of two), and an epilogue thread holds 64 accumulators,
// which leaves registers for ALL of its b values: they are requested at the top of the tile and arrive while the K loop
// runs, instead of four serial batches after it.  Same products, same window-aligned chunks, same epilogue arithmetic:
// bit-identical to dft_gemm_kernel<EPI_RESID>.  (The operand traffic per output rises by a third -- the A tile is read once
// per 128 instead of 256 columns -- which is why the dense transforms stay on the 256-column kernels.)
constexpr int F1_BN = 128;
constexpr int F1_STAGES = 3;
constexpr int F1_STAGE_BYTES = 4 * A_TILE_BYTES;  // A_hi, A_lo, B_hi, B_lo: 4 x (128 rows x 128 B)
constexpr int F1_EPW = 16;                        // epilogue warps: 4 TMEM lane quarters x 4 column quarters
constexpr int F1_COLS = F1_BN / 4;                // accumulator columns per epilogue thread
constexpr int F1_THREADS = 128 + 32 * F1_EPW;
constexpr int F1_WARPS = 4 + F1_EPW;
constexpr int F1_SMEM_BYTES = 1024 + F1_STAGES * F1_STAGE_BYTES + 256 + F1_WARPS * KMASK_WORDS * 4;
static_assert(F1_SMEM_BYTES <= 232448, "shared memory budget exceeded");

// (an epilogue thread of dft_gemm_kernel issues ~60 instructions per element with two warps per scheduler to hide their
//  latencies; sixteen epilogue warps -- four per scheduler -- on 32-column fragments halve the time of the same code)
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void drain_chunk32(uint32_t taddr, float (&acc)[F1_COLS]) {
#pragma unroll
    for (int lh = 0; lh < 2; ++lh) {
        float v[16];
        tmem_ld_16x256b_x4(taddr + ((uint32_t)(16 * lh) << 16), v);
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[lh * 16 + j] += v[j];
    }
}

// b values of this thread's 4 rows x 4 column pairs (requested before the K loop ends; zero where out of range)
template <bool FAST>
__device__ __forceinline__ void resid32_request(const GemmEpilogue& epi, int lane, int row0, int col0, float2 (&bv)[16]) {
    const int quad = lane & 3;
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) {  // ri = 2 lh + rh
        const int row = row0 + (lane >> 2) + 8 * (ri & 1) + 16 * (ri >> 1);
        const float2* bp = reinterpret_cast<const float2*>(epi.b + (size_t)row * epi.ld_out + col0 + 2 * quad);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const bool ok = row < epi.P && (FAST || col0 + 8 * jj + 2 * quad < epi.N);
            bv[ri * 4 + jj] = ok ? __ldcg(bp + 4 * jj) : make_float2(0.f, 0.f);
        }
    }
}

// EPI_RESID on a 32-row x 32-column fragment; same arithmetic, same bits as epilogue_tile<EPI_RESID>.  FAST: the 32
// columns lie inside the matrix and inside one half of the row ([0, m) or [m, 2m)), m even, 8-byte aligned views.
template <bool FAST>
__device__ __forceinline__ void epilogue_resid32(const GemmEpilogue& epi, const float (&acc)[F1_COLS], const float2 (&bv)[16], int lane,
                                                 int row0, int col0) {
    const int quad = lane & 3;
    const int cbase = col0 + 2 * quad;
    const int chan0 = cbase >= epi.m ? cbase - epi.m : cbase;
    const int c8 = cbase >= epi.m ? cbase - epi.m + (epi.ld8 >> 1) : cbase;
    const bool cs_even = (epi.m & 1) == 0 && (reinterpret_cast<uintptr_t>(epi.chan_scale) & 7) == 0;
#pragma unroll
    for (int ri = 0; ri < 4; ++ri) {
        const int lh = ri >> 1, rh = ri & 1;
        const int row = row0 + (lane >> 2) + 8 * rh + 16 * lh;
        const bool row_ok = row < epi.P;
        float dsum = 0.f, rmax = 0.f;
        if (row_ok) {
            const float descale = kTwiddleDescale / __ldg(epi.a_scale + row);
            const float oscale = __ldg(epi.out_scale + row);
            const size_t row_off = (size_t)row * epi.ld_out;
            const float* cs_row = epi.chan_scale ? epi.chan_scale + (epi.chan_per_pixel ? (size_t)row * epi.m : 0) : nullptr;
            __half2* hp = reinterpret_cast<__half2*>(epi.out_hi + row_off + cbase);
            __half2* lp = reinterpret_cast<__half2*>(epi.out_lo + row_off + cbase);
            float2 cs[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                cs[jj] = make_float2(1.f, 1.f);
                if (FAST) {
                    if (cs_row != nullptr) cs[jj] = __ldg(reinterpret_cast<const float2*>(cs_row + chan0) + 4 * jj);
                } else {
                    const int col = cbase + 8 * jj;
                    if (cs_row != nullptr && col < epi.N) {
                        const int c0 = col >= epi.m ? col - epi.m : col;
                        if (cs_even) {
                            cs[jj] = __ldg(reinterpret_cast<const float2*>(cs_row + c0));
                        } else {
                            const int c1 = col + 1 >= epi.m ? col + 1 - epi.m : col + 1;
                            cs[jj] = make_float2(__ldg(cs_row + c0), __ldg(cs_row + c1));
                        }
                    }
                }
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int col = cbase + 8 * jj;
                if (!FAST && col >= epi.N) continue;
                const int ai = lh * 16 + (jj * 2 + rh) * 2;
                const float a0 = acc[ai] * descale, a1 = acc[ai + 1] * descale;
                const float2 b = bv[ri * 4 + jj];
                const float r0 = __fmaf_rn(cs[jj].x, a0, -b.x) * oscale, r1 = __fmaf_rn(cs[jj].y, a1, -b.y) * oscale;
                const __half2 h = __floats2half2_rn(r0, r1);
                const float2 hf = __half22float2(h);
                const __half2 l = __floats2half2_rn(r0 - hf.x, r1 - hf.y);
                hp[4 * jj] = h;
                lp[4 * jj] = l;
                dsum += fabsf(hf.x) + fabsf(hf.y);
                rmax = fmaxf(rmax, fmaxf(fabsf(r0), fabsf(r1)));
                if (epi.out8 != nullptr) {
                    const float2 v8 = make_float2(hf.x * kOperandScale8, hf.y * kOperandScale8);
                    if (fmaxf(fabsf(v8.x), fabsf(v8.y)) > 448.f) dsum = __int_as_float(0x7f800000);  // unscreenable row
                    const __nv_fp8x2_storage_t pk = __nv_cvt_float2_to_fp8x2(v8, __NV_SATFINITE, __NV_E4M3);
                    unsigned char* o8 = epi.out8 + (size_t)row * epi.ld8;
                    if (FAST) {
                        reinterpret_cast<__nv_fp8x2_storage_t*>(o8 + c8)[4 * jj] = pk;
                    } else {
                        const int h8 = epi.ld8 >> 1;
                        if (epi.m & 1) {
                            o8[col < epi.m ? col : col - epi.m + h8] = (unsigned char)(pk & 0xff);
                            o8[col + 1 < epi.m ? col + 1 : col + 1 - epi.m + h8] = (unsigned char)(pk >> 8);
                        } else {
                            *reinterpret_cast<__nv_fp8x2_storage_t*>(o8 + (col < epi.m ? col : col - epi.m + h8)) = pk;
                        }
                    }
                }
            }
        }
        if (epi.abs_sum_out != nullptr) {  // kernel-uniform
            dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
            dsum += __shfl_xor_sync(0xffffffffu, dsum, 2);
            if (quad == 0 && row_ok && dsum != 0.f) atomicAdd(epi.abs_sum_out + row, dsum);
        }
        if (epi.row_stat != nullptr) {  // kernel-uniform
            rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 1));
            rmax = fmaxf(rmax, __shfl_xor_sync(0xffffffffu, rmax, 2));
            if (quad == 0 && row_ok && rmax != 0.f)
                atomicMax(reinterpret_cast<unsigned int*>(epi.row_stat + 4 * (size_t)row + 3), __float_as_uint(rmax));
        }
    }
}

__global__ void __launch_bounds__(F1_THREADS, 1)
dft_fwd128_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                  const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, int K, int MT, int NT,
                  int chunk_kb, const GemmEpilogue epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F1_STAGES * F1_STAGE_BYTES);
    uint64_t* full = bars;                    // [F1_STAGES]
    uint64_t* empty = bars + 4;               // [F1_STAGES]
    uint64_t* tmem_full = bars + 8;           // [4]  one per chunk accumulator
    uint64_t* tmem_empty = bars + 12;         // [4]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = MT * NT;
    const int KB = (K + BK - 1) / BK;
    uint32_t* my_mask = reinterpret_cast<uint32_t*>(smem + F1_STAGES * F1_STAGE_BYTES + 256) + warp * KMASK_WORDS;
    const bool skip = epi.a_kflags != nullptr;
    const int kwords = (KB + 1) >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
        for (int s = 0; s < F1_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 4; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], F1_EPW);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // registers: 640 threads start with 96 each (61440 for the CTA); the four control warps drop to 40 and free 7168, the
    // sixteen epilogue warps rise to 104 and take 4096 of them (setmaxnreg only moves registers inside the CTA's allocation:
    // asking for more than the others gave up would wait forever)
    if (warp < 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
      if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (int ti = blockIdx.x; ti < total_tiles; ti += gridDim.x) {
            int mt, nt;
            tile_coords(ti, MT, NT, mt, nt);
            if (skip) build_kmask(epi.a_kflags + (size_t)mt * epi.a_kflag_ld, kwords, KB, my_mask, lane);
            if (lane == 0) {
                for (int kb = skip ? next_active_kb(my_mask, 0, KB) : 0; kb < KB; kb = skip ? next_active_kb(my_mask, kb + 1, KB) : kb + 1) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sb = stage_base + stage * F1_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[stage], F1_STAGE_BYTES);
                    tma_load_2d(sb, &map_a_hi, &full[stage], kb * BK, mt * BM);
                    tma_load_2d(sb + A_TILE_BYTES, &map_a_lo, &full[stage], kb * BK, mt * BM);
                    tma_load_2d(sb + 2 * A_TILE_BYTES, &map_b_hi, &full[stage], kb * BK, nt * F1_BN);
                    tma_load_2d(sb + 3 * A_TILE_BYTES, &map_b_lo, &full[stage], kb * BK, nt * F1_BN);
                    if (++stage == F1_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            __syncwarp();
        }
      } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_f16(BM, F1_BN);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t chunk = 0;  // running chunk counter -> TMEM buffer = chunk & 3, parity = (chunk >> 2) & 1
        unsigned long long issued = 0;
        for (int ti = blockIdx.x; ti < total_tiles; ti += gridDim.x) {
            if (skip) {
                int mt, nt;
                tile_coords(ti, MT, NT, mt, nt);
                build_kmask(epi.a_kflags + (size_t)mt * epi.a_kflag_ld, kwords, KB, my_mask, lane);
            }
            if (lane == 0) {
                int in_chunk = 0, cur_win = -1;
                uint32_t d_tmem = 0;
                for (int kb = skip ? next_active_kb(my_mask, 0, KB) : 0; kb < KB; kb = skip ? next_active_kb(my_mask, kb + 1, KB) : kb + 1) {
                    const int win = kb / chunk_kb;
                    if (in_chunk != 0 && win != cur_win) {
                        umma_commit(&tmem_full[chunk & 3]);
                        ++chunk;
                        in_chunk = 0;
                    }
                    if (in_chunk == 0) {
                        const int buf = chunk & 3;
                        mbar_wait(&tmem_empty[buf], ((chunk >> 2) & 1) ^ 1);
                        tc_fence_after();
                        d_tmem = tmem_base + (uint32_t)(buf * F1_BN);
                        cur_win = win;
                    }
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa_hi = smem_u32(stage_base + stage * F1_STAGE_BYTES);
                    const uint32_t sa_lo = sa_hi + A_TILE_BYTES, sb_hi = sa_hi + 2 * A_TILE_BYTES, sb_lo = sa_hi + 3 * A_TILE_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < BK / 16; ++k4) {
                        const uint64_t da_hi = umma_desc_sw128(sa_hi + k4 * 32), da_lo = umma_desc_sw128(sa_lo + k4 * 32);
                        const uint64_t db_hi = umma_desc_sw128(sb_hi + k4 * 32), db_lo = umma_desc_sw128(sb_lo + k4 * 32);
                        umma_f16(d_tmem, da_hi, db_hi, idesc, (in_chunk != 0 || k4 != 0) ? 1u : 0u);
                        umma_f16(d_tmem, da_lo, db_hi, idesc, 1u);
                        umma_f16(d_tmem, da_hi, db_lo, idesc, 1u);
                    }
                    issued += 3;  // (half units of 128 x 256 x 64)
                    umma_commit(&empty[stage]);
                    if (++stage == F1_STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                    ++in_chunk;
                }
                if (in_chunk != 0) {
                    umma_commit(&tmem_full[chunk & 3]);
                    ++chunk;
                }
            }
            chunk = __shfl_sync(0xffffffffu, chunk, 0);
        }
        if (lane == 0 && epi.mma_count != nullptr && issued != 0) atomicAdd(epi.mma_count + epi.count_slot, (issued + 1) / 2);
      }
    } else {
        // ===================== epilogue warps =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        const int ew = warp - 4;
        const int sub = warp & 3;        // TMEM lane quarter this warp may read (hardware rule: warp id mod 4)
        const int cq = ew >> 2;          // which 32 of the tile's 128 columns
        const bool fast_ok = (epi.m & 1) == 0 && (epi.ld_out & 1) == 0 &&
                             ((reinterpret_cast<uintptr_t>(epi.b) | reinterpret_cast<uintptr_t>(epi.chan_scale)) & 7) == 0 &&
                             ((reinterpret_cast<uintptr_t>(epi.out_hi) | reinterpret_cast<uintptr_t>(epi.out_lo)) & 3) == 0 &&
                             (epi.out8 == nullptr || ((reinterpret_cast<uintptr_t>(epi.out8) | (uintptr_t)epi.ld8) & 1) == 0) &&
                             epi.no_fast == 0;
        uint32_t chunk = 0;
        for (int ti = blockIdx.x; ti < total_tiles; ti += gridDim.x) {
            int mt, nt;
            tile_coords(ti, MT, NT, mt, nt);
            const int erow0 = mt * BM + sub * 32, ecol0 = nt * F1_BN + cq * F1_COLS;
            const bool fast = fast_ok && ecol0 + F1_COLS <= epi.N && (ecol0 + F1_COLS <= epi.m || ecol0 >= epi.m);  // (warp-uniform)
            float2 bv[16];
            if (fast) resid32_request<true>(epi, lane, erow0, ecol0, bv);  // in flight while the K loop runs
            else resid32_request<false>(epi, lane, erow0, ecol0, bv);
            float acc[F1_COLS];
#pragma unroll
            for (int j = 0; j < F1_COLS; ++j) acc[j] = 0.f;
            int nchunks = (KB + chunk_kb - 1) / chunk_kb;
            if (skip) {
                build_kmask(epi.a_kflags + (size_t)mt * epi.a_kflag_ld, kwords, KB, my_mask, lane);
                int c = 0;
                if (lane == 0) {
                    int cur_win = -1;
                    for (int kb = next_active_kb(my_mask, 0, KB); kb < KB; kb = next_active_kb(my_mask, kb + 1, KB)) {
                        const int win = kb / chunk_kb;
                        c += (win != cur_win);
                        cur_win = win;
                    }
                }
                nchunks = __shfl_sync(0xffffffffu, c, 0);
            }
            for (int c = 0; c < nchunks; ++c, ++chunk) {
                const int buf = chunk & 3;
                mbar_wait(&tmem_full[buf], (chunk >> 2) & 1);
                tc_fence_after();
                drain_chunk32(tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(buf * F1_BN + cq * F1_COLS), acc);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
            }
            if (fast) epilogue_resid32<true>(epi, acc, bv, lane, erow0, ecol0);
            else epilogue_resid32<false>(epi, acc, bv, lane, erow0, ecol0);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}


// ---- the direct residual epilogue on sixteen warps ---------------------------------------------------------------------
// Sparse forward tiles (one accumulation chunk, GemmEpilogue::group_sparse) spend their time in the epilogue: with eight
// epilogue warps -- two per scheduler -- the 2.7 k instructions a warp issues per tile run at a fifth of the issue rate (time
// line: 16 k clocks of body, 10 k of requests against an MMA span of 19 k).  dft_fwd_direct_kernel processes ONLY those tiles
// (single_rt[mt] = 1, written by tile_single_kernel from the K-block flags; dft_gemm_kernel<EPI_RESID> skips them) with sixteen
// epilogue warps: a warp owns 32 rows x 64 columns, reads the accumulator straight from TMEM in two 16-row pieces, transposes
// 4-row x 64-column sub-pieces through a 1.1 KB shared-memory buffer so that its global accesses are row-contiguous (16-byte
// loads of b, 8-byte stores of r_hi / r_lo, 4-byte stores of the fp8 copy), and keeps four sub-pieces of b in flight.  Same
// element arithmetic as resid_direct_rows: same bits.  Needs m % 4 == 0 and channel factors shared by all rows (or none).
constexpr int D16_EPW = 16;
constexpr int D16_THREADS = 128 + 32 * D16_EPW;
constexpr int D16_XP_ROW = 64 + 8;                         // floats; rows padded by 32 bytes
constexpr int D16_XP_BYTES = 4 * D16_XP_ROW * 4;           // 4 rows per sub-piece
constexpr int D16_SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 256 + 2 * KMASK_WORDS * 4 + D16_EPW * D16_XP_BYTES;
static_assert(D16_SMEM_BYTES <= 232448, "shared memory budget exceeded");
// single_rt[mt] = 1: the pixel tile has at most group_max active K blocks (and at least one when K blocks are skipped)
__global__ void tile_single_kernel(const unsigned int* __restrict__ flags, int flag_ld, int kwords, int KB, int MT, int group_max, int skip,
                                   unsigned char* __restrict__ single_rt) {
    const int mt = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (mt >= MT) return;
    const unsigned int* row = flags + (size_t)mt * flag_ld;
    int count = 0;
    unsigned int last = 0;
    for (int w0 = 0; w0 < kwords; w0 += 32) {
        const int idx = w0 + lane;
        const unsigned int v = idx < kwords ? __ldcg(row + idx) : 0u;
        if (idx == kwords - 1) last = v;
        count += 2 * __popc(__ballot_sync(0xffffffffu, v != 0u));
    }
    last = __shfl_sync(0xffffffffu, last, (kwords - 1) & 31);
    if ((KB & 1) && last != 0u) count -= 1;   // (the last flag word covers a single K block when KB is odd: build_kmask)
    if (lane == 0) {
        const bool single = count <= group_max && !(skip && count == 0);
        single_rt[mt] = (unsigned char)single;
        if (!single) atomicAdd(reinterpret_cast<int*>(single_rt + single_rt_counter_offset(MT)), 1);
    }
}

template <bool CS_PP>   // channel factors of the pixel's own ([P, m], requested like b) instead of shared ones
__global__ void __launch_bounds__(D16_THREADS, 1)
dft_fwd_direct_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                      const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo, int K, int MT, int NT,
                      const GemmEpilogue epi) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* stage_base = smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
    uint64_t* full = bars;            // [STAGES]
    uint64_t* empty = bars + 4;       // [STAGES]
    uint64_t* tmem_full = bars + 8;   // [2]
    uint64_t* tmem_empty = bars + 10; // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);
    uint32_t* masks = reinterpret_cast<uint32_t*>(smem + STAGES * STAGE_BYTES + 256);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = MT * NT;
    const int KB = (K + BK - 1) / BK;
    const bool skip = epi.no_kskip == 0;
    const int kwords = (KB + 1) >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_a_lo);
        tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_b_lo);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tmem_full[b], 1);
            mbar_init(&tmem_empty[b], D16_EPW);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    // registers: 640 threads start with 96 each; the four control warps drop to 40 and free 7168, the sixteen epilogue warps rise
    // to 104 and take 4096 of them (setmaxnreg only moves registers inside the CTA's allocation)
    if (warp < 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
      if (warp == 0) {
        // ===================== TMA producer =====================
        uint32_t* my_mask = masks;
        int stage = 0;
        uint32_t phase = 0;
        for (int ti = blockIdx.x; ti < total_tiles; ti += gridDim.x) {
            int mt, nt;
            tile_coords(ti, MT, NT, mt, nt);
            if (!epi.single_rt[mt]) continue;
            if (skip) build_kmask(epi.a_kflags + (size_t)mt * epi.a_kflag_ld, kwords, KB, my_mask, lane);
            if (lane == 0) {
                for (int kb = skip ? next_active_kb(my_mask, 0, KB) : 0; kb < KB; kb = skip ? next_active_kb(my_mask, kb + 1, KB) : kb + 1) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    uint8_t* sb = stage_base + stage * STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
                    tma_load_2d(sb, &map_a_hi, &full[stage], kb * BK, mt * BM);
                    tma_load_2d(sb + A_TILE_BYTES, &map_a_lo, &full[stage], kb * BK, mt * BM);
                    tma_load_2d(sb + 2 * A_TILE_BYTES, &map_b_hi, &full[stage], kb * BK, nt * BN);
                    tma_load_2d(sb + 2 * A_TILE_BYTES + B_TILE_BYTES, &map_b_lo, &full[stage], kb * BK, nt * BN);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
            __syncwarp();
        }
      } else if (warp == 1) {
        // ===================== MMA issuer: one accumulation chunk per tile =====================
        uint32_t* my_mask = masks + KMASK_WORDS;
        constexpr uint32_t idesc = umma_idesc_f16(BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        uint32_t tile_no = 0;  // -> TMEM buffer = tile_no & 1, parity = (tile_no >> 1) & 1
        unsigned long long issued = 0;
        for (int ti = blockIdx.x; ti < total_tiles; ti += gridDim.x) {
            int mt, nt;
            tile_coords(ti, MT, NT, mt, nt);
            if (!epi.single_rt[mt]) continue;
            if (skip) build_kmask(epi.a_kflags + (size_t)mt * epi.a_kflag_ld, kwords, KB, my_mask, lane);
            if (lane == 0) {
                const int buf = tile_no & 1;
                mbar_wait(&tmem_empty[buf], ((tile_no >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
                int done = 0;
                for (int kb = skip ? next_active_kb(my_mask, 0, KB) : 0; kb < KB; kb = skip ? next_active_kb(my_mask, kb + 1, KB) : kb + 1) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa_hi = smem_u32(stage_base + stage * STAGE_BYTES);
                    const uint32_t sa_lo = sa_hi + A_TILE_BYTES, sb_hi = sa_hi + 2 * A_TILE_BYTES, sb_lo = sb_hi + B_TILE_BYTES;
#pragma unroll
                    for (int k4 = 0; k4 < BK / 16; ++k4) {
                        const uint64_t da_hi = umma_desc_sw128(sa_hi + k4 * 32), da_lo = umma_desc_sw128(sa_lo + k4 * 32);
                        const uint64_t db_hi = umma_desc_sw128(sb_hi + k4 * 32), db_lo = umma_desc_sw128(sb_lo + k4 * 32);
                        umma_f16(d_tmem, da_hi, db_hi, idesc, (done != 0 || k4 != 0) ? 1u : 0u);
                        umma_f16(d_tmem, da_lo, db_hi, idesc, 1u);
                        umma_f16(d_tmem, da_hi, db_lo, idesc, 1u);
                    }
                    issued += 3;
                    umma_commit(&empty[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                    ++done;
                }
                umma_commit(&tmem_full[buf]);
            }
            ++tile_no;
        }
        if (lane == 0 && epi.mma_count != nullptr && issued != 0) atomicAdd(epi.mma_count + epi.count_slot, issued);
      }
    } else {
        // ===================== epilogue warps =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        const int ew = warp - 4;
        const int sub = warp & 3;        // TMEM lane quarter this warp may read (hardware rule: warp id mod 4)
        const int cq = ew >> 2;          // which 64 of the tile's 256 columns
        float* xp = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES + 256 + 2 * KMASK_WORDS * 4) + ew * (D16_XP_BYTES / 4);
        const int lrow = lane >> 4;      // which of the two rows of an access
        const int lcol = 4 * (lane & 15);
        const int h8 = epi.ld8 >> 1;
        uint32_t tile_no = 0;
        for (int ti = blockIdx.x; ti < total_tiles; ti += gridDim.x) {
            int mt, nt;
            tile_coords(ti, MT, NT, mt, nt);
            if (!epi.single_rt[mt]) continue;
            const int buf = tile_no & 1;
            const uint32_t par = (tile_no >> 1) & 1;
            ++tile_no;
            const int row0 = mt * BM + sub * 32, col = nt * BN + cq * 64 + lcol;
            const bool col_ok = col < epi.N;   // (N = 2 m, m % 4 == 0: a lane's four columns are inside or outside together)
            if (nt * BN + cq * 64 >= epi.N) {  // the warp's columns lie beyond the matrix: release the accumulator and move on
                mbar_wait(&tmem_full[buf], par);
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                continue;
            }
            const int chan = col >= epi.m ? col - epi.m : col;
            const int col8 = col < epi.m ? col : col - epi.m + h8;
            // row scalars of the warp's 32 rows: lane l holds row l's
            float my_desc = 0.f, my_osc = 0.f;
            if (row0 + lane < epi.P) {
                my_desc = kTwiddleDescale / __ldg(epi.a_scale + row0 + lane);
                my_osc = __ldg(epi.out_scale + row0 + lane);
            }
            const float4 cf_shared = (!CS_PP && epi.chan_scale != nullptr && col_ok) ? __ldg(reinterpret_cast<const float4*>(epi.chan_scale + chan))
                                                                                     : make_float4(1.f, 1.f, 1.f, 1.f);
            // b (and the pixel's own channel factors): sub-step s = (lh, rh, hs) covers rows 16 lh + 8 rh + 4 hs + lrow + 2 i
            // (i = 0, 1); DEPTH sub-steps in flight -- four of b alone, two of b and channel factors (the same registers)
            constexpr int DEPTH = CS_PP ? 2 : 4;
            float4 bq[DEPTH][2], cq_[CS_PP ? DEPTH : 1][2];
            auto request_b = [&](int s, int slot) {
                const int base = 16 * (s >> 2) + 8 * ((s >> 1) & 1) + 4 * (s & 1) + lrow;
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int row = row0 + base + 2 * i;
                    const bool ok = row < epi.P && col_ok;
                    bq[slot][i] = ok ? __ldcg(reinterpret_cast<const float4*>(epi.b + (size_t)row * epi.ld_out + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    if (CS_PP)
                        cq_[slot][i] = ok ? __ldcg(reinterpret_cast<const float4*>(epi.chan_scale + (size_t)row * epi.m + chan))
                                          : make_float4(1.f, 1.f, 1.f, 1.f);
                }
            };
#pragma unroll
            for (int s = 0; s < DEPTH; ++s) request_b(s, s);
            mbar_wait(&tmem_full[buf], par);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(buf * BN + cq * 64);
#pragma unroll
            for (int lh = 0; lh < 2; ++lh) {
                float v[32];
                tmem_ld_16x256b_x8(taddr + ((uint32_t)(16 * lh) << 16), v);
                if (lh == 1) {   // the accumulator has been read: the MMA thread may reuse the buffer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&tmem_empty[buf]);
                }
#pragma unroll
                for (int rh = 0; rh < 2; ++rh) {
#pragma unroll
                    for (int hs = 0; hs < 2; ++hs) {
                        const int s = (lh * 2 + rh) * 2 + hs;
                        // fragment -> rows: thread T holds (row T/4, columns 8 j + 2 (T%4), + 1); rows 4 hs .. 4 hs + 3 of this pass
                        __syncwarp();
                        if ((lane >> 4) == hs) {
#pragma unroll
                            for (int j8 = 0; j8 < 8; ++j8)
                                *reinterpret_cast<float2*>(xp + ((lane >> 2) & 3) * D16_XP_ROW + 8 * j8 + 2 * (lane & 3)) =
                                    make_float2(v[(j8 * 2 + rh) * 2], v[(j8 * 2 + rh) * 2 + 1]);
                        }
                        __syncwarp();
                        float dsum[2] = {0.f, 0.f}, rmax[2] = {0.f, 0.f};
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int lr = 16 * lh + 8 * rh + 4 * hs + lrow + 2 * i;   // row inside the warp's 32
                            const int row = row0 + lr;
                            const float4 a = *reinterpret_cast<const float4*>(xp + (lrow + 2 * i) * D16_XP_ROW + lcol);
                            const float desc = __shfl_sync(0xffffffffu, my_desc, lr), osc = __shfl_sync(0xffffffffu, my_osc, lr);
                            if (row >= epi.P || !col_ok) continue;
                            const float4 b = bq[s % DEPTH][i];
                            const float4 cf = CS_PP ? cq_[CS_PP ? s % DEPTH : 0][i] : cf_shared;
                            const float r0 = __fmaf_rn(cf.x, a.x * desc, -b.x) * osc, r1 = __fmaf_rn(cf.y, a.y * desc, -b.y) * osc;
                            const float r2 = __fmaf_rn(cf.z, a.z * desc, -b.z) * osc, r3 = __fmaf_rn(cf.w, a.w * desc, -b.w) * osc;
                            const __half2 h01 = __floats2half2_rn(r0, r1), h23 = __floats2half2_rn(r2, r3);
                            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                            const __half2 l01 = __floats2half2_rn(r0 - f01.x, r1 - f01.y), l23 = __floats2half2_rn(r2 - f23.x, r3 - f23.y);
                            const size_t off = (size_t)row * epi.ld_out + col;
                            uint2 hv, lv;
                            hv.x = *reinterpret_cast<const unsigned int*>(&h01);
                            hv.y = *reinterpret_cast<const unsigned int*>(&h23);
                            lv.x = *reinterpret_cast<const unsigned int*>(&l01);
                            lv.y = *reinterpret_cast<const unsigned int*>(&l23);
                            *reinterpret_cast<uint2*>(epi.out_hi + off) = hv;
                            *reinterpret_cast<uint2*>(epi.out_lo + off) = lv;
                            dsum[i] = (fabsf(f01.x) + fabsf(f01.y)) + (fabsf(f23.x) + fabsf(f23.y));
                            rmax[i] = fmaxf(fmaxf(fabsf(r0), fabsf(r1)), fmaxf(fabsf(r2), fabsf(r3)));
                            if (epi.out8 != nullptr) {  // kernel-uniform
                                const float2 v01 = make_float2(f01.x * kOperandScale8, f01.y * kOperandScale8);
                                const float2 v23 = make_float2(f23.x * kOperandScale8, f23.y * kOperandScale8);
                                if (fmaxf(fmaxf(fabsf(v01.x), fabsf(v01.y)), fmaxf(fabsf(v23.x), fabsf(v23.y))) > 448.f)
                                    dsum[i] = __int_as_float(0x7f800000);  // unscreenable row
                                const unsigned int p01 = __nv_cvt_float2_to_fp8x2(v01, __NV_SATFINITE, __NV_E4M3);
                                const unsigned int p23 = __nv_cvt_float2_to_fp8x2(v23, __NV_SATFINITE, __NV_E4M3);
                                *reinterpret_cast<unsigned int*>(epi.out8 + (size_t)row * epi.ld8 + col8) = p01 | (p23 << 16);
                            }
                        }
                        if (s + DEPTH < 8) request_b(s + DEPTH, s % DEPTH);   // the slot is free: the sub-step DEPTH ahead
                        // row totals of this sub-step's two rows per half warp: the 16 lanes that share a row
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            const int row = row0 + 16 * lh + 8 * rh + 4 * hs + lrow + 2 * i;
                            float d = dsum[i], r = rmax[i];
#pragma unroll
                            for (int sh = 1; sh < 16; sh <<= 1) {
                                d += __shfl_xor_sync(0xffffffffu, d, sh);
                                r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, sh));
                            }
                            if ((lane & 15) == 0 && row < epi.P) {
                                if (epi.abs_sum_out != nullptr && d != 0.f) atomicAdd(epi.abs_sum_out + row, d);
                                if (epi.row_stat != nullptr && r != 0.f)
                                    atomicMax(reinterpret_cast<unsigned int*>(epi.row_stat + 4 * (size_t)row + 3), __float_as_uint(r));
                            }
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---- host side --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp16 row-major [rows, cols] matrix with pitch ld -> 2-D map with a {64 x box_rows} box, 128B swizzle.
// Out-of-bounds parts of a box (K tail, ragged last pixel tile) read as zero.
int make_operand_map(CUtensorMap* map, const void* base, int rows, int cols, int ld, int box_rows, int elem_bytes) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        return CSR_ERR_CUDA;
    }
    CSR_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld * elem_bytes) % 16 == 0,
                "operand matrix must be 16-byte aligned with a pitch that is a multiple of 16 bytes (ld=%d)", ld);
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)(BK * 2 / elem_bytes), (cuuint32_t)box_rows};  // one 128-byte swizzle row of K
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                     const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%d cols=%d ld=%d)", (int)r, rows, cols, ld);
        return CSR_ERR_CUDA;
    }
    return CSR_OK;
}

// K blocks accumulated inside the tensor core before the partial sum is moved to registers (accuracy / speed knob,
// measured on B200 at C3 sampling, error = max-norm vs float64):
//   forward (K = 2n):  chunk 1, 2, 4 all run at the same speed; error 7.3e-7 / 7.0e-7 / 1.07e-6      -> 2 (CSR_CHUNK_KB)
//   adjoint (K = 2m):  chunk 1 / 2 / 4 / 8 = 2.97 / 2.85 / 2.63 / 2.51 ms, error 4.7e-7 / 7.8e-7 / 1.8e-6 / 3.5e-6
//                      (short K loop: every chunk boundary costs a TMEM hand-shake)                   -> 4 (CSR_CHUNK_KB_ADJ)
static int gemm_chunk_kb(bool adjoint, bool fista = false) {
    static int v[3] = {0, 0, 0};
    if (v[0] == 0) {
        const char* e = getenv("CSR_CHUNK_KB");
        const char* ea = getenv("CSR_CHUNK_KB_ADJ");
        const char* ef = getenv("CSR_CHUNK_KB_FISTA");
        v[0] = e ? atoi(e) : 2;
        v[1] = ea ? atoi(ea) : (e ? atoi(e) : 4);
        // the in-loop adjoint with the FISTA epilogue: K = 2m is 32 K blocks at m = 1000 -- chunks of 16 make that TWO accumulators,
        // which is what TMEM holds, so the MMA thread never waits for a drain (measured: 0.36 -> 0.31 ms on the listed tiles of
        // C3; error of the solution against the fp64 oracle 2.1e-7 -> 5.2e-7 of its scale, chunks of 32: 1.3e-6).  The plain
        // transforms (csr_adjoint, dirty spectrum, wavelet-basis gradient) keep chunks of 4.
        v[2] = ef ? atoi(ef) : (ea ? atoi(ea) : 16);
        for (int i = 0; i < 3; ++i) v[i] = v[i] < 1 ? 1 : (v[i] > 4096 ? 4096 : v[i]);
    }
    return adjoint ? (fista ? v[2] : v[1]) : v[0];
}

// ---- optional per-launch timing (bench.py's roofline leg): CUDA events on the launching stream ----------
struct ProfiledLaunch {
    int kind;  // mode * 2 + adjoint
    cudaEvent_t start, stop;
};
static bool g_profile = false;
static std::vector<ProfiledLaunch> g_profiled;
static std::vector<cudaEvent_t> g_event_pool;  // events are reused: creating two per launch costs host time

static cudaError_t pooled_event(cudaEvent_t* ev) {
    if (!g_event_pool.empty()) {
        *ev = g_event_pool.back();
        g_event_pool.pop_back();
        return cudaSuccess;
    }
    return cudaEventCreate(ev);
}

bool profile_enabled() { return g_profile; }
// used by the wavelet launchers (kinds 6 = fused synthesis, 7 = fused analysis + update)
bool profile_start(cudaStream_t stream, cudaEvent_t* start, cudaEvent_t* stop) {
    if (!g_profile) return false;
    if (pooled_event(start) != cudaSuccess || pooled_event(stop) != cudaSuccess) return false;
    cudaEventRecord(*start, stream);
    return true;
}
void profile_stop(int kind, cudaStream_t stream, cudaEvent_t start, cudaEvent_t stop) {
    cudaEventRecord(stop, stream);
    ProfiledLaunch rec;
    rec.kind = kind;
    rec.start = start;
    rec.stop = stop;
    g_profiled.push_back(rec);
}

// fp32 row-major [rows, cols] matrix -> map with a {256 x 128} box used only for L2 prefetch of epilogue tiles
static int make_prefetch_map(CUtensorMap* map, const void* base, int rows, int cols, int ld) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
        return CSR_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)BN, (cuuint32_t)BM};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (prefetch map) failed with CUresult %d", (int)r);
        return CSR_ERR_CUDA;
    }
    return CSR_OK;
}

template <int MODE>
static int launch_mode(const csr_plan* plan, bool adjoint, const CUtensorMap& a_hi, const CUtensorMap& a_lo, int P,
                       const GemmEpilogue& epi_in, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        CSR_CUDA(cudaFuncSetAttribute(dft_gemm_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        configured = true;
    }
    const int K = MODE == EPI_SCREEN8 ? plan->ld8 : (adjoint ? 2 * plan->m : 2 * plan->n);  // (fp8 rows are padded)
    const int N = adjoint ? 2 * plan->n : 2 * plan->m;
    const int MT = (P + BM - 1) / BM;
    const int NT = (N + BN - 1) / BN;
    const int grid = min(MT * NT, plan->num_sms);
    CUtensorMap pf0, pf1;
    memset(&pf0, 0, sizeof(pf0));
    memset(&pf1, 0, sizeof(pf1));
    int num_prefetch = 0;
    static const bool use_prefetch = getenv("CSR_PREFETCH") != nullptr;  // measured: no gain on B200, off by default
    static const int prefetch_lead = getenv("CSR_PREFETCH_LEAD") ? atoi(getenv("CSR_PREFETCH_LEAD")) : 8;
    if (use_prefetch && (reinterpret_cast<uintptr_t>(MODE == EPI_FISTA ? (const void*)epi_in.z : (const void*)epi_in.b) & 15) == 0) {
        int rc = CSR_OK;
        if (MODE == EPI_FISTA) {
            rc = make_prefetch_map(&pf0, epi_in.z, P, N, epi_in.ld_out);
            if (!rc) rc = make_prefetch_map(&pf1, epi_in.x, P, N, epi_in.ld_out);
            num_prefetch = 2;
        } else if (MODE == EPI_RESID) {
            rc = make_prefetch_map(&pf0, epi_in.b, P, N, epi_in.ld_out);
            num_prefetch = 1;
        }
        if (rc) return rc;
    }
    ProfiledLaunch rec;
    rec.kind = MODE == EPI_SCREEN ? 8 : (MODE == EPI_SCREEN8 ? 9 : (MODE == EPI_PLAIN1 ? 10 : MODE * 2 + (adjoint ? 1 : 0)));
    GemmEpilogue epi = epi_in;
    epi.count_slot = rec.kind;
    if (g_profile) {
        CSR_CUDA(pooled_event(&rec.start));
        CSR_CUDA(pooled_event(&rec.stop));
        CSR_CUDA(cudaEventRecord(rec.start, stream));
    }
    dft_gemm_kernel<MODE><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(
        a_hi, a_lo, MODE == EPI_SCREEN8 ? plan->map_adj8 : (adjoint ? plan->map_adj_hi : plan->map_fwd_hi),
        adjoint ? plan->map_adj_lo : plan->map_fwd_lo, pf0,
        pf1, num_prefetch, prefetch_lead, K, MT, NT,
        // screening needs no chunked accumulation: its bound covers the truncation of one long accumulation chain
        (MODE == EPI_SCREEN || MODE == EPI_SCREEN8 || MODE == EPI_PLAIN1) ? (1 << 20) : gemm_chunk_kb(adjoint, MODE == EPI_FISTA), epi);
    CSR_LAUNCHED();
    if (g_profile) {
        CSR_CUDA(cudaEventRecord(rec.stop, stream));
        g_profiled.push_back(rec);
    }
    return CSR_OK;
}

// The A-operand maps of the FISTA loop are the same every iteration (workspace pointers, P, K, pitch): encoding them
// costs several microseconds of host time per call, which matters once an iteration is ~1 ms of GPU time.  A map is a
// pure function of these arguments, so a small per-thread cache is safe.
static int cached_operand_map(CUtensorMap* map, const void* base, int rows, int cols, int ld, int elem_bytes) {
    struct Entry {
        const void* base;
        int rows, cols, ld, eb;
        CUtensorMap map;
    };
    static thread_local Entry cache[16];
    static thread_local int next = 0;
    for (int i = 0; i < 16; ++i) {
        const Entry& e = cache[i];
        if (e.base == base && e.rows == rows && e.cols == cols && e.ld == ld && e.eb == elem_bytes) {
            *map = e.map;
            return CSR_OK;
        }
    }
    int rc = make_operand_map(map, base, rows, cols, ld, BM, elem_bytes);
    if (rc) return rc;
    Entry& e = cache[next];
    next = (next + 1) % 16;
    e.base = base;
    e.rows = rows;
    e.cols = cols;
    e.ld = ld;
    e.eb = elem_bytes;
    e.map = *map;
    return CSR_OK;
}

// [rows, cols] row-major matrix with pitch ld -> 2-D map with a {box_cols x box_rows} box whose rows span 64 or 128 bytes
// (swizzled accordingly): the staging tiles of the pair kernel's residual epilogue.  Cached like the operand maps.
static int cached_tile_map(CUtensorMap* map, const void* base, int rows, int cols, int ld, int elem_bytes, int box_cols, int box_rows) {
    struct Entry {
        const void* base;
        int rows, cols, ld, eb, bc, br;
        CUtensorMap map;
    };
    static thread_local Entry cache[8];
    static thread_local int next = 0;
    for (int i = 0; i < 8; ++i) {
        const Entry& e = cache[i];
        if (e.base == base && e.rows == rows && e.cols == cols && e.ld == ld && e.eb == elem_bytes && e.bc == box_cols && e.br == box_rows) {
            *map = e.map;
            return CSR_OK;
        }
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return CSR_ERR_CUDA;
    const int row_bytes = box_cols * elem_bytes;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || ((size_t)ld * elem_bytes) % 16 || (row_bytes != 64 && row_bytes != 128)) return CSR_ERR_ARG;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * elem_bytes};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base),
                     dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return CSR_ERR_CUDA;
    Entry& e = cache[next];
    next = (next + 1) % 8;
    e.base = base;
    e.rows = rows;
    e.cols = cols;
    e.ld = ld;
    e.eb = elem_bytes;
    e.bc = box_cols;
    e.br = box_rows;
    e.map = *map;
    return CSR_OK;
}

template <int MODE, bool TMA_EPI>
static int launch_pair3(const csr_plan* plan, bool adjoint, const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& e_in,
                        const CUtensorMap& e_hi, const CUtensorMap& e_lo, int P, const GemmEpilogue& epi_in, cudaStream_t stream) {
    static int max_clusters = 0;
    if (max_clusters == 0) {
        CSR_CUDA(cudaFuncSetAttribute(dft_gemm_pair_kernel<MODE, TMA_EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, P3_SMEM_BYTES));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * plan->num_sms);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = P3_SMEM_BYTES;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, dft_gemm_pair_kernel<MODE, TMA_EPI>, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = plan->num_sms / 2;
        }
        max_clusters = n;
    }
    const int K = adjoint ? 2 * plan->m : 2 * plan->n;
    const int N = adjoint ? 2 * plan->n : 2 * plan->m;
    const int MT = (P + BM - 1) / BM, NT = (N + BN - 1) / BN;
    const int clusters = min(((MT + 1) / 2) * NT, max_clusters);
    GemmEpilogue epi = epi_in;
    ProfiledLaunch rec;
    rec.kind = MODE * 2 + (adjoint ? 1 : 0);
    epi.count_slot = rec.kind;
    if (g_profile) {
        CSR_CUDA(pooled_event(&rec.start));
        CSR_CUDA(pooled_event(&rec.stop));
        CSR_CUDA(cudaEventRecord(rec.start, stream));
    }
    dft_gemm_pair_kernel<MODE, TMA_EPI><<<2 * clusters, NUM_THREADS, P3_SMEM_BYTES, stream>>>(
        a_hi, a_lo, adjoint ? plan->map_adj_hi_half : plan->map_fwd_hi_half, adjoint ? plan->map_adj_lo_half : plan->map_fwd_lo_half, e_in,
        e_hi, e_lo, K, MT, NT, gemm_chunk_kb(adjoint), epi);
    CSR_LAUNCHED();
    if (g_profile) {
        CSR_CUDA(cudaEventRecord(rec.stop, stream));
        g_profiled.push_back(rec);
    }
    return CSR_OK;
}

int launch_dft_gemm(const csr_plan* plan, bool adjoint, const __half* a_hi, const __half* a_lo, int P,
                    const GemmEpilogue& epi_in, cudaStream_t stream) {
    CSR_REQUIRE(P > 0, "P must be positive");
    GemmEpilogue epi = epi_in;
    static const int debug_flags = getenv("CSR_DEBUG_EPI") ? atoi(getenv("CSR_DEBUG_EPI")) : 0;  // measurement aid only
    epi.debug = debug_flags;
    static const int timeline = getenv("CSR_TIMELINE") ? atoi(getenv("CSR_TIMELINE")) : 0;  // measurement aid only
    epi.timeline = timeline;
    if ((timeline == 1 && epi.mode == EPI_RESID) || (timeline == 2 && epi.mode == EPI_FISTA)) {  // the record is of the last launch only
        void* tl = nullptr;
        CSR_CUDA(cudaGetSymbolAddress(&tl, g_timeline));
        CSR_CUDA(cudaMemsetAsync(tl, 0, sizeof(g_timeline), stream));
    }
    epi.no_fast = options().resid_fast == 0;
    static const int no_rows = getenv("CSR_NO_RESID_ROWS") != nullptr;  // measurement aid: fragment-shaped direct epilogue
    epi.no_rows = no_rows;
    const int K = adjoint ? 2 * plan->m : 2 * plan->n;
    const int ldk = adjoint ? plan->ld2m : plan->ld2n;
    epi.P = P;
    epi.N = adjoint ? 2 * plan->n : 2 * plan->m;
    epi.ld_out = adjoint ? plan->ld2n : plan->ld2m;
    epi.m = plan->m;
    CUtensorMap ma_hi, ma_lo;
    int rc;
    if (epi.mode == EPI_SCREEN8) {  // a_hi is the fp8 copy of the operand (bytes, pitch ld8); there is no lo part
        CSR_REQUIRE(adjoint, "fp8 screening exists for the adjoint only");
        rc = cached_operand_map(&ma_hi, a_hi, P, plan->ld8, plan->ld8, 1);
        ma_lo = ma_hi;
    } else {
        rc = cached_operand_map(&ma_hi, a_hi, P, K, ldk, 2);
        if (rc) return rc;
        rc = cached_operand_map(&ma_lo, a_lo, P, K, ldk, 2);
    }
    if (rc) return rc;
    // sparse forward tiles on the sixteen-epilogue-warp kernel, the others on dft_gemm_kernel<EPI_RESID> (which skips the former)
    static const int fwd16_env = getenv("CSR_FWD16") ? atoi(getenv("CSR_FWD16")) : -1;
    const int fwd16 = fwd16_env >= 0 ? fwd16_env : options().fwd16;
    if (fwd16 != 0 && !adjoint && epi.mode == EPI_RESID && epi.group_sparse > 0 && epi.a_kflags != nullptr && epi.tile_list == nullptr &&
        epi.debug == 0 && timeline == 0 && epi.no_fast == 0 && (epi.chan_per_pixel == 0 || epi.chan_scale != nullptr) && (plan->m & 3) == 0 &&
        (epi.ld_out & 3) == 0 &&
        epi.single_rt != nullptr && epi.b != nullptr && (reinterpret_cast<uintptr_t>(epi.b) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(epi.chan_scale) & 15) == 0 &&
        ((reinterpret_cast<uintptr_t>(epi.out_hi) | reinterpret_cast<uintptr_t>(epi.out_lo)) & 7) == 0 &&
        (epi.out8 == nullptr || ((reinterpret_cast<uintptr_t>(epi.out8) | (uintptr_t)epi.ld8 | (uintptr_t)(epi.ld8 >> 1)) & 3) == 0)) {
        static bool configured = false;
        if (!configured) {
            CSR_CUDA(cudaFuncSetAttribute(dft_fwd_direct_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, D16_SMEM_BYTES));
            CSR_CUDA(cudaFuncSetAttribute(dft_fwd_direct_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, D16_SMEM_BYTES));
            CSR_CUDA(cudaFuncSetAttribute(dft_gemm_kernel<EPI_RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
            configured = true;
        }
        const int MT = (P + BM - 1) / BM, NT = (epi.N + BN - 1) / BN, KB = (K + BK - 1) / BK;
        ProfiledLaunch rec;
        rec.kind = 2;
        epi.count_slot = 2;
        if (g_profile) {
            CSR_CUDA(pooled_event(&rec.start));
            CSR_CUDA(pooled_event(&rec.stop));
            CSR_CUDA(cudaEventRecord(rec.start, stream));
        }
        CSR_CUDA(cudaMemsetAsync(epi.single_rt + single_rt_counter_offset(MT), 0, sizeof(int), stream));
        tile_single_kernel<<<(MT + 7) / 8, 256, 0, stream>>>(epi.a_kflags, epi.a_kflag_ld, (KB + 1) >> 1, KB, MT, epi.group_sparse,
                                                            epi.no_kskip == 0, epi.single_rt);
        CSR_LAUNCHED();
        GemmEpilogue rest = epi;
        rest.skip_single = 1;
        CUtensorMap none;
        memset(&none, 0, sizeof(none));
        dft_gemm_kernel<EPI_RESID><<<min(MT * NT, plan->num_sms), NUM_THREADS, SMEM_BYTES, stream>>>(
            ma_hi, ma_lo, plan->map_fwd_hi, plan->map_fwd_lo, none, none, 0, 8, K, MT, NT, gemm_chunk_kb(false), rest);
        CSR_LAUNCHED();
        if (epi.chan_per_pixel != 0)
            dft_fwd_direct_kernel<true><<<min(MT * NT, plan->num_sms), D16_THREADS, D16_SMEM_BYTES, stream>>>(ma_hi, ma_lo, plan->map_fwd_hi,
                                                                                                            plan->map_fwd_lo, K, MT, NT, epi);
        else
            dft_fwd_direct_kernel<false><<<min(MT * NT, plan->num_sms), D16_THREADS, D16_SMEM_BYTES, stream>>>(ma_hi, ma_lo, plan->map_fwd_hi,
                                                                                                             plan->map_fwd_lo, K, MT, NT, epi);
        CSR_LAUNCHED();
        if (g_profile) {
            CSR_CUDA(cudaEventRecord(rec.stop, stream));
            g_profiled.push_back(rec);
        }
        return CSR_OK;
    }
    // the K-block-skipping forward transform with its residual epilogue: 128 x 128 tiles (dft_fwd128_kernel)
    if (options().fwd128 != 0 && !adjoint && epi.mode == EPI_RESID && epi.a_kflags != nullptr && epi.tile_list == nullptr &&
        epi.debug == 0 && options().pair3 < 2 && epi.b != nullptr && (epi.ld_out & 1) == 0 &&
        (reinterpret_cast<uintptr_t>(epi.b) & 7) == 0 && ((reinterpret_cast<uintptr_t>(epi.out_hi) | reinterpret_cast<uintptr_t>(epi.out_lo)) & 3) == 0) {
        static bool configured = false;
        if (!configured) {
            CSR_CUDA(cudaFuncSetAttribute(dft_fwd128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F1_SMEM_BYTES));
            configured = true;
        }
        const int MT = (P + BM - 1) / BM, NT = (epi.N + F1_BN - 1) / F1_BN;
        ProfiledLaunch rec;
        rec.kind = 2;
        epi.count_slot = 2;
        if (g_profile) {
            CSR_CUDA(pooled_event(&rec.start));
            CSR_CUDA(pooled_event(&rec.stop));
            CSR_CUDA(cudaEventRecord(rec.start, stream));
        }
        dft_fwd128_kernel<<<min(MT * NT, plan->num_sms), F1_THREADS, F1_SMEM_BYTES, stream>>>(ma_hi, ma_lo, plan->map_fwd_hi_half,
                                                                                            plan->map_fwd_lo_half, K, MT, NT,
                                                                                            gemm_chunk_kb(false), epi);
        CSR_LAUNCHED();
        if (g_profile) {
            CSR_CUDA(cudaEventRecord(rec.stop, stream));
            g_profiled.push_back(rec);
        }
        return CSR_OK;
    }
    // Three-product transforms of whole matrices on CTA pairs (a tile list names single tiles: those stay on one CTA each).
    // Measured on B200 (profiles/README.md, round 2): the dense transforms gain 8 % (operand delivery: 64 KiB per K block
    // and SM instead of 96); the K-block-skipping forward transform LOSES (0.27 -> 0.31 ms, 0.39 ms with the TMA-staged
    // epilogue: every 2-block accumulation chunk costs a cross-SM hand-shake, and a warp's four 4 KiB passes of dependent
    // bulk copies are latency bound), so the residual transform takes the pair kernel only in the dense regime or on request
    // (option pair3 = 2).
    const bool pair_resid = options().pair3 >= 2 || epi.a_kflags == nullptr || (epi.no_kskip != 0 && epi.group_sparse == 0);
    if (options().pair3 != 0 && options().pair != 0 && P > BM && epi.tile_list == nullptr && epi.debug == 0 &&
        (epi.mode == EPI_PLAIN || (epi.mode == EPI_RESID && pair_resid))) {
        if (epi.mode == EPI_PLAIN) return launch_pair3<EPI_PLAIN, false>(plan, adjoint, ma_hi, ma_lo, ma_hi, ma_hi, ma_hi, P, epi, stream);
        CUtensorMap e_in, e_hi, e_lo;
        const bool staged = options().tma_epi != 0 && options().pair3 >= 2 && epi.b != nullptr && epi.out_hi != nullptr && epi.out_lo != nullptr &&
                            cached_tile_map(&e_in, epi.b, P, epi.N, epi.ld_out, 4, 32, 32) == CSR_OK &&
                            cached_tile_map(&e_hi, epi.out_hi, P, epi.N, epi.ld_out, 2, 32, 32) == CSR_OK &&
                            cached_tile_map(&e_lo, epi.out_lo, P, epi.N, epi.ld_out, 2, 32, 32) == CSR_OK;
        if (staged) return launch_pair3<EPI_RESID, true>(plan, adjoint, ma_hi, ma_lo, e_in, e_hi, e_lo, P, epi, stream);
        return launch_pair3<EPI_RESID, false>(plan, adjoint, ma_hi, ma_lo, ma_hi, ma_hi, ma_hi, P, epi, stream);
    }
    switch (epi.mode) {
        case EPI_PLAIN: return launch_mode<EPI_PLAIN>(plan, adjoint, ma_hi, ma_lo, P, epi, stream);
        case EPI_PLAIN1:
            CSR_REQUIRE(adjoint, "the one-product plain transform exists for the adjoint only");
            return launch_mode<EPI_PLAIN1>(plan, adjoint, ma_hi, ma_lo, P, epi, stream);
        case EPI_RESID: return launch_mode<EPI_RESID>(plan, adjoint, ma_hi, ma_lo, P, epi, stream);
        case EPI_FISTA: return launch_mode<EPI_FISTA>(plan, adjoint, ma_hi, ma_lo, P, epi, stream);
        case EPI_SCREEN:
        case EPI_SCREEN8:
            CSR_REQUIRE(adjoint && epi.abs_sum && epi.tile_state && epi.out_list && epi.out_count,
                        "screening launch: missing argument");
            if (epi.mode == EPI_SCREEN8) return launch_mode<EPI_SCREEN8>(plan, adjoint, ma_hi, ma_lo, P, epi, stream);
            return launch_mode<EPI_SCREEN>(plan, adjoint, ma_hi, ma_lo, P, epi, stream);
    }
    set_error("unknown epilogue mode %d", epi.mode);
    return CSR_ERR_ARG;
}


extern "C" int csr_debug_timeline(long long* out, int cap) {
    CSR_REQUIRE(out && cap > 0, "csr_debug_timeline: bad argument");
    const size_t n = std::min((size_t)cap, sizeof(g_timeline) / sizeof(long long));
    CSR_CUDA(cudaMemcpyFromSymbol(out, g_timeline, n * sizeof(long long)));
    return (int)n;
}

// ---- helpers of the wavelet-domain screening (fista.cu) ------------------------------------------------------------
// eps[p]: everything the two dropped products and the rounding of both evaluations can add to one output element of
// row p (the bound of the fp16 screening above, + 0.1 S_p for the fp32 arithmetic downstream), in output units
__global__ void screen_eps_kernel(const float* __restrict__ abs_sum, const float* __restrict__ a_scale, int K,
                                  float* __restrict__ eps, int P) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < P) eps[p] = (screen_slack(K) + 0.1f) * abs_sum[p] * (kTwiddleDescale / a_scale[p]) * 1.00001f;
}
int launch_screen_eps(const float* abs_sum, const float* a_scale, int K, float* eps, int P, cudaStream_t stream) {
    screen_eps_kernel<<<(P + 255) / 256, 256, 0, stream>>>(abs_sum, a_scale, K, eps, P);
    CSR_LAUNCHED();
    return CSR_OK;
}

// ---- end of an iteration: row clocks of the incremental screening, operand scales that follow the row maxima ----------
// (see "Incremental screening" above.)  One warp per line of sight.  dz = ||z(it+1) - z(it)||_1 and ||z(it+1)||_1 were
// accumulated by EPI_FISTA, S_p = sum |r_hi(it)| by EPI_RESID; the state flags give the phi cells that hold state now.
// Scales: an operand row is rescaled (power of two, target magnitude 2^6) when its maximum came within 2^4 of the fp16
// range -- diverging runs (unit step below the stability limit) grow by less than that per iteration, so nothing
// saturates; runs that stay in range never change a scale and keep their bits.
__global__ void row_update_kernel(RowUpdate u, int P) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= P) return;
    float* rs = u.row_stat + 4 * (size_t)p;
    const float dz = rs[0], zabs = rs[1];
    const float zmax = rs[2], rmax_scaled = rs[3];   // (bit patterns of non-negative floats)
    __syncwarp();
    if (lane < 4) rs[lane] = 0.f;
    if (u.incr.pathlen != nullptr) {
        // Iteration 0 (the dense starting point does not count: nothing is recorded before iteration 1): the phi cells
        // that hold state now, from the state flags -- block b covers columns [128 b, 128 b + 127] of [Re (n) | Im (n)].
        // Later iterations: EPI_FISTA widens the range itself when a block turns on (GemmEpilogue::ext).
        if (u.iter == 0) {
            int lo = 0x7fffffff, hi = -1;
            if (u.zero_flags != nullptr) {
                const unsigned char* fl = u.zero_flags + (size_t)p * u.flag_ld;
                for (int b = lane; b < u.flag_ld; b += 32) {
                    if (!fl[b]) continue;
                    const int c0 = 128 * b, c1 = min(128 * b + 127, 2 * u.n - 1);
                    if (c1 < u.n) {
                        lo = min(lo, c0), hi = max(hi, c1);
                    } else if (c0 >= u.n) {
                        lo = min(lo, c0 - u.n), hi = max(hi, c1 - u.n);
                    } else {  // the block that straddles the two halves: the last cells of Re and the first of Im
                        lo = 0, hi = u.n - 1;
                    }
                }
            } else {
                lo = 0, hi = u.n - 1;
            }
#pragma unroll
            for (int sft = 16; sft > 0; sft >>= 1) {
                lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, sft));
                hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, sft));
            }
            if (lane == 0) u.incr.ext[p] = make_int2(lo, hi);
        }
        if (lane == 0) {
            const float2 c = u.incr.crow[p];
            const float a = 2.f * (c.x + c.y);                                                    // trivial coefficient
            const float r_l1 = u.abs_sum ? u.abs_sum[p] / u.rscale[p] * 1.001f + a * dz : 0.f;   // >= ||r(it+1)||_1
            const float rr = fmaxf(u.incr.rhat[p], a * zabs * (1.f / 4096.f) + r_l1 * (1.f / 524288.f));
            u.incr.rhat[p] = rr;
            u.incr.pathlen[p] += dz * 1.001f;
            u.incr.qclock[p] = ((float)(u.iter + 1) * u.noise[p] + rr) * 1.0001f;
        }
    }
    if (u.rescale && lane == 0) {
        const float zs = u.zscale_write[p];
        u.zscale_read[p] = zs;                       // the operand EPI_FISTA just stored carries this scale
        if (zmax * zs >= 4096.f) u.zscale_write[p] = operand_scale(zmax);
        if (rmax_scaled >= 4096.f) u.rscale[p] = operand_scale(rmax_scaled / u.rscale[p]);
    }
}
// the same bookkeeping with one thread per line of sight (every iteration but the first)
__global__ void row_update_thread_kernel(RowUpdate u, int P) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    float4* rs = reinterpret_cast<float4*>(u.row_stat) + p;
    const float4 st = *rs;                           // dz, ||z||_1, max |z|, max |r| (scaled)
    *rs = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u.incr.pathlen != nullptr) {
        const float2 c = u.incr.crow[p];
        const float a = 2.f * (c.x + c.y);
        const float r_l1 = u.abs_sum ? u.abs_sum[p] / u.rscale[p] * 1.001f + a * st.x : 0.f;
        const float rr = fmaxf(u.incr.rhat[p], a * st.y * (1.f / 4096.f) + r_l1 * (1.f / 524288.f));
        u.incr.rhat[p] = rr;
        u.incr.pathlen[p] += st.x * 1.001f;
        u.incr.qclock[p] = ((float)(u.iter + 1) * u.noise[p] + rr) * 1.0001f;
    }
    if (u.rescale) {
        const float zs = u.zscale_write[p];
        u.zscale_read[p] = zs;
        if (st.z * zs >= 4096.f) u.zscale_write[p] = operand_scale(st.z);
        if (st.w >= 4096.f) u.rscale[p] = operand_scale(st.w / u.rscale[p]);
    }
}
int launch_row_update(const RowUpdate& u, int P, cudaStream_t stream) {
    cudaEvent_t ev0, ev1;
    const bool prof = profile_start(stream, &ev0, &ev1);
    if (u.iter == 0 || (reinterpret_cast<uintptr_t>(u.row_stat) & 15) != 0)
        row_update_kernel<<<(P * 32 + 255) / 256, 256, 0, stream>>>(u, P);
    else
        row_update_thread_kernel<<<(P + 255) / 256, 256, 0, stream>>>(u, P);
    CSR_LAUNCHED();
    if (prof) profile_stop(13, stream, ev0, ev1);
    return CSR_OK;
}

// |RMTF| of the base channel set at the offsets k cell, k = 0 .. n-1 (fp64 phases like the tables), then its suffix maximum
__global__ void rmtf_abs_kernel(const double* __restrict__ l2, double l2_ref, int m, const float* __restrict__ w, double cell,
                                int n, float* __restrict__ out) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const double d = 2.0 * cell * (double)k;
    double re = 0.0, im = 0.0;
    int base = 0;
    for (int i = 0; i < m; ++i) {
        if (w != nullptr && !(w[i] > 0.f)) continue;
        double sn, cs;
        sincos((l2[i] - l2_ref) * d, &sn, &cs);
        re += cs;
        im += sn;
        ++base;
    }
    out[k] = base > 0 ? (float)(sqrt(re * re + im * im) / (double)base * 1.001 + 1e-6) : 0.f;
}
__global__ void __launch_bounds__(1024) suffix_max_kernel(const float* __restrict__ in, int n, float* __restrict__ out) {
    __shared__ float part[1024];
    // thread t owns the contiguous chunk [t c, (t + 1) c): chunk maxima, then a suffix maximum over the chunks
    const int c = (n + 1023) / 1024, t = threadIdx.x;
    float mx = 0.f;
    for (int k = min(n, (t + 1) * c) - 1; k >= t * c; --k) mx = fmaxf(mx, in[k]);
    part[t] = mx;
    __syncthreads();
    float tail = 0.f;  // maximum over the chunks after mine
    for (int q = t + 1; q < 1024; ++q) tail = fmaxf(tail, part[q]);
    for (int k = min(n, (t + 1) * c) - 1; k >= t * c; --k) {
        tail = fmaxf(tail, in[k]);
        out[k] = tail;
    }
}
int launch_rmtf_envelope(const csr_plan* plan, const float* w, float* env, float* scratch, cudaStream_t stream) {
    rmtf_abs_kernel<<<(plan->n + 127) / 128, 128, 0, stream>>>(plan->d_l2, plan->l2_ref, plan->m, w, plan->cell, plan->n, scratch);
    CSR_LAUNCHED();
    suffix_max_kernel<<<1, 1024, 0, stream>>>(scratch, plan->n, env);
    CSR_LAUNCHED();
    return CSR_OK;
}

// every tile goes to one of two lists (tile ids in the rasterised numbering of tile_coords); the lists are short and their
// order does not matter to the kernels that walk them, so tiles are appended with one atomic per warp (counts zeroed here
// by the first thread of a preceding memset-free protocol: the caller zeroes them)
__global__ void __launch_bounds__(256)
split_tile_lists_kernel(const unsigned char* __restrict__ need, const unsigned char* __restrict__ exclude, int MT, int NT,
                        int* __restrict__ list_a, int* __restrict__ count_a, int* __restrict__ list_b,
                        int* __restrict__ count_b) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    bool a = false, b = false;
    if (t < MT * NT) {
        int mt, nt;
        tile_coords(t, MT, NT, mt, nt);
        a = need[(size_t)mt * NT + nt] != 0 && (exclude == nullptr || exclude[(size_t)mt * NT + nt] == 0);
        b = !a && list_b != nullptr && (exclude == nullptr);
    }
    const unsigned ba = __ballot_sync(0xffffffffu, a), bb = __ballot_sync(0xffffffffu, b);
    int base_a = 0, base_b = 0;
    if (lane == 0) {
        if (ba) base_a = atomicAdd(count_a, __popc(ba));
        if (bb) base_b = atomicAdd(count_b, __popc(bb));
    }
    base_a = __shfl_sync(0xffffffffu, base_a, 0);
    base_b = __shfl_sync(0xffffffffu, base_b, 0);
    if (a) list_a[base_a + __popc(ba & ((1u << lane) - 1))] = t;
    if (b) list_b[base_b + __popc(bb & ((1u << lane) - 1))] = t;
}
int launch_split_tile_lists(const unsigned char* need, const unsigned char* exclude, int P, int N, int* list_a, int* count_a,
                            int* list_b, int* count_b, cudaStream_t stream) {
    const int MT = (P + BM - 1) / BM, NT = (N + BN - 1) / BN;
    cudaEvent_t ev0, ev1;
    const bool prof = profile_start(stream, &ev0, &ev1);
    CSR_CUDA(cudaMemsetAsync(count_a, 0, sizeof(int), stream));
    if (count_b) CSR_CUDA(cudaMemsetAsync(count_b, 0, sizeof(int), stream));
    split_tile_lists_kernel<<<(MT * NT + 255) / 256, 256, 0, stream>>>(need, exclude, MT, NT, list_a, count_a, list_b, count_b);
    CSR_LAUNCHED();
    if (prof) profile_stop(12, stream, ev0, ev1);
    return CSR_OK;
}

template <int PMODE>
static int launch_pair_kernel(const csr_plan* plan, const CUtensorMap& ma, const CUtensorMap& mb, int P, const GemmEpilogue& epi_in,
                              cudaStream_t stream) {
    GemmEpilogue epi = epi_in;
    epi.count_slot = (PMODE == 1 || PMODE == 5) ? 10 : ((PMODE == 2 || PMODE == 4) ? 8 : 9);
    const bool sym = PMODE >= 3;
    // symmetric modes: K = one half row, MT counts groups of 64 lines of sight (128 half rows), NT the tiles of 128
    // non-negative phi (256 half rows of the table) plus the tile of the one cell without a mirror
    const int K = sym ? (PMODE == 3 ? plan->ld8 / 2 : plan->m) : (PMODE == 0 ? plan->ld8 : 2 * plan->m);  // (fp8 rows are padded)
    const int MT = sym ? (P + BM / 2 - 1) / (BM / 2) : (P + BM - 1) / BM;
    const int NT = sym ? (plan->n + BN - 1) / BN + 1 : (epi.N + BN - 1) / BN;
    static int max_clusters = 0;
    if (max_clusters == 0) {
        CSR_CUDA(cudaFuncSetAttribute(screen8_pair_kernel<PMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * plan->num_sms);
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = P2_SMEM_BYTES;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, screen8_pair_kernel<PMODE>, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = plan->num_sms / 2;
        }
        max_clusters = n;
    }
    const int pairs = ((MT + 1) / 2) * NT;
    const int clusters = min(pairs, max_clusters);
    ProfiledLaunch rec;
    if (g_profile) {
        rec.kind = (PMODE == 1 || PMODE == 5) ? 10 : ((PMODE == 2 || PMODE == 4) ? 8 : 9);
        CSR_CUDA(pooled_event(&rec.start));
        CSR_CUDA(pooled_event(&rec.stop));
        CSR_CUDA(cudaEventRecord(rec.start, stream));
    }
    screen8_pair_kernel<PMODE><<<2 * clusters, NUM_THREADS, P2_SMEM_BYTES, stream>>>(ma, mb, K, MT, NT, epi);
    CSR_LAUNCHED();
    if (g_profile) {
        CSR_CUDA(cudaEventRecord(rec.stop, stream));
        g_profiled.push_back(rec);
    }
    return CSR_OK;
}

// fp8 screening on CTA pairs; `a8` is the fp8 copy of the residual operand ([P, ld8] bytes)
int launch_screen8_pair(const csr_plan* plan, const unsigned char* a8, int P, const GemmEpilogue& epi_in, cudaStream_t stream) {
    CSR_REQUIRE(P > 0, "P must be positive");
    GemmEpilogue epi = epi_in;
    epi.P = P;
    epi.N = 2 * plan->n;
    epi.ld_out = plan->ld2n;
    epi.m = plan->m;
    CSR_REQUIRE(epi.abs_sum && epi.tile_state && epi.out_list && epi.out_count, "screening launch: missing argument");
    CUtensorMap ma;
    int rc = cached_operand_map(&ma, a8, P, plan->ld8, plan->ld8, 1);
    if (rc) return rc;
    return launch_pair_kernel<0>(plan, ma, plan->map_adj8_half, P, epi, stream);
}

size_t sym_scratch_bytes(const csr_plan* plan, int P) {
    const size_t MTo = (P + BM - 1) / BM, NTo = (2 * plan->n + BN - 1) / BN;
    const size_t MTg = (P + BM / 2 - 1) / (BM / 2), NTs = (plan->n + BN - 1) / BN + 1;
    return align256(MTo * NTo) + align256(MTg * NTs);
}

// 4-D view of a [P, 2 halves] operand for the symmetric kernels: (k, line of sight within a group of 8, half, group), box
// {one 128-byte row, 8, 2, 8} = 128 tile rows ordered [group][half][line of sight]
static int cached_sym_map(CUtensorMap* map, const void* base, int P, int m, size_t row_bytes, size_t half_bytes, int elem_bytes) {
    struct Entry {
        const void* base;
        int P, m, eb;
        size_t rb, hb;
        CUtensorMap map;
    };
    static thread_local Entry cache[8];
    static thread_local int next = 0;
    for (int i = 0; i < 8; ++i) {
        const Entry& e = cache[i];
        if (e.base == base && e.P == P && e.m == m && e.eb == elem_bytes && e.rb == row_bytes && e.hb == half_bytes) {
            *map = e.map;
            return CSR_OK;
        }
    }
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return CSR_ERR_CUDA;
    if ((reinterpret_cast<uintptr_t>(base) & 15) || row_bytes % 16 || half_bytes % 16) return CSR_ERR_ARG;
    cuuint64_t dims[4] = {(cuuint64_t)m, 8, 2, (cuuint64_t)((P + 7) / 8)};
    cuuint64_t strides[3] = {(cuuint64_t)row_bytes, (cuuint64_t)half_bytes, (cuuint64_t)(8 * row_bytes)};
    cuuint32_t box[4] = {(cuuint32_t)(BK * 2 / elem_bytes), 8, 2, 8};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(map, elem_bytes == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return CSR_ERR_CUDA;  // (the caller falls back to the 2-D view)
    Entry& e = cache[next];
    next = (next + 1) % 8;
    e.base = base;
    e.P = P;
    e.m = m;
    e.eb = elem_bytes;
    e.rb = row_bytes;
    e.hb = half_bytes;
    e.map = *map;
    return CSR_OK;
}

int launch_screen_sym_pair(const csr_plan* plan, const void* a, bool fp8, int P, const GemmEpilogue& epi_in, unsigned char* need,
                           int* pair_list, int* pair_count, const IncrState* incr, bool record, cudaStream_t stream) {
    CSR_REQUIRE(P > 0 && plan->sym_ok && (fp8 || plan->sym16_ok) && need && pair_list && pair_count,
                "symmetric screening: not available for this plan");
    GemmEpilogue epi = epi_in;
    epi.P = P;
    epi.N = 2 * plan->n;
    epi.ld_out = plan->ld2n;
    epi.m = plan->m;
    CSR_REQUIRE(epi.abs_sum && epi.tile_state && epi.out_list && epi.out_count && !epi.tile_list, "pair screening: missing argument");
    const int MTo = (P + BM - 1) / BM, NTo = (2 * plan->n + BN - 1) / BN;
    epi.need = need;
    epi.need_ld = NTo;
    const int MTg = (P + BM / 2 - 1) / (BM / 2), NTs = (plan->n + BN - 1) / BN + 1;
    unsigned char* symq = need + align256((size_t)MTo * NTo);  // (the scratch holds both maps: sym_scratch_bytes)
    epi.symq = symq;
    CUtensorMap ma;
    // rows regrouped so that a thread holds both rows of a line of sight (needs the operand allocated with its row count
    // rounded up to 8: fista.cu); 2-D half-row view otherwise
    static const bool rows8 = getenv("CSR_NO_SYM_ROWS8") == nullptr;
    int rc = CSR_ERR_ARG;
    if (rows8)
        rc = fp8 ? cached_sym_map(&ma, a, P, plan->m, (size_t)plan->ld8, (size_t)plan->ld8 / 2, 1)
                 : cached_sym_map(&ma, a, P, plan->m, (size_t)plan->ld2m * 2, (size_t)plan->m * 2, 2);
    epi.sym_rows8 = rc == CSR_OK ? 1 : 0;
    if (rc != CSR_OK)
        rc = fp8 ? cached_operand_map(&ma, a, 2 * P, plan->m, plan->ld8 / 2, 1) : cached_operand_map(&ma, a, 2 * P, plan->m, plan->m, 2);
    if (rc) return rc;
    // incremental screening needs the per-row test of the regrouped operand view, the row clocks and the bound's inputs
    IncrState st = {};
    if (incr) st = *incr;
    if (!epi.sym_rows8 || !st.pathlen || !st.env || !epi.qclock || !epi.rhat) epi.expiry = nullptr;
    cudaEvent_t ev0, ev1;
    const bool prof = profile_start(stream, &ev0, &ev1);
    CSR_CUDA(cudaMemsetAsync(need, 0, (size_t)MTo * NTo, stream));
    CSR_CUDA(cudaMemsetAsync(pair_count, 0, sizeof(int), stream));
    const int MTP = (MTg + 1) / 2;
    sym_select_kernel<<<(MTP * NTs * 32 + 255) / 256, 256, 0, stream>>>(epi.tile_state, epi.flag_ld, MTg, NTs, plan->n, symq, need,
                                                                       NTo, pair_list, pair_count, epi.expiry, st, P, record ? 1 : 0);
    CSR_LAUNCHED();
    if (prof) profile_stop(12, stream, ev0, ev1);
    if (!record) epi.expiry = nullptr;  // (first iteration: the rounding allowance of the starting point is not known yet)
    epi.tile_list = pair_list;
    epi.tile_count = pair_count;
    rc = fp8 ? launch_pair_kernel<3>(plan, ma, plan->map_sym8, P, epi, stream) : launch_pair_kernel<4>(plan, ma, plan->map_sym16, P, epi, stream);
    if (rc) return rc;
    return launch_split_tile_lists(need, nullptr, P, 2 * plan->n, epi.out_list, epi.out_count, nullptr, nullptr, stream);
}

// the fp16 screening test on CTA pairs over every quiet tile (no fp8 pass in front): for residuals whose l1 norm makes
// the fp8 bound useless (per-pixel weights: the weighted residual stays signal-sized)
int launch_screen16_pair(const csr_plan* plan, const __half* a_hi, int P, const GemmEpilogue& epi_in, cudaStream_t stream) {
    CSR_REQUIRE(P > 0, "P must be positive");
    GemmEpilogue epi = epi_in;
    epi.P = P;
    epi.N = 2 * plan->n;
    epi.ld_out = plan->ld2n;
    epi.m = plan->m;
    CSR_REQUIRE(epi.abs_sum && epi.tile_state && epi.out_list && epi.out_count && !epi.tile_list, "pair screening: missing argument");
    CUtensorMap ma;
    int rc = cached_operand_map(&ma, a_hi, P, 2 * plan->m, plan->ld2m, 2);
    if (rc) return rc;
    return launch_pair_kernel<2>(plan, ma, plan->map_adj_hi_half, P, epi, stream);
}

// the one-product plain adjoint (EPI_PLAIN1) on CTA pairs
int launch_plain1_pair(const csr_plan* plan, const __half* a_hi, int P, const GemmEpilogue& epi_in, cudaStream_t stream) {
    CSR_REQUIRE(P > 0 && epi_in.out != nullptr, "launch_plain1_pair: bad argument");
    GemmEpilogue epi = epi_in;
    epi.mode = EPI_PLAIN;
    epi.P = P;
    epi.N = 2 * plan->n;
    epi.ld_out = plan->ld2n;
    epi.m = plan->m;
    epi.debug = 0;
    CUtensorMap ma;
    int rc = cached_operand_map(&ma, a_hi, P, 2 * plan->m, plan->ld2m, 2);
    if (rc) return rc;
    return launch_pair_kernel<1>(plan, ma, plan->map_adj_hi_half, P, epi, stream);
}

// the one-product plain adjoint, mirror-symmetric (needs plan->sym16_ok and the 4-D operand view); epi.tile_list names
// (pixel tile, symmetric column tile) pairs: launch_sym_pair_list
int launch_plain1_sym_pair(const csr_plan* plan, const __half* a_hi, int P, const GemmEpilogue& epi_in, cudaStream_t stream) {
    CSR_REQUIRE(P > 0 && epi_in.out != nullptr && plan->sym16_ok, "launch_plain1_sym_pair: bad argument");
    GemmEpilogue epi = epi_in;
    epi.mode = EPI_PLAIN;
    epi.P = P;
    epi.N = 2 * plan->n;
    epi.ld_out = plan->ld2n;
    epi.m = plan->m;
    epi.debug = 0;
    epi.sym_rows8 = 1;
    CUtensorMap ma;
    int rc = cached_sym_map(&ma, a_hi, P, plan->m, (size_t)plan->ld2m * 2, (size_t)plan->m * 2, 2);
    if (rc) return rc;
    return launch_pair_kernel<5>(plan, ma, plan->map_sym16, P, epi, stream);
}
bool plain1_sym_available(const csr_plan* plan, const __half* a_hi, int P) {
    static const bool on = getenv("CSR_NO_SYM_PLAIN1") == nullptr;
    CUtensorMap ma;
    return on && plan->sym16_ok && cached_sym_map(&ma, a_hi, P, plan->m, (size_t)plan->ld2m * 2, (size_t)plan->m * 2, 2) == CSR_OK;
}

// (pixel tile, symmetric column tile) pairs that overlap an original tile whose need byte is 0
__global__ void __launch_bounds__(1024)
sym_pair_list_kernel(const unsigned char* __restrict__ need, int MTo, int NTo, int NTs, int nphi, int* __restrict__ list,
                     int* __restrict__ count) {
    __shared__ int warp_n[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int t0 = 0; t0 < MTo * NTs; t0 += 1024) {
        const int t = t0 + threadIdx.x;
        bool a = false;
        if (t < MTo * NTs) {
            int mtp, nt;
            tile_coords(t, MTo, NTs, mtp, nt);
            int lo[4], hi[4];
            const int nr = sym_ranges(nt, NTs, nphi, lo, hi);
            for (int r = 0; r < nr; ++r)
                for (int c = lo[r] >> 8; c <= (hi[r] - 1) >> 8; ++c) a |= need[(size_t)mtp * NTo + c] == 0;
        }
        const unsigned ba = __ballot_sync(0xffffffffu, a);
        if (lane == 0) warp_n[warp] = __popc(ba);
        __syncthreads();
        int oa = base;
        for (int w = 0; w < warp; ++w) oa += warp_n[w];
        if (a) list[oa + __popc(ba & ((1u << lane) - 1))] = t;
        __syncthreads();
        if (threadIdx.x == 0)
            for (int w = 0; w < 32; ++w) base += warp_n[w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}
int launch_sym_pair_list(const csr_plan* plan, const unsigned char* need, int P, int* list, int* count, cudaStream_t stream) {
    sym_pair_list_kernel<<<1, 1024, 0, stream>>>(need, (P + BM - 1) / BM, (2 * plan->n + BN - 1) / BN, (plan->n + BN - 1) / BN + 1,
                                                plan->n, list, count);
    CSR_LAUNCHED();
    return CSR_OK;
}

// pairs of pixel tiles (numbering of tile_coords(tp, ceil(MT/2), NT)) with an existing member whose need byte is 0
__global__ void __launch_bounds__(1024)
pair_list_kernel(const unsigned char* __restrict__ need, int MT, int NT, int* __restrict__ list, int* __restrict__ count) {
    __shared__ int warp_n[32];
    __shared__ int base;
    if (threadIdx.x == 0) base = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int MTP = (MT + 1) / 2;
    for (int t0 = 0; t0 < MTP * NT; t0 += 1024) {
        const int t = t0 + threadIdx.x;
        bool a = false;
        if (t < MTP * NT) {
            int mtp, nt;
            tile_coords(t, MTP, NT, mtp, nt);
            a = need[(size_t)(2 * mtp) * NT + nt] == 0 || (2 * mtp + 1 < MT && need[(size_t)(2 * mtp + 1) * NT + nt] == 0);
        }
        const unsigned ba = __ballot_sync(0xffffffffu, a);
        if (lane == 0) warp_n[warp] = __popc(ba);
        __syncthreads();
        int oa = base;
        for (int w = 0; w < warp; ++w) oa += warp_n[w];
        if (a) list[oa + __popc(ba & ((1u << lane) - 1))] = t;
        __syncthreads();
        if (threadIdx.x == 0)
            for (int w = 0; w < 32; ++w) base += warp_n[w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = base;
}
int launch_pair_list(const unsigned char* need, int P, int N, int* list, int* count, cudaStream_t stream) {
    pair_list_kernel<<<1, 1024, 0, stream>>>(need, (P + BM - 1) / BM, (N + BN - 1) / BN, list, count);
    CSR_LAUNCHED();
    return CSR_OK;
}

}  // namespace csr

using namespace csr;

// kinds: 0 plain-forward, 1 plain-adjoint, 2 resid-forward, 3 resid-adjoint, 4 fista-forward, 5 fista-adjoint,
//        6 fused wavelet synthesis, 7 fused wavelet analysis + update, 8 fp16 screening adjoint, 9 fp8 screening adjoint,
//        10 one-product plain adjoint (wavelet-domain screening), 11 wavelet analysis + update, redo pass
extern "C" int csr_profile(int enable) {
    g_profile = enable != 0;
    if (g_profile) {  // create the events up front: cudaEventCreate inside a timed region is host time per launch
        while (g_event_pool.size() < 8192) {
            cudaEvent_t ev;
            CSR_CUDA(cudaEventCreate(&ev));
            g_event_pool.push_back(ev);
        }
    }
    return CSR_OK;
}

extern "C" int csr_profile_collect(double* ms_by_kind, long long* count_by_kind) {
    CSR_REQUIRE(ms_by_kind && count_by_kind, "csr_profile_collect: null pointer");
    for (int i = 0; i < CSR_PROFILE_KINDS; ++i) {
        ms_by_kind[i] = 0.0;
        count_by_kind[i] = 0;
    }
    for (auto& r : g_profiled) {
        CSR_CUDA(cudaEventSynchronize(r.stop));
        float ms = 0.f;
        CSR_CUDA(cudaEventElapsedTime(&ms, r.start, r.stop));
        ms_by_kind[r.kind] += ms;
        count_by_kind[r.kind] += 1;
        g_event_pool.push_back(r.start);
        g_event_pool.push_back(r.stop);
    }
    g_profiled.clear();
    return CSR_OK;
}
