// sparse.cu -- the L1 solution leaves the device compacted.
//
// FISTA's soft threshold (objectivefunction/priors/l1.py:30-32) leaves > 99 % of a solution slab identically zero; the
// reference's recipe writes it into dense output planes anyway (README.md:374-384).  Shipping a [P, 2n] fp32 slab over
// PCIe (1 GB per 16384 lines of sight at n = 7968) is what limited end-to-end scaling in round 1, so the solution can
// be read back in compressed-row form instead: counts per line of sight, then (column, value) lists in row order.
//   csr_sparse_count  counts[p] = #{c < ncols : x[p, c] != 0}                       one warp per row, 16-byte loads
//   csr_sparse_fill   idx / val [offsets[p] .. offsets[p+1]) = the non-zeros of row p in ascending column order
// offsets = exclusive prefix sum of counts (the caller's: a cumsum is plumbing).  Both kernels stream the slab once at
// HBM speed; rows are independent, so there is no atomics-ordered output and the result is deterministic.
#include "plan.cuh"

namespace csr {

__global__ void sparse_count_kernel(const float* __restrict__ x, int ld, int ncols, int P, int* __restrict__ counts) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= P) return;
    const float* row = x + (size_t)p * ld;
    int n = 0;
    const int n4 = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) ? ncols >> 2 : 0;  // (pitch is a multiple of 8 floats)
    for (int i = lane; i < n4; i += 32) {
        const float4 v = __ldcs(reinterpret_cast<const float4*>(row) + i);
        n += (v.x != 0.f) + (v.y != 0.f) + (v.z != 0.f) + (v.w != 0.f);
    }
    for (int c = 4 * n4 + lane; c < ncols; c += 32) n += row[c] != 0.f;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) n += __shfl_xor_sync(0xffffffffu, n, s);
    if (lane == 0) counts[p] = n;
}

__global__ void sparse_fill_kernel(const float* __restrict__ x, int ld, int ncols, int P, const long long* __restrict__ offsets,
                                   int* __restrict__ idx, float* __restrict__ val, long long capacity) {
    const int p = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (p >= P) return;
    const float* row = x + (size_t)p * ld;
    long long at = offsets[p];
    if (offsets[p + 1] == at) return;  // (an all-zero row: nothing to scan for)
    for (int c0 = 0; c0 < ncols; c0 += 32) {
        const int c = c0 + lane;
        const float v = c < ncols ? __ldcs(row + c) : 0.f;
        const unsigned m = __ballot_sync(0xffffffffu, v != 0.f);
        if (v != 0.f) {
            const long long o = at + __popc(m & ((1u << lane) - 1));
            if (o < capacity) {  // (lists sized before the count was known: the caller compares offsets[P] with it)
                idx[o] = c;
                val[o] = v;
            }
        }
        at += __popc(m);
    }
}

}  // namespace csr

using namespace csr;

extern "C" int csr_sparse_count(const float* x, int ld, int ncols, int P, int* counts, void* stream) {
    CSR_REQUIRE(x && counts && P > 0 && ncols > 0 && ld >= ncols, "csr_sparse_count: bad argument");
    sparse_count_kernel<<<(P * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, ld, ncols, P, counts);
    CSR_LAUNCHED();
    return CSR_OK;
}

extern "C" int csr_sparse_fill(const float* x, int ld, int ncols, int P, const long long* offsets, int* idx, float* val,
                               long long capacity, void* stream) {
    CSR_REQUIRE(x && offsets && idx && val && P > 0 && ncols > 0 && ld >= ncols && capacity >= 0, "csr_sparse_fill: bad argument");
    sparse_fill_kernel<<<(P * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(x, ld, ncols, P, offsets, idx, val, capacity);
    CSR_LAUNCHED();
    return CSR_OK;
}
