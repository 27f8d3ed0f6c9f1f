// Device-side pieces of request staging shared by catalog.cu (staging kernels) and small.cu (one-CTA kernel for tiny
// catalogs): zero-copy fetch of the request from the caller's pinned host block, query normalisation and the profile
// build (lib.py:51-52: mean of the liked rows' unit vectors).  Shared so that every route produces the same bits.
#pragma once

#include "exact.cuh"

namespace rebert {

// Every load below crosses PCIe (~2 us round trip): issue them all before the first use — 16-byte loads, four per
// thread in flight (the pinned block is 16-byte aligned).
__device__ __forceinline__ void fetch_query_zero_copy(const float* __restrict__ q_host, int d, float* s_src) {
    const int d4 = d >> 2;
    const float4* q4 = (const float4*)q_host;
    for (int c0 = threadIdx.x; c0 < d4; c0 += 4 * blockDim.x) {
        float4 v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j * blockDim.x;
            v[j] = c < d4 ? q4[c] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + j * blockDim.x;
            if (c < d4) ((float4*)s_src)[c] = v[j];
        }
    }
    for (int c = (d4 << 2) + threadIdx.x; c < d; c += blockDim.x) s_src[c] = q_host[c];
}

// int32 list in pinned host memory -> dst (device or shared memory), four loads per thread in flight
__device__ __forceinline__ void copy_list_zero_copy(const int32_t* __restrict__ src_host, int n, int32_t* dst) {
    int e[4];
    for (int c0 = threadIdx.x; c0 < n; c0 += 4 * blockDim.x) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { const int c = c0 + j * blockDim.x; e[j] = c < n ? src_host[c] : 0; }
#pragma unroll
        for (int j = 0; j < 4; ++j) { const int c = c0 + j * blockDim.x; if (c < n) dst[c] = e[j]; }
    }
}

// ||q||_2 in fp64 (zero -> 1, sklearn normalize) with the summation order of a 256-thread CTA: thread t sums elements
// t, t + 256, ...; xor-shuffle tree per warp; the 8 warp sums through one more tree.  Callable from larger CTAs (threads
// >= 256 contribute nothing), so every kernel normalises a query to the same bits.  All threads must call.
__device__ __forceinline__ double query_norm_256(const float* s_src, int d, double* red /* shared [32] */) {
    double acc = 0.0;
    if (threadIdx.x < 256) {
        for (int c = threadIdx.x; c < d; c += 256) {
            const double x = (double)s_src[c];
            acc = fma(x, x, acc);
        }
    }
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0 && threadIdx.x < 256) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < 8 ? red[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    double nrm = sqrt(red[0]);
    if (nrm == 0.0) nrm = 1.0;
    return nrm;
}

template <typename T> struct RowChunk;
template <> struct RowChunk<float> {
    static constexpr int EPC = 4;
    __device__ static __forceinline__ void unpack(const uint4& v, double* x) {
        x[0] = (double)__uint_as_float(v.x); x[1] = (double)__uint_as_float(v.y);
        x[2] = (double)__uint_as_float(v.z); x[3] = (double)__uint_as_float(v.w);
    }
};
template <> struct RowChunk<__nv_bfloat16> {
    static constexpr int EPC = 8;
    __device__ static __forceinline__ void unpack(const uint4& v, double* x) {
        x[0] = (double)bf16lo(v.x); x[1] = (double)bf16hi(v.x); x[2] = (double)bf16lo(v.y); x[3] = (double)bf16hi(v.y);
        x[4] = (double)bf16lo(v.z); x[5] = (double)bf16hi(v.z); x[6] = (double)bf16lo(v.w); x[7] = (double)bf16hi(v.w);
    }
};

constexpr int kProfBlock = 512;     // entries staged per round
constexpr int kProfMaxIter = 3;     // 16-byte chunks a thread of a 256-thread CTA may own: rows up to 12 KB

// One CTA builds one profile partial: sum64[c] = sum over the entries whose GLOBAL row `col[i]` lies in this shard of
// w_i * row[c] / ||row|| (DIV: element-wise division first, sklearn's order; else one reciprocal per row, 1 ulp apart),
// entries walked in list order so the sum is deterministic; wsum[0] = sum of ALL weights (w == nullptr: the count).
// col / w may live in device memory or in device-addressable pinned host memory (each entry is read once per CTA).
// INFLIGHT = row loads a thread keeps in flight: 4 when thousands of CTAs run (batched build), 16 for a single request,
// where the one CTA is pure latency.  The accumulation order does not depend on it.
template <typename T, bool DIV, int INFLIGHT = 4>
__device__ __forceinline__ void profile_accumulate_cta(const T* __restrict__ rows, const double* __restrict__ norm64, int64_t n,
                                                       int64_t row_base, int ld, const int32_t* __restrict__ col,
                                                       const float* __restrict__ w, int count, double* __restrict__ sum64,
                                                       double* __restrict__ wsum, int chunk_lo = 0, int chunk_hi = 0x7fffffff) {
    // [chunk_lo, chunk_hi): the 16-byte column chunks this CTA accumulates and writes.  Columns are independent and every column
    // walks the rows in list order, so splitting them over several CTAs (single request: one SM's fp64 unit is the bottleneck)
    // changes no bit; the weight sum is computed by every CTA in full, with the same thread layout.
    constexpr int EPC = RowChunk<T>::EPC;
    __shared__ int s_row[kProfBlock];
    __shared__ double s_w[kProfBlock];      // weight (DIV) or weight / norm (!DIV)
    __shared__ double s_n[kProfBlock];      // norm (DIV only)
    __shared__ double s_y[kProfBlock];      // RN(1 / norm) (DIV only): x / norm comes from it by two FMA corrections (exact.cuh)
    __shared__ double s_red[32];
    const int chunks = ld / EPC;
    double acc[kProfMaxIter][EPC];
#pragma unroll
    for (int it = 0; it < kProfMaxIter; ++it)
#pragma unroll
        for (int i = 0; i < EPC; ++i) acc[it][i] = 0.0;
    double wloc = 0.0;
    for (int b0 = 0; b0 < count; b0 += kProfBlock) {
        const int nb = (count - b0) < kProfBlock ? (count - b0) : kProfBlock;
        __syncthreads();
        for (int i = threadIdx.x; i < nb; i += blockDim.x) {
            const int64_t r = (int64_t)col[b0 + i] - row_base;
            const bool mine = r >= 0 && r < n;
            const double wt = w ? (double)w[b0 + i] : 1.0;
            const double nrm = mine ? norm64[r] : 1.0;
            s_row[i] = mine ? (int)r : -1;
            s_w[i] = DIV ? wt : wt / nrm;
            s_n[i] = nrm;
            if (DIV) s_y[i] = 1.0 / nrm;
            wloc += wt;
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < kProfMaxIter; ++it) {
            const int g = it * (int)blockDim.x + (int)threadIdx.x;
            if (g >= chunks || g < chunk_lo || g >= chunk_hi) continue;
            for (int i = 0; i < nb; i += INFLIGHT) {
                uint4 v[INFLIGHT];
                int rr[INFLIGHT];
#pragma unroll
                for (int j = 0; j < INFLIGHT; ++j) {
                    rr[j] = (i + j < nb) ? s_row[i + j] : -1;
                    REBERT_ASSERT(rr[j] < n);
                    if (rr[j] >= 0) v[j] = __ldg((const uint4*)(rows + (size_t)rr[j] * ld) + g);
                }
#pragma unroll
                for (int j = 0; j < INFLIGHT; ++j) {
                    if (rr[j] < 0) continue;
                    double x[EPC];
                    RowChunk<T>::unpack(v[j], x);
                    const double wt = s_w[i + j], nrm = s_n[i + j], y = DIV ? s_y[i + j] : 0.0;
#pragma unroll
                    for (int t = 0; t < EPC; ++t)
                        acc[it][t] = DIV ? fma(wt, div_by_norm(x[t], nrm, y), acc[it][t]) : fma(x[t], wt, acc[it][t]);
                }
            }
        }
    }
#pragma unroll
    for (int it = 0; it < kProfMaxIter; ++it) {
        const int g = it * (int)blockDim.x + (int)threadIdx.x;
        if (g >= chunks || g < chunk_lo || g >= chunk_hi) continue;
#pragma unroll
        for (int t = 0; t < EPC; ++t) sum64[(int64_t)g * EPC + t] = acc[it][t];
    }
    wloc = warp_sum(wloc);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = wloc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double sw = 0.0;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) sw += s_red[i];
        wsum[0] = sw;
    }
}

}  // namespace rebert
