// Exact (fp64) re-score of one catalog row and the (score desc, row asc) ranking of a candidate list — shared by the
// standalone exact pass (finalize.cu), the fused tail of the single-query kernel (gemv_topk.cu) and the one-CTA kernel
// for tiny catalogs (small.cu), so that every route produces bit-identical scores for the same row and query.
//
// The formula is the oracle's (sklearn cosine_similarity, lib.py:51): x / ||x||, y / ||y||, dot.  DIV = true keeps
// sklearn's operation order (normalize() divides every element by the row norm BEFORE the dot), which is what makes rows
// that tie exactly in the float64 reference (d = 1, scaled one-hot rows) tie here as well.
#pragma once

#include "common.cuh"

namespace rebert {

__device__ __forceinline__ bool better(double sa, int64_t ra, double sb, int64_t rb) {
    return sa > sb || (sa == sb && ra < rb);
}

// x / nrm, correctly rounded, from the row's reciprocal y = RN(1 / nrm) computed ONCE per row: q0 = x y; two Markstein
// corrections q <- q + (x - nrm q) y, each residual exact by FMA.  The second correction of a faithful quotient with a
// correctly rounded reciprocal yields RN(x / nrm) (Markstein 1990); tests/test_exact_division.py checks the sequence against
// IEEE division on 10^7 operand pairs per run (6 x 10^9 during development, fp32 / bf16 numerators, mantissas of all ones
// and near powers of two included: no mismatch).  Five FMA-class operations instead of the ~20 of a full fp64 division —
// the exact pass is fp64-latency-bound, and this is what sklearn's divide-first order costs per element.
__device__ __forceinline__ double div_by_norm(double x, double nrm, double y) {
    double q = x * y;
    q = fma(fma(-nrm, q, x), y, q);
    q = fma(fma(-nrm, q, x), y, q);
    return q;
}
template <bool DIV> __device__ __forceinline__ double unit(double x, double nrm, double y) { return DIV ? div_by_norm(x, nrm, y) : x; }

// 16-byte chunk of a stored row times the matching query slice.  q[p] holds query elements 2p, 2p+1 of the chunk.
template <typename T> struct ChunkDot;
template <> struct ChunkDot<float> {
    static constexpr int EPC = 4;
    static constexpr int PAIRS = 2;
    template <bool DIV>
    __device__ static __forceinline__ double dot(const uint4& v, const double2* q, double acc, double nrm, double y) {
        acc = fma(q[0].x, unit<DIV>((double)__uint_as_float(v.x), nrm, y), acc);
        acc = fma(q[0].y, unit<DIV>((double)__uint_as_float(v.y), nrm, y), acc);
        acc = fma(q[1].x, unit<DIV>((double)__uint_as_float(v.z), nrm, y), acc);
        acc = fma(q[1].y, unit<DIV>((double)__uint_as_float(v.w), nrm, y), acc);
        return acc;
    }
};
template <> struct ChunkDot<__nv_bfloat16> {
    static constexpr int EPC = 8;
    static constexpr int PAIRS = 4;
    template <bool DIV>
    __device__ static __forceinline__ double dot(const uint4& v, const double2* q, double acc, double nrm, double y) {
        acc = fma(q[0].x, unit<DIV>((double)bf16lo(v.x), nrm, y), acc);
        acc = fma(q[0].y, unit<DIV>((double)bf16hi(v.x), nrm, y), acc);
        acc = fma(q[1].x, unit<DIV>((double)bf16lo(v.y), nrm, y), acc);
        acc = fma(q[1].y, unit<DIV>((double)bf16hi(v.y), nrm, y), acc);
        acc = fma(q[2].x, unit<DIV>((double)bf16lo(v.z), nrm, y), acc);
        acc = fma(q[2].y, unit<DIV>((double)bf16hi(v.z), nrm, y), acc);
        acc = fma(q[3].x, unit<DIV>((double)bf16lo(v.w), nrm, y), acc);
        acc = fma(q[3].y, unit<DIV>((double)bf16hi(v.w), nrm, y), acc);
        return acc;
    }
};

// Where the fp64 query comes from.  "Pair planes" in shared memory: element e of chunk ch lives at
// ((e >> 1) * chunks + ch) * 2 + (e & 1), so the 32 lanes of a warp (consecutive chunks) read consecutive 16-byte
// double2's — conflict-free LDS.128.  "Global": the natural [ld] layout, read through the read-only path.
struct QueryPlanes {
    const double* sq;
    int chunks;
    template <int PAIRS> __device__ __forceinline__ void load(int ch, double2* q) const {
#pragma unroll
        for (int p = 0; p < PAIRS; ++p) q[p] = *(const double2*)(sq + ((size_t)p * chunks + ch) * 2);
    }
};
struct QueryGlobal {
    const double* q64;
    template <int PAIRS> __device__ __forceinline__ void load(int ch, double2* q) const {
#pragma unroll
        for (int p = 0; p < PAIRS; ++p) q[p] = __ldg((const double2*)(q64 + (size_t)ch * (2 * PAIRS)) + p);
    }
};
// q [ld] (global) -> pair planes in shared memory.  Eight loads per thread are issued before the first store: a plain
// load-store loop would pay one L2 round trip per iteration.
__device__ __forceinline__ void stage_query_planes(const double* __restrict__ q, int ld, int epc, double* s_q) {
    const int chunks = ld / epc;
    for (int i0 = threadIdx.x; i0 < ld; i0 += 8 * blockDim.x) {
        double v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = i0 + j * blockDim.x;
            v[j] = i < ld ? __ldg(q + i) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = i0 + j * blockDim.x;
            if (i < ld) {
                const int ch = i / epc, e = i - ch * epc;
                s_q[((size_t)(e >> 1) * chunks + ch) * 2 + (e & 1)] = v[j];
            }
        }
    }
}

// One warp, one row: every lane returns the score.  The summation order is fixed — lane l owns chunks l, l + 32, l + 64, ...;
// even-numbered ones accumulate into a0, odd-numbered ones into a1, in ascending order; then a0 + a1 and an xor-shuffle
// tree — so the result depends on nothing but the row and the query.  All of a lane's row chunks (up to NB at a time) are
// requested before the first is used: candidate rows come from HBM (the streaming pass evicts them from L2 first), and one
// round trip per row instead of one per pair of chunks is what makes the exact pass cheap.
template <typename T, bool DIV, typename Q>
__device__ __forceinline__ double exact_score_row(const T* __restrict__ rows, int ld, const double* __restrict__ norm64,
                                                  uint32_t local_row, const Q& qsrc, int lane) {
    constexpr int EPC = ChunkDot<T>::EPC;
    constexpr int PAIRS = ChunkDot<T>::PAIRS;
    constexpr int NB = 6;                                // 6 x 32 chunks: a 1536-element bf16 row in one round, fp32 in two
    const int chunks = ld / EPC;
    const uint4* row = (const uint4*)(rows + (size_t)local_row * ld);
    const double nrm = __ldg(norm64 + local_row);
    const double y = DIV ? 1.0 / nrm : 0.0;              // ONE correctly rounded reciprocal per row (div_by_norm)
    double a0 = 0.0, a1 = 0.0;
    for (int base = lane; base < chunks; base += 32 * NB) {
        uint4 v[NB];
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int ch = base + 32 * j;
            if (ch < chunks) v[j] = __ldg(row + ch);
        }
#pragma unroll
        for (int j = 0; j < NB; ++j) {
            const int ch = base + 32 * j;
            if (ch < chunks) {
                double2 q[PAIRS];
                qsrc.template load<PAIRS>(ch, q);
                if (j & 1) a1 = ChunkDot<T>::template dot<DIV>(v[j], q, a1, nrm, y);
                else       a0 = ChunkDot<T>::template dot<DIV>(v[j], q, a0, nrm, y);
            }
        }
    }
    return DIV ? warp_sum(a0 + a1) : warp_sum(a0 + a1) / nrm;
}

// Runtime-dtype front end (the catalog of record is fp32 or bf16), sklearn operation order.
template <typename Q>
__device__ __forceinline__ double exact_score_row_rt(const void* rows, int dtype, int ld, const double* norm64, uint32_t local_row,
                                                     const Q& qsrc, int lane) {
    if (dtype == REBERT_F32) return exact_score_row<float, true>((const float*)rows, ld, norm64, local_row, qsrc, lane);
    return exact_score_row<__nv_bfloat16, true>((const __nv_bfloat16*)rows, ld, norm64, local_row, qsrc, lane);
}

// fp64 score -> 64-bit key whose unsigned order is the numeric order (and back).  The two zeros compare equal as numbers, so
// -0.0 is folded into +0.0 first; key 0 (the image of one NaN pattern no score takes) marks an empty slot.
__device__ __forceinline__ unsigned long long f64_orderable(double x) {
    x += 0.0;
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double f64_from_orderable(unsigned long long o) {
    const unsigned long long b = (o >> 63) ? (o & 0x7FFFFFFFFFFFFFFFull) : ~o;
    return __longlong_as_double((long long)b);
}

// Rank `kc` re-scored candidates (s_row[c] < 0 = empty slot) by (score desc, row asc): the best k go to o_rows / o_scores
// (-1 / -inf padded), their number to *o_count.  fast_last = fast score of the worst kept candidate when the fast pass's
// list was full (list_full): *o_margin = exact k-th - fast_last - err_mult (4) x max|fast - exact| proves the id set when it exceeds the
// fast pass's error bound; +inf when the list was not full (every allowed row is a candidate), -inf when two distinct
// scores at or above the k-th place sit within 4 ulp and the caller computed them divide-after (NEARTIE).
// Every thread of the CTA must call; outputs may be shared, global or device-addressable pinned host memory.
// s_score is CONSUMED: the scores are replaced in place by their orderable keys, so that the kc x kc comparison pass is integer
// compares on one shared-memory word per pair.  (Comparing doubles and rows pair by pair made every iteration a chain of two
// dependent shared loads and fp64 predicates: 22 us for 256 candidates on the request path, as long as the rest of the
// exact-pass kernel together; profiles/r02_gemv_trace_1p25M.txt.)  Rows are consulted only for exact score ties.
template <bool NEARTIE>
__device__ __forceinline__ void rank_candidates(double* s_score, const int64_t* s_row, int kc, int k, bool list_full, double fast_last,
                                                double maxerr, int64_t* o_rows, double* o_scores, int32_t* o_count,
                                                double* o_margin, int* s_tmp /* 2 ints of shared scratch, s_tmp[0] = s_tmp[1] = 0 on entry */,
                                                double* s_kth /* 1 double of shared scratch */, double err_mult = 4.0) {
    unsigned long long* s_ord = (unsigned long long*)s_score;
    for (int c = threadIdx.x; c < kc; c += blockDim.x) s_ord[c] = s_row[c] < 0 ? 0ull : f64_orderable(s_score[c]);   // own entries only
    __syncthreads();
    int valid_mine = 0;
    for (int c = threadIdx.x; c < kc; c += blockDim.x) {
        const unsigned long long oc = s_ord[c];
        if (oc == 0ull) continue;
        ++valid_mine;
        int gt = 0, eq = 0;
        bool near = false;
#pragma unroll 8
        for (int j = 0; j < kc; ++j) {
            const unsigned long long oj = s_ord[j];
            gt += oj > oc;
            eq += oj == oc;
            if (NEARTIE) {
                const unsigned long long dist = oj > oc ? oj - oc : oc - oj;      // distance in representable doubles
                near |= (dist - 1ull) < 4ull;                                      // distinct scores within 4 ulp
            }
        }
        const int64_t r = s_row[c];
        int rank = gt;
        if (eq > 1)                                                                // exact ties (itself included): ascending row
            for (int j = 0; j < kc; ++j) rank += (s_ord[j] == oc && s_row[j] < r);
        if (NEARTIE && near && rank <= k) s_tmp[1] = 1;    // a divide-after formula cannot be trusted to order these
        REBERT_ASSERT(rank >= 0 && rank < kc);
        if (rank < k) {
            const double sc = f64_from_orderable(oc);
            o_rows[rank] = r;
            o_scores[rank] = sc;
            if (rank == k - 1) *s_kth = sc;
        }
    }
    if (valid_mine) atomicAdd(&s_tmp[0], valid_mine);
    __syncthreads();
    const int valid = s_tmp[0];
    const int nout = valid < k ? valid : k;
    for (int i = nout + threadIdx.x; i < k; i += blockDim.x) {
        o_rows[i] = -1;
        o_scores[i] = -INFINITY;
    }
    if (threadIdx.x == 0) {
        *o_count = nout;
        if (o_margin) {
            double mg = (list_full && valid >= k) ? *s_kth - fast_last - err_mult * maxerr : INFINITY;
            if (NEARTIE && s_tmp[1]) mg = -INFINITY;
            *o_margin = mg;
        }
    }
}

}  // namespace rebert
