// Exact pass and merges.
//   rebert_finalize_topk : fp64 re-score of the fast pass's candidates with the oracle's formula
//                          (sklearn cosine_similarity: x/||x||, y/||y||, dot — lib.py:51), then the
//                          (score desc, row asc) order of lib.py:55,63, best k out.
//   rebert_merge_topk    : the same order across per-shard result lists (row-sharded catalogs).
#include "common.cuh"

namespace rebert {

constexpr int kFinalThreads = 1024;
constexpr int kMaxKc = 1024;

__device__ __forceinline__ bool better(double sa, int64_t ra, double sb, int64_t rb) {
    return sa > sb || (sa == sb && ra < rb);
}

// 16-byte chunk of a stored row -> fp64 dot with the matching slice of the query.  The query sits in shared memory in
// "pair planes": element e of chunk ch lives at ((e >> 1) * chunks + ch) * 2 + (e & 1), so the 32 lanes of a warp
// (consecutive chunks) read consecutive 16-byte double2's — conflict-free LDS.128.
__device__ __forceinline__ double2 q_pair(const double* sq, int chunks, int ch, int p) {
    return *(const double2*)(sq + ((size_t)p * chunks + ch) * 2);
}
template <typename T> struct ChunkDot;
// DIV = true follows sklearn's operation order exactly (normalize() divides every element by the row norm, then the
// dot): needed so that rows which tie EXACTLY in the float64 reference (d = 1, scaled one-hot rows) tie here as well.
template <bool DIV> __device__ __forceinline__ double unit(double x, double nrm) { return DIV ? x / nrm : x; }
template <> struct ChunkDot<float> {
    static constexpr int EPC = 4;
    template <bool DIV>
    __device__ static __forceinline__ double dot(const uint4& v, const double* sq, int chunks, int ch, double acc, double nrm) {
        const double2 q0 = q_pair(sq, chunks, ch, 0), q1 = q_pair(sq, chunks, ch, 1);
        acc = fma(q0.x, unit<DIV>((double)__uint_as_float(v.x), nrm), acc);
        acc = fma(q0.y, unit<DIV>((double)__uint_as_float(v.y), nrm), acc);
        acc = fma(q1.x, unit<DIV>((double)__uint_as_float(v.z), nrm), acc);
        acc = fma(q1.y, unit<DIV>((double)__uint_as_float(v.w), nrm), acc);
        return acc;
    }
};
template <> struct ChunkDot<__nv_bfloat16> {
    static constexpr int EPC = 8;
    template <bool DIV>
    __device__ static __forceinline__ double dot(const uint4& v, const double* sq, int chunks, int ch, double acc, double nrm) {
        const double2 q0 = q_pair(sq, chunks, ch, 0), q1 = q_pair(sq, chunks, ch, 1);
        const double2 q2 = q_pair(sq, chunks, ch, 2), q3 = q_pair(sq, chunks, ch, 3);
        acc = fma(q0.x, unit<DIV>((double)bf16lo(v.x), nrm), acc);
        acc = fma(q0.y, unit<DIV>((double)bf16hi(v.x), nrm), acc);
        acc = fma(q1.x, unit<DIV>((double)bf16lo(v.y), nrm), acc);
        acc = fma(q1.y, unit<DIV>((double)bf16hi(v.y), nrm), acc);
        acc = fma(q2.x, unit<DIV>((double)bf16lo(v.z), nrm), acc);
        acc = fma(q2.y, unit<DIV>((double)bf16hi(v.z), nrm), acc);
        acc = fma(q3.x, unit<DIV>((double)bf16lo(v.w), nrm), acc);
        acc = fma(q3.y, unit<DIV>((double)bf16hi(v.w), nrm), acc);
        return acc;
    }
};

// grid = (b); one CTA per query.  cand_keys [b, kc], q64 [b, ld], outputs [b, k].
// The fp64 query is staged in shared memory once; every warp re-scores candidates with 16-byte row loads.
template <typename T, bool DIV>
__global__ void __launch_bounds__(kFinalThreads, 1)
finalize_topk_kernel(const T* __restrict__ rows, const double* __restrict__ norm64, int64_t row_base, int ld,
                     const double* __restrict__ q64, const uint64_t* __restrict__ cand_keys, int kc, int k,
                     int64_t* __restrict__ out_rows, double* __restrict__ out_scores, int32_t* __restrict__ out_count,
                     double* __restrict__ out_margin) {
    constexpr int EPC = ChunkDot<T>::EPC;
    extern __shared__ __align__(16) double s_q[];   // [ld]
    __shared__ double s_score[kMaxKc];
    __shared__ int64_t s_row[kMaxKc];
    __shared__ int s_valid;
    __shared__ double s_kth;
    __shared__ unsigned long long s_maxerr;      // bits of a non-negative double: integer order == numeric order
    __shared__ int s_neartie;
    const int u = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    pdl_trigger();
    pdl_wait();                                      // the candidates come from the kernel launched just before
    const uint64_t* keys = cand_keys + (size_t)u * kc;
    const double* q = q64 + (size_t)u * ld;
    const int chunks = ld / EPC;
    for (int i = threadIdx.x; i < ld; i += blockDim.x) {
        const int ch = i / EPC, e = i - ch * EPC;
        s_q[((size_t)(e >> 1) * chunks + ch) * 2 + (e & 1)] = q[i];
    }
    if (threadIdx.x == 0) { s_valid = 0; s_kth = 0.0; s_maxerr = 0ull; s_neartie = 0; }
    __syncthreads();

    for (int c = warp; c < kc; c += nwarps) {
        const uint64_t key = keys[c];
        double sc = -INFINITY;
        int64_t gr = -1;
        if (key != 0) {
            const uint32_t lr = key_row(key);
            const uint4* row = (const uint4*)(rows + (size_t)lr * ld);
            const double nrm = norm64[lr];
            double a0 = 0.0, a1 = 0.0;
            int ch = lane;
            for (; ch + 32 < chunks; ch += 64) {                 // two independent 16-byte loads in flight per lane
                const uint4 v0 = __ldg(row + ch), v1 = __ldg(row + ch + 32);
                a0 = ChunkDot<T>::template dot<DIV>(v0, s_q, chunks, ch, a0, nrm);
                a1 = ChunkDot<T>::template dot<DIV>(v1, s_q, chunks, ch + 32, a1, nrm);
            }
            if (ch < chunks) a0 = ChunkDot<T>::template dot<DIV>(__ldg(row + ch), s_q, chunks, ch, a0, nrm);
            sc = DIV ? warp_sum(a0 + a1) : warp_sum(a0 + a1) / nrm;
            gr = row_base + lr;
        }
        if (lane == 0) {
            s_score[c] = sc;
            s_row[c] = gr;
            if (key != 0) {
                atomicAdd(&s_valid, 1);
                atomicMax(&s_maxerr, (unsigned long long)__double_as_longlong(fabs(sc - (double)key_score(key))));
            }
        }
    }
    __syncthreads();
    const int valid = s_valid;
    const int nout = valid < k ? valid : k;
    for (int c = threadIdx.x; c < kc; c += blockDim.x) {
        const int64_t r = s_row[c];
        if (r < 0) continue;
        const double sc = s_score[c];
        int rank = 0;
        bool near = false;
        for (int j = 0; j < kc; ++j) {
            const int64_t rj = s_row[j];
            if (rj < 0) continue;
            const double sj = s_score[j];
            if (better(sj, rj, sc, r)) ++rank;
            const double gap = fabs(sj - sc);
            near |= gap != 0.0 && gap <= 8.9e-16 * fmax(fabs(sc), 1e-300);   // distinct scores within 4 ulp
        }
        if (!DIV && near && rank <= k) s_neartie = 1;    // the divide-after formula cannot be trusted to order these
        if (rank < k) {
            out_rows[(size_t)u * k + rank] = r;
            out_scores[(size_t)u * k + rank] = sc;
            if (rank == k - 1) s_kth = sc;
        }
    }
    for (int i = nout + threadIdx.x; i < k; i += blockDim.x) {
        out_rows[(size_t)u * k + i] = -1;
        out_scores[(size_t)u * k + i] = -INFINITY;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        out_count[u] = nout;
        if (out_margin) {
            const uint64_t last = keys[kc - 1];
            // list full => rows outside it have fast score <= key_score(last); compare against the exact k-th
            // minus 4x the largest fast-vs-exact deviation seen on the candidates themselves: calibrates whatever
            // rounding the fast pass had (fp32 accumulation, bf16 queries on the tensor-core path)
            const double maxerr = __longlong_as_double((long long)s_maxerr);
            double mg = (last != 0 && valid >= k) ? s_kth - (double)key_score(last) - 4.0 * maxerr : INFINITY;
            if (s_neartie) mg = -INFINITY;               // caller re-runs this query through the divide-first exact pass
            out_margin[u] = mg;
        }
    }
}

// One CTA per query: rank-merge `lists` sorted lists of up to k entries each.  List l of query u lives at
// rows[l * rows_stride + u * k + e], scores[l * scores_stride + u * k + e], counts[l * counts_stride + u]
// (strides in elements), which covers both a dense [lists, b, k] layout and the packed per-rank buffers an
// all-gather produces.
__global__ void merge_topk_kernel(const int64_t* __restrict__ rows, const double* __restrict__ scores,
                                  const int32_t* __restrict__ counts, int64_t rows_stride, int64_t scores_stride,
                                  int64_t counts_stride, int lists, int k, int64_t* __restrict__ out_rows,
                                  double* __restrict__ out_scores, int32_t* __restrict__ out_count) {
    const int u = blockIdx.x;
    int total = 0;
    for (int l = 0; l < lists; ++l) total += min(counts[l * counts_stride + u], k);
    const int nout = total < k ? total : k;
    for (int i = threadIdx.x; i < lists * k; i += blockDim.x) {
        const int l = i / k, e = i - l * k;
        if (e >= counts[l * counts_stride + u]) continue;
        const double sc = scores[l * scores_stride + (int64_t)u * k + e];
        const int64_t r = rows[l * rows_stride + (int64_t)u * k + e];
        int rank = e;
        for (int o = 0; o < lists; ++o) {
            if (o == l) continue;
            const int64_t* orow = rows + o * rows_stride + (int64_t)u * k;
            const double* osc = scores + o * scores_stride + (int64_t)u * k;
            int lo = 0, hi = min(counts[o * counts_stride + u], k);      // entries of list o better than (sc, r)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (better(osc[mid], orow[mid], sc, r)) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            out_rows[(size_t)u * k + rank] = r;
            out_scores[(size_t)u * k + rank] = sc;
        }
    }
    for (int i = nout + threadIdx.x; i < k; i += blockDim.x) {
        out_rows[(size_t)u * k + i] = -1;
        out_scores[(size_t)u * k + i] = -INFINITY;
    }
    if (threadIdx.x == 0) out_count[u] = nout;
}


// ---------------------------------------------------------------------------------------------------------
// Fused exchange + merge over NVLink peer memory (row-sharded single-query path; replaces all-gather + merge).
// Every rank's symmetric buffer holds, for both parities of the call sequence number,
//     gather[parity][rank][words_cap]  packed results (rows k | fp64 scores k | count | margin)
//     flag  [parity][rank]            the sequence number of the last result that rank delivered
// One CTA per rank: (1) store my packed result straight into every peer's gather slot (P2P stores through
// NVLink/NVSwitch), (2) system-scope fence, then publish my flag in every peer, (3) acquire-spin on my own flags until
// all ranks delivered, (4) rank-merge the G lists under (score desc, row asc).  Parity double-buffering is enough:
// a rank can be at most one call ahead of a peer, because it needs that peer's flag of the current call to finish it.
// ---------------------------------------------------------------------------------------------------------
constexpr int kMaxPeers = 16;
struct PeerBufs { unsigned long long* p[kMaxPeers]; };

__device__ __forceinline__ size_t xchg_gather_off(int parity, int rank, int world, int words_cap) {
    return ((size_t)parity * world + rank) * words_cap;
}
__device__ __forceinline__ size_t xchg_flag_off(int parity, int rank, int world, int words_cap) {
    return (size_t)2 * world * words_cap + (size_t)parity * world + rank;
}

__global__ void __launch_bounds__(256) exchange_merge_kernel(PeerBufs peers, int world, int my_rank, int k, int words_cap,
                                                             unsigned seq, const unsigned long long* __restrict__ local,
                                                             unsigned long long* __restrict__ out, int* __restrict__ err,
                                                             long long timeout_cycles) {
    const int words = 2 * k + 2;
    const int parity = (int)(seq & 1u);
    pdl_trigger();
    pdl_wait();                                      // `local` is written by the exact pass launched just before
    unsigned long long* mine = peers.p[my_rank];
    // (1) deliver my result to every rank (myself included)
    for (int i = threadIdx.x; i < world * words; i += blockDim.x) {
        const int pr = i / words, w = i - pr * words;
        peers.p[pr][xchg_gather_off(parity, my_rank, world, words_cap) + w] = local[w];
    }
    __threadfence_system();
    __syncthreads();
    // (2) publish, (3) wait for everybody
    if (threadIdx.x < world) {
        unsigned long long* f = peers.p[threadIdx.x] + xchg_flag_off(parity, my_rank, world, words_cap);
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"((unsigned long long)seq) : "memory");
        const unsigned long long* w = mine + xchg_flag_off(parity, threadIdx.x, world, words_cap);
        const long long t0 = clock64();
        unsigned long long v;
        while (true) {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory");
            if (v == (unsigned long long)seq) break;
            if (clock64() - t0 > timeout_cycles) { *(volatile int*)err = 1 + threadIdx.x; break; }   // plain store: err may live in pinned host memory
            __nanosleep(64);
        }
    }
    __syncthreads();
    // (4) merge the `world` sorted lists
    const unsigned long long* g = mine + xchg_gather_off(parity, 0, world, words_cap);
    int total = 0;
    double margin = INFINITY;
    for (int l = 0; l < world; ++l) {
        const unsigned long long* L = g + (size_t)l * words_cap;
        total += min((int)(unsigned)L[2 * k], k);
        margin = fmin(margin, __longlong_as_double((long long)L[2 * k + 1]));
    }
    const int nout = total < k ? total : k;
    for (int i = threadIdx.x; i < world * k; i += blockDim.x) {
        const int l = i / k, e = i - l * k;
        const unsigned long long* L = g + (size_t)l * words_cap;
        if (e >= (int)(unsigned)L[2 * k]) continue;
        const int64_t r = (int64_t)L[e];
        const double sc = __longlong_as_double((long long)L[k + e]);
        int rank = e;
        for (int o = 0; o < world; ++o) {
            if (o == l) continue;
            const unsigned long long* O = g + (size_t)o * words_cap;
            int lo = 0, hi = min((int)(unsigned)O[2 * k], k);
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (better(__longlong_as_double((long long)O[k + mid]), (int64_t)O[mid], sc, r)) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k) {
            out[rank] = (unsigned long long)r;
            out[k + rank] = (unsigned long long)__double_as_longlong(sc);
        }
    }
    for (int i = nout + threadIdx.x; i < k; i += blockDim.x) {
        out[i] = (unsigned long long)(-1ll);
        out[k + i] = (unsigned long long)__double_as_longlong(-INFINITY);
    }
    if (threadIdx.x == 0) {
        out[2 * k] = (unsigned long long)(unsigned)nout;
        out[2 * k + 1] = (unsigned long long)__double_as_longlong(margin);
    }
}

template <typename T, bool DIV>
static int finalize_launch_t(const rebert_catalog_t* cat, const double* q64, const uint64_t* cand_keys, int b, int kc, int k,
                             int64_t* out_rows, double* out_scores, int32_t* out_count, double* out_margin, cudaStream_t st) {
    const size_t smem = (size_t)cat->ld * sizeof(double);
    auto kern = finalize_topk_kernel<T, DIV>;
    { int rc = raise_smem_limit(kern); if (rc != REBERT_OK) return rc; }
    REBERT_CUDA(launch_pdl(kern, dim3(b), dim3(kFinalThreads), smem, st, (const T*)cat->rows, cat->norm64, cat->row_base, cat->ld, q64,
                           cand_keys, kc, k, out_rows, out_scores, out_count, out_margin));
    REBERT_CUDA(cudaGetLastError());
    return REBERT_OK;
}

// exact_order = true: sklearn's operation order (element-wise division first) — the single-request path.
// exact_order = false: one division after the dot (within 1 ulp), near-ties flagged through the margin — the batched path.
int finalize_launch(const rebert_catalog_t* cat, const double* q64, const uint64_t* cand_keys, int b, int kc, int k,
                    int64_t* out_rows, double* out_scores, int32_t* out_count, double* out_margin, cudaStream_t st,
                    bool exact_order) {
    if (cat->dtype == REBERT_F32)
        return exact_order ? finalize_launch_t<float, true>(cat, q64, cand_keys, b, kc, k, out_rows, out_scores, out_count, out_margin, st)
                           : finalize_launch_t<float, false>(cat, q64, cand_keys, b, kc, k, out_rows, out_scores, out_count, out_margin, st);
    return exact_order ? finalize_launch_t<__nv_bfloat16, true>(cat, q64, cand_keys, b, kc, k, out_rows, out_scores, out_count, out_margin, st)
                       : finalize_launch_t<__nv_bfloat16, false>(cat, q64, cand_keys, b, kc, k, out_rows, out_scores, out_count, out_margin, st);
}

}  // namespace rebert

using namespace rebert;

extern "C" {

REBERT_API int rebert_finalize_topk(const rebert_catalog_t* cat, const double* qn64, const uint64_t* cand_keys, int32_t kc,
                         int32_t k, int64_t* out_rows, double* out_scores, int32_t* out_count, double* out_margin,
                         rebert_stream stream) {
    REBERT_REQUIRE(cat && cat->rows && cat->norm64 && qn64 && cand_keys && out_rows && out_scores && out_count,
                   "finalize_topk: null argument");
    REBERT_REQUIRE(kc > 0 && kc <= kMaxKc && k > 0 && k <= kc, "finalize_topk: k=%d kc=%d", k, kc);
    return finalize_launch(cat, qn64, cand_keys, 1, kc, k, out_rows, out_scores, out_count, out_margin,
                           (cudaStream_t)stream, /*exact_order=*/true);
}

REBERT_API int rebert_merge_topk(const int64_t* rows, const double* scores, const int32_t* counts, int64_t rows_stride,
                      int64_t scores_stride, int64_t counts_stride, int32_t lists, int32_t b, int32_t k,
                      int64_t* out_rows, double* out_scores, int32_t* out_count, rebert_stream stream) {
    REBERT_REQUIRE(rows && scores && counts && out_rows && out_scores && out_count, "merge_topk: null argument");
    REBERT_REQUIRE(lists > 0 && b > 0 && k > 0, "merge_topk: lists=%d b=%d k=%d", lists, b, k);
    merge_topk_kernel<<<b, 256, 0, (cudaStream_t)stream>>>(rows, scores, counts, rows_stride, scores_stride, counts_stride,
                                                           lists, k, out_rows, out_scores, out_count);
    REBERT_CUDA(cudaGetLastError());
    return REBERT_OK;
}

REBERT_API size_t rebert_exchange_buffer_bytes(int32_t world, int32_t k_max) {
    if (world <= 0 || world > kMaxPeers || k_max <= 0) return 0;
    const size_t words_cap = (size_t)2 * k_max + 2;
    return ((size_t)2 * world * words_cap + (size_t)2 * world) * 8;
}

REBERT_API int rebert_exchange_merge(const uint64_t* peer_buffers, int32_t world, int32_t rank, int32_t k, int32_t k_max,
                                     uint32_t seq, const int64_t* local_packed, int64_t* out_packed, int32_t* err_flag,
                                     rebert_stream stream) {
    REBERT_REQUIRE(peer_buffers && local_packed && out_packed && err_flag, "exchange_merge: null argument");
    REBERT_REQUIRE(world > 0 && world <= kMaxPeers && rank >= 0 && rank < world, "exchange_merge: world=%d rank=%d", world, rank);
    REBERT_REQUIRE(k > 0 && k <= k_max && seq != 0, "exchange_merge: k=%d k_max=%d seq=%u", k, k_max, seq);
    PeerBufs pb;
    memset(&pb, 0, sizeof(pb));
    for (int i = 0; i < world; ++i) pb.p[i] = (unsigned long long*)(uintptr_t)peer_buffers[i];
    // ~10 s of SM clocks at 2 GHz: only a dead peer gets here.  (A constant: querying the clock rate is a slow driver
    // call and this function sits on the per-request path.)
    const long long timeout_cycles = 20000000000ll;
    REBERT_CUDA(launch_pdl(exchange_merge_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, pb, world, rank, k, 2 * k_max + 2, seq,
                           (const unsigned long long*)local_packed, (unsigned long long*)out_packed, err_flag, timeout_cycles));
    REBERT_CUDA(cudaGetLastError());
    return REBERT_OK;
}

}  // extern "C"
