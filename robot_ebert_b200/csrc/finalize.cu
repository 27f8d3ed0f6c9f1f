// Exact pass and merges.
//   rebert_finalize_topk : fp64 re-score of the fast pass's candidates with the oracle's formula
//                          (sklearn cosine_similarity: x/||x||, y/||y||, dot — lib.py:51), then the
//                          (score desc, row asc) order of lib.py:55,63, best k out.
//   rebert_merge_topk    : the same order across per-shard result lists (row-sharded catalogs).
#include "exchange.cuh"

namespace rebert {

constexpr int kFinalThreads = 1024;
constexpr int kMaxKc = 1024;

// grid = (b); one CTA per query.  cand_keys [b, kc], q64 [b, ld], outputs [b, k].
// The fp64 query is staged in shared memory once; every warp re-scores candidates with 16-byte row loads
// (exact_score_row, exact.cuh — the same function the fused single-request kernel uses, so both give the same bits).
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
    __shared__ unsigned long long s_maxerr;      // bits of a non-negative double: integer order == numeric order
    __shared__ int s_tmp[2];
    __shared__ double s_kth;
    const int u = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    pdl_trigger();
    pdl_wait();                                      // the candidates come from the kernel launched just before
    const uint64_t* keys = cand_keys + (size_t)u * kc;
    stage_query_planes(q64 + (size_t)u * ld, ld, EPC, s_q);
    if (threadIdx.x == 0) s_maxerr = 0ull;
    __syncthreads();
    const QueryPlanes qsrc{s_q, ld / EPC};

    for (int c = warp; c < kc; c += nwarps) {
        const uint64_t key = keys[c];
        double sc = -INFINITY;
        int64_t gr = -1;
        if (key != 0) {
            const uint32_t lr = key_row(key);
            sc = exact_score_row<T, DIV>(rows, ld, norm64, lr, qsrc, lane);
            gr = row_base + lr;
        }
        if (lane == 0) {
            s_score[c] = sc;
            s_row[c] = gr;
            if (key != 0) atomicMax(&s_maxerr, (unsigned long long)__double_as_longlong(fabs(sc - (double)key_score(key))));
        }
    }
    __syncthreads();
    // list full => rows outside it have fast score <= key_score(last); the margin compares that against the exact k-th
    // minus 4x the largest fast-vs-exact deviation seen on the candidates themselves, which calibrates whatever
    // rounding the fast pass had (fp32 accumulation, bf16 queries on the tensor-core path)
    const uint64_t last = keys[kc - 1];
    rank_candidates(s_score, s_row, kc, k, last != 0, last ? (double)key_score(last) : 0.0,
                    __longlong_as_double((long long)s_maxerr), /*neartie_matters=*/!DIV, out_rows + (size_t)u * k,
                    out_scores + (size_t)u * k, out_count + u, out_margin ? out_margin + u : nullptr, s_tmp, &s_kth);
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


// Stand-alone forms of the two exchange steps (exchange.cuh).  The single-request kernel runs the result exchange in its
// own tail (gemv_topk.cu); these kernels serve callers that hold a packed local result / a partial profile already.
__global__ void __launch_bounds__(256) exchange_merge_kernel(Exchange x, int k, const unsigned long long* __restrict__ local,
                                                             unsigned long long* __restrict__ out) {
    pdl_trigger();
    pdl_wait();                                      // `local` is written by the kernel launched just before
    exchange_results(x, k, local, out);
}

// sum64 [ld] partial -> exchanged, summed in rank order, divided by wsum: p32 / p64 (profile_finalize fused in).
__global__ void __launch_bounds__(256) profile_exchange_kernel(Exchange x, int ld, double* __restrict__ sum64, const double* __restrict__ wsum,
                                                               float* __restrict__ p32, double* __restrict__ p64) {
    pdl_trigger();
    pdl_wait();
    exchange_profile(x, ld, sum64, sum64);           // every element is read (pushed) before the barrier inside, written after it
    __syncthreads();
    const double ws = wsum[0];
    for (int c = threadIdx.x; c < ld; c += blockDim.x) {
        const double v = ws != 0.0 ? sum64[c] / ws : 0.0;
        p64[c] = v;
        p32[c] = (float)v;
    }
}

int make_exchange(const rebert_exchange_t* ex, int32_t* err_flag, Exchange* out) {
    REBERT_REQUIRE(ex && ex->peer_buffers && err_flag, "exchange: null argument");
    REBERT_REQUIRE(ex->world > 0 && ex->world <= kMaxPeers && ex->rank >= 0 && ex->rank < ex->world, "exchange: world=%d rank=%d",
                   ex->world, ex->rank);
    REBERT_REQUIRE(ex->k_max > 0 && ex->prof_len >= 0 && ex->channels > 0 && ex->channel >= 0 && ex->channel < ex->channels && ex->seq != 0,
                   "exchange: k_max=%d prof_len=%d channel=%d/%d seq=%u", ex->k_max, ex->prof_len, ex->channel, ex->channels, ex->seq);
    memset(out, 0, sizeof(*out));
    out->world = ex->world;
    out->rank = ex->rank;
    out->words_cap = 2 * ex->k_max + 2;
    out->prof_cap = ex->prof_len + 1;
    out->seq = ex->seq;
    // ~10 s of SM clocks at 2 GHz: only a dead peer gets here.  (A constant: querying the clock rate is a slow driver
    // call and this function sits on the per-request path.)
    out->timeout_cycles = 20000000000ll;
    out->err = err_flag;
    const size_t ch_words = xchg_channel_words(ex->world, out->words_cap, out->prof_cap);
    for (int i = 0; i < ex->world; ++i)
        out->peer[i] = (unsigned long long*)(uintptr_t)ex->peer_buffers[i] + (size_t)ex->channel * ch_words;
    return REBERT_OK;
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

REBERT_API size_t rebert_exchange_buffer_bytes(int32_t world, int32_t k_max, int32_t prof_len, int32_t channels) {
    if (world <= 0 || world > kMaxPeers || k_max <= 0 || prof_len < 0 || channels <= 0) return 0;
    return xchg_channel_words(world, 2 * k_max + 2, prof_len + 1) * 8 * (size_t)channels;
}

REBERT_API int rebert_exchange_merge(const rebert_exchange_t* ex, int32_t k, const int64_t* local_packed, int64_t* out_packed,
                                     int32_t* err_flag, rebert_stream stream) {
    REBERT_REQUIRE(local_packed && out_packed, "exchange_merge: null argument");
    Exchange x;
    int rc = make_exchange(ex, err_flag, &x);
    if (rc != REBERT_OK) return rc;
    REBERT_REQUIRE(k > 0 && k <= ex->k_max, "exchange_merge: k=%d k_max=%d", k, ex->k_max);
    REBERT_CUDA(launch_pdl(exchange_merge_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, x, k,
                           (const unsigned long long*)local_packed, (unsigned long long*)out_packed));
    REBERT_CUDA(cudaGetLastError());
    return REBERT_OK;
}

REBERT_API int rebert_profile_exchange(const rebert_exchange_t* ex, int32_t ld, double* sum64, const double* wsum, float* p32,
                                       double* p64, int32_t* err_flag, rebert_stream stream) {
    REBERT_REQUIRE(sum64 && wsum && p32 && p64 && ld > 0, "profile_exchange: bad arguments");
    Exchange x;
    int rc = make_exchange(ex, err_flag, &x);
    if (rc != REBERT_OK) return rc;
    REBERT_REQUIRE(ld <= ex->prof_len, "profile_exchange: ld=%d exceeds the buffer's prof_len=%d", ld, ex->prof_len);
    REBERT_CUDA(launch_pdl(profile_exchange_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, x, ld, sum64, wsum, p32, p64));
    REBERT_CUDA(cudaGetLastError());
    return REBERT_OK;
}

}  // extern "C"
