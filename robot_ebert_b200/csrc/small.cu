// One-kernel request for catalogs small enough for a single CTA — the reference's production shape
// (movies-collab: 2 269 movies x 32 ALS factors, notebooks/create-embeddings.ipynb:232,:1241; lib.py:43-55).
//
// At this size a request is pure latency: streaming 290 KB takes no time, launches and PCIe round trips are everything.
// So the whole of lib.py:43-55 is ONE launch of ONE CTA: fetch the request zero-copy from the caller's pinned block,
// normalise the query or build the profile (mean of the liked rows' unit vectors), score EVERY allowed row in fp64 with the
// oracle's formula (exact.cuh — no fast pass, so nothing to prove: the result is exact by construction), select the top k
// under (score desc, row asc) and write the packed result straight into pinned host memory.
#include "exchange.cuh"
#include "request.cuh"

namespace rebert {

constexpr int kSmallThreads = 512;
constexpr int kSmallMaxElems = 131072;      // n * ld
constexpr int kSmallMaxRows = 8192;
constexpr int kSmallMaxGather = 2048;       // rows at or above the pruning threshold the fast selection can rank

// dynamic shared memory: s_q [ld] f64 | s_sc [n2] f64 | s_sum [ld] f64 | s_excl [ne4] i32 | s_src [d4] f32 | s_gsc [G] f64 | s_grow [G] i32
__host__ __device__ inline size_t small_smem_bytes(int n, int d, int ld, int n_excl) {
    const size_t n2 = (size_t)(n + 1) & ~(size_t)1, ne4 = (size_t)(n_excl + 3) & ~(size_t)3, d4 = (size_t)(d + 3) & ~(size_t)3;
    return (size_t)ld * 16 + n2 * 8 + ne4 * 4 + d4 * 4 + (size_t)kSmallMaxGather * 12;
}

struct SmallParams {
    const void*   rows;
    const double* norm64;
    int64_t       row_base;
    int           n, d, ld;
    // request, in device-addressable pinned host memory
    const float*   q_host;        // [d] raw query, or nullptr
    const int32_t* liked_host;    // [n_liked] GLOBAL rows
    const float*   w_host;        // [n_liked] or nullptr
    int            n_liked;
    const int32_t* excl_host;     // [n_excl] GLOBAL rows, sorted unique
    int            n_excl;
    DevFilter      filter;        // device-resident predicates (bitmap / genre / year); its exclusion list is unused here
    int            k;
    unsigned long long* out_packed;   // rows k | fp64 scores k | count | margin (+inf)
};

template <typename T>
__global__ void __launch_bounds__(kSmallThreads, 1) small_recommend_kernel(const SmallParams p) {
    constexpr int EPC = ChunkDot<T>::EPC;
    extern __shared__ __align__(16) unsigned char sm[];
    double* s_q = (double*)sm;                              // [ld] pair planes
    double* s_sc = s_q + p.ld;                              // [n] exact score, or -inf for a row that may not be returned
    double* s_sum = s_sc + ((p.n + 1) & ~1);                // [ld] profile sum / staging
    int32_t* s_excl = (int32_t*)(s_sum + p.ld);             // [n_excl]
    float* s_src = (float*)(s_excl + ((p.n_excl + 3) & ~3));    // [d] raw query
    double* s_gsc = (double*)(s_src + ((p.d + 3) & ~3));    // [kSmallMaxGather]
    int* s_grow = (int*)(s_gsc + kSmallMaxGather);          // [kSmallMaxGather]
    __shared__ double red[32];
    __shared__ double s_wsum;
    __shared__ double s_tmax[kSmallThreads];
    __shared__ int s_trow[kSmallThreads];
    __shared__ double s_T;
    __shared__ int s_Trow, s_cnt, s_have;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kSmallThreads / 32;
    const T* rows = (const T*)p.rows;

    pdl_trigger();
    pdl_wait();
    copy_list_zero_copy(p.excl_host, p.n_excl, s_excl);
    if (p.q_host) {
        fetch_query_zero_copy(p.q_host, p.d, s_src);
        __syncthreads();
        const double nrm = query_norm_256(s_src, p.d, red);
        for (int c = threadIdx.x; c < p.ld; c += blockDim.x) s_sum[c] = c < p.d ? (double)s_src[c] / nrm : 0.0;
    } else {
        profile_accumulate_cta<T, true, 8>(rows, p.norm64, p.n, p.row_base, p.ld, p.liked_host, p.w_host, p.n_liked, s_sum, &s_wsum);
        __syncthreads();
        const double ws = s_wsum;
        for (int c = threadIdx.x; c < p.ld; c += blockDim.x) s_sum[c] = ws != 0.0 ? s_sum[c] / ws : 0.0;
    }
    __syncthreads();
    stage_query_planes(s_sum, p.ld, EPC, s_q);
    if (threadIdx.x == 0) { s_cnt = 0; s_have = 0; s_T = -INFINITY; s_Trow = 0x7fffffff; }
    __syncthreads();

    // ---- exact score of every allowed row.  Rows of up to 32 chunks (the production 32-d catalog has 8) are packed
    // several to a warp, four steps in flight, because one row per warp would be nothing but L2 latency; the per-row
    // arithmetic and the reduction tree are those of exact_score_row (lanes that hold no chunk add +0.0), so the bits agree.
    const QueryPlanes qsrc{s_q, p.ld / EPC};
    DevFilter f = p.filter;
    f.n_exclude = p.n_excl;
    const int chunks = p.ld / EPC;
    if (chunks <= 32) {
        constexpr int PAIRS = ChunkDot<T>::PAIRS;
        constexpr int U = 4;
        int G = 2;
        while (G < chunks) G <<= 1;
        const int rpw = 32 / G, sub = lane / G, cl = lane % G;
        double2 q[PAIRS];
        if (cl < chunks) qsrc.template load<PAIRS>(cl, q);
        for (int base = warp * rpw * U; base < p.n; base += nwarps * rpw * U) {
            uint4 v[U];
            double nrm[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = base + u * rpw + sub;
                ok[u] = r < p.n && row_allowed(f, (uint32_t)r, s_excl);
                nrm[u] = 1.0;
                v[u] = make_uint4(0u, 0u, 0u, 0u);
                if (ok[u] && cl < chunks) {
                    v[u] = __ldg((const uint4*)(rows + (size_t)r * p.ld) + cl);
                    nrm[u] = __ldg(p.norm64 + r);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int r = base + u * rpw + sub;
                double a0 = 0.0;
                if (ok[u] && cl < chunks) a0 = ChunkDot<T>::template dot<true>(v[u], q, a0, nrm[u]);
                double sc = a0 + 0.0;
                for (int o = G >> 1; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
                if (cl == 0 && r < p.n) s_sc[r] = ok[u] ? sc : -INFINITY;      // cosines are finite, so -inf marks "not allowed"
            }
        }
    } else {
        for (int r = warp; r < p.n; r += nwarps) {
            const bool ok = row_allowed(f, (uint32_t)r, s_excl);
            double sc = -INFINITY;
            if (ok) sc = exact_score_row<T, true>(rows, p.ld, p.norm64, (uint32_t)r, qsrc, lane);
            if (lane == 0) s_sc[r] = ok ? sc : -INFINITY;
        }
    }
    __syncthreads();

    // ---- top-k under (score desc, row asc).  Prune with the k-th best of the per-thread maxima (k distinct rows reach
    // it, so it bounds the k-th best overall from below), gather the rows at or above it, rank-count those.
    const int k = p.k;
    {
        double bs = -INFINITY;
        int br = 0x7fffffff;
        for (int r = threadIdx.x; r < p.n; r += blockDim.x) {
            const double sc = s_sc[r];
            if (sc != -INFINITY && better(sc, r, bs, br)) { bs = sc; br = r; }
        }
        s_tmax[threadIdx.x] = bs;
        s_trow[threadIdx.x] = br;
    }
    __syncthreads();
    if (k <= kSmallThreads) {
        const double bs = s_tmax[threadIdx.x];
        const int br = s_trow[threadIdx.x];
        if (bs != -INFINITY) {
            int c = 0;
            for (int t = 0; t < kSmallThreads; ++t) c += s_tmax[t] != -INFINITY && better(s_tmax[t], s_trow[t], bs, br);
            if (c == k - 1) { s_T = bs; s_Trow = br; s_have = 1; }
        }
    }
    __syncthreads();
    const bool have = s_have != 0;                          // else fewer than k threads hold a row: every allowed row is gathered
    const double Ts = s_T;
    const int Tr = s_Trow;
    for (int r = threadIdx.x; r < p.n; r += blockDim.x) {
        const double sc = s_sc[r];
        if (sc == -INFINITY) continue;
        if (have && better(Ts, Tr, sc, r)) continue;        // strictly worse than the threshold row
        const int idx = atomicAdd(&s_cnt, 1);
        if (idx < kSmallMaxGather) { s_gsc[idx] = sc; s_grow[idx] = r; }
    }
    __syncthreads();
    const int g = s_cnt;
    int64_t* o_rows = (int64_t*)p.out_packed;
    double* o_scores = (double*)(p.out_packed + k);
    if (g <= kSmallMaxGather) {
        for (int i = threadIdx.x; i < g; i += blockDim.x) {
            const double sc = s_gsc[i];
            const int r = s_grow[i];
            int rank = 0;
            for (int j = 0; j < g; ++j) rank += better(s_gsc[j], s_grow[j], sc, r);
            if (rank < k) { o_rows[rank] = p.row_base + r; o_scores[rank] = sc; }
        }
    } else {
        // mass ties (thousands of rows at the threshold): rank every allowed row against all rows — slow, still exact
        for (int r = threadIdx.x; r < p.n; r += blockDim.x) {
            const double sc = s_sc[r];
            if (sc == -INFINITY) continue;
            int rank = 0;
            for (int j = 0; j < p.n && rank < k; ++j) rank += s_sc[j] != -INFINITY && better(s_sc[j], j, sc, r);
            if (rank < k) { o_rows[rank] = p.row_base + r; o_scores[rank] = sc; }
        }
    }
    const int nout = g < k ? g : k;
    for (int i = nout + threadIdx.x; i < k; i += blockDim.x) { o_rows[i] = -1; o_scores[i] = -INFINITY; }
    if (threadIdx.x == 0) {
        p.out_packed[2 * k] = (unsigned long long)(unsigned)nout;
        p.out_packed[2 * k + 1] = (unsigned long long)__double_as_longlong(INFINITY);   // every row was scored exactly
    }
}

bool small_catalog(const rebert_catalog_t* cat, int n_excl) {
    if (cat->n <= 0 || cat->n > kSmallMaxRows || (int64_t)cat->n * cat->ld > kSmallMaxElems) return false;
    if (cat->dtype != REBERT_F32 && cat->dtype != REBERT_BF16) return false;
    return small_smem_bytes((int)cat->n, cat->d, cat->ld, n_excl) <= 160 * 1024;
}

int small_recommend_launch(const rebert_catalog_t* cat, const float* q_host, const int32_t* liked_host, const float* w_host, int n_liked,
                           const int32_t* excl_host, int n_excl, const rebert_filter_t* device_filter, int k,
                           unsigned long long* out_packed, cudaStream_t st) {
    SmallParams p;
    memset(&p, 0, sizeof(p));
    p.rows = cat->rows;
    p.norm64 = cat->norm64;
    p.row_base = cat->row_base;
    p.n = (int)cat->n;
    p.d = cat->d;
    p.ld = cat->ld;
    p.q_host = q_host;
    p.liked_host = liked_host;
    p.w_host = w_host;
    p.n_liked = n_liked;
    p.excl_host = excl_host;
    p.n_excl = n_excl;
    rebert_filter_t f;
    memset(&f, 0, sizeof(f));
    if (device_filter) { f = *device_filter; f.exclude_rows = nullptr; f.n_exclude = 0; }
    p.filter = make_filter(&f, cat->row_base);
    p.k = k;
    p.out_packed = out_packed;
    const size_t smem = small_smem_bytes((int)cat->n, cat->d, cat->ld, n_excl);
    if (cat->dtype == REBERT_F32) {
        auto kern = small_recommend_kernel<float>;
        { int rc = raise_smem_limit(kern); if (rc != REBERT_OK) return rc; }
        REBERT_CUDA(launch_pdl(kern, dim3(1), dim3(kSmallThreads), smem, st, p));
    } else {
        auto kern = small_recommend_kernel<__nv_bfloat16>;
        { int rc = raise_smem_limit(kern); if (rc != REBERT_OK) return rc; }
        REBERT_CUDA(launch_pdl(kern, dim3(1), dim3(kSmallThreads), smem, st, p));
    }
    REBERT_CUDA(cudaGetLastError());
    return REBERT_OK;
}

}  // namespace rebert
