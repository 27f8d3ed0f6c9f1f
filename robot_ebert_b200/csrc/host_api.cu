// Host-buffer entry point: one C call = one request of the reference's hot path (lib.py:43-55) end to end.
// Host pointers in, host pointers out; staging the request, every kernel and the stream synchronisation all happen
// inside.  A raw query is read by the first kernel straight from the caller's pinned block (zero-copy) and the exact
// pass writes its packed result straight into that block, so a query request involves no copy-engine operation at
// all; a liked-rows request ships its lists with one H2D copy.  rebert_recommend_host_sharded appends the fused NVLink
// exchange + merge for one rank of a row-sharded catalog.  Scratch (pinned host + device) is provided by the caller,
// so the library still allocates nothing and concurrent callers only need their own scratch + stream.
#include "common.cuh"

namespace rebert {

static size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

struct HostLayout {
    // request block (identical offsets in the pinned buffer and at the start of the device scratch)
    size_t off_q, off_rp, off_excl, off_col, off_w, in_bytes;
    // device-only
    size_t off_qn32, off_qn64, off_sum64, off_wsum, off_ws, ws_bytes, off_cand, off_out, out_bytes, dev_bytes;
    // pinned-only: result block after the request block
    size_t pin_out, pin_bytes;
};

static HostLayout host_layout(int ld, int d, int n_liked_cap, int n_excl_cap, int k, int kc, int64_t n) {
    HostLayout L;
    L.off_q = 0;
    L.off_rp = al16((size_t)d * 4);
    L.off_excl = L.off_rp + 16;
    L.off_col = L.off_excl + al16((size_t)n_excl_cap * 4);
    L.off_w = L.off_col + al16((size_t)n_liked_cap * 4);
    L.in_bytes = L.off_w + al16((size_t)n_liked_cap * 4);
    size_t o = al256(L.in_bytes);
    L.off_qn32 = o; o += al256((size_t)ld * 4);
    L.off_qn64 = o; o += al256((size_t)ld * 8);
    L.off_sum64 = o; o += al256((size_t)ld * 8);
    L.off_wsum = o; o += 256;
    L.ws_bytes = rebert_gemv_workspace_bytes(n, kc);
    L.off_ws = o; o += al256(L.ws_bytes);
    L.off_cand = o; o += al256((size_t)kc * 8);
    L.out_bytes = (size_t)(2 * k + 2) * 8;
    L.off_out = o; o += al256(L.out_bytes);
    L.dev_bytes = o;
    L.pin_out = al256(L.in_bytes);
    L.pin_bytes = L.pin_out + al256(L.out_bytes + 8);        // + the exchange kernel's error word (sharded entry point)
    return L;
}

}  // namespace rebert

using namespace rebert;

extern "C" {

REBERT_API int rebert_recommend_host_scratch(const rebert_catalog_t* cat, int32_t n_liked_cap, int32_t n_exclude_cap, int32_t k,
                                             size_t* pinned_bytes, size_t* device_bytes) {
    REBERT_REQUIRE(cat && k > 0 && n_liked_cap >= 0 && n_exclude_cap >= 0, "recommend_host_scratch: bad arguments");
    const int kc_max = 256;
    HostLayout L = host_layout(cat->ld, cat->d, n_liked_cap, n_exclude_cap, k, kc_max, cat->n);
    if (pinned_bytes) *pinned_bytes = L.pin_bytes;
    if (device_bytes) *device_bytes = L.dev_bytes;
    return REBERT_OK;
}

// Optional last step for a row-sharded catalog: the fused NVLink exchange + merge (rebert_exchange_merge).
struct HostExchange {
    const uint64_t* peer_buffers;
    int32_t world, rank, k_max;
    uint32_t seq;
};

static int recommend_host_impl(const rebert_catalog_t* cat, const float* query, const int32_t* liked_rows,
                               const float* liked_w, int32_t n_liked, const int32_t* exclude_rows, int32_t n_exclude,
                               const rebert_filter_t* device_filter, int32_t k, int32_t kc, int32_t n_liked_cap,
                               int32_t n_exclude_cap, void* pinned, size_t pinned_bytes, void* device_scratch,
                               size_t device_bytes, int64_t* out_rows, double* out_scores, int32_t* out_count,
                               double* out_margin, rebert_stream stream, const HostExchange* ex) {
    REBERT_REQUIRE(cat && cat->rows && pinned && device_scratch && out_rows && out_scores && out_count, "recommend_host: null argument");
    REBERT_REQUIRE((query != nullptr) != (liked_rows != nullptr), "recommend_host: pass exactly one of query / liked_rows");
    REBERT_REQUIRE(k > 0 && kc >= k && kc <= 256, "recommend_host: k=%d kc=%d", k, kc);
    REBERT_REQUIRE(n_liked >= 0 && n_liked <= n_liked_cap && n_exclude >= 0 && n_exclude <= n_exclude_cap,
                   "recommend_host: list longer than the scratch capacity");
    if (liked_rows && n_liked == 0) {
        // what sklearn's check_array raises in the reference when no rated movie is liked (lib.py:47,51)
        set_error("Found array with 0 sample(s): user has no liked movies in the catalog");
        return REBERT_ERR_INVALID;
    }
    // device-only regions are placed by the capacities; the request block is packed by the actual list lengths
    const HostLayout Lc = host_layout(cat->ld, cat->d, n_liked_cap, n_exclude_cap, k, 256, cat->n);
    HostLayout L = host_layout(cat->ld, cat->d, liked_rows ? n_liked : 0, n_exclude, k, 256, cat->n);
    L.off_qn32 = Lc.off_qn32; L.off_qn64 = Lc.off_qn64; L.off_sum64 = Lc.off_sum64; L.off_wsum = Lc.off_wsum;
    L.off_ws = Lc.off_ws; L.ws_bytes = Lc.ws_bytes; L.off_cand = Lc.off_cand; L.off_out = Lc.off_out;
    L.dev_bytes = Lc.dev_bytes; L.pin_out = Lc.pin_out; L.pin_bytes = Lc.pin_bytes;
    if (pinned_bytes < L.pin_bytes || device_bytes < L.dev_bytes) {
        set_error("recommend_host: scratch too small (pinned %zu < %zu or device %zu < %zu)", pinned_bytes, L.pin_bytes,
                  device_bytes, L.dev_bytes);
        return REBERT_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* h = (unsigned char*)pinned;
    unsigned char* dv = (unsigned char*)device_scratch;
    float* qn32 = (float*)(dv + L.off_qn32);
    double* qn64 = (double*)(dv + L.off_qn64);
    int rc;
    // ---- the request goes into the pinned block.  A raw query is then read by the staging kernel straight from that
    // block (zero-copy: no copy-engine operation in front of the kernels); a liked-rows request ships with ONE H2D copy.
    if (n_exclude) memcpy(h + L.off_excl, exclude_rows, (size_t)n_exclude * 4);
    if (query) {
        memcpy(h + L.off_q, query, (size_t)cat->d * 4);
        rc = stage_query_launch((const float*)(h + L.off_q), cat->d, cat->ld, qn32, qn64, (const int32_t*)(h + L.off_excl),
                                n_exclude, (int32_t*)(dv + L.off_excl), st);
        if (rc != REBERT_OK) return rc;
    } else {
        int64_t rp[2] = {0, n_liked};
        memcpy(h + L.off_rp, rp, 16);
        memcpy(h + L.off_col, liked_rows, (size_t)n_liked * 4);
        if (liked_w) memcpy(h + L.off_w, liked_w, (size_t)n_liked * 4);
        REBERT_CUDA(cudaMemcpyAsync(dv + L.off_rp, h + L.off_rp, L.in_bytes - L.off_rp, cudaMemcpyHostToDevice, st));
        double* sum64 = (double*)(dv + L.off_sum64);
        double* wsum = (double*)(dv + L.off_wsum);
        rc = rebert_profile_accumulate(cat, (const int64_t*)(dv + L.off_rp), (const int32_t*)(dv + L.off_col),
                                       liked_w ? (const float*)(dv + L.off_w) : nullptr, 1, sum64, wsum, stream);
        if (rc != REBERT_OK) return rc;
        rc = rebert_profile_finalize(sum64, wsum, 1, cat->ld, qn32, qn64, nullptr, stream);
        if (rc != REBERT_OK) return rc;
    }
    rebert_filter_t f;
    memset(&f, 0, sizeof(f));
    if (device_filter) f = *device_filter;
    if (n_exclude) {
        f.exclude_rows = (const int32_t*)(dv + L.off_excl);
        f.n_exclude = n_exclude;
    }
    uint64_t* cand = (uint64_t*)(dv + L.off_cand);
    rc = rebert_gemv_topk(cat, qn32, &f, kc, dv + L.off_ws, L.ws_bytes, cand, stream);
    if (rc != REBERT_OK) return rc;
    // ---- exact pass.  Its packed result (rows[k] | scores[k] | count | margin) is written by the kernel straight into
    // the pinned block (no D2H copy operation) — or, on a row shard, into device scratch for the exchange kernel, which
    // then writes the merged block and its error word into the pinned block.
    unsigned char* r = h + L.pin_out;
    unsigned char* ob = ex ? dv + L.off_out : r;
    rc = rebert_finalize_topk(cat, qn64, cand, kc, k, (int64_t*)ob, (double*)(ob + 8 * (size_t)k), (int32_t*)(ob + 16 * (size_t)k),
                              (double*)(ob + 16 * (size_t)k + 8), stream);
    if (rc != REBERT_OK) return rc;
    int32_t* err_word = (int32_t*)(r + L.out_bytes);
    if (ex) {
        *err_word = 0;
        rc = rebert_exchange_merge(ex->peer_buffers, ex->world, ex->rank, k, ex->k_max, ex->seq, (const int64_t*)ob, (int64_t*)r,
                                   err_word, stream);
        if (rc != REBERT_OK) return rc;
    }
    REBERT_CUDA(cudaStreamSynchronize(st));
    if (ex && *err_word != 0) {
        set_error("recommend_host_sharded: peer %d did not deliver its result to the exchange kernel", *err_word - 1);
        return REBERT_ERR_CUDA;
    }
    const int32_t cnt = *(const int32_t*)(r + 16 * (size_t)k);
    memcpy(out_rows, r, (size_t)k * 8);
    memcpy(out_scores, r + 8 * (size_t)k, (size_t)k * 8);
    *out_count = cnt;
    if (out_margin) *out_margin = *(const double*)(r + 16 * (size_t)k + 8);
    return REBERT_OK;
}

REBERT_API int rebert_recommend_host(const rebert_catalog_t* cat, const float* query, const int32_t* liked_rows,
                                     const float* liked_w, int32_t n_liked, const int32_t* exclude_rows, int32_t n_exclude,
                                     const rebert_filter_t* device_filter, int32_t k, int32_t kc, int32_t n_liked_cap,
                                     int32_t n_exclude_cap, void* pinned, size_t pinned_bytes, void* device_scratch,
                                     size_t device_bytes, int64_t* out_rows, double* out_scores, int32_t* out_count,
                                     double* out_margin, rebert_stream stream) {
    return recommend_host_impl(cat, query, liked_rows, liked_w, n_liked, exclude_rows, n_exclude, device_filter, k, kc, n_liked_cap,
                               n_exclude_cap, pinned, pinned_bytes, device_scratch, device_bytes, out_rows, out_scores, out_count,
                               out_margin, stream, nullptr);
}

REBERT_API int rebert_recommend_host_sharded(const rebert_catalog_t* cat, const float* query, const int32_t* exclude_rows,
                                             int32_t n_exclude, const rebert_filter_t* device_filter, int32_t k, int32_t kc,
                                             int32_t n_exclude_cap, void* pinned, size_t pinned_bytes, void* device_scratch,
                                             size_t device_bytes, const uint64_t* peer_buffers, int32_t world, int32_t rank,
                                             int32_t k_max, uint32_t seq, int64_t* out_rows, double* out_scores,
                                             int32_t* out_count, double* out_margin, rebert_stream stream) {
    REBERT_REQUIRE(query && peer_buffers, "recommend_host_sharded: null argument");
    HostExchange ex;
    ex.peer_buffers = peer_buffers; ex.world = world; ex.rank = rank; ex.k_max = k_max; ex.seq = seq;
    return recommend_host_impl(cat, query, nullptr, nullptr, 0, exclude_rows, n_exclude, device_filter, k, kc, 0, n_exclude_cap, pinned,
                               pinned_bytes, device_scratch, device_bytes, out_rows, out_scores, out_count, out_margin, stream, &ex);
}

}  // extern "C"
