// Batched queries on the tensor cores — placeholder until the tcgen05 kernel lands (fails loudly, no fallback).
#include "common.cuh"
using namespace rebert;
extern "C" {
REBERT_API int rebert_gemm_plan(int64_t, int32_t, int32_t, rebert_gemm_plan_t*) { set_error("gemm path not built yet"); return REBERT_ERR_UNSUPPORTED; }
REBERT_API size_t rebert_gemm_workspace_bytes(const rebert_catalog_t*, const rebert_gemm_plan_t*) { return 0; }
REBERT_API int rebert_gemm_topk(const rebert_catalog_t*, const void*, const double*, const int64_t*, const int32_t*, const rebert_gemm_plan_t*,
                     void*, size_t, int64_t*, double*, int32_t*, int32_t*, rebert_stream) { set_error("gemm path not built yet"); return REBERT_ERR_UNSUPPORTED; }
REBERT_API int rebert_gemm_scores(const rebert_catalog_t*, const void*, int32_t, int64_t, int64_t, float*, rebert_stream) { set_error("gemm path not built yet"); return REBERT_ERR_UNSUPPORTED; }
}
