/* Plain-C embedder of librebert_b200.so — what any non-Python host (C, Go via cgo, Rust via FFI ...) binds.
 *
 *   gcc -std=c99 -Iinclude examples/c_embed.c -o c_embed -Lrobot_ebert_b200 -lrebert_b200 -Wl,-rpath,$PWD/robot_ebert_b200
 *   gcc ... -DWITH_CUDA -I/usr/local/cuda/include -L/usr/local/cuda/lib64 -lcudart      (adds the GPU part)
 *
 * Part 1 (always): layout / planning / argument-validation entry points — no device needed.
 * Part 2 (-DWITH_CUDA, needs a B200): builds a synthetic bf16 catalog in HBM with the library's own kernels and serves ONE
 * request with HOST buffers through rebert_recommend_host — the call that replaces lib.py:43-55 — then prints the rows.
 * Usage of part 2:  c_embed <rows> <dim> <k>      (query = synthetic seed-1 vector, exclusions = rows 0,7,14,...,<700) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "rebert_b200.h"

#ifdef WITH_CUDA
#include <cuda_runtime_api.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 20; } } while (0)
#define RB(x) do { int r_ = (x); if (r_ != REBERT_OK) { fprintf(stderr, "%s -> %d: %s\n", #x, r_, rebert_last_error()); return 21; } } while (0)

static int gpu_part(long long n, int d, int k) {
    int32_t ld = 0, kc = rebert_candidates_for_k(k), cnt = 0, nex = 0, i;
    size_t bytes = 0, pinned_bytes = 0, device_bytes = 0;
    void *rows = NULL, *pinned = NULL, *scratch = NULL, *qdev = NULL;
    float *inv = NULL, *q = NULL;
    double *nrm = NULL, *scores = NULL;
    int64_t* out_rows = NULL;
    int32_t excl[100];
    rebert_catalog_t cat;
    rebert_proof_t proof;
    rebert_request_info_t info;

    RB(rebert_check_device());
    RB(rebert_catalog_layout(n, d, REBERT_BF16, &ld, &bytes));
    CK(cudaMalloc(&rows, bytes));
    CK(cudaMalloc((void**)&inv, (size_t)(n + 4) * sizeof(float)));
    CK(cudaMalloc((void**)&nrm, (size_t)n * sizeof(double)));
    RB(rebert_synth_rows(0, 0, n, d, 1, REBERT_BF16, rows, ld, NULL));             /* the catalog (constants.py:55-56) */
    RB(rebert_catalog_norms(rows, n, ld, REBERT_BF16, inv, nrm, NULL));
    memset(&cat, 0, sizeof(cat));
    cat.rows = rows; cat.inv_norm = inv; cat.norm64 = nrm; cat.n = n; cat.row_base = 0; cat.d = d; cat.ld = ld; cat.dtype = REBERT_BF16;
    /* the query: row 0 of the synthetic stream with seed 1, generated on the device and copied back to a HOST buffer */
    q = (float*)malloc((size_t)d * sizeof(float));
    CK(cudaMalloc(&qdev, (size_t)ld * sizeof(float)));
    RB(rebert_synth_rows(1, 0, 1, d, 0, REBERT_F32, qdev, ld, NULL));
    CK(cudaMemcpy(q, qdev, (size_t)d * sizeof(float), cudaMemcpyDeviceToHost));
    for (i = 0; i < 100 && (long long)i * 7 < n; ++i) excl[nex++] = i * 7;          /* the user's rated movies (lib.py:48) */
    RB(rebert_recommend_host_scratch(&cat, 1024, 1024, k, &pinned_bytes, &device_bytes));
    CK(cudaHostAlloc(&pinned, pinned_bytes, cudaHostAllocDefault));
    CK(cudaMalloc(&scratch, device_bytes));
    CK(cudaMemset(scratch, 0, device_bytes));                                        /* zero once: ticket counter */
    out_rows = (int64_t*)malloc((size_t)k * sizeof(int64_t));
    scores = (double*)malloc((size_t)k * sizeof(double));
    memset(&proof, 0, sizeof(proof));
    proof.fast_eps = 1e-5;                                                            /* > (elements per lane + 12) * 2^-24, DESIGN.md 4.2 */
    proof.widen = 1;
    RB(rebert_recommend_host(&cat, q, NULL, NULL, 0, excl, nex, NULL, k, kc, 1024, 1024, pinned, pinned_bytes, scratch, device_bytes,
                             &proof, NULL, out_rows, scores, &cnt, &info, NULL));
    printf("results %d margin_ok %d\n", cnt, info.proven);
    for (i = 0; i < cnt; ++i) printf("row %lld score %.17g\n", (long long)out_rows[i], scores[i]);
    return 0;
}
#endif

int main(int argc, char** argv) {
    int32_t ld = 0;
    size_t bytes = 0;
    rebert_gemm_plan_t plan;
    rebert_catalog_t cat;
    rebert_filter_t filt;

    if (rebert_abi_version() != REBERT_ABI_VERSION) return 1;
    if (rebert_catalog_layout(1000000, 1536, REBERT_BF16, &ld, &bytes) != REBERT_OK) return 2;
    printf("layout: ld=%d bytes=%zu\n", ld, bytes);
    if (ld != 1536 || bytes != (size_t)1000000 * 1536 * 2) return 3;
    if (rebert_candidates_for_k(10) != 32 || rebert_candidates_for_k(100) != 128) return 4;
    if (rebert_gemm_plan(1000000, 4096, 100, &plan) != REBERT_OK || plan.kc != 128) return 5;
    /* invalid calls report through status + thread-local message, never by crashing */
    memset(&cat, 0, sizeof(cat));
    memset(&filt, 0, sizeof(filt));
    if (rebert_gemv_topk(&cat, NULL, &filt, 32, NULL, 0, NULL, NULL) != REBERT_ERR_INVALID) return 6;
    printf("last error: %s\n", rebert_last_error());
    if (rebert_catalog_layout(10, 0, REBERT_F32, &ld, &bytes) != REBERT_ERR_INVALID) return 7;
    printf("c_embed ok\n");
#ifdef WITH_CUDA
    if (argc == 4) return gpu_part(atoll(argv[1]), atoi(argv[2]), atoi(argv[3]));
#else
    (void)argc; (void)argv;
#endif
    return 0;
}
