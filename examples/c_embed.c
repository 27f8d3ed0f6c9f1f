/* Minimal plain-C embedder of librebert_b200.so: what a non-Python host (the reference has none, but any C / Go-cgo /
 * Rust-FFI service would look like this) binds.  Build:
 *   gcc -std=c99 -Iinclude examples/c_embed.c -o c_embed -Lrobot_ebert_b200 -lrebert_b200 -Wl,-rpath,$PWD/robot_ebert_b200
 * Without a GPU it exercises only the argument-validation / layout entry points; on a B200 it runs one request with
 * HOST buffers through rebert_recommend_host (device memory comes from the CUDA runtime here, from torch in Python). */
#include <stdio.h>
#include <string.h>

#include "rebert_b200.h"

int main(void) {
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
    return 0;
}
