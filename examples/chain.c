/* Plain C99 caller of libdip_b200.so: the reference's grayscale -> GaussianBlur -> threshold -> erosion
 * (opencv/benchmark.py:96,105,97,99) as one dip_apply_chain call on a synthetic BGR image.
 *
 *   gcc -std=c99 -Iinclude examples/chain.c -Ldip-benchmark_b200 -ldip_b200 -Wl,-rpath,'$ORIGIN/../dip-benchmark_b200' -o examples/chain
 *   examples/chain WIDTH HEIGHT OUT.raw      (writes the H x W x 3 result bytes)
 */
#include <stdio.h>
#include <stdlib.h>

#include "dip.h"

#define CHECK(expr)                                                                         \
    do {                                                                                    \
        if ((expr) < 0) {                                                                   \
            fprintf(stderr, "%s failed: %s\n", #expr, dip_last_error_string(ctx));          \
            return 1;                                                                       \
        }                                                                                   \
    } while (0)

int main(int argc, char** argv) {
    if (argc != 4) {
        fprintf(stderr, "usage: %s WIDTH HEIGHT OUT.raw\n", argv[0]);
        return 2;
    }
    const int w = atoi(argv[1]), h = atoi(argv[2]);
    if (w < 1 || h < 1) return 2;
    dip_ctx* ctx = NULL;
    if (dip_init(&ctx, 0) < 0) {
        fprintf(stderr, "no GPU: %s\n", dip_last_error_string(NULL));
        return 4;
    }
    const size_t n = (size_t)w * h * 3;
    uint8_t* host = (uint8_t*)malloc(n);
    if (!host) return 3;
    for (size_t i = 0; i < n; ++i) host[i] = (uint8_t)((i * 2654435761u) >> 13);  /* any deterministic content */

    dip_image *src = NULL, *dst = NULL;
    CHECK(dip_image_create(ctx, w, h, 3, &src));
    CHECK(dip_image_create(ctx, w, h, 3, &dst));
    CHECK(dip_upload(ctx, src, host, 0));

    dip_op ref[12], chain[4];
    CHECK(dip_reference_ops(ref, 12));
    chain[0] = ref[2];   /* grayscale */
    chain[1] = ref[11];  /* GaussianBlur 3x3 */
    chain[2] = ref[3];   /* threshold(127, 255) */
    chain[3] = ref[5];   /* erosion, 3x3 square */
    int passes = 0;
    CHECK(dip_chain_passes(chain, 4, 3, &passes));
    CHECK(dip_apply_chain(ctx, chain, 4, src, dst));
    CHECK(dip_download(ctx, dst, host, 0));
    CHECK(dip_sync(ctx));

    FILE* f = fopen(argv[3], "wb");
    if (!f || fwrite(host, 1, n, f) != n) return 5;
    fclose(f);
    printf("%dx%d: 4 operations in %d passes\n", w, h, passes);
    free(host);
    dip_image_destroy(src);
    dip_image_destroy(dst);
    dip_shutdown(ctx);
    return 0;
}
