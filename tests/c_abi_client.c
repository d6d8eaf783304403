/* A plain-C client of include/ewk.h — what a binding in any language reduces to.  Built and run by
 * tests/test_abi.py::test_plain_c_client: with a GPU it scores a 440 Hz tone against itself through ewk_set_template /
 * ewk_similarity_batch (WordMatcher.set_reference / matches, wakeword.py:569-639; the reference's own known-answer test:
 * self-similarity == 100.0, tests/test_wakeword_simulated.py:107-118) and exits 0; without one ewk_create must refuse
 * loudly (EWK_ERR_CUDA, "no CPU fallback") and the program exits 3. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ewk.h"

int main(void) {
    ewk_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.max_templates = 1;
    ewk_ctx* ctx = NULL;
    int rc = ewk_create(0, &cfg, &ctx);
    if (rc != EWK_OK) {
        printf("ewk_create: %d: %s\n", rc, ewk_last_error(NULL));
        return rc == EWK_ERR_CUDA ? 3 : 1;
    }
    enum { N = 16000 };
    float* x = (float*)malloc(sizeof(float) * N);
    for (int i = 0; i < N; i++) x[i] = 0.5f * (float)sin(2.0 * 3.14159265358979323846 * 440.0 * i / 16000.0);
    float score = -1.f;
    uint8_t matched = 0;
    const int64_t off = 0, len = N;
    rc = ewk_similarity_batch(ctx, 0, x, EWK_PCM_F32, EWK_HOST, &off, &len, 1, 75.0f, &score, &matched, NULL);
    if (rc != EWK_ERR_NO_TEMPLATE) { printf("expected EWK_ERR_NO_TEMPLATE, got %d\n", rc); return 1; }
    printf("no template: %s\n", ewk_last_error(ctx));
    if ((rc = ewk_set_template(ctx, 0, x, N)) != EWK_OK) { printf("ewk_set_template: %s\n", ewk_last_error(ctx)); return 1; }
    if ((rc = ewk_similarity_batch(ctx, 0, x, EWK_PCM_F32, EWK_HOST, &off, &len, 1, 75.0f, &score, &matched, NULL)) != EWK_OK) {
        printf("ewk_similarity_batch: %s\n", ewk_last_error(ctx));
        return 1;
    }
    printf("abi %d  self-similarity %.4f  matched %d\n", ewk_abi_version(), score, (int)matched);
    ewk_destroy(ctx);
    free(x);
    return (score == 100.0f && matched == 1) ? 0 : 2;
}
