/* Minimal plain-C client of libblr_cuda (include/blr_cuda.h): the README toy of the reference
 * (BayesianLinearRegressors.jl README.md:44-60: D = 2 features [x; 1], N = 10 observations, prior w ~ N(0, I),
 * noise 0.1) through the C ABI -- posterior + logpdf, then marginals on three test points.
 *
 *   gcc -std=c99 -I include examples/minimal_client.c -L bayesianlinearregressors.jl_b200/csrc -lblr_cuda \
 *       -Wl,-rpath,$PWD/bayesianlinearregressors.jl_b200/csrc -o minimal_client && ./minimal_client
 *
 * Needs a B200 to RUN (blr_ctx_create fails otherwise: there is no CPU fallback); it compiles and links anywhere,
 * which is what tests/test_abi.py checks. */
#include <stdio.h>
#include <stdlib.h>

#include "blr_cuda.h"

#define CHECK(call)                                                                       \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != 0) {                                                                   \
            fprintf(stderr, "%s -> %d (%s)\n", #call, rc_, ctx ? blr_last_error(ctx) : ""); \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

int main(void) {
    enum { D = 2, N = 10, NT = 3 };
    blr_ctx* ctx = NULL;
    if (blr_ctx_create(&ctx, 0) != 0) {
        fprintf(stderr, "no usable sm_100 device: libblr_cuda has no CPU fallback\n");
        return 2;
    }
    double X[D * N], y[N], Xt[D * NT] = {-6.0, 1.0, 0.0, 1.0, 6.0, 1.0};
    for (int n = 0; n < N; ++n) {
        const double x = -5.0 + 10.0 * n / (N - 1);
        X[D * n] = x; /* ColVecs: one observation = D contiguous doubles */
        X[D * n + 1] = 1.0;
        y[n] = 0.5 * x - 1.0;
    }
    double mw[D] = {0.0, 0.0}, lam_diag[D] = {1.0, 1.0};
    blr_prior prior;
    prior.mw = mw;
    prior.lambda_kind = BLR_LAMBDA_DIAGONAL;
    prior.lambda = lam_diag;
    prior.ld = D;
    prior.D = D;
    blr_noise noise;
    noise.kind = BLR_NOISE_SCALAR;
    noise.scalar = 0.1;
    noise.vec = NULL;
    noise.dense = NULL;
    noise.dense_ld = 0;

    blr_x *x = NULL, *xt = NULL;
    blr_vec* yv = NULL;
    blr_post* post = NULL;
    CHECK(blr_x_upload(ctx, X, D, N, D, BLR_COLVECS, &x));
    CHECK(blr_vec_upload(ctx, y, N, &yv));
    double logpdf, m_post[D], Lambda_post[D * D];
    CHECK(blr_infer(ctx, &prior, x, yv, &noise, &logpdf, m_post, NULL, Lambda_post, &post));
    printf("logpdf = %.17g\nposterior mean = [%.17g, %.17g]\n", logpdf, m_post[0], m_post[1]);

    double mean[NT], var[NT];
    CHECK(blr_x_upload(ctx, Xt, D, NT, D, BLR_COLVECS, &xt));
    CHECK(blr_mean_var(ctx, post, xt, &noise, mean, var));
    for (int i = 0; i < NT; ++i) printf("x* = %+.1f: mean %.17g, var %.17g\n", Xt[D * i], mean[i], var[i]);

    blr_post_free(ctx, post);
    blr_x_free(ctx, xt);
    blr_x_free(ctx, x);
    blr_vec_free(ctx, yv);
    blr_ctx_destroy(ctx);
    return 0;
}
