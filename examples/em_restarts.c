/*
 * em_restarts.c -- the drop-in boundary from plain C: nothing but include/ldsr_b200.h and
 * libldsr_b200.so (no Python, no torch, no R).  What LDS_EM_restart does for one series
 * (R/LDS_reconstruction.R:42-62): a handful of EM fits from different initial values, the best
 * one selected, its smoothed trajectory returned.
 *
 *   gcc -O2 -Iinclude examples/em_restarts.c -Lldsr_b200 -lldsr_b200 -Wl,-rpath,$PWD/ldsr_b200 -lm -o em_restarts
 *
 * Exit status: 0 on success, 2 when the library reports that there is no CUDA device (the library
 * has no CPU path and says so), 1 on any other error.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "ldsr_b200.h"

enum { T = 120, P = 2, Q = 2, RESTARTS = 8, TH = P + Q + 6 };

/* small deterministic generator: the example must not depend on the platform's rand() */
static unsigned long long state = 88172645463325252ULL;
static double unif(void) {
    state ^= state << 13;
    state ^= state >> 7;
    state ^= state << 17;
    return (double)(state >> 11) / 9007199254740992.0;
}
static double gauss(void) { return sqrt(-2.0 * log(unif() + 1e-300)) * cos(6.283185307179586 * unif()); }

int main(void) {
    static double y[T], u[P * T], v[Q * T], theta0[RESTARTS * TH];
    /* a stable system driven by two inputs; the first 60 steps are not observed (a reconstruction) */
    double x = 0.0;
    for (int t = 0; t < T; t++) {
        for (int j = 0; j < P; j++) u[t * P + j] = v[t * Q + j] = gauss();
        const double yt = 0.6 * x + 0.3 * v[t * Q] - 0.2 * v[t * Q + 1] + 0.3 * gauss();
        y[t] = t < 60 ? NAN : yt;
        x = 0.7 * x + 0.5 * u[t * P] + 0.2 * u[t * P + 1] + 0.4 * gauss();
    }
    /* make_init (R/LDS_reconstruction.R:14-30): A, C in [0,1], B, D in [-1,1], Q = R = 1, mu1 = 0, V1 = 1 */
    for (int r = 0; r < RESTARTS; r++) {
        double *th = theta0 + r * TH;
        th[0] = unif();
        for (int j = 0; j < P; j++) th[1 + j] = 2.0 * unif() - 1.0;
        th[1 + P] = unif();
        for (int j = 0; j < Q; j++) th[2 + P + j] = 2.0 * unif() - 1.0;
        th[2 + P + Q] = 1.0;
        th[3 + P + Q] = 1.0;
        th[4 + P + Q] = 0.0;
        th[5 + P + Q] = 1.0;
    }

    const int Ts[1] = {T}, ps[1] = {P}, qs[1] = {Q}, group_series[1] = {0};
    const double *ys[1] = {y}, *us[1] = {u}, *vs[1] = {v};
    int fit_group[RESTARTS] = {0};
    ldsr_batch b = {0};
    b.n_series = 1;
    b.T = Ts;
    b.p = ps;
    b.q = qs;
    b.y = ys;
    b.u = us;
    b.v = vs;
    b.n_groups = 1;
    b.group_series = group_series;
    b.n_fits = RESTARTS;
    b.fit_group = fit_group;
    b.theta0 = theta0;
    b.theta_stride = TH;

    static double theta[RESTARTS * TH], lik[RESTARTS], X[T], Y[T], V[T], J[T];
    int iters[RESTARTS], status[RESTARTS], best[1];
    ldsr_em_result r = {0};
    r.theta = theta;
    r.lik = lik;
    r.iters = iters;
    r.status = status;
    r.best = best;
    r.X = X;
    r.Y = Y;
    r.V = V;
    r.J = J;

    char err[512] = "";
    const int rc = ldsr_em_batch(NULL, &b, 300, 1e-5, NULL, &r, err, sizeof err);
    if (rc != LDSR_OK) {
        fprintf(stderr, "ldsr_em_batch failed (%d): %s\n", rc, err);
        return rc == LDSR_ERR_CUDA ? 2 : 1;
    }
    for (int k = 0; k < RESTARTS; k++)
        printf("restart %d: lik %.9f after %d E-steps, A %.6f C %.6f status %d\n", k, lik[k], iters[k],
               theta[k * TH], theta[k * TH + 1 + P], status[k]);
    printf("selected restart %d; X[0] %.6f X[%d] %.6f\n", best[0], X[0], T - 1, X[T - 1]);
    /* the winner has C > 0 when any restart has, and the largest likelihood among those */
    int any_pos = 0;
    for (int k = 0; k < RESTARTS; k++) any_pos |= theta[k * TH + 1 + P] > 0.0;
    for (int k = 0; k < RESTARTS; k++)
        if ((!any_pos || theta[k * TH + 1 + P] > 0.0) && lik[k] > lik[best[0]]) {
            fprintf(stderr, "selection rule violated by restart %d\n", k);
            return 1;
        }
    return 0;
}
