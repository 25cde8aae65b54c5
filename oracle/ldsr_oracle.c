/*
 * ldsr_oracle.c -- CPU ORACLE for the ldsr EM hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is a plain-C restatement (scalar doubles, no Armadillo, no R) of the
 * reference's numeric core, written from the reference's behaviour:
 *     Kalman_smoother   /root/reference/src/EM.cpp:22-131
 *     Mstep             /root/reference/src/EM.cpp:139-229
 *     LDS_EM            /root/reference/src/EM.cpp:245-280
 *     propagate         /root/reference/src/EM.cpp:295-356
 *     restart selection /root/reference/R/LDS_reconstruction.R:50-58
 *     one_LDS_rep       /root/reference/R/stochastics.R:18-47
 * It follows the reference's operation ORDER (which quantity is multiplied by which
 * reciprocal, which sums are formed separately and then added) so that it differs from the
 * real Rcpp/Armadillo build only by BLAS summation order.
 *
 * PARITY PIN: this oracle is checked (tests/test_oracle_golden.py) against every known-answer
 * number the reference's own tests hold for this path (tests/testthat/test-LDS-EM.R:26-40,
 * 11 values incl. the 68-iteration stop) and against the NPlds fixture of R/sysdata.rda
 * (lik, X, Q over T=813 with 767 missing steps).  propagate / replicate / selection have no
 * golden vectors in the reference: "parity unpinned" for those three.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (ldsr_b200/) never does.
 *
 * The third-party arithmetic the reference delegates to (RcppArmadillo, unpinned in
 * DESCRIPTION:28-30; LAPACK getrf/getri behind arma::inv) is restated as: 1x1 inv -> 1.0/x,
 * accu -> sequential sum, n x n inv -> Gauss-Jordan with partial pivoting.
 *
 * theta is passed flat in the reference's own vector order (R/LDS_GA.R:6-16):
 *     [A, B_1..B_p, C, D_1..D_q, Q, R, mu1, V1]       length p+q+6
 * u is p x T column-major (u[t*p+j]), v is q x T column-major, y has T entries, NaN = missing.
 * u == NULL / v == NULL stand for the reference's `matrix(0)` one-column sentinel
 * (EM.cpp:50,71,77,157,189): the input term is dropped and B (resp. D) is returned as zeros.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_PI 3.141592653589793238463 /* EM.cpp:3 */

/* ---- missing-data tests, kept distinct as in the reference ------------------------------ */
/* filter: NumericMatrix::is_na == R_isnancpp (NA or NaN)            EM.cpp:61,82            */
static int y_is_na(double y) { return isnan(y); }
/* lik / M-step: arma::find_finite (also drops +-Inf)                EM.cpp:113,147,339      */
static int y_is_obs(double y) { return isfinite(y); }

typedef struct {
    double A, C, Q, R, mu1, V1;
    const double *B, *D;
} theta_view;

static theta_view view_theta(const double *th, int p, int q) {
    theta_view t;
    t.A = th[0];
    t.B = th + 1;
    t.C = th[1 + p];
    t.D = th + 2 + p;
    t.Q = th[2 + p + q];
    t.R = th[3 + p + q];
    t.mu1 = th[4 + p + q];
    t.V1 = th[5 + p + q];
    return t;
}

static double dotn(const double *a, const double *b, int n) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += a[i] * b[i];
    return s;
}

/* n x n general inverse, Gauss-Jordan with partial pivoting (stands in for arma::inv).
 * a is overwritten; inv receives the inverse (row-major; the matrices here are symmetric so
 * the storage order is immaterial).  Returns 0, or 1 if a pivot is exactly zero / non-finite. */
static int inv_general(int n, double *a, double *inv) {
    for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) inv[i * n + j] = (i == j) ? 1.0 : 0.0;
    for (int c = 0; c < n; c++) {
        int piv = c;
        double best = fabs(a[c * n + c]);
        for (int r = c + 1; r < n; r++)
            if (fabs(a[r * n + c]) > best) {
                best = fabs(a[r * n + c]);
                piv = r;
            }
        if (!(best > 0.0) || !isfinite(best)) return 1;
        if (piv != c)
            for (int j = 0; j < n; j++) {
                double t = a[c * n + j];
                a[c * n + j] = a[piv * n + j];
                a[piv * n + j] = t;
                t = inv[c * n + j];
                inv[c * n + j] = inv[piv * n + j];
                inv[piv * n + j] = t;
            }
        double d = 1.0 / a[c * n + c];
        for (int j = 0; j < n; j++) {
            a[c * n + j] *= d;
            inv[c * n + j] *= d;
        }
        for (int r = 0; r < n; r++) {
            if (r == c) continue;
            double f = a[r * n + c];
            if (f == 0.0) continue;
            for (int j = 0; j < n; j++) {
                a[r * n + j] -= f * a[c * n + j];
                inv[r * n + j] -= f * inv[c * n + j];
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * E-step.  EM.cpp:22-131.  Work arrays Xp,Vp,Yp,Xu,Vu (T each) are caller-provided so the EM
 * loop does not allocate.  Outputs X(=Xs), Y(=Ys), V(=Vs), J (T each) and *lik.
 * ---------------------------------------------------------------------------------------- */
static void smoother_core(const double *y, const double *u, const double *v, int T, int p, int q,
                          const double *th, int stdlik, double *Xp, double *Vp, double *Yp,
                          double *Xu, double *Vu, double *X, double *Y, double *V, double *J,
                          double *lik_out) {
    theta_view t = view_theta(th, p, q);
    const double A = t.A, C = t.C, Q = t.Q, R = t.R;

    /* prior at the first step                                           EM.cpp:48-54 */
    Xp[0] = t.mu1;
    Vp[0] = t.V1;
    Yp[0] = v ? C * Xp[0] + dotn(t.D, v, q) : C * Xp[0];
    for (int k = 0; k < T; k++) {
        if (k > 0) {
            /* u enters with a one-step lag, v without                    EM.cpp:71-81 */
            Xp[k] = u ? A * Xu[k - 1] + dotn(t.B, u + (size_t)(k - 1) * p, p) : A * Xu[k - 1];
            Vp[k] = A * Vu[k - 1] * A + Q;
            Yp[k] = v ? C * Xp[k] + dotn(t.D, v + (size_t)k * q, q) : C * Xp[k];
        }
        if (y_is_na(y[k])) { /* EM.cpp:61-63, 82-84 */
            Xu[k] = Xp[k];
            Vu[k] = Vp[k];
        } else { /* EM.cpp:65-67, 86-88: K = Vp*C*inv(C*Vp*C+R) */
            double K = Vp[k] * C * (1.0 / (C * Vp[k] * C + R));
            Xu[k] = Xp[k] + K * (y[k] - Yp[k]);
            Vu[k] = (1.0 - K * C) * Vp[k];
        }
    }

    /* RTS backward sweep                                                 EM.cpp:94-104 */
    memcpy(X, Xu, sizeof(double) * T);
    memcpy(V, Vu, sizeof(double) * T);
    J[T - 1] = Vu[T - 1] * A * (1.0 / (A * Vu[T - 1] * A + Q));
    for (int k = T - 2; k >= 0; k--) {
        J[k] = Vu[k] * A * (1.0 / Vp[k + 1]);
        X[k] = Xu[k] + J[k] * (X[k + 1] - Xp[k + 1]);
        V[k] = Vu[k] + J[k] * (V[k + 1] - Vp[k + 1]) * J[k];
    }
    /* smoothed output                                                    EM.cpp:106-110 */
    for (int k = 0; k < T; k++) Y[k] = v ? C * X[k] + dotn(t.D, v + (size_t)k * q, q) : C * X[k];

    /* innovations-form log-likelihood                                    EM.cpp:113-124 */
    int n_obs = 0;
    double acc = 0.0;
    for (int k = 0; k < T; k++) {
        if (!y_is_obs(y[k])) continue;
        double delta = y[k] - Yp[k];
        double Sigma = C * Vp[k] * C + R;
        acc += delta / Sigma * delta + log(Sigma);
        n_obs++;
    }
    double lik = -0.5 * n_obs * log(2 * ORACLE_PI) - 0.5 * acc;
    if (stdlik) lik = lik / n_obs;
    *lik_out = lik;
}

/* θ-independent Gram blocks of the M-step (EM.cpp:158,161,193): Svv, Syv over observed steps,
 * Tuu over t=0..T-2.  They are recomputed by every reference Mstep call; caching them is
 * bit-identical and only makes the timed CPU baseline faster (fairer). */
typedef struct {
    int valid;
    double *Svv, *Syv, *Tuu;
} gram_cache;

/* M-step.  EM.cpp:139-229.  Returns 0, or 1 when an inverse does not exist. */
static int mstep_core(const double *y, const double *u, const double *v, int T, int p, int q,
                      const double *X, const double *V, const double *J, double *th_out,
                      gram_cache *gc, double *scratch /* >= 2*(max(p,q)+1)^2 + 2*(max(p,q)+1) */) {
    const int n2 = (p > q ? p : q) + 1;
    double *P = scratch, *Pinv = scratch + n2 * n2, *rowv = scratch + 2 * n2 * n2,
           *sol = rowv + n2;

    double *Bo = th_out + 1, *Do = th_out + 2 + p;
    /* ---- C, D, R over observed steps                                   EM.cpp:151-177 */
    double Syx = 0.0, Sxx_xx = 0.0, Sxx_v = 0.0;
    int n_obs = 0;
    for (int k = 0; k < T; k++)
        if (y_is_obs(y[k])) {
            Syx += y[k] * X[k];
            Sxx_xx += X[k] * X[k];
            Sxx_v += V[k];
            n_obs++;
        }
    const double Sxx = Sxx_xx + Sxx_v;
    double Cn;
    if (v) {
        if (!gc->valid) {
            memset(gc->Svv, 0, sizeof(double) * q * q);
            memset(gc->Syv, 0, sizeof(double) * q);
            for (int k = 0; k < T; k++)
                if (y_is_obs(y[k])) {
                    const double *vk = v + (size_t)k * q;
                    for (int a = 0; a < q; a++) {
                        gc->Syv[a] += y[k] * vk[a];
                        for (int b = 0; b < q; b++) gc->Svv[a * q + b] += vk[a] * vk[b];
                    }
                }
        }
        const int n = q + 1;
        double *Sxv = sol; /* reuse */
        for (int a = 0; a < q; a++) Sxv[a] = 0.0;
        for (int k = 0; k < T; k++)
            if (y_is_obs(y[k])) {
                const double *vk = v + (size_t)k * q;
                for (int a = 0; a < q; a++) Sxv[a] += X[k] * vk[a];
            }
        /* P2 = [[Sxx,Sxv],[Svx,Svv]],  CD = [Syx Syv] * inv(P2)          EM.cpp:164-169 */
        P[0] = Sxx;
        for (int a = 0; a < q; a++) {
            P[0 * n + 1 + a] = Sxv[a];
            P[(1 + a) * n + 0] = Sxv[a];
            for (int b = 0; b < q; b++) P[(1 + a) * n + 1 + b] = gc->Svv[a * q + b];
        }
        rowv[0] = Syx;
        for (int a = 0; a < q; a++) rowv[1 + a] = gc->Syv[a];
        if (inv_general(n, P, Pinv)) return 1;
        for (int c = 0; c < n; c++) {
            double s = 0.0;
            for (int r = 0; r < n; r++) s += rowv[r] * Pinv[r * n + c];
            sol[c] = s;
        }
        Cn = sol[0];
        for (int a = 0; a < q; a++) Do[a] = sol[1 + a];
    } else {
        Cn = Syx * (1.0 / Sxx); /* EM.cpp:172 */
        for (int a = 0; a < q; a++) Do[a] = 0.0;
    }
    /* R = (y - y_hat) * y' / n_obs   (one-sided form)                    EM.cpp:170,173,177 */
    double Racc = 0.0;
    for (int k = 0; k < T; k++)
        if (y_is_obs(y[k])) {
            double yhat = v ? Cn * X[k] + dotn(Do, v + (size_t)k * q, q) : Cn * X[k];
            Racc += (y[k] - yhat) * y[k];
        }
    const double Rn = Racc / n_obs;

    /* ---- A, B, Q over all T-1 transitions                              EM.cpp:180-214 */
    double xx1 = 0.0, vj = 0.0, xx_lo = 0.0, v_lo = 0.0, xx_hi = 0.0, v_hi = 0.0;
    for (int k = 0; k < T - 1; k++) {
        xx1 += X[k + 1] * X[k];
        vj += V[k + 1] * J[k];
        xx_lo += X[k] * X[k];
        v_lo += V[k];
        xx_hi += X[k + 1] * X[k + 1];
        v_hi += V[k + 1];
    }
    const double Tx1x = xx1 + vj, Txx = xx_lo + v_lo, Tx1x1 = xx_hi + v_hi;
    double An = 0.0, Qn;
    if (u) {
        if (!gc->valid) {
            memset(gc->Tuu, 0, sizeof(double) * p * p);
            for (int k = 0; k < T - 1; k++) {
                const double *uk = u + (size_t)k * p;
                for (int a = 0; a < p; a++)
                    for (int b = 0; b < p; b++) gc->Tuu[a * p + b] += uk[a] * uk[b];
            }
        }
        const int n = p + 1;
        double *Tx1u = rowv + 1, *Tux = sol; /* rowv = [Tx1x, Tx1u] is P3 */
        for (int a = 0; a < p; a++) Tx1u[a] = Tux[a] = 0.0;
        for (int k = 0; k < T - 1; k++) {
            const double *uk = u + (size_t)k * p;
            for (int a = 0; a < p; a++) {
                Tx1u[a] += X[k + 1] * uk[a];
                Tux[a] += uk[a] * X[k];
            }
        }
        rowv[0] = Tx1x;
        P[0] = Txx;
        for (int a = 0; a < p; a++) {
            P[0 * n + 1 + a] = Tux[a];
            P[(1 + a) * n + 0] = Tux[a];
            for (int b = 0; b < p; b++) P[(1 + a) * n + 1 + b] = gc->Tuu[a * p + b];
        }
        if (inv_general(n, P, Pinv)) return 1;
        /* AB = P3 * inv(P4)                                              EM.cpp:196-201 */
        double Bq = 0.0;
        for (int c = 0; c < n; c++) {
            double s = 0.0;
            for (int r = 0; r < n; r++) s += rowv[r] * Pinv[r * n + c];
            if (c == 0)
                An = s;
            else
                Bo[c - 1] = s;
        }
        for (int a = 0; a < p; a++) Bq += Bo[a] * Tx1u[a];
        Qn = (Tx1x1 - An * Tx1x - Bq) / (T - 1); /* EM.cpp:210 */
    } else {
        An = Tx1x * (1.0 / Txx); /* EM.cpp:212-213 */
        Qn = (Tx1x1 - An * Tx1x) / (T - 1);
        for (int a = 0; a < p; a++) Bo[a] = 0.0;
    }
    gc->valid = 1;
    th_out[0] = An;
    th_out[1 + p] = Cn;
    th_out[2 + p + q] = Qn;
    th_out[3 + p + q] = Rn;
    th_out[4 + p + q] = X[0]; /* mu1, V1                                  EM.cpp:218-219 */
    th_out[5 + p + q] = V[0];
    return 0;
}

/* ---- workspace ------------------------------------------------------------------------- */
typedef struct {
    double *buf, *Xp, *Vp, *Yp, *Xu, *Vu, *X, *Y, *V, *J, *scratch, *th, *th2;
    gram_cache gc;
} workspace;

static int ws_init(workspace *w, int T, int p, int q) {
    const int n2 = (p > q ? p : q) + 1;
    size_t n = (size_t)9 * T + 2 * n2 * n2 + 2 * n2 + 2 * (size_t)(p + q + 6) + (size_t)q * q +
               q + (size_t)p * p + 8;
    w->buf = (double *)malloc(sizeof(double) * n);
    if (!w->buf) return 1;
    double *b = w->buf;
    w->Xp = b; b += T;
    w->Vp = b; b += T;
    w->Yp = b; b += T;
    w->Xu = b; b += T;
    w->Vu = b; b += T;
    w->X = b; b += T;
    w->Y = b; b += T;
    w->V = b; b += T;
    w->J = b; b += T;
    w->scratch = b; b += 2 * n2 * n2 + 2 * n2;
    w->th = b; b += p + q + 6;
    w->th2 = b; b += p + q + 6;
    w->gc.Svv = b; b += (size_t)q * q;
    w->gc.Syv = b; b += q;
    w->gc.Tuu = b;
    w->gc.valid = 0;
    return 0;
}
static void ws_free(workspace *w) { free(w->buf); }

/* ==========================================================================================
 * Exported functions.  All return 0 on success.
 * ======================================================================================== */

/* Kalman_smoother(y,u,v,theta,stdlik)  ->  X,Y,V,J (T each), lik.            EM.cpp:22 */
int ldsr_oracle_kalman_smoother(const double *y, const double *u, const double *v, int T, int p,
                                int q, const double *theta, int stdlik, double *X, double *Y,
                                double *V, double *J, double *lik) {
    workspace w;
    if (T < 1 || ws_init(&w, T, p, q)) return 2;
    smoother_core(y, u, v, T, p, q, theta, stdlik, w.Xp, w.Vp, w.Yp, w.Xu, w.Vu, X, Y, V, J, lik);
    ws_free(&w);
    return 0;
}

/* Mstep(y,u,v,fit{X,V,J}) -> theta.                                          EM.cpp:139 */
int ldsr_oracle_mstep(const double *y, const double *u, const double *v, int T, int p, int q,
                      const double *X, const double *V, const double *J, double *theta_out) {
    workspace w;
    if (T < 2 || ws_init(&w, T, p, q)) return 2;
    int rc = mstep_core(y, u, v, T, p, q, X, V, J, theta_out, &w.gc, w.scratch);
    ws_free(&w);
    return rc;
}

/* LDS_EM(y,u,v,theta0,niter,tol).                                            EM.cpp:245-280
 * Outputs: theta_out (the theta the LAST E-step was run with), X,Y,V,J of that E-step,
 * liks[0..*n_liks) (caller provides niter doubles; may be NULL), *lik = liks[*n_liks-1].
 * niter >= 2 is required exactly as in the reference (it writes lik[1] unconditionally). */
static int em_core(workspace *w, const double *y, const double *u, const double *v, int T, int p,
                   int q, const double *theta0, int niter, double tol, double *theta_out,
                   double *liks, int *n_liks, double *lik_out) {
    const int nth = p + q + 6;
    double l0, l1, l2 = 0.0; /* lik[i], lik[i-1], lik[i-2] */
    double *th = w->th, *thn = w->th2;
    memcpy(th, theta0, sizeof(double) * nth);
    w->gc.valid = 0;
    /* i = 0                                                              EM.cpp:251-253 */
    smoother_core(y, u, v, T, p, q, th, 1, w->Xp, w->Vp, w->Yp, w->Xu, w->Vu, w->X, w->Y, w->V,
                  w->J, &l0);
    if (liks) liks[0] = l0;
    if (mstep_core(y, u, v, T, p, q, w->X, w->V, w->J, thn, &w->gc, w->scratch)) return 1;
    { double *s = th; th = thn; thn = s; }
    /* i = 1                                                              EM.cpp:255-256 */
    l1 = l0;
    smoother_core(y, u, v, T, p, q, th, 1, w->Xp, w->Vp, w->Yp, w->Xu, w->Vu, w->X, w->Y, w->V,
                  w->J, &l0);
    if (liks) liks[1] = l0;
    int last = 2;
    for (int i = 2; i < niter; i++) { /* EM.cpp:259-275 */
        if (mstep_core(y, u, v, T, p, q, w->X, w->V, w->J, thn, &w->gc, w->scratch)) return 1;
        { double *s = th; th = thn; thn = s; }
        l2 = l1;
        l1 = l0;
        smoother_core(y, u, v, T, p, q, th, 1, w->Xp, w->Vp, w->Yp, w->Xu, w->Vu, w->X, w->Y,
                      w->V, w->J, &l0);
        if (liks) liks[i] = l0;
        last++;
        if (fabs(l0 - l1) < tol && fabs(l1 - l2) < tol) break; /* EM.cpp:272 */
    }
    memcpy(theta_out, th, sizeof(double) * nth);
    *n_liks = last;
    *lik_out = l0;
    return 0;
}

int ldsr_oracle_em(const double *y, const double *u, const double *v, int T, int p, int q,
                   const double *theta0, int niter, double tol, double *theta_out, double *X,
                   double *Y, double *V, double *J, double *liks, int *n_liks, double *lik) {
    workspace w;
    if (T < 2 || niter < 2 || ws_init(&w, T, p, q)) return 2;
    int rc = em_core(&w, y, u, v, T, p, q, theta0, niter, tol, theta_out, liks, n_liks, lik);
    if (!rc) {
        if (X) memcpy(X, w.X, sizeof(double) * T);
        if (Y) memcpy(Y, w.Y, sizeof(double) * T);
        if (V) memcpy(V, w.V, sizeof(double) * T);
        if (J) memcpy(J, w.J, sizeof(double) * T);
    }
    ws_free(&w);
    return rc;
}

/* LDS_EM_restart's selection rule.                      R/LDS_reconstruction.R:50-58
 * liks[n], C[n] -> index of the chosen model, or -1 when R would fail (no usable lik).
 *   posC = which(C > 0); if any: first i with liks[i] == max(liks[posC], na.rm) (we restrict
 *   the tie scan to C>0 fits; an exact tie with a C<=0 fit makes the reference index with a
 *   vector and misbehave), else which.max(liks) (first maximum, NaN skipped). */
int ldsr_oracle_select(const double *liks, const double *C, int n) {
    int best = -1, any_pos = 0;
    for (int i = 0; i < n; i++)
        if (C[i] > 0) {
            any_pos = 1;
            if (!isnan(liks[i]) && (best < 0 || liks[i] > liks[best])) best = i;
        }
    if (any_pos) return best;
    for (int i = 0; i < n; i++)
        if (!isnan(liks[i]) && (best < 0 || liks[i] > liks[best])) best = i;
    return best;
}

/* Batched EM with hold-out folds and per-group selection: the CPU statement of what
 * LDS_EM_restart / one_lds_cv / cvLDS do through foreach (R/LDS_reconstruction.R:46,274,373).
 *   series s: T[s], p[s], q[s], y/u/v pointers (u,v may be NULL)
 *   group g : series id, held-out step indices (0-based, into 0..T-1) in CSR form
 *   fit f   : group id, theta0 (flat, stride th_stride)
 * Outputs per fit: theta (stride th_stride), lik, iters (=length(liks)), status (0 ok,
 * 1 singular M-step); per group: best fit index (global fit index, -1 if none).
 * Fits of one group must be contiguous in the fit table.  OpenMP over fits. */
int ldsr_oracle_em_batch(int n_series, const int *T, const int *p, const int *q,
                         const double *const *y, const double *const *u, const double *const *v,
                         int n_groups, const int *group_series, const int *held_ptr,
                         const int *held_idx, int n_fits, const int *fit_group,
                         const double *theta0, int th_stride, int niter, double tol,
                         double *theta_out, double *lik_out, int *iters_out, int *status_out,
                         int *best_out, int n_threads) {
    (void)n_series;
    if (niter < 2) return 2;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#else
    (void)n_threads;
#endif
    int fail = 0;
#pragma omp parallel
    {
        double *yfold = NULL;
        int ycap = 0, cur_group = -1;
#pragma omp for schedule(dynamic, 4)
        for (int f = 0; f < n_fits; f++) {
            const int g = fit_group[f], s = group_series[g];
            const int Ts = T[s], ps = p[s], qs = q[s];
            if (Ts > ycap) {
                free(yfold);
                yfold = (double *)malloc(sizeof(double) * Ts);
                ycap = Ts;
                cur_group = -1;
            }
            if (g != cur_group) { /* y[instPeriod][z] <- NA   R/LDS_reconstruction.R:274 */
                memcpy(yfold, y[s], sizeof(double) * Ts);
                for (int k = held_ptr[g]; k < held_ptr[g + 1]; k++) yfold[held_idx[k]] = NAN;
                cur_group = g;
            }
            workspace w;
            if (ws_init(&w, Ts, ps, qs)) {
                fail = 1;
                continue;
            }
            int nl = 0;
            double lk = NAN;
            int rc = em_core(&w, yfold, u[s], v[s], Ts, ps, qs, theta0 + (size_t)f * th_stride,
                             niter, tol, theta_out + (size_t)f * th_stride, NULL, &nl, &lk);
            lik_out[f] = rc ? NAN : lk;
            iters_out[f] = nl;
            status_out[f] = rc;
            ws_free(&w);
        }
        free(yfold);
    }
    if (fail) return 2;
    if (best_out) {
        int f0 = 0;
        for (int g = 0; g < n_groups; g++) best_out[g] = -1;
        while (f0 < n_fits) {
            int g = fit_group[f0], f1 = f0;
            while (f1 < n_fits && fit_group[f1] == g) f1++;
            const int n = f1 - f0, s = group_series[g];
            double *Cs = (double *)malloc(sizeof(double) * n);
            for (int i = 0; i < n; i++) Cs[i] = theta_out[(size_t)(f0 + i) * th_stride + 1 + p[s]];
            int b = ldsr_oracle_select(lik_out + f0, Cs, n);
            best_out[g] = b < 0 ? -1 : f0 + b;
            free(Cs);
            f0 = f1;
        }
    }
    return 0;
}

/* propagate(theta,u,v,y,stdlik): open-loop prediction.                       EM.cpp:295-356
 * T is the number of columns of u in the reference (EM.cpp:308); callers pass it. */
int ldsr_oracle_propagate(const double *theta, const double *u, const double *v, const double *y,
                          int T, int p, int q, int stdlik, double *X, double *Y, double *V,
                          double *lik_out) {
    if (T < 1) return 2;
    theta_view t = view_theta(theta, p, q);
    X[0] = t.mu1;
    V[0] = t.V1;
    for (int k = 1; k < T; k++) {
        X[k] = u ? t.A * X[k - 1] + dotn(t.B, u + (size_t)(k - 1) * p, p) : t.A * X[k - 1];
        V[k] = t.A * V[k - 1] * t.A + t.Q;
    }
    for (int k = 0; k < T; k++)
        Y[k] = v ? t.C * X[k] + dotn(t.D, v + (size_t)k * q, q) : t.C * X[k];
    int n_obs = 0;
    double acc = 0.0;
    for (int k = 0; k < T; k++) {
        if (!y_is_obs(y[k])) continue;
        double delta = y[k] - Y[k];
        double Sigma = t.C * V[k] * t.C + t.R;
        acc += delta / Sigma * delta + log(Sigma);
        n_obs++;
    }
    double lik = -0.5 * n_obs * log(2 * ORACLE_PI) - 0.5 * acc;
    if (stdlik) lik = lik / n_obs;
    *lik_out = lik;
    return 0;
}

/* one_LDS_rep with caller-supplied standard-normal draws.          R/stochastics.R:18-47
 * z has 1 + 2n entries in the reference's draw order: z[0] for x_1 (mean 0, sd sqrt(V1) --
 * NOT mu1, stochastics.R:23), z[1..n] state noise, z[n+1..2n] observation noise.
 * simX[t], simY[t], simQ[t] for t = 0..n-1;  simQ = exp(simY+mu) if exp_trans else simY+mu. */
int ldsr_oracle_rep(const double *theta, const double *u, const double *v, int n, int p, int q,
                    const double *z, double mu, int exp_trans, double *simX, double *simY,
                    double *simQ) {
    theta_view t = view_theta(theta, p, q);
    const double sQ = sqrt(t.Q), sR = sqrt(t.R);
    double x = z[0] * sqrt(t.V1);
    for (int k = 0; k < n; k++) {
        double qn = z[1 + k] * sQ, rn = z[1 + n + k] * sR;
        simX[k] = x;
        double yk = t.C * x + (v ? dotn(t.D, v + (size_t)k * q, q) : 0.0) + rn;
        simY[k] = yk;
        simQ[k] = exp_trans ? exp(yk + mu) : yk + mu;
        x = t.A * x + (u ? dotn(t.B, u + (size_t)k * p, p) : 0.0) + qn;
    }
    return 0;
}

int ldsr_oracle_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
