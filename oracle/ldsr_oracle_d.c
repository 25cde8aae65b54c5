/*
 * ldsr_oracle_d.c -- CPU ORACLE for the general-state-dimension Kalman / RTS smoother.
 * TEST INFRASTRUCTURE ONLY (same rules as ldsr_oracle.c: tests/, smoke(), bench cpu legs).
 *
 * The reference is scalar-state only ("matrix inversion is treated as /", /root/reference/
 * src/EM.cpp:20).  BASELINE.json's config 5 (state dimension 4, 20 proxies, T = 100 000) is
 * beyond it, so there is NO reference oracle for d > 1: PARITY UNPINNED BY THE REFERENCE.  What
 * pins this file instead:
 *   (i)  at d = 1 it must reproduce ldsr_oracle_kalman_smoother (which IS pinned to the
 *        reference's golden values) -- tests/test_oracle_golden.py::test_general_d_reduces_to_1d;
 *  (ii)  it is the standard recursion, written with the reference's conventions
 *        (src/EM.cpp:43-124): (mu1,V1) is the PREDICTED law of x_0; u enters with one step of
 *        lag (Xp_t = A Xu_{t-1} + B u_{t-1}), v without (Yp_t = C Xp_t + D v_t); NaN = missing;
 *        innovations-form likelihood divided by n_obs when stdlik.
 * SURVEY.md Appendix D.1 lists the formulas.
 *
 * theta flat: [A d*d row-major | B d*p row-major | C d | D q | Q d*d | R | mu1 d | V1 d*d].
 * u is p x T column-major (u[t*p+j]), v is q x T column-major, y has T entries.
 * Outputs: X [T][d], V [T][d][d], Y [T], lik.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_PI 3.141592653589793238463
#define DMAX 8

/* C = A * B (d x d) */
static void mm(int d, const double *A, const double *B, double *C) {
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            double s = 0.0;
            for (int k = 0; k < d; k++) s += A[i * d + k] * B[k * d + j];
            C[i * d + j] = s;
        }
}
/* C = A * B' */
static void mmt(int d, const double *A, const double *B, double *C) {
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            double s = 0.0;
            for (int k = 0; k < d; k++) s += A[i * d + k] * B[j * d + k];
            C[i * d + j] = s;
        }
}
/* in-place inverse by Gauss-Jordan with partial pivoting; returns 0 on success */
static int inv(int d, double *A) {
    double M[DMAX * 2 * DMAX];
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) {
            M[i * 2 * d + j] = A[i * d + j];
            M[i * 2 * d + d + j] = i == j ? 1.0 : 0.0;
        }
    for (int c = 0; c < d; c++) {
        int piv = c;
        for (int r = c + 1; r < d; r++)
            if (fabs(M[r * 2 * d + c]) > fabs(M[piv * 2 * d + c])) piv = r;
        if (M[piv * 2 * d + c] == 0.0) return 1;
        if (piv != c)
            for (int j = 0; j < 2 * d; j++) {
                double t = M[c * 2 * d + j];
                M[c * 2 * d + j] = M[piv * 2 * d + j];
                M[piv * 2 * d + j] = t;
            }
        const double r = 1.0 / M[c * 2 * d + c];
        for (int j = 0; j < 2 * d; j++) M[c * 2 * d + j] *= r;
        for (int i = 0; i < d; i++)
            if (i != c) {
                const double f = M[i * 2 * d + c];
                if (f != 0.0)
                    for (int j = 0; j < 2 * d; j++) M[i * 2 * d + j] -= f * M[c * 2 * d + j];
            }
    }
    for (int i = 0; i < d; i++)
        for (int j = 0; j < d; j++) A[i * d + j] = M[i * 2 * d + d + j];
    return 0;
}

int ldsr_oracle_smoother_d(int d, int T, int p, int q, const double *y, const double *u, const double *v,
                           const double *theta, int stdlik, double *X, double *V, double *Y, double *lik_out) {
    if (d < 1 || d > DMAX || T < 1) return 1;
    const double *A = theta, *B = A + d * d, *Cc = B + d * p, *D = Cc + d, *Q = D + q;
    const double R = Q[d * d];
    const double *mu1 = Q + d * d + 1, *V1 = mu1 + d;
    const int dd = d * d;
    double *Xp = malloc(sizeof(double) * (size_t)T * d), *Vp = malloc(sizeof(double) * (size_t)T * dd);
    double *Xu = malloc(sizeof(double) * (size_t)T * d), *Vu = malloc(sizeof(double) * (size_t)T * dd);
    if (!Xp || !Vp || !Xu || !Vu) return 2;
    double acc = 0.0;
    long n_obs = 0;
    double tmp[DMAX * DMAX], tmp2[DMAX * DMAX];
    for (int t = 0; t < T; t++) {
        double *xp = Xp + (size_t)t * d, *vp = Vp + (size_t)t * dd, *xu = Xu + (size_t)t * d, *vu = Vu + (size_t)t * dd;
        if (t == 0) {
            memcpy(xp, mu1, sizeof(double) * d);
            memcpy(vp, V1, sizeof(double) * dd);
        } else {
            const double *xu1 = xu - d, *vu1 = vu - dd;
            for (int i = 0; i < d; i++) {
                double s = 0.0;
                for (int k = 0; k < d; k++) s += A[i * d + k] * xu1[k];
                if (u)
                    for (int j = 0; j < p; j++) s += B[i * p + j] * u[(size_t)(t - 1) * p + j];
                xp[i] = s;
            }
            mm(d, A, vu1, tmp);
            mmt(d, tmp, A, vp);
            for (int i = 0; i < dd; i++) vp[i] += Q[i];
        }
        if (isnan(y[t])) {
            memcpy(xu, xp, sizeof(double) * d);
            memcpy(vu, vp, sizeof(double) * dd);
        } else {
            double yp = 0.0;
            for (int k = 0; k < d; k++) yp += Cc[k] * xp[k];
            if (v)
                for (int j = 0; j < q; j++) yp += D[j] * v[(size_t)t * q + j];
            double vc[DMAX]; /* Vp C' */
            double S = R;
            for (int i = 0; i < d; i++) {
                double s = 0.0;
                for (int k = 0; k < d; k++) s += vp[i * d + k] * Cc[k];
                vc[i] = s;
            }
            for (int k = 0; k < d; k++) S += Cc[k] * vc[k];
            const double delta = y[t] - yp;
            for (int i = 0; i < d; i++) xu[i] = xp[i] + vc[i] / S * delta;
            /* Vu = (I - K C) Vp = Vp - (Vp C')(C Vp)/S */
            double cv[DMAX];
            for (int j = 0; j < d; j++) {
                double s = 0.0;
                for (int k = 0; k < d; k++) s += Cc[k] * vp[k * d + j];
                cv[j] = s;
            }
            for (int i = 0; i < d; i++)
                for (int j = 0; j < d; j++) vu[i * d + j] = vp[i * d + j] - vc[i] / S * cv[j];
            acc += delta / S * delta + log(S);
            n_obs++;
        }
    }
    /* RTS */
    memcpy(X + (size_t)(T - 1) * d, Xu + (size_t)(T - 1) * d, sizeof(double) * d);
    memcpy(V + (size_t)(T - 1) * dd, Vu + (size_t)(T - 1) * dd, sizeof(double) * dd);
    for (int t = T - 2; t >= 0; t--) {
        const double *vu = Vu + (size_t)t * dd, *xu = Xu + (size_t)t * d;
        const double *vp1 = Vp + (size_t)(t + 1) * dd, *xp1 = Xp + (size_t)(t + 1) * d;
        double J[DMAX * DMAX], ivp[DMAX * DMAX];
        memcpy(ivp, vp1, sizeof(double) * dd);
        if (inv(d, ivp)) {
            free(Xp); free(Vp); free(Xu); free(Vu);
            return 3;
        }
        mmt(d, vu, A, tmp);  /* Vu A' */
        mm(d, tmp, ivp, J);  /* J = Vu A' Vp1^-1 */
        double *xs = X + (size_t)t * d, *vs = V + (size_t)t * dd;
        const double *xs1 = xs + d, *vs1 = vs + dd;
        for (int i = 0; i < d; i++) {
            double s = xu[i];
            for (int k = 0; k < d; k++) s += J[i * d + k] * (xs1[k] - xp1[k]);
            xs[i] = s;
        }
        for (int i = 0; i < dd; i++) tmp[i] = vs1[i] - vp1[i];
        mm(d, J, tmp, tmp2);
        mmt(d, tmp2, J, tmp);
        for (int i = 0; i < dd; i++) vs[i] = vu[i] + tmp[i];
    }
    for (int t = 0; t < T; t++) {
        double s = 0.0;
        for (int k = 0; k < d; k++) s += Cc[k] * X[(size_t)t * d + k];
        if (v)
            for (int j = 0; j < q; j++) s += D[j] * v[(size_t)t * q + j];
        Y[t] = s;
    }
    double lik = -0.5 * (double)n_obs * log(2.0 * ORACLE_PI) - 0.5 * acc;
    if (stdlik) lik /= (double)n_obs;
    *lik_out = lik;
    free(Xp); free(Vp); free(Xu); free(Vu);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * Cross-validation skill metrics (the step right after the EM hot path in cvLDS).
 * Restates calculate_metrics (/root/reference/R/utils.R:56-70) with the metric definitions of
 * /root/reference/src/utils.cpp:13-97 (NSE :13-19, nRMSE :35-38, corr :48-56, KGE :67-77,
 * RE :93-97).  PINNED: tests/test_oracle_golden.py checks it against the reference's own stored
 * cvLDS result (R/sysdata.rda::NPcv: Ycv, Z, target -> metrics.dist, 30 folds x 5 metrics).
 *   sim    [n_folds][n]  model output per fold (already back-transformed if exp_trans == 0)
 *   obs    [n]           target (NaN allowed: dropped from the calibration part, utils.R:58-59)
 *   z_ptr/z_idx          CSR of the 1-based hold-out indices of every fold
 *   out    [n_folds][5]  R2, RE, CE, nRMSE, KGE
 * exp_trans != 0: sim is exp()'d first (cvLDS with transform = 'log', LDS_reconstruction.R:385-386).
 * --------------------------------------------------------------------------------------------- */
static double mean_of(const double *x, int n) {
    double s = 0.0;
    for (int i = 0; i < n; i++) s += x[i];
    return s / n;
}
static double sd_of(const double *x, int n) {
    const double m = mean_of(x, n);
    double s = 0.0;
    for (int i = 0; i < n; i++) s += (x[i] - m) * (x[i] - m);
    return sqrt(s / (n - 1));
}
static double nse_of(const double *yhat, const double *y, int n) { /* utils.cpp:13-19 */
    const double ybar = mean_of(y, n);
    double rss = 0.0, tss = 0.0;
    for (int i = 0; i < n; i++) {
        rss += (y[i] - yhat[i]) * (y[i] - yhat[i]);
        tss += (y[i] - ybar) * (y[i] - ybar);
    }
    return 1.0 - rss / tss;
}
int ldsr_oracle_cv_metrics(int n, int n_folds, const double *sim, const double *obs, const int *z_ptr,
                           const int *z_idx, int exp_trans, double *out) {
    double *s = malloc(sizeof(double) * n), *ts = malloc(sizeof(double) * n), *to = malloc(sizeof(double) * n);
    double *vs = malloc(sizeof(double) * n), *vo = malloc(sizeof(double) * n);
    char *held = malloc(n);
    if (!s || !ts || !to || !vs || !vo || !held) return 2;
    double osum = 0.0;
    int on = 0;
    for (int i = 0; i < n; i++)
        if (!isnan(obs[i])) {
            osum += obs[i];
            on++;
        }
    const double norm = osum / on; /* norm.fun = mean(obs, na.rm = TRUE), utils.R:67 */
    for (int f = 0; f < n_folds; f++) {
        for (int i = 0; i < n; i++) s[i] = exp_trans ? exp(sim[(size_t)f * n + i]) : sim[(size_t)f * n + i];
        memset(held, 0, n);
        int nv = 0, nt = 0;
        for (int k = z_ptr[f]; k < z_ptr[f + 1]; k++) {
            const int i = z_idx[k] - 1;
            if (i < 0 || i >= n) { free(s); free(ts); free(to); free(vs); free(vo); free(held); return 1; }
            held[i] = 1;
            vs[nv] = s[i];
            vo[nv] = obs[i];
            nv++;
        }
        for (int i = 0; i < n; i++)
            if (!held[i] && !isnan(obs[i])) {
                ts[nt] = s[i];
                to[nt] = obs[i];
                nt++;
            }
        double *o = out + (size_t)f * 5;
        o[0] = nse_of(ts, to, nt);                                   /* R2: NSE on the calibration part */
        {                                                            /* RE, utils.cpp:93-97 */
            const double yc = mean_of(to, nt);
            double rss = 0.0, tss = 0.0;
            for (int i = 0; i < nv; i++) {
                rss += (vo[i] - vs[i]) * (vo[i] - vs[i]);
                tss += (vo[i] - yc) * (vo[i] - yc);
            }
            o[1] = 1.0 - rss / tss;
        }
        o[2] = nse_of(vs, vo, nv);                                   /* CE */
        {                                                            /* nRMSE, utils.cpp:35-38 */
            double ss = 0.0;
            for (int i = 0; i < nv; i++) ss += (vo[i] - vs[i]) * (vo[i] - vs[i]);
            o[3] = sqrt(ss / nv) / norm;
        }
        {                                                            /* KGE, utils.cpp:67-77 with corr :48-56 */
            const double mu = mean_of(vo, nv), muh = mean_of(vs, nv), sg = sd_of(vo, nv), sgh = sd_of(vs, nv);
            double r = 0.0;
            for (int i = 0; i < nv; i++) r += (vs[i] - muh) / sgh * ((vo[i] - mu) / sg);
            r /= (nv - 1);
            const double a = sgh / sg, b = muh / mu;
            o[4] = 1.0 - sqrt((r - 1) * (r - 1) + (a - 1) * (a - 1) + (b - 1) * (b - 1));
        }
    }
    free(s); free(ts); free(to); free(vs); free(vo); free(held);
    return 0;
}

/* ---------------------------------------------------------------------------------------------
 * construct_rec + ensemble averaging (the step right after restart selection in
 * LDS_reconstruction).  Restates /root/reference/R/LDS_reconstruction.R:190-212 (construct_rec),
 * /root/reference/R/utils.R:112-125 (inv_boxcox, exp_ci = qlnorm(c(.05,.95), m, s)) and the
 * year-wise ensemble mean of X and Q (:247-248).  PINNED: tests/test_oracle_golden.py rebuilds all
 * six columns of the reference's stored reconstruction (R/sysdata.rda::NPlds$rec, T = 813).
 *   X, V, Y  [n][T]  smoothed state, its variance, smoothed output (before + mu) of every member
 *   C, R     [n]     each member's theta$C, theta$R
 *   transform 0 = none, 1 = log, 2 = boxcox (lambda)
 *   out      [n][6][T]: X, Xl, Xu, Q, Ql, Qu;   mean [2][T]: ensemble mean of X and of Q
 * --------------------------------------------------------------------------------------------- */
#define QNORM_05 (-1.6448536269514722) /* qnorm(0.05); qlnorm(p, m, s) = exp(m + s qnorm(p)) */
static double inv_boxcox_c(double x, double lambda) { return lambda == 0.0 ? exp(x) : pow(x * lambda + 1.0, 1.0 / lambda); }
int ldsr_oracle_construct_rec(int n, int T, const double *X, const double *V, const double *Y, const double *C,
                              const double *R, double mu, int transform, double lambda, double *out, double *mean) {
    for (int t = 0; t < T; t++) mean[t] = mean[T + t] = 0.0;
    for (int m = 0; m < n; m++) {
        double *o = out + (size_t)m * 6 * T;
        for (int t = 0; t < T; t++) {
            const double x = X[(size_t)m * T + t], v = V[(size_t)m * T + t], y = Y[(size_t)m * T + t] + mu;
            const double ciX = 1.96 * sqrt(v), sdY = sqrt(C[m] * v * C[m] + R[m]), ciY = 1.96 * sdY;
            double q, ql, qu;
            if (transform == 1 || (transform == 2 && lambda == 0.0)) { /* exp_ci */
                q = exp(y);
                ql = exp(y + sdY * QNORM_05);
                qu = exp(y - sdY * QNORM_05);
            } else if (transform == 0) {
                q = y;
                ql = y - ciY;
                qu = y + ciY;
            } else {
                q = inv_boxcox_c(y, lambda);
                ql = inv_boxcox_c(y - ciY, lambda);
                qu = inv_boxcox_c(y + ciY, lambda);
            }
            o[t] = x;
            o[T + t] = x - ciX;
            o[2 * T + t] = x + ciX;
            o[3 * T + t] = q;
            o[4 * T + t] = ql;
            o[5 * T + t] = qu;
            mean[t] += x;
            mean[T + t] += q;
        }
    }
    for (int t = 0; t < 2 * T; t++) mean[t] /= n;
    return 0;
}
