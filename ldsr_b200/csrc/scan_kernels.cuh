// scan_kernels.cuh -- Kalman filter / RTS smoother for state dimension D > 1 and long series,
// sequential or as an ASSOCIATIVE SCAN over time (BASELINE.json config 5: d = 4, 20 proxies,
// T = 100 000).  Beyond the reference, which is scalar-state only (src/EM.cpp:20); conventions are
// the reference's (src/EM.cpp:43-124): (mu1,V1) is the predicted law of x_0, u enters with one
// step of lag, v without, NaN = missing, scalar y.  Formulas: SURVEY.md Appendix D (Saerkkae &
// Garcia-Fernandez, "Temporal parallelization of Bayesian smoothers", IEEE TAC 2021).
//
// Time is cut into chunks of L steps, one THREAD per (fit, chunk):
//   prep        c_t = B u_{t-1}, dv_t = D v_t for every t (one thread per step, coalesced rows)
//   filt_agg    each chunk's filtering element a = (A,b,C,eta,J), combined step by step
//   filt_scan   one CTA per fit, three levels: prefix over the chunk elements -> (Xu,Vu) entering every chunk
//   filt_down   ordinary Kalman filter inside each chunk from that state: Xu_t, Vu_t, lik terms
//   smth_agg    each chunk's smoothing element (E,g,L) from Xu,Vu
//   smth_scan   suffix over the chunk elements -> (Xs,Vs) entering every chunk from the right
//   smth_down   ordinary RTS recursion inside each chunk: X, V, Y
// With L >= T there is one chunk per fit and the two "down" kernels ARE the sequential recursion
// (method 0); the scan kernels are skipped.
#pragma once
#include "common.cuh"

namespace ldsr {

struct ScanParams {
    int n_fits, T, p, q, L, n_chunks, stdlik;
    int theta_len;
    const double *y;     // [T]
    const double *u;     // [T][p] or NULL
    const double *v;     // [T][q] or NULL
    const double *theta; // [n_fits][theta_len]
    double *c;           // [n_fits][T][D]   c_t = B u_{t-1} (c_0 = 0)
    double *dv;          // [n_fits][T]      D v_t
    double *Xu, *Vu;     // [n_fits][T][D], [n_fits][T][D*D]
    double *fagg;        // [n_fits][n_chunks][3D^2+2D]
    double *sagg;        // [n_fits][n_chunks][2D^2+D]
    double *pre;         // [n_fits][n_chunks][D+D^2]  filtered state entering the chunk
    double *suf;         // [n_fits][n_chunks][D+D^2]  smoothed state right of the chunk
    double *likp;        // [n_fits][n_chunks][2]      (sum of terms, n_obs)
    // long series: the chunk elements are scanned by n_groups CTAs per fit (SCAN_GROUPS), in three launches
    int n_groups;
    double *fgagg, *fgpre; // [n_fits][n_groups][3D^2+2D] totals of the groups, prefix before each group
    double *sgagg, *sgsuf; // [n_fits][n_groups][2D^2+D]  totals of the groups, suffix right of each group
    double *X, *V, *Y;   // outputs [n_fits][T][D], [n_fits][T][D*D], [n_fits][T]
    double *lik;         // [n_fits]
};

template <int D> struct ThetaD {
    const double *A, *B, *C, *Dd, *Q, *mu1, *V1;
    double R;
    __device__ __forceinline__ ThetaD(const double *th, int p, int q) {
        A = th;
        B = A + D * D;
        C = B + D * p;
        Dd = C + D;
        Q = Dd + q;
        R = Q[D * D];
        mu1 = Q + D * D + 1;
        V1 = mu1 + D;
    }
};

// ---- small dense helpers, fully unrolled, everything in registers -----------------------------
template <int D> __device__ __forceinline__ void mat_load(const double *__restrict__ g, double (&m)[D * D]) {
#pragma unroll
    for (int i = 0; i < D * D; i++) m[i] = g[i];
}
template <int D> __device__ __forceinline__ void vec_load(const double *__restrict__ g, double (&x)[D]) {
#pragma unroll
    for (int i = 0; i < D; i++) x[i] = g[i];
}
template <int D>
__device__ __forceinline__ void mm(const double (&A)[D * D], const double (&B)[D * D], double (&C)[D * D]) {
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j < D; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) s = fma(A[i * D + k], B[k * D + j], s);
            C[i * D + j] = s;
        }
}
template <int D> // C = A * B'
__device__ __forceinline__ void mmt(const double (&A)[D * D], const double (&B)[D * D], double (&C)[D * D]) {
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j < D; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) s = fma(A[i * D + k], B[j * D + k], s);
            C[i * D + j] = s;
        }
}
template <int D> // C = A' * B
__device__ __forceinline__ void mtm(const double (&A)[D * D], const double (&B)[D * D], double (&C)[D * D]) {
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j < D; j++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < D; k++) s = fma(A[k * D + i], B[k * D + j], s);
            C[i * D + j] = s;
        }
}
template <int D> __device__ __forceinline__ void mv(const double (&A)[D * D], const double (&x)[D], double (&y)[D]) {
#pragma unroll
    for (int i = 0; i < D; i++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) s = fma(A[i * D + k], x[k], s);
        y[i] = s;
    }
}
template <int D> // y = A' x
__device__ __forceinline__ void mtv(const double (&A)[D * D], const double (&x)[D], double (&y)[D]) {
#pragma unroll
    for (int i = 0; i < D; i++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) s = fma(A[k * D + i], x[k], s);
        y[i] = s;
    }
}
// inverse by Gauss-Jordan with partial pivoting; rows are swapped with selects so every index is
// a compile-time constant and the matrix stays in registers
template <int D> __device__ __forceinline__ void mat_inv(const double (&A)[D * D], double (&R)[D * D]) {
    double M[D][2 * D];
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j < D; j++) {
            M[i][j] = A[i * D + j];
            M[i][D + j] = i == j ? 1.0 : 0.0;
        }
#pragma unroll
    for (int c = 0; c < D; c++) {
#pragma unroll
        for (int r = c + 1; r < D; r++) { // bring the largest |entry| of column c up to row c
            const bool sw = fabs(M[r][c]) > fabs(M[c][c]);
#pragma unroll
            for (int j = 0; j < 2 * D; j++) {
                const double a = M[c][j], b = M[r][j];
                M[c][j] = sw ? b : a;
                M[r][j] = sw ? a : b;
            }
        }
        const double rp = 1.0 / M[c][c];
#pragma unroll
        for (int j = 0; j < 2 * D; j++) M[c][j] *= rp;
#pragma unroll
        for (int i = 0; i < D; i++)
            if (i != c) {
                const double f = M[i][c];
#pragma unroll
                for (int j = 0; j < 2 * D; j++) M[i][j] = fma(-f, M[c][j], M[i][j]);
            }
    }
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j < D; j++) R[i * D + j] = M[i][D + j];
}

// ---- filtering element (Appendix D.2) ---------------------------------------------------------
template <int D> struct FiltElem {
    double A[D * D], b[D], C[D * D], eta[D], J[D * D];
    static constexpr int LEN = 3 * D * D + 2 * D;
    __device__ __forceinline__ void store(double *__restrict__ g) const {
#pragma unroll
        for (int i = 0; i < D * D; i++) {
            g[i] = A[i];
            g[D * D + D + i] = C[i];
            g[2 * D * D + 2 * D + i] = J[i];
        }
#pragma unroll
        for (int i = 0; i < D; i++) {
            g[D * D + i] = b[i];
            g[2 * D * D + D + i] = eta[i];
        }
    }
    __device__ __forceinline__ void load(const double *__restrict__ g) {
#pragma unroll
        for (int i = 0; i < D * D; i++) {
            A[i] = g[i];
            C[i] = g[D * D + D + i];
            J[i] = g[2 * D * D + 2 * D + i];
        }
#pragma unroll
        for (int i = 0; i < D; i++) {
            b[i] = g[D * D + i];
            eta[i] = g[2 * D * D + D + i];
        }
    }
};

// the neutral element of filt_combine: (I, 0, 0, 0, 0) -- both combine(a, id) = a and combine(id, e) = e hold
// exactly (the inverse taken is that of the identity)
template <int D> __device__ __forceinline__ void filt_identity(FiltElem<D> &a) {
#pragma unroll
    for (int i = 0; i < D * D; i++) {
        a.A[i] = (i % (D + 1) == 0) ? 1.0 : 0.0;
        a.C[i] = 0.0;
        a.J[i] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < D; i++) a.b[i] = a.eta[i] = 0.0;
}
// a <- a (earlier) combined with e (later)
template <int D> __device__ __forceinline__ void filt_combine(FiltElem<D> &a, const FiltElem<D> &e) {
    double T1[D * D], M[D * D], N[D * D];
    mm<D>(a.C, e.J, T1); // C_i J_j
#pragma unroll
    for (int i = 0; i < D; i++) T1[i * D + i] += 1.0;
    mat_inv<D>(T1, M); // M = (I + C_i J_j)^-1 ;  N = (I + J_j C_i)^-1 = M'
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j < D; j++) N[i * D + j] = M[j * D + i];
    double AjM[D * D];
    mm<D>(e.A, M, AjM);
    // b = A_j M (b_i + C_i eta_j) + b_j
    double t[D], t2[D], nb[D];
    mv<D>(a.C, e.eta, t);
#pragma unroll
    for (int i = 0; i < D; i++) t[i] += a.b[i];
    mv<D>(AjM, t, nb);
#pragma unroll
    for (int i = 0; i < D; i++) nb[i] += e.b[i];
    // eta = A_i' N (eta_j - J_j b_i) + eta_i
    mv<D>(e.J, a.b, t);
#pragma unroll
    for (int i = 0; i < D; i++) t[i] = e.eta[i] - t[i];
    mv<D>(N, t, t2);
    double neta[D];
    mtv<D>(a.A, t2, neta);
#pragma unroll
    for (int i = 0; i < D; i++) neta[i] += a.eta[i];
    // C = A_j M C_i A_j' + C_j
    double T2[D * D], nC[D * D];
    mm<D>(AjM, a.C, T2);
    mmt<D>(T2, e.A, nC);
#pragma unroll
    for (int i = 0; i < D * D; i++) nC[i] += e.C[i];
    // J = A_i' N J_j A_i + J_i
    double nJ[D * D];
    mm<D>(N, e.J, T1);
    mm<D>(T1, a.A, T2);
    mtm<D>(a.A, T2, nJ);
#pragma unroll
    for (int i = 0; i < D * D; i++) nJ[i] += a.J[i];
    // A = A_j M A_i
    double nA[D * D];
    mm<D>(AjM, a.A, nA);
#pragma unroll
    for (int i = 0; i < D * D; i++) {
        a.A[i] = nA[i];
        a.C[i] = 0.5 * (nC[i] + nC[(i % D) * D + i / D]); // keep the covariances symmetric
        a.J[i] = 0.5 * (nJ[i] + nJ[(i % D) * D + i / D]);
    }
#pragma unroll
    for (int i = 0; i < D; i++) {
        a.b[i] = nb[i];
        a.eta[i] = neta[i];
    }
}

// measurement update of (x,V) with scalar y (EM.cpp:61-68 generalised); returns the likelihood term
template <int D>
__device__ __forceinline__ double meas_update(const double (&Cc)[D], double R, double r, double (&x)[D],
                                              double (&V)[D * D]) {
    double vc[D], S = R, yp = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) s = fma(V[i * D + k], Cc[k], s);
        vc[i] = s;
        yp = fma(Cc[i], x[i], yp);
    }
#pragma unroll
    for (int k = 0; k < D; k++) S = fma(Cc[k], vc[k], S);
    const double rS = 1.0 / S, delta = r - yp;
#pragma unroll
    for (int i = 0; i < D; i++) x[i] = fma(vc[i] * rS, delta, x[i]);
    // Vu = Vp - (Vp C')(C Vp)/S ; Vp symmetric so C Vp = (Vp C')'
#pragma unroll
    for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j < D; j++) V[i * D + j] = fma(-vc[i] * rS, vc[j], V[i * D + j]);
    return delta * rS * delta + log(S);
}

// filtering element of step t > 0
template <int D>
__device__ __forceinline__ void filt_element(const double (&F)[D * D], const double (&Q)[D * D], const double (&Cc)[D],
                                             double R, bool obs, double r, const double (&c)[D], FiltElem<D> &e) {
    if (!obs) {
#pragma unroll
        for (int i = 0; i < D * D; i++) {
            e.A[i] = F[i];
            e.C[i] = Q[i];
            e.J[i] = 0.0;
        }
#pragma unroll
        for (int i = 0; i < D; i++) {
            e.b[i] = c[i];
            e.eta[i] = 0.0;
        }
        return;
    }
    // r = y - D v ; innovation of the zero-state response: r - H c
    double qh[D], S = R, hc = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) s = fma(Q[i * D + k], Cc[k], s);
        qh[i] = s;
        hc = fma(Cc[i], c[i], hc);
    }
#pragma unroll
    for (int k = 0; k < D; k++) S = fma(Cc[k], qh[k], S);
    const double rS = 1.0 / S, rr = r - hc;
    double hf[D]; // H F
#pragma unroll
    for (int j = 0; j < D; j++) {
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) s = fma(Cc[k], F[k * D + j], s);
        hf[j] = s;
    }
#pragma unroll
    for (int i = 0; i < D; i++) {
        const double K = qh[i] * rS;
        e.b[i] = fma(K, rr, c[i]);
        e.eta[i] = hf[i] * rS * rr;
#pragma unroll
        for (int j = 0; j < D; j++) {
            e.A[i * D + j] = fma(-K, hf[j], F[i * D + j]); // (I - K H) F
            e.C[i * D + j] = fma(-K, qh[j], Q[i * D + j]); // (I - K H) Q
            e.J[i * D + j] = hf[i] * rS * hf[j];
        }
    }
}

// a <- a combined with the element of ONE step (what filt_agg does L times per chunk), without forming that
// element: the generic combine needs (I + C_a J_e)^-1, a d x d inverse, but a step with a scalar observation has
// J_e = s w w' (w = (H F)', s = 1/S) and Sherman-Morrison gives it in closed form,
//   (I + s C_a w w')^-1 = I - kappa g w',  g = C_a w,  kappa = s / (1 + s w'g),
// which turns combine(a, e) into (K = Q H'/S the step's gain, A_e = F - K w', z = A_a' w, Y = C_a - kappa g g')
//   A   <- F A_a - (K + kappa A_e g) z'
//   b   <- A_e (t - kappa g (w't)) + c + K rr,            t = b_a + s rr g,  rr = r - H c
//   C   <- A_e Y A_e' + Q - K (Q H')'
//   eta <- eta_a + kappa (rr - w'b_a) z
//   J   <- J_a + kappa z z'
// (three d x d products and a dozen matrix-vector ones instead of eight products and an inverse; no element, no
// inverse, a third of the live values -- filt_agg spilled 1 008 B per thread).  An unobserved step has J_e = 0,
// eta_e = 0: A <- F A_a, b <- F b_a + c, C <- F C_a F' + Q, eta and J unchanged.  Same map as filt_combine(a,
// filt_element(..)) in exact arithmetic.
template <int D>
__device__ __forceinline__ void filt_append_step(FiltElem<D> &a, const double (&F)[D * D], const double (&Q)[D * D],
                                                 const double (&Cc)[D], double R, bool obs, double r,
                                                 const double (&c)[D]) {
    double FA[D * D], P[D * D], nC[D * D];
    mm<D>(F, a.A, FA);
    if (!obs) {
        double nb[D];
        mv<D>(F, a.b, nb);
        mm<D>(F, a.C, P);
        mmt<D>(P, F, nC);
#pragma unroll
        for (int i = 0; i < D * D; i++) {
            a.A[i] = FA[i];
            nC[i] += Q[i];
        }
#pragma unroll
        for (int i = 0; i < D * D; i++) a.C[i] = 0.5 * (nC[i] + nC[(i % D) * D + i / D]);
#pragma unroll
        for (int i = 0; i < D; i++) a.b[i] = nb[i] + c[i];
        return;
    }
    double qh[D], w[D], S = R, hc = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        double sq = 0.0, sw = 0.0;
#pragma unroll
        for (int k = 0; k < D; k++) {
            sq = fma(Q[i * D + k], Cc[k], sq);
            sw = fma(Cc[k], F[k * D + i], sw);
        }
        qh[i] = sq; // Q H'
        w[i] = sw;  // (H F)'
        hc = fma(Cc[i], c[i], hc);
    }
#pragma unroll
    for (int k = 0; k < D; k++) S = fma(Cc[k], qh[k], S);
    const double s = 1.0 / S, rr = r - hc;
    double K[D], g[D], z[D];
    mv<D>(a.C, w, g);
    mtv<D>(a.A, w, z);
    double wg = 0.0, wb = 0.0;
#pragma unroll
    for (int i = 0; i < D; i++) {
        K[i] = qh[i] * s;
        wg = fma(w[i], g[i], wg);
        wb = fma(w[i], a.b[i], wb);
    }
    const double kappa = s / fma(s, wg, 1.0);
    // b: A_e x = F x - K (w'x)
    {
        double t[D], Ft[D], wt = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) t[i] = fma(s * rr, g[i], a.b[i]);
#pragma unroll
        for (int i = 0; i < D; i++) wt = fma(w[i], t[i], wt);
#pragma unroll
        for (int i = 0; i < D; i++) t[i] = fma(-kappa * wt, g[i], t[i]);
        wt = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) wt = fma(w[i], t[i], wt);
        mv<D>(F, t, Ft);
#pragma unroll
        for (int i = 0; i < D; i++) a.b[i] = fma(K[i], rr - wt, Ft[i]) + c[i];
    }
    // eta, J (rank one)
    {
        const double ke = kappa * (rr - wb);
#pragma unroll
        for (int i = 0; i < D; i++) {
            a.eta[i] = fma(ke, z[i], a.eta[i]);
#pragma unroll
            for (int j = 0; j < D; j++) a.J[i * D + j] = fma(kappa * z[i], z[j], a.J[i * D + j]);
        }
    }
    // A = F A_a - (K + kappa A_e g) z',  A_e g = F g - K (w'g)
    {
        double Fg[D];
        mv<D>(F, g, Fg);
#pragma unroll
        for (int i = 0; i < D; i++) {
            const double col = fma(kappa, fma(-K[i], wg, Fg[i]), K[i]);
#pragma unroll
            for (int j = 0; j < D; j++) a.A[i * D + j] = fma(-col, z[j], FA[i * D + j]);
        }
    }
    // C = A_e Y A_e' + Q - K qh',  Y = C_a - kappa g g' (symmetric), yv = Y w = g / (1 + s w'g)
    //   A_e Y A_e' = P F' - (F yv) K' - K (F yv)' + (yv'w) K K',  P = F Y
    {
        double Y[D * D], yv[D], Fy[D];
        const double sh = 1.0 / fma(s, wg, 1.0);
#pragma unroll
        for (int i = 0; i < D; i++) {
            yv[i] = g[i] * sh;
#pragma unroll
            for (int j = 0; j < D; j++) Y[i * D + j] = fma(-kappa * g[i], g[j], a.C[i * D + j]);
        }
        mm<D>(F, Y, P);
        mmt<D>(P, F, nC);
        mv<D>(F, yv, Fy);
        const double yw = wg * sh;
#pragma unroll
        for (int i = 0; i < D; i++)
#pragma unroll
            for (int j = 0; j < D; j++) {
                double x = nC[i * D + j] + Q[i * D + j];
                x = fma(-Fy[i], K[j], x);
                x = fma(-K[i], Fy[j], x);
                x = fma(yw * K[i], K[j], x);
                x = fma(-K[i], qh[j], x);
                nC[i * D + j] = x;
            }
#pragma unroll
        for (int i = 0; i < D * D; i++) a.C[i] = 0.5 * (nC[i] + nC[(i % D) * D + i / D]);
    }
}

// ---- kernels --------------------------------------------------------------------------------
// c_t = B u_{t-1}, dv_t = D v_t : one thread per (fit, t)
template <int D> __global__ void scan_prep_kernel(const ScanParams P) {
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)P.n_fits * P.T) return;
    const int f = (int)(gid / P.T), t = (int)(gid % P.T);
    const ThetaD<D> th(P.theta + (size_t)f * P.theta_len, P.p, P.q);
    double c[D];
#pragma unroll
    for (int i = 0; i < D; i++) c[i] = 0.0;
    if (P.u && t > 0) {
        const double *__restrict__ row = P.u + (size_t)(t - 1) * P.p;
        for (int j = 0; j < P.p; j++) {
            const double uj = row[j];
#pragma unroll
            for (int i = 0; i < D; i++) c[i] = fma(th.B[i * P.p + j], uj, c[i]);
        }
    }
    double dv = 0.0;
    if (P.v) {
        const double *__restrict__ row = P.v + (size_t)t * P.q;
        for (int j = 0; j < P.q; j++) dv = fma(th.Dd[j], row[j], dv);
    }
#pragma unroll
    for (int i = 0; i < D; i++) P.c[((size_t)f * P.T + t) * D + i] = c[i];
    P.dv[(size_t)f * P.T + t] = dv;
}

// element of a whole chunk: one thread per (fit, chunk)
template <int D> __global__ void scan_filt_agg_kernel(const ScanParams P) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= P.n_fits * P.n_chunks) return;
    const int f = gid / P.n_chunks, ch = gid % P.n_chunks;
    const ThetaD<D> th(P.theta + (size_t)f * P.theta_len, P.p, P.q);
    double F[D * D], Q[D * D], Cc[D];
    mat_load<D>(th.A, F);
    mat_load<D>(th.Q, Q);
    vec_load<D>(th.C, Cc);
    const int t0 = ch * P.L, t1 = min(P.T, t0 + P.L);
    FiltElem<D> a;
    for (int t = t0; t < t1; t++) {
        const double yt = P.y[t];
        const bool obs = yt == yt;
        const double r = obs ? yt - P.dv[(size_t)f * P.T + t] : 0.0;
        if (t == 0) { // (0, Xu_0, Vu_0, 0, 0): the ordinary update of the prior
            vec_load<D>(th.mu1, a.b);
            mat_load<D>(th.V1, a.C);
            if (obs) meas_update<D>(Cc, th.R, r, a.b, a.C);
#pragma unroll
            for (int i = 0; i < D * D; i++) a.A[i] = a.J[i] = 0.0;
#pragma unroll
            for (int i = 0; i < D; i++) a.eta[i] = 0.0;
        } else {
            double c[D];
            vec_load<D>(P.c + ((size_t)f * P.T + t) * D, c);
            if (t == t0)
                filt_element<D>(F, Q, Cc, th.R, obs, r, c, a); // the chunk's first step: its element
            else
                filt_append_step<D>(a, F, Q, Cc, th.R, obs, r, c); // == filt_combine(a, element of step t)
        }
    }
    a.store(P.fagg + ((size_t)f * P.n_chunks + ch) * FiltElem<D>::LEN);
}

// prefix over the chunk elements: one CTA of NT threads per fit (MODE 0) or per group of chunks (MODES 1, 2), three levels.
//   (1) thread t combines its contiguous range of chunk elements serially;
//   (2) Kogge-Stone over the 32 thread aggregates of a warp through shared memory (5 rounds), warp totals;
//   (3) every thread combines the totals of the warps before its own (< NT / 32 of them) with the inclusive
//       prefix of the lane before it, then walks its range again writing the filtered state entering each chunk.
// Serial depth n/NT + 5 + NT/32 + 1 + n/NT combines for n elements (round 1: one warp per fit, n/32 + 32 + n/32).
// Empty ranges hold the neutral element, so no lane needs a special case.
// A combine is ~600 multiply-adds per thread, so ONE CTA is bound by the FP64 pipe of its one SM (T = 100 000,
// L = 32: 39 combines deep, 0.20 ms).  Long series therefore cut the chunk elements into SCAN_GROUPS groups:
//   MODE 1  one CTA per (fit, group): levels (1), (2) and the group's total -> fgagg          (many SMs)
//   top     one warp per fit: exclusive prefix over the group totals -> fgpre                  (5 rounds)
//   MODE 2  one CTA per (fit, group) again: levels (1)-(3) starting from the group's prefix    (many SMs)
constexpr int SCAN_NT = 256, SCAN_NT_GROUP = 128, SCAN_GROUPS = 32, SCAN_GROUPS_FROM = 512; // chunks from which groups pay
template <int D, int NT, int MODE>
__global__ void __launch_bounds__(NT) scan_filt_scan_kernel(const ScanParams P) {
    extern __shared__ double sh[]; // [NT][LEN] thread aggregates -> inclusive prefixes in the warp; [NT/32][LEN] warp totals
    constexpr int LEN = FiltElem<D>::LEN;
    const int f = blockIdx.x, g = MODE == 0 ? 0 : blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = MODE == 0 ? P.n_chunks : (P.n_chunks + P.n_groups - 1) / P.n_groups; // chunks per group
    const int c_lo = min(P.n_chunks, g * S), n = min(P.n_chunks, c_lo + S) - c_lo;
    const int per = (n + NT - 1) / NT;
    const int k0 = c_lo + min(n, tid * per), k1 = c_lo + min(n, tid * per + per);
    const double *__restrict__ agg = P.fagg + (size_t)f * P.n_chunks * LEN;
    FiltElem<D> a, e;
    filt_identity<D>(a);
    for (int k = k0; k < k1; k++) {
        e.load(agg + (size_t)k * LEN);
        filt_combine<D>(a, e);
    }
    double *const mine = sh + (size_t)tid * LEN, *const wt = sh + (size_t)NT * LEN;
    a.store(mine);
    __syncwarp();
    for (int o = 1; o < 32; o <<= 1) {
        const bool act = lane >= o;
        if (act) { // new_i = old_{i-o} o old_i
            e.load(sh + (size_t)(tid - o) * LEN);
            filt_combine<D>(e, a);
            a = e;
        }
        __syncwarp();
        if (act) a.store(mine);
        __syncwarp();
    }
    if (lane == 31) a.store(wt + (size_t)warp * LEN);
    __syncthreads();
    if (MODE == 1) { // the group's total
        if (tid == 0) {
            e.load(wt);
            for (int w = 1; w < NT / 32; ++w) {
                a.load(wt + (size_t)w * LEN);
                filt_combine<D>(e, a);
            }
            e.store(P.fgagg + ((size_t)f * P.n_groups + g) * LEN);
        }
        return;
    }
    FiltElem<D> run; // exclusive prefix before my range
    if (MODE == 2)
        run.load(P.fgpre + ((size_t)f * P.n_groups + g) * LEN);
    else
        filt_identity<D>(run);
    for (int w = 0; w < warp; ++w) {
        e.load(wt + (size_t)w * LEN);
        filt_combine<D>(run, e);
    }
    if (lane > 0) {
        e.load(sh + (size_t)(tid - 1) * LEN);
        filt_combine<D>(run, e);
    }
    // walk the range again: the filtered state entering chunk k is (b, C) of the prefix before it
    double *__restrict__ pre = P.pre + (size_t)f * P.n_chunks * (D + D * D);
    for (int k = k0; k < k1; k++) {
        if (k > 0) {
#pragma unroll
            for (int i = 0; i < D; i++) pre[(size_t)k * (D + D * D) + i] = run.b[i];
#pragma unroll
            for (int i = 0; i < D * D; i++) pre[(size_t)k * (D + D * D) + D + i] = run.C[i];
        }
        e.load(agg + (size_t)k * LEN);
        filt_combine<D>(run, e);
    }
}
// exclusive prefix over the SCAN_GROUPS group totals of a fit: one warp, Kogge-Stone through shared memory
template <int D> __global__ void __launch_bounds__(SCAN_GROUPS) scan_filt_top_kernel(const ScanParams P) {
    static_assert(SCAN_GROUPS == 32, "one warp");
    extern __shared__ double sh[]; // [32][LEN]
    constexpr int LEN = FiltElem<D>::LEN;
    const int f = blockIdx.x, lane = threadIdx.x;
    FiltElem<D> a, e;
    if (lane < P.n_groups)
        a.load(P.fgagg + ((size_t)f * P.n_groups + lane) * LEN);
    else
        filt_identity<D>(a);
    double *const mine = sh + (size_t)lane * LEN;
    a.store(mine);
    __syncwarp();
    for (int o = 1; o < 32; o <<= 1) {
        const bool act = lane >= o;
        if (act) {
            e.load(sh + (size_t)(lane - o) * LEN);
            filt_combine<D>(e, a);
            a = e;
        }
        __syncwarp();
        if (act) a.store(mine);
        __syncwarp();
    }
    if (lane < P.n_groups) {
        if (lane > 0)
            e.load(sh + (size_t)(lane - 1) * LEN);
        else
            filt_identity<D>(e);
        e.store(P.fgpre + ((size_t)f * P.n_groups + lane) * LEN);
    }
}

// ordinary Kalman filter inside a chunk (EM.cpp:70-90 generalised)
template <int D> __global__ void scan_filt_down_kernel(const ScanParams P) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= P.n_fits * P.n_chunks) return;
    const int f = gid / P.n_chunks, ch = gid % P.n_chunks;
    const ThetaD<D> th(P.theta + (size_t)f * P.theta_len, P.p, P.q);
    double F[D * D], Q[D * D], Cc[D];
    mat_load<D>(th.A, F);
    mat_load<D>(th.Q, Q);
    vec_load<D>(th.C, Cc);
    const int t0 = ch * P.L, t1 = min(P.T, t0 + P.L);
    double x[D], V[D * D];
    if (ch > 0) {
        const double *__restrict__ pre = P.pre + ((size_t)f * P.n_chunks + ch) * (D + D * D);
        vec_load<D>(pre, x);
        mat_load<D>(pre + D, V);
    }
    double acc = 0.0, nobs = 0.0;
    for (int t = t0; t < t1; t++) {
        if (t == 0) {
            vec_load<D>(th.mu1, x);
            mat_load<D>(th.V1, V);
        } else { // predict
            double c[D], xn[D], T1[D * D];
            vec_load<D>(P.c + ((size_t)f * P.T + t) * D, c);
            mv<D>(F, x, xn);
#pragma unroll
            for (int i = 0; i < D; i++) x[i] = xn[i] + c[i];
            mm<D>(F, V, T1);
            mmt<D>(T1, F, V);
#pragma unroll
            for (int i = 0; i < D * D; i++) V[i] += Q[i];
        }
        const double yt = P.y[t];
        if (yt == yt) {
            acc += meas_update<D>(Cc, th.R, yt - P.dv[(size_t)f * P.T + t], x, V);
            nobs += 1.0;
        }
#pragma unroll
        for (int i = 0; i < D; i++) P.Xu[((size_t)f * P.T + t) * D + i] = x[i];
#pragma unroll
        for (int i = 0; i < D * D; i++) P.Vu[((size_t)f * P.T + t) * D * D + i] = V[i];
    }
    P.likp[((size_t)f * P.n_chunks + ch) * 2] = acc;
    P.likp[((size_t)f * P.n_chunks + ch) * 2 + 1] = nobs;
}

// ---- smoothing element (Appendix D.3) -----------------------------------------------------------
template <int D> struct SmthElem {
    double E[D * D], g[D], L[D * D];
    static constexpr int LEN = 2 * D * D + D;
    __device__ __forceinline__ void store(double *__restrict__ o) const {
#pragma unroll
        for (int i = 0; i < D * D; i++) {
            o[i] = E[i];
            o[D * D + D + i] = L[i];
        }
#pragma unroll
        for (int i = 0; i < D; i++) o[D * D + i] = g[i];
    }
    __device__ __forceinline__ void load(const double *__restrict__ o) {
#pragma unroll
        for (int i = 0; i < D * D; i++) {
            E[i] = o[i];
            L[i] = o[D * D + D + i];
        }
#pragma unroll
        for (int i = 0; i < D; i++) g[i] = o[D * D + i];
    }
};
template <int D> __device__ __forceinline__ void smth_identity(SmthElem<D> &a) { // neutral element: (I, 0, 0)
#pragma unroll
    for (int i = 0; i < D * D; i++) {
        a.E[i] = (i % (D + 1) == 0) ? 1.0 : 0.0;
        a.L[i] = 0.0;
    }
#pragma unroll
    for (int i = 0; i < D; i++) a.g[i] = 0.0;
}
// a (earlier) <- a combined with e (later): E = E_i E_j, g = E_i g_j + g_i, L = E_i L_j E_i' + L_i
template <int D> __device__ __forceinline__ void smth_combine(SmthElem<D> &a, const SmthElem<D> &e) {
    double nE[D * D], t[D], T1[D * D], nL[D * D];
    mm<D>(a.E, e.E, nE);
    mv<D>(a.E, e.g, t);
    mm<D>(a.E, e.L, T1);
    mmt<D>(T1, a.E, nL);
#pragma unroll
    for (int i = 0; i < D; i++) a.g[i] += t[i];
#pragma unroll
    for (int i = 0; i < D * D; i++) {
        a.L[i] += 0.5 * (nL[i] + nL[(i % D) * D + i / D]);
        a.E[i] = nE[i];
    }
}
// RTS gain and predicted moments of step t+1 from the filtered moments of step t
template <int D>
__device__ __forceinline__ void rts_gain(const double (&F)[D * D], const double (&Q)[D * D], const double (&xu)[D],
                                         const double (&Vu)[D * D], const double (&c1)[D], double (&E)[D * D],
                                         double (&xp1)[D], double (&Vp1)[D * D]) {
    double T1[D * D], iV[D * D], VFt[D * D];
    mm<D>(F, Vu, T1);
    mmt<D>(T1, F, Vp1);
#pragma unroll
    for (int i = 0; i < D * D; i++) Vp1[i] += Q[i];
    mat_inv<D>(Vp1, iV);
    mmt<D>(Vu, F, VFt); // Vu F'
    mm<D>(VFt, iV, E);  // J_t = Vu_t A' Vp_{t+1}^-1   (EM.cpp:100)
    mv<D>(F, xu, xp1);
#pragma unroll
    for (int i = 0; i < D; i++) xp1[i] += c1[i];
}

template <int D> __global__ void scan_smth_agg_kernel(const ScanParams P) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= P.n_fits * P.n_chunks) return;
    const int f = gid / P.n_chunks, ch = gid % P.n_chunks;
    const ThetaD<D> th(P.theta + (size_t)f * P.theta_len, P.p, P.q);
    double F[D * D], Q[D * D];
    mat_load<D>(th.A, F);
    mat_load<D>(th.Q, Q);
    const int t0 = ch * P.L, t1 = min(P.T, t0 + P.L);
    SmthElem<D> a;
    for (int t = t0; t < t1; t++) {
        double xu[D], Vu[D * D];
        vec_load<D>(P.Xu + ((size_t)f * P.T + t) * D, xu);
        mat_load<D>(P.Vu + ((size_t)f * P.T + t) * D * D, Vu);
        SmthElem<D> e;
        if (t == P.T - 1) {
#pragma unroll
            for (int i = 0; i < D * D; i++) {
                e.E[i] = 0.0;
                e.L[i] = Vu[i];
            }
#pragma unroll
            for (int i = 0; i < D; i++) e.g[i] = xu[i];
        } else {
            double c1[D], xp1[D], Vp1[D * D], tv[D], T1[D * D];
            vec_load<D>(P.c + ((size_t)f * P.T + t + 1) * D, c1);
            rts_gain<D>(F, Q, xu, Vu, c1, e.E, xp1, Vp1);
            mv<D>(e.E, xp1, tv);
#pragma unroll
            for (int i = 0; i < D; i++) e.g[i] = xu[i] - tv[i]; // g = Xu - E (F Xu + c)
            // L = Vu - E Vp1 E'  (= Vu - E F Vu)
            mm<D>(e.E, Vp1, T1);
            double T2[D * D];
            mmt<D>(T1, e.E, T2);
#pragma unroll
            for (int i = 0; i < D * D; i++) e.L[i] = Vu[i] - 0.5 * (T2[i] + T2[(i % D) * D + i / D]);
        }
        if (t == t0)
            a = e;
        else
            smth_combine<D>(a, e);
    }
    a.store(P.sagg + ((size_t)f * P.n_chunks + ch) * SmthElem<D>::LEN);
}

// suffix over the chunk elements: (g, L) of the suffix right of chunk k = smoothed state of its first step.
// Same levels and modes as scan_filt_scan_kernel, mirrored.
template <int D, int NT, int MODE>
__global__ void __launch_bounds__(NT) scan_smth_scan_kernel(const ScanParams P) {
    extern __shared__ double sh[];
    constexpr int LEN = SmthElem<D>::LEN;
    const int f = blockIdx.x, g = MODE == 0 ? 0 : blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int S = MODE == 0 ? P.n_chunks : (P.n_chunks + P.n_groups - 1) / P.n_groups;
    const int c_lo = min(P.n_chunks, g * S), n = min(P.n_chunks, c_lo + S) - c_lo;
    const int per = (n + NT - 1) / NT;
    const int k0 = c_lo + min(n, tid * per), k1 = c_lo + min(n, tid * per + per);
    const double *__restrict__ agg = P.sagg + (size_t)f * P.n_chunks * LEN;
    SmthElem<D> a, e;
    smth_identity<D>(a);
    for (int k = k1 - 1; k >= k0; k--) { // a = e_k0 o ... o e_{k1-1}, built from the right
        e.load(agg + (size_t)k * LEN);
        smth_combine<D>(e, a);
        a = e;
    }
    double *const mine = sh + (size_t)tid * LEN, *const wt = sh + (size_t)NT * LEN;
    a.store(mine);
    __syncwarp();
    for (int o = 1; o < 32; o <<= 1) {
        const bool act = lane + o < 32;
        if (act) { // new_i = old_i o old_{i+o}
            e.load(sh + (size_t)(tid + o) * LEN);
            smth_combine<D>(a, e);
        }
        __syncwarp();
        if (act) a.store(mine);
        __syncwarp();
    }
    if (lane == 0) a.store(wt + (size_t)warp * LEN);
    __syncthreads();
    if (MODE == 1) { // the group's total: e_first o ... o e_last
        if (tid == 0) {
            a.load(wt + (size_t)(NT / 32 - 1) * LEN);
            for (int w = NT / 32 - 2; w >= 0; --w) {
                e.load(wt + (size_t)w * LEN);
                smth_combine<D>(e, a);
                a = e;
            }
            a.store(P.sgagg + ((size_t)f * P.n_groups + g) * LEN);
        }
        return;
    }
    SmthElem<D> run; // exclusive suffix right of my range
    if (MODE == 2)
        run.load(P.sgsuf + ((size_t)f * P.n_groups + g) * LEN);
    else
        smth_identity<D>(run);
    for (int w = NT / 32 - 1; w > warp; --w) {
        e.load(wt + (size_t)w * LEN);
        smth_combine<D>(e, run);
        run = e;
    }
    if (lane < 31) {
        e.load(sh + (size_t)(tid + 1) * LEN);
        smth_combine<D>(e, run);
        run = e;
    }
    double *__restrict__ suf = P.suf + (size_t)f * P.n_chunks * (D + D * D);
    for (int k = k1 - 1; k >= k0; k--) {
        if (k < P.n_chunks - 1) {
#pragma unroll
            for (int i = 0; i < D; i++) suf[(size_t)k * (D + D * D) + i] = run.g[i];
#pragma unroll
            for (int i = 0; i < D * D; i++) suf[(size_t)k * (D + D * D) + D + i] = run.L[i];
        }
        e.load(agg + (size_t)k * LEN);
        smth_combine<D>(e, run);
        run = e;
    }
}
// exclusive suffix over the group totals of a fit: one warp
template <int D> __global__ void __launch_bounds__(SCAN_GROUPS) scan_smth_top_kernel(const ScanParams P) {
    extern __shared__ double sh[];
    constexpr int LEN = SmthElem<D>::LEN;
    const int f = blockIdx.x, lane = threadIdx.x;
    SmthElem<D> a, e;
    if (lane < P.n_groups)
        a.load(P.sgagg + ((size_t)f * P.n_groups + lane) * LEN);
    else
        smth_identity<D>(a);
    double *const mine = sh + (size_t)lane * LEN;
    a.store(mine);
    __syncwarp();
    for (int o = 1; o < 32; o <<= 1) {
        const bool act = lane + o < 32;
        if (act) {
            e.load(sh + (size_t)(lane + o) * LEN);
            smth_combine<D>(a, e);
        }
        __syncwarp();
        if (act) a.store(mine);
        __syncwarp();
    }
    if (lane < P.n_groups) {
        if (lane < 31)
            e.load(sh + (size_t)(lane + 1) * LEN);
        else
            smth_identity<D>(e);
        e.store(P.sgsuf + ((size_t)f * P.n_groups + lane) * LEN);
    }
}

// ordinary RTS recursion inside a chunk (EM.cpp:99-110 generalised)
template <int D> __global__ void scan_smth_down_kernel(const ScanParams P) {
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= P.n_fits * P.n_chunks) return;
    const int f = gid / P.n_chunks, ch = gid % P.n_chunks;
    const ThetaD<D> th(P.theta + (size_t)f * P.theta_len, P.p, P.q);
    double F[D * D], Q[D * D], Cc[D];
    mat_load<D>(th.A, F);
    mat_load<D>(th.Q, Q);
    vec_load<D>(th.C, Cc);
    const int t0 = ch * P.L, t1 = min(P.T, t0 + P.L);
    double xs1[D], Vs1[D * D];
    if (ch < P.n_chunks - 1) {
        const double *__restrict__ suf = P.suf + ((size_t)f * P.n_chunks + ch) * (D + D * D);
        vec_load<D>(suf, xs1);
        mat_load<D>(suf + D, Vs1);
    }
    for (int t = t1 - 1; t >= t0; t--) {
        double xu[D], Vu[D * D], xs[D], Vs[D * D];
        vec_load<D>(P.Xu + ((size_t)f * P.T + t) * D, xu);
        mat_load<D>(P.Vu + ((size_t)f * P.T + t) * D * D, Vu);
        if (t == P.T - 1) {
#pragma unroll
            for (int i = 0; i < D; i++) xs[i] = xu[i];
#pragma unroll
            for (int i = 0; i < D * D; i++) Vs[i] = Vu[i];
        } else {
            double c1[D], xp1[D], Vp1[D * D], E[D * D], dx[D], T1[D * D], T2[D * D];
            vec_load<D>(P.c + ((size_t)f * P.T + t + 1) * D, c1);
            rts_gain<D>(F, Q, xu, Vu, c1, E, xp1, Vp1);
#pragma unroll
            for (int i = 0; i < D; i++) dx[i] = xs1[i] - xp1[i];
            mv<D>(E, dx, xs);
#pragma unroll
            for (int i = 0; i < D; i++) xs[i] += xu[i];
#pragma unroll
            for (int i = 0; i < D * D; i++) T1[i] = Vs1[i] - Vp1[i];
            mm<D>(E, T1, T2);
            mmt<D>(T2, E, T1);
#pragma unroll
            for (int i = 0; i < D * D; i++) Vs[i] = Vu[i] + 0.5 * (T1[i] + T1[(i % D) * D + i / D]);
        }
        double ys = P.dv[(size_t)f * P.T + t];
#pragma unroll
        for (int i = 0; i < D; i++) {
            P.X[((size_t)f * P.T + t) * D + i] = xs[i];
            ys = fma(Cc[i], xs[i], ys);
            xs1[i] = xs[i];
        }
#pragma unroll
        for (int i = 0; i < D * D; i++) {
            P.V[((size_t)f * P.T + t) * D * D + i] = Vs[i];
            Vs1[i] = Vs[i];
        }
        P.Y[(size_t)f * P.T + t] = ys;
    }
}

// likelihood: the chunk terms of a fit are added by one warp, lane-strided then a shuffle tree --
// a fixed order, so the result does not depend on the launch (EM.cpp:115-124)
// sum of the chunks' likelihood terms: one CTA of SCAN_LIK_NT threads per fit (strided partial sums with the loads
// of four chunks in flight, then a fixed-order tree: the result does not depend on the launch)
constexpr int SCAN_LIK_NT = 256;
template <int D> __global__ void __launch_bounds__(SCAN_LIK_NT) scan_lik_kernel(const ScanParams P) {
    __shared__ double sa[SCAN_LIK_NT / 32], sn[SCAN_LIK_NT / 32];
    const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double2 *__restrict__ lp = reinterpret_cast<const double2 *>(P.likp) + (size_t)f * P.n_chunks;
    double acc = 0.0, n = 0.0;
    int k = tid;
    for (; k + 3 * SCAN_LIK_NT < P.n_chunks; k += 4 * SCAN_LIK_NT) {
        const double2 a0 = lp[k], a1 = lp[k + SCAN_LIK_NT], a2 = lp[k + 2 * SCAN_LIK_NT], a3 = lp[k + 3 * SCAN_LIK_NT];
        acc += (a0.x + a1.x) + (a2.x + a3.x);
        n += (a0.y + a1.y) + (a2.y + a3.y);
    }
    for (; k < P.n_chunks; k += SCAN_LIK_NT) {
        const double2 a0 = lp[k];
        acc += a0.x;
        n += a0.y;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(FULL, acc, o);
        n += __shfl_xor_sync(FULL, n, o);
    }
    if (lane == 0) {
        sa[warp] = acc;
        sn[warp] = n;
    }
    __syncthreads();
    if (tid == 0) {
        acc = n = 0.0;
#pragma unroll
        for (int w = 0; w < SCAN_LIK_NT / 32; w++) {
            acc += sa[w];
            n += sn[w];
        }
        double lik = -0.5 * n * LOG_2PI - 0.5 * acc;
        if (P.stdlik) lik /= n;
        P.lik[f] = lik;
    }
}

cudaError_t scan_smoother_launch(int D, const ScanParams &P, cudaStream_t st); // scan_inst.cu
int scan_groups_for(int n_chunks); // CTAs per fit in the scan stages (ScanParams.n_groups): 1 or SCAN_GROUPS

} // namespace ldsr
