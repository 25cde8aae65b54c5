// em_split_kernel.cuh -- the LDS_EM loop (src/EM.cpp:245-280) with the TIME AXIS SPLIT ACROSS THE
// WARPS OF A CTA.  lane = fit (32 fits of one series per CTA), warp = pieces of the time axis.
//
// Why: with one lane per fit and one warp walking all T steps (em_kernel.cuh) a batch of N fits is
// only N/32 serial instruction streams -- 313 for the 10 000-fit cvLDS job, on a machine with 592
// SM sub-partitions -- and each stream is latency-bound.  Here every fit is advanced by NW warps at
// once, so the same batch is NW x more streams, each 1/NW as long.
//
// How the recursions are cut (all of it exact algebra on EM.cpp:70-90, 99-104, no approximation):
//  * The time axis is tiled into UNITS: UW steps no fit of the CTA observes (type U), or MSEG
//    steps (type M).  The unit list is cut into 2*NW PIECES; warp w owns piece w and piece NW+w
//    (the cut between the two halves is put where the series changes from mostly-unobserved to
//    mostly-observed, so every warp gets the same mix of cheap and expensive units and the
//    forward and the backward phase are both balanced).
//  * P1  each warp composes the VARIANCE map of its pieces.  In 1-D the Riccati step is a Moebius
//        map of Vp, i.e. a 2x2 matrix acting on homogeneous coordinates (n,d), Vp = n/d:
//          observed   [[A^2 R + Q C^2, Q R],[C^2, R]]      unobserved [[A^2, Q],[0, 1]]
//        (a U unit is the closed form [[A^2UW, Q sum A^2k],[0,1]]).          -> barrier 1
//        Every warp then applies the maps of the pieces to the left of its own to (V1,1).
//  * P2  forward over each piece with the true variances.  The MEAN is carried as an affine
//        function of the (still unknown) incoming mean x_in:  Xp_t = P_t x_in + q_t, so the
//        innovations are affine and sum delta^2/Sigma is a quadratic (l0,l1,l2) in x_in.  In the
//        same sweep the piece's BACKWARD map is composed: the RTS step is affine,
//        Xs_t = J_t Xs_{t+1} + g_t, Vs_t = J_t^2 Vs_{t+1} + L_t, and over a U unit it telescopes
//        to J = A^UW Vp_first / Vp_last.  Checkpoints (Vp, q, P) per unit.      -> barrier 2
//        Every warp chains the (P,q) of all pieces -> x_in of every piece, the likelihood and the
//        stop rule (EM.cpp:272; identical arithmetic in every warp, so no broadcast is needed),
//        and the backward maps of the pieces to the right -> smoothed state entering its pieces.
//        The backward chain starts from the PRIOR of the virtual step T, which makes
//        Xs_{T-1} = Xu_{T-1} (EM.cpp:94-95) fall out of the ordinary recursion.   -> barrier 2'
//  * P4  backward over each piece, unit by unit: M units are recomputed from their checkpoint
//        and smoothed as in em_kernel.cuh; U units are STREAMED FORWARD because inside a run of
//        unobserved steps  Xs_t = Xp_t + Vp_t A^(r-t) c,  Vs_t = Vp_t + Vp_t^2 A^(2(r-t)) h  with
//        c = (Xs_r - Xp_r)/Vp_r, h = (Vs_r - Vp_r)/Vp_r^2 taken once at the right end r of the run
//        -- no reciprocal and no dependency chain per step.  M-step sums in registers.
//                                                                                -> barrier 3
//  * every warp adds the NW partial sums in the same order and does the M-step (EM.cpp:139-229).
// Shared memory: series blob (TMA bulk copy), 3 doubles per unit and lane of checkpoints, and the
// exchange buffers; nothing O(T) per fit is ever written to global memory.
#pragma once
#include "em_kernel.cuh"

namespace ldsr {

constexpr int SPLIT_NCH = 10; // per-piece values exchanged after P2

// number of M-step partial sums a warp publishes
template <int PQ> __host__ __device__ constexpr int split_nstat() { return 11 + 3 * PQ; }

// dynamic shared memory of em_split_kernel, in bytes, after the series blob
// The Tx1u identity of smooth_word needs the unit's Horner vector W from P2, kept per U unit in
// shared memory.  Narrow inputs only: at PQ = 10 it was measured slower (a PQ x PQ product per unit,
// more spills; W would not fit shared memory and went through an L2-resident scratch): 1.38 -> 1.44 s
// on the 480 000-fit job.
__host__ __device__ constexpr bool split_w_in_smem(int pq) { return pq <= 4; }
__host__ __device__ inline size_t split_smem_bytes(int pq, int nw, int max_units, int max_uunits) {
    size_t b = 0;
    b += (size_t)max_units * 3 * 32 * 8; // checkpoints
    b += (size_t)2 * nw * 4 * 32 * 8;    // variance maps of the pieces
    const bool pair = nw % 2 == 0 && nw * (11 + 3 * pq) > 2 * nw * SPLIT_NCH; // split_pairwise<PQ,NW>()
    const size_t ch = (size_t)2 * nw * SPLIT_NCH * 32 * 8, st = (size_t)(pair ? nw / 2 : nw) * (11 + 3 * pq) * 32 * 8;
    b += ch > st ? ch : st;                           // piece summaries, later the partial sums
    b += ((size_t)max_units * 4 + 15) & ~size_t(15); // unit table
    b += 128;                                        // piece bounds
    b += (size_t)8 * 32 * 8;                         // closed-form variance-sum coefficients
    if (split_w_in_smem(pq)) b += (size_t)max_uunits * pq * 32 * 8; // W of every U unit
    if (theta_bd_in_smem(pq)) b += (size_t)2 * pq * 32 * 8;          // B and D of the CTA's fits
    return b;
}

// Unit types.  U: UW steps nobody observes.  M: MSEG steps.  M1: a single step -- the tail of the
// series (everything from the last full M unit to step T-1) is cut into single steps so that the
// hot M code never sees a partial unit and step T-1 (no transition after it) is a unit of its own.
// US: MSEG steps inside a mixed window that no fit of the CTA observes (NP: 352..359 before the first
// observation, 408..411 after the last): handled like a short U unit, about half the cycles of an M unit.
constexpr int UNIT_M = 1 << 30, UNIT_M1 = 1 << 29, UNIT_US = 1 << 28, UNIT_T0 = UNIT_US - 1;

// M / M1 units of the non-U window [t0, t0+uw): calls f(t, is_single)
template <class F> __host__ __device__ inline void split_window_units(int t0, int T, int mseg, int uw, F f) {
    const int end = T < t0 + uw ? T : t0 + uw;
    int t = t0;
    while (t < end) {
        if (t + mseg <= T - 1 && t + mseg <= end) {
            f(t, false);
            t += mseg;
        } else {
            f(t, true);
            t += 1;
        }
    }
}

// upper bound on the number of units of a series from its finite(y) mask (hold-outs only remove
// observations, and a window with no observation is one unit instead of several)
inline int split_units_upper_bound(const double *y, int T, int mseg, int uw) {
    int n = 0;
    for (int t0 = 0; t0 < T; t0 += uw) {
        bool any = false;
        for (int t = t0; t < T && t < t0 + uw; t++) any = any || (y[t] == y[t]);
        const bool inside = t0 + uw <= T - 1;
        if (!any && inside)
            n += 1;
        else
            split_window_units(t0, T, mseg, uw, [&](int, bool) { n++; });
    }
    return n;
}

// N consecutive doubles from shared memory; 16-byte vector loads when N is even (the callers
// guarantee 16-byte alignment: unit starts are multiples of 4 steps, blob offsets are even)
template <int N> __device__ __forceinline__ void load_vec(const double *__restrict__ p, double (&dst)[N]) {
    if constexpr (N % 2 == 0) {
        const double2 *__restrict__ p2 = reinterpret_cast<const double2 *>(p);
#pragma unroll
        for (int i = 0; i < N / 2; i++) {
            const double2 t = p2[i];
            dst[2 * i] = t.x;
            dst[2 * i + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; i++) dst[i] = p[i];
    }
}

__device__ __forceinline__ void rescale4(double &a, double &b, double &c, double &d) {
    const int e = ((__double2hiint(a + b + c + d) >> 20) & 0x7ff) - 1023;
    const double sc = __hiloint2double((1023 - e) << 20, 0);
    a *= sc;
    b *= sc;
    c *= sc;
    d *= sc;
}

// per-iteration constants of a fit.  Kept to the few values every unit needs: ptxas would rather
// recompute a longer table of powers inside the unit loops than hold it in registers (seen in the
// SASS: a chain of 4 dependent DMULs at the top of every 8-step block).
template <int PQ, int UW> struct SplitConst {
    static_assert(UW == 16 || UW == 32, "U unit is 16 or 32 steps");
    double A, A2, Q;
    double A8, A16;    // A^8, A^16
    double aVW, bVW;   // Vp' = aVW Vp + bVW over an unobserved unit
    MixedConst<PQ> mc;
    __device__ __forceinline__ double AW() const { return UW == 32 ? A16 * A16 : A16; } // A^UW
    __device__ __forceinline__ void set(const Theta<PQ> &th) {
        A = th.A;
        A2 = A * A;
        Q = th.Q;
        const double A4 = A2 * A2;
        A8 = A4 * A4;
        A16 = A8 * A8;
        const double aw = AW();
        aVW = aw * aw;
        // sum_{k<8} A2^k = (1 + A2)(1 + A2^2)(1 + A2^4)
        const double sv = ((1.0 + A2) * (1.0 + A4)) * (1.0 + A8);
        // sum_{k<16} A2^k = sv (1 + A2^8), sum_{k<32} = sv (1 + A2^8)(1 + A2^16);  A2^8 = A^16
        double s = sv * (1.0 + A16);
        if (UW == 32) s *= 1.0 + A16 * A16;
        bVW = Q * s;
        mc.set(th, A2);
    }
};

// ---- P1: variance map of one M unit, M <- S_{N-1} ... S_0 M -----------------------------------
template <int PQ, int UW, int N>
__device__ __forceinline__ void compose_var_unit(const Theta<PQ> &th, const SplitConst<PQ, UW> &k, unsigned bits,
                                                 double &m11, double &m12, double &m21, double &m22) {
#pragma unroll
    for (int j = 0; j < N; j++) {
        const bool obs = (bits >> j) & 1u;
        const double s11 = obs ? k.mc.a11 : k.A2, s12 = obs ? k.mc.a12 : k.Q;
        const double s21 = obs ? k.mc.C2 : 0.0, s22 = obs ? th.R : 1.0;
        const double n11 = fma(s11, m11, s12 * m21), n12 = fma(s11, m12, s12 * m22);
        const double n21 = fma(s21, m11, s22 * m21), n22 = fma(s21, m12, s22 * m22);
        m11 = n11;
        m12 = n12;
        m21 = n21;
        m22 = n22;
    }
    rescale4(m11, m12, m21, m22);
}

// ---- gains of one M unit -----------------------------------------------------------------------
// The variance recursion in homogeneous coordinates (see em_kernel.cuh, mixed_forward), written
// stage by stage: the (n,d) chain first (two multiply-adds per step), then all reciprocals -- they
// are independent of each other -- then everything that hangs off them.
template <int N> struct UnitGains {
    double K[N], rS[N], alpha[N], J[N], L[N];
    double Vnext; // prior variance of the step after the unit
    double dend;  // final d: prod_obs Sigma
};
template <int PQ, int UW, int N>
__device__ __forceinline__ void unit_gains(const Theta<PQ> &th, const SplitConst<PQ, UW> &k, unsigned bits, double Vq,
                                           UnitGains<N> &G) {
    double nj[N], dj[N], nn[N], dd[N];
    double n = Vq, d = 1.0;
#pragma unroll
    for (int j = 0; j < N; j++) {
        const bool obs = (bits >> j) & 1u;
        const double m11 = obs ? k.mc.a11 : k.A2, m12 = obs ? k.mc.a12 : k.Q;
        const double m21 = obs ? k.mc.C2 : 0.0, m22 = obs ? th.R : 1.0;
        nj[j] = n;
        dj[j] = d;
        nn[j] = fma(m11, n, m12 * d);
        dd[j] = fma(m21, n, m22 * d);
        n = nn[j];
        d = dd[j];
    }
    double rho[N], rn[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        rho[j] = fast_rcp(dd[j]);
        rn[j] = fast_rcp(nn[j]);
    }
#pragma unroll
    for (int j = 0; j < N; j++) {
        const bool obs = (bits >> j) & 1u;
        const double K = obs ? th.C * nj[j] * rho[j] : 0.0;
        const double nu = obs ? th.R * nj[j] : nj[j];
        const double vu = nu * rho[j];
        const double J = k.A * nu * rn[j];
        G.K[j] = K;
        G.rS[j] = obs ? dj[j] * rho[j] : 0.0; // 1/Sigma
        G.alpha[j] = fma(-k.mc.AC, K, k.A);
        G.J[j] = J;
        G.L[j] = vu * fma(-k.A, J, 1.0); // Vu - J^2 Vp' with J Vp' = A Vu
    }
    G.Vnext = nn[N - 1] * rho[N - 1];
    G.dend = d;
}

// inputs of one M unit: Bu_j = B u_j, ymd_j = y_j - D v_j (y taken as 0 where unobserved)
template <int PQ, int N>
__device__ __forceinline__ void unit_inputs(const Theta<PQ> &th, unsigned bits, const double *__restrict__ yseg,
                                            const double *__restrict__ useg, const double *__restrict__ vseg,
                                            double (&Bu)[N], double (&ymd)[N], double (&yo)[N],
                                            double (&ur)[N * PQ], double (&vr)[N * PQ]) {
    double yr[N];
    load_vec<N * PQ>(useg, ur);
    load_vec<N * PQ>(vseg, vr);
    load_vec<N>(yseg, yr);
#pragma unroll
    for (int j = 0; j < N; j++) {
        double b = 0.0, dv = 0.0;
#pragma unroll
        for (int i = 0; i < PQ; i++) {
            b = fma(th.b(i), ur[j * PQ + i], b);
            dv = fma(th.d(i), vr[j * PQ + i], dv);
        }
        const bool obs = (bits >> j) & 1u;
        yo[j] = obs ? yr[j] : 0.0; // y is NaN where missing
        Bu[j] = b;
        ymd[j] = yo[j] - dv;
    }
}

// piece state carried through P2
struct PieceFwd {
    double Vq, P, q;            // prior variance; prior mean = P x_in + q
    double l0, l1, l2;          // sum_obs delta^2/Sigma = l0 - 2 C x l1 + C^2 x^2 l2
    double dprod;               // product of the units' final d (sum_obs log Sigma = log dprod + shift ln 2)
    int shift;
    double PJ, PJ2, G0, GG, Lc; // backward map: Xs_first = PJ Xs_in + G0 + GG x_in, Vs_first = PJ2 Vs_in + Lc
};

// ---- P2 over one M unit -------------------------------------------------------------------------
// Same recursion as mixed_forward (em_kernel.cuh) with the mean in (P,q) form and the backward map
// accumulated on the fly.
template <int PQ, int UW, int N>
__device__ __forceinline__ void forward_unit_basis(const Theta<PQ> &th, const SplitConst<PQ, UW> &k, unsigned bits,
                                                   const double *__restrict__ yseg, const double *__restrict__ useg,
                                                   const double *__restrict__ vseg, PieceFwd &c) {
    double Bu[N], ymd[N], yo[N];
    {
        double ur[N * PQ], vr[N * PQ];
        unit_inputs<PQ, N>(th, bits, yseg, useg, vseg, Bu, ymd, yo, ur, vr);
    }
    UnitGains<N> G;
    unit_gains<PQ, UW, N>(th, k, bits, c.Vq, G);
    // mean chain in (P,q) form: one FMA / one MUL per step
    double q[N + 1], P[N + 1];
    q[0] = c.q;
    P[0] = c.P;
#pragma unroll
    for (int j = 0; j < N; j++) {
        const double beta = fma(k.A * G.K[j], ymd[j], Bu[j]);
        q[j + 1] = fma(G.alpha[j], q[j], beta);
        P[j + 1] = G.alpha[j] * P[j];
    }
    // innovations (quadratic in x_in) and the affine backward steps
    double g0[N], gP[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        const double d0 = fma(-th.C, q[j], ymd[j]); // innovation for x_in = 0
        const double w0 = G.rS[j] * d0;
        c.l0 = fma(w0, d0, c.l0);
        c.l1 = fma(w0, P[j], c.l1);
        c.l2 = fma(G.rS[j] * P[j], P[j], c.l2);
        const double xu0 = fma(G.K[j], d0, q[j]);
        const double xuP = P[j] * fma(-G.K[j], th.C, 1.0);
        g0[j] = fma(-G.J[j], q[j + 1], xu0);
        gP[j] = fma(-G.J[j], P[j + 1], xuP);
    }
    // backward map of the piece so far
    double pj = c.PJ, pj2 = c.PJ2;
#pragma unroll
    for (int j = 0; j < N; j++) {
        c.G0 = fma(pj, g0[j], c.G0);
        c.GG = fma(pj, gP[j], c.GG);
        c.Lc = fma(pj2, G.L[j], c.Lc);
        pj *= G.J[j];
        pj2 *= G.J[j] * G.J[j];
    }
    c.PJ = pj;
    c.PJ2 = pj2;
    c.q = q[N];
    c.P = P[N];
    c.Vq = G.Vnext;
    c.dprod *= G.dend;
    {
        const int e = ((__double2hiint(c.dprod) >> 20) & 0x7ff) - 1023;
        c.dprod *= __hiloint2double((1023 - e) << 20, 0);
        c.shift += e;
    }
}

// ---- P2 over one unobserved unit ------------------------------------------------------------------
template <int PQ, int UW>
__device__ __forceinline__ void forward_word_basis(const Theta<PQ> &th, const SplitConst<PQ, UW> &k,
                                                   const double *__restrict__ useg, double *__restrict__ wk,
                                                   PieceFwd &c) {
    // zero-state response of the unit, sum_t A^(r-1-t) B u_t.  With KEEP_W it is formed as B . W with
    // the Horner VECTOR W = sum_t A^(r-1-t) u_t, which P4 needs again (smooth_word) and costs one FMA
    // per step less than dot-then-Horner.
    constexpr bool KEEP_W = split_w_in_smem(PQ);
    const double A4 = k.A2 * k.A2;
    double hh = 0.0;
    double W[KEEP_W ? PQ : 1];
    if (KEEP_W) {
#pragma unroll
        for (int i = 0; i < PQ; i++) W[i] = 0.0;
    }
#pragma unroll 1
    for (int b = 0; b < UW / 8; b++) {
        double ur[8 * PQ];
        load_vec<8 * PQ>(useg + b * 8 * PQ, ur);
        if (KEEP_W) {
#pragma unroll
            for (int i = 0; i < PQ; i++) { // per component: two Horner chains of four steps
                double lo = ur[i], hi = ur[4 * PQ + i];
#pragma unroll
                for (int j = 1; j < 4; j++) {
                    lo = fma(k.A, lo, ur[j * PQ + i]);
                    hi = fma(k.A, hi, ur[(4 + j) * PQ + i]);
                }
                W[i] = fma(W[i], k.A8, fma(lo, A4, hi));
            }
        } else {
            double Bu[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                double acc = 0.0;
#pragma unroll
                for (int i = 0; i < PQ; i++) acc = fma(th.b(i), ur[j * PQ + i], acc);
                Bu[j] = acc;
            }
            double lo = Bu[0], hi = Bu[4];
#pragma unroll
            for (int j = 1; j < 4; j++) {
                lo = fma(k.A, lo, Bu[j]);
                hi = fma(k.A, hi, Bu[4 + j]);
            }
            hh = fma(hh, k.A8, fma(lo, A4, hi));
        }
    }
    if (KEEP_W) {
#pragma unroll
        for (int i = 0; i < PQ; i++) {
            hh = fma(th.b(i), W[i], hh);
            wk[i * 32] = W[i];
        }
    }
    const double AW = k.AW();
    const double qn = fma(AW, c.q, hh);
    const double Pn = AW * c.P;
    const double Vn = fma(k.aVW, c.Vq, k.bVW);
    // backward map of the unit: J = prod J_t = A^UW Vp_first / Vp_last  (EM.cpp:100 telescoped)
    const double Jc = AW * c.Vq * fast_rcp(Vn);
    const double g0 = fma(-Jc, qn, c.q);
    const double gP = fma(-Jc, Pn, c.P);
    const double L = c.Vq * fma(-AW, Jc, 1.0);
    c.G0 = fma(c.PJ, g0, c.G0);
    c.GG = fma(c.PJ, gP, c.GG);
    c.Lc = fma(c.PJ2, L, c.Lc);
    c.PJ *= Jc;
    c.PJ2 *= Jc * Jc;
    c.q = qn;
    c.P = Pn;
    c.Vq = Vn;
}

// ---- P4 over one M unit -------------------------------------------------------------------------
// The filter is recomputed from the checkpoint (Xq,Vq), then the backward recursion
// Xs = J Xs1 + g, Vs = J^2 Vs1 + L runs over it and feeds the sums of EM.cpp:151-161, 180-193.
// `last` (N = 1 only): the unit is step T-1, which has no transition after it.
template <int PQ, int UW, int N>
__device__ __forceinline__ void smooth_unit(const Theta<PQ> &th, const SplitConst<PQ, UW> &k, unsigned bits, bool last,
                                            const double *__restrict__ yseg, const double *__restrict__ useg,
                                            const double *__restrict__ vseg, double Xq, double Vq, double &Xs1,
                                            double &Vs1, Stats<PQ> &st) {
    constexpr bool KEEP_ROWS = N * PQ <= 16; // keep u, v rows in registers for the sums, else reload
    double Bu[N], ymd[N], yo[N];
    double ur[N * PQ], vr[N * PQ];
    unit_inputs<PQ, N>(th, bits, yseg, useg, vseg, Bu, ymd, yo, ur, vr);
    UnitGains<N> G;
    unit_gains<PQ, UW, N>(th, k, bits, Vq, G);
    double xq[N + 1], g[N];
    xq[0] = Xq;
#pragma unroll
    for (int j = 0; j < N; j++) xq[j + 1] = fma(G.alpha[j], xq[j], fma(k.A * G.K[j], ymd[j], Bu[j]));
#pragma unroll
    for (int j = 0; j < N; j++) {
        const double delta = fma(-th.C, xq[j], ymd[j]);
        const double xu = fma(G.K[j], delta, xq[j]);
        g[j] = fma(-G.J[j], xq[j + 1], xu);
    }
    double Xs[N + 1], Vs[N + 1];
    Xs[N] = Xs1;
    Vs[N] = Vs1;
#pragma unroll
    for (int j = N - 1; j >= 0; j--) {
        Xs[j] = fma(G.J[j], Xs[j + 1], g[j]);
        Vs[j] = fma(G.J[j] * G.J[j], Vs[j + 1], G.L[j]);
    }
    if (!KEEP_ROWS) {
        load_vec<N * PQ>(useg, ur);
        load_vec<N * PQ>(vseg, vr);
    }
#pragma unroll
    for (int j = 0; j < N; j++) {
        if (N == 1 && last) {
            st.XT = Xs[j];
            st.VT = Vs[j];
        } else {
            st.Tx1x = fma(Xs[j + 1], Xs[j], st.Tx1x);
            st.Tx1xv = fma(Vs[j + 1], G.J[j], st.Tx1xv);
            st.Txx = fma(Xs[j], Xs[j], st.Txx);
            st.Txxv += Vs[j];
#pragma unroll
            for (int i = 0; i < PQ; i++) {
                st.Tx1u[i] = fma(Xs[j + 1], ur[j * PQ + i], st.Tx1u[i]);
                st.Tux[i] = fma(ur[j * PQ + i], Xs[j], st.Tux[i]);
            }
        }
        const bool obs = (bits >> j) & 1u;
        const double xo = obs ? Xs[j] : 0.0;
        st.Syx = fma(yo[j], xo, st.Syx);
        st.Sxx = fma(xo, xo, st.Sxx);
        st.Sxxv += obs ? Vs[j] : 0.0;
#pragma unroll
        for (int i = 0; i < PQ; i++) st.Sxv[i] = fma(xo, vr[j * PQ + i], st.Sxv[i]);
    }
    Xs1 = Xs[0];
    Vs1 = Vs[0];
}

// ---- short unobserved units (UNIT_US): N steps nobody observes, inside a mixed window ------------
template <int N> struct ShortConst {
    double AN, aV, bV; // A^N ; Vp' = aV Vp + bV over the unit
    __device__ __forceinline__ ShortConst(double A, double A2, double Q) {
        double an = 1.0, av = 1.0, sv = 0.0;
#pragma unroll
        for (int j = 0; j < N; j++) {
            an *= A;
            sv = fma(sv, A2, 1.0);
            av *= A2;
        }
        AN = an;
        aV = av;
        bV = Q * sv;
    }
};
// P2: the unit as one affine step of the mean and of the variance, and its backward map (telescoped)
template <int PQ, int N>
__device__ __forceinline__ void forward_short_basis(const Theta<PQ> &th, double A, double A2, double Q,
                                                    const double *__restrict__ useg, PieceFwd &c) {
    const ShortConst<N> sc(A, A2, Q);
    double hh = 0.0;
#pragma unroll
    for (int j = 0; j < N; j++) {
        double bu = 0.0;
#pragma unroll
        for (int i = 0; i < PQ; i++) bu = fma(th.b(i), useg[j * PQ + i], bu);
        hh = fma(A, hh, bu);
    }
    const double qn = fma(sc.AN, c.q, hh), Pn = sc.AN * c.P, Vn = fma(sc.aV, c.Vq, sc.bV);
    const double Jc = sc.AN * c.Vq * fast_rcp(Vn);
    const double g0 = fma(-Jc, qn, c.q), gP = fma(-Jc, Pn, c.P), L = c.Vq * fma(-sc.AN, Jc, 1.0);
    c.G0 = fma(c.PJ, g0, c.G0);
    c.GG = fma(c.PJ, gP, c.GG);
    c.Lc = fma(c.PJ2, L, c.Lc);
    c.PJ *= Jc;
    c.PJ2 *= Jc * Jc;
    c.q = qn;
    c.P = Pn;
    c.Vq = Vn;
}
// P4: streamed forward like smooth_word (same run constants cG, cH at the right end), sums taken per step
template <int PQ, int N>
__device__ __forceinline__ void smooth_short(const Theta<PQ> &th, double A, double A2, double Q,
                                             const double *__restrict__ useg, double Xq, double Vq, double &cG,
                                             double &cH, double &Xs1, double &Vs1, Stats<PQ> &st) {
    double Gs[N], Hs[N]; // Q G and H at steps 1..N of the unit
    Gs[N - 1] = Q * cG;
    Hs[N - 1] = cH;
#pragma unroll
    for (int j = N - 2; j >= 0; j--) {
        Gs[j] = A * Gs[j + 1];
        Hs[j] = A2 * Hs[j + 1];
    }
    const ShortConst<N> sc(A, A2, Q);
    const double G0 = sc.AN * cG, H0 = sc.aV * cH;
    double Xn[N + 1], vp[N + 1], Vn[N + 1];
    Xn[0] = fma(Vq, G0, Xq);
    vp[0] = Vq;
    Vn[0] = fma(Vq, Vq * H0, Vq);
    double tv = 0.0;
#pragma unroll
    for (int j = 0; j < N; j++) {
        double inp = Gs[j];
#pragma unroll
        for (int i = 0; i < PQ; i++) inp = fma(th.b(i), useg[j * PQ + i], inp);
        Xn[j + 1] = fma(A, Xn[j], inp); // Xs_{t+1} = A Xs_t + B u_t + Q G_{t+1}
        vp[j + 1] = fma(A2, vp[j], Q);
        const double t1 = vp[j + 1] * Hs[j];
        Vn[j + 1] = fma(vp[j + 1], t1, vp[j + 1]);
        st.Tx1x = fma(Xn[j + 1], Xn[j], st.Tx1x);
        st.Txx = fma(Xn[j], Xn[j], st.Txx);
        st.Txxv += Vn[j];
        tv = fma(vp[j], 1.0 + t1, tv); // V_{t+1} J_t = A Vp_t (1 + Vp_{t+1} H_{t+1})
#pragma unroll
        for (int i = 0; i < PQ; i++) {
            st.Tx1u[i] = fma(Xn[j + 1], useg[j * PQ + i], st.Tx1u[i]);
            st.Tux[i] = fma(useg[j * PQ + i], Xn[j], st.Tux[i]);
        }
    }
    st.Tx1xv = fma(A, tv, st.Tx1xv);
    Xs1 = Xn[0];
    Vs1 = Vn[0];
    cG = G0;
    cH = H0;
}

// ---- variance sums of an unobserved unit in closed form ------------------------------------------
// Inside a U unit of n = UW steps that starts with prior variance v0 and whose right end carries
// H = hR:   Vp_s = a^s v0 + Q sig_s  (a = A^2, sig_s = 1 + a + ... + a^(s-1)),   H_s = a^(n-s) hR,
//   sum_s Vs_s                    = sum Vp_s + hR sum Vp_s^2 a^(n-s)          (-> Txx's variance part)
//   sum_s Vp_s (1 + Vp_{s+1} H_{s+1}) = sum Vp_s + hR sum Vp_s Vp_{s+1} a^(n-1-s)  (-> Tx1x's, times A)
// are quadratics in v0 whose seven coefficients depend on (A, Q) only -- the same for every U unit of
// the iteration.  They are built once per iteration by exact recurrences (no 1/(1-a): nothing
// degenerates as A -> 1) by a warp that has no P1 work, and shared through shared memory.
constexpr int SPLIT_NUV = 7;
template <int UW> __device__ __forceinline__ void uvar_constants(double a, double Q, double aW, double *o) {
    double sig = 0.0, b1 = 0.0, W = 0.0, Z = 0.0, pw = 1.0, pw_prev = 1.0;
#pragma unroll 4
    for (int s = 0; s < UW; s++) {
        const double sn = fma(a, sig, 1.0); // sig_{s+1}
        b1 += sig;
        W = fma(a, W, sig * sig);           // sum_{k<=s} sig_k^2 a^(s-k)
        Z = fma(a, Z, sig * sn);            // sum_{k<=s} sig_k sig_{k+1} a^(s-k)
        sig = sn;
        pw_prev = pw;
        pw *= a;
    }
    const double Q2 = Q * Q;
    o[0 * 32] = sig;                                         // sum Vp = K0 v0 + K1
    o[1 * 32] = Q * b1;
    o[2 * 32] = aW * sig;                                    // v0^2 coefficient of both quadratics
    o[3 * 32] = 2.0 * Q * aW * b1;                           // sum Vp^2 a^(n-s): v0 coefficient
    o[4 * 32] = Q2 * (a * W);                                //                   constant
    o[5 * 32] = Q * fma(pw_prev, b1 + sig, aW * b1);         // sum Vp Vp' a^(n-1-s): v0 coefficient
    o[6 * 32] = Q2 * Z;                                      //                       constant
}

// ---- P4 over one unobserved unit ------------------------------------------------------------------
// (Xq,Vq): prior at the first step of the unit.  (cG,cH): the run constants AT THE RIGHT END of the
// unit; on return they are the constants at its left end (= right end of the unit before it).
template <int PQ, int UW>
__device__ __forceinline__ void smooth_word(const Theta<PQ> &th, const SplitConst<PQ, UW> &k,
                                            const double *__restrict__ useg, const double *__restrict__ uv,
                                            const double *__restrict__ wk, const double *__restrict__ tuu_w,
                                            double Xq, double Vq, double &cG, double &cH, double &Xs1, double &Vs1,
                                            Stats<PQ> &st) {
    constexpr int NB = UW / 8;
    constexpr bool KEEP_W = split_w_in_smem(PQ);
    // Q G at the right end of each 8-step block, last block first
    const double qg0 = k.Q * cG;
    const double g1 = k.A8 * qg0, g2 = k.A16 * qg0, g3 = k.A8 * g2;
    const double G0 = (NB == 4 ? k.A8 * k.A16 * k.A8 : k.A16) * cG; // G at the left end: A^UW cG
    const double H0 = k.aVW * cH;                                    // A^(2 UW) hR
    const double Xfirst = fma(Vq, G0, Xq), Vfirst = fma(Vq, Vq * H0, Vq);
    { // the variance sums of the whole unit, closed form (uvar_constants)
        const double sumVp = fma(uv[0 * 32], Vq, uv[1 * 32]);
        const double qa = uv[2 * 32] * Vq;
        st.Txxv += fma(cH, fma(Vq, qa + uv[3 * 32], uv[4 * 32]), sumVp);
        st.Tx1xv = fma(k.A, fma(cH, fma(Vq, qa + uv[5 * 32], uv[6 * 32]), sumVp), st.Tx1xv);
    }
    // Inside the run the smoothed mean obeys its own forward recursion,
    //     Xs_{t+1} = A Xs_t + B u_t + Q G_{t+1}
    // (from Xs = Xp + Vp G, G_t = A G_{t+1}, Vp_{t+1} = A^2 Vp_t + Q), so neither Xp nor Vp is needed
    // per step.
    double Xs = Xfirst;
    double tux[KEEP_W ? PQ : 1]; // this unit's sum u_t Xs_t (for the Tx1u identity below)
    if (KEEP_W) {
#pragma unroll
        for (int i = 0; i < PQ; i++) tux[i] = 0.0;
    }
#pragma unroll 1
    for (int b = 0; b < NB; b++) {
        const int r = NB - 1 - b; // blocks to the right of this one
        const double Gb = r == 0 ? qg0 : (r == 1 ? g1 : (r == 2 ? g2 : g3));
        const double *__restrict__ blk = useg + b * 8 * PQ;
        // stage by stage: every stage is a batch of independent operations except the Xs chain
        constexpr bool KEEP_U = PQ <= 4; // keep the rows in registers for the sums, else reload them
        double uk[8 * PQ];
        load_vec<8 * PQ>(blk, uk);
        double Gs[8];
        Gs[7] = Gb;
#pragma unroll
        for (int j = 6; j >= 0; j--) Gs[j] = k.A * Gs[j + 1];
        double inp[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
            double acc = Gs[j]; // Q G_{t+1}
#pragma unroll
            for (int i = 0; i < PQ; i++) acc = fma(th.b(i), uk[j * PQ + i], acc);
            inp[j] = acc;
        }
        double Xn[9];
        Xn[0] = Xs;
#pragma unroll
        for (int j = 0; j < 8; j++) Xn[j + 1] = fma(k.A, Xn[j], inp[j]);
        // the sums of EM.cpp:180-193
        if (!KEEP_U) load_vec<8 * PQ>(blk, uk);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            st.Tx1x = fma(Xn[j + 1], Xn[j], st.Tx1x);
            st.Txx = fma(Xn[j], Xn[j], st.Txx);
#pragma unroll
            for (int i = 0; i < PQ; i++) {
                if (KEEP_W) {
                    tux[i] = fma(uk[j * PQ + i], Xn[j], tux[i]);
                } else {
                    st.Tx1u[i] = fma(Xn[j + 1], uk[j * PQ + i], st.Tx1u[i]);
                    st.Tux[i] = fma(uk[j * PQ + i], Xn[j], st.Tux[i]);
                }
            }
        }
        Xs = Xn[8];
    }
    if (KEEP_W) {
        // sum_t Xs_{t+1} u_t' = A sum_t Xs_t u_t' + B sum_t u_t u_t' + Q cG W'   (the recursion above times u_t')
#pragma unroll
        for (int i = 0; i < PQ; i++) {
            double acc = fma(qg0, wk[i * 32], k.A * tux[i]);
#pragma unroll
            for (int j = 0; j < PQ; j++) acc = fma(th.b(j), tuu_w[j * PQ + i], acc);
            st.Tx1u[i] += acc;
            st.Tux[i] += tux[i];
        }
    }
    Xs1 = Xfirst;
    Vs1 = Vfirst;
    cG = G0;
    cH = H0;
}

// partial M-step sums <-> shared memory ([NST][32] doubles per slot, pointer already offset by lane)
template <int PQ> __device__ __forceinline__ void stats_store(const Stats<PQ> &st, double *o) {
    o[0 * 32] = st.Syx;
    o[1 * 32] = st.Sxx;
    o[2 * 32] = st.Sxxv;
    o[3 * 32] = st.Tx1x;
    o[4 * 32] = st.Tx1xv;
    o[5 * 32] = st.Txx;
    o[6 * 32] = st.Txxv;
    o[7 * 32] = st.X0;
    o[8 * 32] = st.V0;
    o[9 * 32] = st.XT;
    o[10 * 32] = st.VT;
#pragma unroll
    for (int i = 0; i < PQ; i++) {
        o[(11 + i) * 32] = st.Sxv[i];
        o[(11 + PQ + i) * 32] = st.Tx1u[i];
        o[(11 + 2 * PQ + i) * 32] = st.Tux[i];
    }
}
template <int PQ> __device__ __forceinline__ void stats_add(Stats<PQ> &st, const double *o) {
    st.Syx += o[0 * 32];
    st.Sxx += o[1 * 32];
    st.Sxxv += o[2 * 32];
    st.Tx1x += o[3 * 32];
    st.Tx1xv += o[4 * 32];
    st.Txx += o[5 * 32];
    st.Txxv += o[6 * 32];
    st.X0 += o[7 * 32];
    st.V0 += o[8 * 32];
    st.XT += o[9 * 32];
    st.VT += o[10 * 32];
#pragma unroll
    for (int i = 0; i < PQ; i++) {
        st.Sxv[i] += o[(11 + i) * 32];
        st.Tx1u[i] += o[(11 + PQ + i) * 32];
        st.Tux[i] += o[(11 + 2 * PQ + i) * 32];
    }
}
// With many inputs the NW slots of partial sums would be the largest shared-memory consumer
// (PQ = 10: 42 KB, one CTA per SM).  Then the sums are reduced in two stages through NW/2 slots:
// the upper half of the warps publish, the lower half add their own and publish, everybody sums.
template <int PQ, int NW> __host__ __device__ constexpr bool split_pairwise() {
    return NW % 2 == 0 && NW * (11 + 3 * PQ) > 2 * NW * SPLIT_NCH;
}
template <int PQ, int NW> __host__ __device__ constexpr int split_st_slots() {
    return split_pairwise<PQ, NW>() ? NW / 2 : NW;
}

struct SplitParams {
    EmParams em;
    int max_units;      // capacity of the unit table / checkpoint area
    int max_uunits;     // number of UW-step windows of the longest series
    int blob_smem;      // bytes reserved for the series blob at the start of dynamic shared memory
    int cost_u, cost_m; // relative cost of a U unit and an M unit (piece balancing)
    // Iteration-level sharing of the tasks between the CTAs of a CO-RESIDENT grid (cooperative launch), for batches of
    // a little more than one wave; NULL: CTA b takes tasks b, b + gridDim.x, ... whole.  See the kernel's task loop.
    int *flags; // [n_tasks] flags[j] == epoch: the first part of task j has been written back
    int epoch;  // differs from every value left in flags by earlier launches
    // With at most one task per CTA (and exactly two CTAs per SM): which CTA takes which task is decided by the
    // progress of the tasks (order[]: most advanced first) and by where the CTAs landed (ctl: SHARE_CTL_*), and
    // CTAs that are done early keep iterating until every CTA has done its `chunk`.  NULL: CTA b takes task b.
    int *ctl;
    const int *order;
    int n_sm;
#ifdef LDSR_PHASE_CLOCKS
    long long *clk; // development build: [CTA][NW][21] cycles per phase / unit type of the iteration loop
#endif
};
#ifdef LDSR_PHASE_CLOCKS
#define LDSR_PHASE_MARK(kk)                       \
    do {                                          \
        const long long c_ = clock64();           \
        pc[kk] += c_ - tprev;                     \
        tprev = c_;                               \
    } while (0)
#define LDSR_UNIT_BEGIN() const long long ub_ = clock64()
#define LDSR_UNIT_END(base, u0) pc[(base) + (((u0) & UNIT_M) ? 1 : ((u0) & UNIT_M1) ? 2 : ((u0) & UNIT_US) ? 3 : 0)] += clock64() - ub_
#else
#define LDSR_PHASE_MARK(kk) \
    do {                    \
    } while (0)
#define LDSR_UNIT_BEGIN() \
    do {                  \
    } while (0)
#define LDSR_UNIT_END(base, u0) \
    do {                        \
    } while (0)
#endif

// cut units [a,b) into NW pieces of about equal cost: bounds[0..NW]
__device__ __forceinline__ int unit_cost(int u, int cost_u, int cost_m, int mseg) {
    if (u & UNIT_US) return cost_m / 5;
    return (u & UNIT_M) ? cost_m : ((u & UNIT_M1) ? (3 * cost_m) / (2 * mseg) : cost_u);
}
template <int NW>
__device__ inline void split_range(const int *units, int a, int b, int cost_u, int cost_m, int mseg, int *bounds) {
    long long total = 0;
    for (int i = a; i < b; ++i) total += unit_cost(units[i], cost_u, cost_m, mseg);
    bounds[0] = a;
    long long acc = 0;
    int kq = a;
    for (int w = 1; w < NW; ++w) {
        const long long target = (total * w + NW / 2) / NW;
        while (kq < b) {
            const int cu = unit_cost(units[kq], cost_u, cost_m, mseg);
            if (acc + cu / 2 >= target) break;
            acc += cu;
            kq++;
        }
        bounds[w] = kq;
    }
    bounds[NW] = b;
}

template <int PQ, int NW, int MINB, int MSEG, int UW>
__global__ void __launch_bounds__(NW * 32, MINB) em_split_kernel(const SplitParams SP) {
    static_assert(MSEG == 4 || MSEG == 8, "M unit is 4 or 8 steps");
    const EmParams &P = SP.em;
    LDSR_DYN_SMEM(smem_raw);
    LDSR_STATIC_SMEM(__align__(8) uint64_t, bar);
    constexpr int NST = split_nstat<PQ>();
    constexpr int NP = 2 * NW; // pieces
    constexpr bool PAIR = split_pairwise<PQ, NW>();
    constexpr int ST_SLOTS = split_st_slots<PQ, NW>();
    constexpr size_t CH_DOUBLES = (size_t)NP * SPLIT_NCH * 32, ST_DOUBLES = (size_t)ST_SLOTS * NST * 32;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    // The grid is an upper bound on (or, with fewer CTAs than tasks, a divisor of) the task count,
    // which the launch reads from device memory: CTA b takes tasks b, b + gridDim.x, ...
    //
    // With SP.flags the grid is co-resident (cooperative launch) and shares the tasks by ITERATIONS, so that a batch
    // of a little more than one wave costs its share of a wave instead of two (10 000 fits are 313 tasks for 296
    // slots).  A fit's state lives in global memory between launches anyway, so a task can stop after any
    // iteration and continue elsewhere with identical results.  McNaughton's wrap-around rule: the n * chunk
    // task-iterations are laid on a line, task j on [j chunk, (j+1) chunk), and cut into gridDim.x slots of
    // M = max(chunk, ceil(n chunk / gridDim.x)); a task cut by a slot boundary runs its FIRST iterations at
    // the start of the later slot and its remaining ones at the end of the earlier slot, which waits for the flag
    // the first part raises (M >= chunk: the two parts never overlap in a balanced run, so the wait is short).
    const int n_tasks = *P.n_tasks;
    unsigned phase = 0;
    //
    // With at most one task per CTA (later launches: fits have converged, the tasks were re-packed) the launch
    // takes as long as an SM that still holds two live CTAs needs for `chunk` iterations (1.37 ms against 1.0 ms
    // for a CTA that has its SM to itself), so (SP.ctl) (a) the CTAs find out where they landed (%smid, an
    // arrival counter per SM) and the 2x most advanced tasks go to x = n - n_sm SMs, one task each to the others:
    // the fits with the longest way to go run alone; (b) a CTA that has done its `chunk` iterations keeps going
    // until every CTA has (a counter in global memory, looked at once per iteration): the extra iterations are
    // the fits' own next iterations, done now instead of in the next launch.  How many iterations a fit
    // advances per launch thus depends on the run; its results do not (the state written back is exact).
    // (few registers may live across the iteration loop: only w_lo and n_it do; the rest is recomputed)
    int w_lo = 0; // where on the line of task-iterations my next segment starts (the host keeps n chunk < 2^31)
    const auto slot_len = [&]() -> int {
        const int W = n_tasks * P.chunk, M = (W + (int)gridDim.x - 1) / (int)gridDim.x;
        return M < P.chunk ? P.chunk : M;
    };
    if (SP.flags) w_lo = (int)blockIdx.x * slot_len();
    const bool ranked = SP.ctl != nullptr && n_tasks <= (int)gridDim.x;
    LDSR_STATIC_SMEM(int, s_ti);
    LDSR_STATIC_SMEM(int, s_more);
    LDSR_STATIC_SMEM(int, s_on_sm);
    for (int seg = 0;; ++seg, phase ^= 1u) {
    int ti, n_it = P.chunk;
    bool wait_first = false;
    if (ranked) {
        if (seg > 0) break;
        if (threadIdx.x == 0) {
            const unsigned sm = sm_id(2);
            if (sm >= (unsigned)SHARE_MAX_SMID) share_trap();
            const int local = atomicAdd(SP.ctl + SHARE_CTL_SLOT + sm, 1); // 0: first CTA on this SM, 1: second
            int r;
            if (local == 0) {
                r = atomicAdd(SP.ctl + SHARE_CTL_SMS, 1);
                flag_raise(SP.ctl + SHARE_CTL_RANK + sm, r + 1);
            } else {
                while ((r = ld_acquire(SP.ctl + SHARE_CTL_RANK + sm)) == 0) {
                }
                r -= 1;
            }
            // exactly two CTAs on each of n_sm SMs, or tasks would be left out: fail loudly (the host launches this
            // mode only with a co-resident grid of 2 n_sm CTAs of a kernel that fits twice on an SM)
            if (local >= 2 || r >= SP.n_sm) share_trap();
            // SMs of rank < x hold two tasks (the 2x most advanced), the others one
            const int x = n_tasks > SP.n_sm ? n_tasks - SP.n_sm : 0;
            int qi = -1;
            if (local < 2 && r < SP.n_sm) qi = r < x ? 2 * r + local : (local == 0 ? x + r : -1);
            s_ti = (qi >= 0 && qi < n_tasks) ? SP.order[qi] : -1;
            s_on_sm = r < x ? 2 : 1; // CTAs with a task on my SM
            s_more = 1;
        }
        __syncthreads();
        ti = s_ti;
        if (ti < 0) break;
    } else if (SP.flags) {
        const int W = n_tasks * P.chunk, hi = ((int)blockIdx.x + 1) * slot_len(), w_hi = hi < W ? hi : W;
        if (w_lo >= w_hi) break;
        ti = w_lo / P.chunk;
        const int t_end = (ti + 1) * P.chunk, end = t_end < w_hi ? t_end : w_hi;
        n_it = end - w_lo;
        // starting inside a task: its first n_it iterations (the slot before mine does the rest, and waits for the
        // flag raised below); stopping inside a task: its last n_it iterations, after the next slot's CTA has done
        // the first ones
        wait_first = w_lo == ti * P.chunk && end < t_end;
    } else {
        ti = (int)blockIdx.x + seg * (int)gridDim.x;
        if (ti >= n_tasks) break;
    }
    LDSR_CHECK(ti >= 0 && ti < n_tasks && n_it >= 1 && n_it <= P.chunk); // my segment is a piece of one task's chunk
    if (seg > 0) __syncthreads(); // the previous task's shared memory is dead
    if (wait_first) {
        if (threadIdx.x == 0) flag_wait(SP.flags + ti, SP.epoch);
        __syncthreads();
    }
    const int4 task = P.tasks[ti];
    const SeriesDev S = P.series[task.x];
    const int T = S.T;
    if (threadIdx.x == 0) stage_blob(smem_raw, P.blobs + S.blob_off, (unsigned)S.blob_doubles * 8u, &bar);

    const double *__restrict__ ser = reinterpret_cast<const double *>(smem_raw);
    const double *__restrict__ ys = ser + S.y_off;
    const double *__restrict__ us = ser + S.u_off;
    const double *__restrict__ vs = ser + S.v_off;
    // shared-memory carve-up after the blob; every per-lane array is [..][32] doubles
    double *const ck = reinterpret_cast<double *>(smem_raw + SP.blob_smem) + lane; // [unit][3]
    double *const MC = ck + (size_t)SP.max_units * 96;                             // [NP][4]
    double *const CH = MC + (size_t)NP * 4 * 32;                                   // [NP][SPLIT_NCH]
    double *const ST = CH; // [NW][NST]: the piece summaries are dead after barrier 2'
    int *const units =
        reinterpret_cast<int *>(CH - lane + (CH_DOUBLES > ST_DOUBLES ? CH_DOUBLES : ST_DOUBLES)); // [max_units]
    int *const pbound = units + ((SP.max_units + 3) & ~3);                                        // [NP+1]
    double *const UV = reinterpret_cast<double *>(pbound + 32) + lane;                            // [SPLIT_NUV][32]
    double *const WK = UV + 8 * 32; // [window][PQ][32]: Horner vectors of the U units (split_w_in_smem)
    double *const TB = WK + (split_w_in_smem(PQ) ? (size_t)SP.max_uunits * PQ * 32 : 0); // [PQ][32] B, then D
    double *const TD = TB + PQ * 32;                                                      // (theta_bd_in_smem)
    constexpr bool BD_SMEM = theta_bd_in_smem(PQ);

    // ---- per-lane fit state: every warp holds the same 32 fits
    const bool valid = lane < task.z;
    const int fit = P.active[task.y + (valid ? lane : 0)];
    const int grp = P.f_group[fit];
    const unsigned *__restrict__ mw = P.masks + P.g_mask_off[grp];
    const double *__restrict__ gc = P.gconst + (size_t)grp * gconst_stride(PQ);
    const double *__restrict__ tuu_inv = P.sconst + S.sconst_off;
    const double n_obs = gc[1], inv_n_obs = 1.0 / n_obs;
    constexpr int TL = theta_pad_len<PQ>();
    // a fit's state may have been written by another CTA of this launch (SP.flags): read it past L1
    Theta<PQ> th;
    {
        double g[TL];
#pragma unroll
        for (int i = 0; i < TL; i++) g[i] = ld_l2(P.theta + (size_t)fit * TL + i);
        load_theta<PQ>(th, g);
    }
    th.sb = TB;
    th.sd = TD;
    if (BD_SMEM && warp == 0) { // visible to the other warps after the barrier that ends the set-up
#pragma unroll
        for (int i = 0; i < PQ; i++) {
            TB[i * 32] = th.B[i];
            TD[i * 32] = th.D[i];
        }
    }
    double l1 = ld_l2(P.l1 + fit), l2 = ld_l2(P.l2 + fit), lik = ld_l2(P.lik + fit);
    int ne = ld_l2(P.ne + fit);
    bool live = valid && (ld_l2(P.done + fit) == 0);
    if (live && P.g_status[grp] != 0) { // Gram block not invertible: the reference would throw
        live = false;
        lik = __longlong_as_double(0x7ff8000000000000ULL);
    }

    // ---- unit table and piece bounds (warp 0).  Units are classified from the SERIES (is y finite?),
    //      not from the hold-out masks of the fits that happen to share the CTA: which code path a fit
    //      takes must not depend on its neighbours, or its result would change in the last bits with
    //      the order of the batch and with the compaction between launches.
    if (warp == 0) {
        mbar_wait(&bar, phase); // y is needed
        auto any_finite = [&](int t0, int n) -> bool { // some y_t, t0 <= t < t0+n (n <= 32), is observed
            const int t = t0 + lane;
            const double yt = (lane < n && t < T) ? ys[t] : __longlong_as_double(0x7ff8000000000000ULL);
            return __any_sync(FULL, yt == yt);
        };
        int nu = 0;
        for (int t0 = 0; t0 < T; t0 += UW) {
            const bool any = any_finite(t0, UW);
            const bool inside = t0 + UW <= T - 1;
            if (!any && inside) {
                if (lane == 0) units[nu] = t0;
                nu++;
            } else {
                split_window_units(t0, T, MSEG, UW, [&](int t, bool single) {
                    int type = single ? UNIT_M1 : UNIT_M;
#ifndef LDSR_NO_US
                    if (!single && !any_finite(t, MSEG)) type = UNIT_US;
#endif
                    if (lane == 0) units[nu] = t | type;
                    nu++;
                });
            }
        }
        __syncwarp();
        if (lane == 0) {
            // cut between the halves: where "U before, M after" (or the reverse) fits best; ties go
            // to the cut closest to the middle
            int nM = 0;
            for (int i = 0; i < nu; ++i) nM += (units[i] & (UNIT_M | UNIT_M1)) ? 1 : 0;
            int best = nu / 2, best_mis = 1 << 30, mb = 0;
            for (int s = 0; s <= nu; ++s) { // mb = M units before s
                const int ub = s - mb, ma = nM - mb, uaft = (nu - s) - ma;
                int mis = min(mb + uaft, ub + ma);
                if (s == 0 || s == nu) mis = 1 << 29; // both halves must exist when possible
                if (mis < best_mis || (mis == best_mis && abs(2 * s - nu) < abs(2 * best - nu))) {
                    best_mis = mis;
                    best = s;
                }
                if (s < nu && (units[s] & (UNIT_M | UNIT_M1))) mb++;
            }
            split_range<NW>(units, 0, best, SP.cost_u, SP.cost_m, MSEG, pbound);
            split_range<NW>(units, best, nu, SP.cost_u, SP.cost_m, MSEG, pbound + NW);
            LDSR_CHECK(nu <= SP.max_units && (T + UW - 1) / UW <= SP.max_uunits); // what the host sized the carve-up from
            for (int w = 0; w < 2 * NW; ++w) LDSR_CHECK(pbound[w] >= 0 && pbound[w] <= pbound[w + 1] && pbound[w + 1] <= nu);
        }
    }
    mbar_wait(&bar, phase);
    __syncthreads();

#ifdef LDSR_PHASE_CLOCKS
    long long pc[21] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
    bool counted = false; // (ranked) this CTA has reported its `chunk` iterations done
    for (int it = 0;; ++it) {
        if (it >= n_it) {
            if (!ranked) break;
            if (!counted) {
                if (threadIdx.x == 0) {
                    atomicAdd(SP.ctl + SHARE_CTL_FINISHED, 1);
                    // the SM's other CTA still at its share: leave the SM to it (it runs 1.4x faster alone)
                    if (atomicAdd(SP.ctl + SHARE_CTL_DONE + sm_id(2), 1) + 1 < s_on_sm) s_more = 0;
                }
                counted = true;
                __syncthreads();
            }
            if (!s_more) break; // written before the last barrier of the previous iteration
        }
        if (!__any_sync(FULL, live)) break;
        SplitConst<PQ, UW> k;
        k.set(th);

        // ================= P1: variance maps of my pieces =================
        if (warp == NW - 1) uvar_constants<UW>(k.A2, k.Q, k.aVW, UV); // this warp has no second map to build
        LDSR_PHASE_MARK(9);
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int pj = h * NW + warp;
            if (pj == NP - 1) break; // nothing is to the right of the last piece
            double m11 = 1.0, m12 = 0.0, m21 = 0.0, m22 = 1.0;
            const int ua = pbound[pj], ue = pbound[pj + 1];
            for (int un = ua; un < ue; ++un) {
                const int u0 = units[un];
                const int t0 = u0 & UNIT_T0;
                if (u0 & UNIT_M) {
                    compose_var_unit<PQ, UW, MSEG>(th, k, seg_bits(mw, t0, MSEG), m11, m12, m21, m22);
                } else if (u0 & UNIT_M1) {
                    compose_var_unit<PQ, UW, 1>(th, k, seg_bits(mw, t0, 1), m11, m12, m21, m22);
                } else if (u0 & UNIT_US) {
                    const ShortConst<MSEG> sc(k.A, k.A2, k.Q);
                    m11 = fma(sc.aV, m11, sc.bV * m21);
                    m12 = fma(sc.aV, m12, sc.bV * m22);
                } else {
                    m11 = fma(k.aVW, m11, k.bVW * m21);
                    m12 = fma(k.aVW, m12, k.bVW * m22);
                }
            }
            MC[(pj * 4 + 0) * 32] = m11;
            MC[(pj * 4 + 1) * 32] = m12;
            MC[(pj * 4 + 2) * 32] = m21;
            MC[(pj * 4 + 3) * 32] = m22;
        }
        LDSR_PHASE_MARK(0);
        __syncthreads();
        LDSR_PHASE_MARK(1);
        double VinA = th.V1, VinB = th.V1; // prior variance entering piece warp / piece NW+warp
        {
            double n = th.V1, d = 1.0;
#pragma unroll
            for (int pj = 0; pj < NP - 1; ++pj) {
                const double m11 = MC[(pj * 4 + 0) * 32], m12 = MC[(pj * 4 + 1) * 32];
                const double m21 = MC[(pj * 4 + 2) * 32], m22 = MC[(pj * 4 + 3) * 32];
                // every map was normalised to entries summing to [1,2): the chain cannot overflow
                const double nn = fma(m11, n, m12 * d), dd = fma(m21, n, m22 * d);
                n = nn;
                d = dd;
                if (pj + 1 == warp) VinA = n * fast_rcp(d);
                if (pj + 1 == NW + warp) VinB = n * fast_rcp(d);
            }
        }

        LDSR_PHASE_MARK(10);
        // ================= P2: forward over my pieces =================
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int pj = h * NW + warp;
            const int ua = pbound[pj], ue = pbound[pj + 1];
            PieceFwd c;
            c.Vq = h ? VinB : VinA;
            c.P = 1.0;
            c.q = 0.0;
            c.l0 = c.l1 = c.l2 = 0.0;
            c.dprod = 1.0;
            c.shift = 0;
            c.PJ = c.PJ2 = 1.0;
            c.G0 = c.GG = c.Lc = 0.0;
            bool any_m = false;
            for (int un = ua; un < ue; ++un) {
                const int u0 = units[un];
                const int t0 = u0 & UNIT_T0;
                ck[(un * 3 + 0) * 32] = c.Vq;
                ck[(un * 3 + 1) * 32] = c.q;
                ck[(un * 3 + 2) * 32] = c.P;
                LDSR_UNIT_BEGIN();
                if (u0 & UNIT_M) {
                    any_m = true;
                    forward_unit_basis<PQ, UW, MSEG>(th, k, seg_bits(mw, t0, MSEG), ys + t0, us + t0 * PQ,
                                                     vs + t0 * PQ, c);
                } else if (u0 & UNIT_M1) {
                    any_m = true;
                    forward_unit_basis<PQ, UW, 1>(th, k, seg_bits(mw, t0, 1), ys + t0, us + t0 * PQ, vs + t0 * PQ, c);
                } else if (u0 & UNIT_US) {
                    forward_short_basis<PQ, MSEG>(th, k.A, k.A2, k.Q, us + t0 * PQ, c);
                } else {
                    forward_word_basis<PQ, UW>(th, k, us + t0 * PQ, WK + (size_t)(t0 / UW) * PQ * 32, c);
                }
                LDSR_UNIT_END(13, u0);
            }
            double ld = 0.0;
            if (any_m) ld = fma((double)c.shift, 0.693147180559945309417, log(c.dprod));
            double *o = CH + (size_t)pj * SPLIT_NCH * 32;
            o[0 * 32] = c.P;
            o[1 * 32] = c.q;
            o[2 * 32] = c.l0 + ld; // x_in-independent part of sum_obs (delta^2/Sigma + log Sigma)
            o[3 * 32] = c.l1;
            o[4 * 32] = c.l2;
            o[5 * 32] = c.PJ;
            o[6 * 32] = c.G0;
            o[7 * 32] = c.GG;
            o[8 * 32] = c.Lc;
            o[9 * 32] = c.Vq;
        }
        LDSR_PHASE_MARK(2);
        __syncthreads();
        LDSR_PHASE_MARK(3);

        // ---- chain the pieces: x_in of every piece, likelihood (identical in every warp)
        double gk[NP];                   // full backward offset of each piece
        double xinA = 0.0, xinB = 0.0;   // prior mean entering my pieces
        double XrA = 0.0, XrB = 0.0, VrA = 0.0, VrB = 0.0; // prior right of my pieces
        double acc = 0.0;
        double x = th.mu1; // prior of step 0 (EM.cpp:48)
        double vend = th.V1;
        {
#pragma unroll
            for (int pj = 0; pj < NP; ++pj) {
                const double *o = CH + (size_t)pj * SPLIT_NCH * 32;
                if (pj == warp) xinA = x;
                if (pj == NW + warp) xinB = x;
                gk[pj] = fma(o[7 * 32], x, o[6 * 32]);
                const double tC = th.C * x;
                acc += fma(tC, fma(tC, o[4 * 32], -2.0 * o[3 * 32]), o[2 * 32]);
                x = fma(o[0 * 32], x, o[1 * 32]);
                vend = o[9 * 32];
                if (pj == warp) {
                    XrA = x;
                    VrA = vend;
                }
                if (pj == NW + warp) {
                    XrB = x;
                    VrB = vend;
                }
            }
        }
        const double lik_new = (-0.5 * n_obs * LOG_2PI - 0.5 * acc) * inv_n_obs; // EM.cpp:122-124 (x 1/n, rounded once per task)

        // ================= stop rule (EM.cpp:259-275) =================
        if (live) {
            lik = lik_new;
            ne += 1;
            if (warp == 0 && P.liks) P.liks[(size_t)P.f_user[fit] * P.niter + (ne - 1)] = lik_new;
            const bool conv = (ne >= 3) && (fabs(lik_new - l1) < P.tol) && (fabs(l1 - l2) < P.tol);
            if (conv || ne >= P.niter) live = false;
        }
        if (!__any_sync(FULL, live)) break;

        // ---- smoothed state entering my pieces from the right.  The chain starts from the prior of
        //      the virtual step T: Xs_{T-1} = Xu + J (Xp_T - Xp_T) = Xu_{T-1}   (EM.cpp:94-95)
        double Xs1A = 0.0, Vs1A = 0.0, Xs1B = 0.0, Vs1B = 0.0;
        {
            double Xs = x, Vs = vend;
#pragma unroll
            for (int pj = NP - 1; pj >= 0; --pj) {
                if (pj == NW + warp) {
                    Xs1B = Xs;
                    Vs1B = Vs;
                }
                if (pj == warp) {
                    Xs1A = Xs;
                    Vs1A = Vs;
                }
                const double *o = CH + (size_t)pj * SPLIT_NCH * 32;
                const double pjv = o[5 * 32];
                Xs = fma(pjv, Xs, gk[pj]);
                Vs = fma(pjv * pjv, Vs, o[8 * 32]);
            }
        }
        LDSR_PHASE_MARK(4);
        __syncthreads(); // the piece summaries are dead: their space becomes the partial sums
        LDSR_PHASE_MARK(5);

        // ================= P4: backward over my pieces, M-step sums =================
        Stats<PQ> st;
        st.zero();
#pragma unroll 1
        for (int h = 1; h >= 0; --h) {
            const int pj = h * NW + warp;
            const int ua = pbound[pj], ue = pbound[pj + 1];
            const double xin = h ? xinB : xinA;
            double Xs1 = h ? Xs1B : Xs1A, Vs1 = h ? Vs1B : Vs1A;
            double Xr = h ? XrB : XrA, Vr = h ? VrB : VrA; // prior at the first step right of the piece
            double cG = 0.0, cH = 0.0;
            bool in_run = false;
            for (int un = ue - 1; un >= ua; --un) {
                const int u0 = units[un];
                const int t0 = u0 & UNIT_T0;
                const double Vq = ck[(un * 3 + 0) * 32];
                const double Xq = fma(ck[(un * 3 + 2) * 32], xin, ck[(un * 3 + 1) * 32]);
                LDSR_UNIT_BEGIN();
                if (u0 & UNIT_M) {
                    smooth_unit<PQ, UW, MSEG>(th, k, seg_bits(mw, t0, MSEG), false, ys + t0, us + t0 * PQ,
                                              vs + t0 * PQ, Xq, Vq, Xs1, Vs1, st);
                    in_run = false;
                } else if (u0 & UNIT_M1) {
                    smooth_unit<PQ, UW, 1>(th, k, seg_bits(mw, t0, 1), t0 == T - 1, ys + t0, us + t0 * PQ,
                                           vs + t0 * PQ, Xq, Vq, Xs1, Vs1, st);
                    in_run = false;
                } else {
                    if (!in_run) {
                        const double rv = fast_rcp(Vr);
                        cG = (Xs1 - Xr) * rv;
                        cH = (Vs1 - Vr) * rv * rv;
                        in_run = true;
                    }
                    if (u0 & UNIT_US)
                        smooth_short<PQ, MSEG>(th, k.A, k.A2, k.Q, us + t0 * PQ, Xq, Vq, cG, cH, Xs1, Vs1, st);
                    else
                    smooth_word<PQ, UW>(th, k, us + t0 * PQ, UV, WK + (size_t)(t0 / UW) * PQ * 32,
                                        P.uwin + S.uwin_off + (size_t)(t0 / UW) * PQ * PQ, Xq, Vq, cG, cH, Xs1, Vs1, st);
                }
                LDSR_UNIT_END(17, u0);
                Xr = Xq;
                Vr = Vq;
            }
            if (ua == 0 && ue > 0) {
                st.X0 = Xs1;
                st.V0 = Vs1;
            }
        }
        LDSR_PHASE_MARK(6);
        // (ranked) may another iteration follow the `chunk`-th?  While some CTA still has not done its share.
        if (ranked && it + 1 >= n_it && threadIdx.x == 0)
            s_more = ld_relaxed(SP.ctl + SHARE_CTL_FINISHED) + (counted ? 0 : 1) < n_tasks;
        if (!PAIR) {
            stats_store<PQ>(st, ST + (size_t)warp * NST * 32);
            __syncthreads();
            LDSR_PHASE_MARK(7);
            // ============ M-step (EM.cpp:139-229), same arithmetic in every warp ============
            st.zero();
#pragma unroll
            for (int w = 0; w < NW; ++w) stats_add<PQ>(st, ST + (size_t)w * NST * 32);
        } else {
            if (warp >= NW / 2) stats_store<PQ>(st, ST + (size_t)(warp - NW / 2) * NST * 32);
            __syncthreads();
            if (warp < NW / 2) {
                stats_add<PQ>(st, ST + (size_t)warp * NST * 32);
                stats_store<PQ>(st, ST + (size_t)warp * NST * 32);
            }
            __syncthreads();
            if (warp == 0) {
                st.zero();
#pragma unroll
                for (int w = 0; w < NW / 2; ++w) stats_add<PQ>(st, ST + (size_t)w * NST * 32);
            }
        }
        LDSR_PHASE_MARK(11);
        if (!PAIR) {
            if (live) mstep_from_stats<PQ, true>(st, gc, tuu_inv, T, th); // reciprocal multiplies: a serial section
        } else {
            // Wide inputs: the M-step is two (PQ+1)-dimensional solves, thousands of instructions.  Done once,
            // by warp 0, and published through shared memory (the partial-sum slots are dead once warp 0
            // has read them): the other warps' issue slots go to the CTA sharing the SM instead of
            // to three redundant copies.
            double g[TL];
            if (warp == 0) {
                if (BD_SMEM) { // B, D live in shared memory between M-steps, not in registers
#pragma unroll
                    for (int i = 0; i < PQ; i++) {
                        th.B[i] = TB[i * 32];
                        th.D[i] = TD[i * 32];
                    }
                }
                if (live) mstep_from_stats<PQ, true>(st, gc, tuu_inv, T, th); // reciprocal multiplies: a serial section
                if (BD_SMEM) {
#pragma unroll
                    for (int i = 0; i < PQ; i++) {
                        TB[i * 32] = th.B[i];
                        TD[i * 32] = th.D[i];
                    }
                }
                store_theta<PQ>(th, g);
#pragma unroll
                for (int i = 0; i < TL; i++) ST[i * 32] = g[i];
            }
            __syncthreads();
            if (warp != 0) {
#pragma unroll
                for (int i = 0; i < TL; i++) g[i] = ST[i * 32];
                load_theta<PQ>(th, g);
            }
        }
        if (live) {
            l2 = l1;
            l1 = lik;
        }
        LDSR_PHASE_MARK(8);
    }
#ifdef LDSR_PHASE_CLOCKS
    if (lane == 0 && SP.clk)
        for (int i = 0; i < 21; i++) SP.clk[((size_t)blockIdx.x * NW + warp) * 21 + i] = pc[i];
#endif

    if (ranked && !counted && threadIdx.x == 0) { // left the loop early: every fit of the task is done
        atomicAdd(SP.ctl + SHARE_CTL_FINISHED, 1);
        atomicAdd(SP.ctl + SHARE_CTL_DONE + sm_id(2), 1);
    }
    if (BD_SMEM && warp == 0) {
#pragma unroll
        for (int i = 0; i < PQ; i++) {
            th.B[i] = TB[i * 32];
            th.D[i] = TD[i * 32];
        }
    }
    if (warp == 0 && valid) {
        store_theta<PQ>(th, P.theta + (size_t)fit * TL);
        P.l1[fit] = l1;
        P.l2[fit] = l2;
        P.lik[fit] = lik;
        P.ne[fit] = ne;
        P.done[fit] = live ? 0 : 1;
    }
    if (SP.flags) {
        if (w_lo % P.chunk != 0 && warp == 0) { // the first part of a task: release every lane's stores, then the flag
            __threadfence();
            __syncwarp();
            if (lane == 0) flag_raise(SP.flags + w_lo / P.chunk, SP.epoch);
        }
        w_lo += n_it;
    }
    } // task loop
}

} // namespace ldsr
