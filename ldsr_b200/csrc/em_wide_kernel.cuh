// em_wide_kernel.cuh -- the LDS_EM loop (src/EM.cpp:245-280) for WIDE inputs (padded width PQ >= 5).
//
// Same cut of the recursions as em_split_kernel.cuh (lane = fit, the warps of the CTA share the time
// axis, variance maps / affine mean maps chained across pieces), but the work that touches the input
// ROWS is taken out of the recursions:
//
//   * With p = q = 10 a time step has 3 PQ = 30 multiply-adds on input rows (B.u_t, sum Xs_{t+1} u_t',
//     sum u_t Xs_t; at observed steps also D.v_t and sum Xs_t v_t') against about a dozen on the scalar
//     state.  Fused into the recursions (em_split_kernel) the 3 PQ running sums, B, D and the unit's rows
//     need well over 255 registers: 456 B of stack per thread, 96 MB of spill traffic per launch,
//     FP64 pipe 41 % (profiles/em_r01_pq10.txt).
//   * Here every iteration runs
//       phase A  Bu_t = B.u_t for all t, Dv_t = D.v_t for the steps of observed units     -> shared memory
//       P1..P4   the SCALAR recursions of em_split_kernel (variance maps, forward in the (P,q) basis,
//                chain over the pieces, backward with the closed forms), reading Bu_t / Dv_t and leaving
//                the smoothed mean Xs_t where Bu_t was
//       phase C  Tux = sum u_t Xs_t, Tx1u = sum u_t Xs_{t+1}, Sxv = sum_obs v_t Xs_t       from shared memory
//     Phases A and C have no dependency along t: every warp takes an equal slice of the time axis and
//     streams independent multiply-adds (rows by broadcast LDS.128, B / D or 2 PQ accumulators in
//     registers); the recursions carry scalars only.  Nothing spills: no phase holds more than
//     2 PQ + O(1) doubles plus one unit's scalars.
//   * One double per (fit, step) in shared memory (T x 32 x 8 B = 100 KB at T = 400) => one CTA of
//     NW = 8 warps per SM, 255 registers per thread.
//
// Phases and barriers of one iteration:
//   [A | P1 | variance-sum constants] B1  prefix of the variance maps, P2  B2  chain, likelihood, stop
//   rule, backward chain, P4  B3  C  B4  partial sums -> shared memory (over the dead trajectory)  B5
//   M-step: warp 0 the observation block (C, D, R), warp 1 the transition block (A, B, Q, mu1, V1)  B6
#pragma once
#include "em_split_kernel.cuh"

namespace ldsr {

constexpr int WIDE_NCH = 10; // per-piece values exchanged after P2 (as SPLIT_NCH)

struct WideParams {
    EmParams em;
    int max_units;  // capacity of the unit table / checkpoint area
    int max_msteps; // capacity of the Dv rows: steps inside observed units (wide_count_units)
    int max_T;      // longest series of the plan
    int blob_smem;  // bytes reserved for the series blob at the start of dynamic shared memory
    int cost_u, cost_m;
#ifdef LDSR_PHASE_CLOCKS
    long long *clk; // development build: [CTA][NW][21] cycles per phase of the iteration loop
#endif
};

// Where phases A and C read the input rows from.  0: the whole series blob (y, u, v) is staged in shared memory;
// 1: only y is staged and the rows come from global memory (warp-uniform addresses, L1/L2 hits: every CTA of a
// series reads the same 32 KB) -- which frees 32 KB per CTA for more warps (pieces).  Build-time switch.
#ifndef LDSR_WIDE_ROWS_GLOBAL
#define LDSR_WIDE_ROWS_GLOBAL 0
#endif
__host__ __device__ constexpr bool wide_rows_global() { return LDSR_WIDE_ROWS_GLOBAL != 0; }
// bytes reserved at the start of dynamic shared memory for the staged part of the blob
__host__ __device__ inline size_t wide_blob_smem(size_t max_blob_bytes, int max_T) {
    const size_t b = wide_rows_global() ? (size_t)((max_T + 1) & ~1) * 8 : max_blob_bytes;
    return (b + 127) & ~size_t(127);
}

// Rows of the trajectory area: one per step.  After phase C the partial sums (NW slots), their totals and the
// 3 PQ matrix-vector rows of the M-step alias the trajectory AND the two areas that follow it in the carve-up
// (Dv rows, checkpoints: both dead by then); the trajectory is only enlarged when even that is too small.
__host__ __device__ inline size_t wide_traj_rows(int pq, int nw, int max_T, int max_units, int max_msteps) {
    const size_t need = (size_t)(nw + 1) * (11 + 3 * pq) + 3 * pq;
    const size_t behind = (size_t)(max_msteps > 0 ? max_msteps : 1) + (size_t)3 * max_units;
    const size_t extra = need > behind ? need - behind : 0;
    return (size_t)max_T > extra ? (size_t)max_T : extra;
}
// dynamic shared memory after the series blob, in bytes
__host__ __device__ inline size_t wide_smem_bytes(int pq, int nw, int max_T, int max_units, int max_msteps) {
    size_t b = 0;
    b += wide_traj_rows(pq, nw, max_T, max_units, max_msteps) * 256; // TR: Bu_t, then Xs_t; later the partial sums
    b += (size_t)(max_msteps > 0 ? max_msteps : 1) * 256; // YM: Dv of the steps of observed units
    b += (size_t)max_units * 3 * 256;           // CK: checkpoints (Vq, q, P) per unit
    b += (size_t)nw * 4 * 256;                  // MC: variance maps of the pieces
    b += (size_t)nw * WIDE_NCH * 256;           // CH: piece summaries
    b += (size_t)8 * 256;                       // UV: closed-form variance-sum coefficients
    b += (size_t)2 * pq * 256;                  // TB, TD: B and D of the CTA's fits
    b += (size_t)6 * 256;                       // TH: A, C, Q, R, mu1, V1 after the M-step
    b += (size_t)((max_T + 31) / 32) * 128;     // MW: observed-bit words of the CTA's fits
    b += 2 * (((size_t)max_units * 4 + 15) & ~size_t(15)); // unit table, Dv row of each unit
    b += 256;                                   // piece bounds, slice bounds
    b += ((size_t)(max_msteps > 0 ? max_msteps : 1) * 4 + 15) & ~size_t(15); // time step of every Dv row
    return b;
}

// (number of units, number of steps inside observed units M / M1) of a series, from its finite(y)
// pattern -- exactly the classification the kernel makes (hold-outs only remove observations: a
// group's units are the series' units)
inline void wide_count_units(const double *y, int T, int mseg, int uw, int *n_units, int *n_msteps) {
    int nu = 0, nm = 0;
    for (int t0 = 0; t0 < T; t0 += uw) {
        bool any = false;
        for (int t = t0; t < T && t < t0 + uw; t++) any = any || (y[t] == y[t]);
        const bool inside = t0 + uw <= T - 1;
        if (!any && inside) {
            nu += 1;
            continue;
        }
        split_window_units(t0, T, mseg, uw, [&](int t, bool single) {
            nu++;
            if (single) {
                nm += 1;
                return;
            }
            bool a = false;
            for (int j = 0; j < mseg; j++) a = a || (y[t + j] == y[t + j]);
            if (a) nm += mseg;
        });
    }
    *n_units = nu;
    *n_msteps = nm;
}

// ---- P2 over one observed unit: Bu_t and Dv_t come from shared memory ----------------------------
template <int PQ, int UW, int N>
__device__ __forceinline__ void wide_forward_unit(const Theta<PQ> &th, const SplitConst<PQ, UW> &k, unsigned bits,
                                                  const double *__restrict__ yseg, const double *__restrict__ bu,
                                                  const double *__restrict__ dv, PieceFwd &c) {
    double Bu[N], ymd[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        const bool obs = (bits >> j) & 1u;
        Bu[j] = bu[j * 32];
        ymd[j] = (obs ? yseg[j] : 0.0) - dv[j * 32]; // y is NaN where missing
    }
    UnitGains<N> G;
    unit_gains<PQ, UW, N>(th, k, bits, c.Vq, G);
    double q[N + 1], P[N + 1];
    q[0] = c.q;
    P[0] = c.P;
#pragma unroll
    for (int j = 0; j < N; j++) {
        const double beta = fma(k.A * G.K[j], ymd[j], Bu[j]);
        q[j + 1] = fma(G.alpha[j], q[j], beta);
        P[j + 1] = G.alpha[j] * P[j];
    }
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

// the tail every unobserved unit shares in P2: the unit as one affine step (AN, hh) of the mean and
// (aV, bV) of the variance, and its telescoped backward map
__device__ __forceinline__ void wide_unobserved_tail(double AN, double aV, double bV, double hh, PieceFwd &c) {
    const double qn = fma(AN, c.q, hh), Pn = AN * c.P, Vn = fma(aV, c.Vq, bV);
    const double Jc = AN * c.Vq * fast_rcp(Vn); // J = prod J_t = A^n Vp_first / Vp_last  (EM.cpp:100 telescoped)
    const double g0 = fma(-Jc, qn, c.q), gP = fma(-Jc, Pn, c.P), L = c.Vq * fma(-AN, Jc, 1.0);
    c.G0 = fma(c.PJ, g0, c.G0);
    c.GG = fma(c.PJ, gP, c.GG);
    c.Lc = fma(c.PJ2, L, c.Lc);
    c.PJ *= Jc;
    c.PJ2 *= Jc * Jc;
    c.q = qn;
    c.P = Pn;
    c.Vq = Vn;
}
// ---- P2 over one unobserved unit of UW steps: zero-state response by Horner on Bu_t ---------------
template <int PQ, int UW>
__device__ __forceinline__ void wide_forward_word(const SplitConst<PQ, UW> &k, const double *__restrict__ bu,
                                                  PieceFwd &c) {
    const double A4 = k.A2 * k.A2;
    double hh = 0.0;
#pragma unroll
    for (int b = 0; b < UW / 8; b++) {
        double lo = bu[(b * 8) * 32], hi = bu[(b * 8 + 4) * 32];
#pragma unroll
        for (int j = 1; j < 4; j++) {
            lo = fma(k.A, lo, bu[(b * 8 + j) * 32]);
            hi = fma(k.A, hi, bu[(b * 8 + 4 + j) * 32]);
        }
        hh = fma(hh, k.A8, fma(lo, A4, hi));
    }
    wide_unobserved_tail(k.AW(), k.aVW, k.bVW, hh, c);
}
template <int N>
__device__ __forceinline__ void wide_forward_short(double A, double A2, double Q, const double *__restrict__ bu,
                                                   PieceFwd &c) {
    const ShortConst<N> sc(A, A2, Q);
    double hh = 0.0;
#pragma unroll
    for (int j = 0; j < N; j++) hh = fma(A, hh, bu[j * 32]);
    wide_unobserved_tail(sc.AN, sc.aV, sc.bV, hh, c);
}

// the scalar part of the M-step sums (the row sums are taken in phase C)
struct WideSums {
    double Syx, Sxx, Sxxv, Tx1x, Tx1xv, Txx, Txxv, X0, V0, XT, VT;
    __device__ __forceinline__ void zero() { Syx = Sxx = Sxxv = Tx1x = Tx1xv = Txx = Txxv = X0 = V0 = XT = VT = 0.0; }
};

// ---- P4 over one observed unit: filter recomputed from the checkpoint, backward recursion, scalar
//      sums of EM.cpp:151-152, 180-183; the smoothed means replace Bu_t in shared memory ------------
template <int PQ, int UW, int N>
__device__ __forceinline__ void wide_smooth_unit(const Theta<PQ> &th, const SplitConst<PQ, UW> &k, unsigned bits,
                                                 bool last, const double *__restrict__ yseg, double *__restrict__ bu,
                                                 const double *__restrict__ dv, double Xq, double Vq, double &Xs1,
                                                 double &Vs1, WideSums &st) {
    double Bu[N], ymd[N], yo[N];
#pragma unroll
    for (int j = 0; j < N; j++) {
        const bool obs = (bits >> j) & 1u;
        Bu[j] = bu[j * 32];
        yo[j] = obs ? yseg[j] : 0.0;
        ymd[j] = yo[j] - dv[j * 32];
    }
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
#pragma unroll
    for (int j = 0; j < N; j++) {
        bu[j * 32] = Xs[j];
        if (N == 1 && last) {
            st.XT = Xs[j];
            st.VT = Vs[j];
        } else {
            st.Tx1x = fma(Xs[j + 1], Xs[j], st.Tx1x);
            st.Tx1xv = fma(Vs[j + 1], G.J[j], st.Tx1xv);
            st.Txx = fma(Xs[j], Xs[j], st.Txx);
            st.Txxv += Vs[j];
        }
        const bool obs = (bits >> j) & 1u;
        const double xo = obs ? Xs[j] : 0.0;
        st.Syx = fma(yo[j], xo, st.Syx);
        st.Sxx = fma(xo, xo, st.Sxx);
        st.Sxxv += obs ? Vs[j] : 0.0;
    }
    Xs1 = Xs[0];
    Vs1 = Vs[0];
}

// ---- P4 over one unobserved unit of UW steps (see smooth_word in em_split_kernel.cuh: streamed forward
//      with the run constants cG, cH taken at the right end; variance sums in closed form) ------------
template <int PQ, int UW>
__device__ __forceinline__ void wide_smooth_word(const SplitConst<PQ, UW> &k, double *__restrict__ bu,
                                                 const double *__restrict__ uv, double Xq, double Vq, double &cG,
                                                 double &cH, double &Xs1, double &Vs1, WideSums &st) {
    constexpr int NB = UW / 8;
    const double qg0 = k.Q * cG;
    const double g1 = k.A8 * qg0, g2 = k.A16 * qg0, g3 = k.A8 * g2;
    const double G0 = (NB == 4 ? k.A8 * k.A16 * k.A8 : k.A16) * cG; // G at the left end: A^UW cG
    const double H0 = k.aVW * cH;
    const double Xfirst = fma(Vq, G0, Xq), Vfirst = fma(Vq, Vq * H0, Vq);
    {
        const double sumVp = fma(uv[0 * 32], Vq, uv[1 * 32]);
        const double qa = uv[2 * 32] * Vq;
        st.Txxv += fma(cH, fma(Vq, qa + uv[3 * 32], uv[4 * 32]), sumVp);
        st.Tx1xv = fma(k.A, fma(cH, fma(Vq, qa + uv[5 * 32], uv[6 * 32]), sumVp), st.Tx1xv);
    }
    double Xs = Xfirst;
#pragma unroll
    for (int b = 0; b < NB; b++) {
        const int r = NB - 1 - b; // blocks to the right of this one
        const double Gb = r == 0 ? qg0 : (r == 1 ? g1 : (r == 2 ? g2 : g3));
        double *__restrict__ blk = bu + (b * 8) * 32;
        double inp[8];
        {
            double Gs = Gb;
#pragma unroll
            for (int j = 7; j >= 0; j--) {
                inp[j] = blk[j * 32] + Gs; // B u_t + Q G_{t+1}
                Gs *= k.A;
            }
        }
        double Xn[9];
        Xn[0] = Xs;
#pragma unroll
        for (int j = 0; j < 8; j++) Xn[j + 1] = fma(k.A, Xn[j], inp[j]); // Xs_{t+1} = A Xs_t + B u_t + Q G_{t+1}
#pragma unroll
        for (int j = 0; j < 8; j++) {
            blk[j * 32] = Xn[j];
            st.Tx1x = fma(Xn[j + 1], Xn[j], st.Tx1x);
            st.Txx = fma(Xn[j], Xn[j], st.Txx);
        }
        Xs = Xn[8];
    }
    Xs1 = Xfirst;
    Vs1 = Vfirst;
    cG = G0;
    cH = H0;
}
template <int N>
__device__ __forceinline__ void wide_smooth_short(double A, double A2, double Q, double *__restrict__ bu, double Xq,
                                                  double Vq, double &cG, double &cH, double &Xs1, double &Vs1,
                                                  WideSums &st) {
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
        Xn[j + 1] = fma(A, Xn[j], bu[j * 32] + Gs[j]);
        vp[j + 1] = fma(A2, vp[j], Q);
        const double t1 = vp[j + 1] * Hs[j];
        Vn[j + 1] = fma(vp[j + 1], t1, vp[j + 1]);
        st.Tx1x = fma(Xn[j + 1], Xn[j], st.Tx1x);
        st.Txx = fma(Xn[j], Xn[j], st.Txx);
        st.Txxv += Vn[j];
        tv = fma(vp[j], 1.0 + t1, tv); // V_{t+1} J_t = A Vp_t (1 + Vp_{t+1} H_{t+1})
    }
#pragma unroll
    for (int j = 0; j < N; j++) bu[j * 32] = Xn[j];
    st.Tx1xv = fma(A, tv, st.Tx1xv);
    Xs1 = Xn[0];
    Vs1 = Vn[0];
    cG = G0;
    cH = H0;
}

// ---- phases A and C: register blocking over fits --------------------------------------------------
// With one fit per lane a broadcast row load delivers 16 bytes to each of 32 lanes -- 512 B through the
// 128 B/clk shared-memory port, 4 cycles per LDS.128 -- for two multiply-adds per lane: the row phases
// were bound by that port (8000 cycles of loads per iteration and phase against 2000-4000 cycles of FP64
// pipe; profiles/em_r02_wide_phase_clocks.txt).  So in these phases a lane takes FPL fits and a
// sub-warp GROUP of 32 / FPL lanes one time step: an LDS.128 then brings FPL different rows to the warp
// and every loaded double feeds FPL multiply-adds.  Lane l of group g = l / LG (LG = 32 / FPL lanes) holds
// the fits (l % LG) + k LG, k < FPL; the groups split the warp's slice of the time axis.  Sums over time
// come back to the lane-per-fit layout by a transposing shuffle reduction (lane l ends with fit l).
#ifdef LDSR_WIDE_FPL // development: force the blocking factor
__host__ __device__ constexpr int wide_fpl(int) { return LDSR_WIDE_FPL; }
#else
__host__ __device__ constexpr int wide_fpl(int pq) { return pq <= 16 ? 2 : 1; } // measured at PQ = 10: 2 beats 4 and 1
#endif

// sum over the FPL groups; on return out[i] is the total of MY fit (k = my group index) in every lane
template <int FPL, int NV>
__device__ __forceinline__ void group_transpose_sum(double (&v)[FPL][NV], double (&out)[NV], int lane) {
    static_assert(FPL == 1 || FPL == 2 || FPL == 4, "fits per lane");
    if constexpr (FPL == 1) {
#pragma unroll
        for (int i = 0; i < NV; i++) out[i] = v[0][i];
    } else if constexpr (FPL == 2) {
        const bool upper = (lane & 16) != 0;
#pragma unroll
        for (int i = 0; i < NV; i++) {
            const double send = upper ? v[0][i] : v[1][i], keep = upper ? v[1][i] : v[0][i];
            out[i] = keep + __shfl_xor_sync(FULL, send, 16);
        }
    } else {
        const bool upper = (lane & 16) != 0, up2 = (lane & 8) != 0;
#pragma unroll
        for (int i = 0; i < NV; i++) {
            // groups {0,1} keep the fits k = 0, 1 and groups {2,3} k = 2, 3 ...
            const double s0 = upper ? v[0][i] : v[2][i], k0 = upper ? v[2][i] : v[0][i];
            const double s1 = upper ? v[1][i] : v[3][i], k1 = upper ? v[3][i] : v[1][i];
            const double w0 = k0 + __shfl_xor_sync(FULL, s0, 16), w1 = k1 + __shfl_xor_sync(FULL, s1, 16);
            // ... then each group keeps its own
            const double send = up2 ? w0 : w1, keep = up2 ? w1 : w0;
            out[i] = keep + __shfl_xor_sync(FULL, send, 8);
        }
    }
}

// dst[m][fit] = coef[fit] . rows[t(m)] for m = 0 .. n-1, t(m) = list ? list[m] : m  (FPL fits per lane, one row per
// group at a time)
//   coef: [PQ][32] in shared memory (B or D of the CTA's fits); rows: the series' [t][PQ] rows; dst: [m][32]
template <int PQ, int FPL>
__device__ __forceinline__ void rows_times_coef(const double *__restrict__ coef, const double *__restrict__ rows,
                                                const int *__restrict__ list, double *__restrict__ dst, int n, int lane) {
    constexpr int LG = 32 / FPL;
    const int g = lane / LG, lg = lane % LG;
    double cf[FPL][PQ];
#pragma unroll
    for (int k = 0; k < FPL; k++)
#pragma unroll
        for (int i = 0; i < PQ; i++) cf[k][i] = coef[i * 32 + lg + k * LG];
    const int len = (n + FPL - 1) / FPL; // contiguous sub-slices, one per group
    const int lo = g * len, hi = (lo + len < n) ? lo + len : n;
    int m = lo;
#pragma unroll 1
    for (; m + 2 <= hi; m += 2) { // two rows in flight: 4 FPL independent chains, the loads of one row behind the other's math
        const int t0 = list ? list[m] : m, t1 = list ? list[m + 1] : m + 1;
        double r0[PQ], r1[PQ];
        load_vec<PQ>(rows + (size_t)t0 * PQ, r0);
        load_vec<PQ>(rows + (size_t)t1 * PQ, r1);
#pragma unroll
        for (int k = 0; k < FPL; k++) {
            double a0 = cf[k][0] * r0[0], a1 = cf[k][1] * r0[1], b0 = cf[k][0] * r1[0], b1 = cf[k][1] * r1[1];
#pragma unroll
            for (int i = 2; i < PQ; i++) {
                if (i & 1) {
                    a1 = fma(cf[k][i], r0[i], a1);
                    b1 = fma(cf[k][i], r1[i], b1);
                } else {
                    a0 = fma(cf[k][i], r0[i], a0);
                    b0 = fma(cf[k][i], r1[i], b0);
                }
            }
            dst[(size_t)m * 32 + lg + k * LG] = a0 + a1;
            dst[(size_t)(m + 1) * 32 + lg + k * LG] = b0 + b1;
        }
    }
    if (m < hi) {
        const int t = list ? list[m] : m;
        double r[PQ];
        load_vec<PQ>(rows + (size_t)t * PQ, r);
#pragma unroll
        for (int k = 0; k < FPL; k++) {
            double a0 = cf[k][0] * r[0], a1 = cf[k][1] * r[1];
#pragma unroll
            for (int i = 2; i < PQ; i++) {
                if (i & 1)
                    a1 = fma(cf[k][i], r[i], a1);
                else
                    a0 = fma(cf[k][i], r[i], a0);
            }
            dst[(size_t)m * 32 + lg + k * LG] = a0 + a1;
        }
    }
}

template <int PQ, int NW, int MSEG, int UW>
__global__ void __launch_bounds__(NW * 32, 1) em_wide_kernel(const WideParams WP) {
    static_assert(MSEG == 4 || MSEG == 8, "observed unit is 4 or 8 steps");
    static_assert(NW <= 16, "piece and slice bounds share one 256-byte block");
    static_assert(PQ >= 2, "the row products are split into two partial sums");
    const EmParams &P = WP.em;
    LDSR_DYN_SMEM(smem_raw);
    LDSR_STATIC_SMEM(__align__(8) uint64_t, bar);
    constexpr int NST = split_nstat<PQ>();
    constexpr int NP = NW; // one piece per warp

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int n_tasks = *P.n_tasks;
    unsigned phase = 0;
    for (int ti = blockIdx.x; ti < n_tasks; ti += gridDim.x, phase ^= 1u) {
    if (ti != (int)blockIdx.x) __syncthreads(); // the previous task's shared memory is dead
    const int4 task = P.tasks[ti];
    const SeriesDev S = P.series[task.x];
    const int T = S.T;
    constexpr bool ROWS_GLOBAL = wide_rows_global();
    // y sits first in the blob (u_off doubles, even): with ROWS_GLOBAL only that part is staged
    if (threadIdx.x == 0)
        stage_blob(smem_raw, P.blobs + S.blob_off, (unsigned)(ROWS_GLOBAL ? S.u_off : S.blob_doubles) * 8u, &bar);

    const double *__restrict__ ser = reinterpret_cast<const double *>(smem_raw);
    const double *__restrict__ gser = P.blobs + S.blob_off;
    const double *__restrict__ ys = ser + S.y_off;
    const double *__restrict__ us = (ROWS_GLOBAL ? gser : ser) + S.u_off;
    const double *__restrict__ vs = (ROWS_GLOBAL ? gser : ser) + S.v_off;
    // shared-memory carve-up after the blob; every per-lane array is [..][32] doubles, already offset by lane
    double *const TR = reinterpret_cast<double *>(smem_raw + WP.blob_smem) + lane;  // [rows]
    double *const YM = TR + wide_traj_rows(PQ, NW, WP.max_T, WP.max_units, WP.max_msteps) * 32; // [max_msteps]
    double *const CK = YM + (size_t)(WP.max_msteps > 0 ? WP.max_msteps : 1) * 32;   // [unit][3]
    double *const MC = CK + (size_t)WP.max_units * 96;                              // [NP][4]
    double *const CH = MC + (size_t)NP * 4 * 32;                                    // [NP][WIDE_NCH]
    double *const UV = CH + (size_t)NP * WIDE_NCH * 32;                             // [8]
    double *const TB = UV + 8 * 32;                                                 // [PQ] B
    double *const TD = TB + PQ * 32;                                                // [PQ] D
    double *const TH = TD + PQ * 32;                                                // [6] A, C, Q, R, mu1, V1
    unsigned *const MW = reinterpret_cast<unsigned *>(TH - lane + 6 * 32) + lane;   // [word][32]
    const int n_words = (WP.max_T + 31) / 32;
    int *const units = reinterpret_cast<int *>(MW - lane + (size_t)n_words * 32);   // [max_units]
    int *const ubase = units + ((WP.max_units + 3) & ~3);                           // [max_units] Dv row of the unit
    int *const pbound = ubase + ((WP.max_units + 3) & ~3);                          // [NP + 1]
    int *const sbound = pbound + 32;                                                // [NW + 1]
    int *const mlist = sbound + 32;                                                 // [max_msteps] time step of Dv row m
    double *const ST = TR; // [NW][NST]: the trajectory (and the Dv rows and checkpoints behind it) is dead after phase C

    // ---- per-lane fit state: every warp holds the same 32 fits
    const bool valid = lane < task.z;
    const int fit = P.active[task.y + (valid ? lane : 0)];
    const int grp = P.f_group[fit];
    const double *__restrict__ gc = P.gconst + (size_t)grp * gconst_stride(PQ);
    const double *__restrict__ tuu_inv = P.sconst + S.sconst_off;
    const double n_obs = gc[1], inv_n_obs = 1.0 / n_obs;
    constexpr int TL = theta_pad_len<PQ>();
    Theta<PQ> th;
    th.sb = TB;
    th.sd = TD;
    {
        const double *g = P.theta + (size_t)fit * TL;
        th.A = g[0];
        th.C = g[1 + PQ];
        th.Q = g[2 + 2 * PQ];
        th.R = g[3 + 2 * PQ];
        th.mu1 = g[4 + 2 * PQ];
        th.V1 = g[5 + 2 * PQ];
        if (warp == 0) { // B, D and the mask words live in shared memory; visible after the set-up barrier
#pragma unroll
            for (int i = 0; i < PQ; i++) {
                TB[i * 32] = g[1 + i];
                TD[i * 32] = g[2 + PQ + i];
            }
            const unsigned *__restrict__ mw = P.masks + P.g_mask_off[grp];
            const int nw_series = (T + 31) / 32;
            for (int w = 0; w < nw_series; ++w) MW[w * 32] = mw[w];
        }
    }
    double l1 = P.l1[fit], l2 = P.l2[fit], lik = P.lik[fit];
    int ne = P.ne[fit];
    bool live = valid && (P.done[fit] == 0);
    if (live && P.g_status[grp] != 0) { // Gram block not invertible: the reference would throw
        live = false;
        lik = __longlong_as_double(0x7ff8000000000000ULL);
    }
    auto mask_bits = [&](int t0, int n) -> unsigned { // observed bits of steps t0 .. t0+n-1 (inside one word)
        const unsigned w = MW[(t0 >> 5) * 32];
        return (w >> (t0 & 31)) & ((n == 32) ? 0xffffffffu : ((1u << n) - 1u));
    };

    // ---- unit table, piece bounds, slice bounds (warp 0).  Units are classified from the SERIES
    //      (is y finite?), never from the masks of the fits that share the CTA (see em_split_kernel.cuh).
    if (warp == 0) {
        mbar_wait(&bar, phase); // y is needed
        auto any_finite = [&](int t0, int n) -> bool {
            const int t = t0 + lane;
            const double yt = (lane < n && t < T) ? ys[t] : __longlong_as_double(0x7ff8000000000000ULL);
            return __any_sync(FULL, yt == yt);
        };
        int nu = 0, nm = 0;
        for (int t0 = 0; t0 < T; t0 += UW) {
            const bool any = any_finite(t0, UW);
            const bool inside = t0 + UW <= T - 1;
            if (!any && inside) {
                if (lane == 0) {
                    units[nu] = t0;
                    ubase[nu] = 0;
                }
                nu++;
            } else {
                split_window_units(t0, T, MSEG, UW, [&](int t, bool single) {
                    int type = single ? UNIT_M1 : UNIT_M;
                    if (!single && !any_finite(t, MSEG)) type = UNIT_US;
                    if (lane == 0) {
                        units[nu] = t | type;
                        ubase[nu] = nm;
                    }
                    nu++;
                    if (type != UNIT_US) {
                        const int len = single ? 1 : MSEG;
                        if (lane < len) mlist[nm + lane] = t + lane;
                        nm += len;
                    }
                });
            }
        }
        __syncwarp();
        if (lane == 0) {
            split_range<NW>(units, 0, nu, WP.cost_u, WP.cost_m, MSEG, pbound);
            // phases A and C: equal slices of the time axis, a step with an observation counting 3/2
            int total = 0;
            for (int t = 0; t < T; ++t) total += (ys[t] == ys[t]) ? 3 : 2;
            int acc = 0, t = 0;
            sbound[0] = 0;
            for (int w = 1; w < NW; ++w) {
                const int target = (int)(((long long)total * w) / NW);
                while (t < T && (acc < target || (t & 3))) { // slices start at multiples of 4 steps (16-byte rows)
                    acc += (ys[t] == ys[t]) ? 3 : 2;
                    ++t;
                }
                sbound[w] = t;
            }
            sbound[NW] = T;
            pbound[NP + 1] = nu; // number of units
            pbound[NP + 2] = nm; // number of Dv rows = steps inside observed units
            // the host sized the carve-up from these (wide_count_units, wide_smem_bytes)
            LDSR_CHECK(nu <= WP.max_units && nm <= WP.max_msteps && T <= WP.max_T);
            LDSR_CHECK((size_t)(NW + 1) * NST + 3 * PQ <=
                       wide_traj_rows(PQ, NW, WP.max_T, WP.max_units, WP.max_msteps) +
                           (size_t)(WP.max_msteps > 0 ? WP.max_msteps : 1) + (size_t)3 * WP.max_units);
            for (int w = 0; w < NW; ++w) {
                LDSR_CHECK(pbound[w] >= 0 && pbound[w] <= pbound[w + 1] && pbound[w + 1] <= nu);
                LDSR_CHECK(sbound[w] >= 0 && sbound[w] <= sbound[w + 1] && sbound[w + 1] <= T && (sbound[w] & 3) == 0);
            }
            for (int i = 0; i < nu; ++i)
                if (units[i] & (UNIT_M | UNIT_M1))
                    LDSR_CHECK(ubase[i] >= 0 && ubase[i] + ((units[i] & UNIT_M) ? MSEG : 1) <= nm);
            for (int i = 0; i < nm; ++i) LDSR_CHECK(mlist[i] >= 0 && mlist[i] < T);
        }
    }
    mbar_wait(&bar, phase);
    __syncthreads();
    const int n_msteps = pbound[NP + 2];
    const int ma = (int)(((long long)n_msteps * warp) / NW), mb = (int)(((long long)n_msteps * (warp + 1)) / NW);
    const int sa = sbound[warp], sb = sbound[warp + 1];
    const int ua = pbound[warp], ue = pbound[warp + 1];

#ifdef LDSR_PHASE_CLOCKS
    long long pc[21] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
    for (int it = 0; it < P.chunk; ++it) {
        if (!__any_sync(FULL, live)) break;
        SplitConst<PQ, UW> k;
        k.set(th);

        // ================= phase A: Bu_t = B.u_t over my slice, Dv_t = D.v_t of the observed units =================
        constexpr int FPL = wide_fpl(PQ);
        double *const TRb = TR - lane, *const YMb = YM - lane; // un-offset bases: here a lane is not one fit
        rows_times_coef<PQ, FPL>(TB - lane, us + (size_t)sa * PQ, nullptr, TRb + (size_t)sa * 32, sb - sa, lane);
        rows_times_coef<PQ, FPL>(TD - lane, vs, mlist + ma, YMb + (size_t)ma * 32, mb - ma, lane); // my share of the Dv rows

        LDSR_PHASE_MARK(16);
        // ================= P1: variance map of my piece =================
        if (warp == NW - 1) {
            uvar_constants<UW>(k.A2, k.Q, k.aVW, UV); // nothing is to the right of the last piece
        } else {
            double m11 = 1.0, m12 = 0.0, m21 = 0.0, m22 = 1.0;
            for (int un = ua; un < ue; ++un) {
                const int u0 = units[un];
                const int t0 = u0 & UNIT_T0;
                if (u0 & UNIT_M) {
                    compose_var_unit<PQ, UW, MSEG>(th, k, mask_bits(t0, MSEG), m11, m12, m21, m22);
                } else if (u0 & UNIT_M1) {
                    compose_var_unit<PQ, UW, 1>(th, k, mask_bits(t0, 1), m11, m12, m21, m22);
                } else if (u0 & UNIT_US) {
                    const ShortConst<MSEG> sc(k.A, k.A2, k.Q);
                    m11 = fma(sc.aV, m11, sc.bV * m21);
                    m12 = fma(sc.aV, m12, sc.bV * m22);
                } else {
                    m11 = fma(k.aVW, m11, k.bVW * m21);
                    m12 = fma(k.aVW, m12, k.bVW * m22);
                }
            }
            MC[(warp * 4 + 0) * 32] = m11;
            MC[(warp * 4 + 1) * 32] = m12;
            MC[(warp * 4 + 2) * 32] = m21;
            MC[(warp * 4 + 3) * 32] = m22;
        }
        LDSR_PHASE_MARK(0);
        __syncthreads(); // B1: Bu, Dv, the maps and the variance-sum constants are in place
        LDSR_PHASE_MARK(1);
        double Vin = th.V1; // prior variance entering my piece
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
                if (pj + 1 == warp) Vin = n * fast_rcp(d);
            }
        }

        // ================= P2: forward over my piece =================
        {
            PieceFwd c;
            c.Vq = Vin;
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
                CK[(un * 3 + 0) * 32] = c.Vq;
                CK[(un * 3 + 1) * 32] = c.q;
                CK[(un * 3 + 2) * 32] = c.P;
                if (u0 & UNIT_M) {
                    any_m = true;
                    wide_forward_unit<PQ, UW, MSEG>(th, k, mask_bits(t0, MSEG), ys + t0, TR + t0 * 32,
                                                    YM + (size_t)ubase[un] * 32, c);
                } else if (u0 & UNIT_M1) {
                    any_m = true;
                    wide_forward_unit<PQ, UW, 1>(th, k, mask_bits(t0, 1), ys + t0, TR + t0 * 32,
                                                 YM + (size_t)ubase[un] * 32, c);
                } else if (u0 & UNIT_US) {
                    wide_forward_short<MSEG>(k.A, k.A2, k.Q, TR + t0 * 32, c);
                } else {
                    wide_forward_word<PQ, UW>(k, TR + t0 * 32, c);
                }
            }
            double ld = 0.0;
            if (any_m) ld = fma((double)c.shift, 0.693147180559945309417, log(c.dprod));
            double *o = CH + (size_t)warp * WIDE_NCH * 32;
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
        __syncthreads(); // B2
        LDSR_PHASE_MARK(3);

        // ---- chain the pieces: the prior mean entering every piece, the likelihood, the smoothed state entering
        //      every piece from the right.  ONE warp walks the two chains and publishes what each piece needs in the
        //      piece summaries themselves (their P, q, l0, l1, l2 are dead once read): the chains are serial and short,
        //      and eight warps walking them redundantly were 8 x 80 shared-memory loads of 256 B through one port,
        //      1 400 - 2 900 cycles of an iteration (profiles/em_r02_wide_v4_balance_and_clocks.txt).
        if (warp == 0) {
            double gk[NP];
            double acc = 0.0;
            double x = th.mu1; // prior of step 0 (EM.cpp:48)
            double vend = th.V1;
#pragma unroll
            for (int pj = 0; pj < NP; ++pj) {
                double *o = CH + (size_t)pj * WIDE_NCH * 32;
                const double x_in = x;
                gk[pj] = fma(o[7 * 32], x, o[6 * 32]);
                const double tC = th.C * x;
                acc += fma(tC, fma(tC, o[4 * 32], -2.0 * o[3 * 32]), o[2 * 32]);
                x = fma(o[0 * 32], x, o[1 * 32]);
                vend = o[9 * 32];
                o[0 * 32] = x_in; // prior mean entering the piece
                o[1 * 32] = x;    // prior mean of the first step right of it
            }
            // EM.cpp:122-124 (x 1/n, rounded once per task)
            CH[4 * 32] = (-0.5 * n_obs * LOG_2PI - 0.5 * acc) * inv_n_obs;
            // the backward chain starts from the prior of the virtual step T: Xs_{T-1} = Xu + J (Xp_T - Xp_T) = Xu_{T-1}
            // (EM.cpp:94-95)
            double Xs = x, Vs = vend;
#pragma unroll
            for (int pj = NP - 1; pj >= 0; --pj) {
                double *o = CH + (size_t)pj * WIDE_NCH * 32;
                o[2 * 32] = Xs; // smoothed state of the first step right of the piece
                o[3 * 32] = Vs;
                const double pjv = o[5 * 32];
                Xs = fma(pjv, Xs, gk[pj]);
                Vs = fma(pjv * pjv, Vs, o[8 * 32]);
            }
        }
        __syncthreads(); // B2b
        const double lik_new = CH[4 * 32];
        double xin, Xr, Vr, Xs1, Vs1; // prior mean entering my piece; prior right of it; smoothed state right of it
        {
            const double *o = CH + (size_t)warp * WIDE_NCH * 32;
            xin = o[0 * 32];
            Xr = o[1 * 32];
            Vr = o[9 * 32];
            Xs1 = o[2 * 32];
            Vs1 = o[3 * 32];
        }

        // ================= stop rule (EM.cpp:259-275) =================
        if (live) {
            lik = lik_new;
            ne += 1;
            if (warp == 0 && P.liks) P.liks[(size_t)P.f_user[fit] * P.niter + (ne - 1)] = lik_new;
            const bool conv = (ne >= 3) && (fabs(lik_new - l1) < P.tol) && (fabs(l1 - l2) < P.tol);
            if (conv || ne >= P.niter) live = false;
        }
        if (!__any_sync(FULL, live)) break;

        LDSR_PHASE_MARK(17);
        // ================= P4: backward over my piece: scalar sums, Xs_t -> shared memory =================
        WideSums ws;
        ws.zero();
        {
            double cG = 0.0, cH = 0.0;
            bool in_run = false;
            for (int un = ue - 1; un >= ua; --un) {
                const int u0 = units[un];
                const int t0 = u0 & UNIT_T0;
                const double Vq = CK[(un * 3 + 0) * 32];
                const double Xq = fma(CK[(un * 3 + 2) * 32], xin, CK[(un * 3 + 1) * 32]);
                if (u0 & UNIT_M) {
                    wide_smooth_unit<PQ, UW, MSEG>(th, k, mask_bits(t0, MSEG), false, ys + t0, TR + t0 * 32,
                                                   YM + (size_t)ubase[un] * 32, Xq, Vq, Xs1, Vs1, ws);
                    in_run = false;
                } else if (u0 & UNIT_M1) {
                    wide_smooth_unit<PQ, UW, 1>(th, k, mask_bits(t0, 1), t0 == T - 1, ys + t0, TR + t0 * 32,
                                                YM + (size_t)ubase[un] * 32, Xq, Vq, Xs1, Vs1, ws);
                    in_run = false;
                } else {
                    if (!in_run) {
                        const double rv = fast_rcp(Vr);
                        cG = (Xs1 - Xr) * rv;
                        cH = (Vs1 - Vr) * rv * rv;
                        in_run = true;
                    }
                    if (u0 & UNIT_US)
                        wide_smooth_short<MSEG>(k.A, k.A2, k.Q, TR + t0 * 32, Xq, Vq, cG, cH, Xs1, Vs1, ws);
                    else
                        wide_smooth_word<PQ, UW>(k, TR + t0 * 32, UV, Xq, Vq, cG, cH, Xs1, Vs1, ws);
                }
                Xr = Xq;
                Vr = Vq;
            }
            if (ua == 0 && ue > 0) {
                ws.X0 = Xs1;
                ws.V0 = Vs1;
            }
        }
        LDSR_PHASE_MARK(4);
        __syncthreads(); // B3: every Xs_t is in place
        LDSR_PHASE_MARK(5);

        // ================= phase C: the row sums of EM.cpp:153-161, 184-193 =================
        Stats<PQ> st;
        st.zero();
        {
            constexpr int LG = 32 / FPL;
            const int g = lane / LG, lg = lane % LG;
            { // transitions t = 0 .. T-2 of my slice: Tux += u_t Xs_t, Tx1u += u_t Xs_{t+1}
                const int te = sb < T - 1 ? sb : T - 1;
                const int n = te > sa ? te - sa : 0, len = (n + FPL - 1) / FPL;
                const int lo = sa + g * len, hi = (lo + len < sa + n) ? lo + len : sa + n;
                double acc[FPL][2 * PQ];
                double z[FPL];
#pragma unroll
                for (int k = 0; k < FPL; k++) {
#pragma unroll
                    for (int i = 0; i < 2 * PQ; i++) acc[k][i] = 0.0;
                    z[k] = lo < hi ? TRb[(size_t)lo * 32 + lg + k * LG] : 0.0;
                }
                int t = lo;
#pragma unroll 1
                for (; t + 2 <= hi; t += 2) { // two steps per trip: the second row's loads behind the first row's math
                    double r0[PQ], r1[PQ];
                    load_vec<PQ>(us + (size_t)t * PQ, r0);
                    load_vec<PQ>(us + (size_t)(t + 1) * PQ, r1);
#pragma unroll
                    for (int k = 0; k < FPL; k++) {
                        const double z1 = TRb[(size_t)(t + 1) * 32 + lg + k * LG];
                        const double z2 = TRb[(size_t)(t + 2) * 32 + lg + k * LG];
#pragma unroll
                        for (int i = 0; i < PQ; i++) {
                            acc[k][i] = fma(r0[i], z[k], acc[k][i]);
                            acc[k][PQ + i] = fma(r0[i], z1, acc[k][PQ + i]);
                        }
#pragma unroll
                        for (int i = 0; i < PQ; i++) {
                            acc[k][i] = fma(r1[i], z1, acc[k][i]);
                            acc[k][PQ + i] = fma(r1[i], z2, acc[k][PQ + i]);
                        }
                        z[k] = z2;
                    }
                }
                if (t < hi) {
                    double r[PQ];
                    load_vec<PQ>(us + (size_t)t * PQ, r);
#pragma unroll
                    for (int k = 0; k < FPL; k++) {
                        const double z1 = TRb[(size_t)(t + 1) * 32 + lg + k * LG];
#pragma unroll
                        for (int i = 0; i < PQ; i++) {
                            acc[k][i] = fma(r[i], z[k], acc[k][i]);
                            acc[k][PQ + i] = fma(r[i], z1, acc[k][PQ + i]);
                        }
                        z[k] = z1;
                    }
                }
                double tot[2 * PQ];
                group_transpose_sum<FPL, 2 * PQ>(acc, tot, lane);
#pragma unroll
                for (int i = 0; i < PQ; i++) {
                    st.Tux[i] = tot[i];
                    st.Tx1u[i] = tot[PQ + i];
                }
            }
            { // my share of the steps inside observed units: Sxv += v_t Xs_t where the fit observes t
                double acc[FPL][PQ];
#pragma unroll
                for (int k = 0; k < FPL; k++)
#pragma unroll
                    for (int i = 0; i < PQ; i++) acc[k][i] = 0.0;
                const unsigned *const MWb = MW - lane;
                const int n = mb - ma, len = (n + FPL - 1) / FPL;
                const int lo = ma + g * len, hi = (lo + len < mb) ? lo + len : mb;
#pragma unroll 1
                for (int m = lo; m < hi; ++m) {
                    const int t = mlist[m];
                    double r[PQ];
                    load_vec<PQ>(vs + (size_t)t * PQ, r);
#pragma unroll
                    for (int k = 0; k < FPL; k++) {
                        const int f = lg + k * LG;
                        const bool obs = (MWb[(t >> 5) * 32 + f] >> (t & 31)) & 1u;
                        const double xo = obs ? TRb[(size_t)t * 32 + f] : 0.0;
#pragma unroll
                        for (int i = 0; i < PQ; i++) acc[k][i] = fma(xo, r[i], acc[k][i]);
                    }
                }
                double tot[PQ];
                group_transpose_sum<FPL, PQ>(acc, tot, lane);
#pragma unroll
                for (int i = 0; i < PQ; i++) st.Sxv[i] = tot[i];
            }
        }
        LDSR_PHASE_MARK(18);
        __syncthreads(); // B4: nobody reads the trajectory any more: it becomes the partial sums
        LDSR_PHASE_MARK(7);
        st.Syx = ws.Syx;
        st.Sxx = ws.Sxx;
        st.Sxxv = ws.Sxxv;
        st.Tx1x = ws.Tx1x;
        st.Tx1xv = ws.Tx1xv;
        st.Txx = ws.Txx;
        st.Txxv = ws.Txxv;
        st.X0 = ws.X0;
        st.V0 = ws.V0;
        st.XT = ws.XT;
        st.VT = ws.VT;
        stats_store<PQ>(st, ST + (size_t)warp * NST * 32);
        LDSR_PHASE_MARK(8);
        __syncthreads(); // B5
        LDSR_PHASE_MARK(9);

        // ================= M-step (EM.cpp:139-229), spread over the warps =================
        // (1) totals: warp w adds the NW partial sums of the entries w, w + NW, ...
        double *const TOT = ST + (size_t)NW * NST * 32; // [NST]
        double *const ZW = TOT + (size_t)NST * 32;      // [3 PQ]: SvvInv Sxv | TuuInv Tux | TuuInv Tx1u
#pragma unroll 2
        for (int j = warp; j < NST; j += NW) {
            double a = ST[j * 32];
#pragma unroll
            for (int w = 1; w < NW; ++w) a += ST[((size_t)w * NST + j) * 32];
            TOT[j * 32] = a;
        }
        LDSR_PHASE_MARK(10);
        __syncthreads(); // B5b
        LDSR_PHASE_MARK(11);
        // (2) the three matrix-vector products of the block elimination (lds_math.cuh), one row per warp at a time
        {
            const double *__restrict__ svv_inv = gc + 2 + 2 * PQ;
#pragma unroll 2 // two rows in flight (measured: 1 -> 2: config 3 1.128 -> 1.117 s; 4: 1.125 s)
            for (int rr = warp; rr < 3 * PQ; rr += NW) {
                const int which = rr / PQ, a = rr - which * PQ;
                const double *__restrict__ m = (which == 0 ? svv_inv : tuu_inv) + a * PQ;
                // the vector, in stats_store order: Sxv at 11, Tx1u at 11 + PQ, Tux at 11 + 2 PQ
                const double *__restrict__ xv = TOT + (size_t)(which == 0 ? 11 : (which == 1 ? 11 + 2 * PQ : 11 + PQ)) * 32;
                double a0 = m[0] * xv[0], a1 = m[1] * xv[32];
#pragma unroll
                for (int b2 = 2; b2 < PQ; b2++) {
                    if (b2 & 1)
                        a1 = fma(m[b2], xv[b2 * 32], a1);
                    else
                        a0 = fma(m[b2], xv[b2 * 32], a0);
                }
                ZW[rr * 32] = a0 + a1;
            }
        }
        LDSR_PHASE_MARK(12);
        __syncthreads(); // B5c
        LDSR_PHASE_MARK(13);
        // (3) the scalars: warp 0 the observation block (C, D, R), warp 1 the transition block (A, B, Q, mu1, V1)
        if (warp == 0) {
            if (live) {
                const double Syy = gc[0];
                const double *Syv = gc + 2, *wy = gc + 2 + PQ;
                const double Syx = TOT[0 * 32], Sxx = TOT[1 * 32] + TOT[2 * 32]; // EM.cpp:152
                double num = Syx, den = Sxx;
#pragma unroll
                for (int a = 0; a < PQ; a++) {
                    const double sxv = TOT[(11 + a) * 32];
                    num = fma(-wy[a], sxv, num);
                    den = fma(-sxv, ZW[a * 32], den);
                }
                const double Cn = mstep_div<true>(num, den); // reciprocal multiplies: a serial section
                double racc = fma(-Cn, Syx, Syy);
#pragma unroll
                for (int a = 0; a < PQ; a++) {
                    const double d = fma(-Cn, ZW[a * 32], wy[a]);
                    TD[a * 32] = d;
                    racc = fma(-d, Syv[a], racc);
                }
                th.C = Cn;
                th.R = racc * inv_n_obs;
            }
            TH[1 * 32] = th.C;
            TH[3 * 32] = th.R;
        } else if (warp == 1) {
            if (live) {
                const double Txx = TOT[5 * 32] + TOT[6 * 32], Tx1x = TOT[3 * 32] + TOT[4 * 32]; // EM.cpp:180,181
                const double X0 = TOT[7 * 32], V0 = TOT[8 * 32], XT = TOT[9 * 32], VT = TOT[10 * 32];
                double num = Tx1x, den = Txx;
#pragma unroll
                for (int a = 0; a < PQ; a++) {
                    const double z = ZW[(PQ + a) * 32];
                    num = fma(-TOT[(11 + PQ + a) * 32], z, num);     // Tx1u . z
                    den = fma(-TOT[(11 + 2 * PQ + a) * 32], z, den); // Tux . z
                }
                const double An = mstep_div<true>(num, den);
                // Tx1x1 = sum_{t=1}^{T-1} (X_t^2+V_t) = Txx - (X_0^2+V_0) + (X_{T-1}^2+V_{T-1})   (EM.cpp:181,183)
                const double Tx1x1 = Txx - fma(X0, X0, V0) + fma(XT, XT, VT);
                double qacc = fma(-An, Tx1x, Tx1x1);
#pragma unroll
                for (int a = 0; a < PQ; a++) {
                    const double bb = fma(-An, ZW[(PQ + a) * 32], ZW[(2 * PQ + a) * 32]);
                    TB[a * 32] = bb;
                    qacc = fma(-bb, TOT[(11 + PQ + a) * 32], qacc);
                }
                th.A = An;
                th.Q = mstep_div<true>(qacc, (double)(T - 1));
                th.mu1 = X0; // EM.cpp:218-219
                th.V1 = V0;
            }
            TH[0 * 32] = th.A;
            TH[2 * 32] = th.Q;
            TH[4 * 32] = th.mu1;
            TH[5 * 32] = th.V1;
        }
        LDSR_PHASE_MARK(14);
        __syncthreads(); // B6: the new theta is published
        LDSR_PHASE_MARK(15);
        th.A = TH[0 * 32];
        th.C = TH[1 * 32];
        th.Q = TH[2 * 32];
        th.R = TH[3 * 32];
        th.mu1 = TH[4 * 32];
        th.V1 = TH[5 * 32];
        if (live) {
            l2 = l1;
            l1 = lik;
        }
    }

#ifdef LDSR_PHASE_CLOCKS
    if (lane == 0 && WP.clk)
        for (int i = 0; i < 21; i++) WP.clk[((size_t)blockIdx.x * NW + warp) * 21 + i] = pc[i];
#endif
    if (warp == 0 && valid) {
        double *g = P.theta + (size_t)fit * TL;
        g[0] = th.A;
        g[1 + PQ] = th.C;
        g[2 + 2 * PQ] = th.Q;
        g[3 + 2 * PQ] = th.R;
        g[4 + 2 * PQ] = th.mu1;
        g[5 + 2 * PQ] = th.V1;
#pragma unroll
        for (int i = 0; i < PQ; i++) {
            g[1 + i] = TB[i * 32];
            g[2 + PQ + i] = TD[i * 32];
        }
        P.l1[fit] = l1;
        P.l2[fit] = l2;
        P.lik[fit] = lik;
        P.ne[fit] = ne;
        P.done[fit] = live ? 0 : 1;
    }
    } // task loop
}

} // namespace ldsr
