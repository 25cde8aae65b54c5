// lds_math.cuh -- per-fit (one lane = one fit) scalar-state Kalman / RTS / M-step arithmetic.
//
// What is computed follows the reference (src/EM.cpp:22-131 E-step, :139-229 M-step) term for
// term; HOW it is computed is organised for the GPU: theta and the sufficient statistics live
// in registers (PQ is a compile-time width), u/v/y are read through a warp-uniform pointer
// (shared memory after a TMA stage, or global), and the two (p+1)^2 / (q+1)^2 inverses of the
// reference are replaced by block elimination against the theta-independent Gram blocks
// (Tuu = sum u u', Svv = sum_obs v v') whose inverses are precomputed once per series / group.
#pragma once
#include "common.cuh"

namespace ldsr {

// Wide inputs: the time-split kernel keeps B and D in shared memory during the unit loops (one copy
// per CTA, [i][lane]) instead of 2 PQ registers per lane -- at PQ = 10 the kernel needs well over
// 255 registers and spills (profiles/em_r01_pq10.txt); only that kernel sets sb / sd and calls b() / d().
__host__ __device__ constexpr bool theta_bd_in_smem(int pq) { return pq >= 8; }
template <int PQ> struct Theta {
    double A, C, Q, R, mu1, V1;
    double B[PQ], D[PQ];
    const double *sb, *sd;
    __device__ __forceinline__ double b(int i) const {
        if constexpr (theta_bd_in_smem(PQ)) return sb[i * 32];
        else return B[i];
    }
    __device__ __forceinline__ double d(int i) const {
        if constexpr (theta_bd_in_smem(PQ)) return sd[i * 32];
        else return D[i];
    }
};

// device-side flat layout of a padded theta: [A, B[PQ], C, D[PQ], Q, R, mu1, V1]
template <int PQ> __host__ __device__ constexpr int theta_pad_len() { return 2 * PQ + 6; }

template <int PQ> __device__ __forceinline__ void load_theta(Theta<PQ> &t, const double *g) {
    t.A = g[0];
#pragma unroll
    for (int j = 0; j < PQ; j++) t.B[j] = g[1 + j];
    t.C = g[1 + PQ];
#pragma unroll
    for (int j = 0; j < PQ; j++) t.D[j] = g[2 + PQ + j];
    t.Q = g[2 + 2 * PQ];
    t.R = g[3 + 2 * PQ];
    t.mu1 = g[4 + 2 * PQ];
    t.V1 = g[5 + 2 * PQ];
}
template <int PQ> __device__ __forceinline__ void store_theta(const Theta<PQ> &t, double *g) {
    g[0] = t.A;
#pragma unroll
    for (int j = 0; j < PQ; j++) g[1 + j] = t.B[j];
    g[1 + PQ] = t.C;
#pragma unroll
    for (int j = 0; j < PQ; j++) g[2 + PQ + j] = t.D[j];
    g[2 + 2 * PQ] = t.Q;
    g[3 + 2 * PQ] = t.R;
    g[4 + 2 * PQ] = t.mu1;
    g[5 + 2 * PQ] = t.V1;
}

// Reciprocal for the filter/smoother gains: MUFU.RCP64H seed + two Newton steps (~1 ulp, not
// correctly rounded; the parity budget is 1e-9).  ~3x the throughput of the IEEE division
// sequence (measured: 18 vs 125 issue cycles per warp, profiles/microbench_r01.txt).
// Arguments are variances (Vp, Sigma): positive and far from the subnormal range in any fit that
// still has a finite likelihood.
__device__ __forceinline__ double fast_rcp(double x) {
    double r;
#ifndef LDSR_HOST_SIM
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#else
    r = 1.0 / x; // CPU emulation (tests/host_simt): the Newton steps below leave it unchanged
#endif
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}

// Natural logarithm of a POSITIVE NORMAL double (no zero / subnormal / infinity / NaN handling), 2 ulp, straight-line
// code.  x = 2^e m with m in [sqrt(1/2), sqrt(2)); log m = 2 atanh(s), s = (m - 1)/(m + 1), |s| <= 0.172: odd series
// up to s^21.
__device__ __forceinline__ double log_pos(double x) {
    int hi = __double2hiint(x);
    const int lo = __double2loint(x);
    int e = (hi >> 20) - 1023;
    hi = (hi & 0x000fffff) | 0x3ff00000;
    const bool upper = hi >= 0x3ff6a09e; // m > sqrt(2) (to 20 bits): halve it
    hi = upper ? hi - 0x00100000 : hi;
    e = upper ? e + 1 : e;
    const double m = __hiloint2double(hi, lo);
    const double s = (m - 1.0) * fast_rcp(m + 1.0), z = s * s;
    double p = 1.0 / 21.0;
    p = fma(p, z, 1.0 / 19.0);
    p = fma(p, z, 1.0 / 17.0);
    p = fma(p, z, 1.0 / 15.0);
    p = fma(p, z, 1.0 / 13.0);
    p = fma(p, z, 1.0 / 11.0);
    p = fma(p, z, 1.0 / 9.0);
    p = fma(p, z, 1.0 / 7.0);
    p = fma(p, z, 1.0 / 5.0);
    p = fma(p, z, 1.0 / 3.0);
    const double r = fma(2.0 * s * z, p, 2.0 * s);
    return fma((double)e, 0.693147180559945309417, r);
}

template <int PQ> __device__ __forceinline__ double dot_row(const double (&w)[PQ], const double *__restrict__ row) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < PQ; j++) s = fma(w[j], row[j], s);
    return s;
}

// Measurement update at one step (EM.cpp:61-68, 82-89).  `obs` is per lane; the caller only
// enters when at least one lane of the warp observes this step.  Returns the innovation terms
// for the likelihood (EM.cpp:115-122) through dq (delta^2/Sigma) and S (Sigma).
template <int PQ>
__device__ __forceinline__ void measurement_update(const Theta<PQ> &th, bool obs, double y, const double *__restrict__ vrow,
                                                   double Xp, double Vp, double &Xu, double &Vu, double &dq,
                                                   double &S) {
    const double Dv = dot_row<PQ>(th.D, vrow);
    S = fma(th.C * Vp, th.C, th.R);      // C*Vp*C + R
    const double rS = 1.0 / S;
    const double K = Vp * th.C * rS;     // Vp*C*inv(Sigma)
    const double delta = y - fma(th.C, Xp, Dv);
    Xu = obs ? fma(K, delta, Xp) : Xp;
    Vu = obs ? (1.0 - K * th.C) * Vp : Vp;
    dq = delta * rS * delta;
}

// Sufficient statistics of one E-step (EM.cpp:151-161, 180-193), theta-dependent part only.
template <int PQ> struct Stats {
    double Syx, Sxx, Sxxv;          // over observed steps: sum yX, sum X^2, sum V
    double Tx1x, Tx1xv, Txx, Txxv;  // over transitions t=0..T-2: sum X1 X, sum V1 J, sum X^2, sum V
    double Sxv[PQ], Tx1u[PQ], Tux[PQ];
    double X0, V0, XT, VT; // smoothed state at t=0 and t=T-1
    __device__ __forceinline__ void zero() {
        Syx = Sxx = Sxxv = Tx1x = Tx1xv = Txx = Txxv = 0.0;
#pragma unroll
        for (int j = 0; j < PQ; j++) Sxv[j] = Tx1u[j] = Tux[j] = 0.0;
        X0 = V0 = XT = VT = 0.0;
    }
};

// M-step from the statistics (EM.cpp:164-177, 196-214) by block elimination:
//   [A B] = [Tx1x Tx1u] inv([[Txx,Txu],[Tux,Tuu]])
//     z = TuuInv Tux, w = TuuInv Tx1u, A = (Tx1x - Tx1u.z)/(Txx - Tux.z), B = w - A z
//   [C D] = [Syx Syv] inv([[Sxx,Sxv],[Svx,Svv]])
//     z = SvvInv Sxv, C = (Syx - wy.Sxv)/(Sxx - Sxv.z), D = wy - C z,  wy = SvvInv Syv
//   R = sum_obs (y - yhat) y / n_obs = (Syy - C Syx - D.Syv)/n_obs         (EM.cpp:177)
//   Q = (Tx1x1 - A Tx1x - B.Tx1u)/(T-1)                                     (EM.cpp:210)
// With TuuInv = 0 (no u) this reduces to A = Tx1x/Txx, B = 0 (EM.cpp:212-213); likewise C, D.
// gc -> [Syy, n_obs, Syv[PQ], wy[PQ], SvvInv[PQ*PQ]] (per group), tuu_inv -> [PQ*PQ] (per series).
// a / b.  FAST: a * fast_rcp(b) (about 1 ulp; for the latency path of the scan kernel, where the M-step is a
// serial section of one thread and an IEEE division costs ~130 cycles)
template <bool FAST> __device__ __forceinline__ double mstep_div(double a, double b) {
    if constexpr (FAST)
        return a * fast_rcp(b);
    else
        return a / b;
}

// The observation block: C, D, R (EM.cpp:164-177).
template <int PQ, bool FAST = false>
__device__ __forceinline__ void mstep_obs_block(const Stats<PQ> &s, const double *__restrict__ gc, Theta<PQ> &th) {
    const double Syy = gc[0], n_obs = gc[1];
    const double *Syv = gc + 2, *wy = gc + 2 + PQ, *svv_inv = gc + 2 + 2 * PQ;
    double z[PQ];
    const double Sxx = s.Sxx + s.Sxxv; // EM.cpp:152
    double num = s.Syx, den = Sxx;
#pragma unroll
    for (int a = 0; a < PQ; a++) {
        double acc = 0.0;
#pragma unroll
        for (int b = 0; b < PQ; b++) acc = fma(svv_inv[a * PQ + b], s.Sxv[b], acc);
        z[a] = acc;
    }
#pragma unroll
    for (int a = 0; a < PQ; a++) {
        num = fma(-wy[a], s.Sxv[a], num);
        den = fma(-s.Sxv[a], z[a], den);
    }
    const double Cn = mstep_div<FAST>(num, den);
    double racc = fma(-Cn, s.Syx, Syy);
#pragma unroll
    for (int a = 0; a < PQ; a++) {
        const double d = fma(-Cn, z[a], wy[a]);
        th.D[a] = d;
        racc = fma(-d, Syv[a], racc);
    }
    th.C = Cn;
    th.R = mstep_div<FAST>(racc, n_obs);
}
// The transition block: A, B, Q, mu1, V1 (EM.cpp:180-219).
template <int PQ, bool FAST = false>
__device__ __forceinline__ void mstep_trans_block(const Stats<PQ> &s, const double *__restrict__ tuu_inv, int T,
                                                  Theta<PQ> &th) {
    const double Txx = s.Txx + s.Txxv, Tx1x = s.Tx1x + s.Tx1xv; // EM.cpp:180,181
    double z[PQ], w[PQ];
    double num = Tx1x, den = Txx;
#pragma unroll
    for (int a = 0; a < PQ; a++) {
        double acc = 0.0, acw = 0.0;
#pragma unroll
        for (int b = 0; b < PQ; b++) {
            const double m = tuu_inv[a * PQ + b];
            acc = fma(m, s.Tux[b], acc);
            acw = fma(m, s.Tx1u[b], acw);
        }
        z[a] = acc;
        w[a] = acw;
    }
#pragma unroll
    for (int a = 0; a < PQ; a++) {
        num = fma(-s.Tx1u[a], z[a], num);
        den = fma(-s.Tux[a], z[a], den);
    }
    const double An = mstep_div<FAST>(num, den);
    // Tx1x1 = sum_{t=1}^{T-1} (X_t^2+V_t) = Txx - (X_0^2+V_0) + (X_{T-1}^2+V_{T-1})   (EM.cpp:181,183)
    const double Tx1x1 = Txx - fma(s.X0, s.X0, s.V0) + fma(s.XT, s.XT, s.VT);
    double qacc = fma(-An, Tx1x, Tx1x1);
#pragma unroll
    for (int a = 0; a < PQ; a++) {
        const double b = fma(-An, z[a], w[a]);
        th.B[a] = b;
        qacc = fma(-b, s.Tx1u[a], qacc);
    }
    th.A = An;
    th.Q = mstep_div<FAST>(qacc, (double)(T - 1));
    th.mu1 = s.X0; // EM.cpp:218-219
    th.V1 = s.V0;
}
template <int PQ, bool FAST = false>
__device__ __forceinline__ void mstep_from_stats(const Stats<PQ> &s, const double *__restrict__ gc,
                                                 const double *__restrict__ tuu_inv, int T, Theta<PQ> &th) {
    mstep_obs_block<PQ, FAST>(s, gc, th);
    mstep_trans_block<PQ, FAST>(s, tuu_inv, T, th);
}

} // namespace ldsr
