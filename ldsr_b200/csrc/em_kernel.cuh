// em_kernel.cuh -- the LDS_EM loop (src/EM.cpp:245-280) as a batched kernel: ONE LANE = ONE FIT.
//
// A CTA takes up to 32*W live fits of one series.  The series blob (y,u,v) is staged into shared
// memory once per launch by a TMA bulk copy and read by every fit through warp-uniform
// (broadcast) loads.  Each lane runs, per EM iteration,
//   pass 1  forward Kalman filter: log-likelihood + a checkpoint (Xp,Vp) every SEG steps;
//   stop rule (EM.cpp:272) -- a fit that stops keeps the theta this E-step ran with;
//   pass 2  RTS smoother, segment by segment from the end: the segment's filter states are
//           recomputed from its checkpoint into REGISTERS (Xu,Vu,Xp1,Vp1 for SEG steps), then the
//           backward recursion runs over them and accumulates the M-step sums on the fly.  The
//           recursion is written in affine form Xs = J*Xs1 + g, Vs = J^2*Vs1 + L with J, g, L
//           computed off the dependency chain, so one warp alone keeps its FP64 pipe busy (the
//           segment bodies are straight-line code: no votes, no bounds checks);
//   M-step  closed form from the sums (lds_math.cuh).
// No O(T) trajectory ever goes to memory: per fit and iteration the kernel touches 2 doubles per
// SEG steps of checkpoint.  Steps no lane of the warp observes (warp vote on the mask bits) take a
// short path without the measurement update.
// A launch runs at most `chunk` iterations; live fits are re-compacted between launches
// (compact_kernel) so finished fits do not hold lanes.
#pragma once
#include "lds_math.cuh"

namespace ldsr {

struct EmParams {
    const SeriesDev *series;
    const double *blobs;
    const double *sconst;     // per series TuuInv
    const int *g_series;
    const unsigned *masks;    // per group observed-bit words
    const long long *g_mask_off;
    const double *gconst;     // per group constants, stride gconst_stride(PQ)
    const int *g_status;      // per group LDSR_FIT_SINGULAR or 0
    const int *f_group;
    double *theta;            // [n_fits][2PQ+6] current theta (in/out)
    double *l1, *l2;          // previous two likelihoods (in/out)
    double *lik;              // last likelihood (out)
    int *ne;                  // E-steps done (in/out)
    int *done;                // 0 live, 1 finished (in/out)
    double *liks;             // optional [n_fits][niter] trace (in user fit order via f_user)
    const int *f_user;        // internal -> user fit index (for liks rows)
    const int *active;        // compacted live fit ids, grouped by series
    const int4 *tasks;        // per CTA: x = series, y = first index into active, z = count
    double *ckpt;             // checkpoint scratch: [global warp][seg][2][32]
    int max_seg;              // segments per warp slot in ckpt
    int niter, chunk;
    double tol;
    int blob_in_smem;         // 1: stage with TMA; 0: series too large, read it from global
};

__device__ __forceinline__ unsigned seg_bits(const unsigned *__restrict__ mw, int t0, int seg_len) {
    // mask bits of steps t0 .. t0+seg_len-1 (seg_len divides 32, so they sit in one word)
    const unsigned w = mw[t0 >> 5];
    return (w >> (t0 & 31)) & ((seg_len == 32) ? 0xffffffffu : ((1u << seg_len) - 1u));
}

// ---- pass 1 over one segment -----------------------------------------------------------------
// All lanes unobserved: pure prediction (EM.cpp:72-76 with Xu=Xp, Vu=Vp).
template <int PQ, int SEG>
__device__ __forceinline__ void fwd_unobserved(const Theta<PQ> &th, double A, double A2, double Q,
                                               const double *__restrict__ useg, double &Xp, double &Vp) {
    double Bu[SEG];
#pragma unroll
    for (int j = 0; j < SEG; j++) Bu[j] = dot_row<PQ>(th.B, useg + j * PQ);
#pragma unroll
    for (int j = 0; j < SEG; j++) {
        Xp = fma(A, Xp, Bu[j]);
        Vp = fma(A2, Vp, Q);
    }
}

// Some lane observes some step: measurement update on every step, per-lane predicated, plus the
// likelihood terms (EM.cpp:86-88, 115-122).  cnt (warp-uniform) <= SEG steps are valid.
template <int PQ, int SEG>
__device__ __forceinline__ void fwd_mixed(const Theta<PQ> &th, double A, double A2, double Q, unsigned bits, int cnt,
                                          const double *__restrict__ yseg, const double *__restrict__ useg,
                                          const double *__restrict__ vseg, double &Xp, double &Vp, double &acc) {
#pragma unroll
    for (int j = 0; j < SEG; j++) {
        if (j < cnt) {
            const bool obs = (bits >> j) & 1u;
            const double Bu = dot_row<PQ>(th.B, useg + j * PQ);
            const double Dv = dot_row<PQ>(th.D, vseg + j * PQ);
            const double S = fma(th.C * Vp, th.C, th.R); // C*Vp*C + R
            const double rS = fast_rcp(S);
            const double K = Vp * th.C * rS;
            const double delta = yseg[j] - fma(th.C, Xp, Dv);
            const double Xu = obs ? fma(K, delta, Xp) : Xp;
            const double Vu = obs ? (1.0 - K * th.C) * Vp : Vp;
            const double term = delta * rS * delta + log(S);
            acc += obs ? term : 0.0;
            Xp = fma(A, Xu, Bu); // u lags one step (EM.cpp:74)
            Vp = fma(A2, Vu, Q); // A*Vu*A + Q      (EM.cpp:76)
        }
    }
}

// ---- pass 2 over one segment -----------------------------------------------------------------
// MIXED   : the segment holds observed steps for some lane -> measurement updates + observed sums
// GUARDED : the segment is the last one: only cnt steps are valid and its last step is T-1
template <int PQ, int SEG, bool MIXED, bool GUARDED>
__device__ __forceinline__ void smooth_segment(const Theta<PQ> &th, double A, double A2, double Q, unsigned bits,
                                               int cnt, const double *__restrict__ yseg,
                                               const double *__restrict__ useg, const double *__restrict__ vseg,
                                               double Xq, double Vq, double &Xs1, double &Vs1, Stats<PQ> &st) {
    // Recompute the filter over the segment (EM.cpp:70-90) and, in the same sweep, the RTS gain and
    // the affine form of the backward recursion (EM.cpp:100-102) -- all of it off the dependency
    // chain of the smoother:
    //      Xs_t = Xu + J (Xs1 - Xp1)   = J Xs1 + g,    g = Xu - J Xp1
    //      Vs_t = Vu + J (Vs1 - Vp1) J = J^2 Vs1 + L,  L = Vu - J^2 Vp1
    // Only J, g, L (3*SEG doubles) stay in registers.
    double Jt[SEG], g[SEG], L[SEG];
#pragma unroll
    for (int j = 0; j < SEG; j++) {
        if (!GUARDED || j < cnt) {
            double xu = Xq, vu = Vq;
            if (MIXED) {
                const bool obs = (bits >> j) & 1u;
                const double Dv = dot_row<PQ>(th.D, vseg + j * PQ);
                const double S = fma(th.C * Vq, th.C, th.R);
                const double K = Vq * th.C * fast_rcp(S);
                const double delta = yseg[j] - fma(th.C, Xq, Dv);
                xu = obs ? fma(K, delta, Xq) : Xq;
                vu = obs ? (1.0 - K * th.C) * Vq : Vq;
            }
            Xq = fma(A, xu, dot_row<PQ>(th.B, useg + j * PQ));
            Vq = fma(A2, vu, Q);
            if (GUARDED && j == cnt - 1) { // t == T-1: smoothed = filtered (EM.cpp:94-95)
                Jt[j] = 0.0;
                g[j] = xu;
                L[j] = vu;
            } else {
                const double J = vu * A * fast_rcp(Vq);
                Jt[j] = J;
                g[j] = fma(-J, Xq, xu);
                L[j] = fma(-J * J, Vq, vu);
            }
        }
    }
    // ---- the chain and the sums of EM.cpp:151-161, 180-193
#pragma unroll
    for (int j = SEG - 1; j >= 0; j--) {
        if (!GUARDED || j < cnt) {
            const double J = Jt[j];
            const double Xs = fma(J, Xs1, g[j]);
            const double Vs = fma(J * J, Vs1, L[j]);
            if (GUARDED && j == cnt - 1) {
                st.XT = Xs;
                st.VT = Vs;
            } else {
                st.Tx1x = fma(Xs1, Xs, st.Tx1x);
                st.Tx1xv = fma(Vs1, J, st.Tx1xv);
                st.Txx = fma(Xs, Xs, st.Txx);
                st.Txxv += Vs;
#pragma unroll
                for (int k = 0; k < PQ; k++) {
                    const double uk = useg[j * PQ + k];
                    st.Tx1u[k] = fma(Xs1, uk, st.Tx1u[k]);
                    st.Tux[k] = fma(uk, Xs, st.Tux[k]);
                }
            }
            if (MIXED) {
                const bool obs = (bits >> j) & 1u;
                const double xo = obs ? Xs : 0.0, yo = obs ? yseg[j] : 0.0; // y is NaN where missing
                st.Syx = fma(yo, xo, st.Syx);
                st.Sxx = fma(xo, xo, st.Sxx);
                st.Sxxv += obs ? Vs : 0.0;
#pragma unroll
                for (int k = 0; k < PQ; k++) st.Sxv[k] = fma(xo, vseg[j * PQ + k], st.Sxv[k]);
            }
            Xs1 = Xs;
            Vs1 = Vs;
        }
    }
}

template <int PQ, int SEG, int W, bool STAGED>
__global__ void __launch_bounds__(W * 32, 1) em_chunk_kernel(const EmParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;

    const int4 task = P.tasks[blockIdx.x];
    const SeriesDev S = P.series[task.x];
    const int T = S.T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (STAGED) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) stage_blob(smem_raw, P.blobs + S.blob_off, (unsigned)S.blob_doubles * 8u, &bar);
    }
    const double *__restrict__ ser =
        STAGED ? reinterpret_cast<const double *>(smem_raw) : (P.blobs + S.blob_off);
    const double *__restrict__ ys = ser + S.y_off;
    const double *__restrict__ us = ser + S.u_off;
    const double *__restrict__ vs = ser + S.v_off;

    // ---- per-lane fit state (loaded while the TMA copy is in flight)
    const int slot = warp * 32 + lane;
    const bool valid = slot < task.z;
    const int fit = P.active[task.y + (valid ? slot : 0)];
    const int grp = P.f_group[fit];
    const unsigned *__restrict__ mw = P.masks + P.g_mask_off[grp];
    const double *__restrict__ gc = P.gconst + (size_t)grp * gconst_stride(PQ);
    const double *__restrict__ tuu_inv = P.sconst + S.sconst_off;
    const double n_obs = gc[1];
    constexpr int TL = theta_pad_len<PQ>();
    Theta<PQ> th;
    load_theta<PQ>(th, P.theta + (size_t)fit * TL);
    double l1 = P.l1[fit], l2 = P.l2[fit], lik = P.lik[fit];
    int ne = P.ne[fit];
    bool live = valid && (P.done[fit] == 0);
    if (live && P.g_status[grp] != 0) { // Gram block not invertible: the reference would throw
        live = false;
        lik = __longlong_as_double(0x7ff8000000000000ULL);
    }
    double *__restrict__ ck = P.ckpt + ((size_t)(blockIdx.x * W + warp) * P.max_seg) * 64 + lane;
    const int nseg = (T + SEG - 1) / SEG;       // the last segment (full or not) holds step T-1
    const int cnt_last = T - (nseg - 1) * SEG;  // 1..SEG valid steps in it

    if (STAGED) mbar_wait(&bar, 0);
    if (warp * 32 >= task.z) return; // whole warp has no fit

    for (int it = 0; it < P.chunk; ++it) {
        if (!__any_sync(FULL, live)) break;
        const double A = th.A, A2 = th.A * th.A, Q = th.Q;

        // ================= pass 1: forward filter, likelihood, checkpoints =================
        double Xp = th.mu1, Vp = th.V1; // prior of step 0 (EM.cpp:48-49)
        double acc = 0.0;               // sum_obs delta^2/Sigma + log Sigma (EM.cpp:122)
        for (int sg = 0; sg < nseg; ++sg) {
            const int t0 = sg * SEG;
            ck[(size_t)sg * 64] = Xp;
            ck[(size_t)sg * 64 + 32] = Vp;
            const unsigned bits = seg_bits(mw, t0, SEG);
            const double *__restrict__ useg = us + t0 * PQ;
            if (sg == nseg - 1 || __any_sync(FULL, bits != 0u))
                fwd_mixed<PQ, SEG>(th, A, A2, Q, bits, sg == nseg - 1 ? cnt_last : SEG, ys + t0, useg, vs + t0 * PQ,
                                   Xp, Vp, acc);
            else
                fwd_unobserved<PQ, SEG>(th, A, A2, Q, useg, Xp, Vp);
        }
        // lik = (-0.5 n log 2pi - 0.5 acc)/n      (EM.cpp:122-124, stdlik = TRUE)
        const double lik_new = (-0.5 * n_obs * LOG_2PI - 0.5 * acc) / n_obs;

        // ================= stop rule (EM.cpp:259-275) =================
        if (live) {
            lik = lik_new;
            ne += 1;
            if (P.liks) P.liks[(size_t)P.f_user[fit] * P.niter + (ne - 1)] = lik_new;
            const bool conv = (ne >= 3) && (fabs(lik_new - l1) < P.tol) && (fabs(l1 - l2) < P.tol);
            if (conv || ne >= P.niter) live = false;
        }
        if (!__any_sync(FULL, live)) break;

        // ================= pass 2: RTS smoother + M-step sums, segment by segment =================
        Stats<PQ> st;
        st.zero();
        double Xs1 = 0.0, Vs1 = 0.0; // smoothed state of step t+1
        {
            const int sg = nseg - 1, t0 = sg * SEG;
            smooth_segment<PQ, SEG, true, true>(th, A, A2, Q, seg_bits(mw, t0, SEG), cnt_last, ys + t0, us + t0 * PQ,
                                                vs + t0 * PQ, ck[(size_t)sg * 64], ck[(size_t)sg * 64 + 32], Xs1, Vs1,
                                                st);
        }
        for (int sg = nseg - 2; sg >= 0; --sg) {
            const int t0 = sg * SEG;
            const double Xq = ck[(size_t)sg * 64], Vq = ck[(size_t)sg * 64 + 32]; // (Xp,Vp) entering the segment
            const unsigned bits = seg_bits(mw, t0, SEG);
            if (__any_sync(FULL, bits != 0u))
                smooth_segment<PQ, SEG, true, false>(th, A, A2, Q, bits, SEG, ys + t0, us + t0 * PQ, vs + t0 * PQ, Xq,
                                                     Vq, Xs1, Vs1, st);
            else
                smooth_segment<PQ, SEG, false, false>(th, A, A2, Q, bits, SEG, ys + t0, us + t0 * PQ, vs + t0 * PQ, Xq,
                                                      Vq, Xs1, Vs1, st);
        }
        st.X0 = Xs1;
        st.V0 = Vs1;

        // ================= M-step (EM.cpp:139-229) =================
        if (live) {
            mstep_from_stats<PQ>(st, gc, tuu_inv, T, th);
            l2 = l1;
            l1 = lik;
        }
    }

    if (valid) {
        store_theta<PQ>(th, P.theta + (size_t)fit * TL);
        P.l1[fit] = l1;
        P.l2[fit] = l2;
        P.lik[fit] = lik;
        P.ne[fit] = ne;
        P.done[fit] = live ? 0 : 1;
    }
}

} // namespace ldsr
