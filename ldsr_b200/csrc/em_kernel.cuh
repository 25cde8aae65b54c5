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
// SEG steps of checkpoint.  Segments in which the series has no observation at all take a short path
// without the measurement update.
// A launch runs at most `chunk` iterations; live fits are re-compacted between launches
// (compact_kernel) so finished fits do not hold lanes.
#pragma once
#include "lds_math.cuh"

namespace ldsr {

struct EmParams {
    const SeriesDev *series;
    const double *blobs;
    const double *sconst;     // per series TuuInv
    const double *uwin;       // per series, per window of SPLIT_UW steps: sum u u' (time-split kernel)
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
    const int *n_tasks;       // device-side task count of this launch: CTAs beyond it exit at once
    double *ckpt;             // checkpoint scratch: [global warp][seg][2][32]
    int max_seg;              // segments per warp slot in ckpt
    int ckpt_smem_off;        // MODE 2: byte offset of the checkpoint area in dynamic shared memory
    int niter, chunk;
    double tol;
    int mode;                 // kernel MODE (see em_chunk_kernel)
    // em_scan_kernel<PQ, L, EMIT = true> only: the smoothed trajectories (EM.cpp:94-110 outputs) of n_jobs (group, fit)
    // pairs -- the winners of the groups -- one CTA per job; theta is read from theta[job_theta[job]] (< 0: NaN rows)
    int n_jobs;
    const int *job_group, *job_theta;
    const long long *job_row; // first element of the job's rows in tX / tY / tV / tJ
    double *tX, *tY, *tV, *tJ;
};

__device__ __forceinline__ unsigned seg_bits(const unsigned *__restrict__ mw, int t0, int seg_len) {
    // mask bits of steps t0 .. t0+seg_len-1 (seg_len divides 32, so they sit in one word)
    const unsigned w = mw[t0 >> 5];
    return (w >> (t0 & 31)) & ((seg_len == 32) ? 0xffffffffu : ((1u << seg_len) - 1u));
}

// ---- pass 1 over one segment -----------------------------------------------------------------
// All lanes unobserved: pure prediction (EM.cpp:72-76 with Xu=Xp, Vu=Vp) over SEG steps.  Pass 1
// only needs the state at the end of the segment, so the SEG-step recursion is collapsed:
//   Vp' = A2^SEG Vp + Q (1 + A2 + ... + A2^(SEG-1))            (one FMA, constants per iteration)
//   Xp' = Ah (Ah Xp + h1) + h2,  Ah = A^(SEG/2), h1/h2 = Horner sums of B u over each half
// which leaves two dependent FMAs on Xp instead of SEG.  (Pass 2 recomputes every step from the
// checkpoint with the plain recursion.)
template <int SEG> struct UnobsConst {
    double Ah, aV, bV; // A^(SEG/2), A2^SEG, Q*sum_{k<SEG} A2^k
    __device__ __forceinline__ void set(double A, double A2, double Q) {
        double ah = 1.0, av = 1.0, sv = 0.0;
#pragma unroll
        for (int k = 0; k < SEG / 2; k++) ah *= A;
#pragma unroll
        for (int k = 0; k < SEG; k++) {
            sv = fma(sv, A2, 1.0);
            av *= A2;
        }
        Ah = ah;
        aV = av;
        bV = Q * sv;
    }
};
template <int PQ, int SEG>
__device__ __forceinline__ void fwd_unobserved(const Theta<PQ> &th, double A, const UnobsConst<SEG> &uc,
                                               const double *__restrict__ useg, double &Xp, double &Vp) {
    static_assert(SEG % 2 == 0, "SEG must be even");
    double h1 = 0.0, h2 = 0.0;
#pragma unroll
    for (int j = 0; j < SEG / 2; j++) {
        h1 = fma(A, h1, dot_row<PQ>(th.B, useg + j * PQ));
        h2 = fma(A, h2, dot_row<PQ>(th.B, useg + (SEG / 2 + j) * PQ));
    }
    Xp = fma(uc.Ah, fma(uc.Ah, Xp, h1), h2);
    Vp = fma(uc.aV, Vp, uc.bV);
}

// ---- segments in which some lane observes some step ----------------------------------------
// The variance recursion is run in homogeneous coordinates Vp = n/d (start of segment: n=Vp, d=1):
//   observed   : d' = C^2 n + R d (= Sigma d),  n' = (A^2 R + Q C^2) n + Q R d        (EM.cpp:76,86-88)
//   unobserved : d' = d,                        n' = A^2 n + Q d
// so the dependency chain of a step is two multiply-adds instead of a reciprocal, and everything
// else hangs off it:  1/Sigma = d/d',  K = C n/d',  Vu = nu/d' (nu = R n or n),  Vp' = n'/d',
// J = A Vu/Vp' = A nu/n',  and  sum_obs log Sigma = log(d_end) (telescoping; d is rescaled by exact
// powers of two mid-segment, the shift is added back).  The mean recursion becomes
// Xp' = alpha Xp + beta with alpha = A(1-KC), beta = A K (y - D v) + B u computed off the chain.
template <int PQ> struct MixedConst {
    double C2, a11, a12, AC; // C^2, A^2 R + Q C^2, Q R, A C
    __device__ __forceinline__ void set(const Theta<PQ> &th, double A2) {
        C2 = th.C * th.C;
        a11 = fma(A2, th.R, th.Q * C2);
        a12 = th.Q * th.R;
        AC = th.A * th.C;
    }
};

__device__ __forceinline__ void rescale_pow2(double &n, double &d, int &shift) {
    const int e = ((__double2hiint(d) >> 20) & 0x7ff) - 1023;
    const double sc = __hiloint2double((1023 - e) << 20, 0);
    n *= sc;
    d *= sc;
    shift += e;
}

// PASS2 = false: pass 1 (likelihood terms, state at the end of the segment)
// PASS2 = true : recompute for the smoother: fills J, g, L (see smooth_segment)
template <int PQ, int SEG, bool GUARDED, bool PASS2>
__device__ __forceinline__ void mixed_forward(const Theta<PQ> &th, double A, double A2, double Q,
                                              const MixedConst<PQ> &mc, unsigned bits, int cnt,
                                              const double *__restrict__ yseg, const double *__restrict__ useg,
                                              const double *__restrict__ vseg, double &Xq, double &Vq, double &acc,
                                              double (&Jt)[SEG], double (&g)[SEG], double (&L)[SEG]) {
    double n = Vq, d = 1.0;
    int shift = 0;
#pragma unroll
    for (int j = 0; j < SEG; j++) {
        if (!GUARDED || j < cnt) {
            const bool obs = (bits >> j) & 1u;
            const double m11 = obs ? mc.a11 : A2, m12 = obs ? mc.a12 : Q;
            const double m21 = obs ? mc.C2 : 0.0, m22 = obs ? th.R : 1.0;
            const double nn = fma(m11, n, m12 * d);
            const double dd = fma(m21, n, m22 * d);
            const double rho = fast_rcp(dd);
            const double Bu = dot_row<PQ>(th.B, useg + j * PQ);
            const double Dv = dot_row<PQ>(th.D, vseg + j * PQ);
            const double K = obs ? th.C * n * rho : 0.0;
            const double ymd = (obs ? yseg[j] : 0.0) - Dv; // y is NaN where missing
            const double delta = fma(-th.C, Xq, ymd);      // y - (C Xp + D v)
            const double alpha = fma(-mc.AC, K, A);
            const double beta = fma(A * K, ymd, Bu);
            if (!PASS2) {
                const double w = delta * (d * rho); // delta / Sigma
                acc = fma(obs ? w : 0.0, delta, acc);
            }
            const double Xn = fma(alpha, Xq, beta); // prior mean of the next step
            if (PASS2) {
                const double xu = fma(K, delta, Xq);
                const double nu = obs ? th.R * n : n;
                const double vu = nu * rho;
                if (GUARDED && j == cnt - 1) { // t == T-1: smoothed = filtered (EM.cpp:94-95)
                    Jt[j] = 0.0;
                    g[j] = xu;
                    L[j] = vu;
                } else {
                    const double J = A * nu * fast_rcp(nn);
                    Jt[j] = J;
                    g[j] = fma(-J, Xn, xu);
                    L[j] = vu * fma(-A, J, 1.0); // Vu - J^2 Vp' with J Vp' = A Vu
                }
            }
            Xq = Xn;
            if (j == (GUARDED ? cnt - 1 : SEG - 1)) Vq = nn * rho; // prior variance entering the next segment
            n = nn;
            d = dd;
            if (SEG > 4 && j == SEG / 2 - 1) rescale_pow2(n, d, shift);
        }
    }
    if (!PASS2) acc += fma((double)shift, 0.693147180559945309417, log(d));
}

// ---- pass 2 over one segment -----------------------------------------------------------------
// MIXED   : the segment holds observed steps for some lane -> measurement updates + observed sums
// GUARDED : the segment is the last one: only cnt steps are valid and its last step is T-1
// The filter is recomputed over the segment (EM.cpp:70-90) and, in the same sweep, the RTS gain and
// the affine form of the backward recursion (EM.cpp:100-102) -- all off the smoother's chain:
//      Xs_t = Xu + J (Xs1 - Xp1)   = J Xs1 + g,    g = Xu - J Xp1
//      Vs_t = Vu + J (Vs1 - Vp1) J = J^2 Vs1 + L,  L = Vu - J^2 Vp1
// Only J, g, L (3*SEG doubles) stay in registers.
template <int PQ, int SEG, bool MIXED, bool GUARDED>
__device__ __forceinline__ void smooth_segment(const Theta<PQ> &th, double A, double A2, double Q,
                                               const MixedConst<PQ> &mc, unsigned bits, int cnt,
                                               const double *__restrict__ yseg, const double *__restrict__ useg,
                                               const double *__restrict__ vseg, double Xq, double Vq, double &Xs1,
                                               double &Vs1, Stats<PQ> &st) {
    double Jt[SEG], g[SEG], L[SEG];
    if (MIXED) {
        double unused = 0.0;
        mixed_forward<PQ, SEG, GUARDED, true>(th, A, A2, Q, mc, bits, cnt, yseg, useg, vseg, Xq, Vq, unused, Jt, g, L);
    } else {
#pragma unroll
        for (int j = 0; j < SEG; j++) {
            const double xu = Xq, vu = Vq;
            Xq = fma(A, xu, dot_row<PQ>(th.B, useg + j * PQ));
            Vq = fma(A2, vu, Q);
            const double J = vu * A * fast_rcp(Vq);
            Jt[j] = J;
            g[j] = fma(-J, Xq, xu);
            L[j] = vu * fma(-A, J, 1.0);
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

// MODE 0: series and checkpoints in global memory (series too large for shared memory)
//      1: series staged in shared memory by TMA, checkpoints in global memory
//      2: series and checkpoints in shared memory
template <int PQ, int SEG, int W, int MODE>
__global__ void __launch_bounds__(W * 32, 1) em_chunk_kernel(const EmParams P) {
    constexpr bool STAGED = MODE >= 1;
    LDSR_DYN_SMEM(smem_raw);
    LDSR_STATIC_SMEM(__align__(8) uint64_t, bar);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (STAGED) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
    }
    // The grid is an upper bound on (or, with fewer CTAs than tasks, a divisor of) the task count,
    // which the launch reads from device memory: CTA b takes tasks b, b + gridDim.x, ...
    const int n_tasks = *P.n_tasks;
    unsigned phase = 0;
    for (int ti = blockIdx.x; ti < n_tasks; ti += gridDim.x, phase ^= 1u) {
    if (ti != (int)blockIdx.x) __syncthreads(); // the previous task's shared memory is dead
    const int4 task = P.tasks[ti];
    const SeriesDev S = P.series[task.x];
    const int T = S.T;
    if (STAGED) {
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
    // checkpoints: [warp][segment][Xp|Vp][lane]
    double *__restrict__ ck =
        (MODE == 2 ? reinterpret_cast<double *>(smem_raw + P.ckpt_smem_off) + (size_t)warp * P.max_seg * 64
                   : P.ckpt + ((size_t)(blockIdx.x * W + warp) * P.max_seg) * 64) +
        lane;
    const int nseg = (T + SEG - 1) / SEG;       // the last segment (full or not) holds step T-1
    const int cnt_last = T - (nseg - 1) * SEG;  // 1..SEG valid steps in it

    if (STAGED) mbar_wait(&bar, phase);
    const bool warp_has_fits = warp * 32 < task.z; // else the whole warp idles through this task

    // Which segments hold an observed step: taken from the SERIES (is y finite?), not from a vote over
    // the hold-out masks of the 32 fits that happen to share the warp -- the code path a fit takes must
    // not depend on its neighbours, or its last bits (and, near the tolerance, its iteration count)
    // would change with the order of the batch and with the compaction between launches.  Fixed for the
    // whole launch, so it is computed once here (bitmap in two registers covers 128 segments =
    // T <= 128*SEG; longer series test the segment's y values each time).
    auto seg_has_obs = [&](int sg) -> bool {
        bool m = false;
        for (int t = sg * SEG; t < T && t < (sg + 1) * SEG; ++t) m = m || (ys[t] == ys[t]);
        return m;
    };
    unsigned long long mix0 = 0ull, mix1 = 0ull;
    const bool have_map = nseg <= 128;
    if (have_map)
        for (int sg = 0; sg < nseg; ++sg) {
            if (seg_has_obs(sg)) {
                if (sg < 64)
                    mix0 |= 1ull << sg;
                else
                    mix1 |= 1ull << (sg - 64);
            }
        }
    auto seg_is_mixed = [&](int sg) -> bool {
        if (have_map) return ((sg < 64 ? mix0 >> sg : mix1 >> (sg - 64)) & 1ull) != 0ull;
        return seg_has_obs(sg);
    };

    for (int it = 0; warp_has_fits && it < P.chunk; ++it) {
        if (!__any_sync(FULL, live)) break;
        const double A = th.A, A2 = th.A * th.A, Q = th.Q;
        UnobsConst<SEG> uc;
        uc.set(A, A2, Q);
        MixedConst<PQ> mc;
        mc.set(th, A2);

        // ================= pass 1: forward filter, likelihood, checkpoints =================
        double Xp = th.mu1, Vp = th.V1; // prior of step 0 (EM.cpp:48-49)
        double acc = 0.0;               // sum_obs delta^2/Sigma + log Sigma (EM.cpp:122)
        for (int sg = 0; sg < nseg; ++sg) {
            const int t0 = sg * SEG;
            ck[(size_t)sg * 64] = Xp;
            ck[(size_t)sg * 64 + 32] = Vp;
            const double *__restrict__ useg = us + t0 * PQ;
            if (sg == nseg - 1) {
                double J_[SEG], g_[SEG], L_[SEG];
                mixed_forward<PQ, SEG, true, false>(th, A, A2, Q, mc, seg_bits(mw, t0, SEG), cnt_last, ys + t0, useg,
                                                    vs + t0 * PQ, Xp, Vp, acc, J_, g_, L_);
            } else if (seg_is_mixed(sg)) {
                double J_[SEG], g_[SEG], L_[SEG];
                mixed_forward<PQ, SEG, false, false>(th, A, A2, Q, mc, seg_bits(mw, t0, SEG), SEG, ys + t0, useg,
                                                     vs + t0 * PQ, Xp, Vp, acc, J_, g_, L_);
            } else {
                fwd_unobserved<PQ, SEG>(th, A, uc, useg, Xp, Vp);
            }
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
            smooth_segment<PQ, SEG, true, true>(th, A, A2, Q, mc, seg_bits(mw, t0, SEG), cnt_last, ys + t0, us + t0 * PQ,
                                                vs + t0 * PQ, ck[(size_t)sg * 64], ck[(size_t)sg * 64 + 32], Xs1, Vs1,
                                                st);
        }
        for (int sg = nseg - 2; sg >= 0; --sg) {
            const int t0 = sg * SEG;
            const double Xq = ck[(size_t)sg * 64], Vq = ck[(size_t)sg * 64 + 32]; // (Xp,Vp) entering the segment
            if (seg_is_mixed(sg))
                smooth_segment<PQ, SEG, true, false>(th, A, A2, Q, mc, seg_bits(mw, t0, SEG), SEG, ys + t0,
                                                     us + t0 * PQ, vs + t0 * PQ, Xq, Vq, Xs1, Vs1, st);
            else
                smooth_segment<PQ, SEG, false, false>(th, A, A2, Q, mc, 0u, SEG, ys + t0, us + t0 * PQ, vs + t0 * PQ,
                                                      Xq, Vq, Xs1, Vs1, st);
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
    } // task loop
}

} // namespace ldsr
