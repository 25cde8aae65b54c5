// em_kernel.cuh -- the LDS_EM loop (src/EM.cpp:245-280) as a batched kernel: ONE LANE = ONE FIT.
//
// A CTA takes up to 32*W live fits of one series.  The series blob (y,u,v) is staged into shared
// memory once per launch by a TMA bulk copy and read by every fit through warp-uniform
// (broadcast) loads.  Each lane runs, per EM iteration,
//   pass 1  forward Kalman filter: log-likelihood + a checkpoint (Xp,Vp) every SEG steps;
//   stop rule (EM.cpp:272) -- a fit that stops keeps the theta this E-step ran with;
//   pass 2  RTS smoother, segment by segment from the end: the segment's filter states are
//           recomputed from its checkpoint into REGISTERS (Xu,Vu,Xp1 for SEG steps), then the
//           backward recursion runs over them and accumulates the M-step sums on the fly;
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

template <int PQ, int SEG, int W>
__global__ void __launch_bounds__(W * 32) em_chunk_kernel(const EmParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;

    const int4 task = P.tasks[blockIdx.x];
    const SeriesDev S = P.series[task.x];
    const int T = S.T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    const double *ser;
    if (P.blob_in_smem) {
        if (threadIdx.x == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (threadIdx.x == 0) stage_blob(smem_raw, P.blobs + S.blob_off, (unsigned)S.blob_doubles * 8u, &bar);
        ser = reinterpret_cast<const double *>(smem_raw);
    } else {
        ser = P.blobs + S.blob_off;
    }
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
    const int nseg = (T + SEG - 1) / SEG;

    if (P.blob_in_smem) mbar_wait(&bar, 0);
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
            if (__any_sync(FULL, bits != 0u)) {
#pragma unroll
                for (int j = 0; j < SEG; j++) {
                    const int t = t0 + j;
                    if (t < T) {
                        const bool obs = (bits >> j) & 1u;
                        double Xu = Xp, Vu = Vp;
                        if (__any_sync(FULL, obs)) {
                            double dq, Sg;
                            measurement_update<PQ>(th, obs, ys[t], vs + (size_t)t * PQ, Xp, Vp, Xu, Vu, dq, Sg);
                            if (obs) acc += dq + log(Sg);
                        }
                        Xp = fma(A, Xu, dot_row<PQ>(th.B, us + (size_t)t * PQ)); // u lags one step (EM.cpp:74)
                        Vp = fma(A2, Vu, Q);                                      // A*Vu*A + Q (EM.cpp:76)
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < SEG; j++) {
                    const int t = t0 + j;
                    if (t < T) {
                        Xp = fma(A, Xp, dot_row<PQ>(th.B, us + (size_t)t * PQ));
                        Vp = fma(A2, Vp, Q);
                    }
                }
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
        for (int sg = nseg - 1; sg >= 0; --sg) {
            const int t0 = sg * SEG;
            double Xq = ck[(size_t)sg * 64], Vq = ck[(size_t)sg * 64 + 32]; // (Xp,Vp) entering the segment
            const unsigned bits = seg_bits(mw, t0, SEG);
            const bool anyobs = __any_sync(FULL, bits != 0u);
            double Xu[SEG], Vu[SEG], Xp1[SEG];
            // ---- recompute the filter over the segment into registers
#pragma unroll
            for (int j = 0; j < SEG; j++) {
                const int t = t0 + j;
                if (t < T) {
                    const bool obs = (bits >> j) & 1u;
                    double xu = Xq, vu = Vq;
                    if (anyobs && __any_sync(FULL, obs)) {
                        double dq, Sg;
                        measurement_update<PQ>(th, obs, ys[t], vs + (size_t)t * PQ, Xq, Vq, xu, vu, dq, Sg);
                    }
                    Xu[j] = xu;
                    Vu[j] = vu;
                    Xq = fma(A, xu, dot_row<PQ>(th.B, us + (size_t)t * PQ));
                    Vq = fma(A2, vu, Q);
                    Xp1[j] = Xq;
                }
            }
            // ---- backward over the segment (EM.cpp:99-104) with the sums of EM.cpp:151-161,180-193
#pragma unroll
            for (int j = SEG - 1; j >= 0; j--) {
                const int t = t0 + j;
                if (t < T) {
                    double Xs, Vs;
                    if (t == T - 1) {
                        Xs = Xu[j];
                        Vs = Vu[j];
                        st.XT = Xs;
                        st.VT = Vs;
                    } else {
                        const double Vp1 = fma(A2, Vu[j], Q);
                        const double J = Vu[j] * A * (1.0 / Vp1);
                        Xs = fma(J, Xs1 - Xp1[j], Xu[j]);
                        Vs = fma(J * (Vs1 - Vp1), J, Vu[j]);
                        st.Tx1x = fma(Xs1, Xs, fma(Vs1, J, st.Tx1x));
                        st.Txx += fma(Xs, Xs, Vs);
                        const double *__restrict__ ut = us + (size_t)t * PQ;
#pragma unroll
                        for (int k = 0; k < PQ; k++) {
                            st.Tx1u[k] = fma(Xs1, ut[k], st.Tx1u[k]);
                            st.Tux[k] = fma(ut[k], Xs, st.Tux[k]);
                        }
                    }
                    if (anyobs) {
                        const bool obs = (bits >> j) & 1u;
                        if (__any_sync(FULL, obs)) {
                            const double yo = obs ? ys[t] : 0.0;
                            const double xo = obs ? Xs : 0.0;
                            st.Syx = fma(yo, xo, st.Syx);
                            st.Sxx += obs ? fma(Xs, Xs, Vs) : 0.0;
                            const double *__restrict__ vt = vs + (size_t)t * PQ;
#pragma unroll
                            for (int k = 0; k < PQ; k++) st.Sxv[k] = fma(xo, vt[k], st.Sxv[k]);
                        }
                    }
                    Xs1 = Xs;
                    Vs1 = Vs;
                }
            }
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
