// em_scan_kernel.cuh -- the LDS_EM loop (src/EM.cpp:245-280) for SMALL batches: ONE CTA = ONE FIT,
// one thread = L consecutive time steps, the recursions over time done as parallel scans.
//
// Why: the batched kernels put a fit on a lane, so a call with few fits -- LDS_EM (1 fit),
// LDS_EM_restart / LDS_reconstruction with 20..100 restarts (R/LDS_reconstruction.R:42-62, the call
// every user makes) -- fills a handful of CTAs and pays the full serial latency of an iteration
// (10 us at T = 413: 10 ms for 1000 iterations whether 1 or 4000 fits run).  Here the time axis of ONE
// fit is spread over the threads of a CTA and every dependency along t becomes a scan of depth
// log2(threads) (hard parts (ii); Appendix D.4):
//
//   P1  each thread composes the VARIANCE map of its L steps -- in 1-D the Riccati step is a Moebius map
//       of Vp, a 2x2 matrix on homogeneous coordinates (em_split_kernel.cuh) -- exclusive scan of 2x2
//       matrices (warp shuffles, then the warp totals through shared memory)  -> Vp entering every thread
//   P2  forward over the L steps with the true variances; the mean in the (P, q) basis of its unknown
//       incoming value; scan of the affine maps -> incoming mean of every thread, the likelihood terms
//       (a quadratic in the incoming mean) summed over the CTA -> stop rule (EM.cpp:272)
//   P3  suffix scan of the threads' affine BACKWARD maps (Xs_first = PJ Xs_in + g, Vs_first = PJ^2 Vs_in
//       + L), started from the prior of the virtual step T (which makes Xs_{T-1} = Xu_{T-1}, EM.cpp:94-95)
//   P4  backward over the L steps (gains kept in registers from P2), M-step sums; a transposing
//       shuffle reduction brings the 7 + 3 PQ sums of the CTA together; M-step by two threads
//       (observation block / transition block); theta back to every thread through shared memory.
//
// The thread's rows of y, u, v stay in REGISTERS for the whole launch (they never change); steps at or
// beyond T (the last thread's ragged end, idle threads) are exact identities of every map.
// Five barriers per iteration; nothing per step goes to shared or global memory.
#pragma once
#include "em_split_kernel.cuh"
#ifdef LDSR_PHASE_CLOCKS
#include <cstdio>
#endif

namespace ldsr {

constexpr int SCAN_MAX_WARPS = 8;

template <int PQ> __host__ __device__ constexpr int scan_nsum() { return 7 + 3 * PQ; }
template <int PQ> __host__ __device__ constexpr int scan_nsum_pad() {
    return scan_nsum<PQ>() <= 16 ? 16 : (scan_nsum<PQ>() <= 32 ? 32 : 48); // reduced as 16, 32 or 32 + 16 values
}
constexpr int SCAN_SUM_ROW = 64; // doubles per warp in the partial-sum area
// wide inputs (PQ > SCAN_SHARE_UV_FROM) keep ONE set of rows in registers: the plan must have v == u
constexpr int SCAN_SHARE_UV_FROM = 5;

// ---- warp scans (Hillis-Steele over the 32 lanes) -------------------------------------------------
// inclusive prefix product of 2x2 matrices: lane l ends with M_l M_{l-1} ... M_0
__device__ __forceinline__ void scan_mobius_incl(double &m11, double &m12, double &m21, double &m22, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double p11 = __shfl_up_sync(FULL, m11, o), p12 = __shfl_up_sync(FULL, m12, o);
        const double p21 = __shfl_up_sync(FULL, m21, o), p22 = __shfl_up_sync(FULL, m22, o);
        if (lane >= o) {
            const double n11 = fma(m11, p11, m12 * p21), n12 = fma(m11, p12, m12 * p22);
            const double n21 = fma(m21, p11, m22 * p21), n22 = fma(m21, p12, m22 * p22);
            m11 = n11;
            m12 = n12;
            m21 = n21;
            m22 = n22;
        }
        // only ratios matter: keep the common scale near 1 (the normalisation factors of 32 maps would
        // otherwise multiply up to 2^-32 per map and leave the double range on extreme Q / R)
        if (o == 2 || o == 8) rescale4(m11, m12, m21, m22);
    }
    rescale4(m11, m12, m21, m22);
}
// inclusive prefix composition of affine maps x -> P x + q (lane l: maps of lanes 0..l applied in order)
__device__ __forceinline__ void scan_affine_incl(double &P, double &q, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double Pp = __shfl_up_sync(FULL, P, o), qp = __shfl_up_sync(FULL, q, o);
        if (lane >= o) {
            q = fma(P, qp, q);
            P *= Pp;
        }
    }
}
// inclusive SUFFIX composition of backward maps (x, v) -> (PJ x + g, PJ^2 v + Lc): lane l ends with its own
// map applied after those of lanes l+1 .. 31
__device__ __forceinline__ void scan_backward_incl(double &PJ, double &g, double &Lc, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double PJr = __shfl_down_sync(FULL, PJ, o), gr = __shfl_down_sync(FULL, g, o);
        const double Lr = __shfl_down_sync(FULL, Lc, o);
        if (lane + o < 32) {
            g = fma(PJ, gr, g);
            Lc = fma(PJ * PJ, Lr, Lc);
            PJ *= PJr;
        }
    }
}
__device__ __forceinline__ double scan_warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}
// Transposing reduction of N (16 or 32) values per lane over the warp: halving exchanges, so that after
// log2(N) rounds every lane holds ONE value summed over a group of lanes, then the remaining rounds
// finish it.  Returns the warp total of value index (lane * N) / 32  (N = 16: lane >> 1; N = 32: lane).
// 2 N - 2 (+ 32/N - 1) value exchanges instead of 5 N.
template <int N> __device__ __forceinline__ double scan_reduce_many(double (&v)[N], int lane) {
    static_assert(N == 16 || N == 32, "N is 16 or 32");
    int off = 16;
#pragma unroll
    for (int h = N / 2; h >= 1; h >>= 1, off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < h; i++) {
            const double send = upper ? v[i] : v[i + h];
            const double keep = upper ? v[i + h] : v[i];
            v[i] = keep + __shfl_xor_sync(FULL, send, off);
        }
    }
    double r = v[0];
    if (N == 16) r += __shfl_xor_sync(FULL, r, 1);
    return r;
}

// element idx of the thread's v rows: its own copy, or the u rows when the plan has v == u (no pointer games:
// both arrays must stay plain register arrays)
template <bool SHARE, int N>
__device__ __forceinline__ double scan_vrow(const double (&u)[N], const double (&v)[SHARE ? 1 : N], int idx) {
    if constexpr (SHARE)
        return u[idx];
    else
        return v[idx];
}

// EMIT = true: not EM but its E-step once, for the winners of the groups: the smoothed X, V, the gains J and
// Y = C X + D v of every step are written out (what smoother_kernel computes with one thread per job in
// 2 T dependent steps -- 0.12 ms at T = 413 whatever the number of jobs; here about 3 us).
#ifdef LDSR_PHASE_CLOCKS // development build: cycles per phase of the iteration, threads 0 and 32 of CTA 0 (printf)
#define SCAN_MARK(kk)                                \
    do {                                             \
        const long long c_ = clock64();              \
        pc[kk] += c_ - tprev;                        \
        tprev = c_;                                  \
    } while (0)
#else
#define SCAN_MARK(kk) ((void)0)
#endif

// DENSE = true: compiled for three CTAs of at most 128 threads per SM (168 registers, 32 B of stack at PQ = 3) instead
// of two (244 registers): for batches of more than two fits per SM the third resident CTA hides more latency
// than the spills cost (NP-413, 1000 iterations: 400 / 900 fits 4.63 -> 3.92 / 8.76 -> 7.49 ms), for smaller
// ones it is 2.5 % slower (200 fits 3.11 -> 3.19 ms) and not used.
template <int PQ, int L, bool EMIT = false, bool DENSE = false>
__global__ void __launch_bounds__(DENSE ? 128 : SCAN_MAX_WARPS * 32, DENSE ? 3 : 1) em_scan_kernel(const EmParams P) {
    static_assert(L == 1 || L == 2 || L == 4 || L == 8, "steps per thread");
    constexpr int NS = scan_nsum<PQ>(), NSP = scan_nsum_pad<PQ>();
    static_assert(NS <= 48 && NSP <= SCAN_SUM_ROW, "7 + 3 PQ sums are reduced as 32 + 16 values at most");
    constexpr bool SHARE_UV = PQ >= SCAN_SHARE_UV_FROM;
    constexpr int TL = theta_pad_len<PQ>();
    LDSR_STATIC_SMEM(double, S1[SCAN_MAX_WARPS * 4]);  // variance-map totals of the warps
    LDSR_STATIC_SMEM(double, S2[SCAN_MAX_WARPS * 2]);  // mean-map totals
    LDSR_STATIC_SMEM(double, S3[SCAN_MAX_WARPS]);      // likelihood partial sums
    LDSR_STATIC_SMEM(double, S3M[SCAN_MAX_WARPS]);     // prod Sigma of the warps: mantissa product ...
    LDSR_STATIC_SMEM(int, S3E[SCAN_MAX_WARPS]);        // ... and exponent sum
    LDSR_STATIC_SMEM(double, S_LIK[1]);                // the iteration's likelihood and whether the fit goes on
    LDSR_STATIC_SMEM(int, S_LIVE[1]);
    LDSR_STATIC_SMEM(double, S4[SCAN_MAX_WARPS * 3]);  // backward-map totals
    LDSR_STATIC_SMEM(double, S5[SCAN_MAX_WARPS * SCAN_SUM_ROW]); // M-step partial sums [warp][NSP]
    LDSR_STATIC_SMEM(double, TOTS[2 * SCAN_SUM_ROW]);            // their totals, one copy per M-step warp
    LDSR_STATIC_SMEM(double, ENDS[4]);                 // X0, V0, XT, VT
    LDSR_STATIC_SMEM(double, THS[TL]);                 // theta after the M-step

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int n_tasks = EMIT ? P.n_jobs : *P.n_tasks;
    for (int ti = blockIdx.x; ti < n_tasks; ti += gridDim.x) {
    if (ti != (int)blockIdx.x) __syncthreads();
    const int4 task = EMIT ? make_int4(P.g_series[P.job_group[ti]], 0, 1, 0) : P.tasks[ti];
    const SeriesDev S = P.series[task.x];
    const int T = S.T;
    LDSR_CHECK((int)blockDim.x * L >= T && nw <= SCAN_MAX_WARPS && task.z == 1); // the block covers the series
    LDSR_CHECK(!SHARE_UV || S.same_uv);
    const int fit = EMIT ? P.job_theta[ti] : P.active[task.y];
    if (EMIT && fit < 0) { // no restart of the group qualified: NaN rows (uniform over the CTA)
        const double nan = __longlong_as_double(0x7ff8000000000000ULL);
        const long long row = P.job_row[ti];
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            P.tX[row + t] = P.tY[row + t] = P.tV[row + t] = nan;
            if (P.tJ) P.tJ[row + t] = nan;
        }
        continue;
    }
    const int grp = EMIT ? P.job_group[ti] : P.f_group[fit];
    const double *__restrict__ gc = P.gconst + (size_t)grp * gconst_stride(PQ);
    const double *__restrict__ tuu_inv = P.sconst + S.sconst_off;
    const double n_obs = gc[1], inv_n_obs = 1.0 / n_obs;

    // ---- this thread's steps: rows and mask bits, in registers for the whole launch
    const int t0 = (int)threadIdx.x * L;
    double yr[L], ur[L * PQ], vr_own[SHARE_UV ? 1 : L * PQ]; // v == u (wide inputs): one set of rows, see scan_vrow
    unsigned bits = 0u; // observed steps (the group's mask: finite(y) minus the hold-outs)
    {
        const double *__restrict__ blob = P.blobs + S.blob_off;
        const unsigned *__restrict__ mw = P.masks + P.g_mask_off[grp];
#pragma unroll
        for (int j = 0; j < L; j++) {
            const int t = t0 + j;
            const bool real = t < T;
            yr[j] = real ? blob[S.y_off + t] : 0.0;
#pragma unroll
            for (int i = 0; i < PQ; i++) {
                ur[j * PQ + i] = real ? blob[S.u_off + (size_t)t * PQ + i] : 0.0;
                if constexpr (!SHARE_UV) vr_own[j * PQ + i] = real ? blob[S.v_off + (size_t)t * PQ + i] : 0.0;
            }
            if (real && ((mw[t >> 5] >> (t & 31)) & 1u)) bits |= 1u << j;
        }
#pragma unroll
        for (int j = 0; j < L; j++)
            if (!((bits >> j) & 1u)) yr[j] = 0.0; // y is NaN where missing: never let it into the arithmetic
    }
    Theta<PQ> th;
    load_theta<PQ>(th, P.theta + (size_t)fit * TL);
    double l1 = 0.0, l2 = 0.0, lik = 0.0;
    int ne = 0;
    bool live = true;
    if constexpr (!EMIT) {
        l1 = P.l1[fit], l2 = P.l2[fit], lik = P.lik[fit];
        ne = P.ne[fit];
        live = P.done[fit] == 0;
        if (live && P.g_status[grp] != 0) { // Gram block not invertible: the reference would throw
            live = false;
            lik = __longlong_as_double(0x7ff8000000000000ULL);
        }
    }

#ifdef LDSR_PHASE_CLOCKS
    long long pc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
    int n_it_done = 0;
#endif
    for (int it = 0; live && it < (EMIT ? 1 : P.chunk); ++it) {
        SCAN_MARK(0);
        const double A = th.A, Q = th.Q, Cc = th.C, R = th.R;

        // the input terms of my steps do not wait for anything: issued here, they overlap the first scan
        double Bu[L], ymd[L];
#pragma unroll
        for (int j = 0; j < L; j++) {
            double b = 0.0, dv = 0.0;
#pragma unroll
            for (int i = 0; i < PQ; i++) {
                b = fma(th.B[i], ur[j * PQ + i], b);
                dv = fma(th.D[i], scan_vrow<SHARE_UV, L * PQ>(ur, vr_own, j * PQ + i), dv);
            }
            Bu[j] = b;
            ymd[j] = yr[j] - dv;
        }

        // ================= P1: variance map of my steps, scan =================
        double e11, e12, e21, e22; // exclusive prefix inside the warp
        // Per-step coefficients of this iteration.  An unobserved step is the observed one with C = 0, R = 1
        // ([[A^2 R + Q C^2, Q R],[C^2, R]] becomes [[A^2, Q],[0, 1]], K = 0, Vu = Vp), and a step at or beyond T
        // is the unobserved one with A = 1, Q = 0 (the identity): two selects each instead of one per formula.
        double Aj[L], Cj[L], Rj[L], s11[L], s12[L], s21[L];
        {
            double m11 = 1.0, m12 = 0.0, m21 = 0.0, m22 = 1.0;
#pragma unroll
            for (int j = 0; j < L; j++) {
                const bool real = t0 + j < T, obs = (bits >> j) & 1u;
                Aj[j] = real ? A : 1.0;
                const double Qj = real ? Q : 0.0;
                Cj[j] = obs ? Cc : 0.0;
                Rj[j] = obs ? R : 1.0;
                s21[j] = Cj[j] * Cj[j];
                s11[j] = fma(Aj[j] * Aj[j], Rj[j], Qj * s21[j]);
                s12[j] = Qj * Rj[j];
                const double n11 = fma(s11[j], m11, s12[j] * m21), n12 = fma(s11[j], m12, s12[j] * m22);
                const double n21 = fma(s21[j], m11, Rj[j] * m21), n22 = fma(s21[j], m12, Rj[j] * m22);
                m11 = n11;
                m12 = n12;
                m21 = n21;
                m22 = n22;
            }
            rescale4(m11, m12, m21, m22);
            scan_mobius_incl(m11, m12, m21, m22, lane);
            if (lane == 31) {
                S1[warp * 4 + 0] = m11;
                S1[warp * 4 + 1] = m12;
                S1[warp * 4 + 2] = m21;
                S1[warp * 4 + 3] = m22;
            }
            e11 = __shfl_up_sync(FULL, m11, 1);
            e12 = __shfl_up_sync(FULL, m12, 1);
            e21 = __shfl_up_sync(FULL, m21, 1);
            e22 = __shfl_up_sync(FULL, m22, 1);
            if (lane == 0) {
                e11 = e22 = 1.0;
                e12 = e21 = 0.0;
            }
        }
        SCAN_MARK(1);
        __syncthreads(); // B1
        SCAN_MARK(2);
        // prior variance entering my steps in homogeneous coordinates (Vin = ni / di: P2 only uses ratios, so the
        // division stays off the critical path); prior variance of the virtual step after the last
        double ni, di, vend;
        {
            // the warps' maps first (independent loads), then the chain; every map was normalised to entries
            // summing to [1,2), so at most SCAN_MAX_WARPS of them cannot leave the double range
            double a11[SCAN_MAX_WARPS], a12[SCAN_MAX_WARPS], a21[SCAN_MAX_WARPS], a22[SCAN_MAX_WARPS];
#pragma unroll
            for (int w = 0; w < SCAN_MAX_WARPS; ++w) {
                const bool on = w < nw;
                a11[w] = on ? S1[w * 4 + 0] : 1.0;
                a12[w] = on ? S1[w * 4 + 1] : 0.0;
                a21[w] = on ? S1[w * 4 + 2] : 0.0;
                a22[w] = on ? S1[w * 4 + 3] : 1.0;
            }
            double n = th.V1, d = 1.0, nwp = th.V1, dwp = 1.0;
#pragma unroll
            for (int w = 0; w < SCAN_MAX_WARPS; ++w) {
                if (w == warp) {
                    nwp = n;
                    dwp = d;
                }
                const double nn = fma(a11[w], n, a12[w] * d), dd = fma(a21[w], n, a22[w] * d);
                n = nn;
                d = dd;
            }
            vend = n * fast_rcp(d);
            ni = fma(e11, nwp, e12 * dwp);
            di = fma(e21, nwp, e22 * dwp);
        }

        SCAN_MARK(3);
        // ================= P2: forward over my steps (EM.cpp:70-90), mean in the (P, q) basis =================
        double Kg[L], Jg[L], Lg[L], g0[L], gP[L]; // kept for P4
        double Pth, qth;                           // my affine map of the mean
        double acc_l0, acc_l1, acc_l2;             // sum_obs delta^2/Sigma = l0 - 2 C x l1 + C^2 x^2 l2
        double dratio;                             // prod_obs Sigma over my steps
        double PJ, G0, GG, Lc;                     // my backward map
        {
            // gains: the variance in homogeneous coordinates (n, d), all reciprocals after the chain
            double nj[L], dj[L], nn[L], dd[L];
            {
                double n = ni, d = di;
#pragma unroll
                for (int j = 0; j < L; j++) {
                    nj[j] = n;
                    dj[j] = d;
                    nn[j] = fma(s11[j], n, s12[j] * d);
                    dd[j] = fma(s21[j], n, Rj[j] * d);
                    n = nn[j];
                    d = dd[j];
                }
            }
            double rS[L], alpha[L];
            double dend = dd[L - 1];
#pragma unroll
            for (int j = 0; j < L; j++) {
                const bool real = t0 + j < T, obs = (bits >> j) & 1u;
                const double rho = fast_rcp(dd[j]), rn = fast_rcp(nn[j]);
                const double K = Cj[j] * nj[j] * rho;
                const double nu = Rj[j] * nj[j];
                const double vu = nu * rho;
                const double J = real ? Aj[j] * nu * rn : 1.0;
                Kg[j] = K;
                rS[j] = obs ? dj[j] * rho : 0.0; // 1/Sigma
                alpha[j] = fma(-Aj[j] * Cj[j], K, Aj[j]);
                Jg[j] = J;
                Lg[j] = real ? vu * fma(-Aj[j], J, 1.0) : 0.0; // Vu - J^2 Vp' with J Vp' = A Vu
            }
            double q[L + 1], Pc[L + 1];
            q[0] = 0.0;
            Pc[0] = 1.0;
#pragma unroll
            for (int j = 0; j < L; j++) {
                const double beta = fma(Aj[j] * Kg[j], ymd[j], Bu[j]); // rows at or beyond T are zero
                q[j + 1] = fma(alpha[j], q[j], beta);
                Pc[j + 1] = alpha[j] * Pc[j];
            }
            acc_l0 = acc_l1 = acc_l2 = 0.0;
            PJ = 1.0;
            G0 = GG = Lc = 0.0;
            double pj2 = 1.0;
#pragma unroll
            for (int j = 0; j < L; j++) {
                const double d0 = fma(-Cc, q[j], ymd[j]); // innovation for x_in = 0
                const double w0 = rS[j] * d0;
                acc_l0 = fma(w0, d0, acc_l0);
                acc_l1 = fma(w0, Pc[j], acc_l1);
                acc_l2 = fma(rS[j] * Pc[j], Pc[j], acc_l2);
                const double xu0 = fma(Kg[j], d0, q[j]);
                const double xuP = Pc[j] * fma(-Kg[j], Cc, 1.0);
                g0[j] = fma(-Jg[j], q[j + 1], xu0);
                gP[j] = fma(-Jg[j], Pc[j + 1], xuP);
                G0 = fma(PJ, g0[j], G0);
                GG = fma(PJ, gP[j], GG);
                Lc = fma(pj2, Lg[j], Lc);
                PJ *= Jg[j];
                pj2 *= Jg[j] * Jg[j];
            }
            Pth = Pc[L];
            qth = q[L];
            // sum_obs log Sigma = log prod Sigma = log(d_end / d_in): the ratios are multiplied up over the CTA
            // (mantissas and exponents apart) and ONE thread takes ONE logarithm, beside the M-step (see there)
            dratio = bits ? dend * fast_rcp(di) : 1.0;
        }
        SCAN_MARK(4);
        double xin, xend; // prior mean entering my steps; prior mean of the virtual step after the last
        {
            double Pi = Pth, qi = qth;
            scan_affine_incl(Pi, qi, lane);
            if (lane == 31) {
                S2[warp * 2 + 0] = Pi;
                S2[warp * 2 + 1] = qi;
            }
            double Pe = __shfl_up_sync(FULL, Pi, 1), qe = __shfl_up_sync(FULL, qi, 1);
            if (lane == 0) {
                Pe = 1.0;
                qe = 0.0;
            }
            SCAN_MARK(5);
            __syncthreads(); // B2
            double wp[SCAN_MAX_WARPS], wq[SCAN_MAX_WARPS];
#pragma unroll
            for (int w = 0; w < SCAN_MAX_WARPS; ++w) {
                wp[w] = w < nw ? S2[w * 2 + 0] : 1.0;
                wq[w] = w < nw ? S2[w * 2 + 1] : 0.0;
            }
            double x = th.mu1, xw = th.mu1; // prior of step 0 (EM.cpp:48)
#pragma unroll
            for (int w = 0; w < SCAN_MAX_WARPS; ++w) {
                if (w == warp) xw = x;
                x = fma(wp[w], x, wq[w]);
            }
            xend = x;
            xin = fma(Pe, xw, qe);
        }
        SCAN_MARK(6);
        // ---- likelihood partial sums and the backward maps, one barrier for both
        const double gk = fma(GG, xin, G0);
        double PJs = PJ, gs = gk, Ls = Lc;
        {
            const double tC = Cc * xin;
            const double accw = scan_warp_sum(fma(tC, fma(tC, acc_l2, -2.0 * acc_l1), acc_l0));
            // prod Sigma over the warp: 32 mantissas in [1,2) (product < 2^32) and the sum of the exponents
            const int dhi = __double2hiint(dratio);
            int pe = ((dhi >> 20) & 0x7ff) - 1023;
            double pm = __hiloint2double((dhi & 0x000fffff) | 0x3ff00000, __double2loint(dratio));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                pm *= __shfl_xor_sync(FULL, pm, o);
                pe += __shfl_xor_sync(FULL, pe, o);
            }
            scan_backward_incl(PJs, gs, Ls, lane);
            if (lane == 0) {
                S3M[warp] = pm;
                S3E[warp] = pe;
                S3[warp] = accw;
                S4[warp * 3 + 0] = PJs;
                S4[warp * 3 + 1] = gs;
                S4[warp * 3 + 2] = Ls;
            }
        }
        double PJe = __shfl_down_sync(FULL, PJs, 1), ge = __shfl_down_sync(FULL, gs, 1);
        double Le = __shfl_down_sync(FULL, Ls, 1);
        if (lane == 31) {
            PJe = 1.0;
            ge = 0.0;
            Le = 0.0;
        }
        SCAN_MARK(7);
        __syncthreads(); // B3
        // (the likelihood and the stop rule are computed by one thread beside the M-step and applied after barrier 5:
        //  nothing of the backward pass depends on them)
        SCAN_MARK(8);
        // ---- smoothed state entering my steps from the right: the chain starts from the prior of the
        //      virtual step after the last one, Xs_{T-1} = Xu_{T-1} (EM.cpp:94-95)
        double Xs1, Vs1;
        {
            double X = xend, V = vend;
#pragma unroll
            for (int w = SCAN_MAX_WARPS - 1; w > 0; --w) {
                const bool on = w < nw && w > warp;
                const double pj = on ? S4[w * 3 + 0] : 1.0;
                X = fma(pj, X, on ? S4[w * 3 + 1] : 0.0);
                V = fma(pj * pj, V, on ? S4[w * 3 + 2] : 0.0);
            }
            Xs1 = fma(PJe, X, ge);
            Vs1 = fma(PJe * PJe, V, Le);
        }

        // ================= P4: backward over my steps (EM.cpp:99-104), M-step sums =================
        double sums[NSP];
#pragma unroll
        for (int i = 0; i < NSP; i++) sums[i] = 0.0;
        // layout: 0 Syx, 1 Sxx, 2 Sxxv, 3 Tx1x, 4 Tx1xv, 5 Txx, 6 Txxv, 7.. Sxv, 7+PQ.. Tx1u, 7+2PQ.. Tux
        {
            double Xs[L + 1], Vs[L + 1];
            Xs[L] = Xs1;
            Vs[L] = Vs1;
#pragma unroll
            for (int j = L - 1; j >= 0; j--) {
                Xs[j] = fma(Jg[j], Xs[j + 1], fma(gP[j], xin, g0[j]));
                Vs[j] = fma(Jg[j] * Jg[j], Vs[j + 1], Lg[j]);
            }
            if constexpr (EMIT) { // the outputs of EM.cpp:94-110 for my steps, and that is all
                const long long row = P.job_row[ti];
#pragma unroll
                for (int j = 0; j < L; j++) {
                    const int t = t0 + j;
                    if (t < T) {
                        double dv = 0.0;
#pragma unroll
                        for (int i = 0; i < PQ; i++) dv = fma(th.D[i], scan_vrow<SHARE_UV, L * PQ>(ur, vr_own, j * PQ + i), dv);
                        P.tX[row + t] = Xs[j];
                        P.tV[row + t] = Vs[j];
                        P.tY[row + t] = fma(Cc, Xs[j], dv);
                        if (P.tJ) P.tJ[row + t] = Jg[j]; // A Vu_t / Vp_{t+1}, also at t = T - 1 (EM.cpp:98)
                    }
                }
            }
            if constexpr (!EMIT) {
#pragma unroll
            for (int j = 0; j < L; j++) {
                const int t = t0 + j;
                const bool trans = t < T - 1, obs = (bits >> j) & 1u;
                const double xa = trans ? Xs[j] : 0.0, xb = trans ? Xs[j + 1] : 0.0;
                sums[3] = fma(xb, xa, sums[3]);
                sums[4] = fma(trans ? Vs[j + 1] : 0.0, Jg[j], sums[4]);
                sums[5] = fma(xa, xa, sums[5]);
                sums[6] += trans ? Vs[j] : 0.0;
                const double xo = obs ? Xs[j] : 0.0;
                sums[0] = fma(yr[j], xo, sums[0]);
                sums[1] = fma(xo, xo, sums[1]);
                sums[2] += obs ? Vs[j] : 0.0;
#pragma unroll
                for (int i = 0; i < PQ; i++) {
                    sums[7 + i] = fma(xo, scan_vrow<SHARE_UV, L * PQ>(ur, vr_own, j * PQ + i), sums[7 + i]);
                    sums[7 + PQ + i] = fma(xb, ur[j * PQ + i], sums[7 + PQ + i]);
                    sums[7 + 2 * PQ + i] = fma(ur[j * PQ + i], xa, sums[7 + 2 * PQ + i]);
                }
                if (t == T - 1) {
                    ENDS[2] = Xs[j];
                    ENDS[3] = Vs[j];
                }
            }
            if (threadIdx.x == 0) {
                ENDS[0] = Xs[0];
                ENDS[1] = Vs[0];
            }
            } // !EMIT
        }
        if constexpr (EMIT) break; // nothing else of the iteration is needed (no barrier is pending)
        SCAN_MARK(9);
        {
            double *const row = S5 + warp * SCAN_SUM_ROW;
            if constexpr (NSP == 16) {
                const double r = scan_reduce_many<16>(sums, lane);
                if ((lane & 1) == 0) row[lane >> 1] = r;
            } else {
                const double r = scan_reduce_many<32>(*reinterpret_cast<double(*)[32]>(&sums[0]), lane);
                row[lane] = r;
                if constexpr (NSP == 48) {
                    const double r2 = scan_reduce_many<16>(*reinterpret_cast<double(*)[16]>(&sums[32]), lane);
                    if ((lane & 1) == 0) row[32 + (lane >> 1)] = r2;
                }
            }
        }
        SCAN_MARK(10);
        __syncthreads(); // B4
        SCAN_MARK(11);

        // ================= M-step (EM.cpp:139-229): warp 0 the observation block, warp 1 the transition block ======
        // (one warp: warp 0 does both).  Lane a < PQ takes row a of the matrix-vector products of the block
        // elimination (lds_math.cuh); the dot products over a are warp sums.
        if constexpr (PQ <= 4) {
        // narrow inputs: thread 0 does the observation block, thread 32 the transition block (measured on NP-413:
        // 2.91 ms per 1000 iterations against 3.15 ms with the rows spread over lanes and warp sums)
        const bool do_obs = warp == 0, do_trans = nw > 1 ? warp == 1 : warp == 0; // one warp: warp 0 does both
        if (do_obs || do_trans) {
            if (lane < NS) {
                double a = 0.0;
#pragma unroll
                for (int w = 0; w < SCAN_MAX_WARPS; ++w) a += w < nw ? S5[w * SCAN_SUM_ROW + lane] : 0.0; // loads first
                TOTS[warp * SCAN_SUM_ROW + lane] = a;
            }
            __syncwarp();
            if (lane == 0) {
                Stats<PQ> st;
                const double *o = TOTS + warp * SCAN_SUM_ROW;
                st.Syx = o[0];
                st.Sxx = o[1];
                st.Sxxv = o[2];
                st.Tx1x = o[3];
                st.Tx1xv = o[4];
                st.Txx = o[5];
                st.Txxv = o[6];
#pragma unroll
                for (int i = 0; i < PQ; i++) {
                    st.Sxv[i] = o[7 + i];
                    st.Tx1u[i] = o[7 + PQ + i];
                    st.Tux[i] = o[7 + 2 * PQ + i];
                }
                st.X0 = ENDS[0];
                st.V0 = ENDS[1];
                st.XT = ENDS[2];
                st.VT = ENDS[3];
                Theta<PQ> tn = th;
                if (do_obs) {
                    mstep_obs_block<PQ, true>(st, gc, tn);
                    THS[1 + PQ] = tn.C;
#pragma unroll
                    for (int i = 0; i < PQ; i++) THS[2 + PQ + i] = tn.D[i];
                    THS[3 + 2 * PQ] = tn.R;
                }
                if (do_trans) {
                    mstep_trans_block<PQ, true>(st, tuu_inv, T, tn);
                    THS[0] = tn.A;
#pragma unroll
                    for (int i = 0; i < PQ; i++) THS[1 + i] = tn.B[i];
                    THS[2 + 2 * PQ] = tn.Q;
                    THS[4 + 2 * PQ] = tn.mu1;
                    THS[5 + 2 * PQ] = tn.V1;
                }
            }
        }
        } else {
        const bool do_obs = warp == 0, do_trans = nw > 1 ? warp == 1 : warp == 0;
        if (do_obs || do_trans) {
            double *const o = TOTS + warp * SCAN_SUM_ROW;
            for (int i = lane; i < NS; i += 32) {
                double a = 0.0;
                for (int w = 0; w < nw; ++w) a += S5[w * SCAN_SUM_ROW + i];
                o[i] = a;
            }
            __syncwarp();
            const bool row = lane < PQ;
            const int la = row ? lane : 0;
            if (do_obs) {
                const double *Syv = gc + 2, *wy = gc + 2 + PQ, *svv_inv = gc + 2 + 2 * PQ;
                const double sxv = row ? o[7 + la] : 0.0, wya = row ? wy[la] : 0.0, syva = row ? Syv[la] : 0.0;
                double z = 0.0;
#pragma unroll
                for (int b2 = 0; b2 < PQ; b2++) z = fma(svv_inv[la * PQ + b2], o[7 + b2], z);
                if (!row) z = 0.0;
                const double num = o[0] - scan_warp_sum(wya * sxv);
                const double den = (o[1] + o[2]) - scan_warp_sum(sxv * z); // Sxx + sum V over observed steps (EM.cpp:152)
                const double Cn = num / den;
                const double d = fma(-Cn, z, wya);
                const double racc = fma(-Cn, o[0], gc[0]) - scan_warp_sum(d * syva);
                if (row) THS[2 + PQ + la] = d;
                if (lane == 0) {
                    THS[1 + PQ] = Cn;
                    THS[3 + 2 * PQ] = racc / n_obs;
                }
            }
            if (do_trans) {
                const double tx1u = row ? o[7 + PQ + la] : 0.0, tux = row ? o[7 + 2 * PQ + la] : 0.0;
                double z = 0.0, w2 = 0.0;
#pragma unroll
                for (int b2 = 0; b2 < PQ; b2++) {
                    const double m = tuu_inv[la * PQ + b2];
                    z = fma(m, o[7 + 2 * PQ + b2], z);
                    w2 = fma(m, o[7 + PQ + b2], w2);
                }
                if (!row) z = w2 = 0.0;
                const double Txx = o[5] + o[6], Tx1x = o[3] + o[4]; // EM.cpp:180,181
                const double num = Tx1x - scan_warp_sum(tx1u * z), den = Txx - scan_warp_sum(tux * z);
                const double An = num / den;
                // Tx1x1 = sum_{t=1}^{T-1} (X_t^2+V_t) = Txx - (X_0^2+V_0) + (X_{T-1}^2+V_{T-1})   (EM.cpp:181,183)
                const double Tx1x1 = Txx - fma(ENDS[0], ENDS[0], ENDS[1]) + fma(ENDS[2], ENDS[2], ENDS[3]);
                const double bb = fma(-An, z, w2);
                const double qacc = fma(-An, Tx1x, Tx1x1) - scan_warp_sum(bb * tx1u);
                if (row) THS[1 + la] = bb;
                if (lane == 0) {
                    THS[0] = An;
                    THS[2 + 2 * PQ] = qacc / (double)(T - 1);
                    THS[4 + 2 * PQ] = ENDS[0]; // EM.cpp:218-219
                    THS[5 + 2 * PQ] = ENDS[1];
                }
            }
        }
        }
        SCAN_MARK(12);
        // ---- likelihood (EM.cpp:122-124) and stop rule (EM.cpp:259-275): the last thread of the last warp, while
        //      threads 0 and 32 do the M-step -- off everybody's critical path (as part of every thread's own
        //      instruction stream the logarithm alone cost 230 cycles of an iteration of 5 100).  The M-step is
        //      speculative: when the fit has converged or is out of iterations, its theta is simply not loaded.
        if constexpr (!EMIT) {
            if (warp == nw - 1 && lane == 31) {
                double acc = 0.0, pmt = 1.0; // in warp order
                int pet = 0;
#pragma unroll
                for (int w = 0; w < SCAN_MAX_WARPS; ++w) {
                    acc += w < nw ? S3[w] : 0.0;
                    pmt *= w < nw ? S3M[w] : 1.0; // < 2^(32 SCAN_MAX_WARPS)
                    pet += w < nw ? S3E[w] : 0;
                }
                acc += fma((double)pet, 0.693147180559945309417, log_pos(pmt));
                const double lik_new = (-0.5 * n_obs * LOG_2PI - 0.5 * acc) * inv_n_obs;
                if (P.liks) P.liks[(size_t)P.f_user[fit] * P.niter + ne] = lik_new;
                const bool conv = (ne + 1 >= 3) && (fabs(lik_new - l1) < P.tol) && (fabs(l1 - l2) < P.tol);
                S_LIK[0] = lik_new;
                S_LIVE[0] = (conv || ne + 1 >= P.niter) ? 0 : 1;
            }
        }
        __syncthreads(); // B5: the new theta and the stop decision are published
        if constexpr (!EMIT) {
            lik = S_LIK[0];
            ne += 1;
            live = S_LIVE[0] != 0;
            if (!live) break; // theta stays as it entered this iteration
        }
        load_theta<PQ>(th, THS);
        l2 = l1;
        l1 = lik;
        SCAN_MARK(13);
#ifdef LDSR_PHASE_CLOCKS
        n_it_done++;
#endif
    }
#ifdef LDSR_PHASE_CLOCKS
    if (!EMIT && blockIdx.x == 0 && (threadIdx.x == 0 || threadIdx.x == 32) && n_it_done > 0) {
        static const char *names[14] = {"loop top", "Bu + P1 compose + scan", "B1 wait", "warp chain + Vin", "P2", "affine scan",
                                        "B2 wait + xin", "lik sum + backward scan", "B3 wait + lik + stop", "Xs1 + P4 + sums",
                                        "reduce", "B4 wait", "M-step", "B5 wait + theta"};
        for (int i = 0; i < 14; i++)
            printf("[scan clocks] thread %2d  %-26s %8.0f cycles/iteration\n", (int)threadIdx.x, names[i],
                   (double)pc[i] / n_it_done);
    }
#endif

    if (!EMIT && threadIdx.x == 0) {
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
