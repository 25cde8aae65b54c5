// em_split_kernel.cuh -- the LDS_EM loop (src/EM.cpp:245-280) with the TIME AXIS SPLIT ACROSS THE
// WARPS OF A CTA.  lane = fit (32 fits of one series per CTA), warp = a contiguous chunk of time.
//
// Why: with one lane per fit and one warp walking all T steps (em_kernel.cuh) a batch of N fits is
// only N/32 serial instruction streams -- 313 for the 10 000-fit cvLDS job, on a machine with 592
// SM sub-partitions -- and each stream is latency-bound.  Here every fit is advanced by NW warps at
// once, so the same batch is NW x more streams, each 1/NW as long.
//
// How the recursions are cut (all of it exact algebra on EM.cpp:70-90, 99-104, no approximation):
//  * The time axis is tiled into UNITS: a 32-step word no fit of the CTA observes (type U), or an
//    8-step segment (type M).  Warp w owns units [ub[w], ub[w+1]) (cost-balanced at launch).
//  * P1  each warp composes the VARIANCE map of its chunk.  In 1-D the Riccati step is a Moebius
//        map of Vp, i.e. a 2x2 matrix acting on homogeneous coordinates (n,d), Vp = n/d:
//          observed   [[A^2 R + Q C^2, Q R],[C^2, R]]      unobserved [[A^2, Q],[0, 1]]
//        (a U word is the closed form [[A^64, Q sum A^2k],[0,1]]).          -> barrier 1
//        Every warp then applies the maps of the chunks to its left to (V1,1): its incoming Vp.
//  * P2  forward over the chunk with the true variances.  The MEAN is carried as an affine
//        function of the (still unknown) incoming mean x_in:  Xp_t = P_t x_in + q_t, so the
//        innovations are affine and sum delta^2/Sigma is a quadratic (l0,l1,l2) in x_in.  In the
//        same sweep the chunk's BACKWARD map is composed: the RTS step is affine,
//        Xs_t = J_t Xs_{t+1} + g_t, Vs_t = J_t^2 Vs_{t+1} + L_t, and over a U word it telescopes
//        to J = A^32 Vp_first / Vp_last.  Checkpoints (Vp, q, P) per unit.      -> barrier 2
//        Every warp chains the (P,q) of all chunks -> x_in of every chunk, the likelihood and the
//        stop rule (EM.cpp:272; identical arithmetic in every warp, so no broadcast is needed),
//        and the backward maps of the chunks to its right -> smoothed state entering its chunk.
//  * P4  backward over the chunk, unit by unit: M segments are recomputed from their checkpoint
//        and smoothed as in em_kernel.cuh; U words are STREAMED FORWARD because inside a run of
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

constexpr int SPLIT_NCH = 10; // per-chunk values exchanged after P2

// number of M-step partial sums a warp publishes
template <int PQ> __host__ __device__ constexpr int split_nstat() { return 11 + 3 * PQ; }

// dynamic shared memory of em_split_kernel, in bytes, after the series blob
__host__ __device__ inline size_t split_smem_bytes(int pq, int nw, int max_units) {
    size_t b = 0;
    b += (size_t)max_units * 3 * 32 * 8;             // checkpoints
    b += (size_t)nw * 4 * 32 * 8;                    // variance maps
    b += (size_t)nw * SPLIT_NCH * 32 * 8;            // chunk summaries
    b += (size_t)nw * (11 + 3 * pq) * 32 * 8;        // partial sums
    b += ((size_t)max_units * 4 + 15) & ~size_t(15); // unit table
    b += 64;                                         // chunk bounds
    return b;
}

// upper bound on the number of units of a series from its finite(y) mask (hold-outs only remove
// observations, and a word with no observation is one unit instead of four)
inline int split_units_upper_bound(const double *y, int T) {
    int n = 0;
    for (int w = 0; w * 32 < T; w++) {
        bool any = false;
        for (int t = w * 32; t < T && t < w * 32 + 32; t++) any = any || (y[t] == y[t]);
        const bool inside = 32 * (w + 1) <= T - 1;
        if (!any && inside)
            n += 1;
        else
            n += (std::min(T, w * 32 + 32) - w * 32 + 7) / 8;
    }
    return n;
}

__device__ __forceinline__ void rescale4(double &a, double &b, double &c, double &d) {
    const int e = ((__double2hiint(a + b + c + d) >> 20) & 0x7ff) - 1023;
    const double sc = __hiloint2double((1023 - e) << 20, 0);
    a *= sc;
    b *= sc;
    c *= sc;
    d *= sc;
}

// per-iteration constants of a fit
template <int PQ> struct SplitConst {
    double A, A2, Q;
    double Ap[9];            // A^0 .. A^8
    double A16, A32;         // A^16, A^32
    double aV32, bV32;       // Vp' = aV32 Vp + bV32 over an unobserved word
    MixedConst<PQ> mc;
    __device__ __forceinline__ void set(const Theta<PQ> &th) {
        A = th.A;
        A2 = A * A;
        Q = th.Q;
        Ap[0] = 1.0;
#pragma unroll
        for (int k = 1; k <= 8; k++) Ap[k] = Ap[k - 1] * A;
        A16 = Ap[8] * Ap[8];
        A32 = A16 * A16;
        aV32 = A32 * A32;
        double sv = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) sv = fma(sv, A2, 1.0); // sum_{k<8} A2^k
        // sum_{k<32} A2^k = sv (1 + A2^8)(1 + A2^16),  A2^8 = A^16
        bV32 = Q * (sv * ((1.0 + A16) * (1.0 + A32)));
        mc.set(th, A2);
    }
};

// ---- P1: variance map of one M segment, M <- S_j ... S_0 M ----------------------------------
template <int PQ>
__device__ __forceinline__ void compose_var_segment(const Theta<PQ> &th, const SplitConst<PQ> &k, unsigned bits,
                                                    int cnt, double &m11, double &m12, double &m21, double &m22) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (j < cnt) {
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
    }
    rescale4(m11, m12, m21, m22);
}

// chunk state carried through P2
struct ChunkFwd {
    double Vq, P, q;            // prior variance; prior mean = P x_in + q
    double l0, l1, l2;          // sum_obs delta^2/Sigma = l0 - 2 C x l1 + C^2 x^2 l2
    double dprod;               // product of the segments' final d (sum_obs log Sigma = log dprod + shift ln 2)
    int shift;
    double PJ, PJ2, G0, GG, Lc; // backward map: Xs_first = PJ Xs_in + G0 + GG x_in, Vs_first = PJ2 Vs_in + Lc
};

// ---- P2 over one M segment --------------------------------------------------------------------
// Same recursion as mixed_forward (em_kernel.cuh) with the mean in (P,q) form and the backward map
// accumulated on the fly.
template <int PQ, bool GUARDED>
__device__ __forceinline__ void forward_segment_basis(const Theta<PQ> &th, const SplitConst<PQ> &k, unsigned bits,
                                                      int cnt, const double *__restrict__ yseg,
                                                      const double *__restrict__ useg,
                                                      const double *__restrict__ vseg, ChunkFwd &c) {
    double n = c.Vq, d = 1.0;
    int shift = 0;
    const double A = k.A;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (!GUARDED || j < cnt) {
            const bool obs = (bits >> j) & 1u;
            const double m11 = obs ? k.mc.a11 : k.A2, m12 = obs ? k.mc.a12 : k.Q;
            const double m21 = obs ? k.mc.C2 : 0.0, m22 = obs ? th.R : 1.0;
            const double nn = fma(m11, n, m12 * d);
            const double dd = fma(m21, n, m22 * d);
            const double rho = fast_rcp(dd);
            const double Bu = dot_row<PQ>(th.B, useg + j * PQ);
            const double Dv = dot_row<PQ>(th.D, vseg + j * PQ);
            const double K = obs ? th.C * n * rho : 0.0;
            const double rS = obs ? d * rho : 0.0; // 1/Sigma
            const double ymd = (obs ? yseg[j] : 0.0) - Dv;
            const double d0 = fma(-th.C, c.q, ymd); // innovation for x_in = 0
            const double w0 = rS * d0;
            c.l0 = fma(w0, d0, c.l0);
            c.l1 = fma(w0, c.P, c.l1);
            c.l2 = fma(rS * c.P, c.P, c.l2);
            const double alpha = fma(-k.mc.AC, K, A);
            const double beta = fma(A * K, ymd, Bu);
            const double xu0 = fma(K, d0, c.q);
            const double xuP = c.P * fma(-K, th.C, 1.0);
            const double qn = fma(alpha, c.q, beta);
            const double Pn = alpha * c.P;
            const double nu = obs ? th.R * n : n;
            const double vu = nu * rho;
            double J, g0, gP, L;
            if (GUARDED && j == cnt - 1) { // t == T-1: smoothed = filtered (EM.cpp:94-95)
                J = 0.0;
                g0 = xu0;
                gP = xuP;
                L = vu;
            } else {
                J = A * nu * fast_rcp(nn);
                g0 = fma(-J, qn, xu0);
                gP = fma(-J, Pn, xuP);
                L = vu * fma(-A, J, 1.0);
            }
            c.G0 = fma(c.PJ, g0, c.G0);
            c.GG = fma(c.PJ, gP, c.GG);
            c.Lc = fma(c.PJ2, L, c.Lc);
            c.PJ *= J;
            c.PJ2 *= J * J;
            c.q = qn;
            c.P = Pn;
            if (j == (GUARDED ? cnt - 1 : 7)) c.Vq = nn * rho;
            n = nn;
            d = dd;
            if (j == 3) rescale_pow2(n, d, shift);
        }
    }
    c.dprod *= d;
    {
        const int e = ((__double2hiint(c.dprod) >> 20) & 0x7ff) - 1023;
        c.dprod *= __hiloint2double((1023 - e) << 20, 0);
        c.shift += shift + e;
    }
}

// ---- P2 over one unobserved 32-step word ------------------------------------------------------
template <int PQ>
__device__ __forceinline__ void forward_word_basis(const Theta<PQ> &th, const SplitConst<PQ> &k,
                                                   const double *__restrict__ useg, ChunkFwd &c) {
    double h[4];
#pragma unroll
    for (int b = 0; b < 4; b++) {
        double hb = 0.0;
#pragma unroll
        for (int j = 0; j < 8; j++) hb = fma(k.A, hb, dot_row<PQ>(th.B, useg + (b * 8 + j) * PQ));
        h[b] = hb;
    }
    const double A8 = k.Ap[8];
    const double hh = fma(fma(fma(h[0], A8, h[1]), A8, h[2]), A8, h[3]);
    const double qn = fma(k.A32, c.q, hh);
    const double Pn = k.A32 * c.P;
    const double Vn = fma(k.aV32, c.Vq, k.bV32);
    // backward map of the word: J = prod J_t = A^32 Vp_first / Vp_last  (EM.cpp:100 telescoped)
    const double Jc = k.A32 * c.Vq * fast_rcp(Vn);
    const double g0 = fma(-Jc, qn, c.q);
    const double gP = fma(-Jc, Pn, c.P);
    const double L = c.Vq * fma(-k.A32, Jc, 1.0);
    c.G0 = fma(c.PJ, g0, c.G0);
    c.GG = fma(c.PJ, gP, c.GG);
    c.Lc = fma(c.PJ2, L, c.Lc);
    c.PJ *= Jc;
    c.PJ2 *= Jc * Jc;
    c.q = qn;
    c.P = Pn;
    c.Vq = Vn;
}

// ---- P4 over one unobserved 32-step word ------------------------------------------------------
// (Xq,Vq): prior at the first step of the word.  (cG,cH): the run constants AT THE RIGHT END of the
// word; on return they are the constants at its left end (= right end of the word before it).
template <int PQ>
__device__ __forceinline__ void smooth_word(const Theta<PQ> &th, const SplitConst<PQ> &k,
                                            const double *__restrict__ useg, double Xq, double Vq, double &cG,
                                            double &cH, double &Xs1, double &Vs1, Stats<PQ> &st) {
    const double A8 = k.Ap[8];
    double Gb[4], Hb[4]; // constants at the right end of each 8-step block
    Gb[3] = cG;
    Gb[2] = A8 * cG;
    Gb[1] = k.A16 * cG;
    Gb[0] = A8 * Gb[1];
    Hb[3] = cH;
    Hb[2] = k.A16 * cH;
    Hb[1] = k.A32 * cH;
    Hb[0] = k.A16 * Hb[1];
    const double G0 = A8 * Gb[0], H0 = k.A16 * Hb[0];
    double xp = Xq, vp = Vq;
    double Xs = fma(vp, G0, xp);
    double Vs = fma(vp, vp * H0, vp);
    const double Xfirst = Xs, Vfirst = Vs;
    double tv = 0.0;
#pragma unroll
    for (int b = 0; b < 4; b++) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double *__restrict__ row = useg + (b * 8 + j) * PQ;
            double uk[PQ];
#pragma unroll
            for (int i = 0; i < PQ; i++) uk[i] = row[i];
            double Bu = 0.0;
#pragma unroll
            for (int i = 0; i < PQ; i++) Bu = fma(th.B[i], uk[i], Bu);
            const double xpn = fma(k.A, xp, Bu);
            const double vpn = fma(k.A2, vp, k.Q);
            const double pw = k.Ap[7 - j];
            const double Gn = pw * Gb[b];
            const double Hn = (pw * pw) * Hb[b];
            const double Xsn = fma(vpn, Gn, xpn);
            const double t1 = vpn * Hn;
            const double Vsn = fma(vpn, t1, vpn);
            st.Tx1x = fma(Xsn, Xs, st.Tx1x);
            st.Txx = fma(Xs, Xs, st.Txx);
            st.Txxv += Vs;
            tv = fma(vp, 1.0 + t1, tv); // V_{t+1} J_t = A Vp_t (1 + Vp_{t+1} H_{t+1})
#pragma unroll
            for (int i = 0; i < PQ; i++) {
                st.Tx1u[i] = fma(Xsn, uk[i], st.Tx1u[i]);
                st.Tux[i] = fma(uk[i], Xs, st.Tux[i]);
            }
            xp = xpn;
            vp = vpn;
            Xs = Xsn;
            Vs = Vsn;
        }
    }
    st.Tx1xv = fma(k.A, tv, st.Tx1xv);
    Xs1 = Xfirst;
    Vs1 = Vfirst;
    cG = G0;
    cH = H0;
}

struct SplitParams {
    EmParams em;
    int max_units;   // capacity of the unit table / checkpoint area
    int blob_smem;   // bytes reserved for the series blob at the start of dynamic shared memory
    int cost_u, cost_m; // relative cost of a U word and an M segment (chunk balancing)
};

constexpr int UNIT_M = 1 << 30;

template <int PQ, int NW, int MINB>
__global__ void __launch_bounds__(NW * 32, MINB) em_split_kernel(const SplitParams SP) {
    const EmParams &P = SP.em;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    constexpr int NST = split_nstat<PQ>();

    const int4 task = P.tasks[blockIdx.x];
    const SeriesDev S = P.series[task.x];
    const int T = S.T;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) stage_blob(smem_raw, P.blobs + S.blob_off, (unsigned)S.blob_doubles * 8u, &bar);

    const double *__restrict__ ser = reinterpret_cast<const double *>(smem_raw);
    const double *__restrict__ ys = ser + S.y_off;
    const double *__restrict__ us = ser + S.u_off;
    const double *__restrict__ vs = ser + S.v_off;
    // shared-memory carve-up after the blob; every per-lane array is [..][32] doubles
    double *const ck = reinterpret_cast<double *>(smem_raw + SP.blob_smem) + lane;   // [unit][3]
    double *const MC = ck - lane + (size_t)SP.max_units * 96 + lane;                  // [NW][4]
    double *const CH = MC - lane + (size_t)NW * 4 * 32 + lane;                        // [NW][SPLIT_NCH]
    double *const ST = CH - lane + (size_t)NW * SPLIT_NCH * 32 + lane;                // [NW][NST]
    int *const units = reinterpret_cast<int *>(ST - lane + (size_t)NW * NST * 32);    // [max_units]
    int *const ubound = units + ((SP.max_units + 3) & ~3);                            // [NW+1]

    // ---- per-lane fit state: every warp holds the same 32 fits
    const bool valid = lane < task.z;
    const int fit = P.active[task.y + (valid ? lane : 0)];
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

    // ---- unit table and chunk bounds (warp 0; the vote is over the CTA's 32 fits)
    if (warp == 0) {
        int nu = 0, cost = 0;
        for (int w = 0; w * 32 < T; ++w) {
            const bool any = __any_sync(FULL, mw[w] != 0u);
            const bool inside = 32 * (w + 1) <= T - 1;
            if (!any && inside) {
                if (lane == 0) units[nu] = w * 32;
                nu++;
                cost += SP.cost_u;
            } else {
                for (int t0 = w * 32; t0 < T && t0 < w * 32 + 32; t0 += 8) {
                    if (lane == 0) units[nu] = t0 | UNIT_M;
                    nu++;
                    cost += SP.cost_m;
                }
            }
        }
        __syncwarp();
        if (lane == 0) {
            ubound[0] = 0;
            int acc = 0, kq = 0;
            for (int w = 1; w < NW; ++w) {
                const int target = (int)(((long long)cost * w + NW / 2) / NW);
                while (kq < nu) {
                    const int cu = (units[kq] & UNIT_M) ? SP.cost_m : SP.cost_u;
                    if (acc + cu / 2 >= target) break;
                    acc += cu;
                    kq++;
                }
                ubound[w] = kq;
            }
            ubound[NW] = nu;
        }
    }
    mbar_wait(&bar, 0);
    __syncthreads();
    const int ua = ubound[warp], ue = ubound[warp + 1];

    for (int it = 0; it < P.chunk; ++it) {
        if (!__any_sync(FULL, live)) break;
        SplitConst<PQ> k;
        k.set(th);

        // ================= P1: variance map of the chunk =================
        if (warp < NW - 1) { // nobody is to the right of the last chunk
            double m11 = 1.0, m12 = 0.0, m21 = 0.0, m22 = 1.0;
            for (int un = ua; un < ue; ++un) {
                const int u0 = units[un];
                if (u0 & UNIT_M) {
                    const int t0 = u0 & (UNIT_M - 1);
                    compose_var_segment<PQ>(th, k, seg_bits(mw, t0, 8), min(8, T - t0), m11, m12, m21, m22);
                } else {
                    m11 = fma(k.aV32, m11, k.bV32 * m21);
                    m12 = fma(k.aV32, m12, k.bV32 * m22);
                }
            }
            MC[(warp * 4 + 0) * 32] = m11;
            MC[(warp * 4 + 1) * 32] = m12;
            MC[(warp * 4 + 2) * 32] = m21;
            MC[(warp * 4 + 3) * 32] = m22;
        }
        __syncthreads();
        ChunkFwd c;
        {
            double n = th.V1, d = 1.0;
            for (int w = 0; w < warp; ++w) {
                const double m11 = MC[(w * 4 + 0) * 32], m12 = MC[(w * 4 + 1) * 32];
                const double m21 = MC[(w * 4 + 2) * 32], m22 = MC[(w * 4 + 3) * 32];
                const double nn = fma(m11, n, m12 * d), dd = fma(m21, n, m22 * d);
                n = nn;
                d = dd;
                int sh = 0;
                rescale_pow2(n, d, sh);
            }
            c.Vq = warp == 0 ? th.V1 : n * fast_rcp(d);
        }

        // ================= P2: forward over the chunk =================
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
            const int t0 = u0 & (UNIT_M - 1);
            ck[(un * 3 + 0) * 32] = c.Vq;
            ck[(un * 3 + 1) * 32] = c.q;
            ck[(un * 3 + 2) * 32] = c.P;
            if (u0 & UNIT_M) {
                any_m = true;
                if (t0 + 8 >= T)
                    forward_segment_basis<PQ, true>(th, k, seg_bits(mw, t0, 8), T - t0, ys + t0, us + t0 * PQ,
                                                    vs + t0 * PQ, c);
                else
                    forward_segment_basis<PQ, false>(th, k, seg_bits(mw, t0, 8), 8, ys + t0, us + t0 * PQ,
                                                     vs + t0 * PQ, c);
            } else {
                forward_word_basis<PQ>(th, k, us + t0 * PQ, c);
            }
        }
        {
            double ld = 0.0;
            if (any_m) ld = fma((double)c.shift, 0.693147180559945309417, log(c.dprod));
            double *o = CH + (size_t)warp * SPLIT_NCH * 32;
            o[0 * 32] = c.P;
            o[1 * 32] = c.q;
            o[2 * 32] = c.l0;
            o[3 * 32] = c.l1;
            o[4 * 32] = c.l2;
            o[5 * 32] = ld;
            o[6 * 32] = c.PJ;
            o[7 * 32] = c.G0;
            o[8 * 32] = c.GG;
            o[9 * 32] = c.Lc;
        }
        __syncthreads();

        // ---- chain the chunks: x_in of every chunk, likelihood (identical in every warp)
        double xk[NW];
        double acc = 0.0;
        {
            double x = th.mu1; // prior of step 0 (EM.cpp:48)
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const double *o = CH + (size_t)w * SPLIT_NCH * 32;
                xk[w] = x;
                const double tC = th.C * x;
                acc += fma(tC, fma(tC, o[4 * 32], -2.0 * o[3 * 32]), o[2 * 32]) + o[5 * 32];
                x = fma(o[0 * 32], x, o[1 * 32]);
            }
        }
        const double lik_new = (-0.5 * n_obs * LOG_2PI - 0.5 * acc) / n_obs; // EM.cpp:122-124

        // ================= stop rule (EM.cpp:259-275) =================
        if (live) {
            lik = lik_new;
            ne += 1;
            if (warp == 0 && P.liks) P.liks[(size_t)P.f_user[fit] * P.niter + (ne - 1)] = lik_new;
            const bool conv = (ne >= 3) && (fabs(lik_new - l1) < P.tol) && (fabs(l1 - l2) < P.tol);
            if (conv || ne >= P.niter) live = false;
        }
        if (!__any_sync(FULL, live)) break;

        // ---- smoothed state entering the chunk from the right
        double xin = 0.0, Xs1 = 0.0, Vs1 = 0.0;
#pragma unroll
        for (int w = NW - 1; w >= 0; --w) {
            if (w == warp) xin = xk[w];
            if (w > warp) {
                const double *o = CH + (size_t)w * SPLIT_NCH * 32;
                const double pj = o[6 * 32];
                Xs1 = fma(pj, Xs1, fma(o[8 * 32], xk[w], o[7 * 32]));
                Vs1 = fma(pj * pj, Vs1, o[9 * 32]);
            }
        }

        // ================= P4: backward over the chunk, M-step sums =================
        Stats<PQ> st;
        st.zero();
        {
            double Xr = fma(c.P, xin, c.q), Vr = c.Vq; // prior at the first step right of the chunk
            double cG = 0.0, cH = 0.0;
            bool in_run = false;
            for (int un = ue - 1; un >= ua; --un) {
                const int u0 = units[un];
                const int t0 = u0 & (UNIT_M - 1);
                const double Vq = ck[(un * 3 + 0) * 32];
                const double Xq = fma(ck[(un * 3 + 2) * 32], xin, ck[(un * 3 + 1) * 32]);
                if (u0 & UNIT_M) {
                    if (t0 + 8 >= T)
                        smooth_segment<PQ, 8, true, true>(th, k.A, k.A2, k.Q, k.mc, seg_bits(mw, t0, 8), T - t0,
                                                          ys + t0, us + t0 * PQ, vs + t0 * PQ, Xq, Vq, Xs1, Vs1, st);
                    else
                        smooth_segment<PQ, 8, true, false>(th, k.A, k.A2, k.Q, k.mc, seg_bits(mw, t0, 8), 8, ys + t0,
                                                           us + t0 * PQ, vs + t0 * PQ, Xq, Vq, Xs1, Vs1, st);
                    in_run = false;
                } else {
                    if (!in_run) {
                        const double rv = fast_rcp(Vr);
                        cG = (Xs1 - Xr) * rv;
                        cH = (Vs1 - Vr) * rv * rv;
                        in_run = true;
                    }
                    smooth_word<PQ>(th, k, us + t0 * PQ, Xq, Vq, cG, cH, Xs1, Vs1, st);
                }
                Xr = Xq;
                Vr = Vq;
            }
            if (ua == 0 && ue > 0) {
                st.X0 = Xs1;
                st.V0 = Vs1;
            }
        }
        {
            double *o = ST + (size_t)warp * NST * 32;
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
        __syncthreads();

        // ================= M-step (EM.cpp:139-229), same arithmetic in every warp =================
        st.zero();
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const double *o = ST + (size_t)w * NST * 32;
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
        if (live) {
            mstep_from_stats<PQ>(st, gc, tuu_inv, T, th);
            l2 = l1;
            l1 = lik;
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
}

} // namespace ldsr
