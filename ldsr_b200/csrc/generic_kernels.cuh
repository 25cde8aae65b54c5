// generic_kernels.cuh -- kernels that do not depend on the padded width PQ at compile time:
// set-up (hold-out masks, theta-independent Gram blocks and their inverses), live-fit compaction,
// CTA task list, restart selection, transposition.  Included by ldsr_abi.cu only.
#pragma once
#include "lds_math.cuh"

namespace ldsr {

// ------------------------------------------------------------------------------------------
// In-place inverse of a symmetric positive-definite n x n matrix held in shared memory with
// leading dimension ld, by Cholesky (one thread).  Returns false if a pivot is not > 0.
// Stands in for the Gram-block part of arma::inv(P2)/inv(P4) (EM.cpp:166,198).
// ------------------------------------------------------------------------------------------
__device__ inline bool spd_inverse(double *a, int n, int ld, double *work /* n*n */) {
    // a = L L'
    for (int j = 0; j < n; j++) {
        double d = a[j * ld + j];
        for (int k = 0; k < j; k++) d -= a[j * ld + k] * a[j * ld + k];
        if (!(d > 0.0) || !isfinite(d)) return false;
        d = sqrt(d);
        a[j * ld + j] = d;
        for (int i = j + 1; i < n; i++) {
            double s = a[i * ld + j];
            for (int k = 0; k < j; k++) s -= a[i * ld + k] * a[j * ld + k];
            a[i * ld + j] = s / d;
        }
    }
    // work = inv(L) (lower)
    for (int j = 0; j < n; j++) {
        for (int i = 0; i < n; i++) work[i * n + j] = 0.0;
        work[j * n + j] = 1.0 / a[j * ld + j];
        for (int i = j + 1; i < n; i++) {
            double s = 0.0;
            for (int k = j; k < i; k++) s -= a[i * ld + k] * work[k * n + j];
            work[i * n + j] = s / a[i * ld + i];
        }
    }
    // a = inv(L)' inv(L)
    for (int i = 0; i < n; i++)
        for (int j = 0; j <= i; j++) {
            double s = 0.0;
            for (int k = i; k < n; k++) s += work[k * n + i] * work[k * n + j];
            a[i * ld + j] = s;
            a[j * ld + i] = s;
        }
    return true;
}

struct SetupParams {
    const SeriesDev *series;
    const double *blobs;
    int n_series, n_groups;
    const int *g_series;
    const int *held_ptr; // may be null
    const int *held_idx;
    unsigned *masks;
    const long long *g_mask_off;
    double *gconst;  // per group
    double *sconst;  // per series (at SeriesDev.sconst_off)
    int *g_status;
    int *g_nobs;
    int pq;
};

// One block per group (blockIdx < n_groups) or per series (the rest).
//  group : observed-bit mask = finite(y) minus the hold-out list (R/LDS_reconstruction.R:274);
//          Syy, Syv, Svv over the observed steps (EM.cpp:158,161), SvvInv, wy = SvvInv Syv.
//  series: Tuu = sum_{t<T-1} u u' (EM.cpp:193) and its inverse.
// Dynamic shared memory: (2*pq*pq + 2*pq + 4) doubles.
__global__ void setup_kernel(const SetupParams P) {
    extern __shared__ __align__(16) double sh[];
    const int pq = P.pq;
    double *M = sh, *work = sh + pq * pq, *vec = work + pq * pq, *wy = vec + pq;
    __shared__ int ok_flag;
    const bool is_group = (int)blockIdx.x < P.n_groups;
    const int g = blockIdx.x, s = is_group ? P.g_series[g] : (int)blockIdx.x - P.n_groups;
    const SeriesDev S = P.series[s];
    const double *ys = P.blobs + S.blob_off + S.y_off;
    const double *rows = P.blobs + S.blob_off + (is_group ? S.v_off : S.u_off);
    const int T = S.T, nwords = (T + 31) / 32;
    const int n_real = is_group ? (S.has_v ? S.q : 0) : (S.has_u ? S.p : 0);
    unsigned *mw = is_group ? P.masks + P.g_mask_off[g] : nullptr;

    if (is_group) {
        for (int w = threadIdx.x; w < nwords; w += blockDim.x) {
            unsigned bits = 0;
            for (int b = 0; b < 32; b++) {
                const int t = w * 32 + b;
                if (t < T && isfinite(ys[t])) bits |= 1u << b;
            }
            mw[w] = bits;
        }
        __syncthreads();
        if (P.held_ptr)
            for (int k = P.held_ptr[g] + threadIdx.x; k < P.held_ptr[g + 1]; k += blockDim.x) {
                const int t = P.held_idx[k];
                atomicAnd(&mw[t >> 5], ~(1u << (t & 31)));
            }
        __syncthreads();
    }
    // Gram entries: thread e handles (a,b) of the pq x pq block; thread pq*pq+a handles Syv[a];
    // thread pq*pq+pq handles Syy and n_obs.  Each entry is summed over t IN ORDER, as the reference does
    // (EM.cpp:158,161): cutting the time axis into slices added in a fixed order is 0.1 ms faster per call but
    // changes the constants in the last bit, which 1000 EM iterations on a flat likelihood amplify to 1.2e-6 in one
    // theta entry of test_cv_np413_sample (bar: 1e-6) -- measured and not kept.
    const int n_ent = pq * pq + pq + 1;
    for (int e = threadIdx.x; e < n_ent; e += blockDim.x) {
        double acc = 0.0;
        int cnt = 0;
        const int a = e < pq * pq ? e / pq : e - pq * pq, b = e % pq;
        const int t_end = is_group ? T : T - 1;
        for (int t = 0; t < t_end; t++) {
            if (is_group && !((mw[t >> 5] >> (t & 31)) & 1u)) continue;
            if (e < pq * pq)
                acc = fma(rows[(size_t)t * pq + a], rows[(size_t)t * pq + b], acc);
            else if (e < pq * pq + pq)
                acc = is_group ? fma(ys[t], rows[(size_t)t * pq + a], acc) : 0.0;
            else {
                acc = is_group ? fma(ys[t], ys[t], acc) : 0.0;
                cnt++;
            }
        }
        if (e < pq * pq)
            M[e] = acc;
        else if (e < pq * pq + pq)
            vec[a] = acc;
        else {
            wy[pq] = acc;             // Syy
            wy[pq + 1] = (double)cnt; // n_obs
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        bool ok = true;
        if (n_real > 0) ok = spd_inverse(M, n_real, pq, work);
        // rows/cols beyond the real width belong to zero padding: their inverse block is zero
        for (int a = 0; a < pq; a++)
            for (int b = 0; b < pq; b++)
                if (a >= n_real || b >= n_real) M[a * pq + b] = 0.0;
        ok_flag = ok ? 1 : 0;
    }
    __syncthreads();
    if (is_group) {
        double *gc = P.gconst + (size_t)g * gconst_stride(pq);
        for (int a = threadIdx.x; a < pq; a += blockDim.x) {
            double acc = 0.0;
            for (int b = 0; b < pq; b++) acc = fma(M[a * pq + b], vec[b], acc);
            gc[2 + a] = vec[a];
            gc[2 + pq + a] = acc;
        }
        for (int e = threadIdx.x; e < pq * pq; e += blockDim.x) gc[2 + 2 * pq + e] = M[e];
        if (threadIdx.x == 0) {
            gc[0] = wy[pq];
            gc[1] = wy[pq + 1];
            P.g_nobs[g] = (int)wy[pq + 1];
            P.g_status[g] = ok_flag ? 0 : 1;
        }
    } else {
        double *sc = P.sconst + S.sconst_off;
        for (int e = threadIdx.x; e < pq * pq; e += blockDim.x) sc[e] = M[e];
        if (threadIdx.x == 0) sc[pq * pq] = ok_flag ? 0.0 : 1.0; // series-level singular flag
    }
}

// A series whose Tuu is singular makes every group of it singular.
__global__ void merge_status_kernel(const SeriesDev *series, const double *sconst, const int *g_series, int n_groups,
                                    int pq, int *g_status) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const SeriesDev S = series[g_series[g]];
    if (sconst[S.sconst_off + pq * pq] != 0.0) g_status[g] = 1;
}

// ------------------------------------------------------------------------------------------
// Restart selection, one warp per group (R/LDS_reconstruction.R:50-58): among fits with C > 0
// the first with the largest non-NaN lik; if no fit has C > 0, the first largest non-NaN lik.
// Also writes the per-fit status.  Lanes stride over the group's fits; "first" is kept by breaking
// likelihood ties towards the lower fit index in the lane merge.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void select_merge(double &l, int &f, double lo, int fo) {
    // (lo, fo) replaces (l, f) when it is a valid candidate that is larger, or equal and earlier
    if (fo >= 0 && (f < 0 || lo > l || (lo == l && fo < f))) {
        l = lo;
        f = fo;
    }
}
__global__ void select_kernel(int n_groups, const int *g_fit_ptr, const double *theta, int theta_len, int c_index,
                              const double *lik, const int *g_status, int *best, int *f_status) {
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (g >= n_groups) return; // whole warps leave together
    const int gs = g_status[g];
    double lpos = 0.0, lall = 0.0;
    int bpos = -1, ball = -1;
    bool any_pos = false;
    for (int f = g_fit_ptr[g] + lane; f < g_fit_ptr[g + 1]; f += 32) {
        const double l = lik[f], c = theta[(size_t)f * theta_len + c_index];
        f_status[f] = gs ? 1 : (isfinite(l) ? 0 : 2);
        // posC is decided on C alone, NaN liks included (which(allC > 0))
        if (c > 0.0) any_pos = true;
        if (l != l) continue;
        if (c > 0.0 && (bpos < 0 || l > lpos)) {
            lpos = l;
            bpos = f;
        }
        if (ball < 0 || l > lall) {
            lall = l;
            ball = f;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        select_merge(lpos, bpos, __shfl_xor_sync(FULL, lpos, o), __shfl_xor_sync(FULL, bpos, o));
        select_merge(lall, ball, __shfl_xor_sync(FULL, lall, o), __shfl_xor_sync(FULL, ball, o));
    }
    any_pos = __any_sync(FULL, any_pos);
    if (lane == 0) best[g] = any_pos ? bpos : ball;
}

// ---- compaction of live fits, per series (fits of a series are contiguous) --------------------
// One block per series: writes the ids of its live fits to active[fit_begin ..) in order and the
// count to n_live[series].  Each thread owns a contiguous run of fits (count, block scan, write).
// The last block to finish turns the per-series counts into the CTA task list:
// counts[0] = number of CTA tasks, counts[1] = number of live fits; *ticket is left at 0 again.
// With share_ctl (time-split kernel, ranked task assignment): the control block is zeroed and order[] lists the
// tasks by progress, most advanced first (key: E-steps done by the task's least advanced fit; ties in task order).
// The keys are gathered while compacting (ne[f] is loaded next to done[f]; a shared-memory atomicMin per live fit)
// and handed to the last block through task_key[fit_begin/fits_per_cta + series + local task] -- slots of different series
// never collide and stay below n_fits/32 + n_series.
__global__ void compact_kernel(const SeriesDev *series, int n_series, const int *done, int *active, int *n_live,
                               int fits_per_cta, int4 *tasks, int *task_off, int *counts, unsigned *ticket,
                               const int *ne = nullptr, int *share_ctl = nullptr, int *order = nullptr,
                               int *task_key = nullptr) {
    __shared__ int warp_sums[32];
    __shared__ int keys[1024];
    __shared__ bool last;
    const SeriesDev S = series[blockIdx.x];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int n = S.fit_end - S.fit_begin, per = (n + blockDim.x - 1) / blockDim.x;
    const int lo = S.fit_begin + min(n, (int)threadIdx.x * per), hi = S.fit_begin + min(n, ((int)threadIdx.x + 1) * per);
    const bool rank_tasks = share_ctl != nullptr && (n + fits_per_cta - 1) / fits_per_cta <= 1024 && blockDim.x == 1024;
    if (rank_tasks) keys[threadIdx.x] = 0x7fffffff;
    int mine = 0;
    for (int f = lo; f < hi; f++) mine += done[f] == 0;
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    int off = incl - mine, total = 0;
    for (int w = 0; w < nw; w++) {
        if (w < warp) off += warp_sums[w];
        total += warp_sums[w];
    }
    for (int f = lo; f < hi; f++)
        if (done[f] == 0) {
            if (rank_tasks) atomicMin(&keys[off / fits_per_cta], ne[f]);
            active[S.fit_begin + off++] = f;
        }
    if (rank_tasks) {
        __syncthreads();
        const int my_tasks = (total + fits_per_cta - 1) / fits_per_cta;
        if ((int)threadIdx.x < my_tasks) task_key[S.fit_begin / fits_per_cta + blockIdx.x + threadIdx.x] = keys[threadIdx.x];
    }
    __syncthreads(); // every thread's writes, before thread 0's fence and ticket
    if (threadIdx.x == 0) {
        n_live[blockIdx.x] = total;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    if (threadIdx.x == 0) {
        int nt = 0, nl = 0;
        for (int s = 0; s < n_series; s++) {
            const int live = ((volatile int *)n_live)[s];
            task_off[s] = nt;
            nt += (live + fits_per_cta - 1) / fits_per_cta;
            nl += live;
        }
        task_off[n_series] = nt;
        counts[0] = nt;
        counts[1] = nl;
        *ticket = 0;
    }
    __syncthreads();
    for (int s = 0; s < n_series; s++) {
        const int live = ((volatile int *)n_live)[s], t0 = task_off[s], nt = task_off[s + 1] - t0;
        for (int i = threadIdx.x; i < nt; i += blockDim.x) {
            const int first = i * fits_per_cta;
            tasks[t0 + i] = make_int4(s, series[s].fit_begin + first, min(fits_per_cta, live - first), 0);
        }
    }
    if (!share_ctl) return;
    for (int i = threadIdx.x; i < SHARE_CTL_LEN; i += blockDim.x) share_ctl[i] = 0;
    __syncthreads(); // the task list of this block's threads
    const int nt = task_off[n_series];
    if (nt > 1024 || blockDim.x < 1024) return; // more tasks than any co-resident grid: the ranked mode is off
    const int t = threadIdx.x;
    if (t < nt) {
        const int s = tasks[t].x;
        keys[t] = __ldcg(task_key + series[s].fit_begin / fits_per_cta + s + (t - task_off[s]));
    }
    __syncthreads();
    if (t < nt) {
        int rank = 0;
        const int mine_key = keys[t];
        for (int j = 0; j < nt; j++) rank += (keys[j] > mine_key) || (keys[j] == mine_key && j < t);
        order[rank] = t;
    }
}


// ---- results in the caller's order and layout, ready for one device-to-host copy ---------------
// theta rows are unpadded to the caller's (p, q) (B / D come back as zeros when the series has no
// u / v, EM.cpp:186,154); fit-indexed outputs move from internal to caller order; best[] is mapped
// to caller fit ids.  One thread per fit, the first n_groups threads also do a group.
struct PackParams {
    int n_fits, n_groups, theta_len, pq, stride; // theta_len = 2 pq + 6 (device), stride = caller's row length
    const SeriesDev *series;
    const int *g_series, *f_group, *f_user, *g_user;
    const double *theta, *lik;
    const int *iters, *status, *best;
    double *theta_u, *lik_u;
    int *iters_u, *status_u, *best_u;
};
__global__ void pack_results_kernel(const PackParams P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.n_fits) {
        const SeriesDev S = P.series[P.g_series[P.f_group[i]]];
        const int fu = P.f_user[i];
        const double *src = P.theta + (size_t)i * P.theta_len;
        double *dst = P.theta_u + (size_t)fu * P.stride;
        dst[0] = src[0];
        for (int j = 0; j < S.p; j++) dst[1 + j] = S.has_u ? src[1 + j] : 0.0;
        dst[1 + S.p] = src[1 + P.pq];
        for (int j = 0; j < S.q; j++) dst[2 + S.p + j] = S.has_v ? src[2 + P.pq + j] : 0.0;
        for (int k = 0; k < 4; k++) dst[2 + S.p + S.q + k] = src[2 + 2 * P.pq + k];
        P.lik_u[fu] = P.lik[i];
        P.iters_u[fu] = P.iters[i];
        P.status_u[fu] = P.status[i];
    }
    if (i < P.n_groups) {
        const int b = P.best[i];
        P.best_u[P.g_user[i]] = b < 0 ? -1 : P.f_user[b];
    }
}

// ---- cross-validation skill metrics (the epilogue of cvLDS) --------------------------------------
// calculate_metrics (R/utils.R:56-70) with the definitions of src/utils.cpp:13-97, one warp per
// fold: R2 = NSE on the calibration part, RE, CE = NSE on the hold-out, nRMSE (normalised by
// mean(obs)), KGE.  held [n_folds][n] marks the hold-out points; sim is exp()'d first when exp_trans
// (cvLDS with transform = 'log', R/LDS_reconstruction.R:385-386).  Two passes (means, then moments
// about the means) with shuffle reductions.
__device__ __forceinline__ double warp_sum(double x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULL, x, o);
    return x;
}
__global__ void cv_metrics_kernel(int n, int n_folds, const double *__restrict__ sim, const double *__restrict__ obs,
                                  const unsigned char *__restrict__ held, int exp_trans, double *__restrict__ out) {
    const int f = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (f >= n_folds) return;
    const double *__restrict__ s = sim + (size_t)f * n;
    const unsigned char *__restrict__ h = held + (size_t)f * n;
    double so = 0, sto = 0, nt = 0, svs = 0, svo = 0, nv = 0, no = 0;
    for (int i = lane; i < n; i += 32) {
        const double o = obs[i], x = exp_trans ? exp(s[i]) : s[i];
        const bool fin = o == o;
        if (fin) {
            so += o;
            no += 1;
        }
        if (h[i]) {
            svs += x;
            svo += o;
            nv += 1;
        } else if (fin) {
            sto += o;
            nt += 1;
        }
    }
    so = warp_sum(so), no = warp_sum(no), sto = warp_sum(sto), nt = warp_sum(nt);
    svs = warp_sum(svs), svo = warp_sum(svo), nv = warp_sum(nv);
    const double norm = so / no, yc = sto / nt, mvs = svs / nv, mvo = svo / nv;
    double t_rss = 0, t_tss = 0, v_rss = 0, v_tss = 0, v_tssc = 0, v_ss = 0, v_sc = 0;
    for (int i = lane; i < n; i += 32) {
        const double o = obs[i], x = exp_trans ? exp(s[i]) : s[i];
        if (h[i]) {
            v_rss = fma(o - x, o - x, v_rss);
            v_tss = fma(o - mvo, o - mvo, v_tss);
            v_tssc = fma(o - yc, o - yc, v_tssc);
            v_ss = fma(x - mvs, x - mvs, v_ss);
            v_sc = fma(x - mvs, o - mvo, v_sc);
        } else if (o == o) {
            t_rss = fma(o - x, o - x, t_rss);
            t_tss = fma(o - yc, o - yc, t_tss);
        }
    }
    t_rss = warp_sum(t_rss), t_tss = warp_sum(t_tss), v_rss = warp_sum(v_rss), v_tss = warp_sum(v_tss);
    v_tssc = warp_sum(v_tssc), v_ss = warp_sum(v_ss), v_sc = warp_sum(v_sc);
    if (lane == 0) {
        const double sg = sqrt(v_tss / (nv - 1)), sgh = sqrt(v_ss / (nv - 1));
        const double r = v_sc / (nv - 1) / (sg * sgh), a = sgh / sg, b = mvs / mvo;
        double *o5 = out + (size_t)f * 5;
        o5[0] = 1.0 - t_rss / t_tss;
        o5[1] = 1.0 - v_rss / v_tssc;
        o5[2] = 1.0 - v_rss / v_tss;
        o5[3] = sqrt(v_rss / nv) / norm;
        o5[4] = 1.0 - sqrt((r - 1) * (r - 1) + (a - 1) * (a - 1) + (b - 1) * (b - 1));
    }
}


// ---- construct_rec + ensemble mean (the step after restart selection in LDS_reconstruction) --------
// R/LDS_reconstruction.R:190-212 with exp_ci / inv_boxcox of R/utils.R:112-125; one thread per year:
// it walks the members in order, so the ensemble mean (:247-248) is a fixed-order sum.
//   X,V,Y [n][T]; Cm,Rm [n]; out [n][6][T] = X, Xl, Xu, Q, Ql, Qu; mean [2][T] = mean X, mean Q.
__global__ void construct_rec_kernel(int n, int T, const double *__restrict__ X, const double *__restrict__ V,
                                     const double *__restrict__ Y, const double *__restrict__ Cm,
                                     const double *__restrict__ Rm, double mu, int transform, double lambda,
                                     double *__restrict__ out, double *__restrict__ mean) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    const double QNORM_05 = -1.6448536269514722; // qnorm(0.05): qlnorm(p, m, s) = exp(m + s qnorm(p))
    double sx = 0.0, sq = 0.0;
    for (int m = 0; m < n; m++) {
        const double x = X[(size_t)m * T + t], v = V[(size_t)m * T + t], y = Y[(size_t)m * T + t] + mu;
        const double ciX = 1.96 * sqrt(v), sdY = sqrt(Cm[m] * v * Cm[m] + Rm[m]), ciY = 1.96 * sdY;
        double q, ql, qu;
        if (transform == 1 || (transform == 2 && lambda == 0.0)) {
            q = exp(y);
            ql = exp(y + sdY * QNORM_05);
            qu = exp(y - sdY * QNORM_05);
        } else if (transform == 0) {
            q = y;
            ql = y - ciY;
            qu = y + ciY;
        } else {
            const double il = 1.0 / lambda;
            q = pow(y * lambda + 1.0, il);
            ql = pow((y - ciY) * lambda + 1.0, il);
            qu = pow((y + ciY) * lambda + 1.0, il);
        }
        double *o = out + (size_t)m * 6 * T;
        o[t] = x;
        o[T + t] = x - ciX;
        o[2 * T + t] = x + ciX;
        o[3 * T + t] = q;
        o[4 * T + t] = ql;
        o[5 * T + t] = qu;
        sx += x;
        sq += q;
    }
    mean[t] = sx / n;
    mean[T + t] = sq / n;
}

// initial EM state: theta <- theta0, no E-step done, lik = NaN
__global__ void init_state_kernel(int n_fits, int theta_len, const double *theta0, double *theta, double *l1,
                                  double *l2, double *lik, int *ne, int *done) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_fits * theta_len) theta[i] = theta0[i];
    if (i < n_fits) {
        l1[i] = 0.0;
        l2[i] = 0.0;
        lik[i] = __longlong_as_double(0x7ff8000000000000ULL);
        ne[i] = 0;
        done[i] = 0;
    }
}

__global__ void fill_nan_kernel(double *p, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = __longlong_as_double(0x7ff8000000000000ULL);
}

__global__ void sum_int_kernel(const int *v, int n, unsigned long long *out) {
    unsigned long long s = 0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) s += (unsigned)v[i];
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(FULL, s, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, s);
}

// FP64 FMA throughput probe: 8 independent dependent-chains per thread, no memory traffic.
__global__ void __launch_bounds__(1024) dfma_peak_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            x0 = fma(x0, a, b);
            x1 = fma(x1, a, b);
            x2 = fma(x2, a, b);
            x3 = fma(x3, a, b);
            x4 = fma(x4, a, b);
            x5 = fma(x5, a, b);
            x6 = fma(x6, a, b);
            x7 = fma(x7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.6789) out[0] = s; // never true; keeps the chains alive
}

} // namespace ldsr
