// r_rng.cuh -- the reference's random numbers, bit for bit: what R's default generators return
// after set.seed(seed) (RNGkind "Mersenne-Twister", normal.kind "Inversion").
//
// The reference draws its restarts' initial values with runif (R/LDS_reconstruction.R:14-30) and its
// stochastic replicates with rnorm (R/stochastics.R:23-26, one stream across lapply, :60-61); a
// drop-in that is to reproduce `set.seed(s); LDS_rep(...)` must reproduce that stream.  R itself is
// not in /root/reference (it is the host language); what is restated here is its published algorithm:
//   * set.seed: 50 rounds of the LCG seed <- 69069 seed + 1, then 625 more fill (mti, mt[0..623]);
//     mti is then forced to 624 (RNG.c: Randomize / FixupSeeds)
//   * MT19937 (Matsumoto & Nishimura 1998) with the standard tempering; unif_rand = y * 2^-32 moved
//     into the open interval (RNG.c: MT_genrand, fixup)
//   * norm_rand, INVERSION: u = unif_rand(); u = (int)(2^27 u) + unif_rand(); qnorm(u / 2^27)
//     (snorm.c), qnorm = Wichura's AS 241 PPND16 (qnorm.c)
// Pinned by tests/test_r_rng.py to values every R user knows (set.seed(42); runif(3) = 0.9148060
// 0.9370754 0.2861395; set.seed(1); rnorm(3) = -0.6264538 0.1836433 -0.8356286; set.seed(123) ...),
// and AS 241 to scipy's ndtri over the whole range (1e-15).
//
// Host: RMersenne, a plain sequential generator (initial values: a few thousand draws).
// Device: r_rnorm_device -- one CTA walks the twister (one sequential stream: the parallelism is
// inside a regeneration of the 624 words), the whole grid then turns output pairs into normals.
// The arithmetic of qnorm is written with the non-contracting intrinsics (__dmul_rn, __dadd_rn,
// __ddiv_rn, __dsqrt_rn) so that it rounds like R's compiled C does on x86-64; log() is CUDA's
// (<= 1 ulp), so tail draws (|p - 0.5| > 0.425, 15 % of them) may differ from R's in the last bit.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

namespace ldsr {

constexpr int R_MT_N = 624, R_MT_M = 397;

__host__ __device__ inline uint32_t r_mt_temper(uint32_t y) {
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
}
__host__ __device__ inline uint32_t r_mt_twist(uint32_t cur, uint32_t next, uint32_t far) {
    const uint32_t y = (cur & 0x80000000u) | (next & 0x7fffffffu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
// MT_genrand's scaling and unif_rand's fixup into (0, 1)
__host__ __device__ inline double r_unif_from_u32(uint32_t y) {
    const double v = (double)y * 2.3283064365386963e-10;
    const double i2_32m1 = 2.328306437080797e-10;
    if (v <= 0.0) return 0.5 * i2_32m1;
    if (1.0 - v <= 0.0) return 1.0 - 0.5 * i2_32m1;
    return v;
}

#ifdef __CUDA_ARCH__
#define LDSR_R_MUL(a, b) __dmul_rn((a), (b))
#define LDSR_R_ADD(a, b) __dadd_rn((a), (b))
#define LDSR_R_DIV(a, b) __ddiv_rn((a), (b))
#define LDSR_R_SQRT(a) __dsqrt_rn(a)
#else
#define LDSR_R_MUL(a, b) ((a) * (b))
#define LDSR_R_ADD(a, b) ((a) + (b))
#define LDSR_R_DIV(a, b) ((a) / (b))
#define LDSR_R_SQRT(a) std::sqrt(a)
#endif

// Horner step acc * r + c without contraction
__host__ __device__ inline double r_horner(double acc, double r, double c) { return LDSR_R_ADD(LDSR_R_MUL(acc, r), c); }

// qnorm5(p, 0, 1, lower_tail = TRUE, log_p = FALSE) for 0 < p < 1 (AS 241, PPND16)
__host__ __device__ inline double r_qnorm(double p) {
    const double q = LDSR_R_ADD(p, -0.5);
    if (fabs(q) <= 0.425) {
        const double r = LDSR_R_ADD(0.180625, -LDSR_R_MUL(q, q));
        double num = 2509.0809287301226727;
        num = r_horner(num, r, 33430.575583588128105);
        num = r_horner(num, r, 67265.770927008700853);
        num = r_horner(num, r, 45921.953931549871457);
        num = r_horner(num, r, 13731.693765509461125);
        num = r_horner(num, r, 1971.5909503065514427);
        num = r_horner(num, r, 133.14166789178437745);
        num = r_horner(num, r, 3.387132872796366608);
        double den = 5226.495278852545925;
        den = r_horner(den, r, 28729.085735721942674);
        den = r_horner(den, r, 39307.89580009271061);
        den = r_horner(den, r, 21213.794301586595867);
        den = r_horner(den, r, 5394.1960214247511077);
        den = r_horner(den, r, 687.1870074920579083);
        den = r_horner(den, r, 42.313330701600911252);
        den = r_horner(den, r, 1.0);
        return LDSR_R_DIV(LDSR_R_MUL(q, num), den);
    }
    double r = q < 0.0 ? p : LDSR_R_ADD(1.0, -p);
    r = LDSR_R_SQRT(-log(r));
    double val;
    if (r <= 5.0) {
        r = LDSR_R_ADD(r, -1.6);
        double num = 7.7454501427834140764e-4;
        num = r_horner(num, r, 0.0227238449892691845833);
        num = r_horner(num, r, 0.24178072517745061177);
        num = r_horner(num, r, 1.27045825245236838258);
        num = r_horner(num, r, 3.64784832476320460504);
        num = r_horner(num, r, 5.7694972214606914055);
        num = r_horner(num, r, 4.6303378461565452959);
        num = r_horner(num, r, 1.42343711074968357734);
        double den = 1.05075007164441684324e-9;
        den = r_horner(den, r, 5.475938084995344946e-4);
        den = r_horner(den, r, 0.0151986665636164571966);
        den = r_horner(den, r, 0.14810397642748007459);
        den = r_horner(den, r, 0.68976733498510000455);
        den = r_horner(den, r, 1.6763848301838038494);
        den = r_horner(den, r, 2.05319162663775882187);
        den = r_horner(den, r, 1.0);
        val = LDSR_R_DIV(num, den);
    } else {
        r = LDSR_R_ADD(r, -5.0);
        double num = 2.01033439929228813265e-7;
        num = r_horner(num, r, 2.71155556874348757815e-5);
        num = r_horner(num, r, 0.0012426609473880784386);
        num = r_horner(num, r, 0.026532189526576123093);
        num = r_horner(num, r, 0.29656057182850489123);
        num = r_horner(num, r, 1.7848265399172913358);
        num = r_horner(num, r, 5.4637849111641143699);
        num = r_horner(num, r, 6.6579046435011037772);
        double den = 2.04426310338993978564e-15;
        den = r_horner(den, r, 1.4215117583164458887e-7);
        den = r_horner(den, r, 1.8463183175100546818e-5);
        den = r_horner(den, r, 7.868691311456132591e-4);
        den = r_horner(den, r, 0.0148753612908506148525);
        den = r_horner(den, r, 0.13692988092273580531);
        den = r_horner(den, r, 0.59983220655588793769);
        den = r_horner(den, r, 1.0);
        val = LDSR_R_DIV(num, den);
    }
    return q < 0.0 ? -val : val;
}

// norm_rand() of the INVERSION kind from two consecutive unif_rand() values
__host__ __device__ inline double r_norm_from_unifs(double u1, double u2) {
    const double BIG = 134217728.0; // 2^27
    const double u = LDSR_R_ADD((double)(int)LDSR_R_MUL(BIG, u1), u2);
    return r_qnorm(LDSR_R_DIV(u, BIG));
}

// set.seed(seed): initial scrambling and fill (RNG.c: RNG_Init + FixupSeeds)
__host__ __device__ inline void r_mt_seed(uint32_t seed, uint32_t *mt /* [624] */) {
    for (int j = 0; j < 50; j++) seed = 69069u * seed + 1u;
    seed = 69069u * seed + 1u; // i_seed[0]: the stored position, overwritten with 624 by FixupSeeds
    for (int j = 0; j < R_MT_N; j++) {
        seed = 69069u * seed + 1u;
        mt[j] = seed;
    }
}

// ---- host: sequential generator ----------------------------------------------------------------
struct RMersenne {
    uint32_t mt[R_MT_N];
    int mti;
    explicit RMersenne(uint32_t seed) : mti(R_MT_N) { r_mt_seed(seed, mt); }
    uint32_t next_u32() {
        if (mti >= R_MT_N) {
            int k = 0;
            for (; k < R_MT_N - R_MT_M; k++) mt[k] = r_mt_twist(mt[k], mt[k + 1], mt[k + R_MT_M]);
            for (; k < R_MT_N - 1; k++) mt[k] = r_mt_twist(mt[k], mt[k + 1], mt[k + (R_MT_M - R_MT_N)]);
            mt[R_MT_N - 1] = r_mt_twist(mt[R_MT_N - 1], mt[0], mt[R_MT_M - 1]);
            mti = 0;
        }
        return r_mt_temper(mt[mti++]);
    }
    double unif() { return r_unif_from_u32(next_u32()); }
    double norm() {
        const double u1 = unif();
        const double u2 = unif();
        return r_norm_from_unifs(u1, u2);
    }
    // R_unif_index (RNG.c), sample.kind "Rejection" (default since R 3.6.0): an integer below dn
    // from ceil(log2(dn)) random bits, 16 per unif_rand(), redrawn until it is below dn
    long long unif_index(double dn) {
        if (dn <= 0) return 0;
        const int bits = (int)std::ceil(std::log2(dn));
        for (;;) {
            long long v = 0;
            for (int n = 0; n <= bits; n += 16) v = 65536 * v + (long long)std::floor(unif() * 65536);
            if (bits < 64) v &= (1LL << bits) - 1;
            if ((double)v < dn) return v;
        }
    }
    // sample.int(n, k) without replacement (do_sample): partial Fisher-Yates, 1-based results
    void sample_int(int n, int k, int *out) {
        std::vector<int> x(n);
        for (int i = 0; i < n; i++) x[i] = i;
        for (int i = 0; i < k; i++) {
            const int j = (int)unif_index(n);
            out[i] = x[j] + 1;
            x[j] = x[--n];
        }
    }
};

#ifdef __CUDACC__
// ---- device: set.seed(seed); rnorm(n) ----------------------------------------------------------
// Two kernels.  r_mt_stream_kernel: ONE CTA walks the twister (it is a single sequential stream) and
// writes the tempered 32-bit outputs, two per normal, into the 8 bytes the normal will occupy.  A
// regeneration of the 624 words has three phases (words 0..226 need only old words, 227..453 need
// the new 0..226, 454..623 the new 227..396 and, for the last one, the new word 0); inside a phase
// every word is independent, and with the state double-buffered (read old, write new) a phase is one
// barrier.  r_norm_from_words_kernel: every pair of outputs becomes a normal, in place, over the
// whole grid (unif_rand scaling and fixup, the 2^27 splice, AS 241).
__global__ void __launch_bounds__(256) r_mt_stream_kernel(uint32_t seed, long long n_words, uint32_t *__restrict__ out) {
    __shared__ uint32_t st[2][R_MT_N];
    const int tid = threadIdx.x;
    if (tid == 0) r_mt_seed(seed, st[0]);
    __syncthreads();
    int cur = 0;
    for (long long base = 0; base < n_words; base += R_MT_N, cur ^= 1) {
        const uint32_t *__restrict__ o = st[cur];
        uint32_t *__restrict__ nw = st[cur ^ 1];
        if (tid < 227) nw[tid] = r_mt_twist(o[tid], o[tid + 1], o[tid + R_MT_M]);
        __syncthreads();
        if (tid < 227) nw[227 + tid] = r_mt_twist(o[227 + tid], o[228 + tid], nw[tid]);
        __syncthreads();
        if (tid < 170) {
            const int k = 454 + tid;
            nw[k] = r_mt_twist(o[k], k == R_MT_N - 1 ? nw[0] : o[k + 1], nw[k - 227]);
        }
        __syncthreads();
        for (int j = tid; j < R_MT_N; j += 256)
            if (base + j < n_words) out[base + j] = r_mt_temper(nw[j]);
        // the next regeneration writes st[cur] only after its first barrier... which no thread passes
        // before every thread has left this loop: st[cur] (the old state) is not read here
    }
}
__global__ void r_norm_from_words_kernel(long long n, double *__restrict__ inout) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 w = reinterpret_cast<const uint2 *>(inout)[i];
    inout[i] = r_norm_from_unifs(r_unif_from_u32(w.x), r_unif_from_u32(w.y));
}
// set.seed(seed); rnorm(n) into d_out (device), on stream st
inline cudaError_t r_rnorm_device(uint32_t seed, long long n, double *d_out, cudaStream_t st) {
    r_mt_stream_kernel<<<1, 256, 0, st>>>(seed, 2 * n, reinterpret_cast<uint32_t *>(d_out));
    r_norm_from_words_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, d_out);
    return cudaGetLastError();
}
#endif

} // namespace ldsr
