"""R's default random numbers after set.seed(seed), restated in pure Python -- TEST INFRASTRUCTURE ONLY
(nothing in the product imports this; the product's generators are ldsr_b200/csrc/r_rng.cuh).

The reference is an R package: its restarts draw runif (R/LDS_reconstruction.R:14-30) and its
replicates draw rnorm (R/stochastics.R:23-26) from R's own generator, which is not under
/root/reference.  Restated from the published algorithm: set.seed scrambling (RNG.c RNG_Init,
FixupSeeds), MT19937 + tempering, unif_rand's scaling and fixup, norm_rand INVERSION (snorm.c) and
qnorm = AS 241 PPND16 (qnorm.c).  Pinned in tests/test_r_rng.py by values any R session prints
(set.seed(42); runif(3); set.seed(1); rnorm(3); ...) and by scipy's ndtri for qnorm."""
import numpy as np, math
N, M = 624, 397
class RMT:
    def __init__(self, seed):
        s = np.uint32(seed)
        s = int(s)
        for _ in range(50):
            s = (69069 * s + 1) & 0xFFFFFFFF
        self.mt = []
        dummy0 = None
        for j in range(625):
            s = (69069 * s + 1) & 0xFFFFFFFF
            if j == 0: dummy0 = s
            else: self.mt.append(s)
        self.mti = N  # FixupSeeds: dummy[0] = 624
    def genrand(self):
        mt = self.mt
        if self.mti >= N:
            for kk in range(N - M):
                y = (mt[kk] & 0x80000000) | (mt[kk+1] & 0x7fffffff)
                mt[kk] = mt[kk+M] ^ (y >> 1) ^ (0x9908b0df if y & 1 else 0)
            for kk in range(N - M, N - 1):
                y = (mt[kk] & 0x80000000) | (mt[kk+1] & 0x7fffffff)
                mt[kk] = mt[kk+(M-N)] ^ (y >> 1) ^ (0x9908b0df if y & 1 else 0)
            y = (mt[N-1] & 0x80000000) | (mt[0] & 0x7fffffff)
            mt[N-1] = mt[M-1] ^ (y >> 1) ^ (0x9908b0df if y & 1 else 0)
            self.mti = 0
        y = mt[self.mti]; self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9d2c5680
        y ^= (y << 15) & 0xefc60000
        y ^= y >> 18
        return (y & 0xFFFFFFFF) * 2.3283064365386963e-10
    def unif(self):
        v = self.genrand()
        i2 = 2.328306437080797e-10
        if v <= 0.0: return 0.5 * i2
        if 1.0 - v <= 0.0: return 1.0 - 0.5 * i2
        return v
    def norm(self):
        BIG = 134217728.0
        u = self.unif()
        u = int(BIG * u) + self.unif()
        return qnorm(u / BIG)

def qnorm(p):
    q = p - 0.5
    if abs(q) <= 0.425:
        r = .180625 - q * q
        return q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r
                        + 45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r
                      + 133.14166789178437745) * r + 3.387132872796366608) / \
                   (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r
                        + 21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r
                      + 42.313330701600911252) * r + 1.)
    r = p if q < 0 else 1.0 - p
    r = math.sqrt(-math.log(r))
    if r <= 5.:
        r -= 1.6
        val = (((((((r * 7.7454501427834140764e-4 + .0227238449892691845833) * r + .24178072517745061177) * r
                   + 1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r
                 + 4.6303378461565452959) * r + 1.42343711074968357734) / \
              (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + .0151986665636164571966) * r
                   + .14810397642748007459) * r + .68976733498510000455) * r + 1.6763848301838038494) * r
                 + 2.05319162663775882187) * r + 1.)
    else:
        r -= 5.
        val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + .0012426609473880784386) * r
                   + .026532189526576123093) * r + .29656057182850489123) * r + 1.7848265399172913358) * r
                 + 5.4637849111641143699) * r + 6.6579046435011037772) / \
              (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r
                   + 7.868691311456132591e-4) * r + .0148753612908506148525) * r + .13692988092273580531) * r
                 + .59983220655588793769) * r + 1.)
    return -val if q < 0 else val



def unif_index(g, dn):
    """R_unif_index (RNG.c), sample.kind = "Rejection" (the default since R 3.6.0)."""
    if dn <= 0:
        return 0
    bits = int(math.ceil(math.log2(dn)))
    while True:
        v = 0
        n = 0
        while n <= bits:
            v = 65536 * v + int(math.floor(g.unif() * 65536))
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        if v < dn:
            return v


def sample_int(g, n, k):
    """sample.int(n, k) without replacement (do_sample, unique.c / random.c): partial Fisher-Yates."""
    x = list(range(n))
    out = []
    for _ in range(k):
        j = unif_index(g, n)
        out.append(x[j] + 1)
        n -= 1
        x[j] = x[n]
    return out
