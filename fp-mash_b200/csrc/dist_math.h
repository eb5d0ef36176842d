// dist_math.h -- Jaccard -> Mash distance and the binomial-tail p-value, shared by the
// dist kernels (device) and the host mirror of compareSketches (host), so both sides run
// the same double-precision algorithm.
//
// Replaces (fp-mash mash/src/mash/):
//   CommandDistance.cpp:402-414  distance = -log(2j/(1+j))/k with the 0 / 1 special cases
//   CommandDistance.cpp:433-450  pValue(): P[Binomial(n=denom, r) >= x] via
//                                gsl_cdf_binomial_Q(x-1, r, n) (or Boost) -- third-party,
//                                not vendored.  We evaluate the mathematically exact tail
//                                by summing pmf terms from the tail start outward with a
//                                ratio recurrence (all terms positive and decreasing), the
//                                first term built as a scaled product so nothing under- or
//                                overflows before the final ldexp.  Target: <=1e-12
//                                relative to the exact value (tests pin it to mpmath).
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define FPM_HD __host__ __device__ __forceinline__
#else
#define FPM_HD inline
#endif

namespace fpm {

FPM_HD double mash_distance(uint64_t common, uint64_t denom, int kmer_size)
{
    if (common == denom) return 0.;          // avoid -0 (also 0/0: both sketches empty)
    if (common == 0) return 1.;              // avoid inf
    double jaccard = double(common) / double(denom);
    double d = -log(2 * jaccard / (1. + jaccard)) / kmer_size;
    return d > 1 ? 1. : d;
}

// C(n,i) r^i (1-r)^(n-i) as mant * 2^ex, mant in [0.5,1).
FPM_HD void binom_pmf_scaled(uint64_t n, uint64_t i, double r, double& mant, long long& ex)
{
    double m = 1.0;
    long long e = 0;
    uint64_t j = 1;
    while (j <= i) {
        // up to 8 factors between renormalisations: each factor is within [r, n*r],
        // i.e. > 1e-30 and < 1e30 for every r this path sees, so 8 of them cannot
        // leave the double range starting from [0.5,1).
        uint64_t stop = j + 8 <= i + 1 ? j + 8 : i + 1;
        for (; j < stop; j++) m *= (double(n - i + j) / double(j)) * r;
        int t;
        m = frexp(m, &t);
        e += t;
    }
    // (1-r)^(n-i) = 2^(t), t = (n-i) * log2(1-r), integer part goes to the exponent
    double t = double(n - i) * (log1p(-r) * 1.4426950408889634074);
    double ti = floor(t);
    m *= exp2(t - ti);
    int q;
    m = frexp(m, &q);
    e += q + (long long)ti;
    mant = m;
    ex = e;
}

FPM_HD double ldexp_clamped(double m, long long e)
{
    if (e < -1200) return 0.;
    if (e > 1200) e = 1200;
    return ldexp(m, (int)e);
}

// P[Binomial(n, r) >= x]
FPM_HD double binom_tail_ge(uint64_t x, uint64_t n, double r)
{
    if (x == 0) return 1.;
    if (x > n) return 0.;
    if (!(r > 0.)) return 0.;
    if (!(r < 1.)) return 1.;
    double odds = r / (1. - r);
    double mean = double(n + 1) * r;
    double m;
    long long e;
    if (double(x) >= mean) {
        // Cheap exit for hopeless tails (the common case between related genomes: hundreds of shared
        // hashes, p far below the smallest double).  tail <= (n+1) * C(n,x) r^x <= (n+1) * (e*n*r/x)^x;
        // below 2^-1200 the full evaluation returns 0 as well (ldexp_clamped), so results are identical.
        if (double(x) * log2(2.718281828459045 * double(n) * r / double(x)) + log2(double(n) + 1.) < -1200.) return 0.;
        binom_pmf_scaled(n, x, r, m, e);
        double sum = m, term = m;
        for (uint64_t i = x; i < n; i++) {
            term *= (double(n - i) / double(i + 1)) * odds;
            sum += term;
            if (term < sum * 1e-18) break;
        }
        return ldexp_clamped(sum, e);
    }
    binom_pmf_scaled(n, x - 1, r, m, e);
    double sum = m, term = m;
    for (uint64_t i = x - 1; i > 0; i--) {
        term *= (double(i) / double(n - i + 1)) / odds;
        sum += term;
        if (term < sum * 1e-18) break;
    }
    return 1. - ldexp_clamped(sum, e);
}

FPM_HD double mash_pvalue(uint64_t x, uint64_t len_ref, uint64_t len_qry, double kmer_space, uint64_t n)
{
    if (x == 0) return 1.;
    double pX = 1. / (1. + kmer_space / double(len_ref));
    double pY = 1. / (1. + kmer_space / double(len_qry));
    double r = pX * pY / (pX + pY - pX * pY);
    return binom_tail_ge(x, n, r);
}

}  // namespace fpm
