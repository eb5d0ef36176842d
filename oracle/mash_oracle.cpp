// mash_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the fp-mash hot path (mash sketch / mash dist) used as the
// parity oracle for the CUDA kernels.  Nothing in the product path (fp-mash_b200/)
// links, imports or executes this file; only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may.
//
// Parity status: PINNED.  tests/test_oracle_golden.py checks this file against the
// reference's own fixtures: test/ref/reads.json (1000 hashes +
// length), new_data/reads/reads.msh (counts), test_sequence.msh, DNA{1,2,3}-sketch
// .json/.msh (fp mode), test/ref/genomes.dist + tutorials.rst rows (dist), the
// Appendix-C pocket vectors, and -- in this container -- against oracle/_ref, the
// reference's own hash.cpp/MurmurHash3.cpp/MinHashHeap.cpp compiled unmodified.
//
// Every function cites the reference file:line it restates (paths relative to
// mash/src/mash/ of UmbertoDellaMonica/fp-mash).
//
// The p-value has no reference source in the tree (GSL gsl_cdf_binomial_Q or Boost,
// version unpinned: CommandDistance.cpp:446-448, configure.ac:11-17).  The oracle
// evaluates the mathematically exact binomial tail in 80-bit long double; tests pin it
// to mpmath at 50 digits and to the 6-significant-digit rows of test/ref/genomes.dist.

#include <cstdint>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <queue>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

// ---------------------------------------------------------------------------
// MurmurHash3_x64_128  (MurmurHash3.cpp:255-331)
// ---------------------------------------------------------------------------
inline uint64_t rol64(uint64_t v, int r) { return (v << r) | (v >> (64 - r)); }

inline uint64_t avalanche64(uint64_t v)  // fmix64, MurmurHash3.cpp:81-90
{
    v ^= v >> 33;
    v *= 0xff51afd7ed558ccdULL;
    v ^= v >> 33;
    v *= 0xc4ceb9fe1a85ec53ULL;
    v ^= v >> 33;
    return v;
}

const uint64_t MC1 = 0x87c37b91114253d5ULL;
const uint64_t MC2 = 0x4cf5ad432745937fULL;

inline uint64_t mix_lane1(uint64_t k) { k *= MC1; k = rol64(k, 31); k *= MC2; return k; }
inline uint64_t mix_lane2(uint64_t k) { k *= MC2; k = rol64(k, 33); k *= MC1; return k; }

void murmur3_x64_128(const uint8_t* key, int len, uint32_t seed, uint64_t out[2])
{
    uint64_t ha = seed, hb = seed;
    int full = len / 16;
    for (int b = 0; b < full; b++) {                 // body, :270-282
        uint64_t ka, kb;
        memcpy(&ka, key + 16 * b, 8);
        memcpy(&kb, key + 16 * b + 8, 8);
        ha ^= mix_lane1(ka);
        ha = rol64(ha, 27); ha += hb; ha = ha * 5 + 0x52dce729ULL;
        hb ^= mix_lane2(kb);
        hb = rol64(hb, 31); hb += ha; hb = hb * 5 + 0x38495ab5ULL;
    }
    // tail, :287-314 -- the fall-through switch is "load the remaining bytes little-endian
    // into a zero-padded 16-byte block"; lane 2 is mixed only if >8 bytes remain, lane 1
    // if >0 remain.
    int rem = len & 15;
    if (rem) {
        uint8_t pad[16] = {0};
        memcpy(pad, key + 16 * full, rem);
        uint64_t ka, kb;
        memcpy(&ka, pad, 8);
        memcpy(&kb, pad + 8, 8);
        if (rem > 8) hb ^= mix_lane2(kb);
        ha ^= mix_lane1(ka);
    }
    ha ^= (uint64_t)(int64_t)len; hb ^= (uint64_t)(int64_t)len;   // finalization, :318-330
    ha += hb; hb += ha;
    ha = avalanche64(ha); hb = avalanche64(hb);
    ha += hb; hb += ha;
    out[0] = ha; out[1] = hb;
}

// getHash (hash.cpp:12-40): low 64 (use64) or low 32 bits of h1.
inline uint64_t get_hash(const uint8_t* s, int len, uint32_t seed, bool use64)
{
    uint64_t o[2];
    murmur3_x64_128(s, len, seed, o);
    return use64 ? o[0] : (o[0] & 0xffffffffULL);
}

// ---------------------------------------------------------------------------
// MinHashHeap (MinHashHeap.cpp:7-146, MinHashHeap.h:44-47) -- literal restatement of
// tryInsert with std containers standing in for robin_hood (iteration order is never
// observable: toHashList sorts, HashSet.cpp:78-118).  No Bloom filter (-b is out of
// scope).
// ---------------------------------------------------------------------------
struct Heap {
    bool use64;
    uint64_t cap;      // cardinalityMaximum
    uint64_t mincov;   // multiplicityMinimum
    uint64_t multsum;
    std::unordered_map<uint64_t, uint32_t> kept, pending;
    std::priority_queue<uint64_t> keptq, pendq;

    Heap(bool u, uint64_t c, uint64_t m) : use64(u), cap(c), mincov(m), multsum(0) {}

    uint32_t cnt(const std::unordered_map<uint64_t, uint32_t>& m, uint64_t h) const
    {
        auto it = m.find(h);
        return it == m.end() ? 0 : it->second;
    }

    void offer(uint64_t h)   // tryInsert, MinHashHeap.cpp:68-146
    {
        if (!(kept.size() < cap || h < keptq.top())) return;           // :70-74
        if (cnt(kept, h) == 0) {                                       // :76
            if (mincov == 1 || cnt(pending, h) == mincov - 1) {        // :96
                kept[h] += (uint32_t)mincov;
                keptq.push(h);
                multsum += mincov;
                if (mincov > 1) pending.erase(h);                      // :102-108
            } else {
                if (cnt(pending, h) == 0) pendq.push(h);               // :112-115
                pending[h] += 1;
            }
        } else {
            kept[h] += 1;                                              // :121-124
            multsum++;
        }
        if (kept.size() > cap) {                                       // :126
            multsum -= cnt(kept, keptq.top());
            kept.erase(keptq.top());
            while (!pendq.empty() && keptq.top() < pendq.top()) {      // :133-141
                pending.erase(pendq.top());
                pendq.pop();
            }
            keptq.pop();
        }
    }

    double set_size() const   // estimateSetSize, MinHashHeap.h:45
    {
        if (kept.empty()) return 0;
        return pow(2.0, use64 ? 64.0 : 32.0) * (double)kept.size() / (double)keptq.top();
    }
    double multiplicity() const   // estimateMultiplicity, MinHashHeap.h:44
    {
        return kept.empty() ? 0 : (double)multsum / kept.size();
    }
};

// reverseComplement (Sketch.cpp:1223-1258).  The table covers 'A'..'Z'; the reference
// indexes it out of bounds for other bytes, which only ever happens inside windows
// that are skipped, so any value works there: we use 'N'.
const char COMP[27] = "TVGHNNCDNNMNKNNNNYSAABWNRN";

// addMinHashes (Sketch.cpp:664-735)
void add_min_hashes(Heap& heap, uint8_t* seq, uint64_t length, int k, uint32_t seed,
                    bool noncanonical, bool preserve_case, const uint8_t* alphabet,
                    std::vector<uint64_t>* trace)
{
    if (!preserve_case)                                                // :676-682
        for (uint64_t i = 0; i < length; i++)
            if (seq[i] > 96 && seq[i] < 123) seq[i] -= 32;

    std::vector<uint8_t> rev;
    if (!noncanonical) {                                               // :686-690
        rev.resize(length);
        for (uint64_t i = 0; i < length; i++) {
            int c = (int)seq[length - 1 - i] - 'A';
            rev[i] = (c >= 0 && c < 26) ? (uint8_t)COMP[c] : (uint8_t)'N';
        }
    }
    if (length < (uint64_t)k) return;

    // :692-729.  The reference's j-scan skips every window touching a byte outside the
    // alphabet; equivalently window i is hashed iff the run of valid bytes ending at
    // i+k-1 is at least k long.
    uint64_t run = 0;
    for (uint64_t e = 0; e < length; e++) {
        run = alphabet[seq[e]] ? run + 1 : 0;
        if (run < (uint64_t)k) continue;
        uint64_t i = e + 1 - k;
        const uint8_t* fwd = seq + i;
        const uint8_t* pick = fwd;
        if (!noncanonical) {
            const uint8_t* rc = rev.data() + length - i - k;
            if (memcmp(fwd, rc, k) > 0) pick = rc;                     // :721-723
        }
        uint64_t h = get_hash(pick, k, seed, heap.use64);              // :726
        if (trace) trace->push_back(h);
        heap.offer(h);                                                 // :728
    }
}

// ---------------------------------------------------------------------------
// pValue (CommandDistance.cpp:433-450): P[Binomial(n, r) >= x], exact tail in long
// double.  Terms are generated from the mode-side outward so every partial sum is a sum
// of decreasing positive terms (no cancellation except the final 1-L when x is below
// the mean, where L <= ~0.5).
// ---------------------------------------------------------------------------
// Scaled product C(n,i) r^i (1-r)^(n-i) as mantissa * 2^exp, built from i exact-ish
// multiplications (each ~0.5 ulp of 64-bit mantissa) rather than lgamma.
void binom_pmf_scaled(uint64_t n, uint64_t i, long double r, long double& mant, long& ex)
{
    mant = 1.0L; ex = 0;
    for (uint64_t j = 1; j <= i; j++) {
        mant *= ((long double)(n - i + j) / (long double)j) * r;
        int e; mant = frexpl(mant, &e); ex += e;
    }
    // (1-r)^(n-i) via exp2(m*log2(1-r)) split into integer and fractional parts
    long double t = (long double)(n - i) * (log1pl(-r) / logl(2.0L));
    long double ti = floorl(t);
    mant *= exp2l(t - ti);
    int e; mant = frexpl(mant, &e); ex += e + (long)ti;
}

long double binom_tail_ge(uint64_t x, uint64_t n, long double r)
{
    if (x == 0) return 1.0L;
    if (x > n) return 0.0L;
    if (r <= 0) return 0.0L;
    if (r >= 1) return 1.0L;
    long double odds = r / (1.0L - r);
    long double mean = (long double)(n + 1) * r;
    if ((long double)x >= mean) {
        // upper tail: sum_{i>=x} t_i, ratio t_{i+1}/t_i = (n-i)/(i+1)*odds < 1
        long double m; long e;
        binom_pmf_scaled(n, x, r, m, e);
        long double sum = m, term = m;
        for (uint64_t i = x; i < n; i++) {
            term *= ((long double)(n - i) / (long double)(i + 1)) * odds;
            sum += term;
            if (term < sum * 1e-25L) break;
        }
        return ldexpl(sum, (int)std::max<long>(e, -20000));
    }
    // lower tail complement: 1 - sum_{i<=x-1} t_i, descending from i = x-1
    long double m; long e;
    binom_pmf_scaled(n, x - 1, r, m, e);
    long double sum = m, term = m;
    for (uint64_t i = x - 1; i > 0; i--) {
        term *= ((long double)i / (long double)(n - i + 1)) / odds;
        sum += term;
        if (term < sum * 1e-25L) break;
    }
    return 1.0L - ldexpl(sum, (int)std::max<long>(e, -20000));
}

double p_value(uint64_t x, uint64_t len_ref, uint64_t len_qry, double kmer_space, uint64_t n)
{
    if (x == 0) return 1.;                                            // :435-438
    double pX = 1. / (1. + kmer_space / len_ref);                      // :440-441
    double pY = 1. / (1. + kmer_space / len_qry);
    double r = pX * pY / (pX + pY - pX * pY);                          // :443
    return (double)binom_tail_ge(x, n, (long double)r);                // :446-448
}

}  // namespace

extern "C" {

struct orc_pair {
    uint64_t numer, denom;
    double distance, pvalue;
    int pass;
};

void orc_murmur3_x64_128(const void* key, int len, uint32_t seed, uint64_t* out2)
{
    murmur3_x64_128((const uint8_t*)key, len, seed, out2);
}

uint64_t orc_get_hash(const void* seq, int len, uint32_t seed, int use64)
{
    return get_hash((const uint8_t*)seq, len, seed, use64 != 0);
}

// getHashFingerPrint (hash.cpp:45-73): Murmur over the raw little-endian bytes of the
// uint64 token vector, n*8 bytes.
uint64_t orc_fp_hash(const uint64_t* tokens, int n_tokens, uint32_t seed, int use64)
{
    return get_hash((const uint8_t*)tokens, n_tokens * 8, seed, use64 != 0);
}

void* orc_heap_new(int use64, uint64_t sketch_size, uint64_t min_cov)
{
    return new Heap(use64 != 0, sketch_size, min_cov);
}
void orc_heap_free(void* h) { delete (Heap*)h; }
void orc_heap_offer(void* h, const uint64_t* hashes, uint64_t n)
{
    Heap* H = (Heap*)h;
    for (uint64_t i = 0; i < n; i++) H->offer(hashes[i]);
}

// One addMinHashes call (one record).  seq is modified in place like the reference.
// If trace/trace_cap are given, the hash stream is copied out (for closed-form tests).
uint64_t orc_heap_add_sequence(void* h, char* seq, uint64_t length, int k, uint32_t seed,
                               int noncanonical, int preserve_case, const uint8_t* alphabet256,
                               uint64_t* trace, uint64_t trace_cap)
{
    std::vector<uint64_t> tr;
    add_min_hashes(*(Heap*)h, (uint8_t*)seq, length, k, seed, noncanonical != 0,
                   preserve_case != 0, alphabet256, trace ? &tr : nullptr);
    if (trace) memcpy(trace, tr.data(), std::min<uint64_t>(tr.size(), trace_cap) * 8);
    return tr.size();
}

uint64_t orc_heap_size(void* h) { return ((Heap*)h)->kept.size(); }
double orc_heap_set_size(void* h) { return ((Heap*)h)->set_size(); }
double orc_heap_multiplicity(void* h) { return ((Heap*)h)->multiplicity(); }

// toHashList (HashSet.cpp:78-118): ascending hashes with parallel counts.
uint64_t orc_heap_result(void* h, uint64_t* hashes, uint32_t* counts)
{
    Heap* H = (Heap*)h;
    std::vector<std::pair<uint64_t, uint32_t>> v(H->kept.begin(), H->kept.end());
    std::sort(v.begin(), v.end());
    for (size_t i = 0; i < v.size(); i++) {
        if (hashes) hashes[i] = v[i].first;
        if (counts) counts[i] = v[i].second;
    }
    return v.size();
}

double orc_pvalue(uint64_t x, uint64_t len_ref, uint64_t len_qry, double kmer_space, uint64_t n)
{
    return p_value(x, len_ref, len_qry, kmer_space, n);
}

double orc_binom_tail(uint64_t x, uint64_t n, double r) { return (double)binom_tail_ge(x, n, r); }

// compareSketches (CommandDistance.cpp:365-430), literal loop (also defines the result on
// unsorted / repeating fp-mode lists).  Hash lists are passed widened to u64.
void orc_compare_sketches(const uint64_t* ref, uint64_t n_ref, const uint64_t* qry, uint64_t n_qry,
                          uint64_t len_ref, uint64_t len_qry, uint64_t sketch_size, int kmer_size,
                          double kmer_space, double max_distance, double max_pvalue, orc_pair* out)
{
    uint64_t i = 0, j = 0, common = 0, denom = 0;
    out->pass = 0;
    out->numer = out->denom = 0; out->distance = out->pvalue = 0;
    while (denom < sketch_size && i < n_ref && j < n_qry) {            // :376-387
        if (ref[i] < qry[j]) i++;
        else if (qry[j] < ref[i]) j++;
        else { i++; j++; common++; }
        denom++;
    }
    if (denom < sketch_size) {                                         // :389-400
        if (i < n_ref) denom += n_ref - i;
        if (j < n_qry) denom += n_qry - j;
        if (denom > sketch_size) denom = sketch_size;
    }
    double distance;
    double jaccard = double(common) / denom;                           // :403
    if (common == denom) distance = 0;                                 // :405-414
    else if (common == 0) distance = 1.;
    else {
        distance = -log(2 * jaccard / (1. + jaccard)) / kmer_size;
        if (distance > 1) distance = 1;
    }
    out->numer = common; out->denom = denom; out->distance = distance;
    if (max_distance >= 0 && distance > max_distance) return;          // :416-419
    out->pvalue = p_value(common, len_ref, len_qry, kmer_space, denom);
    if (max_pvalue >= 0 && out->pvalue > max_pvalue) return;           // :425-428
    out->pass = 1;
}

}  // extern "C"
