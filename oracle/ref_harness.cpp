// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/mash_oracle.cpp header).
//
// Thin C-ABI harness around the reference's OWN translation units, compiled where they
// lie under /root/reference/mash/src/mash (hash.cpp, MurmurHash3.cpp, MinHashHeap.cpp,
// HashSet.cpp, HashList.cpp, HashPriorityQueue.cpp; headers kseq.h, robin_hood.h,
// bloom_filter.hpp) by oracle/Makefile into oracle/_ref/libmashref.so.  No reference
// source is copied into this repository.
//
// Sketch.cpp / CommandDistance.cpp cannot be compiled here (they need libcapnp, the
// generated MinHash.capnp.h and GSL -- SURVEY.md section 8c), so the driver loops around
// the reference classes are restated below, each citing the lines it follows.  The
// harness is used (1) to pin oracle/mash_oracle.cpp, (2) as the "reference" CPU baseline
// of bench.py (all host threads, same -p fan-out discipline: one work item per file /
// per <=4096-pair chunk).

#include "hash.h"
#include "MinHashHeap.h"
#include "HashList.h"
#include <zlib.h>
#include <stdio.h>
#include <unistd.h>
#include "kseq.h"

#include <atomic>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

KSEQ_INIT(gzFile, gzread)

namespace {

const char COMPL[27] = "TVGHNNCDNNMNKNNNNYSAABWNRN";   // table at Sketch.cpp:1223-1250

// addMinHashes (Sketch.cpp:664-735) restated around the reference's getHash and
// MinHashHeap::tryInsert, keeping its j-scan literally.
void add_min_hashes_ref(MinHashHeap& heap, char* seq, uint64_t length, int k, uint32_t seed,
                        bool use64, bool noncanonical, bool preserveCase, const bool* alphabet)
{
    for (uint64_t i = 0; i < length; i++)
        if (!preserveCase && seq[i] > 96 && seq[i] < 123) seq[i] -= 32;
    char* rev = 0;
    if (!noncanonical) {
        rev = new char[length];
        for (uint64_t i = 0; i < length; i++) {
            int c = (int)seq[length - i - 1] - 'A';
            rev[i] = (c >= 0 && c < 26) ? COMPL[c] : 'N';
        }
    }
    uint64_t j = 0;
    for (uint64_t i = 0; i + k <= length; i++) {
        bool bad = false;
        for (; j < i + k && i + k <= length; j++) {
            if (!alphabet[(unsigned char)seq[j]]) { i = j++; bad = true; break; }
        }
        if (bad) continue;
        if (i + k > length) break;
        const char* f = seq + i;
        const char* r = rev + length - i - k;
        const char* km = (noncanonical || memcmp(f, r, k) <= 0) ? f : r;
        heap.tryInsert(getHash(km, k, seed, use64));
    }
    delete[] rev;
}

}  // namespace

extern "C" {

uint64_t ref_get_hash(const char* seq, int len, uint32_t seed, int use64)
{
    hash_u h = getHash(seq, len, seed, use64 != 0);
    return use64 ? h.hash64 : (uint64_t)h.hash32;
}

uint64_t ref_fp_hash(const uint64_t* tokens, int n, uint32_t seed, int use64)
{
    std::vector<uint64_t> v(tokens, tokens + n);
    hash_u h = getHashFingerPrint(v, n * 8, seed, use64 != 0);
    return use64 ? h.hash64 : (uint64_t)h.hash32;
}

void* ref_heap_new(int use64, uint64_t s, uint64_t mincov) { return new MinHashHeap(use64 != 0, s, mincov); }
void ref_heap_free(void* h) { delete (MinHashHeap*)h; }
void ref_heap_offer(void* h, const uint64_t* hashes, uint64_t n, int use64)
{
    MinHashHeap* H = (MinHashHeap*)h;
    for (uint64_t i = 0; i < n; i++) {
        hash_u u;
        u.hash64 = 0;
        if (use64) u.hash64 = hashes[i]; else u.hash32 = (uint32_t)hashes[i];
        H->tryInsert(u);
    }
}
void ref_heap_add_sequence(void* h, char* seq, uint64_t length, int k, uint32_t seed, int use64,
                           int noncanonical, int preserveCase, const uint8_t* alphabet256)
{
    bool alpha[256];
    for (int i = 0; i < 256; i++) alpha[i] = alphabet256[i] != 0;
    add_min_hashes_ref(*(MinHashHeap*)h, seq, length, k, seed, use64 != 0, noncanonical != 0, preserveCase != 0, alpha);
}
double ref_heap_set_size(void* h) { return ((MinHashHeap*)h)->estimateSetSize(); }
double ref_heap_multiplicity(void* h) { return ((MinHashHeap*)h)->estimateMultiplicity(); }
uint64_t ref_heap_result(void* h, int use64, uint64_t* hashes, uint32_t* counts, uint64_t cap)
{
    HashList list(use64 != 0);
    std::vector<uint32_t> c;
    ((MinHashHeap*)h)->toHashList(list, c);
    uint64_t n = list.size();
    for (uint64_t i = 0; i < n && i < cap; i++) {
        hashes[i] = use64 ? list.at(i).hash64 : (uint64_t)list.at(i).hash32;
        counts[i] = c[i];
    }
    return n;
}

// kseq.h record reader (kseq.h:170-208) exposed for parser parity tests: calls
// cb(user, name, comment_cstr, comment_len, seq, len) per record; returns last kseq code.
typedef void (*ref_record_cb)(void*, const char*, const char*, uint64_t, const char*, int);
int ref_parse_file(const char* path, ref_record_cb cb, void* user)
{
    gzFile fp = gzopen(path, "r");
    if (!fp) return -100;
    kseq_t* ks = kseq_init(fp);
    int l;
    while ((l = kseq_read(ks)) >= 0)
        cb(user, ks->name.s, ks->comment.s ? ks->comment.s : "", ks->comment.l, ks->seq.s, l);
    kseq_destroy(ks);
    gzclose(fp);
    return l;
}

// ---- CPU baseline: sketch ---------------------------------------------------------
// Restates Sketch::initFromFiles' fan-out (Sketch.cpp:353-355: one work item per input,
// -p threads) + sketchFile's per-record loop (Sketch.cpp:1352-1422) for in-memory
// records: genome g is the byte range [offsets[g], offsets[g+1]) of `seqs`, one record.
// out_hashes is [n_genomes][sketch_size] (u64), out_counts likewise (may be null),
// out_n[g] = number of hashes.  Returns total k-mer windows hashed is not tracked here;
// callers count windows themselves.
void ref_sketch_batch(char* seqs, const uint64_t* offsets, uint64_t n_genomes, int k, uint64_t s,
                      uint32_t seed, int use64, int noncanonical, uint64_t mincov, int threads,
                      uint64_t* out_hashes, uint32_t* out_counts, uint64_t* out_n)
{
    bool alpha[256];
    memset(alpha, 0, sizeof alpha);
    alpha['A'] = alpha['C'] = alpha['G'] = alpha['T'] = true;   // alphabetNucleotide, Sketch.h:25
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        for (;;) {
            uint64_t g = next.fetch_add(1);
            if (g >= n_genomes) return;
            MinHashHeap heap(use64 != 0, s, mincov);
            add_min_hashes_ref(heap, seqs + offsets[g], offsets[g + 1] - offsets[g], k, seed,
                               use64 != 0, noncanonical != 0, false, alpha);
            HashList list(use64 != 0);
            std::vector<uint32_t> c;
            heap.toHashList(list, c);                              // setMinHashesForReference, :1291-1297
            uint64_t n = list.size();
            out_n[g] = n;
            for (uint64_t i = 0; i < n; i++) {
                out_hashes[g * s + i] = use64 ? list.at(i).hash64 : (uint64_t)list.at(i).hash32;
                if (out_counts) out_counts[g * s + i] = c[i];
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
}

// ---- CPU baseline: dist -----------------------------------------------------------
// compare()/compareSketches() (CommandDistance.cpp:335-430) over dense panels
// ref[n_ref][s], qry[n_qry][s] with per-sketch sizes, query-major pair order, chunks of
// <=4096 pairs per work item (CommandDistance.cpp:224-261), HashList::at() bounds checks
// kept by going through the reference's HashList.  Emits numer/denom/distance only (the
// p-value's GSL call is not available; the oracle adds it).
void ref_dist_batch(const uint64_t* ref, const uint32_t* ref_n, uint64_t n_ref,
                    const uint64_t* qry, const uint32_t* qry_n, uint64_t n_qry,
                    uint64_t s, int k, int threads, uint32_t* out_numer, uint32_t* out_denom, double* out_dist)
{
    std::vector<HashList> R(n_ref), Q(n_qry);
    for (uint64_t i = 0; i < n_ref; i++) { R[i].setUse64(true); for (uint32_t j = 0; j < ref_n[i]; j++) R[i].push_back64(ref[i * s + j]); }
    for (uint64_t i = 0; i < n_qry; i++) { Q[i].setUse64(true); for (uint32_t j = 0; j < qry_n[i]; j++) Q[i].push_back64(qry[i * s + j]); }
    uint64_t pairs = n_ref * n_qry;
    const uint64_t chunk = 0x1000;
    std::atomic<uint64_t> next(0);
    auto work = [&]() {
        for (;;) {
            uint64_t c0 = next.fetch_add(chunk);
            if (c0 >= pairs) return;
            uint64_t c1 = std::min(pairs, c0 + chunk);
            for (uint64_t p = c0; p < c1; p++) {
                const HashList& a = R[p % n_ref];
                const HashList& b = Q[p / n_ref];
                uint64_t i = 0, j = 0, common = 0, denom = 0;
                while (denom < s && i < (uint64_t)a.size() && j < (uint64_t)b.size()) {
                    if (hashLessThan(a.at(i), b.at(j), true)) i++;
                    else if (hashLessThan(b.at(j), a.at(i), true)) j++;
                    else { i++; j++; common++; }
                    denom++;
                }
                if (denom < s) {
                    if (i < (uint64_t)a.size()) denom += a.size() - i;
                    if (j < (uint64_t)b.size()) denom += b.size() - j;
                    if (denom > s) denom = s;
                }
                double jac = double(common) / denom, d;
                if (common == denom) d = 0;
                else if (common == 0) d = 1.;
                else { d = -log(2 * jac / (1. + jac)) / k; if (d > 1) d = 1; }
                out_numer[p] = (uint32_t)common; out_denom[p] = (uint32_t)denom; out_dist[p] = d;
            }
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++) pool.emplace_back(work);
    for (auto& t : pool) t.join();
}

}  // extern "C"
