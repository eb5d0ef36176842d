// Sketch.cpp -- host mirror of mash/src/mash/Sketch.cpp for the sketch/dist hot path.
// File parsing, naming rules and the .msh container live here; every hash, every bottom-s
// selection and every comparison is done by the CUDA library behind include/fpmash_b200.h.
#include "Sketch.h"

#include <stdio.h>
#include <stdlib.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#include <fstream>
#include <iostream>
#include <atomic>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <thread>
#include <sstream>

#include "fastx.h"
#include <chrono>
#include "fpmash_b200.h"
#include "msh.h"

using namespace std;

static const uint64_t kLimitReadFingerprint = 1000000;   // LIMIT_READ_FINGERPRINT, Sketch.cpp:37
static const uint64_t kFlushBytes = 256ull << 20;        // sequence bytes per GPU batch (the pinned staging buffer is sized for it)
static const uint64_t kGzFlushBytes = 4096ull << 20;     // COMPRESSED bytes per batch of .gz files inflated on the GPU: the parallelism there is
                                                         // one warp per file and a stream's rate is fixed, so a batch should hold thousands of
                                                         // files when there are that many (~14 GB inflated, in HBM only)

void fpmTick(const char* label)
{
    static const bool on = getenv("FPMASH_TIMING") != nullptr;
    if (!on) return;
    static const auto t0 = std::chrono::steady_clock::now();
    static auto last = t0;
    const auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[timing] %8.1f ms (+%7.1f)  %s\n", std::chrono::duration<double, std::milli>(now - t0).count(),
            std::chrono::duration<double, std::milli>(now - last).count(), label);
    last = now;
}

// All GPUs this process may use, created on first use (context creation costs a few hundred ms per device on boxes without
// persistence mode, so jobs below gpuMulti()'s thresholds only ever touch the first one).  FPMASH_DEVICE pins one
// device; FPMASH_GPUS limits (or, beyond the number of devices present, repeats -- tests) the devices used.
static std::vector<int> gpuDeviceList()
{
    std::vector<int> list;
    const int have = fpm_device_count();
    if (const char* dev = getenv("FPMASH_DEVICE")) { list.push_back(atoi(dev)); return list; }
    int want = have;
    if (const char* g = getenv("FPMASH_GPUS")) want = atoi(g);
    if (want < 1) want = 1;
    for (int i = 0; i < want; i++) list.push_back(have > 0 ? i % have : i);
    return list;
}

static fpm_multi* g_multi = nullptr;
static fpm_ctx* g_ctx = nullptr;

fpm_ctx* gpuContext()
{
    if (g_multi) return fpm_multi_ctx(g_multi, 0);
    if (!g_ctx) {
        const std::vector<int> devs = gpuDeviceList();
        fpmTick("before fpm_ctx_create");
        int rc = fpm_ctx_create(devs[0], &g_ctx);
        fpmTick("fpm_ctx_create done");
        if (rc != FPM_OK) {
            cerr << "ERROR: " << fpm_last_error() << endl;
            exit(1);
        }
    }
    return g_ctx;
}

// The several-GPU handle (fpm_multi_*, csrc/dist_multi.cu), or nullptr when this process has one device / the job is too
// small to be worth the other devices' start-up.  The reference spreads the same work over -p threads
// (CommandDistance.cpp:224-261, Sketch.cpp:353-355).
fpm_multi* gpuMulti(uint64_t work /* pairs, or sequence bytes */, uint64_t threshold)
{
    if (g_multi) return g_multi;
    if (const char* e = getenv("FPMASH_MULTI_MIN")) threshold = strtoull(e, nullptr, 10);   // tests: small jobs on several GPUs
    if (work < threshold) return nullptr;
    const std::vector<int> devs = gpuDeviceList();
    if (devs.size() < 2) return nullptr;
    if (g_ctx) { fpm_ctx_destroy(g_ctx); g_ctx = nullptr; }          // its scratch memory would only compete with the handle's own context
    fpmTick("before fpm_multi_create");
    int rc = fpm_multi_create(devs.data(), (int)devs.size(), &g_multi);
    fpmTick("fpm_multi_create done");
    if (rc != FPM_OK) {
        cerr << "ERROR: " << fpm_last_error() << endl;
        exit(1);
    }
    return g_multi;
}

static void gpuCheck(int rc)
{
    if (rc != FPM_OK) {
        cerr << "ERROR: " << fpm_last_error() << endl;
        exit(1);
    }
}

bool hasSuffix(string const& whole, string const& suffix)
{
    return whole.length() >= suffix.length() && 0 == whole.compare(whole.length() - suffix.length(), suffix.length(), suffix);
}

// Sketch.cpp:1260-1289
void setAlphabetFromString(Sketch::Parameters& parameters, const char* characters)
{
    parameters.alphabetSize = 0;
    memset(parameters.alphabet, 0, 256);
    for (const char* c = characters; *c != 0; c++) {
        char upper = *c;
        if (!parameters.preserveCase && upper > 96 && upper < 123) upper -= 32;
        parameters.alphabet[(unsigned char)upper] = true;
    }
    for (int i = 0; i < 256; i++)
        if (parameters.alphabet[i]) parameters.alphabetSize++;
    parameters.use64 = pow(parameters.alphabetSize, parameters.kmerSize) > pow(2, 32);
}

// ------------------------------------------------------------------------------------------
// a batch of sketches on its way to the GPU: sequence bytes (records each followed by 0x00,
// the layout fpm_sketch_batch defines) + per-sketch metadata
// ------------------------------------------------------------------------------------------
// where parsed records go: the pinned GPU batch, or a per-file buffer filled by a parser thread
struct SeqSink {
    virtual void addRecord(const char* s, uint64_t l) = 0;
    virtual void closeGroup(const Sketch::Reference& meta) = 0;
    virtual void expect(uint64_t /*bytes*/) {}   // size hint before a file is parsed
    virtual ~SeqSink() {}
};

struct Sketch::Batch : SeqSink {
    uint8_t* seq = nullptr;      // pinned
    uint64_t used = 0, cap = 0;
    vector<uint64_t> goff{0};
    vector<Reference> metas;

    ~Batch() { if (seq) fpm_host_free(seq); }

    void reserve(uint64_t extra)
    {
        if (used + extra <= cap) return;
        uint64_t ncap = max<uint64_t>(cap ? cap * 2 : (kFlushBytes + (64ull << 20)), used + extra);
        void* p = nullptr;
        gpuCheck(fpm_host_alloc(ncap, &p));
        fpmTick("pinned batch allocated");
        if (used) memcpy(p, seq, used);
        if (seq) fpm_host_free(seq);
        seq = (uint8_t*)p;
        cap = ncap;
    }
    void addRecord(const char* s, uint64_t l) override
    {
        reserve(l + 1);
        memcpy(seq + used, s, l);
        seq[used + l] = 0;
        used += l + 1;
    }
    void closeGroup(const Reference& meta) override
    {
        goff.push_back(used);
        metas.push_back(meta);
    }
    // append a whole parsed file (records already followed by their separators)
    void append(const uint8_t* bytes, uint64_t n, const vector<uint64_t>& offs, const vector<Reference>& ms)
    {
        reserve(n);
        if (n) memcpy(seq + used, bytes, n);
        for (size_t g = 0; g < ms.size(); g++) { goff.push_back(used + offs[g + 1]); metas.push_back(ms[g]); }
        used += n;
    }
    void clear()
    {
        used = 0;
        goff.assign(1, 0);
        metas.clear();
    }
};

static void fillSketchParams(const Sketch::Parameters& p, fpm_sketch_params& sp, bool readsGroup)
{
    memset(&sp, 0, sizeof sp);
    sp.kmer_size = p.kmerSize;
    sp.sketch_size = (uint32_t)p.minHashesPerWindow;
    sp.seed = p.seed;
    sp.min_cov = readsGroup ? p.minCov : 1;                    // Sketch.cpp:1308, 1513
    sp.noncanonical = p.noncanonical;
    sp.preserve_case = p.preserveCase;
    sp.use64 = p.use64;
    sp.want_counts = p.counts;
    for (int i = 0; i < 256; i++) sp.alphabet[i] = p.alphabet[i];
}

void Sketch::flushBatch(Batch& b)
{
    if (b.metas.empty()) return;
    const uint32_t n = (uint32_t)b.metas.size();
    const uint64_t s = parameters.minHashesPerWindow;
    fpm_sketch_params sp;
    fillSketchParams(parameters, sp, parameters.reads);
    vector<uint64_t> hashes((size_t)n * s);
    vector<uint32_t> counts(parameters.counts ? (size_t)n * s : 0);
    vector<uint32_t> outn(n);
    fpmTick("batch parsed");
    // several sketches and enough bytes: whole sketches go to all GPUs (contiguous byte-balanced ranges), else one GPU
    if (fpm_multi* multi = n >= 2 ? gpuMulti(b.used, 96ull << 20) : nullptr)
        gpuCheck(fpm_sketch_batch_multi(multi, &sp, b.seq, b.used, b.goff.data(), n, hashes.data(), parameters.counts ? counts.data() : nullptr, outn.data(), nullptr));
    else
        gpuCheck(fpm_sketch_batch(gpuContext(), &sp, b.seq, b.used, b.goff.data(), n, hashes.data(),
                                  parameters.counts ? counts.data() : nullptr, outn.data(), nullptr));
    fpmTick("fpm_sketch_batch done");
    for (uint32_t g = 0; g < n; g++) {
        Reference& r = b.metas[g];
        r.hashesSorted.setUse64(parameters.use64);
        r.hashesSorted.values.assign(hashes.begin() + (size_t)g * s, hashes.begin() + (size_t)g * s + outn[g]);
        if (parameters.counts) r.counts.assign(counts.begin() + (size_t)g * s, counts.begin() + (size_t)g * s + outn[g]);
        r.countsSorted = true;                                  // setMinHashesForReference, Sketch.cpp:1296
        references.push_back(r);
    }
    b.clear();
}

static void parseSequenceFile(const string& file, const Sketch::Parameters& parameters, SeqSink& sink);

// ------------------------------------------------------------------------------------------
// GPU ingestion of plain FASTA files: the raw bytes of whole files go to the device, the
// parser there (csrc/fasta_parse.cu) does what FastxReader::next does byte by byte, and only the record table
// comes back.  The metadata rules of sketchFile (Sketch.cpp:1318-1422,1436-1444) are applied to that table.
// ------------------------------------------------------------------------------------------
struct Sketch::RawBatch {
    uint8_t* raw = nullptr;      // pinned: file bytes, each file followed by one 0x00 -- or, gz: the compressed files back to back
    uint64_t used = 0, cap = 0;
    bool gz = false;             // the files are gzip streams, inflated on the device (csrc/gunzip.cu); set while the batch is empty
    vector<string> files;
    vector<uint64_t> fileEnd;    // offset of each file's 0x00 (gz: end of its compressed bytes)

    ~RawBatch() { if (raw) fpm_host_free(raw); }
    void reserve(uint64_t extra)
    {
        if (used + extra <= cap) return;
        uint64_t ncap = max<uint64_t>(cap ? cap * 2 : (kFlushBytes + (64ull << 20)), used + extra);
        void* p = nullptr;
        gpuCheck(fpm_host_alloc(ncap, &p));
        fpmTick("pinned raw batch allocated");
        if (used) memcpy(p, raw, used);
        if (raw) fpm_host_free(raw);
        raw = (uint8_t*)p;
        cap = ncap;
    }
    // Whole files -> batch, read by up to `threads` threads straight into the pinned buffer.  Returns how many of
    // them (a prefix of `names`) were appended: it stops before the first file that cannot go to the GPU parser
    // (gzip behind a plain name, unreadable, not a regular file, contains 0x00, changed size while being read).
    size_t addFiles(const vector<string>& names, int threads)
    {
        const size_t n = names.size();
        vector<uint64_t> size(n, 0), off(n, 0);
        vector<char> ok(n, 1);
        uint64_t total = 0;
        for (size_t i = 0; i < n; i++) {
            struct stat st;
            if (stat(names[i].c_str(), &st) != 0 || !S_ISREG(st.st_mode)) { ok[i] = 0; continue; }
            size[i] = (uint64_t)st.st_size;
            off[i] = used + total;
            total += size[i] + (gz ? 0 : 1);
        }
        reserve(total);
        auto readOne = [&](size_t i) {
            if (!ok[i]) return;
            FILE* f = fopen(names[i].c_str(), "rb");
            if (!f) { ok[i] = 0; return; }
            uint8_t* dst = raw + off[i];
            const uint64_t got = size[i] ? fread(dst, 1, size[i], f) : 0;
            const bool more = fgetc(f) != EOF;
            fclose(f);
            if (got != size[i] || more) { ok[i] = 0; return; }
            const bool magic = size[i] >= 2 && dst[0] == 0x1f && dst[1] == 0x8b;
            if (gz) { if (!magic) ok[i] = 0; return; }                                          // zlib reads anything else through unchanged: host reader
            if (magic) { ok[i] = 0; return; }                                                  // gzip behind a plain name
            if (size[i] && memchr(dst, 0, size[i]) != nullptr) { ok[i] = 0; return; }
            dst[size[i]] = 0;
        };
        const int nt = (int)min<size_t>((size_t)max(threads, 1), n);
        if (nt <= 1) {
            for (size_t i = 0; i < n; i++) readOne(i);
        } else {
            atomic<size_t> next(0);
            vector<thread> pool;
            for (int t = 0; t < nt; t++)
                pool.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) readOne(i); });
            for (auto& t : pool) t.join();
        }
        size_t good = 0;
        while (good < n && ok[good]) good++;
        for (size_t i = 0; i < good; i++) {
            files.push_back(names[i]);
            fileEnd.push_back(off[i] + size[i]);
            used = off[i] + size[i] + (gz ? 0 : 1);
        }
        return good;
    }
    void clear() { used = 0; files.clear(); fileEnd.clear(); }
};

namespace {
// name / comment of one header, as FastxReader::next reads them (kseq.h:175-182): [p, e) is the text after '>',
// up to but excluding the '\n' that ended it (or up to the end of the file when atEof).
void parseHeader(const uint8_t* p, const uint8_t* e, bool atEof, string& name, string& comment, string& commentCstr)
{
    name.clear();
    comment.clear();
    const uint8_t* q = p;
    while (q < e && !isspace(*q)) q++;
    name.assign((const char*)p, (size_t)(q - p));
    if (q == e) return;                                    // delimiter was the '\n' (or the end of the file): no comment, buffer untouched
    q++;                                                   // the delimiter
    if (q == e && atEof) return;                           // nothing after it before the end of the file: kseq leaves the buffer untouched
    comment.assign((const char*)q, (size_t)(e - q));
    commentCstr = comment;
}
}  // namespace

void Sketch::flushRawBatch(RawBatch& rb, Batch& hostBatch)
{
    if (rb.files.empty()) return;
    fpmTick("raw batch read");
    uint64_t nRec = 0, seqBytes = 0;
    int status = 0;
    if (rb.gz) {
        // gzip streams: inflated on the device (one warp per file); the raw batch never exists on the host
        vector<uint64_t> gzOff{0};
        gzOff.insert(gzOff.end(), rb.fileEnd.begin(), rb.fileEnd.end());
        uint64_t total = 0;
        gpuCheck(fpm_gunzip_batch(gpuContext(), rb.raw, gzOff.data(), (uint32_t)rb.files.size(), rb.fileEnd.data(), &total, &status));
        fpmTick("gzip batch inflated on the GPU");
        if (status == FPM_GUNZIP_OK) gpuCheck(fpm_fasta_parse(gpuContext(), nullptr, total, &nRec, &seqBytes, &status));
        else status = FPM_FASTA_NOT_PLAIN;                 // a damaged or unusual stream: zlib (the host reader) defines what the reference reads
    } else {
        gpuCheck(fpm_fasta_parse(gpuContext(), rb.raw, rb.used, &nRec, &seqBytes, &status));
    }
    if (status != FPM_FASTA_OK) {
        // not plain FASTA somewhere in this batch (FASTQ records): the host reader defines the result
        for (const string& f : rb.files) parseSequenceFile(f, parameters, hostBatch);
        flushBatch(hostBatch);
        rb.clear();
        return;
    }
    vector<fpm_fasta_record> recs(nRec);
    if (nRec) gpuCheck(fpm_fasta_records(gpuContext(), recs.data()));
    // header text: in the pinned batch, or (gz) gathered from the device
    vector<uint8_t> hdrBytes;
    vector<uint64_t> hdrOff;
    if (rb.gz) {
        hdrOff.assign(nRec + 1, 0);
        for (uint64_t i = 0; i < nRec; i++) hdrOff[i + 1] = hdrOff[i] + (recs[i].hdr_end - recs[i].hdr_begin);
        hdrBytes.resize(hdrOff[nRec] + 1);
        if (nRec) gpuCheck(fpm_fasta_headers(gpuContext(), hdrOff.data(), hdrBytes.data()));
    }
    fpmTick("fasta parsed on the GPU");
    vector<uint64_t> goff{0};
    vector<Reference> metas;
    vector<char> keep;                                     // -i: groups of records shorter than k are dropped after sketching
    string name, comment, commentCstr;
    size_t r = 0;
    for (size_t f = 0; f < rb.files.size(); f++) {
        const uint64_t fEnd = rb.fileEnd[f];
        commentCstr.clear();                               // one reader (one kseq_t) per file
        Reference reference;
        reference.length = 0;
        reference.name = rb.files[f];
        uint64_t count = 0;
        bool skipped = false;
        for (; r < nRec && recs[r].hdr_begin < fEnd; r++) {
            const uint64_t seqEnd = r + 1 < nRec ? recs[r + 1].seq_begin : seqBytes;
            const uint64_t l = seqEnd - recs[r].seq_begin - 1;
            const bool atEof = recs[r].hdr_end == fEnd;
            if (atEof && recs[r].hdr_end == recs[r].hdr_begin + 1) {          // a lone '>' as the file's last byte: no record for kseq
                if (!parameters.concatenated) { goff.push_back(seqEnd); keep.push_back(0); metas.emplace_back(); }
                continue;
            }
            if (rb.gz) parseHeader(hdrBytes.data() + hdrOff[r] + 1, hdrBytes.data() + hdrOff[r + 1], atEof, name, comment, commentCstr);
            else parseHeader(rb.raw + recs[r].hdr_begin + 1, rb.raw + recs[r].hdr_end, atEof, name, comment, commentCstr);
            if (parameters.concatenated) {
                if (l < (uint64_t)parameters.kmerSize) { skipped = true; continue; }
                if (count == 0) {
                    reference.comment = name;
                    reference.comment.append(" ");
                    reference.comment.append(commentCstr.c_str());            // read as a C string (Sketch.cpp:1383-1392)
                }
                count++;
                reference.length += l;
            } else {
                Reference one;
                one.length = l;
                one.name = name;
                one.comment = comment;
                goff.push_back(seqEnd);
                keep.push_back(l >= (uint64_t)parameters.kmerSize);
                metas.push_back(one);
            }
        }
        if (parameters.concatenated) {
            if (count > 1) {                                                  // Sketch.cpp:1436-1444
                reference.comment.insert(0, " seqs] ");
                reference.comment.insert(0, to_string(count));
                reference.comment.insert(0, "[");
                reference.comment.append(" [...]");
            }
            if (reference.length == 0) {
                if (skipped) cerr << "\nWARNING: All fasta records in input files were shorter than the k-mer size (" << parameters.kmerSize << ")." << endl;
                else cerr << "\nERROR: Did not find fasta records in \"input files\"." << endl;
                exit(1);
            }
            const uint64_t gEnd = r < nRec ? recs[r].seq_begin : seqBytes;   // this file's records end where the next file's begin
            goff.push_back(gEnd);
            keep.push_back(1);
            metas.push_back(reference);
        }
    }
    if (goff.back() != seqBytes) {                          // files without any record leave no bytes: cannot happen, but never trust a table
        cerr << "ERROR: inconsistent FASTA record table." << endl;
        exit(1);
    }
    const uint32_t n = (uint32_t)metas.size();
    if (n) {
        const uint64_t s = parameters.minHashesPerWindow;
        fpm_sketch_params sp;
        fillSketchParams(parameters, sp, false);
        vector<uint64_t> hashes((size_t)n * s);
        vector<uint32_t> counts(parameters.counts ? (size_t)n * s : 0);
        vector<uint32_t> outn(n);
        gpuCheck(fpm_sketch_parsed(gpuContext(), &sp, goff.data(), n, hashes.data(), parameters.counts ? counts.data() : nullptr, outn.data()));
        fpmTick("fpm_sketch_parsed done");
        for (uint32_t g = 0; g < n; g++) {
            if (!keep[g]) continue;
            Reference& ref = metas[g];
            ref.hashesSorted.setUse64(parameters.use64);
            ref.hashesSorted.values.assign(hashes.begin() + (size_t)g * s, hashes.begin() + (size_t)g * s + outn[g]);
            if (parameters.counts) ref.counts.assign(counts.begin() + (size_t)g * s, counts.begin() + (size_t)g * s + outn[g]);
            ref.countsSorted = true;
            references.push_back(ref);
        }
    }
    rb.clear();
}

// sketchFile's reading loop (Sketch.cpp:1318-1422) for one sketch made of one or more files:
// records are taken one per file in rotation, a record shorter than k is skipped WITHOUT
// advancing the rotation, name/comment come from the first valid record.
static void readGroup(const vector<string>& fileNames, const Sketch::Parameters& parameters, Sketch::Reference& reference,
                      uint64_t& count, bool& skipped, SeqSink& sink)
{
    int fileCount = (int)fileNames.size();
    vector<gzFile> fps(fileCount);
    vector<unique_ptr<FastxReader>> readers;
    reference.length = 0;
    for (int f = 0; f < fileCount; f++) {
        if (fileNames[f] == "-") {
            if (f > 1) {
                cerr << "ERROR: '-' for stdin must be first input" << endl;
                exit(1);
            }
            fps[f] = gzdopen(fileno(stdin), "r");
        } else {
            if (reference.name == "" && fileNames[f] != "-") reference.name = fileNames[f];
            fps[f] = gzopen(fileNames[f].c_str(), "r");
            if (fps[f] == 0) {
                cerr << "ERROR: could not open " << fileNames[f] << endl;
                exit(1);
            }
        }
        gzbuffer(fps[f], 1 << 17);   // smaller than the reader's 1 MB requests: plain files are then read directly
        readers.emplace_back(new FastxReader(fps[f]));
    }
    int64_t l = -1;
    size_t it = 0;
    count = 0;
    skipped = false;
    while (!readers.empty()) {
        FastxReader& rd = *readers[it];
        l = rd.next();
        if (l < -1) break;                                       // error
        if (l == -1) {                                           // eof
            readers.erase(readers.begin() + it);
            if (it == readers.size()) it = 0;
            continue;
        }
        if (l < parameters.kmerSize) {
            skipped = true;
            continue;
        }
        if (count == 0) {
            if (fileNames[0] == "-") {
                reference.name = rd.name;
                reference.comment = rd.comment_cstr;
            } else {
                reference.comment = rd.name;
                reference.comment.append(" ");
                reference.comment.append(rd.comment_cstr);
            }
        }
        count++;
        if (!parameters.reads) reference.length += l;
        sink.addRecord(rd.seq.data(), (uint64_t)l);
        it++;
        if (it == readers.size()) it = 0;
    }
    for (int i = 0; i < fileCount; i++) gzclose(fps[i]);
    if (count > 1) {                                             // Sketch.cpp:1436-1444
        reference.comment.insert(0, " seqs] ");
        reference.comment.insert(0, to_string(count));
        reference.comment.insert(0, "[");
        reference.comment.append(" [...]");
    }
    if (l != -1) {
        cerr << "\nERROR: reading input files." << endl;
        exit(1);
    }
}


// One input file -> records in `sink`: one sketch per file (sketchFile, Sketch.cpp:1299-1488) or, with -i, one
// per record (sketchFileBySequence + sketchSequence, Sketch.cpp:478-522, 1490-1517).  Errors print and exit
// like the reference's workers do.
static void parseSequenceFile(const string& file, const Sketch::Parameters& parameters, SeqSink& sink)
{
    struct stat st;
    if (file != "-" && stat(file.c_str(), &st) == 0 && !hasSuffix(file, ".gz")) sink.expect((uint64_t)st.st_size);
    if (parameters.concatenated) {
        Sketch::Reference reference;
        uint64_t count;
        bool skipped;
        readGroup(vector<string>(1, file), parameters, reference, count, skipped, sink);
        if (reference.length == 0) {
            if (skipped) cerr << "\nWARNING: All fasta records in input files were shorter than the k-mer size (" << parameters.kmerSize << ")." << endl;
            else cerr << "\nERROR: Did not find fasta records in \"input files\"." << endl;
            exit(1);
        }
        sink.closeGroup(reference);
        return;
    }
    gzFile fp = file == "-" ? gzdopen(fileno(stdin), "r") : gzopen(file.c_str(), "r");
    if (!fp) {
        cerr << "ERROR: could not open " << file << " for reading." << endl;
        exit(1);
    }
    gzbuffer(fp, 1 << 17);
    FastxReader rd(fp);
    int64_t l;
    while ((l = rd.next()) >= 0) {
        if (l < parameters.kmerSize) continue;
        Sketch::Reference reference;
        reference.length = l;
        reference.name = rd.name;
        reference.comment = rd.comment;
        sink.addRecord(rd.seq.data(), (uint64_t)l);
        sink.closeGroup(reference);
    }
    gzclose(fp);
    if (l != -1) {
        cerr << "\nERROR: reading " << file << "." << endl;
        exit(1);
    }
}

// A parsed file held in ordinary memory until the main thread splices it into the GPU batch.
struct FileBuffer : SeqSink {
    vector<uint8_t> seq;
    vector<uint64_t> goff{0};
    vector<Sketch::Reference> metas;
    void expect(uint64_t bytes) override { seq.reserve(seq.size() + bytes + 16); }
    void addRecord(const char* s, uint64_t l) override
    {
        seq.insert(seq.end(), (const uint8_t*)s, (const uint8_t*)s + l);
        seq.push_back(0);
    }
    void closeGroup(const Sketch::Reference& meta) override
    {
        goff.push_back(seq.size());
        metas.push_back(meta);
    }
};

// -p: files are parsed by a pool of threads (the reference's ThreadPool fans out whole files the same way,
// Sketch.cpp:353-355) with a bounded look-ahead; results are consumed strictly in submission order.
class ParsePool {
public:
    ParsePool(const vector<string>& jobFiles, const Sketch::Parameters& p, int nThreads)
        : files(jobFiles), parameters(p), slots(jobFiles.size()), ready(jobFiles.size(), 0), window((size_t)nThreads * 4)
    {
        for (int t = 0; t < nThreads; t++) threads.emplace_back([this] { work(); });
    }
    ~ParsePool()
    {
        {
            lock_guard<mutex> lk(m);
            stop = true;
        }
        cvWork.notify_all();
        for (auto& t : threads) t.join();
    }
    unique_ptr<FileBuffer> take(size_t j)
    {
        unique_lock<mutex> lk(m);
        cvDone.wait(lk, [&] { return ready[j] != 0; });
        unique_ptr<FileBuffer> r = move(slots[j]);
        consumed = j + 1;
        lk.unlock();
        cvWork.notify_all();
        return r;
    }

private:
    void work()
    {
        for (;;) {
            size_t j;
            {
                unique_lock<mutex> lk(m);
                cvWork.wait(lk, [&] { return stop || (next < files.size() && next < consumed + window); });
                if (stop || next >= files.size()) return;
                j = next++;
            }
            unique_ptr<FileBuffer> fb(new FileBuffer());
            parseSequenceFile(files[j], parameters, *fb);
            {
                lock_guard<mutex> lk(m);
                slots[j] = move(fb);
                ready[j] = 1;
            }
            cvDone.notify_all();
        }
    }
    const vector<string>& files;
    const Sketch::Parameters& parameters;
    vector<unique_ptr<FileBuffer>> slots;
    vector<char> ready;
    size_t window, next = 0, consumed = 0;
    bool stop = false;
    mutex m;
    condition_variable cvWork, cvDone;
    vector<thread> threads;
};

static void checkUnsupported(const Sketch::Parameters& p)
{
    if (p.windowed) { cerr << "ERROR: windowed sketching is not part of the accelerated path." << endl; exit(1); }
    if (p.memoryBound > 0) { cerr << "ERROR: the Bloom-filter option (-b) is not part of the accelerated path; use -m for exact filtering." << endl; exit(1); }
    if (p.targetCov > 0) { cerr << "ERROR: target coverage (-c) is inherently sequential and not part of the accelerated path." << endl; exit(1); }
}

// ------------------------------------------------------------------------------------------

void Sketch::getAlphabetAsString(string& alphabet) const
{
    for (int i = 0; i < 256; i++)
        if (parameters.alphabet[i]) alphabet.append(1, (char)i);
}

int Sketch::getMinKmerSize(uint64_t reference) const
{
    return ceil(log(references[reference].length * (1 - parameters.warning) / parameters.warning) / log(parameters.alphabetSize));
}

double Sketch::getRandomKmerChance(uint64_t reference) const
{
    return 1. / (kmerSpace / references[reference].length + 1.);
}

void Sketch::getReferenceHistogram(uint64_t index, map<uint32_t, uint64_t>& histogram) const
{
    const Reference& reference = references.at(index);
    histogram.clear();
    for (uint64_t i = 0; i < reference.counts.size(); i++) histogram[reference.counts.at(i)]++;
}

uint64_t Sketch::getReferenceIndex(string id) const
{
    auto it = referenceIndecesById.find(id);
    return it == referenceIndecesById.end() ? (uint64_t)-1 : (uint64_t)it->second;
}

void Sketch::createIndex()   // Sketch.cpp:644-662
{
    for (size_t i = 0; i < references.size(); i++) referenceIndecesById[references[i].id] = (int)i;
    kmerSpace = pow(parameters.alphabetSize, parameters.kmerSize);
}

// Sketch::initFromFingerprints (Sketch.cpp:56-151): one hash per line, appended in line order
// (never sorted, truncated or deduplicated); consecutive equal ids form one Reference.
void Sketch::initFromFingerprints(const vector<string>& files, const Parameters& parametersNew)
{
    parameters = parametersNew;
    uint64_t counterLine = 0;
    string lastID = "";
    vector<uint64_t> tokens, lineOff{0};
    vector<size_t> lineRef;                 // reference index of every hashed line

    cout << "Initializing from fingerprints..." << endl;
    for (const string& file : files) {
        cout << "Processing file: " << file << endl;
        ifstream inputFile(file);
        if (!inputFile) {
            cerr << "ERROR: Could not open fingerprint file " << file << " for reading." << endl;
            exit(1);
        }
        string line;
        bool haveCurrent = false;           // the reference code keeps a per-file current pointer
        while (getline(inputFile, line) && counterLine < kLimitReadFingerprint) {
            counterLine++;
            istringstream ss(line);
            uint64_t number;
            string id;
            ss >> id;
            size_t first = tokens.size();
            while (ss >> number) {
                tokens.push_back(number);
                if (ss.peek() == ' ') ss.ignore();
            }
            size_t n = tokens.size() - first;
            if (id != lastID) {
                Reference r;
                r.id = id;
                r.length = n;
                r.name = "" + id;
                r.comment = "FingerPrint : " + r.id;
                r.hashesSorted.setUse64(parameters.use64);
                references.push_back(r);
                lastID = id;
                haveCurrent = true;
            } else if (!haveCurrent) {
                // The reference dereferences a null pointer here (a file that starts with an empty
                // line or with the id the previous file ended on).
                cerr << "ERROR: fingerprint file " << file << " starts with an empty id or with the id that ended the previous file." << endl;
                exit(1);
            }
            references.back().length += n;
            lineOff.push_back(tokens.size());
            lineRef.push_back(references.size() - 1);
        }
    }
    const uint64_t nLines = lineRef.size();
    if (nLines) {
        vector<uint64_t> hashes(nLines);
        gpuCheck(fpm_fp_hash_batch(gpuContext(), tokens.data(), lineOff.data(), nLines, parameters.seed, parameters.use64, hashes.data()));
        for (uint64_t i = 0; i < nLines; i++) references[lineRef[i]].hashesSorted.add(hashes[i]);
    }
    createIndex();
    cout << "Initialization complete." << endl;
}

// `mash sketch -r reads.fastq` (or reads.fa) without the host reader: pieces, each cut at a record boundary, go to
// fpm_fastq_stream_append (four-line FASTQ) or to fpm_fasta_parse + fpm_sketch_stream_append_parsed (FASTA reads).
// Returns false (nothing usable was produced) if the file is neither clean four-line FASTQ nor plain FASTA.
// Metadata as sketchFile keeps it (Sketch.cpp:1380-1393,1436-1444): comment from the first read of at least k bases
// (with kseq's stale-comment quirk, so the headers before it are walked too), count of such reads.
static bool readsViaGpu(const string& file, const Sketch::Parameters& parameters, Sketch::Reference& reference, uint64_t& count, bool& skipped)
{
    FILE* f = fopen(file.c_str(), "rb");
    if (!f) return false;
    struct stat st;
    if (fstat(fileno(f), &st) != 0 || !S_ISREG(st.st_mode)) { fclose(f); return false; }
    const uint64_t fileSize = (uint64_t)st.st_size;
    const int fd = fileno(f);
    const char* pieceEnv = getenv("FPMASH_FASTQ_PIECE");                                  // tests: small pieces exercise the boundary search
    // 64 MB pieces: page-locking the staging buffer costs ~0.2 ms per MB, a piece takes a few ms on the device
    const uint64_t kPiece = pieceEnv && atoll(pieceEnv) >= 4096 ? (uint64_t)atoll(pieceEnv) : 64ull << 20;
    void* pinned = nullptr;
    gpuCheck(fpm_host_alloc(kPiece + 64, &pinned));
    uint8_t* buf = (uint8_t*)pinned;
    uint64_t used = 0, fileOff = 0;
    bool eof = false, ok = true, first = true, haveComment = false, fasta = false;
    vector<fpm_fasta_record> recs;
    string name, comment, commentCstr;
    reference.name = file;
    reference.length = 0;
    count = 0;
    skipped = false;
    vector<uint64_t> ends;
    while (ok && !(eof && used == 0)) {
        if (!eof) {
            // the next stretch of the file, read by the -p threads straight into the pinned piece
            const uint64_t want = min<uint64_t>(kPiece - used, fileSize - fileOff);
            const int nt = (int)max<uint64_t>(1, min<uint64_t>((uint64_t)max(parameters.parallelism, 1), want >> 20));
            atomic<bool> readOk(true);
            auto readPart = [&](int t) {
                uint64_t a = want * t / nt, b = want * (t + 1) / nt;
                while (a < b) {
                    const ssize_t r = pread(fd, buf + used + a, b - a, (off_t)(fileOff + a));
                    if (r <= 0) { readOk = false; return; }
                    a += (uint64_t)r;
                }
            };
            if (nt == 1) readPart(0);
            else {
                vector<thread> pool;
                for (int t = 0; t < nt; t++) pool.emplace_back(readPart, t);
                for (auto& t : pool) t.join();
            }
            if (!readOk) { ok = false; break; }
            used += want;
            fileOff += want;
            if (fileOff >= fileSize) eof = true;
        }
        if (first && used >= 2 && buf[0] == 0x1f && buf[1] == 0x8b) { ok = false; break; }     // gzip behind a plain name
        if (first) {
            // kseq looks for the first '>' or '@' (kseq.h:172): that byte decides which parser takes the file
            const uint8_t* gt = (const uint8_t*)memchr(buf, '>', used);
            const uint8_t* at = (const uint8_t*)memchr(buf, '@', used);
            fasta = gt && (!at || gt < at);
        }
        first = false;
        if (fasta) {
            // ---- FASTA reads: pieces cut at "\n>" (always a record start), parsed by fpm_fasta_parse, appended device to device ----
            uint64_t cut = used;
            if (!eof) {
                cut = 0;
                for (uint64_t p = used; p > 1; p--)
                    if (buf[p - 1] == '>' && buf[p - 2] == '\n') { cut = p - 1; break; }
                if (cut == 0) { ok = false; break; }                                          // one record larger than a piece: host reader
            }
            if (cut == 0) break;
            if (memchr(buf, 0, cut) != nullptr) { ok = false; break; }                        // 0x00 separates files for the device parser
            const uint8_t saved = buf[cut];
            buf[cut] = 0;
            uint64_t nRec = 0, seqBytes = 0;
            int status = 0;
            gpuCheck(fpm_fasta_parse(gpuContext(), buf, cut + 1, &nRec, &seqBytes, &status));
            buf[cut] = saved;
            if (status != FPM_FASTA_OK) { ok = false; break; }
            recs.resize(nRec);
            if (nRec) gpuCheck(fpm_fasta_records(gpuContext(), recs.data()));
            gpuCheck(fpm_sketch_stream_append_parsed(gpuContext()));
            for (uint64_t r = 0; r < nRec; r++) {
                const uint64_t seqEnd = r + 1 < nRec ? recs[r + 1].seq_begin : seqBytes;
                const uint64_t l = seqEnd - recs[r].seq_begin - 1;
                const bool atEof = recs[r].hdr_end == cut;                                    // header ended by the end of the file, not by '\n'
                if (atEof && recs[r].hdr_end == recs[r].hdr_begin + 1) continue;              // a lone '>' as the last byte: no record for kseq
                if (!haveComment) parseHeader(buf + recs[r].hdr_begin + 1, buf + recs[r].hdr_end, atEof, name, comment, commentCstr);
                if (l < (uint64_t)parameters.kmerSize) { skipped = true; continue; }
                if (!haveComment) {
                    reference.comment = name;
                    reference.comment.append(" ");
                    reference.comment.append(commentCstr.c_str());
                    haveComment = true;
                }
                count++;
            }
            memmove(buf, buf + cut, used - cut);
            used -= cut;
            continue;
        }
        uint64_t cut = used;
        if (eof) {
            if (used && buf[used - 1] != '\n') buf[used++] = '\n';                            // a last line without newline parses the same
            cut = used;
        } else {
            // the last header line that is followed by a complete record: line starts s with buf[s] == '@' whose line
            // two further down starts with '+' (a quality line starting with '@' is followed, two lines down, by a
            // sequence line, and clean sequence lines never start with '+')
            uint64_t starts[12];
            int ns = 0;
            for (uint64_t p = used; p > 0 && ns < 12; p--)
                if (buf[p - 1] == '\n') starts[ns++] = p;                                      // starts[0] is the latest line start
            cut = 0;
            for (int i = 4; i < ns; i++)
                if (starts[i] < used && buf[starts[i]] == '@' && buf[starts[i - 2]] == '+') { cut = starts[i]; break; }
            if (cut == 0) { ok = false; break; }                                              // no record boundary in sight
        }
        if (cut == 0) break;
        int status = 0;
        uint64_t info[FPM_FASTQ_INFO_WORDS];
        gpuCheck(fpm_fastq_stream_append(gpuContext(), buf, cut, (uint32_t)parameters.kmerSize, &status, info));
        if (status != FPM_FASTA_OK) { ok = false; break; }
        const uint64_t nReads = info[0], nValid = info[1], firstValid = info[3];
        if (!haveComment && nReads) {
            // walk the headers up to the first read that counts (all of them if none does: the stale comment carries on)
            const uint64_t upto = nValid ? firstValid : nReads - 1;
            ends.resize(4 * upto + 1);
            gpuCheck(fpm_fastq_line_ends(gpuContext(), 0, 4 * upto + 1, ends.data()));
            for (uint64_t r = 0; r <= upto; r++) {
                const uint64_t b = r ? ends[4 * r - 1] + 1 : 0, e = ends[4 * r];
                parseHeader(buf + b + 1, buf + e, false, name, comment, commentCstr);
            }
            if (nValid) {
                reference.comment = name;
                reference.comment.append(" ");
                reference.comment.append(commentCstr.c_str());
                haveComment = true;
            }
        }
        count += nValid;
        skipped |= nReads > nValid;
        memmove(buf, buf + cut, used - cut);
        used -= cut;
    }
    fclose(f);
    fpm_host_free(pinned);
    if (!ok) return false;
    fpmTick("reads parsed on the GPU");
    if (count > 1) {                                                                          // Sketch.cpp:1436-1444
        reference.comment.insert(0, " seqs] ");
        reference.comment.insert(0, to_string(count));
        reference.comment.insert(0, "[");
        reference.comment.append(" [...]");
    }
    return true;
}

void Sketch::initFromReads(const vector<string>& files, const Parameters& parametersNew)   // Sketch.cpp:203-210
{
    parameters = parametersNew;
    checkUnsupported(parameters);
    // A read set becomes ONE sketch and can be far larger than a staging buffer: records are streamed to the
    // GPU in pinned pieces (fpm_sketch_stream_*), accumulate in HBM and are sketched once at the end.
    struct StreamSink : SeqSink {
        // two pinned staging buffers filled in turn: while one crosses PCIe (fpm_sketch_stream_append_async) the reader fills the other
        uint8_t* stages[2] = {nullptr, nullptr};
        uint64_t tickets[2] = {0, 0};
        bool inFlight[2] = {false, false};
        int cur = 0;
        uint8_t* stage = nullptr;
        uint64_t used = 0;
        const uint64_t kStage = 32ull << 20;
        StreamSink() { gpuCheck(fpm_sketch_stream_begin(gpuContext())); }
        ~StreamSink()
        {
            for (int i = 0; i < 2; i++) {
                if (inFlight[i]) fpm_sketch_stream_wait(gpuContext(), tickets[i]);
                if (stages[i]) fpm_host_free(stages[i]);
            }
        }
        void push()
        {
            if (!used) return;
            gpuCheck(fpm_sketch_stream_append_async(gpuContext(), stage, used, &tickets[cur]));
            inFlight[cur] = true;
            used = 0;
            cur ^= 1;                                                       // go on in the other buffer, once its last copy is through
            if (inFlight[cur]) { gpuCheck(fpm_sketch_stream_wait(gpuContext(), tickets[cur])); inFlight[cur] = false; }
            if (!stages[cur]) { void* p = nullptr; gpuCheck(fpm_host_alloc(kStage, &p)); stages[cur] = (uint8_t*)p; }
            stage = stages[cur];
        }
        void addRecord(const char* s, uint64_t l) override
        {
            if (!stage) { void* p = nullptr; gpuCheck(fpm_host_alloc(kStage, &p)); stages[cur] = stage = (uint8_t*)p; }   // only the host-reader route needs it
            uint64_t done = 0;                                // records longer than the stage go through in pieces
            while (done < l) {
                if (used == kStage) push();
                uint64_t n = min<uint64_t>(l - done, kStage - used);
                memcpy(stage + used, s + done, n);
                used += n; done += n;
            }
            if (used == kStage) push();
            stage[used++] = 0;
        }
        void closeGroup(const Sketch::Reference&) override { push(); gpuCheck(fpm_sketch_stream_end_group(gpuContext())); }
    } sink;
    Reference reference;
    uint64_t count = 0;
    bool skipped = false;
    // One plain FASTQ file: its four-line records are parsed on the GPU, piece by piece, straight into the HBM
    // stream (csrc/fasta_parse.cu).  Anything else -- several files feeding one sketch (their records interleave,
    // Sketch.cpp:1352-1422), gzip, stdin, FASTQ that is not four lines per record (LF or CRLF) -- takes the host reader; if the
    // GPU route gives up half way the stream is restarted from scratch.  (FASTA reads: the FASTA parser, piece by piece.)
    const char* gpuParseEnv = getenv("FPMASH_GPU_PARSE");
    bool viaGpu = !(gpuParseEnv && gpuParseEnv[0] == '0') && files.size() == 1 && files[0] != "-" && !hasSuffix(files[0], ".gz");
    if (viaGpu) {
        viaGpu = readsViaGpu(files[0], parameters, reference, count, skipped);
        if (!viaGpu) {
            gpuCheck(fpm_sketch_stream_begin(gpuContext()));
            reference = Reference();
        }
    }
    if (!viaGpu) readGroup(files, parameters, reference, count, skipped, sink);
    sink.closeGroup(reference);
    {
        const uint64_t s = parameters.minHashesPerWindow;
        fpm_sketch_params sp;
        fillSketchParams(parameters, sp, parameters.reads);
        vector<uint64_t> hashes(s);
        vector<uint32_t> counts(s), outn(1);
        gpuCheck(fpm_sketch_stream_finish(gpuContext(), &sp, hashes.data(), parameters.counts ? counts.data() : nullptr, outn.data(), nullptr));
        reference.hashesSorted.setUse64(parameters.use64);
        reference.hashesSorted.values.assign(hashes.begin(), hashes.begin() + outn[0]);
        if (parameters.counts) reference.counts.assign(counts.begin(), counts.begin() + outn[0]);
        reference.countsSorted = true;
        references.push_back(reference);
    }
    Reference& ref = references.back();
    // estimateSetSize / estimateMultiplicity (MinHashHeap.h:44-45)
    double setSize = 0, multiplicity = 0;
    size_t n = ref.hashesSorted.size();
    if (n) {
        setSize = pow(2.0, parameters.use64 ? 64.0 : 32.0) * (double)n / (double)ref.hashesSorted.at(n - 1);
        uint64_t sum = 0;
        for (uint32_t c : ref.counts) sum += c;
        multiplicity = (double)sum / n;
    }
    if (parameters.reads) ref.length = parameters.genomeSize != 0 ? parameters.genomeSize : (uint64_t)setSize;   // Sketch.cpp:1424-1434
    if (ref.length == 0) {
        if (skipped) cerr << "\nWARNING: All fasta records in input files were shorter than the k-mer size (" << parameters.kmerSize << ")." << endl;
        else cerr << "\nERROR: Did not find fasta records in \"input files\"." << endl;
        exit(1);
    }
    if (parameters.reads) {
        cerr << "Estimated genome size: " << setSize << endl;
        cerr << "Estimated coverage:    " << multiplicity << endl;
    }
    createIndex();
}

void Sketch::loadSketchFile(const string& file)   // loadCapnp, Sketch.cpp:1059-1219
{
    vector<uint8_t> bytes;
    if (!msh::read_file(file, bytes)) return;
    msh::File mf;
    string err;
    if (!msh::decode(bytes.data(), bytes.size(), parameters.use64, parameters.minHashesPerWindow, mf, err)) {
        cerr << "ERROR: " << file << " is not a valid sketch file (" << err << ")." << endl;
        exit(1);
    }
    for (msh::RefRecord& r : mf.refs) {
        Reference ref;
        ref.name = r.name;
        ref.comment = r.comment;
        ref.length = r.length;
        ref.hashesSorted.setUse64(parameters.use64);
        ref.hashesSorted.values.swap(r.hashes);
        if (r.has_counts) ref.counts.swap(r.counts);
        ref.countsSorted = r.counts_sorted;
        references.push_back(ref);
    }
}

uint64_t Sketch::initParametersFromCapnp(const char* file)   // Sketch.cpp:401-470
{
    vector<uint8_t> bytes;
    if (!msh::read_file(file, bytes)) {
        cerr << "ERROR: could not open \"" << file << "\" for reading." << endl;
        exit(1);
    }
    msh::Header h;
    uint64_t referenceCount = 0;
    bool firstHasCounts = false;
    string err;
    if (!msh::decode_header(bytes.data(), bytes.size(), h, referenceCount, firstHasCounts, err)) {
        cerr << "ERROR: \"" << file << "\" is not a valid sketch file (" << err << ")." << endl;
        exit(1);
    }
    parameters.kmerSize = h.kmer_size;
    parameters.error = h.error;
    parameters.minHashesPerWindow = h.min_hashes_per_window;
    parameters.windowSize = h.window_size;
    parameters.concatenated = h.concatenated;
    parameters.noncanonical = h.noncanonical;
    parameters.preserveCase = h.preserve_case;
    parameters.counts = firstHasCounts;
    parameters.seed = h.hash_seed;
    setAlphabetFromString(parameters, h.has_alphabet ? h.alphabet.c_str() : alphabetNucleotide);
    return referenceCount;
}


// The checks initFromFiles applies to a sketch file among its inputs (Sketch.cpp:266-313): false = the file is skipped, with
// the reference's warning.
bool Sketch::sketchFileCompatible(const Parameters& parameters, const string& file, bool contain)
{
    Sketch sketchTest;
    sketchTest.initParametersFromCapnp(file.c_str());
    Sketch current;
    current.parameters = parameters;
    string alphabet, alphabetTest;
    current.getAlphabetAsString(alphabet);
    sketchTest.getAlphabetAsString(alphabetTest);
    if (alphabet != alphabetTest) {
        cerr << "\nWARNING: The sketch file " << file << " has different alphabet (" << alphabetTest << ") than the current alphabet (" << alphabet << "). This file will be skipped." << endl << endl;
        return false;
    }
    if (sketchTest.getHashSeed() != parameters.seed) {
        cerr << "\nWARNING: The sketch " << file << " has a seed size (" << sketchTest.getHashSeed() << ") that does not match the current seed (" << parameters.seed << "). This file will be skipped." << endl << endl;
        return false;
    }
    if (sketchTest.getKmerSize() != parameters.kmerSize) {
        cerr << "\nWARNING: The sketch " << file << " has a kmer size (" << sketchTest.getKmerSize() << ") that does not match the current kmer size (" << parameters.kmerSize << "). This file will be skipped." << endl << endl;
        return false;
    }
    if (!contain && sketchTest.getMinHashesPerWindow() < parameters.minHashesPerWindow) {
        cerr << "\nWARNING: The sketch file " << file << " has a target sketch size (" << sketchTest.getMinHashesPerWindow() << ") that is smaller than the current sketch size (" << parameters.minHashesPerWindow << "). This file will be skipped." << endl << endl;
        return false;
    }
    if (sketchTest.getNoncanonical() != parameters.noncanonical) {
        cerr << "\nWARNING: The sketch file " << file << " is " << (sketchTest.getNoncanonical() ? "noncanonical" : "canonical") << ", which is incompatible with the current setting. This file will be skipped." << endl << endl;
        return false;
    }
    if (sketchTest.getMinHashesPerWindow() > parameters.minHashesPerWindow) {
        cerr << "\nWARNING: The sketch file " << file << " has a target sketch size (" << sketchTest.getMinHashesPerWindow() << ") that is larger than the current sketch size (" << parameters.minHashesPerWindow << "). Its sketches will be reduced." << endl << endl;
    }
    return true;
}

int Sketch::initFromFiles(const vector<string>& files, const Parameters& parametersNew, int verbosity, bool enforceParameters, bool contain)
{
    parameters = parametersNew;
    Batch batch;
    RawBatch rawBatch;

    // Plain FASTA files are parsed on the GPU (FPMASH_GPU_PARSE=0 turns that off): not read sets (-r: FASTQ, and
    // several files may feed one sketch), not .gz, not stdin.  The content checks happen when the file is read.
    const char* gpuParseEnv = getenv("FPMASH_GPU_PARSE");
    const bool gpuParse = !(gpuParseEnv && gpuParseEnv[0] == '0') && !parametersNew.reads && !contain;
    vector<char> rawCandidate(files.size(), 0);
    size_t rawDone = 0;                                          // candidates below this index are already in rawBatch
    for (size_t i = 0; i < files.size(); i++)
        rawCandidate[i] = gpuParse && !hasSuffix(files[i], suffixSketch) && files[i] != "-" && !hasSuffix(files[i], ".gz");
    // .gz files are inflated on the GPU (csrc/gunzip.cu: one warp per file) when a run of them keeps the device busier than
    // the -p host threads would be: one stream inflates at ~28 MB/s there (a host core's zlib: ~240 MB/s), thousands run at
    // once, and the batch takes as long as its largest file.  So: at least 32 files in a row, and
    // largest / 28 MB/s < total / (240 MB/s x p).  FPMASH_GPU_GUNZIP=0 turns the route off, =1 takes it for any run.
    const char* gunzipEnv = getenv("FPMASH_GPU_GUNZIP");
    const bool gunzipOff = gunzipEnv && gunzipEnv[0] == '0', gunzipForce = gunzipEnv && gunzipEnv[0] == '1';
    vector<char> gzCandidate(files.size(), 0);
    for (size_t i = 0; i < files.size() && gpuParse && !gunzipOff;) {
        size_t j = i;
        uint64_t sumBytes = 0, maxBytes = 0;
        while (j < files.size() && hasSuffix(files[j], ".gz") && files[j] != "-") {
            struct stat st;
            const uint64_t sz = stat(files[j].c_str(), &st) == 0 ? (uint64_t)st.st_size : 0;
            sumBytes += sz;
            maxBytes = max(maxBytes, sz);
            j++;
        }
        const double hostThreads = (double)max(parametersNew.parallelism, 1);
        if (j > i && (gunzipForce || (j - i >= 32 && (double)maxBytes * 8.6 * hostThreads < (double)sumBytes)))
            for (size_t t = i; t < j; t++) gzCandidate[t] = rawCandidate[t] = 1;
        i = max(j, i + 1);
    }

    // the other sequence files (not sketches, not stdin) can be parsed ahead by -p threads
    vector<string> jobFiles;
    vector<size_t> jobOf(files.size(), 0);
    for (size_t i = 0; i < files.size(); i++) {
        if (!hasSuffix(files[i], suffixSketch) && files[i] != "-" && !rawCandidate[i]) { jobOf[i] = jobFiles.size(); jobFiles.push_back(files[i]); }
    }
    bool anyStdin = false;
    for (const string& f : files) anyStdin |= f == "-";
    unique_ptr<ParsePool> pool;
    const Parameters poolParameters = parameters;   // sketch files given first may still change `parameters`
    bool sketchFirst = !files.empty() && hasSuffix(files[0], suffixSketch) && !enforceParameters;
    if (parameters.parallelism > 1 && jobFiles.size() > 1 && !anyStdin && !sketchFirst) {
        for (const string& f : jobFiles) {
            FILE* probe = fopen(f.c_str(), "r");
            if (probe == NULL) {
                cerr << "ERROR: could not open " << f << " for reading." << endl;
                exit(1);
            }
            fclose(probe);
        }
        checkUnsupported(parameters);
        pool.reset(new ParsePool(jobFiles, poolParameters, min<int>(parameters.parallelism, (int)jobFiles.size())));
    }

    for (size_t i = 0; i < files.size(); i++) {
        bool isSketch = hasSuffix(files[i], suffixSketch);
        if (isSketch) {
            flushRawBatch(rawBatch, batch);
            flushBatch(batch);                                   // keep submission order (ThreadPool output queue)
            if (i == 0 && !enforceParameters) initParametersFromCapnp(files[i].c_str());
            if (!sketchFileCompatible(parameters, files[i], contain)) continue;
            loadSketchFile(files[i]);
            continue;
        }

        checkUnsupported(parameters);
        if (files[i] == "-") {
            if (verbosity > 0) cerr << "Sketching from stdin..." << endl;
        } else {
            if (verbosity > 0) cerr << "Sketching " << files[i] << "..." << endl;
            FILE* probe = fopen(files[i].c_str(), "r");
            if (probe == NULL) {
                cerr << "ERROR: could not open " << files[i] << " for reading." << endl;
                exit(1);
            }
            fclose(probe);
        }
        if (rawCandidate[i]) {
            flushBatch(batch);                                   // references keep the order of the inputs
            if (rawDone > i) continue;                           // already read with an earlier file of its run
            // the run of GPU-parse candidates starting here, up to one batch of bytes, read by the -p threads
            const bool gz = gzCandidate[i] != 0;
            if (rawBatch.gz != gz) { flushRawBatch(rawBatch, batch); rawBatch.gz = gz; }
            const uint64_t flushAt = gz ? kGzFlushBytes : kFlushBytes;
            vector<string> run;
            uint64_t runBytes = 0;
            for (size_t j = i; j < files.size() && rawCandidate[j] && (gzCandidate[j] != 0) == gz && (run.empty() || rawBatch.used + runBytes < flushAt); j++) {
                struct stat st;
                run.push_back(files[j]);
                runBytes += stat(files[j].c_str(), &st) == 0 ? (uint64_t)st.st_size + 1 : 1;
            }
            const size_t good = rawBatch.addFiles(run, parameters.parallelism);
            rawDone = i + good;
            if (rawBatch.used >= flushAt) flushRawBatch(rawBatch, batch);
            if (good > 0) continue;
            flushRawBatch(rawBatch, batch);                      // gzip behind a plain name, 0x00 inside, ...: host reader
            parseSequenceFile(files[i], parameters, batch);
            rawDone = i + 1;
        } else {
            flushRawBatch(rawBatch, batch);
            if (pool && !rawCandidate[i]) {
                unique_ptr<FileBuffer> fb = pool->take(jobOf[i]);
                batch.append(fb->seq.data(), fb->seq.size(), fb->goff, fb->metas);
            } else {
                parseSequenceFile(files[i], parameters, batch);
            }
        }
        if (batch.used >= kFlushBytes) flushBatch(batch);
    }
    flushRawBatch(rawBatch, batch);
    flushBatch(batch);
    createIndex();
    return 0;
}

int Sketch::writeToCapnp(const char* file) const   // Sketch.cpp:536-642
{
    msh::File mf;
    mf.use64 = parameters.use64;
    mf.header.kmer_size = parameters.kmerSize;
    mf.header.hash_seed = parameters.seed;
    mf.header.error = (float)parameters.error;
    mf.header.min_hashes_per_window = (uint32_t)parameters.minHashesPerWindow;
    mf.header.window_size = (uint32_t)parameters.windowSize;
    mf.header.concatenated = parameters.concatenated;
    mf.header.noncanonical = parameters.noncanonical;
    mf.header.preserve_case = parameters.preserveCase;
    getAlphabetAsString(mf.header.alphabet);
    mf.refs.resize(references.size());
    for (size_t i = 0; i < references.size(); i++) {
        msh::RefRecord& r = mf.refs[i];
        r.name = references[i].name;
        r.comment = references[i].comment;
        r.length = references[i].length;
        r.hashes = references[i].hashesSorted.values;
        r.counts = references[i].counts;
    }
    vector<uint8_t> bytes = msh::encode(mf, parameters.counts);
    if (!msh::write_file(file, bytes)) {
        cerr << "ERROR: could not open " << file << " for writing.\n";
        exit(1);
    }
    return 0;
}
