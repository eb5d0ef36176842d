// fastx.h -- FASTA/FASTQ (optionally gzipped) record reader with the semantics of the
// reference's kseq.h (2009 version, kseq.h:170-208), which the sketching results depend on:
//   * a record starts at the next '>' or '@';
//   * name = up to the first whitespace; if that delimiter was not '\n', comment = rest of the
//     line (a '\r' stays in it);
//   * sequence = every isgraph() byte up to the next '>', '+' or '@' (ANYWHERE, not only at a
//     line start); other bytes (space, '\r', '\n') are dropped;
//   * after '+': skip that line, then read quality bytes (33..127) until as many as bases;
//     one more byte is consumed after the last quality byte; shorter quality => error (-2);
//   * the comment buffer keeps its previous content when a record has no comment (only its
//     length is reset) -- sketchFile reads it as a C string (Sketch.cpp:1383-1392), so that
//     stale text is observable and is mirrored here as `comment_cstr`.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <zlib.h>
#include <string>
#include <vector>

class FastxReader {
public:
    explicit FastxReader(gzFile f);
    ~FastxReader();
    // >= 0 sequence length, -1 end of file, -2 truncated quality
    int64_t next();
    std::string name;
    std::string comment;        // this record's comment (empty if none)
    std::string comment_cstr;   // what kseq's comment.s would hold (stale when no comment)
    // sequence bytes (not NUL-terminated); a plain growable buffer (no zero fill on growth)
    struct CharBuf {
        char* p = nullptr;
        size_t n = 0, cap = 0;
        ~CharBuf() { free(p); }
        char* data() { return p; }
        const char* data() const { return p; }
        size_t size() const { return n; }
        void clear() { n = 0; }
        void reserve_more(size_t extra)
        {
            if (n + extra <= cap) return;
            size_t nc = cap ? cap * 2 : (1 << 16);
            while (nc < n + extra) nc *= 2;
            p = (char*)realloc(p, nc);
            cap = nc;
        }
        const char* begin() const { return p; }
        const char* end() const { return p + n; }
    } seq;
private:
    int getc();
    gzFile fp;
    std::vector<unsigned char> buf;
    size_t begin = 0, end = 0;
    bool eof = false;
    int last_char = 0;
};
