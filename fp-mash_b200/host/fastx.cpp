// fastx.cpp -- see fastx.h.
#include "fastx.h"

#include <ctype.h>
#include <string.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace {
// byte classes while reading sequence lines: 0 keep (isgraph), 1 drop, 2 terminator ('>', '+', '@')
struct SeqClass {
    unsigned char t[256];
    SeqClass()
    {
        for (int c = 0; c < 256; c++) t[c] = (c >= 33 && c <= 126) ? 0 : 1;
        t[(unsigned char)'>'] = t[(unsigned char)'+'] = t[(unsigned char)'@'] = 2;
    }
};
const SeqClass kClass;
}  // namespace

FastxReader::FastxReader(gzFile f) : fp(f), buf(1 << 20) {}
FastxReader::~FastxReader() {}

int FastxReader::getc()
{
    if (begin >= end) {
        if (eof) return -1;
        begin = 0;
        int n = gzread(fp, buf.data(), (unsigned)buf.size());
        if (n < (int)buf.size()) eof = true;
        if (n <= 0) { end = 0; return -1; }
        end = (size_t)n;
    }
    return buf[begin++];
}

int64_t FastxReader::next()
{
    int c;
    if (last_char == 0) {
        while ((c = getc()) != -1 && c != '>' && c != '@') {}
        if (c == -1) return -1;
        last_char = c;
    }
    seq.clear();
    comment.clear();
    // name: up to the first whitespace
    if (begin >= end && eof) return -1;
    name.clear();
    int delim = 0;
    while ((c = getc()) != -1) {
        if (isspace(c)) { delim = c; break; }
        name.push_back((char)c);
    }
    if (delim != '\n') {
        // rest of the line is the comment; at EOF kseq leaves the buffer untouched
        if (!(begin >= end && eof)) {
            while ((c = getc()) != -1 && c != '\n') comment.push_back((char)c);
            comment_cstr = comment;
        }
    }
    // sequence: every isgraph byte up to the next '>', '+' or '@'
    c = -1;
    for (;;) {
        if (begin >= end) {
            int first = getc();
            if (first == -1) { c = -1; break; }
            begin--;                       // un-read: let the bulk loop classify it
        }
        const unsigned char* p = buf.data() + begin;
        const unsigned char* e = buf.data() + end;
        seq.reserve_more((size_t)(e - p));
        char* out = seq.data() + seq.size();
        bool stop = false;
        while (p < e) {
#if defined(__SSE2__)
            // 16 bytes at a time while they are all plain sequence characters (33..126, none of > + @)
            while (p + 16 <= e) {
                __m128i x = _mm_loadu_si128((const __m128i*)p);
                __m128i bad = _mm_or_si128(_mm_cmpgt_epi8(_mm_set1_epi8(33), x), _mm_cmpgt_epi8(x, _mm_set1_epi8(126)));
                bad = _mm_or_si128(bad, _mm_cmpeq_epi8(x, _mm_set1_epi8('>')));
                bad = _mm_or_si128(bad, _mm_cmpeq_epi8(x, _mm_set1_epi8('+')));
                bad = _mm_or_si128(bad, _mm_cmpeq_epi8(x, _mm_set1_epi8('@')));
                int mask = _mm_movemask_epi8(bad);
                _mm_storeu_si128((__m128i*)out, x);        // harmless over-write: the buffer has room for e - p bytes
                if (mask) {
                    int good = __builtin_ctz(mask);          // bytes before the first special one
                    p += good; out += good;
                    break;
                }
                p += 16; out += 16;
            }
            if (p >= e) break;
#endif
            unsigned char ch = *p++;
            unsigned char cls = kClass.t[ch];
            if (cls == 0) *out++ = (char)ch;
            else if (cls == 2) { c = ch; stop = true; break; }
        }
        seq.n = (size_t)(out - seq.data());
        begin = (size_t)(p - buf.data());
        if (stop) break;
    }
    if (c == '>' || c == '@') last_char = c;
    if (c != '+') return (int64_t)seq.size();
    while ((c = getc()) != -1 && c != '\n') {}
    if (c == -1) return -2;
    // quality: skip bytes until as many printable ones (33..127) as bases were seen, then one more byte
    size_t qual = 0;
    for (;;) {
        if (begin >= end) {
            c = getc();
            if (c == -1) break;
            begin--;
        }
        if (qual >= seq.size()) { c = buf[begin++]; break; }     // the byte consumed after the last quality byte
        const unsigned char* p = buf.data() + begin;
        const unsigned char* e = buf.data() + end;
        const size_t need = seq.size() - qual;
        // fast path: the next `need` bytes are one clean quality line
        if ((size_t)(e - p) > need && memchr(p, '\n', need) == nullptr && memchr(p, '\r', need) == nullptr &&
            memchr(p, ' ', need) == nullptr && memchr(p, '\t', need) == nullptr) {
            bool clean = true;
            for (size_t i = 0; i < need; i++) clean &= (p[i] >= 33) & (p[i] <= 127);
            if (clean) { qual += need; begin += need; continue; }
        }
        unsigned char ch = buf[begin++];
        if (ch >= 33 && ch <= 127) qual++;
    }
    last_char = 0;
    if (qual != seq.size()) return -2;
    return (int64_t)seq.size();
}
