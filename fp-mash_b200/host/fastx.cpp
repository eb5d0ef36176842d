// fastx.cpp -- see fastx.h.
#include "fastx.h"

#include <ctype.h>
#include <string.h>

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
        size_t old = seq.size();
        seq.resize(old + (size_t)(e - p));
        char* out = seq.data() + old;
        bool stop = false;
        while (p < e) {
            unsigned char ch = *p++;
            unsigned char cls = kClass.t[ch];
            if (cls == 0) *out++ = (char)ch;
            else if (cls == 2) { c = ch; stop = true; break; }
        }
        seq.resize((size_t)(out - seq.data()));
        begin = (size_t)(p - buf.data());
        if (stop) break;
    }
    if (c == '>' || c == '@') last_char = c;
    if (c != '+') return (int64_t)seq.size();
    while ((c = getc()) != -1 && c != '\n') {}
    if (c == -1) return -2;
    size_t qual = 0;
    while ((c = getc()) != -1 && qual < seq.size())
        if (c >= 33 && c <= 127) qual++;
    last_char = 0;
    if (qual != seq.size()) return -2;
    return (int64_t)seq.size();
}
