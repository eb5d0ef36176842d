// msh.cpp -- see msh.h.  Wire layout and allocator behaviour: SURVEY.md section 5.1 / Appendix D.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include "msh.h"

#include <stdio.h>
#include <string.h>
#include <algorithm>

namespace msh {

// ------------------------------------------------------------------------------------------
// writer: an emulation of MallocMessageBuilder (first segment 1024 words, GROW_HEURISTICALLY)
// + WireHelpers::allocate (object goes into its pointer's segment if it fits, else
// landing pad + object into the newest segment, reached through a far pointer).
// ------------------------------------------------------------------------------------------
namespace {

struct Loc { int seg; size_t idx; };

class Arena {
public:
    struct Seg { std::vector<uint64_t> w; size_t used = 0; };
    std::vector<Seg> segs;
    size_t next_size = 1024;

    uint64_t& at(Loc l) { return segs[l.seg].w[l.idx]; }

    bool try_in(int seg, size_t n, size_t& off)
    {
        Seg& s = segs[seg];
        if (s.w.size() - s.used < n) return false;
        off = s.used;
        s.used += n;
        return true;
    }

    int new_segment(size_t minimum)
    {
        size_t size = std::max(minimum, next_size);
        segs.emplace_back();
        segs.back().w.assign(size, 0);
        if (segs.size() > 1) next_size += size;     // the first segment does not grow the heuristic
        return (int)segs.size() - 1;
    }

    Loc alloc_any(size_t n)
    {
        size_t off;
        if (!segs.empty() && try_in((int)segs.size() - 1, n, off)) return Loc{(int)segs.size() - 1, off};
        int s = new_segment(n);
        try_in(s, n, off);
        return Loc{s, off};
    }

    // Allocate an n-word object for the pointer stored at `ptr`; `kind` = 0 struct / 1 list,
    // `hi` = the pointer's upper 32 bits.  Returns where the object starts.
    Loc alloc_obj(Loc ptr, size_t n, uint32_t kind, uint32_t hi)
    {
        size_t off;
        if (try_in(ptr.seg, n, off)) {
            int64_t rel = (int64_t)off - (int64_t)(ptr.idx + 1);
            at(ptr) = ((uint64_t)hi << 32) | (uint32_t)(((uint32_t)rel << 2) | kind);
            return Loc{ptr.seg, off};
        }
        Loc pad = alloc_any(n + 1);
        at(ptr) = ((uint64_t)(uint32_t)pad.seg << 32) | (uint32_t)(((uint32_t)pad.idx << 3) | 2u);
        at(pad) = ((uint64_t)hi << 32) | kind;       // offset 0: the object follows its landing pad
        return Loc{pad.seg, pad.idx + 1};
    }

    void set_text(Loc ptr, const std::string& s)
    {
        size_t bytes = s.size() + 1;
        Loc o = alloc_obj(ptr, (bytes + 7) / 8, 1, (uint32_t)((bytes << 3) | 2u));
        memcpy((char*)&segs[o.seg].w[o.idx], s.data(), s.size());
    }
};

}  // namespace

std::vector<uint8_t> encode(const File& f, bool write_counts)
{
    Arena A;
    Loc rootptr = A.alloc_any(1);
    Loc root = A.alloc_obj(rootptr, 7, 0, 3u | (4u << 16));                 // MinHash: 3 data words, 4 pointers
    auto root_ptr = [&](int i) { return Loc{root.seg, root.idx + 3 + i}; };

    // referenceListOld (@4, pointer 0) when the seed is the schema default, else referenceList (@11, pointer 3)
    Loc rl = A.alloc_obj(root_ptr(f.header.hash_seed == 42 ? 0 : 3), 1, 0, 0u | (1u << 16));
    const size_t n = f.refs.size();
    Loc list = A.alloc_obj(Loc{rl.seg, rl.idx}, 1 + 9 * n, 1, (uint32_t)(((9 * n) << 3) | 7u));
    A.at(list) = ((uint64_t)(2u | (7u << 16)) << 32) | (uint32_t)(n << 2);   // tag: n elements of (2 data, 7 ptr)
    for (size_t i = 0; i < n; i++) {
        const RefRecord& r = f.refs[i];
        Loc e{list.seg, list.idx + 1 + 9 * i};
        auto eptr = [&](int k) { return Loc{e.seg, e.idx + 2 + k}; };
        A.set_text(eptr(2), r.name);
        A.set_text(eptr(3), r.comment);
        A.at(Loc{e.seg, e.idx + 1}) = r.length;                              // length64
        if (!r.hashes.empty()) {
            if (f.use64) {
                Loc h = A.alloc_obj(eptr(5), r.hashes.size(), 1, (uint32_t)((r.hashes.size() << 3) | 5u));
                memcpy(&A.segs[h.seg].w[h.idx], r.hashes.data(), r.hashes.size() * 8);
            } else {
                Loc h = A.alloc_obj(eptr(4), (r.hashes.size() + 1) / 2, 1, (uint32_t)((r.hashes.size() << 3) | 4u));
                uint32_t* d = (uint32_t*)&A.segs[h.seg].w[h.idx];
                for (size_t j = 0; j < r.hashes.size(); j++) d[j] = (uint32_t)r.hashes[j];
            }
            if (!r.counts.empty() && write_counts) {
                Loc c = A.alloc_obj(eptr(6), (r.counts.size() + 1) / 2, 1, (uint32_t)((r.counts.size() << 3) | 4u));
                memcpy(&A.segs[c.seg].w[c.idx], r.counts.data(), r.counts.size() * 4);
                A.at(e) |= 1ull << 32;                                       // counts32Sorted
            }
        }
    }
    Loc ll = A.alloc_obj(root_ptr(1), 1, 0, 0u | (1u << 16));               // LocusList
    Loc loci = A.alloc_obj(Loc{ll.seg, ll.idx}, 1, 1, (0u << 3) | 7u);       // empty composite list: tag only
    A.at(loci) = (uint64_t)(3u | (0u << 16)) << 32;                          // Locus: 3 data words
    uint64_t w0 = (uint64_t)f.header.kmer_size | ((uint64_t)f.header.window_size << 32);
    uint64_t w1 = (uint64_t)f.header.min_hashes_per_window |
                  ((uint64_t)((f.header.concatenated ? 1u : 0u) | (f.header.noncanonical ? 2u : 0u) | (f.header.preserve_case ? 4u : 0u)) << 32);
    uint32_t errbits;
    memcpy(&errbits, &f.header.error, 4);
    uint64_t w2 = (uint64_t)errbits | ((uint64_t)(f.header.hash_seed ^ 42u) << 32);
    A.at(root) = w0;
    A.at(Loc{root.seg, root.idx + 1}) = w1;
    A.at(Loc{root.seg, root.idx + 2}) = w2;
    A.set_text(root_ptr(2), f.header.alphabet);

    // stream framing: u32 nseg-1, u32 used-size per segment, pad to 8 bytes, segments
    std::vector<uint8_t> out;
    size_t nseg = A.segs.size();
    std::vector<uint32_t> table;
    table.push_back((uint32_t)(nseg - 1));
    for (auto& s : A.segs) table.push_back((uint32_t)s.used);
    if (table.size() & 1) table.push_back(0);
    out.resize(table.size() * 4);
    memcpy(out.data(), table.data(), out.size());
    for (auto& s : A.segs) {
        size_t o = out.size();
        out.resize(o + s.used * 8);
        memcpy(out.data() + o, s.w.data(), s.used * 8);
    }
    return out;
}

// ------------------------------------------------------------------------------------------
// reader
// ------------------------------------------------------------------------------------------
namespace {

struct Msg {
    std::vector<const uint64_t*> seg;
    std::vector<size_t> len;
    std::string err;

    bool in(int s, size_t i, size_t n = 1) const { return s >= 0 && (size_t)s < seg.size() && i <= len[s] && n <= len[s] - i; }
};

struct Obj {          // a resolved pointer
    int kind = -1;    // -1 null, 0 struct, 1 list
    int seg = 0;
    size_t idx = 0;
    uint32_t hi = 0;  // struct: data | ptrs<<16 ; list: count<<3 | elemcode
};

bool resolve(Msg& m, int s, size_t i, Obj& o)
{
    o = Obj();
    if (!m.in(s, i)) { m.err = "pointer outside its segment"; return false; }
    uint64_t w = m.seg[s][i];
    if (w == 0) return true;                       // null
    uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
    int kind = lo & 3;
    if (kind == 2) {                               // far pointer
        size_t pad = lo >> 3;
        int ts = (int)hi;
        bool dbl = (lo >> 2) & 1;
        if (!m.in(ts, pad, dbl ? 2 : 1)) { m.err = "far pointer out of range"; return false; }
        if (!dbl) {
            // the landing pad of a single-far pointer is an ordinary (near) pointer: exactly one hop, no recursion
            if ((m.seg[ts][pad] & 3) == 2) { m.err = "far pointer landing pad is itself a far pointer"; return false; }
            return resolve(m, ts, pad, o);
        }
        uint64_t p0 = m.seg[ts][pad], tag = m.seg[ts][pad + 1];
        if (((uint32_t)p0 & 3) != 2) { m.err = "bad double-far landing pad"; return false; }
        o.kind = (uint32_t)tag & 3; o.seg = (int)(p0 >> 32); o.idx = (uint32_t)p0 >> 3; o.hi = (uint32_t)(tag >> 32);
        if (!m.in(o.seg, o.idx, 0)) { m.err = "double-far target out of range"; return false; }
        return true;
    }
    if (kind == 3) { m.err = "capability pointer in a .msh file"; return false; }
    int32_t off = (int32_t)lo >> 2;
    int64_t target = (int64_t)i + 1 + off;
    if (target < 0 || !m.in(s, (size_t)target, 0)) { m.err = "pointer target out of range"; return false; }
    o.kind = kind; o.seg = s; o.idx = (size_t)target; o.hi = hi;
    return true;
}

struct StructView { int seg; size_t idx; uint32_t nd, np; };

bool as_struct(Msg& m, const Obj& o, StructView& v)
{
    if (o.kind != 0) { v = StructView{0, 0, 0, 0}; return o.kind == -1; }
    v.seg = o.seg; v.idx = o.idx; v.nd = o.hi & 0xffff; v.np = o.hi >> 16;
    if (!m.in(v.seg, v.idx, (size_t)v.nd + v.np)) { m.err = "struct out of range"; return false; }
    return true;
}

uint64_t data_word(const Msg& m, const StructView& v, uint32_t k) { return k < v.nd ? m.seg[v.seg][v.idx + k] : 0; }

bool ptr_field(Msg& m, const StructView& v, uint32_t k, Obj& o)
{
    o = Obj();
    if (k >= v.np) return true;
    return resolve(m, v.seg, v.idx + v.nd + k, o);
}

bool read_text(Msg& m, const Obj& o, std::string& s)
{
    s.clear();
    if (o.kind == -1) return true;
    if (o.kind != 1 || (o.hi & 7) != 2) { m.err = "text field is not a byte list"; return false; }
    size_t n = o.hi >> 3;
    if (!m.in(o.seg, o.idx, (n + 7) / 8)) { m.err = "text out of range"; return false; }
    const char* p = (const char*)(m.seg[o.seg] + o.idx);
    s.assign(p, n ? n - 1 : 0);                     // drop the NUL
    return true;
}

bool open(const uint8_t* data, size_t size, Msg& m)
{
    if (size < 8 || (size & 7)) { m.err = "not a Cap'n Proto stream (size)"; return false; }
    uint32_t nseg;
    memcpy(&nseg, data, 4);
    nseg += 1;
    if (nseg == 0 || nseg > 1u << 20) { m.err = "bad segment count"; return false; }
    size_t hdr = 4 + 4 * (size_t)nseg;
    hdr = (hdr + 7) & ~(size_t)7;
    if (hdr > size) { m.err = "truncated segment table"; return false; }
    size_t pos = hdr;
    for (uint32_t i = 0; i < nseg; i++) {
        uint32_t w;
        memcpy(&w, data + 4 + 4 * i, 4);
        if ((size - pos) / 8 < w) { m.err = "truncated segment"; return false; }
        m.seg.push_back((const uint64_t*)(data + pos));
        m.len.push_back(w);
        pos += (size_t)w * 8;
    }
    return true;
}

struct RootView { StructView root; Obj refs; size_t n_refs = 0; size_t elem_nd = 0, elem_np = 0; };

bool open_root(Msg& m, RootView& rv)
{
    Obj o;
    if (!resolve(m, 0, 0, o) || o.kind != 0) { if (m.err.empty()) m.err = "root is not a struct"; return false; }
    if (!as_struct(m, o, rv.root)) return false;
    // reader.getReferenceList().getReferences().size() ? referenceList : referenceListOld  (Sketch.cpp:444, 1084)
    for (int which : {3, 0}) {
        Obj rl, lst;
        StructView rls;
        if (!ptr_field(m, rv.root, which, rl) || !as_struct(m, rl, rls)) return false;
        if (!ptr_field(m, rls, 0, lst)) return false;
        rv.refs = lst; rv.n_refs = 0;
        if (lst.kind == 1) {
            if ((lst.hi & 7) != 7) { m.err = "reference list is not a composite list"; return false; }
            if (!m.in(lst.seg, lst.idx, 1)) { m.err = "reference list out of range"; return false; }
            uint64_t tag = m.seg[lst.seg][lst.idx];
            rv.n_refs = (uint32_t)tag >> 2;
            rv.elem_nd = (tag >> 32) & 0xffff; rv.elem_np = tag >> 48;
            if (!m.in(lst.seg, lst.idx + 1, rv.n_refs * (rv.elem_nd + rv.elem_np))) { m.err = "reference list out of range"; return false; }
        }
        if (rv.n_refs) break;
    }
    return true;
}

void fill_header(const Msg& m, const StructView& r, Header& h)
{
    uint64_t w0 = data_word(m, r, 0), w1 = data_word(m, r, 1), w2 = data_word(m, r, 2);
    h.kmer_size = (uint32_t)w0; h.window_size = (uint32_t)(w0 >> 32);
    h.min_hashes_per_window = (uint32_t)w1;
    h.concatenated = (w1 >> 32) & 1; h.noncanonical = (w1 >> 33) & 1; h.preserve_case = (w1 >> 34) & 1;
    uint32_t eb = (uint32_t)w2;
    memcpy(&h.error, &eb, 4);
    h.hash_seed = (uint32_t)(w2 >> 32) ^ 42u;
}

}  // namespace

bool decode_header(const uint8_t* data, size_t size, Header& h, uint64_t& ref_count, bool& first_has_counts, std::string& err)
{
    Msg m;
    RootView rv;
    if (!open(data, size, m) || !open_root(m, rv)) { err = m.err; return false; }
    fill_header(m, rv.root, h);
    Obj a;
    if (!ptr_field(m, rv.root, 2, a)) { err = m.err; return false; }
    h.has_alphabet = a.kind != -1;
    if (!read_text(m, a, h.alphabet)) { err = m.err; return false; }
    ref_count = rv.n_refs;
    first_has_counts = false;
    if (rv.n_refs) {
        StructView e{rv.refs.seg, rv.refs.idx + 1, (uint32_t)rv.elem_nd, (uint32_t)rv.elem_np};
        Obj c;
        if (!ptr_field(m, e, 6, c)) { err = m.err; return false; }
        first_has_counts = c.kind != -1;
    }
    return true;
}

bool decode(const uint8_t* data, size_t size, bool use64, uint64_t max_hashes, File& out, std::string& err)
{
    Msg m;
    RootView rv;
    if (!open(data, size, m) || !open_root(m, rv)) { err = m.err; return false; }
    fill_header(m, rv.root, out.header);
    Obj a;
    if (!ptr_field(m, rv.root, 2, a) || !read_text(m, a, out.header.alphabet)) { err = m.err; return false; }
    out.header.has_alphabet = a.kind != -1;
    out.use64 = use64;
    out.refs.assign(rv.n_refs, RefRecord());
    const size_t esz = rv.elem_nd + rv.elem_np;
    for (size_t i = 0; i < rv.n_refs; i++) {
        StructView e{rv.refs.seg, rv.refs.idx + 1 + i * esz, (uint32_t)rv.elem_nd, (uint32_t)rv.elem_np};
        RefRecord& r = out.refs[i];
        Obj o;
        if (!ptr_field(m, e, 2, o) || !read_text(m, o, r.name)) { err = m.err; return false; }
        if (!ptr_field(m, e, 3, o) || !read_text(m, o, r.comment)) { err = m.err; return false; }
        uint64_t d0 = data_word(m, e, 0), len64 = data_word(m, e, 1);
        r.length = len64 ? len64 : (uint32_t)d0;                              // Sketch.cpp:1099-1106
        r.counts_sorted = (d0 >> 32) & 1;
        if (!ptr_field(m, e, use64 ? 5 : 4, o)) { err = m.err; return false; }
        size_t hn = 0;
        if (o.kind == 1) {
            uint32_t code = o.hi & 7;
            hn = o.hi >> 3;
            if (code != (use64 ? 5u : 4u)) { err = "hash list has the wrong element size"; return false; }
            if (!m.in(o.seg, o.idx, use64 ? hn : (hn + 1) / 2)) { err = "hash list out of range"; return false; }
            if (max_hashes && hn > max_hashes) hn = max_hashes;               // Sketch.cpp:1117-1120
            r.hashes.resize(hn);
            if (use64) memcpy(r.hashes.data(), m.seg[o.seg] + o.idx, hn * 8);
            else {
                const uint32_t* p = (const uint32_t*)(m.seg[o.seg] + o.idx);
                for (size_t j = 0; j < hn; j++) r.hashes[j] = p[j];
            }
        }
        if (!ptr_field(m, e, 6, o)) { err = m.err; return false; }
        r.has_counts = o.kind != -1;
        if (o.kind == 1) {
            size_t cn = o.hi >> 3;
            if ((o.hi & 7) != 4 || !m.in(o.seg, o.idx, (cn + 1) / 2)) { err = "counts list malformed"; return false; }
            const uint32_t* p = (const uint32_t*)(m.seg[o.seg] + o.idx);
            r.counts.resize(hn);                                              // Sketch.cpp:1161-1168: hashCount entries
            for (size_t j = 0; j < hn; j++) r.counts[j] = j < cn ? p[j] : 0;
        }
    }
    return true;
}

// ---- PanelReader -------------------------------------------------------------------------------------------------
namespace {
struct ReaderState { Msg m; RootView rv; };
}  // namespace

PanelReader::~PanelReader()
{
    delete (ReaderState*)state_;
    if (map_) munmap((void*)map_, size_);
}

bool PanelReader::open(const std::string& path, std::string& err)
{
    int fd = ::open(path.c_str(), O_RDONLY);
    if (fd < 0) { err = "could not open the file"; return false; }
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) { ::close(fd); err = "empty or unreadable file"; return false; }
    size_ = (size_t)st.st_size;
    void* p = mmap(nullptr, size_, PROT_READ, MAP_PRIVATE, fd, 0);
    ::close(fd);
    if (p == MAP_FAILED) { map_ = nullptr; err = "mmap failed"; return false; }
    map_ = (const uint8_t*)p;
    madvise(p, size_, MADV_SEQUENTIAL);
    ReaderState* stt = new ReaderState();
    state_ = stt;
    if (!msh::open(map_, size_, stt->m) || !open_root(stt->m, stt->rv)) { err = stt->m.err; return false; }
    fill_header(stt->m, stt->rv.root, header_);
    Obj a;
    if (!ptr_field(stt->m, stt->rv.root, 2, a) || !read_text(stt->m, a, header_.alphabet)) { err = stt->m.err; return false; }
    header_.has_alphabet = a.kind != -1;
    n_refs_ = stt->rv.n_refs;
    return true;
}

static bool hash_list_of(Msg& m, const RootView& rv, uint64_t i, bool use64, Obj& o, size_t& n, std::string& err)
{
    const size_t esz = rv.elem_nd + rv.elem_np;
    StructView e{rv.refs.seg, rv.refs.idx + 1 + i * esz, (uint32_t)rv.elem_nd, (uint32_t)rv.elem_np};
    if (!ptr_field(m, e, use64 ? 5 : 4, o)) { err = m.err; return false; }
    n = 0;
    if (o.kind == 1) {
        n = o.hi >> 3;
        if ((o.hi & 7) != (use64 ? 5u : 4u)) { err = "hash list has the wrong element size"; return false; }
        if (!m.in(o.seg, o.idx, use64 ? n : (n + 1) / 2)) { err = "hash list out of range"; return false; }
    }
    return true;
}

bool PanelReader::max_list(bool use64, uint64_t max_hashes, uint64_t& out, std::string& err) const
{
    ReaderState* stt = (ReaderState*)state_;
    out = 0;
    for (uint64_t i = 0; i < n_refs_; i++) {
        Obj o;
        size_t n;
        if (!hash_list_of(stt->m, stt->rv, i, use64, o, n, err)) return false;
        if (max_hashes && n > max_hashes) n = max_hashes;
        if (n > out) out = n;
    }
    return true;
}

bool PanelReader::fill(uint64_t i0, uint64_t i1, bool use64, uint64_t max_hashes, uint64_t* hashes, uint64_t stride, uint32_t* sizes, uint64_t* lengths,
                       std::string& err) const
{
    ReaderState* stt = (ReaderState*)state_;
    Msg& m = stt->m;
    const RootView& rv = stt->rv;
    const size_t esz = rv.elem_nd + rv.elem_np;
    if (i1 > n_refs_ || i0 > i1) { err = "row range outside the file"; return false; }
    for (uint64_t i = i0; i < i1; i++) {
        StructView e{rv.refs.seg, rv.refs.idx + 1 + i * esz, (uint32_t)rv.elem_nd, (uint32_t)rv.elem_np};
        const uint64_t d0 = data_word(m, e, 0), len64 = data_word(m, e, 1);
        lengths[i - i0] = len64 ? len64 : (uint32_t)d0;                      // Sketch.cpp:1099-1106
        Obj o;
        size_t n;
        if (!hash_list_of(m, rv, i, use64, o, n, err)) return false;
        if (max_hashes && n > max_hashes) n = max_hashes;                     // Sketch.cpp:1117-1120
        if (n > stride) { err = "a sketch is longer than the panel stride"; return false; }
        uint64_t* dst = hashes + (i - i0) * stride;
        if (use64) memcpy(dst, m.seg[o.seg] + o.idx, n * 8);
        else {
            const uint32_t* p = n ? (const uint32_t*)(m.seg[o.seg] + o.idx) : nullptr;
            for (size_t j = 0; j < n; j++) dst[j] = p[j];
        }
        sizes[i - i0] = (uint32_t)n;
    }
    return true;
}

bool PanelReader::meta(uint64_t i, std::string& name, std::string& comment, std::string& err) const
{
    ReaderState* stt = (ReaderState*)state_;
    Msg& m = stt->m;
    const RootView& rv = stt->rv;
    if (i >= n_refs_) { err = "row outside the file"; return false; }
    const size_t esz = rv.elem_nd + rv.elem_np;
    StructView e{rv.refs.seg, rv.refs.idx + 1 + i * esz, (uint32_t)rv.elem_nd, (uint32_t)rv.elem_np};
    Obj o;
    if (!ptr_field(m, e, 2, o) || !read_text(m, o, name)) { err = m.err; return false; }
    if (!ptr_field(m, e, 3, o) || !read_text(m, o, comment)) { err = m.err; return false; }
    return true;
}

bool read_file(const std::string& path, std::vector<uint8_t>& bytes)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    bytes.resize(n > 0 ? (size_t)n : 0);
    size_t got = n > 0 ? fread(bytes.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    return got == bytes.size();
}

bool write_file(const std::string& path, const std::vector<uint8_t>& bytes)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    size_t put = bytes.empty() ? 0 : fwrite(bytes.data(), 1, bytes.size(), f);
    fclose(f);
    return put == bytes.size();
}

}  // namespace msh
