// gunzip_sim.cpp -- test harness (never shipped): runs the sequential half of the GPU inflate, csrc/gunzip_core.cuh, on the
// CPU exactly as lane 0 of gunzip_kernel runs it, with the output side (token placement, CRC join) done serially.
// usage: gunzip_sim <in.gz> <out>   exit status 0 = clean stream, 2 = refused.  tests/test_host_cpu.py compares with zlib.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../fp-mash_b200/csrc/gunzip_core.cuh"

using namespace fpm;

int main(int argc, char** argv)
{
    if (argc != 3) return 64;
    FILE* f = fopen(argv[1], "rb");
    if (!f) return 65;
    std::vector<uint8_t> in;
    // the file starts at an odd offset of the buffer, like a file in the middle of a batch
    const size_t lead = 21;
    in.resize(lead, 0xAA);
    uint8_t buf[65536];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) in.insert(in.end(), buf, buf + got);
    fclose(f);
    const uint64_t in_begin = lead, in_end = in.size();
    in.resize(in.size() + 8192 + 16, 0);
    // 16-byte aligned base, as the device buffer is
    std::vector<uint8_t> store(in.size() + 32);
    uint8_t* base = store.data();
    while (((uintptr_t)base) & 15) base++;
    memcpy(base, in.data(), in.size());

    static GzShared sh;
    for (uint32_t i = 0; i < 256; i++) sh.crc_tab[i] = gz_crc_table_entry(i);
    GzStream s;
    s.init(sh, base, in_begin, in_end);
    std::vector<uint8_t> out;
    uint64_t member_begin = 0;
    int state = GZ_S_HEADER;
    while (state != GZ_S_DONE && state != GZ_S_BAD) {
        // the warp's top-up
        while ((int32_t)(s.b.whi - s.b.wnext()) <= GZ_RING_WORDS - 128) {
            for (int w = 0; w < 128; w++) {
                uint32_t v;
                memcpy(&v, s.b.base + 4ull * (s.b.whi + w), 4);
                sh.ring[(s.b.whi + w) & (GZ_RING_WORDS - 1)] = v;
            }
            s.b.whi += 128;
        }
        uint32_t ntok = 0, ready = 0;
        int next = state;
        if (state == GZ_S_HEADER) next = gz_read_header(s);
        else if (state == GZ_S_BLOCK) next = gz_read_block(s, sh);
        else if (state == GZ_S_CODES) next = gz_decode_batch(s, sh, &ntok, &ready);
        for (uint32_t t = 0; t < ntok && next != GZ_S_BAD; t++) {
            const uint32_t tok = ((ready >> t) & 1u) ? sh.tok[t] : gz_token_at(sh, sh.tok[t]);       // what lane t does
            if (!(tok >> 31)) { out.push_back((uint8_t)tok); continue; }
            const uint32_t len = (tok >> 16) & 0x1ffu, d = (tok & 0xffffu) + 1u;
            if (d > out.size() - member_begin) { next = GZ_S_BAD; break; }
            for (uint32_t j = 0; j < len; j++) out.push_back(out[out.size() - d]);
        }
        if (next == GZ_S_STORED) {
            const uint64_t src0 = s.in_pos();
            if (src0 + s.stored_len > in_end) next = GZ_S_BAD;
            else {
                out.insert(out.end(), base + src0, base + src0 + s.stored_len);
                s.seek(src0 + s.stored_len);
                next = s.last_block ? GZ_S_TRAILER : GZ_S_BLOCK;
            }
        }
        if (next == GZ_S_TRAILER) {
            uint32_t crc_stored = 0, isize = 0;
            if (!gz_read_trailer(s, &crc_stored, &isize) || isize != (uint32_t)(out.size() - member_begin)) next = GZ_S_BAD;
            else {
                // the warp's CRC: 32 chunk CRCs joined by multiplication with x^(8 len)
                const uint64_t n = out.size() - member_begin, per = ((n + 31) / 32 + 3) & ~3ull;
                uint32_t total = 0;
                for (int l = 0; l < 32; l++) {
                    const uint64_t lo = std::min<uint64_t>(member_begin + per * l, out.size()), hi = std::min<uint64_t>(lo + per, out.size());
                    if (hi == lo) break;
                    uint32_t c = 0xffffffffu;
                    for (uint64_t i = lo; i < hi; i++) c = gz_crc_byte(sh.crc_tab, c, out[i]);
                    c = ~c;
                    total = l == 0 ? c : gz_multmodp(gz_x8n_modp(hi - lo), total) ^ c;
                }
                if (total != crc_stored) next = GZ_S_BAD;
                else {
                    const bool more = gz_more_members(s);
                    member_begin = out.size();
                    next = more ? GZ_S_HEADER : GZ_S_DONE;
                }
            }
        }
        state = next;
    }
    if (state == GZ_S_BAD) return 2;
    FILE* o = fopen(argv[2], "wb");
    if (!o) return 66;
    if (!out.empty()) fwrite(out.data(), 1, out.size(), o);
    fclose(o);
    return 0;
}
