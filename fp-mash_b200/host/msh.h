// msh.h -- the .msh container (unpacked Cap'n Proto stream of capnp/MinHash.capnp), hand-written
// because libcapnp is not part of this build.  Replaces Sketch::writeToCapnp / loadCapnp /
// initParametersFromCapnp (Sketch.cpp:536-642, 1059-1219, 401-470) at the byte level: the writer
// reproduces MallocMessageBuilder's allocation order and far-pointer placement
// so files are byte-identical with the reference's on the same content.
#pragma once
#include <stdint.h>
#include <string>
#include <vector>

namespace msh {

struct RefRecord {
    std::string name, comment;
    uint64_t length = 0;
    std::vector<uint64_t> hashes;      // u64, or u32 values zero-extended
    std::vector<uint32_t> counts;      // empty = no counts32 list
    bool has_counts = false;           // counts32 pointer present (reader) / to be written (writer)
    bool counts_sorted = false;
};

struct Header {
    uint32_t kmer_size = 0, window_size = 0, min_hashes_per_window = 0, hash_seed = 42;
    bool concatenated = false, noncanonical = false, preserve_case = false;
    float error = 0;
    bool has_alphabet = false;
    std::string alphabet;
};

struct File {
    Header header;
    std::vector<RefRecord> refs;
    bool use64 = true;                 // which hash list the writer emits / the reader prefers
};

// Serialise exactly like writeToCapnp (referenceListOld when seed == 42, empty locus list, alphabet).
// write_counts mirrors `parameters.counts`.
std::vector<uint8_t> encode(const File& f, bool write_counts);

// Parse a .msh image.  use64 selects hashes64 vs hashes32 like loadCapnp; max_hashes truncates each
// list to its first max_hashes entries (0 = keep all).  Returns false and sets err on malformed input.
bool decode(const uint8_t* data, size_t size, bool use64, uint64_t max_hashes, File& out, std::string& err);
bool decode_header(const uint8_t* data, size_t size, Header& h, uint64_t& ref_count, bool& first_has_counts, std::string& err);

// Streaming reader for `mash dist` on large sketch files: the file is mapped, nothing is materialised as
// Reference objects; rows of sketches are copied straight from the mapping into a dense panel (the layout fpm_dist_tile
// takes: [rows][stride] u64 + sizes + lengths), e.g. a pinned staging buffer, chunk by chunk.  Same rules as loadCapnp
// (Sketch.cpp:1059-1219): referenceList else referenceListOld, hashes64 / hashes32 by use64, lists truncated to
// max_hashes, length64 else length.
class PanelReader {
public:
    PanelReader() = default;
    ~PanelReader();
    PanelReader(const PanelReader&) = delete;
    PanelReader& operator=(const PanelReader&) = delete;
    bool open(const std::string& path, std::string& err);
    const Header& header() const { return header_; }
    uint64_t count() const { return n_refs_; }
    // largest list length over all references after truncation to max_hashes (0 = no truncation): the panel stride
    bool max_list(bool use64, uint64_t max_hashes, uint64_t& out, std::string& err) const;
    // rows [i0, i1) -> hashes[(i - i0) * stride + j], sizes[i - i0], lengths[i - i0]; unused slots are left untouched
    bool fill(uint64_t i0, uint64_t i1, bool use64, uint64_t max_hashes, uint64_t* hashes, uint64_t stride, uint32_t* sizes, uint64_t* lengths,
              std::string& err) const;
    bool meta(uint64_t i, std::string& name, std::string& comment, std::string& err) const;

private:
    const uint8_t* map_ = nullptr;
    size_t size_ = 0;
    Header header_;
    uint64_t n_refs_ = 0;
    void* state_ = nullptr;          // parsed segment table + reference list position
};

bool read_file(const std::string& path, std::vector<uint8_t>& bytes);
bool write_file(const std::string& path, const std::vector<uint8_t>& bytes);

}  // namespace msh
