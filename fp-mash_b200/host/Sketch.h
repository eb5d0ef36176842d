// Sketch.h -- host-side mirror of the reference's Sketch container (mash/src/mash/Sketch.h:30-288):
// same Parameters / Reference fields, same init* entry points and getters, same error behaviour
// (message on cerr + exit(1)).  The arithmetic behind it is NOT here: sequences are batched and
// handed to the CUDA library through the C ABI (include/fpmash_b200.h); there is no CPU path.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

static const char* const suffixSketch = ".msh";
static const char* const alphabetNucleotide = "ACGT";
static const char* const alphabetProtein = "ACDEFGHIKLMNPQRSTVWY";

// HashList (HashList.h:7-40): hashes kept as u64 (32-bit hashes zero-extended) + the width flag.
class HashList {
public:
    HashList() : use64(true) {}
    uint64_t at(size_t i) const { return values.at(i); }
    size_t size() const { return values.size(); }
    void clear() { values.clear(); }
    void add(uint64_t h) { values.push_back(h); }
    void setUse64(bool u) { use64 = u; }
    bool get64() const { return use64; }
    std::vector<uint64_t> values;
private:
    bool use64;
};

class Sketch {
public:
    struct Parameters {   // Sketch.h:40-113
        int parallelism = 1;
        int kmerSize = 0;
        bool alphabet[256];
        uint32_t alphabetSize = 0;
        bool preserveCase = false;
        bool use64 = false;
        uint32_t seed = 0;
        double error = 0;
        double warning = 0;
        uint64_t minHashesPerWindow = 0;
        uint64_t windowSize = 0;
        bool windowed = false;
        bool concatenated = false;
        bool noncanonical = false;
        bool reads = false;
        uint64_t memoryBound = 0;
        uint32_t minCov = 1;
        double targetCov = 0;
        uint64_t genomeSize = 0;
        bool counts = false;
        bool fingerprint = false;
        Parameters() { memset(alphabet, 0, sizeof alphabet); }
    };

    struct Reference {    // Sketch.h:177-186
        std::string id, name, comment;
        uint64_t length = 0;
        HashList hashesSorted;
        std::vector<uint32_t> counts;
        bool countsSorted = false;
    };

    void initFromFingerprints(const std::vector<std::string>& files, const Parameters& parametersNew);
    static bool sketchFileCompatible(const Parameters& parameters, const std::string& file, bool contain);
    int initFromFiles(const std::vector<std::string>& files, const Parameters& parametersNew, int verbosity = 0,
                      bool enforceParameters = false, bool contain = false);
    void initFromReads(const std::vector<std::string>& files, const Parameters& parametersNew);
    uint64_t initParametersFromCapnp(const char* file);
    int writeToCapnp(const char* file) const;

    void getAlphabetAsString(std::string& alphabet) const;
    uint32_t getAlphabetSize() const { return parameters.alphabetSize; }
    bool getConcatenated() const { return parameters.concatenated; }
    float getError() const { return parameters.error; }
    uint32_t getHashSeed() const { return parameters.seed; }
    float getMinHashesPerWindow() const { return parameters.minHashesPerWindow; }   // float, like Sketch.h:234
    int getMinKmerSize(uint64_t reference) const;
    bool getPreserveCase() const { return parameters.preserveCase; }
    double getRandomKmerChance(uint64_t reference) const;
    const Reference& getReference(uint64_t index) const { return references.at(index); }
    uint64_t getReferenceCount() const { return references.size(); }
    void getReferenceHistogram(uint64_t index, std::map<uint32_t, uint64_t>& histogram) const;
    uint64_t getReferenceIndex(std::string id) const;
    int getKmerSize() const { return parameters.kmerSize; }
    double getKmerSpace() const { return kmerSpace; }
    bool getUse64() const { return parameters.use64; }
    uint64_t getWindowSize() const { return parameters.windowSize; }
    bool getNoncanonical() const { return parameters.noncanonical; }
    bool hasHashCounts() const { return references.size() > 0 && references.at(0).counts.size() > 0; }
    void setReferenceName(int i, const std::string name) { references[i].name = name; }
    void setReferenceComment(int i, const std::string comment) { references[i].comment = comment; }
    const Parameters& getParameters() const { return parameters; }
    struct Batch;   // sequences on their way to the GPU (Sketch.cpp)

private:
    void createIndex();
    void flushBatch(Batch& b);
    struct RawBatch;   // raw FASTA files on their way to the GPU parser (Sketch.cpp)
    void flushRawBatch(RawBatch& rb, Batch& hostBatch);
    void loadSketchFile(const std::string& file);

    std::vector<Reference> references;
    std::unordered_map<std::string, int> referenceIndecesById;
    Parameters parameters;
    double kmerSpace = 0;
};

bool hasSuffix(std::string const& whole, std::string const& suffix);
void setAlphabetFromString(Sketch::Parameters& parameters, const char* characters);

// The process-wide CUDA context (created on first use; failure is fatal: no CPU fallback).
struct fpm_ctx;
struct fpm_multi;
fpm_ctx* gpuContext();
// all GPUs of the box for jobs of at least `threshold` units of `work` (pairs / sequence bytes); nullptr: use gpuContext()
fpm_multi* gpuMulti(uint64_t work, uint64_t threshold);
static const uint64_t kMultiGpuPairs = 64ull << 20;   // dist jobs below this many pairs stay on one GPU
// FPMASH_TIMING=1: print elapsed wall-clock milestones to stderr (where does a CLI run spend its time?)
void fpmTick(const char* label);
