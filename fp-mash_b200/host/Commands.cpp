// Commands.cpp -- `mash sketch`, `mash dist` (+ `paste`, `info`) with the reference's option
// surface and output formats (CommandSketch.cpp:19-122, CommandDistance.cpp:38-333,
// sketchParameterSetup.cpp:9-126, Command.cpp).  All comparison arithmetic runs on the GPU
// through fpm_dist_tile; this file only parses options, moves buffers and prints rows.
#include "Command.h"

#include <math.h>
#include <stdlib.h>
#include <unistd.h>
#include <algorithm>
#include <fstream>
#include <iostream>

#include "Sketch.h"
#include "msh.h"
#include "fpmash_b200.h"

using namespace std;

namespace mash {

// ------------------------------------------------------------------------------------------
// Option / Command base
// ------------------------------------------------------------------------------------------
Command::Option::Option(Type t, string id, string cat, string desc, string def, float mn, float mx)
    : type(t), identifier(id), category(cat), description(desc), argumentDefault(def), argumentMin(mn), argumentMax(mx)
{
    setArgument(argumentDefault);
}

void Command::Option::setArgument(string argumentNew)   // Command.cpp:55-168
{
    argument = argumentNew;
    if (type == Number || type == Integer) {
        if (argument.empty()) { argumentAsNumber = 0; return; }
        bool failed = false;
        try {
            argumentAsNumber = stof(argument);
            if (argumentMin != argumentMax && (argumentAsNumber < argumentMin || argumentAsNumber > argumentMax)) failed = true;
            else if (type == Integer && static_cast<uint64_t>(argumentAsNumber) != argumentAsNumber) failed = true;
        } catch (const exception&) {
            failed = true;
        }
        if (failed) {
            cerr << "ERROR: Argument to -" << identifier << " must be a" << (type == Integer ? "n integer" : " number");
            if (argumentMin != argumentMax) cerr << " between " << argumentMin << " and " << argumentMax;
            cerr << " (" << argument << " given)" << endl;
            exit(1);
        }
    } else if (type == Size) {
        if (argument.empty()) { argumentAsNumber = 0; return; }
        char suffix = argument.back();
        uint64_t factor = 1;
        if (suffix < '0' || suffix > '9') {
            switch (suffix) {
                case 'k': case 'K': factor = 1000; break;
                case 'm': case 'M': factor = 1000000; break;
                case 'g': case 'G': factor = 1000000000; break;
                case 't': case 'T': factor = 1000000000000; break;
                default:
                    cerr << "ERROR: Unrecognized unit (\"" << suffix << "\") in argument to -" << identifier
                         << ". If specified, unit must be one of [kKmMgGtT]." << endl;
                    exit(1);
            }
            argument.pop_back();
        }
        bool fail = false;
        try { argumentAsNumber = stof(argument); } catch (const exception&) { fail = true; }
        if (argumentAsNumber <= 0 || static_cast<uint64_t>(argumentAsNumber) != argumentAsNumber) fail = true;
        if (fail) {
            cerr << "ERROR: Argument to -" << identifier << " must be a whole number, optionally followed by one of [kKmMgGtT]." << endl;
            exit(1);
        }
        argumentAsNumber *= factor;
    }
}

Command::Command()   // Command.cpp:179-238 (the options this path uses)
{
    auto avail = [&](string n, Option o) { optionsAvailable[n] = o; };
    avail("help", Option(Option::Boolean, "h", "", "Help", ""));
    avail("kmer", Option(Option::Integer, "k", "Sketch", "K-mer size. Hashes will be based on strings of this many nucleotides. Canonical nucleotides are used by default (see Alphabet options below).", "21", 1, 32));
    avail("sketchSize", Option(Option::Integer, "s", "Sketch", "Sketch size. Each sketch will have at most this many non-redundant min-hashes.", "1000"));
    avail("individual", Option(Option::Boolean, "i", "Sketch", "Sketch individual sequences, rather than whole files, e.g. for multi-fastas of single-chromosome genomes or pair-wise gene comparisons.", ""));
    avail("warning", Option(Option::Number, "w", "Sketch", "Probability threshold for warning about low k-mer size.", "0.01", 0, 1));
    avail("reads", Option(Option::Boolean, "r", "Sketch", "Input is a read set. See Reads options below. Implies -M. Incompatible with -i.", ""));
    avail("seed", Option(Option::Integer, "S", "Sketch", "Seed to provide to the hash function.", "42", 0, 0xFFFFFFFF));
    avail("memory", Option(Option::Size, "b", "Reads", "Use a Bloom filter of this size (raw bytes or with K/M/G/T) to filter out unique k-mers. Not part of the accelerated path (use -m). Implies -r."));
    avail("minCov", Option(Option::Integer, "m", "Reads", "Minimum copies of each k-mer required to pass noise filter for reads. Implies -r.", "1"));
    avail("targetCov", Option(Option::Number, "c", "Reads", "Target coverage. Not part of the accelerated path. Implies -r."));
    avail("genome", Option(Option::Size, "g", "Reads", "Genome size (raw bases or with K/M/G/T). If specified, will be used for p-value calculation instead of an estimated size from k-mer content. Implies -r."));
    avail("noncanonical", Option(Option::Boolean, "n", "Alphabet", "Preserve strand (by default, strand is ignored by using canonical DNA k-mers, which are alphabetical minima of forward-reverse pairs). Implied if an alphabet is specified with -a or -z.", ""));
    avail("protein", Option(Option::Boolean, "a", "Alphabet", "Use amino acid alphabet (A-Z, except BJOUXZ). Implies -n, -k 9.", ""));
    avail("alphabet", Option(Option::String, "z", "Alphabet", "Alphabet to base hashes on (case ignored by default; see -Z). K-mers with other characters will be ignored. Implies -n.", ""));
    avail("case", Option(Option::Boolean, "Z", "Alphabet", "Preserve case in k-mers and alphabet (case is ignored by default). Sequence letters whose case is not in the current alphabet will be skipped when sketching.", ""));
    avail("threads", Option(Option::Integer, "p", "", "Parallelism. Accepted for compatibility: the GPU path batches all inputs.", "1"));
}

void Command::addOption(string name, Option option)
{
    options[name] = option;
    optionNamesByIdentifier[option.identifier] = name;
    optionOrder.push_back(name);
}

void Command::useSketchOptions()   // Command.cpp:385-410
{
    for (const char* n : {"threads", "kmer", "noncanonical", "protein", "alphabet", "case", "sketchSize", "individual", "seed",
                          "warning", "reads", "memory", "minCov", "targetCov", "genome"})
        useOption(n);
}

void Command::print() const
{
    cout << endl << "Usage:" << endl << endl << "  mash " << name << " [options] " << argumentString << endl << endl;
    cout << "Description:" << endl << endl << "  " << description << endl << endl << "Options:" << endl << endl;
    for (const string& n : optionOrder) {
        const Option& o = options.at(n);
        cout << "  -" << o.identifier;
        if (o.type != Option::Boolean) cout << " <" << (o.type == Option::Integer ? "int" : o.type == Option::Number ? "num" : o.type == Option::Size ? "size" : "text") << ">";
        cout << "  " << o.description;
        if (!o.argumentDefault.empty()) cout << " [" << o.argumentDefault << "]";
        cout << endl;
    }
    cout << endl;
}

int Command::run(int argc, const char** argv)   // Command.cpp:341-376
{
    for (int i = 0; i < argc; i++) {
        if (argv[i][0] == '-' && argv[i][1] != 0) {
            if (optionNamesByIdentifier.count(argv[i] + 1) == 0) {
                cerr << "ERROR: Unrecognized option: " << argv[i] << endl;
                return 1;
            }
            Option& option = options.at(optionNamesByIdentifier.at(argv[i] + 1));
            option.active = true;
            if (option.type != Option::Boolean) {
                i++;
                if (i == argc) {
                    cerr << "ERROR: -" << option.identifier << " requires an argument" << endl;
                    return 1;
                }
                option.setArgument(argv[i]);
            }
        } else {
            arguments.push_back(argv[i]);
        }
    }
    return run();
}

void splitFile(const string& file, vector<string>& lines)   // Command.cpp:430-445
{
    string line;
    ifstream in(file);
    if (in.fail()) {
        cerr << "ERROR: Could not open " << file << ".\n";
        exit(1);
    }
    while (getline(in, line)) lines.push_back(line);
}

// sketchParameterSetup.cpp:9-106
static int sketchParameterSetup(Sketch::Parameters& parameters, const Command& command)
{
    parameters.kmerSize = command.getOption("kmer").getArgumentAsNumber();
    parameters.minHashesPerWindow = command.getOption("sketchSize").getArgumentAsNumber();
    parameters.concatenated = !command.getOption("individual").active;
    parameters.noncanonical = command.getOption("noncanonical").active;
    parameters.seed = command.getOption("seed").getArgumentAsNumber();
    parameters.reads = command.getOption("reads").active;
    parameters.minCov = command.getOption("minCov").getArgumentAsNumber();
    parameters.targetCov = command.getOption("targetCov").getArgumentAsNumber();
    parameters.fingerprint = command.getOption("fingerprint").active;
    parameters.parallelism = command.getOption("threads").getArgumentAsNumber();
    parameters.preserveCase = command.getOption("case").active;
    if (command.hasOption("warning")) parameters.warning = command.getOption("warning").getArgumentAsNumber();
    if (command.getOption("memory").active) {
        parameters.reads = true;
        parameters.memoryBound = command.getOption("memory").getArgumentAsNumber();
        if (command.getOption("minCov").active) {
            cerr << "ERROR: The option " << command.getOption("minCov").identifier << " cannot be used with " << command.getOption("memory").identifier << "." << endl;
            return 1;
        }
    }
    if (command.getOption("minCov").active || command.getOption("targetCov").active) parameters.reads = true;
    if (command.getOption("genome").active) {
        parameters.reads = true;
        parameters.genomeSize = command.getOption("genome").getArgumentAsNumber();
    }
    if (parameters.reads) parameters.counts = true;
    if (parameters.reads && command.getOption("threads").active)
        cerr << "WARNING: The option " << command.getOption("threads").identifier << " will be ignored with " << command.getOption("reads").identifier << "." << endl;
    if (parameters.reads && !parameters.concatenated) {
        cerr << "ERROR: The option " << command.getOption("individual").identifier << " cannot be used with " << command.getOption("reads").identifier << "." << endl;
        return 1;
    }
    if (parameters.fingerprint) {
        parameters.kmerSize = 1;
        parameters.noncanonical = true;
        setAlphabetFromString(parameters, "0123456789");
    } else if (command.getOption("protein").active) {
        parameters.noncanonical = true;
        setAlphabetFromString(parameters, alphabetProtein);
        if (!command.getOption("kmer").active) parameters.kmerSize = 9;
    } else if (command.getOption("alphabet").active) {
        parameters.noncanonical = true;
        setAlphabetFromString(parameters, command.getOption("alphabet").argument.c_str());
    } else {
        setAlphabetFromString(parameters, alphabetNucleotide);
    }
    return 0;
}

// sketchParameterSetup.cpp:108-126
static void warnKmerSize(const Sketch::Parameters& parameters, const Command& command, uint64_t lengthMax, const string& lengthMaxName,
                         double randomChance, int kMin, int warningCount)
{
    cerr << "\nWARNING: For the k-mer size used (" << parameters.kmerSize << "), the random match probability (" << randomChance
         << ") is above the specified warning threshold (" << parameters.warning << ") for the sequence \"" << lengthMaxName
         << "\" of size " << lengthMax;
    if (warningCount > 1) cerr << " (and " << (warningCount - 1) << " others)";
    cerr << ". Distances to " << (warningCount == 1 ? "this sequence" : "these sequences")
         << " may be underestimated as a result. To meet the threshold of " << parameters.warning << ", a k-mer size of at least " << kMin
         << " is required. See: -" << command.getOption("kmer").identifier << ", -" << command.getOption("warning").identifier << "." << endl << endl;
}

// ------------------------------------------------------------------------------------------
// mash sketch (CommandSketch.cpp)
// ------------------------------------------------------------------------------------------
CommandSketch::CommandSketch() : Command()
{
    name = "sketch";
    summary = "Create sketches (reduced representations for fast operations).";
    description = "Create a sketch file, which is a reduced representation of a sequence or set of sequences (based on min-hashes) that can be used for fast distance estimations. Inputs can be fasta or fastq files (gzipped or not), and \"-\" can be given to read from standard input. Input files can also be files of file names (see -l). For output, one sketch file will be generated, but it can have multiple sketches within it, divided by sequences or files (see -i). By default, the output file name will be the first input file with a '.msh' extension, or 'stdin.msh' if standard input is used (see -o).";
    argumentString = "<input> [<input>] ...";
    useOption("help");
    addOption("list", Option(Option::Boolean, "l", "Input", "List input. Lines in each <input> specify paths to sequence files, one per line.", ""));
    addOption("prefix", Option(Option::File, "o", "Output", "Output prefix (first input file used if unspecified). The suffix '.msh' will be appended.", ""));
    addOption("id", Option(Option::File, "I", "Sketch", "ID field for sketch of reads (instead of first sequence ID).", ""));
    addOption("comment", Option(Option::File, "C", "Sketch", "Comment for a sketch of reads (instead of first sequence comment).", ""));
    addOption("counts", Option(Option::Boolean, "M", "Sketch", "Store multiplicity of each k-mer in each sketch.", ""));
    addOption("fingerprint", Option(Option::Boolean, "fp", "Input", "Indicates that the input files are fingerprints instead of sequences.", ""));
    useSketchOptions();
}

int CommandSketch::run() const
{
    if (arguments.size() == 0 || options.at("help").active) {
        print();
        return 0;
    }
    int verbosity = 1;
    bool list = options.at("list").active;
    bool fingerprint = options.at("fingerprint").active;
    Sketch::Parameters parameters;
    parameters.counts = options.at("counts").active;
    if (sketchParameterSetup(parameters, *this)) return 1;
    vector<string> files;
    for (size_t i = 0; i < arguments.size(); i++) {
        if (list) splitFile(arguments[i], files);
        else files.push_back(arguments[i]);
    }
    Sketch sketch;
    if (parameters.reads) sketch.initFromReads(files, parameters);
    else if (fingerprint) sketch.initFromFingerprints(files, parameters);
    else sketch.initFromFiles(files, parameters, verbosity);
    if (getOption("id").active) sketch.setReferenceName(0, getOption("id").argument);
    if (getOption("comment").active) sketch.setReferenceComment(0, getOption("comment").argument);
    string prefix;
    if (options.at("prefix").argument.length() > 0) prefix = options.at("prefix").argument;
    else prefix = arguments[0] == "-" ? "stdin" : arguments[0];
    if (!hasSuffix(prefix, suffixSketch)) prefix += suffixSketch;
    cerr << "Writing to " << prefix << "..." << endl;
    sketch.writeToCapnp(prefix.c_str());
    return 0;
}

// ------------------------------------------------------------------------------------------
// mash dist (CommandDistance.cpp)
// ------------------------------------------------------------------------------------------
CommandDistance::CommandDistance() : Command()
{
    name = "dist";
    summary = "Estimate the distance of query sequences to references.";
    description = "Estimate the distance of each query sequence to the reference. Both the reference and queries can be fasta or fastq, gzipped or not, or Mash sketch files (.msh) with matching k-mer sizes. Query files can also be files of file names (see -l). Whole files are compared by default (see -i). The output fields are [reference-ID, query-ID, distance, p-value, shared-hashes].";
    argumentString = "<reference> <query> [<query>] ...";
    useOption("help");
    addOption("list", Option(Option::Boolean, "l", "Input", "List input. Lines in each <query> specify paths to sequence files, one per line. The reference file is not affected.", ""));
    addOption("table", Option(Option::Boolean, "t", "Output", "Table output (will not report p-values, but fields will be blank if they do not meet the p-value threshold).", ""));
    addOption("pvalue", Option(Option::Number, "v", "Output", "Maximum p-value to report.", "1.0", 0., 1.));
    addOption("distance", Option(Option::Number, "d", "Output", "Maximum distance to report.", "1.0", 0., 1.));
    addOption("comment", Option(Option::Boolean, "C", "Output", "Show comment fields with reference/query names (denoted with ':').", "1.0", 0., 1.));
    addOption("fingerprint", Option(Option::Boolean, "fp", "Input", "Indicates that the input files are fingerprints instead of sequences.", ""));
    useSketchOptions();
}

static bool containsTag(const vector<string>& v, const char* tag)   // containsMSH / containsTXT, CommandDistance.cpp:453-475
{
    bool flag = false;
    for (const auto& s : v) flag = s.find(tag) != string::npos;
    return flag;
}

struct HostPanel {
    vector<uint64_t> hashes;
    vector<uint32_t> sizes;
    vector<uint64_t> lengths;
    uint64_t stride = 0;
    fpm_panel view(uint64_t first, uint64_t n) const
    {
        fpm_panel p;
        p.hashes = hashes.data() + first * stride;
        p.sizes = sizes.data() + first;
        p.lengths = lengths.data() + first;
        p.n = n;
        p.stride = stride;
        return p;
    }
};

static void buildPanel(const Sketch& sk, HostPanel& p)
{
    uint64_t n = sk.getReferenceCount();
    p.stride = 1;
    for (uint64_t i = 0; i < n; i++) p.stride = max<uint64_t>(p.stride, sk.getReference(i).hashesSorted.size());
    p.hashes.assign(n * p.stride, 0);
    p.sizes.resize(n);
    p.lengths.resize(n);
    for (uint64_t i = 0; i < n; i++) {
        const Sketch::Reference& r = sk.getReference(i);
        copy(r.hashesSorted.values.begin(), r.hashesSorted.values.end(), p.hashes.begin() + i * p.stride);
        p.sizes[i] = (uint32_t)r.hashesSorted.size();
        p.lengths[i] = r.length;
    }
}

// `mash dist ref.msh query1.msh query2.msh ...` without materialising the queries: chunks of rows from the mapped files ->
// pinned panel -> fpm_dist_tile / fpm_dist_hits against the resident reference panel -> writeOutput (CommandDistance.cpp:276-333).
static int streamQueries(const Sketch& sketchRef, const vector<string>& queryFiles, const Sketch::Parameters& parameters, bool table, bool comment,
                         double distanceMax, double pValueMax)
{
    const uint64_t nRef = sketchRef.getReferenceCount();
    HostPanel pr;
    buildPanel(sketchRef, pr);
    fpm_dist_params dp;
    // compare(): sketchSize = min of the two getMinHashesPerWindow(); the query side's is `parameters` (enforced on its files)
    dp.sketch_size = (uint32_t)min<uint64_t>(parameters.minHashesPerWindow, sketchRef.getMinHashesPerWindow());
    dp.kmer_size = sketchRef.getKmerSize();
    dp.kmer_space = sketchRef.getKmerSpace();
    dp.max_distance = distanceMax;
    dp.max_pvalue = pValueMax;
    dp.sorted_unique = 1;
    {
        fpm_panel vr = pr.view(0, nRef);
        if (fpm_dist_set_reference(gpuContext(), &vr) != FPM_OK) {
            cerr << "ERROR: " << fpm_last_error() << endl;
            return 1;
        }
    }
    const bool filtered = !table && ((distanceMax >= 0 && distanceMax < 1.) || (pValueMax >= 0 && pValueMax < 1.));
    void* pinned = nullptr;
    size_t pinnedCap = 0;
    vector<fpm_pair> out;
    vector<fpm_hit> hits;
    vector<string> names, comments;
    for (const string& file : queryFiles) {
        if (!Sketch::sketchFileCompatible(parameters, file, false)) continue;
        msh::PanelReader reader;
        string err;
        uint64_t stride = 0;
        if (!reader.open(file, err) || !reader.max_list(parameters.use64, parameters.minHashesPerWindow, stride, err)) {
            cerr << "ERROR: " << file << " is not a valid sketch file (" << err << ")." << endl;
            return 1;
        }
        stride = max<uint64_t>(stride, 1);
        const uint64_t nQry = reader.count();
        // rows per chunk: ~64 MB of hashes in the staging panel, and at most ~32M pairs of results
        uint64_t rows = max<uint64_t>(1, min<uint64_t>((64ull << 20) / (stride * 8), (32ull << 20) / nRef));
        const size_t need = rows * (stride * 8 + 12) + 64;
        if (need > pinnedCap) {
            if (pinned) fpm_host_free(pinned);
            if (fpm_host_alloc(need, &pinned) != FPM_OK) { cerr << "ERROR: " << fpm_last_error() << endl; return 1; }
            pinnedCap = need;
        }
        uint64_t* ph = (uint64_t*)pinned;
        uint64_t* pl = ph + rows * stride;
        uint32_t* ps = (uint32_t*)(pl + rows);
        for (uint64_t q0 = 0; q0 < nQry; q0 += rows) {
            const uint64_t nq = min(rows, nQry - q0);
            if (!reader.fill(q0, q0 + nq, parameters.use64, parameters.minHashesPerWindow, ph, stride, ps, pl, err)) {
                cerr << "ERROR: " << file << " is not a valid sketch file (" << err << ")." << endl;
                return 1;
            }
            names.resize(nq); comments.resize(nq);
            for (uint64_t i = 0; i < nq; i++)
                if (!reader.meta(q0 + i, names[i], comments[i], err)) { cerr << "ERROR: " << file << " is not a valid sketch file (" << err << ")." << endl; return 1; }
            fpm_panel vq;
            vq.hashes = ph; vq.sizes = ps; vq.lengths = pl; vq.n = nq; vq.stride = stride;
            if (filtered) {
                if (hits.size() < (1u << 16)) hits.resize(max<uint64_t>(1 << 16, 16 * (nRef + nq)));
                uint64_t nHits = 0;
                int rc = fpm_dist_hits(gpuContext(), &dp, nullptr, &vq, hits.data(), hits.size(), &nHits);
                if (rc == FPM_ERR_CAPACITY) {
                    hits.resize(nHits);
                    rc = fpm_dist_hits(gpuContext(), &dp, nullptr, &vq, hits.data(), hits.size(), &nHits);
                }
                if (rc != FPM_OK) { cerr << "ERROR: " << fpm_last_error() << endl; return 1; }
                for (uint64_t h = 0; h < nHits; h++) {
                    const fpm_hit& hit = hits[h];
                    const Sketch::Reference& rref = sketchRef.getReference(hit.ref);
                    cout << rref.name;
                    if (comment) cout << ':' << rref.comment;
                    cout << '\t' << names[hit.query];
                    if (comment) cout << ':' << comments[hit.query];
                    cout << '\t' << hit.distance << '\t' << hit.pvalue << '\t' << hit.numer << '/' << (hit.denom & 0x7fffffffu) << '\n';
                }
                continue;
            }
            out.resize(nq * nRef);
            if (fpm_dist_tile(gpuContext(), &dp, nullptr, &vq, out.data()) != FPM_OK) { cerr << "ERROR: " << fpm_last_error() << endl; return 1; }
            for (uint64_t i = 0; i < nq; i++) {
                if (table) cout << names[i];
                for (uint64_t j = 0; j < nRef; j++) {
                    const fpm_pair& pair = out[i * nRef + j];
                    const bool pass = (pair.denom & FPM_PAIR_PASS) != 0;
                    if (table) {
                        cout << '\t';
                        if (pass) cout << pair.distance;
                    } else if (pass) {
                        const Sketch::Reference& rref = sketchRef.getReference(j);
                        cout << rref.name;
                        if (comment) cout << ':' << rref.comment;
                        cout << '\t' << names[i];
                        if (comment) cout << ':' << comments[i];
                        cout << '\t' << pair.distance << '\t' << pair.pvalue << '\t' << pair.numer << '/' << FPM_PAIR_DENOM(pair) << '\n';
                    }
                }
                if (table) cout << endl;
            }
        }
    }
    if (pinned) fpm_host_free(pinned);
    fpm_dist_set_reference(gpuContext(), nullptr);
    cout.flush();
    return 0;
}

int CommandDistance::run() const
{
    if (arguments.size() < 2 || options.at("help").active) {
        print();
        return 0;
    }
    bool list = options.at("list").active;
    bool table = options.at("table").active;
    bool comment = options.at("comment").active;
    double pValueMax = options.at("pvalue").getArgumentAsNumber();
    double distanceMax = options.at("distance").getArgumentAsNumber();
    bool fingerprint = options.at("fingerprint").active;

    Sketch::Parameters parameters;
    if (sketchParameterSetup(parameters, *this)) return 1;

    Sketch sketchRef;
    uint64_t lengthMax = 0;
    double randomChance = 0;
    int kMin = 0;
    string lengthMaxName;
    int warningCount = 0;
    const string& fileReference = arguments[0];
    bool isSketch = hasSuffix(fileReference, suffixSketch);
    if (isSketch) {
        for (const char* o : {"kmer", "noncanonical", "protein", "alphabet"}) {
            if (options.at(o).active) {
                cerr << "ERROR: The option -" << options.at(o).identifier << " cannot be used when a sketch is provided; it is inherited from the sketch." << endl;
                return 1;
            }
        }
    } else {
        cerr << "Sketching " << fileReference << " (provide sketch file made with \"mash sketch\" to skip)...";
    }
    vector<string> refArgVector(1, fileReference);
    bool tagMSH = containsTag(refArgVector, ".msh");
    bool tagTXT = containsTag(refArgVector, ".txt");
    if (fingerprint && tagMSH) sketchRef.initFromFiles(refArgVector, parameters);
    else if (fingerprint && tagTXT) sketchRef.initFromFingerprints(refArgVector, parameters);
    else sketchRef.initFromFiles(refArgVector, parameters);

    double lengthThreshold = (parameters.warning * sketchRef.getKmerSpace()) / (1. - parameters.warning);
    if (isSketch) {
        if (options.at("sketchSize").active) {
            if (parameters.reads && parameters.minHashesPerWindow != sketchRef.getMinHashesPerWindow()) {
                cerr << "ERROR: The sketch size must match the reference when using a bloom filter (leave this option out to inherit from the reference sketch)." << endl;
                return 1;
            }
        }
        parameters.minHashesPerWindow = sketchRef.getMinHashesPerWindow();
        parameters.kmerSize = sketchRef.getKmerSize();
        parameters.noncanonical = sketchRef.getNoncanonical();
        parameters.preserveCase = sketchRef.getPreserveCase();
        parameters.seed = sketchRef.getHashSeed();
        string alphabet;
        sketchRef.getAlphabetAsString(alphabet);
        setAlphabetFromString(parameters, alphabet.c_str());
    } else {
        for (uint64_t i = 0; i < sketchRef.getReferenceCount(); i++) {
            uint64_t length = sketchRef.getReference(i).length;
            if (length > lengthThreshold) {
                if (warningCount == 0 || length > lengthMax) {
                    lengthMax = length;
                    lengthMaxName = sketchRef.getReference(i).name;
                    randomChance = sketchRef.getRandomKmerChance(i);
                    kMin = sketchRef.getMinKmerSize(i);
                }
                warningCount++;
            }
        }
        cerr << "done.\n";
    }
    if (table) {
        cout << "#query";
        for (uint64_t i = 0; i < sketchRef.getReferenceCount(); i++) cout << '\t' << sketchRef.getReference(i).name;
        cout << endl;
    }
    vector<string> queryFiles;
    for (size_t i = 1; i < arguments.size(); i++) {
        if (list) splitFile(arguments[i], queryFiles);
        else queryFiles.push_back(arguments[i]);
    }
    // Query side given as sketch files only: stream them.  The files are mapped, rows of sketches go from
    // the mapping into a pinned panel chunk and on to the GPU, where the reference panel and its index stay resident; no
    // Reference objects, no second copy of a multi-gigabyte query set.  FPMASH_MSH_STREAM=0 keeps the loading path.
    {
        bool allSketches = !queryFiles.empty() && !fingerprint;
        for (const string& f : queryFiles) allSketches &= hasSuffix(f, suffixSketch);
        const char* e = getenv("FPMASH_MSH_STREAM");
        if (allSketches && !(e && e[0] == '0') && sketchRef.getReferenceCount() > 0) {
            const int rc = streamQueries(sketchRef, queryFiles, parameters, table, comment, distanceMax, pValueMax);
            if (rc == 0 && warningCount > 0 && !parameters.reads) warnKmerSize(parameters, *this, lengthMax, lengthMaxName, randomChance, kMin, warningCount);
            return rc;
        }
    }
    Sketch sketchQuery;
    if (fingerprint && tagMSH) sketchQuery.initFromFiles(queryFiles, parameters);
    else if (fingerprint && tagTXT) sketchQuery.initFromFingerprints(queryFiles, parameters);
    else sketchQuery.initFromFiles(queryFiles, parameters, 0, true);

    // compare(): sketchSize = min of the two (float) getMinHashesPerWindow(), CommandDistance.cpp:342-344
    uint64_t sketchSize = sketchQuery.getMinHashesPerWindow() < sketchRef.getMinHashesPerWindow() ? sketchQuery.getMinHashesPerWindow()
                                                                                                   : sketchRef.getMinHashesPerWindow();
    const uint64_t nRef = sketchRef.getReferenceCount(), nQry = sketchQuery.getReferenceCount();
    HostPanel pr, pq;
    buildPanel(sketchRef, pr);
    buildPanel(sketchQuery, pq);
    fpm_dist_params dp;
    dp.sketch_size = (uint32_t)sketchSize;
    dp.kmer_size = sketchRef.getKmerSize();
    dp.kmer_space = sketchRef.getKmerSpace();
    dp.max_distance = distanceMax;
    dp.max_pvalue = pValueMax;
    dp.sorted_unique = fingerprint ? 0 : 1;   // fp lists are unsorted: the literal loop defines the result

    // Filtered list output (-d / -v exclude some pairs, no table): only the passing pairs come back from the GPU, already
    // in writeOutput's order (CommandDistance.cpp:303-333), from ONE call over all queries -- no n x n matrix anywhere.
    const bool filtered = !table && ((distanceMax >= 0 && distanceMax < 1.) || (pValueMax >= 0 && pValueMax < 1.));
    if (filtered && nRef && nQry) {
        fpm_panel vr = pr.view(0, nRef), vq = pq.view(0, nQry);
        vector<fpm_hit> hits(max<uint64_t>(1 << 16, 16 * (nRef + nQry)));
        uint64_t nHits = 0;
        // large jobs: the pair space is cut into query x reference blocks over all GPUs (fpm_dist_hits_multi); the list comes
        // back merged in the same order
        fpm_multi* multi = fingerprint ? nullptr : gpuMulti(nRef * nQry, kMultiGpuPairs);
        auto call = [&]() { return multi ? fpm_dist_hits_multi(multi, &dp, &vr, &vq, hits.data(), hits.size(), &nHits)
                                         : fpm_dist_hits(gpuContext(), &dp, &vr, &vq, hits.data(), hits.size(), &nHits); };
        int rc = call();
        if (rc == FPM_ERR_CAPACITY) {
            hits.resize(nHits);
            rc = call();
        }
        if (rc != FPM_OK) {
            cerr << "ERROR: " << fpm_last_error() << endl;
            return 1;
        }
        for (uint64_t h = 0; h < nHits; h++) {
            const fpm_hit& hit = hits[h];
            const Sketch::Reference& rref = sketchRef.getReference(hit.ref);
            const Sketch::Reference& qref = sketchQuery.getReference(hit.query);
            cout << rref.name;
            if (comment) cout << ':' << rref.comment;
            cout << '\t' << qref.name;
            if (comment) cout << ':' << qref.comment;
            cout << '\t' << hit.distance << '\t' << hit.pvalue << '\t' << hit.numer << '/' << (hit.denom & 0x7fffffffu) << '\n';
        }
    }
    // query-major tiles of at most ~32M pairs (768 MB of results) per GPU call
    fpm_multi* multiTile = (filtered || fingerprint) ? nullptr : gpuMulti(nRef * nQry, kMultiGpuPairs);
    uint64_t rowsPerCall = nRef ? max<uint64_t>(1, ((32ull << 20) * (multiTile ? fpm_multi_size(multiTile) : 1)) / nRef) : 1;
    vector<fpm_pair> out;
    // one GPU, several query chunks: the reference panel is uploaded and indexed once and stays resident (the reference keeps
    // its whole sketch file loaded while the queries stream past, CommandDistance.cpp:163-266)
    const bool residentRef = !multiTile && !filtered && nRef && nQry > rowsPerCall;
    if (residentRef) {
        fpm_panel vr = pr.view(0, nRef);
        if (fpm_dist_set_reference(gpuContext(), &vr) != FPM_OK) {
            cerr << "ERROR: " << fpm_last_error() << endl;
            return 1;
        }
    }
    for (uint64_t q0 = 0; q0 < nQry && nRef && !filtered; q0 += rowsPerCall) {
        uint64_t nq = min(rowsPerCall, nQry - q0);
        out.resize(nq * nRef);
        fpm_panel vr = pr.view(0, nRef), vq = pq.view(q0, nq);
        if ((multiTile ? fpm_dist_tile_multi(multiTile, &dp, &vr, &vq, out.data()) : fpm_dist_tile(gpuContext(), &dp, residentRef ? nullptr : &vr, &vq, out.data())) != FPM_OK) {
            cerr << "ERROR: " << fpm_last_error() << endl;
            return 1;
        }
        // writeOutput, CommandDistance.cpp:276-333
        for (uint64_t i = 0; i < nq; i++) {
            const Sketch::Reference& qref = sketchQuery.getReference(q0 + i);
            if (table) cout << qref.name;
            for (uint64_t j = 0; j < nRef; j++) {
                const fpm_pair& pair = out[i * nRef + j];
                bool pass = (pair.denom & FPM_PAIR_PASS) != 0;
                if (table) {
                    cout << '\t';
                    if (pass) cout << pair.distance;
                } else if (pass) {
                    const Sketch::Reference& rref = sketchRef.getReference(j);
                    cout << rref.name;
                    if (comment) cout << ':' << rref.comment;
                    cout << '\t' << qref.name;
                    if (comment) cout << ':' << qref.comment;
                    cout << '\t' << pair.distance << '\t' << pair.pvalue << '\t' << pair.numer << '/' << FPM_PAIR_DENOM(pair) << '\n';
                }
            }
            if (table) cout << endl;
        }
    }
    cout.flush();
    if (warningCount > 0 && !parameters.reads) warnKmerSize(parameters, *this, lengthMax, lengthMaxName, randomChance, kMin, warningCount);
    return 0;
}

// ------------------------------------------------------------------------------------------
// mash paste (CommandPaste.cpp:214,242 = initFromFiles over .msh inputs + writeToCapnp)
// ------------------------------------------------------------------------------------------
CommandPaste::CommandPaste() : Command()
{
    name = "paste";
    summary = "Create a single sketch file from multiple sketch files.";
    description = "Create a single sketch file from multiple sketch files.";
    argumentString = "<out_prefix> <sketch> [<sketch>] ...";
    useOption("help");
    addOption("list", Option(Option::Boolean, "l", "", "Input files are lists of file names.", ""));
}

int CommandPaste::run() const
{
    if (arguments.size() < 2 || options.at("help").active) {
        print();
        return 0;
    }
    bool list = options.at("list").active;
    vector<string> files;
    for (size_t i = 1; i < arguments.size(); i++) {
        if (list) splitFile(arguments[i], files);
        else files.push_back(arguments[i]);
    }
    for (const string& file : files) {
        if (!hasSuffix(file, suffixSketch)) {
            cerr << "ERROR: The file \"" << file << "\" does not look like a sketch." << endl;
            return 1;
        }
    }
    Sketch sketch;
    Sketch::Parameters parameters;
    parameters.parallelism = 1;
    sketch.initFromFiles(files, parameters);
    string out = arguments[0];
    if (!hasSuffix(out, suffixSketch)) out += suffixSketch;
    if (access(out.c_str(), F_OK) != -1) {
        cerr << "ERROR: \"" << out << "\" exists; remove to write." << endl;
        exit(1);
    }
    cerr << "Writing " << out << "..." << endl;
    sketch.writeToCapnp(out.c_str());
    return 0;
}

// ------------------------------------------------------------------------------------------
// mash info (CommandInfo.cpp:62-346): header, tabular and the JSON dump the reference's own
// tests diff (the fork's whitespace and its stray first line, CommandInfo.cpp:148, kept).
// ------------------------------------------------------------------------------------------
CommandInfo::CommandInfo() : Command()
{
    name = "info";
    summary = "Display information about sketch files.";
    description = "Displays information about sketch files.";
    argumentString = "<sketch>";
    useOption("help");
    addOption("header", Option(Option::Boolean, "H", "", "Only show header info. Do not list each sketch. Incompatible with -d, -t and -c.", ""));
    addOption("tabular", Option(Option::Boolean, "t", "", "Tabular output (rather than padded), with no header. Incompatible with -d, -H and -c.", ""));
    addOption("counts", Option(Option::Boolean, "c", "", "Show hash count histograms for each sketch. Incompatible with -d, -H and -t.", ""));
    addOption("dump", Option(Option::Boolean, "d", "", "Dump sketches in JSON format. Incompatible with -H, -t, and -c.", ""));
}

int CommandInfo::run() const
{
    if (arguments.size() != 1 || options.at("help").active) {
        print();
        return 0;
    }
    bool header = options.at("header").active, tabular = options.at("tabular").active;
    bool counts = options.at("counts").active, dump = options.at("dump").active;
    if ((header && tabular) || (header && counts) || (tabular && counts) || (dump && (tabular || header || counts))) {
        cerr << "ERROR: The options -H, -t, -c and -d are mutually incompatible." << endl;
        return 1;
    }
    const string& file = arguments[0];
    if (!hasSuffix(file, suffixSketch)) {
        cerr << "ERROR: The file \"" << file << "\" does not look like a sketch." << endl;
        return 1;
    }
    Sketch sketch;
    Sketch::Parameters params;
    params.parallelism = 1;
    uint64_t referenceCount;
    if (header) referenceCount = sketch.initParametersFromCapnp(file.c_str());
    else {
        sketch.initFromFiles(arguments, params);
        referenceCount = sketch.getReferenceCount();
    }
    string alphabet;
    sketch.getAlphabetAsString(alphabet);
    const char* HASH = "MurmurHash3_x64_128";
    if (counts) {
        if (!sketch.hasHashCounts()) {
            cerr << "ERROR: Sketch file does not have hash counts. Re-sketch with counting enabled to use this feature." << endl;
            return 1;
        }
        cout << "#Sketch\tBin\tFrequency" << endl;
        map<uint32_t, uint64_t> histogram;
        for (uint64_t i = 0; i < sketch.getReferenceCount(); i++) {
            sketch.getReferenceHistogram(i, histogram);
            for (auto& e : histogram) cout << sketch.getReference(i).name << '\t' << e.first << '\t' << e.second << endl;
        }
        return 0;
    }
    if (dump) {
        bool use64 = sketch.getUse64();
        cout << "      \"Write JSON information : " << endl;
        cout << "{" << endl;
        cout << "  \"kmer\" : " << sketch.getKmerSize() << ',' << endl;
        cout << "  \"alphabet\" : \"" << alphabet << "\"," << endl;
        cout << "  \"preserveCase\" : " << (sketch.getPreserveCase() ? "true" : "false") << ',' << endl;
        cout << "  \"canonical\" : " << (sketch.getNoncanonical() ? "false" : "true") << ',' << endl;
        cout << "  \"sketchSize\" : " << sketch.getMinHashesPerWindow() << ',' << endl;
        cout << "  \"hashType\" : \"" << HASH << "\"," << endl;
        cout << "  \"hashBits\" : " << (use64 ? 64 : 32) << ',' << endl;
        cout << "  \"hashSeed\" : " << sketch.getHashSeed() << ',' << endl;
        cout << "  \"sketches\" :" << endl;
        cout << "  [" << endl;
        for (uint64_t i = 0; i < referenceCount; i++) {
            const Sketch::Reference& ref = sketch.getReference(i);
            cout << "    {" << endl;
            cout << "      \"name\" : \"" << ref.name << "\"," << endl;
            cout << "      \"length\" : " << ref.length << ',' << endl;
            cout << "      \"comment\" : \"" << ref.comment << "\"," << endl;
            cout << "      \"hashes\" :" << endl;
            cout << "      [" << endl;
            for (size_t j = 0; j < ref.hashesSorted.size(); j++) {
                cout << "        " << ref.hashesSorted.at(j);
                if (j < ref.hashesSorted.size() - 1) cout << ',';
                cout << endl;
            }
            cout << "      ]" << (sketch.hasHashCounts() ? "," : "") << endl;
            if (sketch.hasHashCounts()) {
                cout << "      \"counts\" :" << endl;
                cout << "      [" << endl;
                for (size_t j = 0; j < ref.counts.size(); j++) {
                    cout << "        " << ref.counts.at(j);
                    if (j < ref.counts.size() - 1) cout << ',';
                    cout << endl;
                }
                cout << "      ]" << endl;
            }
            cout << (i < referenceCount - 1 ? "    }," : "    }") << endl;
        }
        cout << "  ]" << endl;
        cout << "}" << endl;
        return 0;
    }
    if (tabular) {
        cout << "#Hashes\tLength\tID\tComment" << endl;
    } else {
        cout << "Header:" << endl;
        cout << "  Hash function (seed):          " << HASH << " (" << sketch.getHashSeed() << ")" << endl;
        cout << "  K-mer size:                    " << sketch.getKmerSize() << " (" << (sketch.getUse64() ? "64" : "32") << "-bit hashes)" << endl;
        cout << "  Alphabet:                      " << alphabet << (sketch.getNoncanonical() ? "" : " (canonical)") << (sketch.getPreserveCase() ? " (case-sensitive)" : "") << endl;
        cout << "  Target min-hashes per sketch:  " << sketch.getMinHashesPerWindow() << endl;
        cout << "  Sketches:                      " << referenceCount << endl;
    }
    if (!header) {
        if (!tabular) cout << endl << "Sketches:" << endl << "  [Hashes]  [Length]  [ID]  [Comment]" << endl;
        for (uint64_t i = 0; i < sketch.getReferenceCount(); i++) {
            const Sketch::Reference& ref = sketch.getReference(i);
            if (tabular) cout << ref.hashesSorted.size() << '\t' << ref.length << '\t' << ref.name << '\t' << ref.comment << endl;
            else cout << "  " << ref.hashesSorted.size() << "  " << ref.length << "  " << ref.name << "  " << ref.comment << endl;
        }
    }
    return 0;
}

// ------------------------------------------------------------------------------------------
// mash triangle (CommandTriangle.cpp:38-263): lower-triangular all-vs-all over one sketch set with
// the same comparison kernel as `dist`; `-fp` uses the fork's positional compareFingerprints.
// ------------------------------------------------------------------------------------------
CommandTriangle::CommandTriangle() : Command()
{
    name = "triangle";
    summary = "Estimate a lower-triangular distance matrix.";
    description = "Estimate the distance of each input sequence or fingerprint to every other input. Outputs a lower-triangular distance matrix in relaxed Phylip format. The input sequences can be fasta or fastq, gzipped or not, or Mash sketch files (.msh) with matching k-mer sizes. Input files can also be files of file names (see -l). If more than one input file is provided, whole files are compared by default (see -i).";
    argumentString = "<seq1> [<seq2>] ...";
    useOption("help");
    addOption("list", Option(Option::Boolean, "l", "Input", "List input. Lines in each <query> specify paths to sequence files, one per line. The reference file is not affected.", ""));
    addOption("comment", Option(Option::Boolean, "C", "Output", "Use comment fields for sequence names instead of IDs.", ""));
    addOption("edge", Option(Option::Boolean, "E", "Output", "Output edge list instead of Phylip matrix, with fields [seq1, seq2, dist, p-val, shared-hashes].", ""));
    addOption("pvalue", Option(Option::Number, "v", "Output", "Maximum p-value to report in edge list. Implies -E.", "1.0", 0., 1.));
    addOption("distance", Option(Option::Number, "d", "Output", "Maximum distance to report in edge list. Implies -E.", "1.0", 0., 1.));
    addOption("fingerprint", Option(Option::Boolean, "fp", "Input", "Indicates that the input files are fingerprints instead of sequences.", ""));
    useSketchOptions();
}

int CommandTriangle::run() const
{
    if (arguments.size() < 1 || options.at("help").active) {
        print();
        return 0;
    }
    bool list = options.at("list").active;
    bool comment = options.at("comment").active;
    bool edge = options.at("edge").active;
    bool fingerprint = options.at("fingerprint").active;
    double pValueMax = options.at("pvalue").getArgumentAsNumber();
    double distanceMax = options.at("distance").getArgumentAsNumber();
    double pValuePeak = 0;
    if (options.at("pvalue").active || options.at("distance").active) edge = true;

    Sketch::Parameters parameters;
    if (sketchParameterSetup(parameters, *this)) return 1;
    if (arguments.size() == 1 && !list) parameters.concatenated = false;

    vector<string> queryFiles;
    for (size_t i = 0; i < arguments.size(); i++) {
        if (list) splitFile(arguments[i], queryFiles);
        else queryFiles.push_back(arguments[i]);
    }
    Sketch sketch;
    if (fingerprint && containsTag(queryFiles, ".msh")) sketch.initFromFiles(queryFiles, parameters);
    else if (fingerprint) sketch.initFromFingerprints(queryFiles, parameters);
    else sketch.initFromFiles(queryFiles, parameters);

    uint64_t lengthMax = 0;
    double randomChance = 0;
    int kMin = 0;
    string lengthMaxName;
    int warningCount = 0;
    double lengthThreshold = (parameters.warning * sketch.getKmerSpace()) / (1. - parameters.warning);
    for (uint64_t i = 0; i < sketch.getReferenceCount(); i++) {
        uint64_t length = sketch.getReference(i).length;
        if (length > lengthThreshold) {
            if (warningCount == 0 || length > lengthMax) {
                lengthMax = length;
                lengthMaxName = sketch.getReference(i).name;
                randomChance = sketch.getRandomKmerChance(i);
                kMin = sketch.getMinKmerSize(i);
            }
            warningCount++;
        }
    }
    const uint64_t n = sketch.getReferenceCount();
    if (n == 0) return 0;
    if (!edge) {
        cout << '\t' << n << endl;
        cout << (comment ? sketch.getReference(0).comment : sketch.getReference(0).name) << endl;
    }
    HostPanel panel;
    buildPanel(sketch, panel);
    fpm_dist_params dp;
    dp.sketch_size = (uint32_t)sketch.getMinHashesPerWindow();
    dp.kmer_size = sketch.getKmerSize();
    dp.kmer_space = sketch.getKmerSpace();
    dp.max_distance = distanceMax;
    dp.max_pvalue = pValueMax;
    dp.sorted_unique = 1;
    // row blocks [i0, i1) against columns [0, i1-1): only the lower triangle (plus block padding) is computed
    vector<fpm_pair> out;
    for (uint64_t i0 = 1; i0 < n;) {
        uint64_t i1 = i0 + 1;
        while (i1 < n && (i1 - i0 + 1) * i1 <= (32ull << 20)) i1++;
        const uint64_t nq = i1 - i0, nr = i1 - 1;
        out.resize(nq * nr);
        fpm_panel vr = panel.view(0, nr), vq = panel.view(i0, nq);
        int rc = fingerprint ? fpm_fp_positional_tile(gpuContext(), &dp, &vr, &vq, out.data())
                             : fpm_dist_tile(gpuContext(), &dp, &vr, &vq, out.data());
        if (rc != FPM_OK) {
            cerr << "ERROR: " << fpm_last_error() << endl;
            return 1;
        }
        for (uint64_t i = i0; i < i1; i++) {                        // writeOutput, CommandTriangle.cpp:202-236
            const Sketch::Reference& ref = sketch.getReference(i);
            if (!edge) cout << (comment ? ref.comment : ref.name);
            for (uint64_t j = 0; j < i; j++) {
                const fpm_pair& pair = out[(i - i0) * nr + j];
                if (edge) {
                    if (pair.denom & FPM_PAIR_PASS) {
                        const Sketch::Reference& qry = sketch.getReference(j);
                        cout << (comment ? ref.comment : ref.name) << '\t' << (comment ? qry.comment : qry.name) << '\t' << pair.distance
                             << '\t' << pair.pvalue << '\t' << pair.numer << '/' << FPM_PAIR_DENOM(pair) << '\n';
                    }
                } else {
                    cout << '\t' << pair.distance;
                }
                if (pair.pvalue > pValuePeak) pValuePeak = pair.pvalue;
            }
            if (!edge) cout << '\n';
        }
        i0 = i1;
    }
    cout.flush();
    if (!edge) cerr << "Max p-value: " << pValuePeak << endl;
    if (warningCount > 0 && !parameters.reads) warnKmerSize(parameters, *this, lengthMax, lengthMaxName, randomChance, kMin, warningCount);
    return 0;
}

// ------------------------------------------------------------------------------------------
// mash fingerprint: lyn2vec's `--type basic --type_factorization CFL --shift shift --rev_comb true`
// (README.md:34-52; lyn2vec/fingerprint_utils.py:256-308,312-376,443-476) with the factorisation on the
// GPU, so the producer of `-fp` inputs no longer needs Python.  Output rows: "<id> n1 n2 ...".
// ------------------------------------------------------------------------------------------
CommandFingerprint::CommandFingerprint() : Command()
{
    name = "fingerprint";
    summary = "Lyndon (CFL / ICFL) fingerprints of sequences, the input of `sketch -fp` (lyn2vec basic mode).";
    description = "For every record of each FASTA input, factorise every circular window of -w characters (-t CFL: Duval's algorithm; ICFL: inverse Lyndon factorisation; CFL_ICFL-<C>: CFL with factors longer than C sub-factorised by ICFL; CFL_COMB, ICFL_COMB, CFL_ICFL_COMB-<C>: the same refined by the factorisation of the window's reverse complement) and write one line of factor lengths per window, in the format of lyn2vec's fingerprint_<type>.txt. The row id is the second word of the record header followed by _0, as lyn2vec writes it with --rev_comb true.";
    argumentString = "<fasta> [<fasta>] ...";
    useOption("help");
    addOption("window", Option(Option::Integer, "w", "", "Window length of the circular shifts.", "100", 1, 256));
    addOption("type", Option(Option::File, "t", "", "Factorisation (lyn2vec --type_factorization): CFL, ICFL, CFL_ICFL-<C>, CFL_COMB, ICFL_COMB or CFL_ICFL_COMB-<C> (lyn2vec offers C = 10, 20, 30).", "CFL"));
    addOption("prefix", Option(Option::File, "o", "Output", "Output file (default: fingerprint_<type>.txt).", ""));
}

int CommandFingerprint::run() const
{
    if (arguments.size() == 0 || options.at("help").active) {
        print();
        return 0;
    }
    const uint32_t window = (uint32_t)options.at("window").getArgumentAsNumber();
    const string type = options.at("type").argument;
    int factorization = FPM_FACT_CFL;
    uint32_t subLen = 0;
    if (type == "CFL") factorization = FPM_FACT_CFL;
    else if (type == "ICFL") factorization = FPM_FACT_ICFL;
    else if (type.compare(0, 9, "CFL_ICFL-") == 0 && type.size() > 9 && type.find_first_not_of("0123456789", 9) == string::npos) {
        factorization = FPM_FACT_CFL_ICFL;
        subLen = (uint32_t)atoi(type.c_str() + 9);
    } else if (type == "CFL_COMB") factorization = FPM_FACT_CFL_COMB;
    else if (type == "ICFL_COMB") factorization = FPM_FACT_ICFL_COMB;
    else if (type.compare(0, 14, "CFL_ICFL_COMB-") == 0 && type.size() > 14 && type.find_first_not_of("0123456789", 14) == string::npos) {
        factorization = FPM_FACT_CFL_ICFL_COMB;
        subLen = (uint32_t)atoi(type.c_str() + 14);
    } else {
        cerr << "ERROR: unknown factorisation \"" << type << "\" (CFL, ICFL, CFL_ICFL-<C>, CFL_COMB, ICFL_COMB, CFL_ICFL_COMB-<C>)." << endl;
        return 1;
    }
    vector<string> ids;
    vector<uint8_t> seq;
    vector<uint64_t> off{0};
    for (const string& file : arguments) {
        ifstream in(file);
        if (!in) {
            cerr << "ERROR: could not open " << file << " for reading." << endl;
            return 1;
        }
        // read_fasta (fingerprint_utils.py:256-308): id = 2nd word of the header; sequence lines joined, upper-cased
        string line;
        bool open = false;
        while (getline(in, line)) {
            if (!line.empty() && line[0] == '>') {
                if (open) off.push_back(seq.size());
                size_t a = line.find_first_not_of(" \t\r", line.find_first_of(" \t\r"));
                string id = a == string::npos ? line.substr(1) : line.substr(a, line.find_first_of(" \t\r", a) - a);
                ids.push_back(id + "_0");
                open = true;
            } else if (open) {
                for (char ch : line) {
                    if (ch == '\r' || ch == '\n') continue;
                    seq.push_back((uint8_t)((ch > 96 && ch < 123) ? ch - 32 : ch));
                }
            }
        }
        if (open) off.push_back(seq.size());
    }
    const uint32_t n = (uint32_t)ids.size();
    if (n == 0) {
        cerr << "ERROR: no FASTA records found." << endl;
        return 1;
    }
    seq.push_back(0);
    vector<uint64_t> woff(n + 1);
    if (fpm_fingerprint_batch(gpuContext(), seq.data(), off.data(), n, window, factorization, subLen, 42, 0, nullptr, nullptr, nullptr, woff.data()) != FPM_OK) {
        cerr << "ERROR: " << fpm_last_error() << endl;
        return 1;
    }
    const uint64_t nWin = woff[n];
    vector<uint16_t> tok(nWin * window), ntok(nWin);
    if (fpm_fingerprint_batch(gpuContext(), seq.data(), off.data(), n, window, factorization, subLen, 42, 0, nullptr, tok.data(), ntok.data(), woff.data()) != FPM_OK) {
        cerr << "ERROR: " << fpm_last_error() << endl;
        return 1;
    }
    const string outName = options.at("prefix").argument != "" ? options.at("prefix").argument : "fingerprint_" + type + ".txt";
    ofstream out(outName);
    if (!out) {
        cerr << "ERROR: could not open " << outName << " for writing." << endl;
        return 1;
    }
    string row;
    for (uint32_t r = 0; r < n; r++) {
        for (uint64_t w = woff[r]; w < woff[r + 1]; w++) {
            row = ids[r];
            row += ' ';                                      // lbl_id_gene = id + ' '; lengths joined by ' '
            for (uint16_t t = 0; t < ntok[w]; t++) {
                if (t) row += ' ';
                row += to_string(tok[w * window + t]);
            }
            row += '\n';
            out << row;
        }
    }
    cerr << "Writing to " << outName << "..." << endl;
    return 0;
}

}  // namespace mash
