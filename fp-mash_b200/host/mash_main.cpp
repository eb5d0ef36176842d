// mash_main.cpp -- `mash <command> [options] ...` dispatcher (mash.cpp:18-40, CommandList.cpp).
// Commands on the accelerated path: sketch, dist (+ paste, info as thin users of the same Sketch).
#include <string.h>
#include <iostream>
#include <map>
#include <memory>

#include "Command.h"
#include "Sketch.h"
#include "fastx.h"

int main(int argc, const char** argv)
{
    std::map<std::string, std::unique_ptr<mash::Command>> commands;
    commands["sketch"].reset(new mash::CommandSketch());
    commands["dist"].reset(new mash::CommandDistance());
    commands["paste"].reset(new mash::CommandPaste());
    commands["info"].reset(new mash::CommandInfo());
    commands["triangle"].reset(new mash::CommandTriangle());
    commands["fingerprint"].reset(new mash::CommandFingerprint());
    if (argc >= 2 && (strcmp(argv[1], "--version") == 0)) {
        std::cout << "2.3 (fp-mash hot path, B200-native)" << std::endl;
        return 0;
    }
    if (argc == 3 && strcmp(argv[1], "debug-parse") == 0) {
        // test hook: dump what the FASTA/FASTQ reader sees (parity with kseq.h is tested against it)
        gzFile fp = gzopen(argv[2], "r");
        if (!fp) return 2;
        FastxReader rd(fp);
        int64_t l;
        while ((l = rd.next()) >= 0) {
            unsigned long long h = 1469598103934665603ull;
            for (char c : rd.seq) h = (h ^ (unsigned char)c) * 1099511628211ull;
            std::cout << rd.name << '\t' << rd.comment << '\t' << rd.comment_cstr << '\t' << l << '\t' << h << '\n';
        }
        std::cout << "END\t" << l << std::endl;
        gzclose(fp);
        return 0;
    }
    if (argc < 2 || commands.count(argv[1]) == 0) {
        std::cout << std::endl << "Mash (fp-mash fork) -- B200-native sketch/dist" << std::endl << std::endl;
        std::cout << "Usage:" << std::endl << std::endl << "  mash <command> [options] [arguments ...]" << std::endl << std::endl << "Commands:" << std::endl << std::endl;
        for (auto& c : commands) std::cout << "  " << c.first << "  " << c.second->summary << std::endl;
        std::cout << std::endl;
        return argc < 2 ? 0 : 1;
    }
    fpmTick("main");
    int rc = commands[argv[1]]->run(argc - 2, argv + 2);
    fpmTick("command done");
    return rc;
}
