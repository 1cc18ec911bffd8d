// `modkit` drop-in entry point for the pileup path: `modkit pileup [flags] <in.bam> <out.bed>`
// (src/bin/main.rs:15-29, src/commands.rs:59-160). Other subcommands are out of scope.
#include "pileup_run.hpp"

extern "C" int mkh_pileup_main(int argc, const char* const* argv);

int main(int argc, char** argv) {
    if (argc < 2 || std::string(argv[1]) != "pileup") {
        fprintf(stderr, "Usage: modkit pileup [OPTIONS] <IN_BAM> <OUT_BED>\n(only the pileup subcommand is provided by this build)\n");
        return 2;
    }
    return mkh_pileup_main(argc - 2, argv + 2);
}
