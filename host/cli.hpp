// cli.hpp — komb2 command line, flag-compatible with the reference's TCLAP setup
// (reference src/komb2.cpp:35-62): -i/--input, -j/--input2, -u/--input-unitigs
// (required, value follows after a space), -l/--readlen, -t/--threads,
// -o/--output, -f/--fulgor (accepted, ignored like the reference), --version,
// -h/--help, and `--` / --ignore_rest.  Parse errors go to stderr and exit(1),
// help/version exit(0) (TCLAP CmdLine.h:473-508 behaviour KOMB.py relies on).
#pragma once

#include <cstdio>
#include <cstdlib>
#include <ctime>
#include <string>
#include <vector>

namespace komb {

struct Options {
    std::string input, input2, input_unitigs, output;
    int readlen = 151;
    int threads = 1;
    bool fulgor = false;
};

inline std::string default_output_dir() {
    // output_<yyyymd>_<hms>, same recipe as the reference (komb2.cpp:27-32)
    time_t now = time(nullptr);
    tm *t = localtime(&now);
    return "output_" + std::to_string(1900 + t->tm_year) + std::to_string(1 + t->tm_mon) + std::to_string(t->tm_mday) +
           "_" + std::to_string(t->tm_hour) + std::to_string(t->tm_min) + std::to_string(t->tm_sec);
}

inline void print_usage(FILE *f, const char *prog, bool full) {
    fprintf(f, "%sUSAGE: \n\n   %s  [-f] [-o <string>] [-t <int>] [-l <int>] -u <string> -j <string> -i\n"
               "          <string> [--] [--version] [-h]\n\n",
            full ? "" : "Brief ", prog);
    if (!full) {
        fprintf(f, "For complete USAGE and HELP type: \n   %s --help\n\n", prog);
        return;
    }
    fprintf(f,
            "Where: \n\n"
            "   -f,  --fulgor\n     Use Fulgor pseudoalignments instead of SAM files\n\n"
            "   -o <string>,  --output <string>\n     Output directory [Default: output_yyyymmdd_hhmmss]\n\n"
            "   -t <int>,  --threads <int>\n     Number of Threads [Default: Max]\n\n"
            "   -l <int>,  --readlen <int>\n     Read Length (can be average) [Default: 151]\n\n"
            "   -u <string>,  --input-unitigs <string>\n     (required)  FASTA file containing unitigs [Default: unitigs.fa]\n\n"
            "   -j <string>,  --input2 <string>\n     (required)  Second input SAM file [Default: alignment2.sam]\n\n"
            "   -i <string>,  --input <string>\n     (required)  Input SAM file [Default: alingment1.sam]\n\n"
            "   --,  --ignore_rest\n     Ignores the rest of the labeled arguments following this flag.\n\n"
            "   --version\n     Displays version information and exits.\n\n"
            "   -h,  --help\n     Displays usage information and exits.\n\n\n"
            "   KOMB: Taxonomy-oblivious characterization of metagenome dynamics\n\n");
}

[[noreturn]] inline void parse_error(const char *prog, const std::string &arg_id, const std::string &what) {
    fprintf(stderr, "PARSE ERROR: %s\n             %s\n\n", arg_id.c_str(), what.c_str());
    print_usage(stderr, prog, false);
    exit(1);
}

inline Options parse_cli(int argc, const char **argv, const char *version, int default_threads) {
    Options o;
    o.threads = default_threads;
    o.output = default_output_dir();
    const char *prog = argc > 0 ? argv[0] : "komb2";
    std::string p(prog);
    size_t slash = p.find_last_of('/');
    if (slash != std::string::npos) p = p.substr(slash + 1);
    bool have_i = false, have_j = false, have_u = false;
    struct Spec { const char *flag; const char *name; };
    auto is = [](const std::string &a, const char *flag, const char *name) {
        return a == std::string("-") + flag || a == std::string("--") + name;
    };
    auto value = [&](int &i, const std::string &id) -> std::string {
        if (i + 1 >= argc) parse_error(p.c_str(), "Argument: " + id, "Missing a value for this argument!");
        return std::string(argv[++i]);
    };
    auto to_int = [&](const std::string &s, const std::string &id) -> int {
        char *end = nullptr;
        long v = strtol(s.c_str(), &end, 10);
        if (end == s.c_str() || *end != '\0') parse_error(p.c_str(), "Argument: " + id, "Couldn't read argument value from string '" + s + "'");
        return (int)v;
    };
    for (int i = 1; i < argc; ++i) {
        std::string a(argv[i]);
        if (a == "--" || a == "--ignore_rest") break;
        if (a == "--version") { printf("\n%s  version: %s\n\n", p.c_str(), version); exit(0); }
        if (is(a, "h", "help")) { printf("\n"); print_usage(stdout, p.c_str(), true); exit(0); }
        if (is(a, "i", "input")) { o.input = value(i, "-i (--input)"); have_i = true; }
        else if (is(a, "j", "input2")) { o.input2 = value(i, "-j (--input2)"); have_j = true; }
        else if (is(a, "u", "input-unitigs")) { o.input_unitigs = value(i, "-u (--input-unitigs)"); have_u = true; }
        else if (is(a, "l", "readlen")) { o.readlen = to_int(value(i, "-l (--readlen)"), "-l (--readlen)"); }
        else if (is(a, "t", "threads")) { o.threads = to_int(value(i, "-t (--threads)"), "-t (--threads)"); }
        else if (is(a, "o", "output")) { o.output = value(i, "-o (--output)"); }
        else if (is(a, "f", "fulgor")) { o.fulgor = true; }
        else parse_error(p.c_str(), "Argument: " + a, "Couldn't find match for argument");
    }
    std::string missing;
    if (!have_i) missing += "input";
    if (!have_j) missing += std::string(missing.empty() ? "" : ", ") + "input2";
    if (!have_u) missing += std::string(missing.empty() ? "" : ", ") + "input-unitigs";
    if (!missing.empty()) parse_error(p.c_str(), "", "Required argument" + std::string(missing.find(',') != std::string::npos ? "s" : "") + " missing: " + missing);
    if (o.threads < 1) o.threads = 1;
    return o;
}

}  // namespace komb
