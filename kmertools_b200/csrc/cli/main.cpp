// kmertools — drop-in for the reference CLI's `comp oligo` subcommand (kmertools/src/args.rs:70-103,
// 242-263).  Same flags, same defaults, same output bytes; compute runs on the GPU through
// libkmertools_b200.so.  Other subcommands of the reference are out of scope.
#include "../../../include/kmertools_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

static void usage() {
    fprintf(stderr,
            "Generate oligonucleotide frequency vectors (GPU)\n\n"
            "Usage: kmertools comp oligo [OPTIONS] --input <INPUT> --output <OUTPUT>\n\n"
            "Options:\n"
            "  -i, --input <INPUT>    Input file path\n"
            "  -o, --output <OUTPUT>  Output vectors path\n"
            "  -c, --counts           Disable normalisation and output raw counts\n"
            "  -k, --k-size <K_SIZE>  Set k-mer size [default: 3]\n"
            "  -r, --raw-count        Raw counts\n"
            "  -p, --preset <PRESET>  Output type to write [default: spc] [possible values: csv, tsv, spc]\n"
            "  -H, --header           Include header (with k-mer in ACGT format)\n"
            "  -t, --threads <N>      Thread count for computations 0=auto [default: 0]\n"
            "      --device <N>       CUDA device ordinal [default: 0]\n");
}

// `kmertools comp cgr -k K ...` (k-mer mode of kmertools/src/args.rs:105-128,264-283)
static int main_cgr(int argc, char **argv) {
    ktb_file_opts o{};
    std::string in, out;
    o.k = 0; o.canonical = 1; o.norm = 1; o.delim = ' '; o.device = 0;
    long vec = -1;
    for (int i = 3; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&](const char *name) -> const char * {
            if (i + 1 >= argc) { fprintf(stderr, "error: a value is required for '%s'\n", name); exit(2); }
            return argv[++i];
        };
        if (a == "-i" || a == "--input") in = val("--input");
        else if (a == "-o" || a == "--output") out = val("--output");
        else if (a == "-c" || a == "--counts") o.norm = 0;
        else if (a == "-k" || a == "--k-size") o.k = atoi(val("--k-size"));
        else if (a == "-v" || a == "--vec-size") vec = atol(val("--vec-size"));
        else if (a == "-t" || a == "--threads") o.threads = atoi(val("--threads"));
        else if (a == "--device") o.device = atoi(val("--device"));
        else { fprintf(stderr, "error: unexpected argument '%s' found\n", a.c_str()); return 2; }
    }
    if (in.empty() || out.empty()) { fprintf(stderr, "error: --input and --output are required\n"); return 2; }
    if (o.k == 0) {
        fprintf(stderr, "error: whole-sequence CGR (no -k) is outside this build's scope; use -k 3..7 for k-mer mode\n");
        return 2;
    }
    if (o.k < 3 || o.k > 7) {
        fprintf(stderr, "error: invalid value '%d' for '--k-size <K_SIZE>': %d is not in 3..=7\n", o.k, o.k);
        return 2;
    }
    if (vec < 0) vec = (long)(o.k * o.k);   // (k as f64).powf(4.0).powf(0.5) as u64
    o.in_path = in.c_str();
    o.out_path = out.c_str();
    if (ktb_comp_cgr_file(&o, (int)vec, nullptr) != KTB_OK) fprintf(stderr, "Error: %s\n", ktb_last_error());
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 3 && !strcmp(argv[1], "comp") && !strcmp(argv[2], "cgr")) return main_cgr(argc, argv);
    if (argc < 3 || strcmp(argv[1], "comp") != 0 || strcmp(argv[2], "oligo") != 0) {
        if (argc >= 2 && (!strcmp(argv[1], "-h") || !strcmp(argv[1], "--help"))) { usage(); return 0; }
        fprintf(stderr, "error: this build provides only `kmertools comp oligo` and `kmertools comp cgr -k` (GPU oligo path)\n");
        usage();
        return 2;
    }
    ktb_file_opts o{};
    std::string in, out;
    o.k = 3; o.canonical = 1; o.norm = 1; o.delim = ' '; o.header = 0; o.threads = 0; o.device = 0;
    bool k_range_checked = true;
    for (int i = 3; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&](const char *name) -> const char * {
            if (i + 1 >= argc) { fprintf(stderr, "error: a value is required for '%s'\n", name); exit(2); }
            return argv[++i];
        };
        if (a == "-i" || a == "--input") in = val("--input");
        else if (a == "-o" || a == "--output") out = val("--output");
        else if (a == "-c" || a == "--counts") o.norm = 0;
        else if (a == "-k" || a == "--k-size") o.k = atoi(val("--k-size"));
        else if (a == "-r" || a == "--raw-count") o.canonical = 0;
        else if (a == "-H" || a == "--header") o.header = 1;
        else if (a == "-t" || a == "--threads") o.threads = atoi(val("--threads"));
        else if (a == "--device") o.device = atoi(val("--device"));
        else if (a == "--any-k") k_range_checked = false;  // superset: allow k outside 3..=7
        else if (a == "-p" || a == "--preset") {
            const std::string p = val("--preset");
            if (p == "csv") o.delim = ',';
            else if (p == "tsv") o.delim = '\t';
            else if (p == "spc") o.delim = ' ';
            else { fprintf(stderr, "error: invalid value '%s' for '--preset <PRESET>' [possible values: csv, tsv, spc]\n", p.c_str()); return 2; }
        } else if (a == "-h" || a == "--help") { usage(); return 0; }
        else { fprintf(stderr, "error: unexpected argument '%s' found\n", a.c_str()); usage(); return 2; }
    }
    if (in.empty() || out.empty()) {
        fprintf(stderr, "error: the following required arguments were not provided:\n%s%s", in.empty() ? "  --input <INPUT>\n" : "",
                out.empty() ? "  --output <OUTPUT>\n" : "");
        return 2;
    }
    if (k_range_checked && (o.k < 3 || o.k > 7)) {  // clap value_parser range(3..=7), args.rs:85
        fprintf(stderr, "error: invalid value '%d' for '--k-size <K_SIZE>': %d is not in 3..=7\n", o.k, o.k);
        return 2;
    }
    o.in_path = in.c_str();
    o.out_path = out.c_str();
    ktb_file_stats st{};
    const int rc = ktb_comp_oligo_file(&o, &st);
    if (rc != KTB_OK) {
        // the reference prints the error and still exits 0 (args.rs:260-262)
        fprintf(stderr, "Error: %s\n", ktb_last_error());
        return 0;
    }
    if (getenv("KTB_VERBOSE"))
        fprintf(stderr, "records %llu bases %llu written %llu B | parse %.1f ms gpu-wait %.1f ms write %.1f ms total %.1f ms\n",
                (unsigned long long)st.records, (unsigned long long)st.bases, (unsigned long long)st.bytes_written,
                st.parse_ms, st.gpu_wait_ms, st.write_ms, st.total_ms);
    return 0;
}
