// `autocycler compress` and `autocycler decompress` with the reference's flags (main.rs:126-162), messages and exit codes
// (misc.rs:130-136: "Error: <text>" on stderr, exit 1), running the B200 path through the C ABI.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/autocycler_gpu.h"

static void usage() {
    fprintf(stderr,
            "Usage: autocycler compress --assemblies_dir <ASSEMBLIES_DIR> --autocycler_dir <AUTOCYCLER_DIR> [OPTIONS]\n\n"
            "Options:\n"
            "  -i, --assemblies_dir <DIR>   Directory containing input assemblies (required)\n"
            "  -a, --autocycler_dir <DIR>   Autocycler directory to be created (required)\n"
            "      --kmer <KMER>            K-mer size for De Bruijn graph [default: 51]\n"
            "      --max_contigs <N>        refuse to run if mean contigs per assembly exceeds this value [default: 25]\n"
            "  -t, --threads <THREADS>      Number of CPU threads (end repair) [default: 8]\n"
            "      --device <ORDINAL>       CUDA device [default: 0]\n"
            "      --devices <A,B,...>      several CUDA devices of this box: the assemblies are sharded by file over them\n");
}

// `autocycler decompress` (main.rs:150-162, decompress.rs:27-57)
static int decompress_main(int argc, char** argv) {
    std::string in, out_dir, out_file; int device = 0;
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        auto value = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "error: a value is required for '%s'\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "-i" || a == "--in_gfa") in = value();
        else if (a == "-o" || a == "--out_dir") out_dir = value();
        else if (a == "-f" || a == "--out_file") out_file = value();
        else if (a == "--device") device = atoi(value());
        else { fprintf(stderr, "error: unexpected argument '%s'\nUsage: autocycler decompress --in_gfa <IN_GFA> [--out_dir <DIR>] [--out_file <FASTA>]\n", a.c_str()); return 2; }
    }
    if (in.empty()) { fprintf(stderr, "Usage: autocycler decompress --in_gfa <IN_GFA> [--out_dir <DIR>] [--out_file <FASTA>]\n"); return 2; }
    fprintf(stderr, "\nStarting autocycler decompress (%s)\n\nSettings:\n  --in_gfa %s\n", ac_version(), in.c_str());
    if (!out_dir.empty()) fprintf(stderr, "  --out_dir %s\n", out_dir.c_str());
    if (!out_file.empty()) fprintf(stderr, "  --out_file %s\n", out_file.c_str());
    fprintf(stderr, "\n");
    const int rc = ac_decompress_gfa(in.c_str(), out_dir.empty() ? nullptr : out_dir.c_str(), out_file.empty() ? nullptr : out_file.c_str(), device, 1);
    if (rc != AC_OK) { fprintf(stderr, "\nError: %s\n", ac_last_error(nullptr)); return 1; }
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && strcmp(argv[1], "decompress") == 0) return decompress_main(argc, argv);
    if (argc < 2 || strcmp(argv[1], "compress") != 0) { usage(); return 2; }
    std::string in, out; unsigned k = 51, max_contigs = 25, threads = 8; int device = 0;
    std::vector<int32_t> devices;
    for (int i = 2; i < argc; ++i) {
        std::string a = argv[i];
        auto value = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "error: a value is required for '%s'\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "-i" || a == "--assemblies_dir") in = value();
        else if (a == "-a" || a == "--autocycler_dir") out = value();
        else if (a == "--kmer") k = (unsigned)strtoul(value(), nullptr, 10);
        else if (a == "--max_contigs") max_contigs = (unsigned)strtoul(value(), nullptr, 10);
        else if (a == "-t" || a == "--threads") threads = (unsigned)strtoul(value(), nullptr, 10);
        else if (a == "--device") device = atoi(value());
        else if (a == "--devices") { devices.clear(); for (const char* q = value(); *q;) { devices.push_back((int32_t)strtol(q, (char**)&q, 10)); if (*q == ',') ++q; else if (*q) { fprintf(stderr, "error: --devices wants a comma-separated list of ordinals\n"); return 2; } } }
        else if (a == "-h" || a == "--help") { usage(); return 0; }
        else { fprintf(stderr, "error: unexpected argument '%s'\n", a.c_str()); usage(); return 2; }
    }
    if (in.empty() || out.empty()) { usage(); return 2; }
    fprintf(stderr, "\nStarting autocycler compress (%s)\n\nSettings:\n  --assemblies_dir %s\n  --autocycler_dir %s\n  --kmer %u\n  --threads %u\n\n",
            ac_version(), in.c_str(), out.c_str(), k, threads);
    if (devices.empty()) devices.push_back(device);
    const int node = ac_bind_host_to_device(devices[0]);   // stay on the socket of the GPU that finishes the graph; harmless when the topology cannot be read
    if (node >= 0) fprintf(stderr, "host threads bound to NUMA node %d (device %d)\n\n", node, devices[0]);
    int rc = ac_compress_dir_devices(in.c_str(), out.c_str(), k, max_contigs, threads, devices.data(), (int32_t)devices.size(), 1);
    if (rc != AC_OK) { fprintf(stderr, "\nError: %s\n", ac_last_error(nullptr)); return 1; }
    return 0;
}
