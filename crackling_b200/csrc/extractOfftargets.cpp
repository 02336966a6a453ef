// extractOfftargets -- drop-in replacement for Crackling's off-target site extractor
// (/root/reference/src/crackling/utils/extractOfftargets.py, run as `python -m crackling.utils.extractOfftargets`),
// host program over the C ABI of libissl_cuda.
//
//   extractOfftargets <output> <inputs...> [--maxOpenFiles N] [--threads N]
//
// Same arguments (a single directory argument means every file in it, :200-206; a single input file takes the
// tool's multi-FASTA path, :208-222), same output: one sorted line of 20 bases per site on either strand.
// --maxOpenFiles and --threads are accepted and ignored: there are no intermediate files and no process pool.
// Matching, reverse complement and the sort run on a B200; there is no CPU path.
//
// Extension: --index <file.issl> [--slice-width W] also builds the ISSL index (isslCreateIndex <output> 20 W <file>)
// from the sites while they are still in HBM; `-` as <output> then skips the text file altogether.
// Environment: ISSL_DEVICE=<ordinal> (default 0), ISSL_TIMING=1.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dirent.h>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

#include "issl_cuda.h"

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

static int usage(const char *argv0)
{
    fprintf(stderr, "usage: %s [-h] [--maxOpenFiles MAXOPENFILES] [--threads THREADS] [--index ISSL] [--slice-width W] output inputs [inputs ...]\n", argv0);
    return 2;   // argparse exits with 2
}

int main(int argc, char **argv)
{
    std::vector<std::string> positional;
    std::string indexPath;
    int sliceWidth = 8;
    for (int a = 1; a < argc; a++) {
        const std::string arg = argv[a];
        auto value = [&](const char *name) -> const char * {
            const size_t n = strlen(name);
            if (arg.compare(0, n, name) == 0 && arg.size() > n && arg[n] == '=') return argv[a] + n + 1;
            if (arg == name && a + 1 < argc) return argv[++a];
            return nullptr;
        };
        if (arg == "-h" || arg == "--help") { usage(argv[0]); return 0; }
        if (arg.rfind("--maxOpenFiles", 0) == 0) { if (!value("--maxOpenFiles")) return usage(argv[0]); continue; }
        if (arg.rfind("--threads", 0) == 0) { if (!value("--threads")) return usage(argv[0]); continue; }
        if (arg.rfind("--index", 0) == 0) { const char *v = value("--index"); if (!v) return usage(argv[0]); indexPath = v; continue; }
        if (arg.rfind("--slice-width", 0) == 0) { const char *v = value("--slice-width"); if (!v) return usage(argv[0]); sliceWidth = atoi(v); continue; }
        positional.push_back(arg);
    }
    if (positional.size() < 2) return usage(argv[0]);
    const std::string output = positional[0];
    std::vector<std::string> inputs(positional.begin() + 1, positional.end());
    const bool timing = getenv("ISSL_TIMING") && atoi(getenv("ISSL_TIMING")) != 0;
    const int device = getenv("ISSL_DEVICE") ? atoi(getenv("ISSL_DEVICE")) : 0;
    const double t0 = now_s();

    struct stat st;
    if (inputs.size() == 1 && stat(inputs[0].c_str(), &st) == 0 && S_ISDIR(st.st_mode)) {   // ref :200-206
        const std::string dir = inputs[0];
        inputs.clear();
        if (DIR *d = opendir(dir.c_str())) {
            while (dirent *e = readdir(d))
                if (e->d_name[0] != '.') inputs.push_back(dir + "/" + e->d_name);   // glob('*') skips dot files
            closedir(d);
        }
        std::sort(inputs.begin(), inputs.end());
        if (inputs.empty()) { fprintf(stderr, "No input files in %s\n", dir.c_str()); return 1; }
    }
    const int singleInput = inputs.size() == 1;   // ref :208-222
    printf("Extracting off-targets on the GPU\n");
    printf("Beginning to process %zu files\n", inputs.size());

    issl_sites *sites = nullptr;
    if (issl_sites_create(device, &sites) != ISSL_OK) { fprintf(stderr, "%s\n", issl_last_error()); return 1; }
    for (const std::string &path : inputs) {
        const int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0 || fstat(fd, &st) != 0) { fprintf(stderr, "Cannot read %s\n", path.c_str()); return 1; }
        if (S_ISDIR(st.st_mode)) { close(fd); continue; }
        const size_t bytes = (size_t)st.st_size;
        if (bytes) {
            void *text = mmap(nullptr, bytes, PROT_READ, MAP_PRIVATE, fd, 0);
            if (text == MAP_FAILED) { fprintf(stderr, "Cannot read %s\n", path.c_str()); return 1; }
            madvise(text, bytes, MADV_SEQUENTIAL);
            const int rc = issl_sites_add_fasta(sites, (const char *)text, bytes, singleInput);
            munmap(text, bytes);
            if (rc != ISSL_OK) { fprintf(stderr, "%s\n", issl_last_error()); return 1; }
        }
        close(fd);
    }
    uint64_t n = 0, chars = 0;
    issl_sites_count(sites, &n, &chars);
    const double t1 = now_s();
    printf("Processing completed. Found %llu targets.\n", (unsigned long long)n);
    if (output != "-" || indexPath.empty()) {
        if (issl_sites_write_text(sites, output.c_str()) != ISSL_OK) { fprintf(stderr, "%s\n", issl_last_error()); return 1; }
    }
    const double t2 = now_s();
    if (!indexPath.empty()) {
        issl_device *dev = nullptr;
        if (issl_device_create_from_sites(sites, (uint32_t)sliceWidth, ISSL_LAYOUT_GATHER /* ids-only lists: all the .issl writer needs */, &dev) != ISSL_OK ||
            issl_device_write_issl(dev, indexPath.c_str()) != ISSL_OK) {
            fprintf(stderr, "%s\n", issl_last_error());
            return 1;
        }
        issl_device_destroy(dev);
    }
    issl_sites_destroy(sites);
    printf("Goodbye.\n");
    if (timing)
        fprintf(stderr, "[issl] %llu sequence characters, %llu sites: extract %.3f s, sort+write %.3f s, index %.3f s, total %.3f s\n",
                (unsigned long long)chars, (unsigned long long)n, t1 - t0, t2 - t1, now_s() - t2, now_s() - t0);
    return 0;
}
