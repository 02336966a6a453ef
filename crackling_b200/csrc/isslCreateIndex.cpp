// isslCreateIndex -- drop-in replacement for Crackling's ISSL index builder
// (/root/reference/src/ISSL/isslCreateIndex.cpp), host program over the C ABI of libissl_cuda.
//
//   isslCreateIndex <offtargetSites.txt> <sequence length> <slice width (bits)> <sissltable>
//
// Same four positional arguments, same sorted fixed-width text input, and a byte-identical .issl
// output (header, score table, signatures, list sizes, list entries: ref :256-289).  The collapse of
// repeated sites into occurrence counts, the slice lists and their serialisation run on a B200; there
// is no CPU path.  stdout carries the reference's four progress lines, stderr the sequence count.
//
// Environment: ISSL_DEVICE=<ordinal> (default 0), ISSL_TIMING=1 (phase timings on stderr).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "issl_cuda.h"

static double now_s()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv)
{
    if (argc < 5) {   // ref :134-137
        fprintf(stderr, "Usage: %s [offtargetSites.txt] [sequence length] [slice width (bits)] [sissltable]\n", argv[0]);
        return 1;
    }
    const bool timing = getenv("ISSL_TIMING") && atoi(getenv("ISSL_TIMING")) != 0;
    const int device = getenv("ISSL_DEVICE") ? atoi(getenv("ISSL_DEVICE")) : 0;
    const double t0 = now_s();

    const int seqLength = atoi(argv[2]);
    if (seqLength > 32) {   // ref :142-145
        fprintf(stderr, "Sequence length is greater than 32, which is the maximum supported currently\n");
        return 1;
    }
    if (seqLength <= 0) {   // the reference divides by garbage here; refuse instead
        fprintf(stderr, "Sequence length must be a positive number\n");
        return 1;
    }
    const int sliceWidth = atoi(argv[3]);

    const int fd = open(argv[1], O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) {   // the reference dereferences a NULL FILE* here
        fprintf(stderr, "Failed to read in file.\n");
        return 1;
    }
    const size_t fileSize = (size_t)st.st_size;
    const size_t seqLineLength = (size_t)seqLength + 1;
    if (fileSize % seqLineLength != 0) {   // ref :147-153
        fprintf(stderr, "fileSize: %zu\n", fileSize);
        fprintf(stderr, "Error: file does is not a multiple of the expected line length (%zu)\n", seqLineLength);
        fprintf(stderr, "The sequence length may be incorrect; alternatively, the line endings\n");
        fprintf(stderr, "may be something other than LF, or there may be junk at the end of the file.\n");
        return 1;
    }
    fprintf(stderr, "Number of sequences: %zu\n", fileSize / seqLineLength);   // ref :155
    if (fileSize == 0) {   // ref :176-179: fread of 0 bytes reports failure
        fprintf(stderr, "Failed to read in file.\n");
        return 1;
    }
    void *text = mmap(nullptr, fileSize, PROT_READ, MAP_PRIVATE, fd, 0);
    if (text == MAP_FAILED) {
        fprintf(stderr, "Failed to read in file.\n");
        return 1;
    }
    madvise(text, fileSize, MADV_SEQUENTIAL);
    const double t1 = now_s();

    issl_device *dev = nullptr;
    if (issl_device_create_from_text((const char *)text, fileSize, (uint32_t)seqLength, (uint32_t)sliceWidth, device,
                                     ISSL_LAYOUT_GATHER, &dev) != ISSL_OK) {   // ids-only lists: all the .issl writer needs (no sub-bucket copies)
        fprintf(stderr, "%s\n", issl_last_error());
        return 1;
    }
    munmap(text, fileSize);
    close(fd);
    const double t2 = now_s();
    // the reference reports its three phases one by one (:209, :237, :254); here they are one device pass
    printf("Finished counting occurrences, now constructing index...\n");
    printf("Finished constructing index, now precalculating scores...\n");
    printf("Finished calculating scores, now preparing to write to disk...\n");
    if (issl_device_write_issl(dev, argv[4]) != ISSL_OK) {
        fprintf(stderr, "%s\n", issl_last_error());
        return 1;
    }
    printf("Writing to disk...\n");   // ref :291-294
    issl_device_destroy(dev);
    printf("Done.\n");
    if (timing)
        fprintf(stderr, "[issl] open %.3f s, build on device %.3f s, write %.3f s, total %.3f s\n", t1 - t0, t2 - t1,
                now_s() - t2, now_s() - t0);
    return 0;
}
