/*
 * samtools_caller.c -- the reference benchmark's "RAW SAMTOOLS" reader
 * (benchmark/flagstats.cpp:490-519 + the report of :577-588) with the per-record
 * flagstat_loop replaced by one FLAGSTAT_cuda_samtools call per block.
 *
 * Plain C against include/flagstats_cuda.h only -- what a maintainer of the reference
 * would write.  The block loop, the 1,024,000-byte block size (:119), the
 * bam_flagstat_t accumulator and the printed report are the reference's; nothing here
 * counts on the CPU (no device => the call fails and so does this program).
 *
 *   samtools_caller FILE.bin            block by block, like the reference
 *   samtools_caller --file FILE.bin     the whole file in one call (pread threads + pinned ring)
 *   samtools_caller --lz4 FILE.lz4      [int32 raw][int32 comp][LZ4 block] container, decoded on the GPU
 *   samtools_caller --zstd FILE.zst     the same around Zstandard frames (zstd_decompress_samtools, :684-728)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "flagstats_cuda.h"

#define BLOCK_BYTES 1024000 /* benchmark/flagstats.cpp:119 */

static int fail(const char* what, int rc)
{
    fprintf(stderr, "samtools_caller: %s: %s (%d)\n", what, FLAGSTAT_cuda_strerror(rc), rc);
    return 1;
}

int main(int argc, char** argv)
{
    if (argc < 2) {
        fprintf(stderr, "usage: %s [--file|--lz4|--zstd] FILE\n", argv[0]);
        return 2;
    }
    FLAGSTAT_cuda_bam_flagstat s; /* bam_flagstat_t, :43-49 */
    memset(&s, 0, sizeof s);
    unsigned long long tot_flags = 0;
    int rc;

    if (argc >= 3 && (strcmp(argv[1], "--file") == 0 || strcmp(argv[1], "--lz4") == 0 ||
                      strcmp(argv[1], "--zstd") == 0)) {
        const int fmt = (strcmp(argv[1], "--lz4") == 0    ? FLAGSTAT_CUDA_FILE_LZ4
                         : strcmp(argv[1], "--zstd") == 0 ? FLAGSTAT_CUDA_FILE_ZSTD
                                                          : FLAGSTAT_CUDA_FILE_RAW) |
                        FLAGSTAT_CUDA_FILE_SAMTOOLS;
        uint64_t f[32], n = 0;
        memset(f, 0, sizeof f);
        if ((rc = FLAGSTAT_cuda_file_u64(argv[2], fmt, f, &n))) return fail(argv[2], rc);
        if ((rc = FLAGSTAT_cuda_samtools_from_counters(f, &s))) return fail("from_counters", rc);
        tot_flags = n;
    } else {
        FILE* fp = fopen(argv[1], "rb");
        if (!fp) {
            perror(argv[1]);
            return 1;
        }
        uint16_t* inflags = (uint16_t*)malloc(BLOCK_BYTES);
        if (!inflags) return 1;
        for (;;) { /* the loop of :499-515 */
            const size_t got = fread(inflags, 1, BLOCK_BYTES, fp);
            const uint32_t N = (uint32_t)(got >> 1); /* :506 */
            if (N == 0) break;
            /* was: for (i = 0; i < N; ++i) flagstat_loop(s, inflags[i]);   :510-512 */
            if ((rc = FLAGSTAT_cuda_samtools(inflags, N, &s))) return fail("FLAGSTAT_cuda_samtools", rc);
            tot_flags += N;
            if (got < BLOCK_BYTES) break;
        }
        free(inflags);
        fclose(fp);
    }
    fprintf(stderr, "[CUDA SAMTOOLS %s] %llu flags, %llu kernel launches\n", argv[argc - 1], tot_flags,
            (unsigned long long)FLAGSTAT_cuda_launch_count());
    char report[2048];
    if ((rc = FLAGSTAT_cuda_samtools_report(&s, report, sizeof report)) < 0) return fail("report", rc);
    fputs(report, stdout); /* :577-588 */
    return 0;
}
