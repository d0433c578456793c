/* main.c -- entry of the `sigfish-b200` binary: `sigfish-b200 dtw [OPTIONS] genome.fa reads.blow5`.
 * Sub-command dispatch and the trailer lines follow reference src/main.c:64-102; `dtw` (the hot
 * path) and `eval` (the scorer the reference's test scripts grade `dtw` output with) exist here. */
#include <stdlib.h>
#include <string.h>

#include "sfhost.h"

static int usage(FILE *fp)
{
    fprintf(fp, "Usage: sigfish-b200 <command> [options]\n\n");
    fprintf(fp, "command:\n");
    fprintf(fp, "         dtw           map raw signal reads to a reference by subsequence DTW on B200 GPUs\n");
    fprintf(fp, "         eval          compare a test set of mappings (PAF) with a truth set\n\n");
    return fp == stdout ? EXIT_SUCCESS : EXIT_FAILURE;
}

int main(int argc, char *argv[])
{
    const double t0 = sf_realtime();
    int ret = 1;
    if (argc < 2)
        return usage(stderr);
    if (!strcmp(argv[1], "dtw")) {
        ret = dtw_main(argc - 1, argv + 1);
    } else if (!strcmp(argv[1], "eval")) {
        ret = eval_main(argc - 1, argv + 1);
    } else if (!strcmp(argv[1], "--version") || !strcmp(argv[1], "-V")) {
        fprintf(stdout, "sigfish %s\n", SFHOST_VERSION);
        return EXIT_SUCCESS;
    } else if (!strcmp(argv[1], "--help") || !strcmp(argv[1], "-h")) {
        return usage(stdout);
    } else {
        fprintf(stderr, "[sigfish-b200] Unrecognised command %s\n", argv[1]);
        return usage(stderr);
    }
    fprintf(stderr, "[%s] Version: %s\n", __func__, SFHOST_VERSION);
    fprintf(stderr, "[%s] CMD:", __func__);
    for (int i = 0; i < argc; i++)
        fprintf(stderr, " %s", argv[i]);
    fprintf(stderr, "\n[%s] Real time: %.3f sec; CPU time: %.3f sec; Peak RAM: %.3f GB\n\n", __func__,
            sf_realtime() - t0, sf_cputime(), sf_peakrss() / 1024.0 / 1024.0 / 1024.0);
    return ret;
}
