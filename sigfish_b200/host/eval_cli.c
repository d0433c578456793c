/* eval_cli.c -- `sigfish eval truth.paf test.paf`: compares a test set of mappings with a truth set.
 *
 * Same command line and report as the reference's eval sub-tool (reference src/eval.c:17-24 options,
 * 219-242 the correctness rule, 270-327 the per-mapping classification, 329-362 the report), which the
 * reference's own test scripts use to grade `dtw` output (test/test.sh:24-43).  Rules:
 *   - truth mappings are grouped by read id; a test mapping of a read absent from the truth set counts as
 *     "only_in_testset";
 *   - a test mapping is correct when some truth mapping of the same read (with --secondary no: of the same
 *     tp:A type) has the same target name and strand and min(|d start|, |d end|) < 100 (--tid-only: name and
 *     strand are enough);
 *   - correct / incorrect counts are also broken down by MAPQ (0..60).
 * Host-only code: nothing here touches the GPU.
 */
#include <getopt.h>
#include <stdlib.h>
#include <string.h>

#include "sfhost.h"

typedef struct {
    char *rid;
    char *tid;
    int32_t target_start, target_end;
    int8_t strand;
    uint8_t mapq;
    char tp;
} paf_t;

typedef struct read_node {
    char *rid;
    paf_t **recs;
    int n, cap;
    struct read_node *next;
} read_node_t;

typedef struct {
    read_node_t **bucket;
    size_t n_bucket, n_reads;
} truth_t;

static struct option eval_options[] = {{"verbose", required_argument, 0, 'v'}, {"help", no_argument, 0, 'h'},
                                       {"version", no_argument, 0, 'V'},       {"output", required_argument, 0, 'o'},
                                       {"secondary", required_argument, 0, 0}, {"tid-only", no_argument, 0, 0},
                                       {0, 0, 0, 0}};

static uint64_t hash_str(const char *s)
{
    uint64_t h = 1469598103934665603ull; /* FNV-1a */
    for (; *s; s++)
        h = (h ^ (unsigned char)*s) * 1099511628211ull;
    return h;
}

static char *next_field(char **save)
{
    char *f = strtok_r(NULL, "\t\r\n", save);
    if (!f) {
        SF_FATAL("%s", "malformed PAF record: fewer than 12 columns");
    }
    return f;
}

static paf_t *parse_paf(char *line)
{
    char *save = NULL;
    char *f = strtok_r(line, "\t\r\n", &save);
    if (!f)
        return NULL; /* blank line */
    paf_t *p = (paf_t *)calloc(1, sizeof(paf_t));
    p->rid = strdup(f);
    next_field(&save); /* query length */
    next_field(&save); /* query start */
    next_field(&save); /* query end */
    f = next_field(&save);
    if (!strcmp(f, "+"))
        p->strand = 0;
    else if (!strcmp(f, "-"))
        p->strand = 1;
    else
        SF_FATAL("malformed PAF record: strand '%s'", f);
    p->tid = strdup(next_field(&save));
    next_field(&save); /* target length */
    p->target_start = atoi(next_field(&save));
    p->target_end = atoi(next_field(&save));
    next_field(&save); /* residue matches */
    next_field(&save); /* block length */
    p->mapq = (uint8_t)atoi(next_field(&save));
    p->tp = 'P';
    while ((f = strtok_r(NULL, "\t\r\n", &save))) {
        if (!strcmp(f, "tp:A:P"))
            p->tp = 'P';
        else if (!strcmp(f, "tp:A:S"))
            p->tp = 'S';
    }
    return p;
}

static void free_paf(paf_t *p)
{
    free(p->rid);
    free(p->tid);
    free(p);
}

static read_node_t *truth_find(const truth_t *t, const char *rid)
{
    for (read_node_t *n = t->bucket[hash_str(rid) % t->n_bucket]; n; n = n->next)
        if (!strcmp(n->rid, rid))
            return n;
    return NULL;
}

static int same_locus(const paf_t *a, const paf_t *b, int tid_only)
{
    if (strcmp(a->tid, b->tid) != 0 || a->strand != b->strand)
        return 0;
    if (tid_only)
        return 1;
    int ds = a->target_start - b->target_start, de = a->target_end - b->target_end;
    if (ds < 0) ds = -ds;
    if (de < 0) de = -de;
    return (de < ds ? de : ds) < 100;
}

int eval_main(int argc, char *argv[])
{
    int longindex = 0, c;
    FILE *fp_help = stderr;
    int use_secondary = 1, tid_only = 0;
    while ((c = getopt_long(argc, argv, "o:hV", eval_options, &longindex)) >= 0) {
        if (c == 'V') {
            fprintf(stdout, "sigfish %s\n", SFHOST_VERSION);
            exit(EXIT_SUCCESS);
        } else if (c == 'h') {
            fp_help = stdout;
        } else if (c == 0 && longindex == 4) {
            if (!strcmp(optarg, "yes") || !strcmp(optarg, "y"))
                use_secondary = 1;
            else if (!strcmp(optarg, "no") || !strcmp(optarg, "n"))
                use_secondary = 0;
            else {
                SF_WARNING("option '--%s' only accepts 'yes' or 'no'.", "secondary");
                use_secondary = 0; /* the reference's yes_or_no() returns 0 here */
            }
        } else if (c == 0 && longindex == 5) {
            tid_only = 1;
        }
    }
    if (argc - optind < 2 || fp_help == stdout) {
        fprintf(fp_help, "Usage: sigfish eval truth.paf test.paf\n");
        fprintf(fp_help, "\nbasic options:\n");
        fprintf(fp_help, "   -h                         help\n");
        fprintf(fp_help, "   --version                  print version\n");
        fprintf(fp_help, "   --secondary STR            consider secondary mappings. yes or no.\n");
        fprintf(fp_help, "   --tid-only                 consider regerence name and strand only\n");
        exit(fp_help == stdout ? EXIT_SUCCESS : EXIT_FAILURE);
    }

    FILE *fp = fopen(argv[optind], "r");
    if (!fp)
        SF_FATAL("cannot open %s. ", argv[optind]);
    truth_t truth;
    truth.n_bucket = 1 << 16;
    truth.n_reads = 0;
    truth.bucket = (read_node_t **)calloc(truth.n_bucket, sizeof(read_node_t *));
    char *line = NULL;
    size_t cap = 0;
    long truth_rec = 0;
    while (getline(&line, &cap, fp) != -1) {
        paf_t *p = parse_paf(line);
        if (!p)
            continue;
        read_node_t *n = truth_find(&truth, p->rid);
        if (!n) {
            n = (read_node_t *)calloc(1, sizeof(read_node_t));
            n->rid = strdup(p->rid);
            const size_t b = hash_str(p->rid) % truth.n_bucket;
            n->next = truth.bucket[b];
            truth.bucket[b] = n;
            truth.n_reads++;
        }
        if (n->n == n->cap) {
            n->cap = n->cap ? n->cap * 2 : 2;
            n->recs = (paf_t **)realloc(n->recs, sizeof(paf_t *) * (size_t)n->cap);
        }
        n->recs[n->n++] = p;
        truth_rec++;
    }
    fclose(fp);
    (void)truth_rec;

    fp = fopen(argv[optind + 1], "r");
    if (!fp)
        SF_FATAL("cannot open %s. ", argv[optind + 1]);
    long test_rec = 0, correct = 0, incorrect = 0, only_in_test = 0;
    long by_mapq[2][61];
    memset(by_mapq, 0, sizeof by_mapq);
    while (getline(&line, &cap, fp) != -1) {
        paf_t *p = parse_paf(line);
        if (!p)
            continue;
        const read_node_t *n = truth_find(&truth, p->rid);
        if (!n) {
            only_in_test++;
        } else {
            int ok = 0;
            for (int i = 0; i < n->n && !ok; i++)
                if (use_secondary || n->recs[i]->tp == p->tp)
                    ok = same_locus(n->recs[i], p, tid_only);
            if (p->mapq > 60)
                SF_FATAL("MAPQ %d out of range in %s", p->mapq, argv[optind + 1]);
            if (ok) {
                correct++;
                by_mapq[0][p->mapq]++;
            } else {
                incorrect++;
                by_mapq[1][p->mapq]++;
            }
        }
        free_paf(p);
        test_rec++;
    }
    fclose(fp);
    free(line);
    fprintf(stderr, "Total mappings in testset: %d\n", (int)test_rec);

    printf("\nComparison between truthset and testset\n"
           "mapped_truthset\t%ld\n"
           "mapped_testset\t%ld (%.2f%%)\n"
           "correct\t%ld (%.2f%%)\n"
           "incorrect\t%ld (%.2f%%)\n"
           "only_in_testset\t%ld\n",
           (long)truth.n_reads, test_rec, test_rec / (float)truth.n_reads * 100, correct, correct / (float)test_rec * 100,
           incorrect, incorrect / (float)test_rec * 100, only_in_test);
    printf("\n#mapq\tcorrect\tincorrect\n");
    for (int q = 60; q >= 0; q--)
        if (by_mapq[0][q] || by_mapq[1][q])
            printf("%d\t%d\t%d\n", q, (int)by_mapq[0][q], (int)by_mapq[1][q]);

    for (size_t b = 0; b < truth.n_bucket; b++) {
        read_node_t *n = truth.bucket[b];
        while (n) {
            read_node_t *nx = n->next;
            for (int i = 0; i < n->n; i++)
                free_paf(n->recs[i]);
            free(n->recs);
            free(n->rid);
            free(n);
            n = nx;
        }
    }
    free(truth.bucket);
    return 0;
}
