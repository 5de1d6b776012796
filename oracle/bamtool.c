/* Tiny helper for the test oracle, linked against the reference's own libbam.a (samtools 0.1.x):
 *   bamtool sam2bam in.sam out.bam    -- what `samtools view -Sb` does in example/seeksv.sh
 *   bamtool index   in.bam            -- what `samtools index` does (writes in.bam.bai)
 *   bamtool bam2sam in.bam out.sam    -- text dump, for debugging fixtures
 *   bamtool depth   in.bam minMapQ    -- the multi-pileup loop of bam2depth.cpp:72-96 on its own: prints
 *                                        "tid pos1 n_plp n_del_or_skip" per covered position (probe for
 *                                        the pileup semantics of the linked libbam: flag mask, =/X ops,
 *                                        the 8000-read cap)
 *   bamtool auxi    in.bam TAG        -- bam_aux2i(bam_aux_get(b, TAG)) per record (probe: how the linked libbam walks
 *                                        the aux block, i.e. which types it knows how to skip)
 *   bamtool calend  pos 10M2D5X...    -- bam_calend() of the linked libbam for a CIGAR (probe: which ops
 *                                        advance the reference end in this build of the library)
 * Test infrastructure only (oracle/): never linked or executed by the product. */
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include "bam.h"
#include "sam.h"
typedef struct { bamFile fp; int min_mapQ; } depth_aux_t;
static int depth_read(void *data, bam1_t *b)
{
    depth_aux_t *aux = (depth_aux_t*)data;
    int ret = bam_read1(aux->fp, b);
    if ((int)b->core.qual < aux->min_mapQ) b->core.flag |= BAM_FUNMAP;
    return ret;
}
static int depth_main(const char *fn, int min_mapQ)
{
    depth_aux_t aux, *auxp = &aux;
    bam_header_t *h;
    bam_mplp_t mplp;
    const bam_pileup1_t *plp;
    int tid, pos, n_plp;
    aux.fp = bam_open(fn, "r");
    if (!aux.fp) return 1;
    aux.min_mapQ = min_mapQ;
    h = bam_header_read(aux.fp);
    mplp = bam_mplp_init(1, depth_read, (void**)&auxp);
    while (bam_mplp_auto(mplp, &tid, &pos, &n_plp, &plp) > 0) {
        int j, m = 0;
        for (j = 0; j < n_plp; ++j) if (plp[j].is_del || plp[j].is_refskip) ++m;
        printf("%d\t%d\t%d\t%d\n", tid, pos + 1, n_plp, m);
    }
    bam_mplp_destroy(mplp);
    bam_header_destroy(h);
    bam_close(aux.fp);
    return 0;
}
int main(int argc, char **argv)
{
    if (argc >= 4 && strcmp(argv[1], "depth") == 0) return depth_main(argv[2], atoi(argv[3]));
    if (argc >= 3 && strcmp(argv[1], "index") == 0) return bam_index_build(argv[2]);
    if (argc >= 4 && (strcmp(argv[1], "sam2bam") == 0 || strcmp(argv[1], "bam2sam") == 0)) {
        int tobam = strcmp(argv[1], "sam2bam") == 0;
        samfile_t *in = samopen(argv[2], tobam ? "r" : "rb", 0);
        if (!in || !in->header) { fprintf(stderr, "bamtool: cannot open %s\n", argv[2]); return 1; }
        samfile_t *out = samopen(argv[3], tobam ? "wb" : "wh", in->header);
        if (!out) { fprintf(stderr, "bamtool: cannot write %s\n", argv[3]); return 1; }
        bam1_t *b = bam_init1();
        long n = 0;
        while (samread(in, b) >= 0) { samwrite(out, b); ++n; }
        bam_destroy1(b);
        samclose(out);
        samclose(in);
        fprintf(stderr, "bamtool: %ld records\n", n);
        return 0;
    }
    if (argc >= 4 && strcmp(argv[1], "auxi") == 0) {  /* bam_aux2i(bam_aux_get(b, TAG)) for every record (probe of the aux walk) */
        samfile_t *in = samopen(argv[2], "rb", 0);
        if (!in || !in->header) { fprintf(stderr, "bamtool: cannot open %s\n", argv[2]); return 1; }
        bam1_t *b = bam_init1();
        while (samread(in, b) >= 0) {
            uint8_t *s = bam_aux_get(b, argv[3]);
            printf("%s\t%d\t%d\n", bam1_qname(b), s ? 1 : 0, (int)bam_aux2i(s));
            fflush(stdout);  /* the walk can run off the record on malformed input: keep what was printed */
        }
        bam_destroy1(b);
        samclose(in);
        return 0;
    }
    if (argc >= 4 && strcmp(argv[1], "calend") == 0) {
        bam1_core_t c;
        uint32_t cig[64];
        int n = 0;
        const char *p = argv[3];
        memset(&c, 0, sizeof(c));
        c.pos = atoi(argv[2]);
        while (*p && n < 64) {
            char *e;
            long len = strtol(p, &e, 10);
            const char *ops = "MIDNSHP=X", *o = strchr(ops, *e);
            if (!o) return 2;
            cig[n++] = (uint32_t)len << 4 | (uint32_t)(o - ops);
            p = e + 1;
        }
        c.n_cigar = n;
        printf("%u\n", bam_calend(&c, cig));
        return 0;
    }
    fprintf(stderr, "usage: bamtool sam2bam in.sam out.bam | bam2sam in.bam out.sam | index in.bam\n");
    return 2;
}
