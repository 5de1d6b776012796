/* Tiny helper for the test oracle, linked against the reference's own libbam.a (samtools 0.1.x):
 *   bamtool sam2bam in.sam out.bam    -- what `samtools view -Sb` does in example/seeksv.sh
 *   bamtool index   in.bam            -- what `samtools index` does (writes in.bam.bai)
 *   bamtool bam2sam in.bam out.sam    -- text dump, for debugging fixtures
 * Test infrastructure only (oracle/): never linked or executed by the product. */
#include <stdio.h>
#include <string.h>
#include "bam.h"
#include "sam.h"
int main(int argc, char **argv)
{
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
    fprintf(stderr, "usage: bamtool sam2bam in.sam out.bam | bam2sam in.bam out.sam | index in.bam\n");
    return 2;
}
