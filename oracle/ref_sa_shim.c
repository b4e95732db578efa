/*
 * oracle/ref_sa_shim.c -- TEST INFRASTRUCTURE ONLY.
 * The one function of the reference's CPU bwa that cannot work as shipped: bwt_restore_sa (bwa_index/bwt.c:501-527) still reads
 * (n_sa - 1) x sizeof(bwtint_t) bytes and never restores pack_size / pack_mask / sa_bits, while bwt_dump_sa (bwa_index/bwt.c:472-487)
 * writes u32 samples + pack_size + the packed high bits and bwt_sa (bwa_index/bwt.c:151-172) reads exactly that layout
 * (`[fread] Unexpected end of file`).  oracle/build_ref.sh compiles the reference's bwt.c with
 * -Dbwt_restore_sa=bwt_restore_sa_as_shipped and links this reader instead, giving oracle/_ref/bwa7p: the reference's own
 * `bwa mem` / `bwa fastmap`, every other line unmodified, able to load the index it builds.  No reference code is copied here.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include "bwt.h"

void bwt_restore_sa(const char *fn, bwt_t *bwt)
{
    FILE *fp = fopen(fn, "rb");
    uint64_t hdr[7];
    if (!fp || fread(hdr, 8, 7, fp) != 7) { fprintf(stderr, "[ref_sa_shim] cannot read %s\n", fn); exit(1); }
    if (hdr[0] != bwt->primary) { fprintf(stderr, "SA-BWT inconsistency: primary is not the same.\n"); exit(1); }
    if (hdr[6] != bwt->seq_len) { fprintf(stderr, "SA-BWT inconsistency: seq_len is not the same.\n"); exit(1); }
    bwt->sa_intv = (int)hdr[5];
    bwt->n_sa = (bwt->seq_len + bwt->sa_intv) / bwt->sa_intv;
    bwt->sa = (uint32_t *)calloc(bwt->n_sa, 4);
    bwt->sa[0] = (uint32_t)-1;
    uint8_t ps = 1;
    if (fread(bwt->sa + 1, 4, bwt->n_sa - 1, fp) != bwt->n_sa - 1 || fread(&ps, 1, 1, fp) != 1) { fprintf(stderr, "[ref_sa_shim] %s is truncated\n", fn); exit(1); }
    bwt->pack_size = ps;
    bwt->pack_mask = ps >= 32 ? 0xffffffffu : ((1u << ps) - 1);
    if ((bwt->seq_len >> 32) == 0) bwt->pack_mask = 0;          /* bwa_index/bwt.c:88-91: msb == 0 */
    size_t nhi = (size_t)ps * bwt->n_sa / 32 + 1;
    bwt->sa_bits = (uint32_t *)calloc(nhi + 1, 4);
    size_t got = fread(bwt->sa_bits, 4, nhi, fp); (void)got;
    fclose(fp);
}
