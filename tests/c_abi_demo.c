/* tests/c_abi_demo.c — the drop-in boundary used from plain C99, as a maintainer of the reference would:
 * sqg_init with a profile_t-shaped struct and a model_t-shaped table, one gen_sig-shaped call, one batch call with
 * svb-zd and ss:Z: text output, one batch of reads named by coordinates against a GPU-resident genome.  Prints "nodevice <code>" and exits 3 when there is no GPU (no CPU fallback), else
 * "ok <samples> <kmers> <svb bytes> <text bytes>".  Built and run by tests/test_abi.py and tests/test_gpu_parity.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sqg.h"

int main(void) {
    /* dna-r9-prom, reference src/sim.c:67-78 */
    sqg_config_t cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.profile.digitisation = 2048; cfg.profile.sample_rate = 4000; cfg.profile.bps = 450; cfg.profile.range = 748.5801;
    cfg.profile.offset_mean = -237.4102; cfg.profile.offset_std = 14.1575;
    cfg.profile.median_before_mean = 214.2890337; cfg.profile.median_before_std = 18.0127916;
    cfg.profile.dwell_mean = 9.0; cfg.profile.dwell_std = 4.0;
    cfg.kmer_size = 6; cfg.num_kmer = 4096; cfg.amp_noise = 1.0f; cfg.seed = 1; cfg.rng_mode = SQG_RNG_PHILOX;
    sqg_model_t *model = (sqg_model_t *)malloc(4096 * sizeof *model);
    for (int i = 0; i < 4096; i++) { model[i].level_mean = 60.0f + (float)(i % 70); model[i].level_stdv = 1.0f + (float)(i % 3); }

    sqg_ctx_t *ctx = NULL;
    int rc = sqg_init(&ctx, &cfg, model);
    if (rc != SQG_OK) {
        printf("nodevice %d %s\n", rc, sqg_last_error(NULL));
        free(model);
        return rc == SQG_ERR_NODEVICE ? 3 : 4;
    }
    const char *read = "ACGTTGCATGCATGCAAACCCGGGTTTACGATCGATCGATTAGCTAGCTAGGATCGATCGGCTAGCTAGCATCGACTGACTAGCTAGCATCGATCGA";
    double offset, median_before;
    int64_t len = 0, ss_n = 0;
    int32_t *ss = NULL;
    int16_t *raw = sqg_gen_sig(ctx, read, (int32_t)strlen(read), &offset, &median_before, &len, 0, &ss, &ss_n);
    if (!raw || len <= 0 || ss_n != (int64_t)strlen(read) - 6 + 1) { printf("gen_sig failed\n"); return 5; }
    int64_t sum = 0;
    for (int64_t i = 0; i < ss_n; i++) sum += ss[i];
    if (sum != len) { printf("dwell sum %lld != len %lld\n", (long long)sum, (long long)len); return 6; }

    int64_t off[3] = {0, (int64_t)strlen(read), (int64_t)(2 * strlen(read))};
    char *two = (char *)malloc(2 * strlen(read) + 1);
    strcpy(two, read); strcat(two, read);
    sqg_result_t r;
    rc = sqg_gen_batch(ctx, 2, two, off, 0, SQG_WANT_SVB | SQG_WANT_SS_TEXT, &r);
    if (rc != SQG_OK || r.n_reads != 2 || r.signal != NULL || !r.svb || !r.ss_text) { printf("gen_batch failed %d\n", rc); return 7; }
    if (r.len_raw_signal[0] != len) { printf("batch read 0 differs from gen_sig\n"); return 8; }
    uint32_t n0;
    memcpy(&n0, r.svb + r.svb_off[0], 4);
    if ((int64_t)n0 != len) { printf("svb header %u\n", n0); return 9; }
    const int64_t svb_bytes = r.svb_len[0], text_bytes = r.ss_text_off[1] - r.ss_text_off[0];  /* r is valid until the next batch */
    /* the same read named by coordinates against a genome resident on the GPU (forward: identical signal; reverse: its
     * reverse complement comes back through SQG_WANT_BASES) */
    const int64_t coff[2] = {0, (int64_t)strlen(read)};
    rc = sqg_genome_load(ctx, 1, read, coff, NULL, NULL);
    if (rc != SQG_OK) { printf("genome_load failed %d %s\n", rc, sqg_last_error(ctx)); return 10; }
    sqg_coord_t co[2];
    memset(co, 0, sizeof co);
    co[0].contig = 0; co[0].pos = 0; co[0].len = (int32_t)strlen(read); co[0].strand = '+';
    co[1] = co[0]; co[1].strand = '-';
    sqg_result_t rc2;
    rc = sqg_gen_batch_coords(ctx, 2, co, 0, 0, SQG_WANT_BASES, &rc2);
    if (rc != SQG_OK || rc2.n_reads != 2 || !rc2.bases || !rc2.signal) { printf("gen_batch_coords failed %d\n", rc); return 11; }
    if (rc2.len_raw_signal[0] != len || memcmp(rc2.signal + rc2.sig_off[0], raw, (size_t)len * sizeof(int16_t)) != 0) {
        printf("coordinate read differs from gen_sig\n"); return 12;
    }
    if (memcmp(rc2.bases + rc2.bases_off[0], read, strlen(read)) != 0) { printf("forward bases differ\n"); return 13; }
    {
        const size_t n = strlen(read);
        const char *rv = rc2.bases + rc2.bases_off[1];
        for (size_t i = 0; i < n; i++) {
            const char f = read[n - 1 - i], want = f == 'A' ? 'T' : f == 'C' ? 'G' : f == 'G' ? 'C' : 'A';
            if (rv[i] != want) { printf("reverse complement differs at %d\n", (int)i); return 14; }
        }
    }
    printf("ok %lld %lld %lld %lld\n", (long long)len, (long long)ss_n, (long long)svb_bytes, (long long)text_bytes);
    free(raw); free(ss); free(two); free(model);
    sqg_destroy(ctx);
    return 0;
}
