/* examples/shard_host.c -- the row-sharded hot path driven from plain C: one process per
 * GPU, NCCL id handed over through a file.  Build:
 *     gcc -std=c99 -I include examples/shard_host.c -L pydem_b200 -lpydem_b200 -lm -o shard_host
 * Run (2 GPUs):  for r in 0 1; do LD_LIBRARY_PATH=pydem_b200 ./shard_host $r 2 /tmp/nccl_id & done; wait
 * It builds a tilted plane with a valley (every cell has an analytic neighbour), runs slope/aspect -> UCA -> TWI over
 * the shards and prints each rank's largest contributing area.  tests/test_abi.py compiles and links it (no GPU needed
 * for that); scripts/c_abi_shard_check.py is the same program with a full comparison against the single-tile result. */
#define _DEFAULT_SOURCE     /* usleep */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include "pydem_b200.h"

#define CHECK(call) do { int rc__ = (call); if (rc__) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, pdm_last_error()); return 1; } } while (0)

int main(int argc, char **argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s rank nranks id_file [rows cols]\n", argv[0]); return 2; }
    const int rank = atoi(argv[1]), nranks = atoi(argv[2]);
    const char *id_file = argv[3];
    const int64_t R = argc > 4 ? atoll(argv[4]) : 2048, C = argc > 5 ? atoll(argv[5]) : 1024;
    CHECK(pdm_init(rank));
    unsigned char id[PDM_COMM_ID_BYTES];
    if (rank == 0) {
        CHECK(pdm_comm_unique_id(id));
        char tmp[512];
        snprintf(tmp, sizeof(tmp), "%s.tmp", id_file);
        FILE *f = fopen(tmp, "wb");
        if (!f || fwrite(id, 1, sizeof(id), f) != sizeof(id)) { perror("id file"); return 1; }
        fclose(f);
        rename(tmp, id_file);
    } else {
        FILE *f = NULL;
        for (int tries = 0; tries < 1200 && !(f = fopen(id_file, "rb")); tries++) usleep(50000);
        if (!f || fread(id, 1, sizeof(id), f) != sizeof(id)) { fprintf(stderr, "no NCCL id from rank 0\n"); return 1; }
        fclose(f);
    }
    CHECK(pdm_comm_init(rank, nranks, id));

    /* my block of rows [r0, r1) + one halo row on every side where another rank continues the grid */
    const int64_t r0 = R * rank / nranks, r1 = R * (rank + 1) / nranks;
    const int top = rank > 0, bot = rank < nranks - 1;
    const int64_t row_off = r0 - top, Rl = (r1 - r0) + top + bot, lo = top, hi = lo + (r1 - r0);
    const double d = 30.0;
    double *dXY = malloc(sizeof(double) * (size_t)R), *th = malloc(sizeof(double) * (size_t)R), *th_row = malloc(sizeof(double) * (size_t)R);
    for (int64_t i = 0; i < R; i++) { dXY[i] = d; th[i] = atan2(d, d); th_row[i] = atan2(d, d); }
    pdm_tile *t = NULL;
    CHECK(pdm_tile_create(Rl, C, NULL, &t));
    CHECK(pdm_tile_set_spacing(t, dXY, dXY, dXY, dXY, th, th));                 /* R-1 fences / R posts of the local rows */
    CHECK(pdm_tile_set_window(t, row_off, R, lo, hi, th_row));
    CHECK(pdm_tile_set_global_spacing(t, dXY, dXY, R - 1));
    CHECK(pdm_shard_connect(t));

    double *elev = malloc(sizeof(double) * (size_t)(Rl * C));
    for (int64_t i = 0; i < Rl; i++)
        for (int64_t j = 0; j < C; j++) {
            const double gi = (double)(row_off + i), centre = C / 2.0 + (C / 8.0) * sin(gi / 40.0);
            elev[i * C + j] = 2.0 * ((double)R - gi) + 50.0 + 0.9 * fabs((double)j - centre);
        }
    CHECK(pdm_tile_upload(t, PDM_F_ELEV, elev));
    pdm_uca_params p;  pdm_default_uca_params(&p);
    pdm_twi_params tw; pdm_default_twi_params(&tw);
    tw.twi_min_area = d * d;
    pdm_uca_stats st;
    int label_rounds = 0;
    CHECK(pdm_shard_run(t, &p, &tw, &st, &label_rounds));
    double *uca = malloc(sizeof(double) * (size_t)(Rl * C));
    CHECK(pdm_tile_download(t, PDM_F_UCA, uca));
    double best = 0.0;
    for (int64_t i = lo; i < hi; i++)
        for (int64_t j = 0; j < C; j++)
            if (uca[i * C + j] > best) best = uca[i * C + j];
    printf("rank %d/%d rows [%lld, %lld): drained %lld cells, undone %lld, largest contributing area %.0f cells\n", rank, nranks,
           (long long)r0, (long long)r1, (long long)st.n_drained, (long long)st.n_undone, best / (d * d));
    CHECK(pdm_shard_disconnect(t));
    CHECK(pdm_tile_destroy(t));
    CHECK(pdm_comm_finalize());
    free(elev); free(uca); free(dXY); free(th); free(th_row);
    return 0;
}
