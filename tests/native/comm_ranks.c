/* The multi-GPU reduce of include/bydb_gpu.h driven from plain C, one PROCESS per rank, no torch, no NCCL -- the way a Go data
 * node would use it through cgo: every rank opens its context, exports its mailbox handle, the handles travel over a pipe (in
 * BanyanDB: the cluster's own gRPC), the ranks connect and run bydb_scan_reduce collectively; rank 0 compares the reduced
 * answer with one context scanning all shards.  usage: comm_ranks <nranks> <ndevices>  (ranks share devices round-robin).
 * CUDA must not be touched before fork(): the parent only forks, relays the handles and collects the exit codes. */
#define _POSIX_C_SOURCE 200809L
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include "bydb_gpu.h"
#include "bydb_synth.h"

#define MAXR 8
#define SERIES_PER_RANK 6
#define POINTS 3000
#define GROUPS 4

static int read_all(int fd, void *buf, size_t n) {
    char *p = buf;
    while (n) {
        ssize_t k = read(fd, p, n);
        if (k <= 0) return -1;
        p += k;
        n -= (size_t)k;
    }
    return 0;
}
static int write_all(int fd, const void *buf, size_t n) {
    const char *p = buf;
    while (n) {
        ssize_t k = write(fd, p, n);
        if (k <= 0) return -1;
        p += k;
        n -= (size_t)k;
    }
    return 0;
}

static bydb_part_image *shard_image(int r) {
    static bydb_synth_field flds[2] = {{"latency", BYDB_SYN_F_LATENCY, 0}, {"calls", BYDB_SYN_I_FLUCT, 0}};
    bydb_synth_spec sp;
    memset(&sp, 0, sizeof sp);
    sp.n_series = SERIES_PER_RANK; sp.n_points = POINTS; sp.sid0 = 1 + (uint64_t)r * SERIES_PER_RANK; sp.sid_step = 1;
    sp.t0 = 1700000000000000000LL; sp.t_step = 60000000000LL; sp.n_fields = 2; sp.fields = flds; sp.seed = 4242;
    bydb_part_image *img = NULL;
    return bydb_synth_part(&sp, &img) == 0 ? img : NULL;
}
static int register_image(bydb_ctx *ctx, uint64_t id, bydb_part_image *img, bydb_part_h *h) {
    bydb_file files[16];
    uint32_t n = bydb_part_image_n_files(img);
    if (n > 16) return -1;
    for (uint32_t i = 0; i < n; ++i) {
        files[i].name = bydb_part_image_file_name(img, i);
        files[i].data = bydb_part_image_file_data(img, i, &files[i].len);
    }
    bydb_part_files pf = {n, files};
    return bydb_part_register(ctx, id, &pf, h);
}
static void fill_query(bydb_query *q, const bydb_part_h *parts, uint32_t n_parts, const uint64_t *sids, const int32_t *grp, uint64_t ns, const bydb_agg *aggs) {
    memset(q, 0, sizeof *q);
    q->parts = parts; q->n_parts = n_parts; q->series_ids = sids; q->series_group = grp; q->n_series = ns; q->n_groups = GROUPS;
    q->aggs = aggs; q->n_aggs = 4; q->tmin = 1700000000000000000LL + 100 * 60000000000LL; q->tmax = 1700000000000000000LL + 2500 * 60000000000LL;
    q->top_n = 3; q->top_agg = 0; q->top_desc = 1;
}

static int rank_main(int rank, int nranks, int ndev, int to_parent, int from_parent) {
    bydb_cfg cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.device = rank % ndev;
    bydb_ctx *ctx = NULL;
    if (bydb_init(&cfg, &ctx) != 0) { fprintf(stderr, "rank %d: init: %s\n", rank, bydb_last_error()); return 2; }
    bydb_agg aggs[4] = {{"latency", BYDB_AGG_SUM, 0}, {"latency", BYDB_AGG_MEAN, 0}, {"calls", BYDB_AGG_MAX, 0}, {"calls", BYDB_AGG_COUNT, 0}};
    uint64_t all_sids[MAXR * SERIES_PER_RANK];
    int32_t all_grp[MAXR * SERIES_PER_RANK];
    for (int i = 0; i < nranks * SERIES_PER_RANK; ++i) { all_sids[i] = 1 + (uint64_t)i; all_grp[i] = i % GROUPS; }
    bydb_query probe;
    fill_query(&probe, NULL, 0, all_sids, all_grp, (uint64_t)nranks * SERIES_PER_RANK, aggs);
    bydb_partials_layout_t lay;
    if (bydb_partials_layout(&probe, &lay) != 0) return 3;
    bydb_comm_handle mine, all[MAXR];
    if (bydb_comm_export(ctx, lay.total_bytes, nranks, &mine) != 0) { fprintf(stderr, "rank %d: export: %s\n", rank, bydb_last_error()); return 4; }
    if (write_all(to_parent, &mine, sizeof mine) || read_all(from_parent, all, sizeof(bydb_comm_handle) * (size_t)nranks)) return 5;
    if (bydb_comm_connect(ctx, rank, nranks, all) != 0) { fprintf(stderr, "rank %d: connect: %s\n", rank, bydb_last_error()); return 6; }
    bydb_part_image *img = shard_image(rank);
    bydb_part_h h = 0;
    if (!img || register_image(ctx, 1, img, &h) != 0) { fprintf(stderr, "rank %d: register: %s\n", rank, bydb_last_error()); return 7; }
    const uint64_t *my_sids = all_sids + rank * SERIES_PER_RANK;
    const int32_t *my_grp = all_grp + rank * SERIES_PER_RANK;
    bydb_query q;
    fill_query(&q, &h, 1, my_sids, my_grp, SERIES_PER_RANK, aggs);
    int fails = 0;
    for (int iter = 0; iter < 5; ++iter) {     /* several epochs: slot parities alternate, the root's `done` word gates the reuse */
        const int root = iter % nranks;
        bydb_result res;
        int rc = bydb_scan_reduce(ctx, &q, root, &res);
        if (rc != 0) { fprintf(stderr, "rank %d iter %d: scan_reduce: %d %s\n", rank, iter, rc, bydb_last_error()); return 8; }
        if (rank != root) {
            if (res.n_rows != 0 || res.stats.blocks_scanned == 0) ++fails;
            bydb_result_free(ctx, &res);
            continue;
        }
        /* the root checks the reduced rows against ONE context scanning all the shards */
        bydb_part_h hs[MAXR];
        bydb_part_image *imgs[MAXR];
        for (int r = 0; r < nranks; ++r) {
            imgs[r] = shard_image(r);
            if (!imgs[r] || register_image(ctx, 100 + (uint64_t)(iter * MAXR + r), imgs[r], &hs[r]) != 0) return 9;
        }
        bydb_query whole;
        fill_query(&whole, hs, (uint32_t)nranks, all_sids, all_grp, (uint64_t)nranks * SERIES_PER_RANK, aggs);
        bydb_result want;
        if (bydb_scan_agg(ctx, &whole, &want) != 0) { fprintf(stderr, "whole scan: %s\n", bydb_last_error()); return 10; }
        if (res.n_rows != want.n_rows || res.n_rows != 3) ++fails;
        for (int i = 0; i < res.n_rows && i < want.n_rows; ++i) {
            if (res.group_id[i] != want.group_id[i] || res.rows[i] != want.rows[i]) ++fails;
            for (int a = 0; a < 4; ++a) {
                const int k = i * 4 + a;
                if (res.is_float[a] != want.is_float[a]) ++fails;
                if (res.is_float[a]) {
                    if (fabs(res.val_f64[k] - want.val_f64[k]) > 1e-12 * fabs(want.val_f64[k])) ++fails;
                } else if (res.val_i64[k] != want.val_i64[k]) {
                    ++fails;
                }
            }
        }
        bydb_result_free(ctx, &want);
        bydb_result_free(ctx, &res);
        for (int r = 0; r < nranks; ++r) {
            bydb_part_release(ctx, hs[r]);
            bydb_part_image_free(imgs[r]);
        }
    }
    bydb_part_release(ctx, h);
    bydb_part_image_free(img);
    bydb_shutdown(ctx);
    if (fails) fprintf(stderr, "rank %d: %d mismatches\n", rank, fails);
    return fails ? 11 : 0;
}

int main(int argc, char **argv) {
    const int nranks = argc > 1 ? atoi(argv[1]) : 2, ndev = argc > 2 ? atoi(argv[2]) : 1;
    if (nranks < 1 || nranks > MAXR || ndev < 1) return 64;
    int up[MAXR][2], down[MAXR][2];
    pid_t pids[MAXR];
    for (int r = 0; r < nranks; ++r) {
        if (pipe(up[r]) || pipe(down[r])) return 65;
        pids[r] = fork();
        if (pids[r] < 0) return 66;
        if (pids[r] == 0) {
            close(up[r][0]);
            close(down[r][1]);
            _exit(rank_main(r, nranks, ndev, up[r][1], down[r][0]));
        }
        close(up[r][1]);
        close(down[r][0]);
    }
    bydb_comm_handle all[MAXR];
    int bad = 0;
    for (int r = 0; r < nranks; ++r)
        if (read_all(up[r][0], &all[r], sizeof all[r])) bad = 1;
    for (int r = 0; r < nranks; ++r)
        if (bad || write_all(down[r][1], all, sizeof(bydb_comm_handle) * (size_t)nranks)) close(down[r][1]);
    int status = 0, worst = bad ? 67 : 0;
    for (int r = 0; r < nranks; ++r) {
        waitpid(pids[r], &status, 0);
        const int code = WIFEXITED(status) ? WEXITSTATUS(status) : 99;
        if (code) { fprintf(stderr, "rank %d exited with %d\n", r, code); worst = code; }
    }
    printf(worst ? "FAILED\n" : "OK %d ranks on %d device(s), 5 collective calls, roots rotated\n", nranks, ndev);
    return worst;
}
