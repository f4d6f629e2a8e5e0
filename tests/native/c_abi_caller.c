/* A plain C99 caller of the drop-in boundary, the way a cgo shim sees it (include/bydb_gpu.h, include/bydb_synth.h): every
 * entry point is referenced with its declared signature and linked against libbydbgpu.so.  Without a GPU bydb_init must
 * fail loudly (no CPU fallback) and the host-only calls (layout, synthetic part writer) must still work. */
#include <stdio.h>
#include <string.h>

#include "bydb_gpu.h"
#include "bydb_synth.h"

int main(void) {
    /* take the address of every entry point with its prototype: a signature drift breaks the build */
    int (*f_init)(const bydb_cfg *, bydb_ctx **) = bydb_init;
    void (*f_shutdown)(bydb_ctx *) = bydb_shutdown;
    int (*f_reg)(bydb_ctx *, uint64_t, const bydb_part_files *, bydb_part_h *) = bydb_part_register;
    int (*f_rel)(bydb_ctx *, bydb_part_h) = bydb_part_release;
    int (*f_info)(bydb_ctx *, bydb_part_h, uint64_t *, uint64_t *, uint64_t *) = bydb_part_info;
    int (*f_fb)(bydb_ctx *, bydb_part_h, uint64_t *, uint64_t *) = bydb_part_fallback_pages;
    int (*f_scan)(bydb_ctx *, const bydb_query *, bydb_result *) = bydb_scan_agg;
    int (*f_host)(bydb_ctx *, uint32_t, const bydb_part_files *, const bydb_query *, bydb_result *) = bydb_scan_agg_host;
    void (*f_free)(bydb_ctx *, bydb_result *) = bydb_result_free;
    int (*f_lay)(const bydb_query *, bydb_partials_layout_t *) = bydb_partials_layout;
    int (*f_part)(bydb_ctx *, const bydb_query *, void *, uint64_t, void *, bydb_stats *) = bydb_scan_partials;
    int (*f_comb)(bydb_ctx *, const bydb_query *, void *, uint32_t, uint64_t, void *) = bydb_partials_combine;
    int (*f_fin)(bydb_ctx *, const bydb_query *, const void *, uint64_t, void *, bydb_result *) = bydb_reduce_finalize;
    int (*f_prep)(bydb_ctx *, const bydb_query *, bydb_prepared **) = bydb_query_prepare;
    int (*f_run)(bydb_ctx *, bydb_prepared *, bydb_result *) = bydb_scan_agg_prepared;
    void (*f_qrel)(bydb_ctx *, bydb_prepared *) = bydb_query_release;
    int (*f_dir)(bydb_ctx *, bydb_part_h, void *, uint64_t, void *, uint64_t, uint64_t *, uint64_t *) = bydb_part_directory;
    int (*f_keyed)(bydb_ctx *, const bydb_query *, const bydb_group_key *, bydb_keyed_result *) = bydb_scan_agg_keyed;
    void (*f_kfree)(bydb_ctx *, bydb_keyed_result *) = bydb_keyed_result_free;
    int (*f_prow)(bydb_ctx *, const bydb_query *, const void *, uint64_t, void *, bydb_partial_rows *) = bydb_partials_rows;
    void (*f_prfree)(bydb_ctx *, bydb_partial_rows *) = bydb_partial_rows_free;
    int (*f_cexp)(bydb_ctx *, uint64_t, int32_t, bydb_comm_handle *) = bydb_comm_export;
    int (*f_ccon)(bydb_ctx *, int32_t, int32_t, const bydb_comm_handle *) = bydb_comm_connect;
    int (*f_red)(bydb_ctx *, const bydb_query *, int32_t, bydb_result *) = bydb_scan_reduce;
    int (*f_redp)(bydb_ctx *, bydb_prepared *, int32_t, bydb_result *) = bydb_scan_reduce_prepared;
    int (*f_redh)(bydb_ctx *, uint32_t, const bydb_part_files *, const bydb_query *, int32_t, bydb_result *) = bydb_scan_reduce_host;
    int (*f_enc)(bydb_ctx *, const bydb_encode_input *, bydb_encoded_pages *) = bydb_encode_pages;
    void (*f_encfree)(bydb_ctx *, bydb_encoded_pages *) = bydb_encoded_pages_free;
    (void)f_enc; (void)f_encfree;
    (void)f_dir; (void)f_prow; (void)f_prfree; (void)f_cexp; (void)f_ccon; (void)f_red; (void)f_redp; (void)f_redh;
    (void)f_prep; (void)f_run; (void)f_qrel;
    (void)f_shutdown; (void)f_reg; (void)f_rel; (void)f_info; (void)f_fb; (void)f_scan; (void)f_host; (void)f_free; (void)f_part; (void)f_comb; (void)f_fin;

    printf("version %s\n", bydb_version());
    /* host-only: the partial-table layout of a query with 3 groups and 2 distinct fields */
    bydb_agg aggs[3];
    memset(aggs, 0, sizeof aggs);
    aggs[0].field = "latency"; aggs[0].func = BYDB_AGG_MEAN;
    aggs[1].field = "calls";   aggs[1].func = BYDB_AGG_MIN;
    aggs[2].field = "latency"; aggs[2].func = BYDB_AGG_MAX;
    uint64_t sids[4] = {1, 2, 3, 4};
    int32_t grp[4] = {0, 1, 2, 0};
    bydb_query q;
    memset(&q, 0, sizeof q);
    q.series_ids = sids; q.n_series = 4; q.series_group = grp; q.n_groups = 3; q.aggs = aggs; q.n_aggs = 3;
    q.tmin = INT64_MIN; q.tmax = INT64_MAX;
    bydb_partials_layout_t lay;
    if (f_lay(&q, &lay) != 0) { printf("layout failed: %s\n", bydb_last_error()); return 1; }
    if (lay.n_sum_f64 != 6 || lay.n_max_f64 != 12 || lay.n_sum_i64 != 15 || lay.n_max_i64 != 14 || lay.total_bytes != 8 * (6 + 12 + 15 + 14)) {
        printf("unexpected layout %llu %llu %llu %llu total %llu\n", (unsigned long long)lay.n_sum_f64, (unsigned long long)lay.n_max_f64,
               (unsigned long long)lay.n_sum_i64, (unsigned long long)lay.n_max_i64, (unsigned long long)lay.total_bytes);
        return 1;
    }
    /* host-only: the synthetic part writer */
    bydb_synth_field fld = {"latency", BYDB_SYN_F_LATENCY, 0};
    bydb_synth_spec sp;
    memset(&sp, 0, sizeof sp);
    sp.n_series = 3; sp.n_points = 50; sp.sid0 = 1; sp.sid_step = 1; sp.t0 = 1700000000000000000LL; sp.t_step = 60000000000LL;
    sp.n_fields = 1; sp.fields = &fld; sp.seed = 7;
    bydb_part_image *img = NULL;
    if (bydb_synth_part(&sp, &img) != 0 || !img) { printf("synth failed\n"); return 1; }
    uint64_t rows = 0, blocks = 0;
    bydb_part_image_counts(img, &rows, &blocks);
    if (rows != 150 || blocks != 3 || bydb_part_image_n_files(img) < 4) { printf("synth counts %llu %llu\n", (unsigned long long)rows, (unsigned long long)blocks); return 1; }
    bydb_part_image_free(img);
    /* device: must be refused without a GPU, never emulated */
    bydb_ctx *ctx = NULL;
    int rc = f_init(NULL, &ctx);
    if (rc == 0) {
        printf("init ok (GPU present)\n");
        /* group-by on a stored tag from plain C: count(latency) per value of default/region over 3 series x 2000 points */
        sp.n_points = 2000; sp.region_values = 4; sp.region_run = 8;
        if (bydb_synth_part(&sp, &img) != 0 || !img) { printf("synth failed\n"); return 1; }
        bydb_file files[16];
        uint32_t nf = bydb_part_image_n_files(img);
        if (nf > 16) return 1;
        for (uint32_t i = 0; i < nf; ++i) {
            files[i].name = bydb_part_image_file_name(img, i);
            files[i].data = bydb_part_image_file_data(img, i, &files[i].len);
        }
        bydb_part_files pf = {nf, files};
        bydb_part_h h = 0;
        if (f_reg(ctx, 42, &pf, &h) != 0) { printf("register failed: %s\n", bydb_last_error()); return 1; }
        bydb_agg cnt = {"latency", BYDB_AGG_COUNT, 0};
        uint64_t s3[3] = {1, 2, 3};
        bydb_query kq;
        memset(&kq, 0, sizeof kq);
        kq.parts = &h; kq.n_parts = 1; kq.series_ids = s3; kq.n_series = 3; kq.aggs = &cnt; kq.n_aggs = 1;
        kq.tmin = INT64_MIN; kq.tmax = INT64_MAX;
        bydb_group_key gk = {"default", "region", 0, 0};
        bydb_keyed_result kr;
        if (f_keyed(ctx, &kq, &gk, &kr) != 0) { printf("keyed scan failed: %s\n", bydb_last_error()); return 1; }
        int64_t total = 0;
        for (int32_t r = 0; r < kr.base.n_rows; ++r) {
            const uint32_t a = kr.key_off[kr.key_id[r]], b = kr.key_off[kr.key_id[r] + 1];
            if (b - a != 2 || kr.key_bytes[a] != 'r' || kr.base.group_id[r] != 0) { printf("bad key row %d\n", r); return 1; }
            total += kr.base.val_i64[r];
        }
        if (kr.base.n_rows != 4 || kr.n_keys != 4 || total != 6000) { printf("keyed result: %d rows, %d keys, %lld\n", kr.base.n_rows, kr.n_keys, (long long)total); return 1; }
        f_kfree(ctx, &kr);
        f_rel(ctx, h);
        bydb_part_image_free(img);
        bydb_shutdown(ctx);
    } else {
        printf("init refused: %d %s\n", rc, bydb_last_error());
        if (ctx != NULL) return 1;
    }
    printf("OK\n");
    return 0;
}
