// Drives the C++ host-side mirror of the reference's operator surface (include/bydb_operator.hpp) the way the reference's own
// operator tests do (pkg/query/vectorized/measure/aggregation_test.go, limit_test.go): schema typing, the
// (batch,nil)/(nil,nil)/(nil,err) contract with a sticky error, idempotent Close -- and, when a GPU is present, a grouped
// aggregation over a synthetic part whose rows are checked against a direct call of the C ABI.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "bydb_operator.hpp"
#include "bydb_synth.h"

using namespace bydb::vectorized;

static int fails = 0;
#define CHECK(cond, what)                                   \
    do {                                                    \
        if (!(cond)) {                                      \
            std::printf("FAIL %s (%s:%d)\n", what, __FILE__, __LINE__); \
            ++fails;                                        \
        }                                                   \
    } while (0)

static BatchSchema input_schema() {
    BatchSchema s;
    s.Columns.push_back({"service_id", ColumnRole::RoleTag, ColumnType::ColumnTypeString, "default"});
    s.Columns.push_back({"latency", ColumnRole::RoleField, ColumnType::ColumnTypeFloat64, ""});
    return s;
}

static std::vector<std::pair<std::string, std::vector<double>>> drain(GPUScanAgg &op, int max_batch, Status *err_out) {
    std::vector<std::pair<std::string, std::vector<double>>> rows;
    for (;;) {
        std::unique_ptr<RecordBatch> b;
        Status s = op.NextBatch(b);
        if (s) {
            *err_out = s;
            break;
        }
        if (!b) break;  // EOF
        CHECK(b->Len > 0 && b->Len <= max_batch, "batch length within (0, batch_size]");
        for (int i = 0; i < b->Len; ++i) {
            std::vector<double> vals;
            for (size_t c = 1; c < b->Columns.size(); ++c) {
                const Column &col = b->Columns[c];
                vals.push_back(col.Type == ColumnType::ColumnTypeFloat64 ? col.Float64[static_cast<size_t>(i)] : static_cast<double>(col.Int64[static_cast<size_t>(i)]));
            }
            rows.emplace_back(b->Columns[0].Bytes[static_cast<size_t>(i)], vals);
        }
    }
    return rows;
}

int main() {
    // ---- schema typing (aggOutputType): COUNT is int64, everything else follows the field
    {
        GPUScanAgg op(nullptr, input_schema(), {0}, {{"s", AggFunc::AggSum, 1}, {"n", AggFunc::AggCount, 1}, {"m", AggFunc::AggMean, 1}}, ScanSpec{});
        const BatchSchema &o = op.OutputSchema();
        CHECK(o.Columns.size() == 4 && o.Columns[0].Name == "service_id" && o.Columns[0].Role == ColumnRole::RoleTag, "tag column first");
        CHECK(o.Columns[1].Type == ColumnType::ColumnTypeFloat64 && o.Columns[2].Type == ColumnType::ColumnTypeInt64 && o.Columns[3].Type == ColumnType::ColumnTypeFloat64,
              "agg output types");
        CHECK(!op.Init(), "Init succeeds");
        std::unique_ptr<RecordBatch> b;
        Status e1 = op.NextBatch(b);
        CHECK(e1 && !b && e1->Code == BYDB_EINVAL, "no context -> (nil, err)");
        Status e2 = op.NextBatch(b);
        CHECK(e2 && e2->Msg == e1->Msg, "the error is sticky");
        CHECK(!op.Close() && !op.Close(), "Close is idempotent");
    }
    {
        GPUScanAgg bad(nullptr, input_schema(), {0}, {{"x", AggFunc::AggSum, 0}}, ScanSpec{});  // aggregating a tag column
        CHECK(bad.Init().has_value(), "AggSpec on a tag column is refused at Init");
    }
    // ---- with a GPU: grouped aggregation against the C ABI called directly
    bydb_ctx *ctx = nullptr;
    if (bydb_init(nullptr, &ctx) != 0) {
        std::printf("%s (no GPU: contract checks only): %s\n", fails ? "FAILED" : "OK host-only", bydb_last_error());
        return fails ? 1 : 0;
    }
    bydb_synth_field fld = {"latency", BYDB_SYN_F_LATENCY, 0};
    bydb_synth_spec sp{};
    sp.n_series = 12;
    sp.n_points = 400;
    sp.sid0 = 10;
    sp.sid_step = 3;
    sp.t0 = 1700000000000000000LL;
    sp.t_step = 60000000000LL;
    sp.n_fields = 1;
    sp.fields = &fld;
    sp.seed = 99;
    bydb_part_image *img = nullptr;
    CHECK(bydb_synth_part(&sp, &img) == 0, "synth part");
    std::vector<bydb_file> files(bydb_part_image_n_files(img));
    for (uint32_t i = 0; i < files.size(); ++i) {
        files[i].name = bydb_part_image_file_name(img, i);
        files[i].data = bydb_part_image_file_data(img, i, &files[i].len);
    }
    bydb_part_files pf{static_cast<uint32_t>(files.size()), files.data()};
    bydb_part_h h = 0;
    CHECK(bydb_part_register(ctx, 1, &pf, &h) == 0, "register");
    ScanSpec scan;
    scan.Parts = {h};
    std::vector<std::string> svc;
    for (int i = 11; i >= 0; --i) {  // index order is not ascending
        scan.SeriesIDs.push_back(10 + 3 * static_cast<uint64_t>(i));
        svc.push_back("svc_" + std::to_string(i % 4));
    }
    scan.SeriesTags[{"default", "service_id"}] = svc;
    std::vector<AggSpec> aggs = {{"sum_v", AggFunc::AggSum, 1}, {"n", AggFunc::AggCount, 1}, {"mean_v", AggFunc::AggMean, 1}};
    {
        GPUScanAgg op(ctx, input_schema(), {0}, aggs, scan, 3);
        CHECK(!op.Init(), "Init");
        Status err;
        auto rows = drain(op, 3, &err);
        CHECK(!err, "no error");
        CHECK(rows.size() == 4, "four services");
        const char *first_seen[4] = {"svc_3", "svc_2", "svc_1", "svc_0"};  // series 11,10,9,8 come first
        double total_n = 0;
        for (size_t r = 0; r < rows.size(); ++r) {
            CHECK(rows[r].first == first_seen[r], "group order = first appearance in scan order");
            if (rows[r].second[1] != 3 * 400.0 || rows[r].second.size() != 3)
                std::printf("row %zu (%s): %zu values: sum %.17g count %.17g mean %.17g\n", r, rows[r].first.c_str(), rows[r].second.size(), rows[r].second[0],
                            rows[r].second.size() > 1 ? rows[r].second[1] : -1.0, rows[r].second.size() > 2 ? rows[r].second[2] : -1.0);
            CHECK(rows[r].second[1] == 3 * 400.0, "count per service");
            const double mean = rows[r].second[0] / rows[r].second[1];
            CHECK(rows[r].second[2] == (mean < 1 ? 1.0 : mean), "MEAN = sum/count with the <1 -> 1 rule");
            total_n += rows[r].second[1];
        }
        CHECK(total_n == 12 * 400.0, "every datapoint counted once");
        std::unique_ptr<RecordBatch> b;
        CHECK(!op.NextBatch(b) && !b, "EOF stays EOF");
        CHECK(!op.Close() && !op.Close(), "Close idempotent");
        // windows and directions over the same stream (limit_test.go:52-129, order-by DESC)
        GPUScanAgg lim(ctx, input_schema(), {0}, aggs, scan, 8, std::nullopt, LimitSpec{1, 2});
        (void)lim.Init();
        auto lr = drain(lim, 8, &err);
        CHECK(lr.size() == 2 && lr[0].first == "svc_2" && lr[1].first == "svc_1", "offset 1, limit 2");
        ScanSpec rev = scan;
        rev.OrderDesc = true;
        GPUScanAgg desc(ctx, input_schema(), {0}, aggs, rev, 8);
        (void)desc.Init();
        auto dr = drain(desc, 8, &err);
        CHECK(dr.size() == 4 && dr[0].first == "svc_0" && dr[3].first == "svc_3", "order-by DESC reverses the first-appearance order");
        GPUScanAgg top(ctx, input_schema(), {0}, aggs, scan, 8, TopSpec{2, 0, true});
        (void)top.Init();
        auto tr = drain(top, 8, &err);
        CHECK(tr.size() == 2 && tr[0].second[0] >= tr[1].second[0], "Top 2 by the sum, descending");
    }
    bydb_part_release(ctx, h);
    bydb_part_image_free(img);
    bydb_shutdown(ctx);
    std::printf("%s\n", fails ? "FAILED" : "OK full");
    return fails ? 1 : 0;
}
