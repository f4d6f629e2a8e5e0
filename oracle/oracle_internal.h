/* oracle_internal.h -- ORACLE (test infrastructure) private structures. */
#ifndef ORACLE_INTERNAL_H
#define ORACLE_INTERNAL_H

#include "bydb_oracle.h"

/* banyand/measure/measure.go:41-46 */
#define OBI_MAX_UNCOMPRESSED_BLOCK (2u * 1024u * 1024u)
#define OBI_MAX_UNCOMPRESSED_PRIMARY (128u * 1024u)
#define OBI_MAX_BLOCK_LENGTH (8u * 1024u)

typedef struct {
    char name[160];
    ob_buf data;
} obi_file;

/* primary_metadata.go:30-35 */
typedef struct {
    uint64_t sid;
    int64_t min_ts, max_ts;
    uint64_t offset, size;
} obi_primary;

/* column_metadata.go:30-34 */
typedef struct {
    char name[128];
    int value_type;
    uint64_t offset, size;
} obi_colmeta;

typedef struct {
    char name[128];
    uint64_t offset, size; /* into <name>.tfm */
} obi_fammeta;

/* block_metadata.go:61-69, 238-246 */
typedef struct {
    uint64_t sid;
    uint64_t uncompressed;
    uint64_t count;
    uint64_t ts_off, ts_size;
    int64_t ts_min, ts_max;
    int ts_enc;
    uint64_t ver_off;
    int64_t ver_first;
    int ver_enc;
    size_t n_fams;
    obi_fammeta *fams;
    size_t n_fields;
    obi_colmeta *fields;
} obi_blockmeta;

struct ob_part {
    int n_files;
    obi_file **files;
    size_t n_pbm;
    obi_primary *pbm;
    uint64_t total_count, blocks_count, uncompressed, compressed;
    int64_t min_ts, max_ts;
};

obi_file *obi_part_file(ob_part *p, const char *name, int create);
int obi_part_read_primary_block(ob_part *p, size_t idx, obi_blockmeta **out, size_t *n_out);
const uint8_t *obi_bm_unmarshal(const uint8_t *p, const uint8_t *end, obi_blockmeta *bm);
const uint8_t *obi_cfm_unmarshal(const uint8_t *p, const uint8_t *end, obi_colmeta **out, size_t *n_out);
void obi_bm_free(obi_blockmeta *bm);
void ob_set_error(const char *msg);

#endif
