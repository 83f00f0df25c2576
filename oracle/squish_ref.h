/* ORACLE — test infrastructure only. Never linked into the product library.
 *
 * CPU restatement of the squishRS v1.2.0 pack / unpack / list data path:
 *   chunk rule        reference src/archive/writer.rs:240-246 (fixed 2 MiB reads)
 *   digest            src/util/chunk.rs:46-49        (xxh3_ref.c)
 *   dedup + encode    src/util/chunk.rs:80-100       (set of digests; zstd level 12 under the shard lock)
 *   chunk record      src/fsutil/writer.rs:17-39     (hash | orig u64 | comp u64 | frame)
 *   archive prefix    src/archive/writer.rs:66-107, src/util/header.rs:10-63,192-233
 *   manifest          src/archive/writer.rs:292-329
 *   index scan        src/archive/reader.rs:46-118
 *   decode            src/archive/reader.rs:259-314  (serial zstd decompress into a digest-keyed map)
 *   rebuild           src/archive/reader.rs:316-413  (per-file concat by digest, MissingChunk)
 *   list              src/archive/reader.rs:155-219
 *
 * zstd itself is the third-party C library libzstd (reference pins 1.5.7 via
 * zstd-sys 2.0.15, Cargo.lock:1637-1638).  Its source is not under
 * /root/reference; this oracle calls the image's /lib/x86_64-linux-gnu/libzstd.so.1
 * (1.5.5) through dlopen — same frozen format, compressed bytes/ratio may differ
 * slightly from 1.5.7.  The reference cannot be compiled here (no Rust toolchain),
 * so there is no oracle/_ref; parity for digests is pinned by the XXH3 KAT table
 * and libxxhash, the format by a byte-for-byte rebuild of the reference's own
 * hand-built archive fixture (src/archive/tests.rs:14-58) in tests/.
 */
#ifndef SQ_ORACLE_SQUISH_REF_H
#define SQ_ORACLE_SQUISH_REF_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define SQO_CHUNK_SIZE (2048u * 1024u) /* chunk.rs:11 */
#define SQO_LEVEL 12                   /* chunk.rs:12 */

/* error codes mirror AppError kinds (src/util/errors.rs:5-66) */
enum {
    SQO_OK = 0, SQO_ERR_IO = -1, SQO_ERR_WRITER = -4, SQO_ERR_READER = -5, SQO_ERR_COMPRESSION = -7,
    SQO_ERR_ARCHIVE = -8, SQO_ERR_FILE_NOT_EXIST = -14, SQO_ERR_ILLEGAL_UTF8 = -15,
    SQO_ERR_MISSING_CHUNK = -16, SQO_ERR_INVALID_CHUNK_SIZE = -17, SQO_ERR_OTHER = -20
};

/* libzstd via dlopen */
int sqo_zstd_available(void);
unsigned sqo_zstd_version(void);
size_t sqo_zstd_bound(size_t n);
/* returns compressed size, or 0 on error */
size_t sqo_zstd_compress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, int level);
/* returns decoded size, or (size_t)-1 on error (capacity too small, garbage, ...) */
size_t sqo_zstd_decompress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap);

/* An input file: either in memory (data != NULL) or on disk (path_on_disk). */
typedef struct {
    const char *rel_path;      /* manifest path (writer.rs:230-231) */
    const char *path_on_disk;  /* used when data == NULL */
    const uint8_t *data;
    uint64_t size;             /* used when data != NULL */
} sqo_file;

typedef struct {
    uint64_t archive_size, unique_chunks, total_chunks, total_input_bytes, payload_bytes;
    double seconds;
} sqo_pack_stats;

/* ArchiveWriter::new + pack (writer.rs:66-195).  deterministic != 0 writes chunk
 * records in ascending global chunk index instead of arrival order (the
 * reference's order is nondeterministic). */
int sqo_pack(const sqo_file *files, uint32_t nfiles, const char *out_path, int threads,
             uint64_t timestamp, int deterministic, sqo_pack_stats *stats);
/* walk_dir (src/fsutil/directory.rs:39-73) + pack */
int sqo_pack_dir(const char *input_dir, const char *out_path, int threads, sqo_pack_stats *stats);

/* Digest-only mode: per-chunk digests in (file, chunk) order + is_new flag under
 * the lowest-global-index winner rule.  digests: 16 B per chunk slot. */
int sqo_digest_map(const sqo_file *files, uint32_t nfiles, uint8_t *digests, uint8_t *is_new,
                   uint64_t max_chunks, uint64_t *n_chunks, uint64_t *n_unique);

typedef struct {
    uint64_t unique_chunks, total_original_size, archive_size, timestamp;
    double compression_ratio; /* archive/original*100, reader.rs:204-208 */
    uint32_t file_count;
    char version[16];
    double decode_seconds, rebuild_seconds;
} sqo_summary;

/* ArchiveReader::new + get_summary.  If paths_out != NULL it receives a malloc'd
 * '\n'-joined "size path" listing the caller frees with sqo_free. */
int sqo_list(const char *archive_path, sqo_summary *out, char **paths_out);
/* ArchiveReader::new + unpack.  parallel_decode == 0 mirrors reader.rs:276-311
 * (single-threaded decode); != 0 decodes with `threads` workers (NOT reference behaviour). */
int sqo_unpack(const char *archive_path, const char *out_dir, int threads, int parallel_decode,
               sqo_summary *out);
void sqo_free(void *p);
const char *sqo_strerror(int code);

#ifdef __cplusplus
}
#endif
#endif
