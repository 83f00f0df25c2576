/*
 * squish_b200.h — C ABI of the B200-native squishRS pack/unpack data path.
 *
 * This is the boundary a Rust `cuda/` crate (build.rs + nvcc, `extern "C"`)
 * would bind.  Every entry point replaces one seam of the reference
 * (SamB032/squishRS v1.2.0); the reference file:line is cited on each.
 * Batch-oriented: one call covers many chunks, because a per-chunk FFI call
 * makes no sense on a GPU.
 *
 * Conventions
 *  - every call returns int32 status: 0 = SQ_OK, negative = an AppError kind
 *    (reference src/util/errors.rs:5-66); sq_last_error(ctx) has the text.
 *  - no exceptions / aborts cross the boundary.
 *  - the caller owns host buffers; pointers named d_* are device pointers the
 *    caller owns (cudaMalloc / torch); the library owns its own scratch.
 *  - `stream` is a cudaStream_t passed as void*; NULL = the context's stream.
 *    Device entry points are asynchronous on that stream unless stated.
 *  - one submitting thread per context.
 *  - there is NO CPU fallback: without a CUDA device sq_create fails with
 *    SQ_ERR_NO_DEVICE and nothing else can be called.
 */
#ifndef SQUISH_B200_H
#define SQUISH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SQ_CHUNK_SIZE (2048u * 1024u)  /* CHUNK_SIZE, reference src/util/chunk.rs:11 */
#define SQ_COMPRESSION_LEVEL 12        /* COMPRESSION_LEVEL, src/util/chunk.rs:12 (parity target for ratio) */
#define SQ_DIGEST_BYTES 16             /* ChunkHash = [u8;16], src/util/chunk.rs:9 */
#define SQ_FORMAT_VERSION "1.2.0"      /* CARGO_PKG_VERSION, src/lib.rs:17, Cargo.toml:3 */

/* status codes: 0 ok, negatives 1:1 with AppError variants (src/util/errors.rs:5-66) */
enum {
    SQ_OK = 0,
    SQ_ERR_IO = -1,                 /* AppError::Io */
    SQ_ERR_READ_DIR = -2,           /* ReadDirError */
    SQ_ERR_READ_ENTRY = -3,         /* ReadEntryError */
    SQ_ERR_WRITER = -4,             /* WriterError */
    SQ_ERR_READER = -5,             /* ReaderError (also: corrupt zstd frame on decode) */
    SQ_ERR_FLUSH = -6,              /* FlushError */
    SQ_ERR_COMPRESSION = -7,        /* Compression */
    SQ_ERR_ARCHIVE = -8,            /* Archive(String) */
    SQ_ERR_ENCODER = -9,            /* EncoderError */
    SQ_ERR_LOCK_POISONED = -10,     /* LockPoisoned */
    SQ_ERR_SENDER = -11,            /* SenderError */
    SQ_ERR_CREATE_DIR = -12,        /* CreateDirError */
    SQ_ERR_CREATE_FILE = -13,       /* CreateFileError */
    SQ_ERR_FILE_NOT_EXIST = -14,    /* FileNotExist */
    SQ_ERR_ILLEGAL_UTF8 = -15,      /* IllegalUTF8 */
    SQ_ERR_MISSING_CHUNK = -16,     /* MissingChunk */
    SQ_ERR_INVALID_CHUNK_SIZE = -17,/* InvalidChunkSize */
    SQ_ERR_CAP_THREADS = -18,       /* CapThreadsError */
    SQ_ERR_INVALID_TIMESTAMP = -19, /* InvalidTimeStamp */
    SQ_ERR_OTHER = -20,             /* Other(String) */
    /* not in the reference: device-side failures */
    SQ_ERR_NO_DEVICE = -100,        /* no CUDA device / driver: there is no CPU fallback */
    SQ_ERR_CUDA = -101,             /* a CUDA runtime call failed */
    SQ_ERR_INVALID_ARG = -102,
    SQ_ERR_CAPACITY = -103          /* caller-provided output buffer too small */
};

typedef struct sq_ctx sq_ctx;

typedef struct {
    int32_t device;             /* CUDA device ordinal */
    uint32_t chunk_size;        /* 0 = SQ_CHUNK_SIZE; must be <= SQ_CHUNK_SIZE */
    uint64_t dedup_capacity;    /* max distinct digests the context will ever hold (0 = 1<<20); must be < 2^30 */
    uint32_t max_batch_chunks;  /* largest n passed to any batch call (0 = 4096) */
    uint32_t flags;             /* SQ_FLAG_* bits, 0 = defaults */
} sq_config;

/* Round 1: "look up every position in the match search" (the default looked up every second one).  The search has looked
 * up every position by default since round 2 (that is what keeps the frames within 3 % of zstd level 12 on real files);
 * the flag is still accepted and changes nothing. */
#define SQ_FLAG_DENSE_SEARCH 1u
/* Record CUDA events between the encoder's kernels so sq_encode_stage_ms can report per-kernel durations (bench.py's
 * roofline line).  Costs four event records per encode call. */
#define SQ_FLAG_STAGE_TIMING 2u
/* Reproducible frames: the same chunk always compresses to the same bytes.  By default the match finder lets the entries of one
 * hash row that fall into the same 512-position tile land in the order of their shared-memory atomics, which can move a few
 * candidates in or out of a search window from run to run (a few hundred bytes in a gigabyte differ; every frame is valid either
 * way).  With this flag the index kernel orders them exactly; a pack then takes about twice as long. */
#define SQ_FLAG_DETERMINISTIC 4u

/* One chunk of a batch: bytes [off, off+len) of the batch buffer.
 * Chunk rule (reference src/archive/writer.rs:240-246): chunk i of a file is
 * bytes [i*2MiB, min((i+1)*2MiB, size)); empty file => no chunk; len in 1..=2MiB. */
typedef struct {
    uint64_t off;
    uint32_t len;
    uint32_t reserved;
} sq_span;

/* Per-chunk result of a pack call == InsertReturn (src/util/chunk.rs:14-17)
 * flattened: hash + Option<frame>.  is_new==1  <=> compressed_data is Some. */
typedef struct {
    uint8_t digest[SQ_DIGEST_BYTES]; /* low64 LE || high64 LE == u128::to_le_bytes (chunk.rs:48) */
    uint64_t frame_off;              /* offset of the zstd frame in the output buffer (is_new only) */
    uint32_t frame_len;              /* bytes (is_new only) */
    uint8_t is_new;
    uint8_t reserved[3];
} sq_chunk_result;

/* One chunk record payload to decode == one (compressed_data, orig_size) pair of
 * read_chunks (src/archive/reader.rs:276-305). */
typedef struct {
    uint64_t src_off;   /* offset of the payload in the compressed buffer */
    uint64_t dst_off;   /* where to write the decoded bytes in the output buffer */
    uint32_t src_len;   /* comp_size */
    uint32_t capacity;  /* orig_size field of the record = decode capacity (reader.rs:286-303) */
} sq_frame;

typedef struct {
    uint32_t out_len;   /* decoded bytes */
    int32_t status;     /* SQ_OK or SQ_ERR_READER (corrupt / exceeds capacity / trailing garbage) */
} sq_frame_result;

/* ---- context ------------------------------------------------------------ */
int32_t sq_create(const sq_config *cfg, sq_ctx **out);
void sq_destroy(sq_ctx *ctx);
const char *sq_last_error(const sq_ctx *ctx);   /* ctx may be NULL: last create error */
const char *sq_strerror(int32_t status);        /* AppError display text (errors.rs:7-65) */
int32_t sq_abi_version(void);
int32_t sq_synchronize(sq_ctx *ctx, void *stream);
/* Returns the device memory the context caches (encoder / decoder scratch, pipeline staging) to the driver; it is allocated
 * again on demand.  Synchronizes the device; fails while a pipeline ticket is in flight. */
int32_t sq_release_scratch(sq_ctx *ctx);
/* number of CUDA kernels this context has launched so far (bench.py's gpu_launches) */
int32_t sq_kernel_launches(sq_ctx *ctx, uint64_t *out);

/* pinned staging that host reader threads fill concurrently (north_star: "pinned,
 * double-buffered cudaMemcpyAsync uploads"); replaces `vec![0u8; CHUNK_SIZE]` (writer.rs:240) */
int32_t sq_host_alloc(sq_ctx *ctx, size_t bytes, void **out);
int32_t sq_host_free(sq_ctx *ctx, void *p);

/* ---- K1: digest  == hash_chunk (src/util/chunk.rs:46-49) ------------------ */
/* d_digests: n*16 bytes.  d_spans: n spans in device memory. */
int32_t sq_digest_device(sq_ctx *ctx, const void *d_data, const sq_span *d_spans, uint32_t n,
                         void *d_digests, void *stream);
/* host convenience (H2D + kernel + D2H, synchronous). */
int32_t sq_digest_host(sq_ctx *ctx, const void *h_data, size_t data_len, const sq_span *h_spans,
                       uint32_t n, uint8_t *h_digests);

/* ---- K2: dedup index  == ChunkStore set (src/util/chunk.rs:19-25,83-99) ---- */
/* Inserts n digests with global chunk indices d_gidx[i] (or gidx_base+i when
 * d_gidx == NULL).  d_is_new[i] = 1 iff i carries the LOWEST global chunk index
 * seen so far for its digest (deterministic stand-in for "first inserter wins";
 * Occupied -> None / Vacant -> Some).  Digest equality is identity (chunk.rs:84-87). */
int32_t sq_dedup_insert_device(sq_ctx *ctx, const void *d_digests, const uint64_t *d_gidx,
                               uint64_t gidx_base, uint32_t n, uint8_t *d_is_new, void *stream);
int32_t sq_dedup_len(sq_ctx *ctx, uint64_t *out);   /* ChunkStore::len (chunk.rs:116-118); synchronizes */
int32_t sq_dedup_reset(sq_ctx *ctx);                /* ChunkStore::new (chunk.rs:52-56) */

/* ---- K2 across the GPUs of one box: the index is sharded by digest prefix (north_star (2)) ----
 * One process per GPU.  Per batch: K1 on local chunks -> sq_route_digests_device buckets
 * {digest, gidx} records by owner = LE64(digest[0..8]) % world into `world` blocks of
 * cap_per_peer 32-byte records (padding records carry gidx = ~0) -> the caller exchanges the
 * blocks with an all-to-all over NCCL -> sq_dedup_insert_routed_device inserts what this rank
 * owns and writes one verdict byte per received record (same layout) -> reverse all-to-all ->
 * sq_unroute_verdicts_device scatters verdicts back to chunk order.  sq_dedup_len on each rank
 * counts the digests that rank owns; the store's len() is their sum.
 * Constraint: count = world * cap_per_peer of a routed insert must not exceed the context's max_batch_chunks (the index
 * keeps one slot reference per record of a batch): create the context with max_batch_chunks >= world * cap_per_peer. */
int32_t sq_route_digests_device(sq_ctx *ctx, const void *d_digests, uint64_t gidx_base, uint32_t n,
                                uint32_t world, uint32_t cap_per_peer, void *d_send,
                                uint32_t *d_send_pos, void *stream);
int32_t sq_dedup_insert_routed_device(sq_ctx *ctx, const void *d_recv, uint32_t count,
                                      uint8_t *d_verdict, void *stream);
int32_t sq_unroute_verdicts_device(sq_ctx *ctx, const uint8_t *d_verdict_back, const uint32_t *d_send_pos,
                                   uint32_t n, uint8_t *d_is_new, void *stream);

/* Buffer slack for the *_device entry points: the kernels read whole aligned words around the bytes they are given.  d_data
 * (encode / digest) and d_comp (decode) must be readable for 16 bytes past the last span / payload and d_comp for 8 bytes
 * before the first payload that does not start at offset 0; a buffer from cudaMalloc with 64 spare bytes at the end, payloads
 * at 16-byte aligned offsets, always satisfies this (what the *_host entry points do).  d_out of a decode: frame outputs at 16-byte
 * aligned dst_off; the decoder reads back whole aligned 32-bit words of what it has written, so the buffer must be readable up to
 * the next multiple of 4 after the last frame's capacity. */

/* ---- K3: encode  == zstd::bulk::compress(chunk, 12) (src/util/chunk.rs:89-90) */
/* worst-case frame bytes for a chunk of `len` bytes (raw-block fallback) */
size_t sq_encode_bound(size_t len);
/* Encodes every chunk i with d_select == NULL || d_select[i] != 0 as ONE standard
 * zstd frame (single segment, FCS present, no checksum, no dictID).  Frames are
 * packed back to back into d_out in ascending i; d_frame_off[i]/d_frame_len[i]
 * receive their placement (len 0 for unselected).  *d_total (device u64) = bytes used.
 * Fails with SQ_ERR_CAPACITY (reported at the next sync point via sq_encode_status)
 * if out_capacity is too small.  Calls on two different streams use two independent
 * scratch sets and may overlap on the device; a further stream takes over the least
 * recently used set and is ordered behind that set's previous call. */
int32_t sq_encode_device(sq_ctx *ctx, const void *d_data, const sq_span *d_spans,
                         const uint8_t *d_select, uint32_t n, void *d_out, uint64_t out_capacity,
                         uint64_t *d_frame_off, uint32_t *d_frame_len, uint64_t *d_total, void *stream);

/* Synchronizes and reports whether the most recent sq_encode_device on this context
 * overflowed d_out (SQ_ERR_CAPACITY) -- the Err arm of compress() (chunk.rs:90). */
int32_t sq_encode_status(sq_ctx *ctx);
/* Durations in ms of the last sq_encode_device call on `stream` (NULL = the context's stream): out[0] match search,
 * out[1] parse (chase), out[2] entropy coding, out[3] frame sizing + placement + emission.  Needs SQ_FLAG_STAGE_TIMING;
 * synchronizes on that call's last event. */
int32_t sq_encode_stage_ms(sq_ctx *ctx, void *stream, float out[4]);

/* ---- K4: decode  == zstd::bulk::decompress(bytes, orig_size) (src/archive/reader.rs:302-303) */
/* Accepts everything stock ZSTD_decompress accepts whole: >=1 concatenated frames,
 * skippable frames, frames without content size.  Errors per frame, not per call. */
int32_t sq_decode_device(sq_ctx *ctx, const void *d_comp, const sq_frame *d_frames, uint32_t n,
                         void *d_out, sq_frame_result *d_results, void *stream);

/* ---- fused hot path: ChunkStore::insert for a whole batch (chunk.rs:80-100) ---- */
/* digest -> dedup -> encode winners.  d_results[n]; frames packed into d_out.
 * out_used (host, may be NULL) is written after an internal synchronize. */
int32_t sq_pack_device(sq_ctx *ctx, const void *d_data, const sq_span *d_spans, uint32_t n,
                       uint64_t gidx_base, sq_chunk_result *d_results, void *d_out,
                       uint64_t out_capacity, uint64_t *out_used, void *stream);
/* Same through HOST buffers: H2D of the batch, kernels, D2H of results + frames.
 * h_data should come from sq_host_alloc (pinned) for full link speed. */
int32_t sq_pack_host(sq_ctx *ctx, const void *h_data, size_t data_len, const sq_span *h_spans,
                     uint32_t n, uint64_t gidx_base, sq_chunk_result *h_results, void *h_out,
                     uint64_t out_capacity, uint64_t *out_used);
/* Asynchronous twin of sq_pack_host for a double-buffered host pipeline (north_star: "pinned,
 * double-buffered cudaMemcpyAsync uploads"): submit returns as soon as the upload, kernels and
 * result download are queued; at most two tickets may be in flight.  sq_pack_wait blocks for
 * one ticket and downloads exactly the frame bytes it produced.  h_data / h_results / h_out
 * must stay valid (and should be pinned) until the ticket is waited on. */
typedef struct sq_ticket sq_ticket;
int32_t sq_pack_submit(sq_ctx *ctx, const void *h_data, size_t data_len, const sq_span *h_spans,
                       uint32_t n, uint64_t gidx_base, sq_chunk_result *h_results, void *h_out,
                       uint64_t out_capacity, sq_ticket **ticket);
int32_t sq_pack_wait(sq_ctx *ctx, sq_ticket *ticket, uint64_t *out_used);
/* Asynchronous twin of sq_unpack_host (same arguments): at most two tickets in flight; the upload of
 * batch k+1 and the download of batch k-1 overlap the decode of batch k.  h_comp / h_frames may be
 * reused once submit returns only if they are pageable; pinned buffers must stay untouched until the
 * ticket is waited on, like h_out / h_results. */
int32_t sq_unpack_submit(sq_ctx *ctx, const void *h_comp, size_t comp_len, const sq_frame *h_frames,
                         uint32_t n, void *h_out, size_t out_len, sq_frame_result *h_results,
                         sq_ticket **ticket);
int32_t sq_unpack_wait(sq_ctx *ctx, sq_ticket *ticket);
/* read_chunks for a batch through HOST buffers: H2D payloads, decode, D2H output. */
int32_t sq_unpack_host(sq_ctx *ctx, const void *h_comp, size_t comp_len, const sq_frame *h_frames,
                       uint32_t n, void *h_out, size_t out_len, sq_frame_result *h_results);

/* ---- archive level: the packer / unpacker objects ----------------------------
 * ArchiveWriter::new + pack (src/archive/writer.rs:66-195), ArchiveReader::new +
 * unpack / get_summary (src/archive/reader.rs:46-244).  Host C++ in this library;
 * the per-chunk work goes through the kernels above. */
typedef struct {
    uint64_t archive_size, unique_chunks, total_chunks, total_input_bytes, payload_bytes;
    uint32_t file_count;
    double seconds_total, seconds_device;
} sq_pack_report;

typedef struct {
    uint64_t unique_chunks, total_original_size, archive_size, timestamp;
    double compression_ratio;      /* archive/original*100 (reader.rs:204-208) */
    uint32_t file_count;
    char version[16];
    double seconds_total, seconds_device;
} sq_summary;

int32_t sq_archive_pack(sq_ctx *ctx, const char *input_dir, const char *output_path, int32_t threads,
                        sq_pack_report *report);
int32_t sq_archive_unpack(sq_ctx *ctx, const char *archive_path, const char *output_dir, int32_t threads,
                          sq_summary *summary);
/* The same over several GPUs of one box: one context per device (sq_create with cfg.device = ordinal), all owned by the
 * calling process; ctxs[0] reports errors.  n_ctx == 1 is exactly the single-device call.
 *  - unpack: chunk records are independent (reader.rs:276-305), so they are cut into contiguous ranges of equal bytes, one
 *    range per device, no exchange between devices.
 *  - pack: the chunk stream is dealt to the devices batch by batch in file order; digests of one round of batches meet in ONE
 *    dedup index (on ctxs[0]'s device, digests travel device to device), so "first occurrence wins" (chunk.rs:83-99) holds
 *    across devices and the archive has the records a single device would write, in the same order. */
/* The building block of the multi-device pack: from now on `ctx` decides "new or duplicate" in `owner`'s index (the owner shares
 * with itself).  K1 runs where the batch is; 16 digest bytes per chunk travel to the owner's device, K2 runs there on one stream
 * in submission order, one verdict byte per chunk travels back; K3 runs where the batch is.  One host thread submits the batches
 * of all sharing contexts in chunk order.  owner == NULL ends the sharing (the context uses its own index again); the owner must
 * outlive every context that shares with it. */
int32_t sq_share_dedup(sq_ctx *ctx, sq_ctx *owner);
int32_t sq_archive_pack_multi(sq_ctx **ctxs, uint32_t n_ctx, const char *input_dir, const char *output_path,
                              int32_t threads, sq_pack_report *report);
int32_t sq_archive_unpack_multi(sq_ctx **ctxs, uint32_t n_ctx, const char *archive_path, const char *output_dir,
                                int32_t threads, sq_summary *summary);
/* list needs no device: ctx may be NULL.  *listing (optional) receives a malloc'd
 * "size path\n" text to release with sq_free. */
int32_t sq_archive_list(const char *archive_path, sq_summary *summary, char **listing);
void sq_free(void *p);

/* ---- synthetic corpora (bench/test support) -------------- */
/* Fills d_out[0..len) with the bytes of payload `payload_id` of corpus class
 * `klass` (0 text, 1 log lines, 2 JSON, 3 binary records, 4 random, 5 zeros).
 * The same generator is available on the host as sq_corpus_fill_host so oracle
 * and device see identical bytes. */
int32_t sq_corpus_fill_device(sq_ctx *ctx, void *d_out, uint64_t len, uint64_t seed, uint64_t payload_id,
                              uint32_t klass, void *stream);
int32_t sq_corpus_fill_host(void *h_out, uint64_t len, uint64_t seed, uint64_t payload_id, uint32_t klass);
/* batched: slot i gets payload ids[i], class klasses[i], at d_out + i*slot_bytes */
int32_t sq_corpus_fill_slots_device(sq_ctx *ctx, void *d_out, uint64_t slot_bytes, const uint64_t *d_ids,
                                    const uint32_t *d_klass, uint32_t n, uint64_t seed, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SQUISH_B200_H */
