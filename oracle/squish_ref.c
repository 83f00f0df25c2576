/* ORACLE — test infrastructure only (see squish_ref.h for the reference file:line map). */
#define _GNU_SOURCE
#include "squish_ref.h"
#include "xxh3_ref.h"

#include <dirent.h>
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

static const char VERSION[] = "1.2.0"; /* Cargo.toml:3 -> CARGO_PKG_VERSION, lib.rs:17 */
static const char PREFIX[] = "squish"; /* header.rs:10 */

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}

/* ------------------------------------------------------------------ libzstd */
typedef size_t (*fn_compress)(void *, size_t, const void *, size_t, int);
typedef size_t (*fn_decompress)(void *, size_t, const void *, size_t);
typedef size_t (*fn_bound)(size_t);
typedef unsigned (*fn_iserr)(size_t);
typedef unsigned (*fn_ver)(void);
static struct { void *h; fn_compress c; fn_decompress d; fn_bound b; fn_iserr e; fn_ver v; } Z;
static pthread_once_t z_once = PTHREAD_ONCE_INIT;
static void z_load(void) {
    const char *names[] = { "libzstd.so.1", "/lib/x86_64-linux-gnu/libzstd.so.1", "libzstd.so" };
    for (int i = 0; i < 3 && !Z.h; i++) Z.h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!Z.h) return;
    Z.c = (fn_compress)dlsym(Z.h, "ZSTD_compress");
    Z.d = (fn_decompress)dlsym(Z.h, "ZSTD_decompress");
    Z.b = (fn_bound)dlsym(Z.h, "ZSTD_compressBound");
    Z.e = (fn_iserr)dlsym(Z.h, "ZSTD_isError");
    Z.v = (fn_ver)dlsym(Z.h, "ZSTD_versionNumber");
}
int sqo_zstd_available(void) { pthread_once(&z_once, z_load); return Z.c && Z.d && Z.b && Z.e; }
unsigned sqo_zstd_version(void) { return sqo_zstd_available() && Z.v ? Z.v() : 0; }
size_t sqo_zstd_bound(size_t n) { return sqo_zstd_available() ? Z.b(n) : 0; }
size_t sqo_zstd_compress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap, int level) {
    if (!sqo_zstd_available()) return 0;
    size_t r = Z.c(dst, cap, src, n, level); /* zstd::bulk::compress(chunk, 12), chunk.rs:89-90 */
    return Z.e(r) ? 0 : r;
}
size_t sqo_zstd_decompress(const uint8_t *src, size_t n, uint8_t *dst, size_t cap) {
    if (!sqo_zstd_available()) return (size_t)-1;
    size_t r = Z.d(dst, cap, src, n); /* zstd::bulk::decompress(bytes, orig_size), reader.rs:302-303 */
    return Z.e(r) ? (size_t)-1 : r;
}
void sqo_free(void *p) { free(p); }
const char *sqo_strerror(int code) {
    switch (code) {
    case SQO_OK: return "ok";
    case SQO_ERR_IO: return "I/O error";
    case SQO_ERR_WRITER: return "Error writing to squish";
    case SQO_ERR_READER: return "Error reading from squish";
    case SQO_ERR_COMPRESSION: return "Compression error";
    case SQO_ERR_ARCHIVE: return "Archive format error";
    case SQO_ERR_FILE_NOT_EXIST: return "Specified file does not exist";
    case SQO_ERR_ILLEGAL_UTF8: return "Illegal UTF8 detected";
    case SQO_ERR_MISSING_CHUNK: return "Missing Chunk for File";
    case SQO_ERR_INVALID_CHUNK_SIZE: return "Invalid chunk size";
    default: return "Unknown error";
    }
}

static void put32(uint8_t *p, uint32_t v) { for (int i = 0; i < 4; i++) p[i] = (uint8_t)(v >> (8 * i)); }
static void put64(uint8_t *p, uint64_t v) { for (int i = 0; i < 8; i++) p[i] = (uint8_t)(v >> (8 * i)); }
static uint32_t get32(const uint8_t *p) { uint32_t v = 0; for (int i = 0; i < 4; i++) v |= (uint32_t)p[i] << (8 * i); return v; }
static uint64_t get64(const uint8_t *p) { uint64_t v = 0; for (int i = 0; i < 8; i++) v |= (uint64_t)p[i] << (8 * i); return v; }

/* ------------------------------------------- ChunkStore: a set of digests
 * chunk.rs:14-25,80-100.  DashMap = N shards each behind a lock; the vacant
 * entry holds its shard lock across compress() (chunk.rs:83-98). */
typedef struct set_node { uint8_t h[16]; uint64_t gidx; struct set_node *next; } set_node;
typedef struct { pthread_mutex_t mu; set_node **buckets; size_t nb, count; } set_shard;
typedef struct { set_shard *shards; size_t nshards; } chunk_store;

static void store_init(chunk_store *s) {
    long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
    size_t n = 1;
    while ((long)n < 4 * (ncpu > 0 ? ncpu : 1)) n <<= 1; /* dashmap default shard amount */
    s->nshards = n;
    s->shards = calloc(n, sizeof(set_shard));
    for (size_t i = 0; i < n; i++) {
        pthread_mutex_init(&s->shards[i].mu, NULL);
        s->shards[i].nb = 1024;
        s->shards[i].buckets = calloc(1024, sizeof(set_node *));
    }
}
static void store_free(chunk_store *s) {
    for (size_t i = 0; i < s->nshards; i++) {
        for (size_t b = 0; b < s->shards[i].nb; b++)
            for (set_node *n = s->shards[i].buckets[b]; n;) { set_node *nx = n->next; free(n); n = nx; }
        free(s->shards[i].buckets);
        pthread_mutex_destroy(&s->shards[i].mu);
    }
    free(s->shards);
}
static uint64_t store_len(chunk_store *s) { /* chunk.rs:116-118 */
    uint64_t n = 0;
    for (size_t i = 0; i < s->nshards; i++) n += s->shards[i].count;
    return n;
}

/* ------------------------------------------------------------ writer thread
 * fsutil/writer.rs:11-39: unbounded MPSC queue drained by one thread. */
typedef struct msg { uint8_t h[16]; uint8_t *comp; uint64_t comp_len, orig, gidx; struct msg *next; } msg;
typedef struct {
    pthread_mutex_t mu; pthread_cond_t cv; msg *head, *tail; int closed;
    FILE *f; int err; int hold; /* hold: keep for deterministic ordering */
    msg **held; size_t nheld, capheld; uint64_t payload;
} wqueue;

static int write_record(FILE *f, const msg *m) {
    uint8_t hdr[32];
    memcpy(hdr, m->h, 16);
    put64(hdr + 16, m->orig);
    put64(hdr + 24, m->comp_len);
    if (fwrite(hdr, 1, 32, f) != 32) return -1;
    if (m->comp_len && fwrite(m->comp, 1, m->comp_len, f) != m->comp_len) return -1;
    return 0;
}
static void *writer_main(void *arg) {
    wqueue *q = arg;
    for (;;) {
        pthread_mutex_lock(&q->mu);
        while (!q->head && !q->closed) pthread_cond_wait(&q->cv, &q->mu);
        msg *m = q->head;
        if (m) { q->head = m->next; if (!q->head) q->tail = NULL; }
        pthread_mutex_unlock(&q->mu);
        if (!m) break;
        q->payload += m->comp_len;
        if (q->hold) {
            if (q->nheld == q->capheld) { q->capheld = q->capheld ? q->capheld * 2 : 1024; q->held = realloc(q->held, q->capheld * sizeof(msg *)); }
            q->held[q->nheld++] = m;
            continue;
        }
        if (write_record(q->f, m)) q->err = 1;
        free(m->comp);
        free(m);
    }
    return NULL;
}
static int cmp_msg(const void *a, const void *b) {
    uint64_t x = (*(msg *const *)a)->gidx, y = (*(msg *const *)b)->gidx;
    return x < y ? -1 : x > y;
}

/* ------------------------------------------------------------------- pack */
typedef struct {
    const sqo_file *files; uint32_t nfiles; uint32_t next; pthread_mutex_t mu;
    chunk_store store; wqueue q;
    uint8_t **file_hashes; uint32_t *file_nchunks; uint64_t *file_sizes; uint64_t *file_first_gidx;
    int err; int digest_only; uint8_t *is_new_out;
} pack_ctx;

/* ChunkStore::insert (chunk.rs:80-100) + the send in process_file (writer.rs:251-260) */
static int store_insert(pack_ctx *pc, const uint8_t *chunk, size_t n, uint64_t gidx, uint8_t hash[16]) {
    sqo_hash_chunk(chunk, n, hash);
    uint64_t k0 = get64(hash), k1 = get64(hash + 8);
    set_shard *sh = &pc->store.shards[(k0 ^ (k1 * 0x9E3779B97F4A7C15ULL)) >> 7 & (pc->store.nshards - 1)];
    pthread_mutex_lock(&sh->mu);
    size_t b = (size_t)(k1 ^ k0 >> 32) & (sh->nb - 1);
    for (set_node *nd = sh->buckets[b]; nd; nd = nd->next)
        if (!memcmp(nd->h, hash, 16)) { /* Entry::Occupied -> None: digest equality is identity */
            if (pc->digest_only && gidx < nd->gidx) nd->gidx = gidx;
            pthread_mutex_unlock(&sh->mu);
            return 0;
        }
    msg *m = NULL;
    if (!pc->digest_only) { /* Entry::Vacant: compress while holding the shard lock */
        size_t cap = sqo_zstd_bound(n);
        uint8_t *comp = malloc(cap ? cap : 1);
        size_t cl = sqo_zstd_compress(chunk, n, comp, cap, SQO_LEVEL);
        if (!cl) { free(comp); pthread_mutex_unlock(&sh->mu); return SQO_ERR_COMPRESSION; }
        m = calloc(1, sizeof(msg));
        memcpy(m->h, hash, 16);
        m->comp = comp; m->comp_len = cl; m->gidx = gidx;
        m->orig = SQO_CHUNK_SIZE; /* writer.rs:255: chunk_buf.len(), not bytes_read */
    }
    set_node *nd = malloc(sizeof(set_node));
    memcpy(nd->h, hash, 16);
    nd->gidx = gidx;
    nd->next = sh->buckets[b];
    sh->buckets[b] = nd;
    sh->count++;
    pthread_mutex_unlock(&sh->mu);
    if (m) {
        pthread_mutex_lock(&pc->q.mu);
        if (pc->q.tail) pc->q.tail->next = m; else pc->q.head = m;
        pc->q.tail = m;
        pthread_cond_signal(&pc->q.cv);
        pthread_mutex_unlock(&pc->q.mu);
    }
    return 1;
}

/* ArchiveWriter::process_file (writer.rs:229-271) */
static int process_file(pack_ctx *pc, uint32_t fi, uint8_t *chunk_buf) {
    const sqo_file *f = &pc->files[fi];
    uint64_t size = f->size, gidx = pc->file_first_gidx ? pc->file_first_gidx[fi] : 0;
    int fd = -1;
    if (!f->data) {
        fd = open(f->path_on_disk, O_RDONLY);
        if (fd < 0) return SQO_ERR_IO;
        struct stat st;
        if (fstat(fd, &st)) { close(fd); return SQO_ERR_IO; }
        size = (uint64_t)st.st_size;
    }
    pc->file_sizes[fi] = size;
    uint32_t cap = (uint32_t)((size + SQO_CHUNK_SIZE - 1) / SQO_CHUNK_SIZE) + 1, n = 0;
    uint8_t *hashes = malloc((size_t)cap * 16);
    uint64_t off = 0;
    int rc = 0;
    for (;;) {
        const uint8_t *slice; size_t got;
        if (f->data) {
            got = size - off < SQO_CHUNK_SIZE ? (size_t)(size - off) : SQO_CHUNK_SIZE;
            slice = f->data + off;
        } else { /* one read() of <= 2 MiB; regular files fill until EOF */
            got = 0;
            while (got < SQO_CHUNK_SIZE) {
                ssize_t r = read(fd, chunk_buf + got, SQO_CHUNK_SIZE - got);
                if (r < 0) { rc = SQO_ERR_READER; break; }
                if (r == 0) break;
                got += (size_t)r;
            }
            slice = chunk_buf;
        }
        if (rc || got == 0) break; /* empty file => 0 chunks; exact multiple => no empty tail */
        if (n == cap) { cap *= 2; hashes = realloc(hashes, (size_t)cap * 16); }
        int r = store_insert(pc, slice, got, gidx + n, hashes + (size_t)n * 16);
        if (r < 0) { rc = r; break; }
        n++;
        off += got;
    }
    if (fd >= 0) close(fd);
    pc->file_hashes[fi] = hashes;
    pc->file_nchunks[fi] = n;
    return rc;
}
static void *pack_worker(void *arg) { /* files.par_iter().map(process_file) (writer.rs:153-165) */
    pack_ctx *pc = arg;
    uint8_t *buf = malloc(SQO_CHUNK_SIZE);
    for (;;) {
        pthread_mutex_lock(&pc->mu);
        uint32_t i = pc->next < pc->nfiles ? pc->next++ : UINT32_MAX;
        pthread_mutex_unlock(&pc->mu);
        if (i == UINT32_MAX) break;
        int r = process_file(pc, i, buf);
        if (r < 0) pc->err = r;
    }
    free(buf);
    return NULL;
}

static int run_pack(pack_ctx *pc, int threads) {
    if (threads < 1) threads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, pack_worker, pc);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    return pc->err;
}

int sqo_pack(const sqo_file *files, uint32_t nfiles, const char *out_path, int threads,
             uint64_t timestamp, int deterministic, sqo_pack_stats *stats) {
    if (!sqo_zstd_available()) return SQO_ERR_OTHER;
    double t0 = now_s();
    FILE *f = fopen(out_path, "wb+");
    if (!f) return SQO_ERR_IO;
    setvbuf(f, NULL, _IOFBF, 1 << 20);
    /* ArchiveWriter::new (writer.rs:72-86): magic+version, timestamp, chunk-count placeholder */
    uint8_t pre[27];
    memcpy(pre, PREFIX, 6);
    memcpy(pre + 6, VERSION, 5);
    put64(pre + 11, timestamp);
    put64(pre + 19, 0);
    if (fwrite(pre, 1, 27, f) != 27) { fclose(f); return SQO_ERR_WRITER; }

    pack_ctx pc;
    memset(&pc, 0, sizeof pc);
    pc.files = files; pc.nfiles = nfiles;
    pthread_mutex_init(&pc.mu, NULL);
    store_init(&pc.store);
    pthread_mutex_init(&pc.q.mu, NULL);
    pthread_cond_init(&pc.q.cv, NULL);
    pc.q.f = f; pc.q.hold = deterministic;
    pc.file_hashes = calloc(nfiles ? nfiles : 1, sizeof(uint8_t *));
    pc.file_nchunks = calloc(nfiles ? nfiles : 1, sizeof(uint32_t));
    pc.file_sizes = calloc(nfiles ? nfiles : 1, sizeof(uint64_t));
    pc.file_first_gidx = calloc(nfiles ? nfiles : 1, sizeof(uint64_t));
    for (uint32_t i = 0; i < nfiles; i++) /* gidx is only a sort key; a stride of 2^32 chunks/file keeps order */
        pc.file_first_gidx[i] = (uint64_t)i << 32;
    pthread_t wt;
    pthread_create(&wt, NULL, writer_main, &pc.q);
    int rc = run_pack(&pc, threads);
    pthread_mutex_lock(&pc.q.mu); /* drop(sender); handle.join() (writer.rs:168-174) */
    pc.q.closed = 1;
    pthread_cond_signal(&pc.q.cv);
    pthread_mutex_unlock(&pc.q.mu);
    pthread_join(wt, NULL);
    if (deterministic) {
        qsort(pc.q.held, pc.q.nheld, sizeof(msg *), cmp_msg);
        for (size_t i = 0; i < pc.q.nheld; i++) {
            if (write_record(f, pc.q.held[i])) pc.q.err = 1;
            free(pc.q.held[i]->comp);
            free(pc.q.held[i]);
        }
        free(pc.q.held);
    }
    if (pc.q.err && !rc) rc = SQO_ERR_WRITER;
    uint64_t total_chunks = 0, total_bytes = 0;
    if (!rc) {
        /* patch_u64(pos, chunk_store.len()) (writer.rs:177-184, header.rs:224-233) */
        uint8_t b8[8];
        put64(b8, store_len(&pc.store));
        fseek(f, 19, SEEK_SET);
        fwrite(b8, 1, 8, f);
        fseek(f, 0, SEEK_END);
        /* write_files_metadata (writer.rs:292-329) */
        uint8_t b4[4];
        put32(b4, nfiles);
        fwrite(b4, 1, 4, f);
        for (uint32_t i = 0; i < nfiles; i++) {
            uint32_t pl = (uint32_t)strlen(files[i].rel_path);
            put32(b4, pl); fwrite(b4, 1, 4, f);
            fwrite(files[i].rel_path, 1, pl, f);
            put64(b8, pc.file_sizes[i]); fwrite(b8, 1, 8, f);
            put32(b4, pc.file_nchunks[i]); fwrite(b4, 1, 4, f);
            fwrite(pc.file_hashes[i], 16, pc.file_nchunks[i], f);
            total_chunks += pc.file_nchunks[i];
            total_bytes += pc.file_sizes[i];
        }
        if (fflush(f) || ferror(f)) rc = SQO_ERR_WRITER;
    }
    if (stats) {
        memset(stats, 0, sizeof *stats);
        fseek(f, 0, SEEK_END);
        stats->archive_size = (uint64_t)ftell(f);
        stats->unique_chunks = store_len(&pc.store);
        stats->total_chunks = total_chunks;
        stats->total_input_bytes = total_bytes;
        stats->payload_bytes = pc.q.payload;
        stats->seconds = now_s() - t0;
    }
    fclose(f);
    for (uint32_t i = 0; i < nfiles; i++) free(pc.file_hashes[i]);
    free(pc.file_hashes); free(pc.file_nchunks); free(pc.file_sizes); free(pc.file_first_gidx);
    store_free(&pc.store);
    return rc;
}

int sqo_digest_map(const sqo_file *files, uint32_t nfiles, uint8_t *digests, uint8_t *is_new,
                   uint64_t max_chunks, uint64_t *n_chunks, uint64_t *n_unique) {
    /* serial in (file, chunk) order: the first inserter of a digest is the lowest global chunk index */
    pack_ctx pc;
    memset(&pc, 0, sizeof pc);
    pc.digest_only = 1;
    store_init(&pc.store);
    uint8_t *buf = malloc(SQO_CHUNK_SIZE);
    int rc = 0;
    uint64_t g = 0;
    for (uint32_t i = 0; i < nfiles && !rc; i++) {
        const sqo_file *f = &files[i];
        uint64_t size = f->size, off = 0;
        int fd = -1;
        if (!f->data) {
            fd = open(f->path_on_disk, O_RDONLY);
            if (fd < 0) { rc = SQO_ERR_IO; break; }
        }
        for (;;) {
            const uint8_t *slice; size_t got = 0;
            if (f->data) { got = size - off < SQO_CHUNK_SIZE ? (size_t)(size - off) : SQO_CHUNK_SIZE; slice = f->data + off; }
            else {
                while (got < SQO_CHUNK_SIZE) {
                    ssize_t r = read(fd, buf + got, SQO_CHUNK_SIZE - got);
                    if (r < 0) { rc = SQO_ERR_READER; break; }
                    if (r == 0) break;
                    got += (size_t)r;
                }
                slice = buf;
            }
            if (rc || !got) break;
            if (g >= max_chunks) { rc = SQO_ERR_OTHER; break; }
            int r = store_insert(&pc, slice, got, g, digests + g * 16);
            if (r < 0) { rc = r; break; }
            if (is_new) is_new[g] = (uint8_t)r;
            g++; off += got;
        }
        if (fd >= 0) close(fd);
    }
    free(buf);
    if (!rc) { *n_chunks = g; *n_unique = store_len(&pc.store); }
    store_free(&pc.store);
    return rc;
}

/* -------------------------------------------------- walk_dir (directory.rs:39-73)
 * iterative stack DFS; directories via stat() (follows symlinks);
 * everything else is a file. */
typedef struct { char **v; size_t n, cap; } strvec;
static void sv_push(strvec *s, char *p) {
    if (s->n == s->cap) { s->cap = s->cap ? s->cap * 2 : 64; s->v = realloc(s->v, s->cap * sizeof(char *)); }
    s->v[s->n++] = p;
}
static int walk_dir(const char *root, strvec *files) {
    strvec stack = { 0 };
    sv_push(&stack, strdup(root));
    int rc = 0;
    while (stack.n) {
        char *dir = stack.v[--stack.n];
        DIR *d = opendir(dir);
        if (!d) { free(dir); rc = SQO_ERR_IO; break; }
        struct dirent *e;
        while ((e = readdir(d))) {
            if (!strcmp(e->d_name, ".") || !strcmp(e->d_name, "..")) continue;
            size_t l = strlen(dir) + strlen(e->d_name) + 2;
            char *p = malloc(l);
            snprintf(p, l, "%s/%s", dir, e->d_name);
            struct stat st;
            if (!stat(p, &st) && S_ISDIR(st.st_mode)) sv_push(&stack, p); else sv_push(files, p);
        }
        closedir(d);
        free(dir);
    }
    for (size_t i = 0; i < stack.n; i++) free(stack.v[i]);
    free(stack.v);
    return rc;
}

int sqo_pack_dir(const char *input_dir, const char *out_path, int threads, sqo_pack_stats *stats) {
    size_t il = strlen(input_dir);
    char *root = strdup(input_dir);
    while (il > 1 && (root[il - 1] == '/' || root[il - 1] == '\\')) root[--il] = 0; /* lib.rs:28 */
    struct stat st;
    if (stat(root, &st) || !S_ISDIR(st.st_mode)) { free(root); return SQO_ERR_IO; }
    strvec paths = { 0 };
    int rc = walk_dir(root, &paths);
    sqo_file *files = calloc(paths.n ? paths.n : 1, sizeof(sqo_file));
    for (size_t i = 0; i < paths.n; i++) {
        files[i].path_on_disk = paths.v[i];
        files[i].rel_path = paths.v[i] + il + 1; /* strip_prefix(input_path) (writer.rs:230) */
    }
    if (!rc) rc = sqo_pack(files, (uint32_t)paths.n, out_path, threads, (uint64_t)time(NULL), 0, stats);
    for (size_t i = 0; i < paths.n; i++) free(paths.v[i]);
    free(paths.v); free(files); free(root);
    return rc;
}

/* ------------------------------------------------------------------ reader */
typedef struct {
    uint8_t *buf; uint64_t size;                    /* whole archive in memory */
    uint64_t timestamp, nchunks, chunk_table_off, file_table_off;
    uint32_t file_count; char version[16];
} reader;

static int read_whole(const char *path, uint8_t **out, uint64_t *size) {
    FILE *f = fopen(path, "rb");
    if (!f) return SQO_ERR_FILE_NOT_EXIST; /* reader.rs:47-48 */
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *b = malloc(n > 0 ? (size_t)n : 1);
    if (n > 0 && fread(b, 1, (size_t)n, f) != (size_t)n) { fclose(f); free(b); return SQO_ERR_READER; }
    fclose(f);
    *out = b; *size = (uint64_t)n;
    return 0;
}
/* verify_header (header.rs:119-163) + ArchiveReader::new (reader.rs:46-118) */
static int reader_open(reader *r, const char *path) {
    memset(r, 0, sizeof *r);
    int rc = read_whole(path, &r->buf, &r->size);
    if (rc) return rc;
    size_t hl = 6 + strlen(VERSION); /* expected_len = the READER's own magic+version length (header.rs:121) */
    if (r->size < hl) return SQO_ERR_IO;
    if (memcmp(r->buf, PREFIX, 6)) return SQO_ERR_ARCHIVE;
    memcpy(r->version, r->buf + 6, hl - 6);
    r->version[hl - 6] = 0;
    /* major.minor string equality */
    char *d1 = strchr(r->version, '.');
    if (!d1) return SQO_ERR_ARCHIVE;
    char *d2 = strchr(d1 + 1, '.');
    size_t mm = d2 ? (size_t)(d2 - r->version) : strlen(r->version);
    const char *c1 = strchr(VERSION, '.'); const char *c2 = strchr(c1 + 1, '.');
    size_t cm = c2 ? (size_t)(c2 - VERSION) : strlen(VERSION);
    if (mm != cm || memcmp(r->version, VERSION, mm)) return SQO_ERR_ARCHIVE;
    uint64_t p = hl;
    if (r->size < p + 16) return SQO_ERR_READER;
    r->timestamp = get64(r->buf + p); p += 8;
    r->nchunks = get64(r->buf + p); p += 8;
    r->chunk_table_off = p;
    for (uint64_t i = 0; i < r->nchunks; i++) { /* serial hop over every record (reader.rs:75-96) */
        if (r->size < p + 32) return SQO_ERR_READER;
        uint64_t cs = get64(r->buf + p + 24);
        p += 32;
        if (cs > r->size - p) return SQO_ERR_READER;
        p += cs;
    }
    if (r->size < p + 4) return SQO_ERR_READER;
    r->file_count = get32(r->buf + p);
    r->file_table_off = p + 4;
    return 0;
}
static int utf8_ok(const uint8_t *s, size_t n) { /* String::from_utf8 (reader.rs:178,350); structural check */
    size_t i = 0;
    while (i < n) {
        uint8_t c = s[i];
        size_t k;
        if (c < 0x80) k = 0;
        else if (c >= 0xC2 && c <= 0xDF) k = 1;
        else if (c >= 0xE0 && c <= 0xEF) k = 2;
        else if (c >= 0xF0 && c <= 0xF4) k = 3;
        else return 0;
        if (i + k >= n + (k == 0 ? 1 : 0) && k) return 0;
        for (size_t j = 1; j <= k; j++) if ((s[i + j] >> 6) != 2) return 0;
        i += k + 1;
    }
    return 1;
}

int sqo_list(const char *archive_path, sqo_summary *out, char **paths_out) {
    reader r;
    int rc = reader_open(&r, archive_path);
    if (rc) { free(r.buf); return rc; }
    uint64_t p = r.file_table_off, total = 0;
    size_t cap = 1 << 16, len = 0;
    char *list = paths_out ? malloc(cap) : NULL;
    for (uint32_t i = 0; i < r.file_count && !rc; i++) { /* get_summary (reader.rs:155-219) */
        if (r.size < p + 4) { rc = SQO_ERR_READER; break; }
        uint32_t pl = get32(r.buf + p); p += 4;
        if (r.size < p + pl + 12) { rc = SQO_ERR_READER; break; }
        if (!utf8_ok(r.buf + p, pl)) { rc = SQO_ERR_ILLEGAL_UTF8; break; }
        const uint8_t *path = r.buf + p; p += pl;
        uint64_t osz = get64(r.buf + p); p += 8;
        uint32_t cc = get32(r.buf + p); p += 4;
        p += (uint64_t)cc * 16;
        total += osz;
        if (list) {
            while (len + pl + 32 > cap) { cap *= 2; list = realloc(list, cap); }
            len += (size_t)snprintf(list + len, cap - len, "%llu ", (unsigned long long)osz);
            memcpy(list + len, path, pl); len += pl;
            list[len++] = '\n';
        }
    }
    if (!rc && out) {
        memset(out, 0, sizeof *out);
        out->unique_chunks = r.nchunks;
        out->total_original_size = total;
        out->archive_size = r.size;
        out->timestamp = r.timestamp;
        out->compression_ratio = total ? (double)r.size / (double)total * 100.0 : 0.0;
        out->file_count = r.file_count;
        strncpy(out->version, r.version, sizeof out->version - 1);
    }
    if (list) { list[len] = 0; if (rc) free(list); else *paths_out = list; }
    free(r.buf);
    return rc;
}

/* digest-keyed map of decoded chunks (reader.rs:268,305) */
typedef struct cm_node { uint8_t h[16]; uint8_t *data; size_t len; struct cm_node *next; } cm_node;
typedef struct { cm_node **b; size_t nb; } chunk_map;
static cm_node *cm_get(chunk_map *m, const uint8_t *h) {
    for (cm_node *n = m->b[(size_t)get64(h) & (m->nb - 1)]; n; n = n->next) if (!memcmp(n->h, h, 16)) return n;
    return NULL;
}
static void cm_put(chunk_map *m, const uint8_t *h, uint8_t *data, size_t len) {
    cm_node *n = cm_get(m, h);
    if (n) { free(n->data); n->data = data; n->len = len; return; } /* HashMap::insert overwrites */
    n = malloc(sizeof *n);
    memcpy(n->h, h, 16); n->data = data; n->len = len;
    size_t b = (size_t)get64(h) & (m->nb - 1);
    n->next = m->b[b]; m->b[b] = n;
}

typedef struct { const uint8_t *rec; uint64_t orig, comp; uint8_t *out; size_t out_len; } dec_job;
typedef struct { dec_job *jobs; uint64_t n, next; pthread_mutex_t mu; int err; } dec_pool;
static int decode_one(dec_job *j) {
    size_t cap = (size_t)j->orig;
    j->out = malloc(cap ? cap : 1);
    size_t got = sqo_zstd_decompress(j->rec + 32, (size_t)j->comp, j->out, cap);
    if (got == (size_t)-1) return SQO_ERR_READER;
    j->out_len = got;
    return 0;
}
static void *dec_worker(void *arg) {
    dec_pool *p = arg;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        uint64_t i = p->next < p->n ? p->next++ : UINT64_MAX;
        pthread_mutex_unlock(&p->mu);
        if (i == UINT64_MAX) break;
        if (decode_one(&p->jobs[i])) p->err = SQO_ERR_READER;
    }
    return NULL;
}

typedef struct { const uint8_t *path; uint32_t pl; const uint8_t *hashes; uint32_t cc; } rb_entry;
typedef struct { rb_entry *e; uint32_t n, next; pthread_mutex_t mu; chunk_map *map; const char *out_dir; int err; } rb_pool;
static void mkdir_p(char *path) {
    for (char *p = path + 1; *p; p++)
        if (*p == '/') { *p = 0; mkdir(path, 0777); *p = '/'; }
}
static void *rb_worker(void *arg) { /* entries.par_iter().try_for_each (reader.rs:380-410) */
    rb_pool *p = arg;
    for (;;) {
        pthread_mutex_lock(&p->mu);
        uint32_t i = p->next < p->n ? p->next++ : UINT32_MAX;
        pthread_mutex_unlock(&p->mu);
        if (i == UINT32_MAX) break;
        rb_entry *e = &p->e[i];
        size_t l = strlen(p->out_dir) + e->pl + 2;
        char *full = malloc(l);
        snprintf(full, l, "%s/%.*s", p->out_dir, (int)e->pl, (const char *)e->path);
        mkdir_p(full);
        FILE *f = fopen(full, "wb");
        if (!f) { p->err = SQO_ERR_IO; free(full); continue; }
        for (uint32_t c = 0; c < e->cc; c++) {
            cm_node *n = cm_get(p->map, e->hashes + (size_t)c * 16);
            if (!n) { p->err = SQO_ERR_MISSING_CHUNK; break; } /* reader.rs:397-401 */
            if (n->len && fwrite(n->data, 1, n->len, f) != n->len) { p->err = SQO_ERR_IO; break; }
        }
        fclose(f);
        free(full);
    }
    return NULL;
}

int sqo_unpack(const char *archive_path, const char *out_dir, int threads, int parallel_decode, sqo_summary *out) {
    if (!sqo_zstd_available()) return SQO_ERR_OTHER;
    if (threads < 1) threads = 1;
    reader r;
    int rc = reader_open(&r, archive_path);
    if (rc) { free(r.buf); return rc; }
    double t0 = now_s();
    /* read_chunks (reader.rs:259-314) */
    dec_job *jobs = calloc(r.nchunks ? r.nchunks : 1, sizeof(dec_job));
    uint64_t p = r.chunk_table_off;
    for (uint64_t i = 0; i < r.nchunks; i++) {
        jobs[i].rec = r.buf + p;
        jobs[i].orig = get64(r.buf + p + 16);
        jobs[i].comp = get64(r.buf + p + 24);
        if (jobs[i].orig > (uint64_t)1 << 40) rc = SQO_ERR_INVALID_CHUNK_SIZE;
        p += 32 + jobs[i].comp;
    }
    if (!rc) {
        if (!parallel_decode || threads == 1) {
            for (uint64_t i = 0; i < r.nchunks && !rc; i++) rc = decode_one(&jobs[i]);
        } else {
            dec_pool dp = { jobs, r.nchunks, 0, PTHREAD_MUTEX_INITIALIZER, 0 };
            pthread_t *th = malloc(sizeof(pthread_t) * (size_t)threads);
            for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, dec_worker, &dp);
            for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
            free(th);
            rc = dp.err;
        }
    }
    chunk_map map;
    map.nb = 1;
    while (map.nb < r.nchunks * 2 + 16) map.nb <<= 1;
    map.b = calloc(map.nb, sizeof(cm_node *));
    for (uint64_t i = 0; i < r.nchunks; i++)
        if (jobs[i].out) { if (!rc) cm_put(&map, jobs[i].rec, jobs[i].out, jobs[i].out_len); else free(jobs[i].out); }
    double t1 = now_s();
    /* rebuild_files (reader.rs:316-413) */
    rb_entry *ent = calloc(r.file_count ? r.file_count : 1, sizeof(rb_entry));
    uint64_t total = 0;
    p = r.file_table_off;
    for (uint32_t i = 0; i < r.file_count && !rc; i++) {
        if (r.size < p + 4) { rc = SQO_ERR_READER; break; }
        ent[i].pl = get32(r.buf + p); p += 4;
        if (r.size < p + ent[i].pl + 12) { rc = SQO_ERR_READER; break; }
        ent[i].path = r.buf + p; p += ent[i].pl;
        if (!utf8_ok(ent[i].path, ent[i].pl)) { rc = SQO_ERR_ILLEGAL_UTF8; break; }
        total += get64(r.buf + p); p += 8;
        ent[i].cc = get32(r.buf + p); p += 4;
        if (r.size < p + (uint64_t)ent[i].cc * 16) { rc = SQO_ERR_READER; break; }
        ent[i].hashes = r.buf + p; p += (uint64_t)ent[i].cc * 16;
    }
    if (!rc) {
        mkdir(out_dir, 0777);
        rb_pool rp = { ent, r.file_count, 0, PTHREAD_MUTEX_INITIALIZER, &map, out_dir, 0 };
        pthread_t *th = malloc(sizeof(pthread_t) * (size_t)threads);
        for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, rb_worker, &rp);
        for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
        free(th);
        rc = rp.err;
    }
    double t2 = now_s();
    if (out) {
        memset(out, 0, sizeof *out);
        out->unique_chunks = r.nchunks; out->total_original_size = total; out->archive_size = r.size;
        out->timestamp = r.timestamp; out->file_count = r.file_count;
        out->compression_ratio = total ? (double)r.size / (double)total * 100.0 : 0.0;
        strncpy(out->version, r.version, sizeof out->version - 1);
        out->decode_seconds = t1 - t0; out->rebuild_seconds = t2 - t1;
    }
    for (size_t b = 0; b < map.nb; b++)
        for (cm_node *n = map.b[b]; n;) { cm_node *nx = n->next; free(n->data); free(n); n = nx; }
    free(map.b); free(ent); free(jobs); free(r.buf);
    return rc;
}
