// Host side of the packer / unpacker: the reference's archive layer re-done around batched GPU calls.
//
//   sq_archive_pack    ArchiveWriter::new + pack      (reference src/archive/writer.rs:66-195,229-329)
//                      + walk_dir                      (src/fsutil/directory.rs:39-73)
//                      + chunk record serialisation    (src/fsutil/writer.rs:17-39)
//   sq_archive_unpack  ArchiveReader::new + unpack     (src/archive/reader.rs:46-118,232-413)
//   sq_archive_list    ArchiveReader::new + get_summary (src/archive/reader.rs:155-219)
//
// The `.squish` byte layout is kept exactly: "squish"+"1.2.0", u64 timestamp,
// u64 unique-chunk count (patched), chunk records {digest[16], orig_size u64 (always 2 MiB, the
// reference's quirk writer.rs:255), comp_size u64, frame}, u32 file count, per-file manifest.
// Differences, all invisible to a reader: chunk records are written in ascending global chunk
// index (the reference's order is channel-arrival order), files are read by a pool of host threads
// into pinned batch buffers while the previous batch is on the GPU, and every per-chunk operation
// (digest, dedup verdict, zstd frame) comes from the CUDA kernels -- never from the CPU.
#include <dirent.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../csrc/common.cuh"

extern "C" int32_t sq_share_dedup(sq_ctx *ctx, sq_ctx *owner);

namespace {

const char kPrefix[] = "squish";           // header.rs:10
const char kVersion[] = SQ_FORMAT_VERSION;  // lib.rs:17

static bool sq_timing() { static int v = -1; if (v < 0) v = getenv("SQ_TIMING") ? 1 : 0; return v == 1; }
#define SQ_T(msg) do { if (sq_timing()) fprintf(stderr, "[sq %8.3f] %s\n", now_s() - t0, msg); } while (0)
double now_s();
double now_s() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + ts.tv_nsec * 1e-9;
}
void put32(uint8_t *p, uint32_t v) { memcpy(p, &v, 4); }  // little-endian host
void put64(uint8_t *p, uint64_t v) { memcpy(p, &v, 8); }
uint32_t get32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
uint64_t get64(const uint8_t *p) { uint64_t v; memcpy(&v, p, 8); return v; }

struct FileItem {
    std::string disk_path, rel_path;
    uint64_t size = 0;
    uint64_t first_chunk = 0;
    uint32_t n_chunks = 0;
};

// walk_dir: iterative stack DFS; directories by stat() (follows symlinks), everything else a file
int32_t walk_dir(sq_ctx *ctx, const std::string &root, std::vector<std::string> *files) {
    struct stat st;
    if (stat(root.c_str(), &st) != 0 || !S_ISDIR(st.st_mode))
        return sq_set_error(ctx, SQ_ERR_READ_DIR, "Failed to read directory %s", root.c_str());
    std::vector<std::string> stack{root};
    while (!stack.empty()) {
        std::string dir = std::move(stack.back());
        stack.pop_back();
        DIR *d = opendir(dir.c_str());
        if (!d) return sq_set_error(ctx, SQ_ERR_READ_DIR, "Failed to read directory %s", dir.c_str());
        while (dirent *e = readdir(d)) {
            if (!strcmp(e->d_name, ".") || !strcmp(e->d_name, "..")) continue;
            std::string p = dir + "/" + e->d_name;
            if (stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode)) stack.push_back(std::move(p));
            else files->push_back(std::move(p));
        }
        closedir(d);
    }
    return SQ_OK;
}

// Host staging for a job: pinned memory moves at full link speed but costs ~0.4 s per GiB to pin, so jobs under a few GiB
// use ordinary page-aligned memory (the driver stages it) and only large jobs pay for pinning once.
struct Staging {
    void *p = nullptr; bool pinned = false;
    int32_t alloc(sq_ctx *ctx, size_t bytes, bool pin) {
        pinned = pin;
        if (pin) return sq_host_alloc(ctx, bytes, &p);
        p = aligned_alloc(4096, (bytes + 4095) & ~(size_t)4095);
        return p ? SQ_OK : sq_set_error(ctx, SQ_ERR_OTHER, "Unknown error: cannot allocate %zu bytes of staging", bytes);
    }
    void release(sq_ctx *ctx) { if (!p) return; if (pinned) sq_host_free(ctx, p); else free(p); p = nullptr; }
};
constexpr uint64_t kPinThreshold = 4ull << 30;

template <class F>
void parallel_for(size_t n, int threads, F fn) {
    if (threads < 1) threads = 1;
    if ((size_t)threads > n) threads = (int)(n ? n : 1);
    std::atomic<size_t> next{0};
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; t++)
        pool.emplace_back([&] { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); });
    for (auto &th : pool) th.join();
}

struct ChunkRef { uint32_t file; uint32_t len; uint64_t file_off; };

// One pinned batch: chunk payloads at 16-byte aligned offsets + spans.
struct Batch {
    uint8_t *pinned = nullptr;
    size_t cap = 0, used = 0;
    std::vector<sq_span> spans;
    uint64_t first_gidx = 0;
    int32_t err = 0;
};

bool utf8_ok(const uint8_t *s, size_t n) {  // String::from_utf8 (reader.rs:178,350)
    size_t i = 0;
    while (i < n) {
        uint8_t c = s[i];
        size_t k;
        if (c < 0x80) k = 0;
        else if (c >= 0xC2 && c <= 0xDF) k = 1;
        else if (c >= 0xE0 && c <= 0xEF) k = 2;
        else if (c >= 0xF0 && c <= 0xF4) k = 3;
        else return false;
        if (i + k >= n + (k ? 0 : 1)) return false;
        for (size_t j = 1; j <= k; j++) if ((s[i + j] >> 6) != 2) return false;
        i += k + 1;
    }
    return true;
}

// ---- archive index (ArchiveReader::new) -----------------------------------------------------
struct Record { const uint8_t *digest; uint64_t orig, comp; const uint8_t *payload; };
struct Archive {
    const uint8_t *buf = nullptr; uint64_t size = 0; int fd = -1;
    uint64_t timestamp = 0, nchunks = 0, file_table_off = 0;
    uint32_t file_count = 0;
    char version[16] = {0};
    std::vector<Record> records;
    ~Archive() { if (buf && size) munmap((void *)buf, size); if (fd >= 0) close(fd); }
};

int32_t open_archive(sq_ctx *ctx, const char *path, Archive *a, bool keep_records) {
    a->fd = open(path, O_RDONLY);
    if (a->fd < 0) return sq_set_error(ctx, SQ_ERR_FILE_NOT_EXIST, "Specified file does not exist: `%s`", path);  // reader.rs:47-48
    struct stat st;
    if (fstat(a->fd, &st)) return sq_set_error(ctx, SQ_ERR_IO, "I/O error: fstat %s", path);
    a->size = (uint64_t)st.st_size;
    if (a->size) {
        void *m = mmap(nullptr, a->size, PROT_READ, MAP_PRIVATE, a->fd, 0);
        if (m == MAP_FAILED) return sq_set_error(ctx, SQ_ERR_IO, "I/O error: mmap %s", path);
        a->buf = (const uint8_t *)m;
    }
    // verify_header (header.rs:119-163): expected_len is the READER's own magic+version length
    const size_t hl = 6 + strlen(kVersion);
    if (a->size < hl) return sq_set_error(ctx, SQ_ERR_IO, "I/O error: failed to fill whole buffer");
    if (memcmp(a->buf, kPrefix, 6)) return sq_set_error(ctx, SQ_ERR_ARCHIVE, "Archive format error: Invalid archive header: prefix mismatch");
    memcpy(a->version, a->buf + 6, hl - 6);
    if (!utf8_ok((const uint8_t *)a->version, hl - 6)) return sq_set_error(ctx, SQ_ERR_ARCHIVE, "Archive format error: Invalid UTF-8 in version string");
    std::string v(a->version), cur(kVersion);
    auto major_minor = [](const std::string &s, std::string *mm) {
        size_t d1 = s.find('.');
        if (d1 == std::string::npos) return false;
        size_t d2 = s.find('.', d1 + 1);
        *mm = s.substr(0, d2);
        return true;
    };
    std::string hmm, cmm;
    if (!major_minor(v, &hmm)) return sq_set_error(ctx, SQ_ERR_ARCHIVE, "Archive format error: Invalid version format in archive header");
    major_minor(cur, &cmm);
    if (hmm != cmm)
        return sq_set_error(ctx, SQ_ERR_ARCHIVE, "Archive format error: Incompatible version: archive %s vs current %s", hmm.c_str(), cmm.c_str());
    uint64_t p = hl;
    if (a->size < p + 16) return sq_set_error(ctx, SQ_ERR_READER, "Error reading from squish: failed to fill whole buffer");
    a->timestamp = get64(a->buf + p); p += 8;
    a->nchunks = get64(a->buf + p); p += 8;
    if (keep_records) a->records.reserve((size_t)std::min<uint64_t>(a->nchunks, 1u << 24));
    for (uint64_t i = 0; i < a->nchunks; i++) {  // reader.rs:75-96
        if (a->size < p + 32) return sq_set_error(ctx, SQ_ERR_READER, "Error reading from squish: failed to fill whole buffer");
        Record r{a->buf + p, get64(a->buf + p + 16), get64(a->buf + p + 24), a->buf + p + 32};
        p += 32;
        if (r.comp > a->size - p) return sq_set_error(ctx, SQ_ERR_READER, "Error reading from squish: failed to fill whole buffer");
        p += r.comp;
        if (keep_records) a->records.push_back(r);
    }
    if (a->size < p + 4) return sq_set_error(ctx, SQ_ERR_READER, "Error reading from squish: failed to fill whole buffer");
    a->file_count = get32(a->buf + p);
    a->file_table_off = p + 4;
    return SQ_OK;
}

struct ManifestEntry { const uint8_t *path; uint32_t path_len; uint64_t orig_size; uint32_t chunk_count; const uint8_t *hashes; };
int32_t read_manifest(sq_ctx *ctx, const Archive &a, std::vector<ManifestEntry> *out, uint64_t *total) {
    uint64_t p = a.file_table_off;
    *total = 0;
    for (uint32_t i = 0; i < a.file_count; i++) {
        ManifestEntry e;
        if (a.size < p + 4) return sq_set_error(ctx, SQ_ERR_READER, "Error reading from squish: failed to fill whole buffer");
        e.path_len = get32(a.buf + p); p += 4;
        if (a.size < p + e.path_len + 12) return sq_set_error(ctx, SQ_ERR_READER, "Error reading from squish: failed to fill whole buffer");
        e.path = a.buf + p; p += e.path_len;
        if (!utf8_ok(e.path, e.path_len)) return sq_set_error(ctx, SQ_ERR_ILLEGAL_UTF8, "Illegal UTF8 detected");
        e.orig_size = get64(a.buf + p); p += 8;
        e.chunk_count = get32(a.buf + p); p += 4;
        if (a.size < p + (uint64_t)e.chunk_count * 16) return sq_set_error(ctx, SQ_ERR_READER, "Error reading from squish: failed to fill whole buffer");
        e.hashes = a.buf + p; p += (uint64_t)e.chunk_count * 16;
        *total += e.orig_size;
        out->push_back(e);
    }
    return SQ_OK;
}

// Upper bound of the decoded size of a record payload from its frame headers (all frames must carry
// a Frame_Content_Size), else `fallback`.  Lets 200k small records avoid 2 MiB of capacity each.
uint64_t payload_decoded_bound(const uint8_t *p, uint64_t n, uint64_t fallback, bool *exact = nullptr) {
    if (exact) *exact = false;
    // only the first frame is inspected: a single-frame payload with FCS is what every writer emits
    if (n < 6 || get32(p) != 0xFD2FB528u) return fallback;
    uint8_t fhd = p[4];
    uint32_t fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
    uint32_t pos = 5 + (single ? 0 : 1) + (did == 3 ? 4 : did);
    uint32_t fcs_bytes = fcs_flag == 0 ? (single ? 1 : 0) : fcs_flag == 1 ? 2 : fcs_flag == 2 ? 4 : 8;
    if (!fcs_bytes || n < pos + fcs_bytes) return fallback;
    uint64_t fcs = 0;
    memcpy(&fcs, p + pos, fcs_bytes);
    if (fcs_flag == 1) fcs += 256;
    // more than one frame in the payload?  walk blocks to the end of frame 1
    uint64_t q = pos + fcs_bytes;
    for (;;) {
        if (n < q + 3) return fallback;
        uint32_t bh = p[q] | p[q + 1] << 8 | p[q + 2] << 16;
        uint32_t type = (bh >> 1) & 3, bsz = bh >> 3;
        q += 3 + (type == 1 ? 1 : type == 3 ? 0 : bsz);
        if (type == 3) return fallback;
        if (bh & 1) break;
    }
    if (fhd & 4) q += 4;
    if (q != n) return fallback;  // trailing frames / garbage: let the decoder decide with the full capacity
    if (exact) *exact = fcs <= fallback;
    return std::min<uint64_t>(fcs, fallback);
}

}  // namespace

extern "C" int32_t sq_archive_list(const char *archive_path, sq_summary *summary, char **listing) {
    if (!archive_path) return SQ_ERR_INVALID_ARG;
    Archive a;
    int32_t rc = open_archive(nullptr, archive_path, &a, false);
    if (rc) return rc;
    std::vector<ManifestEntry> man;
    uint64_t total = 0;
    if ((rc = read_manifest(nullptr, a, &man, &total))) return rc;
    if (summary) {
        memset(summary, 0, sizeof *summary);
        summary->unique_chunks = a.nchunks;
        summary->total_original_size = total;
        summary->archive_size = a.size;
        summary->timestamp = a.timestamp;
        summary->compression_ratio = total ? (double)a.size / (double)total * 100.0 : 0.0;  // reader.rs:204-208
        summary->file_count = a.file_count;
        snprintf(summary->version, sizeof summary->version, "%s", a.version);
    }
    if (listing) {
        std::string s;
        for (auto &e : man) { s += std::to_string(e.orig_size); s += ' '; s.append((const char *)e.path, e.path_len); s += '\n'; }
        char *out = (char *)malloc(s.size() + 1);
        memcpy(out, s.c_str(), s.size() + 1);
        *listing = out;
    }
    return SQ_OK;
}

// One host thread drives every device: batch k of the chunk stream goes to device k % D, pipeline slot (k / D) % 2, and batches
// are waited for in the order they were submitted, so records reach the archive in chunk order.  With D > 1 the contexts share
// ctxs[0]'s dedup index (sq_share_dedup): digests meet there in submission order.
static int32_t pack_impl(sq_ctx **ctxs, uint32_t D, const char *input_dir, const char *output_path, int32_t threads, sq_pack_report *report) {
    sq_ctx *ctx = ctxs[0];
    auto on = [&](uint32_t d) -> sq_ctx * { if (D > 1) cudaSetDevice(ctxs[d]->device); return ctxs[d]; };
    const double t0 = now_s();
    double t_dev = 0;
    if (threads < 1) threads = 1;
    std::string root(input_dir);
    while (root.size() > 1 && (root.back() == '/' || root.back() == '\\')) root.pop_back();  // lib.rs:28
    std::vector<std::string> paths;
    int32_t rc = walk_dir(ctx, root, &paths);
    if (rc) return rc;

    SQ_T("walk done");
    std::vector<FileItem> files(paths.size());
    std::atomic<int32_t> ferr{0};
    parallel_for(paths.size(), threads, [&](size_t i) {
        struct stat st;
        files[i].disk_path = paths[i];
        files[i].rel_path = paths[i].substr(root.size() + 1);  // strip_prefix(input_path) (writer.rs:230)
        if (stat(paths[i].c_str(), &st)) ferr = SQ_ERR_IO; else files[i].size = (uint64_t)st.st_size;
    });
    if (ferr) return sq_set_error(ctx, SQ_ERR_IO, "I/O error: stat failed under %s", root.c_str());
    const uint32_t cs = ctx->chunk_size;
    std::vector<ChunkRef> chunks;
    for (size_t i = 0; i < files.size(); i++) {  // chunk rule writer.rs:240-246
        files[i].first_chunk = chunks.size();
        for (uint64_t off = 0; off < files[i].size; off += cs)
            chunks.push_back({(uint32_t)i, (uint32_t)std::min<uint64_t>(cs, files[i].size - off), off});
        files[i].n_chunks = (uint32_t)(chunks.size() - files[i].first_chunk);
    }
    const uint64_t total_chunks = chunks.size();
    SQ_T("stat + chunk plan done");
    for (uint32_t d = 0; d < D && D > 1; d++) if ((rc = sq_share_dedup(ctxs[d], ctx))) { if (d) snprintf(ctx->err, sizeof ctx->err, "%s", ctxs[d]->err); return rc; }
    on(0);
    if ((rc = sq_dedup_ensure(ctx, total_chunks))) return rc;  // the index is sized to the job, not to the context's maximum
    if ((rc = sq_dedup_reset(ctx))) return rc;  // ChunkStore::new (writer.rs:88)
    SQ_T("dedup index ready");

    FILE *out = fopen(output_path, "wb+");
    if (!out) return sq_set_error(ctx, SQ_ERR_IO, "I/O error: cannot create %s", output_path);
    setvbuf(out, nullptr, _IOFBF, 4 << 20);
    uint8_t pre[32];
    const size_t hl = 6 + strlen(kVersion);
    memcpy(pre, kPrefix, 6); memcpy(pre + 6, kVersion, hl - 6);      // write_header (header.rs:35-38)
    put64(pre + hl, (uint64_t)time(nullptr));                          // write_timestamp (header.rs:55-63)
    put64(pre + hl + 8, 0);                                            // write_placeholder_u64 (header.rs:192-196)
    const long count_pos = (long)hl + 8;
    bool werr = fwrite(pre, 1, hl + 16, out) != hl + 16;

    // batches: up to batch_bytes of chunk payload or max_batch chunks
    // Batches of up to 512 MiB of chunk payload: one chunk occupies one search CTA for its whole encode, so a batch should
    // offer a few hundred chunks; pinned staging is sized to the job, because pinning memory costs ~0.4 s per GiB.
    uint64_t total_bytes_all = 0;
    for (auto &f : files) total_bytes_all += f.size + 16;
    size_t batch_bytes = (size_t)std::min<uint64_t>(512ull << 20, std::max<uint64_t>(total_bytes_all / D, 1u << 20));
    if (const char *bb = getenv("SQ_PACK_BATCH_BYTES")) if (atoll(bb) > 0) batch_bytes = (size_t)atoll(bb);  // tests: many small batches
    const size_t slot_cap = batch_bytes + (size_t)cs;
    const bool pin = total_bytes_all >= kPinThreshold;
    const uint32_t S = 2 * D;  // pipeline slots: two per device
    std::vector<Batch> bufs(S);
    std::vector<Staging> in_stage(S), out_stage(S), res_stage(S);
    auto bail = [&](int32_t code) {  // early exit: give the staging back and do not leave a truncated archive behind
        for (uint32_t i = 0; i < S; i++) { sq_ctx *c = on(i % D); in_stage[i].release(c); out_stage[i].release(c); res_stage[i].release(c); }
        for (uint32_t d = 0; d < D && D > 1; d++) sq_share_dedup(ctxs[d], nullptr);
        on(0);
        fclose(out);
        remove(output_path);
        return code;
    };
    uint64_t out_cap = sq_encode_bound(cs) * (uint64_t)(batch_bytes / cs + 1) + batch_bytes / 16;
    // the per-chunk results are small and always pinned: their download then never blocks sq_pack_submit
    struct Stage { sq_chunk_result *res = nullptr; void *h_out = nullptr; sq_ticket *ticket = nullptr; bool live = false; };
    std::vector<Stage> stages(S);
    // A slot's staging is allocated when the slot is first used: pinning costs ~0.4 s per GiB, and this way the first batch is on
    // its GPU while the memory of the later slots is still being pinned (a job of one batch never pays for the others).
    auto ensure_slot = [&](uint32_t i) -> int32_t {
        if (in_stage[i].p) return SQ_OK;
        sq_ctx *c = on(i % D);
        int32_t r;
        if ((r = in_stage[i].alloc(c, slot_cap, pin))) return r;
        bufs[i].pinned = (uint8_t *)in_stage[i].p; bufs[i].cap = slot_cap;
        if ((r = res_stage[i].alloc(c, (size_t)ctx->max_batch * sizeof(sq_chunk_result), true))) return r;
        stages[i].res = (sq_chunk_result *)res_stage[i].p;
        if ((r = out_stage[i].alloc(c, out_cap, pin))) return r;
        stages[i].h_out = out_stage[i].p;
        return SQ_OK;
    };
    std::vector<uint8_t> digests(total_chunks * 16);

    auto fill = [&](Batch *b, uint64_t first, uint64_t *next) {  // host reader pool -> pinned buffer
        b->spans.clear(); b->used = 0; b->first_gidx = first; b->err = 0;
        uint64_t g = first;
        while (g < total_chunks && b->spans.size() < ctx->max_batch) {
            size_t need = ((size_t)chunks[g].len + 15) & ~(size_t)15;
            if (b->used + need > b->cap || (b->used >= batch_bytes)) break;
            b->spans.push_back({b->used, chunks[g].len, 0});
            b->used += need;
            g++;
        }
        *next = g;
        std::atomic<int32_t> err{0};
        parallel_for(b->spans.size(), threads, [&](size_t i) {
            const ChunkRef &c = chunks[first + i];
            int fd = open(files[c.file].disk_path.c_str(), O_RDONLY);
            if (fd < 0) { err = SQ_ERR_IO; return; }
            size_t got = 0;
            while (got < c.len) {
                ssize_t r = pread(fd, b->pinned + b->spans[i].off + got, c.len - got, (off_t)(c.file_off + got));
                if (r <= 0) { err = SQ_ERR_READER; break; }
                got += (size_t)r;
            }
            close(fd);
        });
        b->err = err;
    };

    // Software pipeline over two slots: while batch k is on the GPU (sq_pack_submit), the host threads read batch k+1
    // into the other pinned buffer; sq_pack_wait(k) then overlaps its frame download with the kernels of batch k+1.
    uint64_t next = 0, unique = 0, payload = 0;
    uint32_t head = 0, tail = 0, inflight = 0;
    while ((next < total_chunks || inflight) && !rc) {
        if (next < total_chunks && inflight < S) {
            if ((rc = ensure_slot(head))) { if (head % D) snprintf(ctx->err, sizeof ctx->err, "%s", ctxs[head % D]->err); break; }
            Batch *b = &bufs[head];
            fill(b, next, &next);
            if (b->err) { rc = sq_set_error(ctx, b->err, "Error reading from squish: input file changed or unreadable"); break; }
            double td = now_s();
            rc = sq_pack_submit(on(head % D), b->pinned, b->used, b->spans.data(), (uint32_t)b->spans.size(), b->first_gidx, stages[head].res,
                                stages[head].h_out, out_cap, &stages[head].ticket);
            t_dev += now_s() - td;
            if (rc && head % D) snprintf(ctx->err, sizeof ctx->err, "%s", ctxs[head % D]->err);
            if (!rc) { stages[head].live = true; head = (head + 1) % S; inflight++; }
            continue;
        }
        Batch *b = &bufs[tail];
        Stage &sg = stages[tail];
        uint64_t used = 0;
        double td = now_s();
        rc = sq_pack_wait(on(tail % D), sg.ticket, &used);
        if (rc && tail % D) snprintf(ctx->err, sizeof ctx->err, "%s", ctxs[tail % D]->err);
        sg.live = false;
        t_dev += now_s() - td;
        if (!rc) {
            const uint32_t n = (uint32_t)b->spans.size();
            for (uint32_t i = 0; i < n; i++) {
                memcpy(&digests[(b->first_gidx + i) * 16], sg.res[i].digest, 16);
                if (!sg.res[i].is_new) continue;
                uint8_t rec[32];  // chunk record (fsutil/writer.rs:21-36)
                memcpy(rec, sg.res[i].digest, 16);
                put64(rec + 16, (uint64_t)SQ_CHUNK_SIZE);  // original_size: chunk_buf.len() (writer.rs:255)
                put64(rec + 24, sg.res[i].frame_len);
                werr |= fwrite(rec, 1, 32, out) != 32;
                werr |= fwrite((uint8_t *)sg.h_out + sg.res[i].frame_off, 1, sg.res[i].frame_len, out) != sg.res[i].frame_len;
                unique++;
                payload += sg.res[i].frame_len;
            }
        }
        tail = (tail + 1) % S; inflight--;
    }
    if (rc) {  // drain whatever is still in flight (through the tickets that were issued) before the buffers go away
        for (uint32_t i = 0; i < S; i++) if (stages[i].live) { uint64_t u; sq_pack_wait(on(i % D), stages[i].ticket, &u); stages[i].live = false; }
    }
    SQ_T("all batches packed + records written");
    if (!rc) {
        uint64_t dl = 0;
        rc = sq_dedup_len(on(0), &dl);  // chunk_store.len() (writer.rs:177-184)
        if (!rc && dl != unique) rc = sq_set_error(ctx, SQ_ERR_OTHER, "Unknown error: dedup index holds %llu digests, %llu records written", (unsigned long long)dl, (unsigned long long)unique);
    }
    if (!rc) {
        uint8_t b8[8], b4[4];
        put64(b8, unique);
        fseek(out, count_pos, SEEK_SET);               // patch_u64 (header.rs:224-233)
        werr |= fwrite(b8, 1, 8, out) != 8;
        fseek(out, 0, SEEK_END);
        put32(b4, (uint32_t)files.size());             // write_files_metadata (writer.rs:292-329)
        werr |= fwrite(b4, 1, 4, out) != 4;
        for (auto &f : files) {
            put32(b4, (uint32_t)f.rel_path.size()); werr |= fwrite(b4, 1, 4, out) != 4;
            werr |= fwrite(f.rel_path.data(), 1, f.rel_path.size(), out) != f.rel_path.size();
            put64(b8, f.size); werr |= fwrite(b8, 1, 8, out) != 8;
            put32(b4, f.n_chunks); werr |= fwrite(b4, 1, 4, out) != 4;
            if (f.n_chunks) werr |= fwrite(&digests[f.first_chunk * 16], 16, f.n_chunks, out) != f.n_chunks;
        }
        if (fflush(out)) werr = true;
        if (werr) rc = sq_set_error(ctx, SQ_ERR_WRITER, "Error writing to squish: %s", output_path);
    }
    SQ_T("manifest written");
    uint64_t asize = 0;
    if (!rc) { fseek(out, 0, SEEK_END); asize = (uint64_t)ftell(out); }
    fclose(out);
    for (uint32_t i = 0; i < S; i++) { sq_ctx *c = on(i % D); in_stage[i].release(c); out_stage[i].release(c); res_stage[i].release(c); }
    for (uint32_t d = 0; d < D && D > 1; d++) sq_share_dedup(ctxs[d], nullptr);  // the sharing lasts for this pack only
    on(0);
    if (!rc && report) {
        memset(report, 0, sizeof *report);
        report->archive_size = asize; report->unique_chunks = unique; report->total_chunks = total_chunks;
        uint64_t tb = 0; for (auto &f : files) tb += f.size;
        report->total_input_bytes = tb; report->payload_bytes = payload; report->file_count = (uint32_t)files.size();
        report->seconds_total = now_s() - t0; report->seconds_device = t_dev;
    }
    return rc;
}

extern "C" int32_t sq_archive_pack(sq_ctx *ctx, const char *input_dir, const char *output_path, int32_t threads, sq_pack_report *report) {
    if (!ctx || !input_dir || !output_path) return SQ_ERR_INVALID_ARG;
    return pack_impl(&ctx, 1, input_dir, output_path, threads, report);
}

extern "C" int32_t sq_archive_pack_multi(sq_ctx **ctxs, uint32_t n_ctx, const char *input_dir, const char *output_path, int32_t threads,
                                         sq_pack_report *report) {
    if (!ctxs || n_ctx == 0 || !input_dir || !output_path) return SQ_ERR_INVALID_ARG;
    for (uint32_t i = 0; i < n_ctx; i++) if (!ctxs[i]) return SQ_ERR_INVALID_ARG;
    return pack_impl(ctxs, n_ctx, input_dir, output_path, threads, report);
}

// ---- unpack, streaming: decoded batches go straight from the staging buffer into the files ------------------------------
// read_chunks + rebuild_files (reader.rs:259-413) fused.  Every frame of the archive states its decoded size (every writer of the
// format emits a Frame_Content_Size), so the byte offset of every chunk in every file is known before anything is decoded: a
// batch that comes back from the GPU is written with pwrite by the host threads while the next batch decodes, and host memory
// holds two staging pairs instead of all unique data (the reference keeps everything: reader.rs:268).
struct UnpackJob {  // what every device of an unpack needs: where each record's bytes go, and the files, already created
    struct Ref { uint32_t file; uint64_t off; };
    std::vector<uint32_t> ref_start;
    std::vector<Ref> refs;
    std::vector<std::string> full;
    bool lazy_create = false;  // the output directory is new: a file is created by whoever writes one of its chunks first
    uint64_t total_out = 0, total_comp = 0;
};

int32_t unpack_prepare(sq_ctx *ctx, Archive &a, std::vector<ManifestEntry> &man, const std::vector<uint64_t> &size, const char *output_dir,
                       int32_t threads, UnpackJob *job) {
    const size_t nrec = a.records.size();
    struct Key { uint64_t a, b; bool operator==(const Key &o) const { return a == o.a && b == o.b; } };
    struct KeyHash { size_t operator()(const Key &k) const { return (size_t)(k.a ^ (k.b * 0x9E3779B97F4A7C15ULL)); } };
    std::unordered_map<Key, size_t, KeyHash> map;  // digest -> record; later records overwrite earlier ones (reader.rs:305)
    map.reserve(nrec * 2);
    for (size_t k = 0; k < nrec; k++) map[{get64(a.records[k].digest), get64(a.records[k].digest + 8)}] = k;
    // where every record's bytes go: (file, offset) references in CSR form
    job->ref_start.assign(nrec + 1, 0);
    std::vector<size_t> chunk_rec;
    size_t total_refs = 0;
    for (auto &e : man) total_refs += e.chunk_count;
    chunk_rec.reserve(total_refs);
    for (size_t fi = 0; fi < man.size(); fi++) {
        const ManifestEntry &e = man[fi];
        for (uint32_t c = 0; c < e.chunk_count; c++) {
            auto it = map.find({get64(e.hashes + (size_t)c * 16), get64(e.hashes + (size_t)c * 16 + 8)});
            if (it == map.end())  // reader.rs:397-401
                return sq_set_error(ctx, SQ_ERR_MISSING_CHUNK, "%s `%.*s`", sq_strerror(SQ_ERR_MISSING_CHUNK), (int)e.path_len, (const char *)e.path);
            chunk_rec.push_back(it->second);
            job->ref_start[it->second + 1]++;
        }
    }
    for (size_t k = 0; k < nrec; k++) job->ref_start[k + 1] += job->ref_start[k];
    job->refs.resize(total_refs);
    {
        std::vector<uint32_t> fill(job->ref_start.begin(), job->ref_start.end() - 1);
        size_t j = 0;
        for (size_t fi = 0; fi < man.size(); fi++) {
            uint64_t off = 0;
            for (uint32_t c = 0; c < man[fi].chunk_count; c++, j++) {
                job->refs[fill[chunk_rec[j]]++] = {(uint32_t)fi, off};
                off += size[chunk_rec[j]];
            }
        }
    }
    // create_dir_all(parent) + File::create for every file (reader.rs:375-383).  Directories are made once each, not once per
    // file.  When the output directory did not exist before, nothing stale can be in it: only the empty files are created here and
    // every other file is created by whoever writes one of its chunks first (O_CREAT) -- two system calls fewer per file, which is
    // most of the host time of an archive of 200 000 small files.  Into an existing directory every file is created and truncated
    // up front, as the reference does.
    std::string outdir(output_dir);
    job->lazy_create = mkdir(outdir.c_str(), 0777) == 0 && !getenv("SQ_UNPACK_EAGER_CREATE");
    job->full.resize(man.size());
    std::atomic<int32_t> err{0};
    std::atomic<size_t> err_idx{0};
    {
        std::unordered_map<std::string, char> seen;
        std::vector<std::string> dirs;  // parents before children
        for (size_t fi = 0; fi < man.size(); fi++) {
            const ManifestEntry &e = man[fi];
            job->full[fi] = outdir + "/" + std::string((const char *)e.path, e.path_len);
            const std::string &f = job->full[fi];
            const size_t last = f.rfind('/');
            if (last == std::string::npos || last <= outdir.size() || seen.count(f.substr(0, last))) continue;
            for (size_t p = outdir.size() + 1; p <= last; p++)
                if (f[p] == '/' && seen.emplace(f.substr(0, p), 1).second) dirs.push_back(f.substr(0, p));
        }
        for (auto &d : dirs) mkdir(d.c_str(), 0777);
    }
    parallel_for(man.size(), threads, [&](size_t fi) {
        if (job->lazy_create && man[fi].chunk_count) return;
        int fd = open(job->full[fi].c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
        if (fd < 0) { err = SQ_ERR_CREATE_FILE; err_idx = fi; return; }
        close(fd);
    });
    if (err) {
        const ManifestEntry &e = man[err_idx];
        return sq_set_error(ctx, err, "%s `%.*s`", sq_strerror(err), (int)e.path_len, (const char *)e.path);
    }
    for (size_t k = 0; k < nrec; k++) { job->total_out += (size[k] + 15) & ~15ull; job->total_comp += (a.records[k].comp + 15) & ~15ull; }
    return SQ_OK;
}

// decodes records [lo, hi) on ctx's device and writes their bytes into the files (one call per device of a multi-GPU unpack)
int32_t unpack_range(sq_ctx *ctx, Archive &a, std::vector<ManifestEntry> &man, const std::vector<uint64_t> &size, const UnpackJob &job,
                     size_t lo, size_t hi, int32_t threads, double *t_dev_out) {
    double t_dev = 0;
    int32_t rc = SQ_OK;
    uint64_t total_out = 0, total_comp = 0;
    for (size_t k = lo; k < hi; k++) { total_out += (size[k] + 15) & ~15ull; total_comp += (a.records[k].comp + 15) & ~15ull; }
    std::atomic<int32_t> err{0};
    std::atomic<size_t> err_idx{0};
    // batches: at most 1 GiB of output per slot, and at least four batches for jobs above 256 MiB so that upload, decode,
    // download and file writes overlap
    const uint64_t batch_out = std::min<uint64_t>(1024ull << 20, std::max<uint64_t>(total_out > (256ull << 20) ? total_out / 4 : total_out, 1u << 20)),
                   batch_in = std::min<uint64_t>(1024ull << 20, std::max<uint64_t>(total_comp, 1u << 20));
    struct UnpackBatch {
        Staging comp, out, res;
        std::vector<sq_frame> frames;
        sq_frame_result *fres = nullptr;
        size_t first = 0;
        sq_ticket *ticket = nullptr; bool live = false;
    } ub[2];
    const bool pin = total_out >= kPinThreshold;
    for (int k = 0; k < 2 && !rc; k++) {
        if ((rc = ub[k].comp.alloc(ctx, batch_in + (4u << 20), pin))) break;
        if ((rc = ub[k].out.alloc(ctx, batch_out + (4u << 20), pin))) break;
        if ((rc = ub[k].res.alloc(ctx, (size_t)ctx->max_batch * sizeof(sq_frame_result), pin))) break;  // pinned with the data: downloads then never block a submit
        ub[k].fres = (sq_frame_result *)ub[k].res.p;
    }
    auto release = [&]() { for (int k = 0; k < 2; k++) { ub[k].comp.release(ctx); ub[k].out.release(ctx); ub[k].res.release(ctx); } };
    if (rc) { release(); return rc; }
    auto finish = [&](UnpackBatch &b) -> int32_t {  // wait for the batch, check every payload, write its bytes where they belong
        b.live = false;
        double td = now_s();
        int32_t r = sq_unpack_wait(ctx, b.ticket);
        t_dev += now_s() - td;
        if (r) return r;
        for (size_t k = 0; k < b.frames.size(); k++)
            if (b.fres[k].status != SQ_OK || b.fres[k].out_len != size[b.first + k])
                return sq_set_error(ctx, SQ_ERR_READER, "Error reading from squish: chunk %zu failed to decode", b.first + k);
        parallel_for(b.frames.size(), threads, [&](size_t k) {
            const size_t ri = b.first + k;
            const uint8_t *src = (const uint8_t *)b.out.p + b.frames[k].dst_off;
            for (uint32_t j = job.ref_start[ri]; j < job.ref_start[ri + 1]; j++) {
                int fd = open(job.full[job.refs[j].file].c_str(), job.lazy_create ? O_WRONLY | O_CREAT : O_WRONLY, 0666);
                if (fd < 0) { err = SQ_ERR_CREATE_FILE; err_idx = job.refs[j].file; return; }
                size_t w = 0;
                while (w < size[ri]) {
                    ssize_t n = pwrite(fd, src + w, size[ri] - w, (off_t)(job.refs[j].off + w));
                    if (n <= 0) { err = SQ_ERR_IO; err_idx = job.refs[j].file; break; }
                    w += (size_t)n;
                }
                close(fd);
            }
        });
        return SQ_OK;
    };
    size_t i = lo;
    int cur = 0;
    while (i < hi && !rc && !err) {
        UnpackBatch &b = ub[cur];
        if (b.live && (rc = finish(b))) break;  // the slot's previous batch must be on disk before its buffers are reused
        b.frames.clear();
        uint64_t so = 0, dof = 0;
        b.first = i;
        while (i < hi && b.frames.size() < ctx->max_batch) {
            const Record &r = a.records[i];
            uint64_t sneed = (r.comp + 15) & ~15ull, dneed = (size[i] + 15) & ~15ull;
            if (!b.frames.empty() && (so + sneed > batch_in || dof + dneed > batch_out)) break;
            if (sneed > batch_in + (4u << 20) || dneed > batch_out + (4u << 20)) { rc = sq_set_error(ctx, SQ_ERR_INVALID_CHUNK_SIZE, "Invalid chunk size: %llu bytes", (unsigned long long)r.orig); break; }
            b.frames.push_back({so, dof, (uint32_t)r.comp, (uint32_t)size[i]});
            so += sneed; dof += dneed;
            i++;
        }
        if (rc) break;
        parallel_for(b.frames.size(), threads, [&](size_t k) { memcpy((uint8_t *)b.comp.p + b.frames[k].src_off, a.records[b.first + k].payload, b.frames[k].src_len); });
        double td = now_s();
        rc = sq_unpack_submit(ctx, b.comp.p, so, b.frames.data(), (uint32_t)b.frames.size(), b.out.p, dof, b.fres, &b.ticket);
        t_dev += now_s() - td;
        if (rc) break;
        b.live = true;
        cur ^= 1;
    }
    for (int k = 0; k < 2; k++) {  // drain, oldest first; on error still wait before the buffers go away
        UnpackBatch &b = ub[cur ^ k];
        if (!b.live) continue;
        if (rc || err) { sq_unpack_wait(ctx, b.ticket); b.live = false; }
        else rc = finish(b);
    }
    release();
    if (t_dev_out) *t_dev_out = t_dev;
    if (rc) return rc;
    if (err) {
        const ManifestEntry &e = man[err_idx];
        return sq_set_error(ctx, err, "%s `%.*s`", sq_strerror(err), (int)e.path_len, (const char *)e.path);
    }
    return SQ_OK;
}

void fill_summary(sq_summary *summary, const Archive &a, uint64_t total, double t0, double t_dev) {
    if (!summary) return;
    memset(summary, 0, sizeof *summary);
    summary->unique_chunks = a.nchunks; summary->total_original_size = total; summary->archive_size = a.size;
    summary->timestamp = a.timestamp; summary->file_count = a.file_count;
    summary->compression_ratio = total ? (double)a.size / (double)total * 100.0 : 0.0;
    snprintf(summary->version, sizeof summary->version, "%s", a.version);
    summary->seconds_total = now_s() - t0; summary->seconds_device = t_dev;
}

int32_t unpack_streaming(sq_ctx *ctx, Archive &a, std::vector<ManifestEntry> &man, uint64_t total, const std::vector<uint64_t> &size,
                         const char *output_dir, int32_t threads, sq_summary *summary, double t0) {
    UnpackJob job;
    int32_t rc = unpack_prepare(ctx, a, man, size, output_dir, threads, &job);
    if (rc) return rc;
    SQ_T("files created");
    double t_dev = 0;
    rc = unpack_range(ctx, a, man, size, job, 0, a.records.size(), threads, &t_dev);
    SQ_T("all chunks decoded and written");
    if (rc) return rc;
    fill_summary(summary, a, total, t0, t_dev);
    return SQ_OK;
}

// Unpack on several GPUs of one box (records are independent, so the path shards with no collective): the records
// are cut into contiguous ranges of equal compressed + restored bytes, one range per context; every context runs the same
// two-slot pipeline on its own device from its own host thread, and all of them write into the files created up front.
extern "C" int32_t sq_archive_unpack_multi(sq_ctx **ctxs, uint32_t n_ctx, const char *archive_path, const char *output_dir, int32_t threads,
                                           sq_summary *summary) {
    if (!ctxs || n_ctx == 0 || !ctxs[0]) return SQ_ERR_INVALID_ARG;
    for (uint32_t i = 0; i < n_ctx; i++) if (!ctxs[i]) return SQ_ERR_INVALID_ARG;
    if (n_ctx == 1) return sq_archive_unpack(ctxs[0], archive_path, output_dir, threads, summary);
    if (!archive_path || !output_dir) return SQ_ERR_INVALID_ARG;
    sq_ctx *ctx = ctxs[0];
    const double t0 = now_s();
    if (threads < 1) threads = 1;
    Archive a;
    int32_t rc = open_archive(ctx, archive_path, &a, true);
    if (rc) return rc;
    std::vector<ManifestEntry> man;
    uint64_t total = 0;
    if ((rc = read_manifest(ctx, a, &man, &total))) return rc;
    const size_t nrec = a.records.size();
    std::vector<uint64_t> bound(nrec);
    bool all_exact = true;
    for (size_t i = 0; i < nrec; i++) {
        const Record &r = a.records[i];
        if (r.orig > (uint64_t)1 << 40) return sq_set_error(ctx, SQ_ERR_INVALID_CHUNK_SIZE, "Invalid chunk size: %llu bytes", (unsigned long long)r.orig);
        if (r.orig > 0xFFFFFFFFull || r.comp > 0xFFFFFFFFull) return sq_set_error(ctx, SQ_ERR_READER, "Error reading from squish: chunk record too large");
        bool ex = false;
        bound[i] = payload_decoded_bound(r.payload, r.comp, r.orig, &ex);
        all_exact = all_exact && ex;
    }
    if (!all_exact) return sq_archive_unpack(ctx, archive_path, output_dir, threads, summary);  // frames without a stated size: the one-device store path
    UnpackJob job;
    if ((rc = unpack_prepare(ctx, a, man, bound, output_dir, threads, &job))) return rc;
    std::vector<size_t> cut(n_ctx + 1, nrec);
    {
        const uint64_t all = job.total_out + job.total_comp;
        uint64_t acc = 0;
        uint32_t k = 1;
        cut[0] = 0;
        for (size_t i = 0; i < nrec && k < n_ctx; i++) {
            acc += ((bound[i] + 15) & ~15ull) + ((a.records[i].comp + 15) & ~15ull);
            while (k < n_ctx && acc * n_ctx >= all * k) cut[k++] = i + 1;
        }
    }
    std::vector<int32_t> rcs(n_ctx, SQ_OK);
    std::vector<double> tdev(n_ctx, 0.0);
    std::vector<std::thread> workers;
    const int32_t per = std::max<int32_t>(1, threads / (int32_t)n_ctx);
    for (uint32_t k = 0; k < n_ctx; k++)
        workers.emplace_back([&, k] {
            if (cut[k] >= cut[k + 1]) return;
            if (cudaSetDevice(ctxs[k]->device) != cudaSuccess) { rcs[k] = sq_set_error(ctxs[k], SQ_ERR_CUDA, "cudaSetDevice(%d) failed", ctxs[k]->device); return; }
            rcs[k] = unpack_range(ctxs[k], a, man, bound, job, cut[k], cut[k + 1], per, &tdev[k]);
        });
    for (auto &w : workers) w.join();
    cudaSetDevice(ctx->device);
    double t_dev = 0;
    for (uint32_t k = 0; k < n_ctx; k++) {
        t_dev = std::max(t_dev, tdev[k]);
        if (rcs[k] && !rc) { rc = rcs[k]; if (k) snprintf(ctx->err, sizeof ctx->err, "%s", ctxs[k]->err); }
    }
    if (rc) return rc;
    fill_summary(summary, a, total, t0, t_dev);
    return SQ_OK;
}

extern "C" int32_t sq_archive_unpack(sq_ctx *ctx, const char *archive_path, const char *output_dir, int32_t threads, sq_summary *summary) {
    if (!ctx || !archive_path || !output_dir) return SQ_ERR_INVALID_ARG;
    const double t0 = now_s();
    double t_dev = 0;
    if (threads < 1) threads = 1;
    Archive a;
    int32_t rc = open_archive(ctx, archive_path, &a, true);
    if (rc) return rc;
    std::vector<ManifestEntry> man;
    uint64_t total = 0;
    if ((rc = read_manifest(ctx, a, &man, &total))) return rc;

    SQ_T("index scan + manifest parsed");
    // read_chunks (reader.rs:259-314): decode every record, keyed by digest; later records overwrite earlier ones
    struct Decoded { uint64_t off; uint32_t len; };
    std::vector<Decoded> dec(a.records.size());
    std::vector<uint64_t> bound(a.records.size());
    uint64_t total_out = 0;
    bool all_exact = true;
    for (size_t i = 0; i < a.records.size(); i++) {
        const Record &r = a.records[i];
        if (r.orig > (uint64_t)1 << 40) return sq_set_error(ctx, SQ_ERR_INVALID_CHUNK_SIZE, "Invalid chunk size: %llu bytes", (unsigned long long)r.orig);
        if (r.orig > 0xFFFFFFFFull || r.comp > 0xFFFFFFFFull) return sq_set_error(ctx, SQ_ERR_READER, "Error reading from squish: chunk record too large");
        bool ex = false;
        bound[i] = payload_decoded_bound(r.payload, r.comp, r.orig, &ex);
        all_exact = all_exact && ex;
        total_out += (bound[i] + 15) & ~15ull;
    }
    if (all_exact && !getenv("SQ_UNPACK_STORE"))
        return unpack_streaming(ctx, a, man, total, bound, output_dir, threads, summary, t0);
    // Fallback for archives whose frames do not state their size (no writer of this format produces them): decode everything
    // into host memory first, like the reference (reader.rs:268), then assemble the files.
    uint8_t *store = nullptr;
    if (total_out) {
        store = (uint8_t *)malloc(total_out);
        if (!store) return sq_set_error(ctx, SQ_ERR_OTHER, "Unknown error: cannot allocate %llu bytes", (unsigned long long)total_out);
    }
    // a frame decodes on one warp: many frames per call keep the GPU full; pinned staging is sized to the job
    uint64_t total_comp = 0;
    for (auto &r : a.records) total_comp += (r.comp + 15) & ~15ull;
    const uint64_t batch_out = std::min<uint64_t>(4096ull << 20, std::max<uint64_t>(total_out, 1u << 20)),
                   batch_in = std::min<uint64_t>(2048ull << 20, std::max<uint64_t>(total_comp, 1u << 20));
    // Two staging pairs drive the two-slot unpack pipeline (sq_unpack_submit / sq_unpack_wait): while batch k decodes, the
    // payloads of batch k+1 are gathered and uploaded and the output of batch k-1 is downloaded and copied into the store.
    struct UnpackBatch {
        Staging comp, out;
        std::vector<sq_frame> frames;
        std::vector<sq_frame_result> fres;
        size_t first = 0; uint64_t dof = 0, store_off = 0;
        sq_ticket *ticket = nullptr; bool live = false;
    } ub[2];
    const bool pin = total_out >= kPinThreshold;
    for (int k = 0; k < 2 && !rc; k++) {
        if ((rc = ub[k].comp.alloc(ctx, batch_in + (4u << 20), pin))) break;
        if ((rc = ub[k].out.alloc(ctx, batch_out + (4u << 20), pin))) break;
        ub[k].fres.resize(ctx->max_batch);
    }
    if (rc) { for (int k = 0; k < 2; k++) { ub[k].comp.release(ctx); ub[k].out.release(ctx); } free(store); return rc; }
    SQ_T("unpack buffers allocated");
    auto finish = [&](UnpackBatch &b) -> int32_t {  // wait for the batch, check every payload, move its bytes into the store
        b.live = false;
        double td = now_s();
        int32_t r = sq_unpack_wait(ctx, b.ticket);
        t_dev += now_s() - td;
        if (r) return r;
        for (size_t k = 0; k < b.frames.size(); k++) {
            if (b.fres[k].status != SQ_OK) return sq_set_error(ctx, SQ_ERR_READER, "Error reading from squish: chunk %zu failed to decode", b.first + k);
            dec[b.first + k] = {b.store_off + b.frames[k].dst_off, b.fres[k].out_len};
        }
        parallel_for(b.frames.size(), threads, [&](size_t k) { memcpy(store + b.store_off + b.frames[k].dst_off, (uint8_t *)b.out.p + b.frames[k].dst_off, b.fres[k].out_len); });
        return SQ_OK;
    };
    uint64_t store_off = 0;
    size_t i = 0;
    int cur = 0;
    while (i < a.records.size() && !rc) {
        UnpackBatch &b = ub[cur];
        if (b.live && (rc = finish(b))) break;  // the slot's previous batch must be home before its buffers are reused
        b.frames.clear();
        uint64_t so = 0, dof = 0;
        b.first = i;
        while (i < a.records.size() && b.frames.size() < ctx->max_batch) {
            const Record &r = a.records[i];
            uint64_t sneed = (r.comp + 15) & ~15ull, dneed = (bound[i] + 15) & ~15ull;
            if (!b.frames.empty() && (so + sneed > batch_in || dof + dneed > batch_out)) break;
            if (sneed > batch_in + (4u << 20) || dneed > batch_out + (4u << 20)) { rc = sq_set_error(ctx, SQ_ERR_INVALID_CHUNK_SIZE, "Invalid chunk size: %llu bytes", (unsigned long long)r.orig); break; }
            b.frames.push_back({so, dof, (uint32_t)r.comp, (uint32_t)bound[i]});
            so += sneed; dof += dneed;
            i++;
        }
        if (rc) break;
        parallel_for(b.frames.size(), threads, [&](size_t k) { memcpy((uint8_t *)b.comp.p + b.frames[k].src_off, a.records[b.first + k].payload, b.frames[k].src_len); });
        double td = now_s();
        rc = sq_unpack_submit(ctx, b.comp.p, so, b.frames.data(), (uint32_t)b.frames.size(), b.out.p, dof, b.fres.data(), &b.ticket);
        t_dev += now_s() - td;
        if (rc) break;
        b.live = true; b.dof = dof; b.store_off = store_off;
        store_off += dof;
        cur ^= 1;
    }
    for (int k = 0; k < 2; k++) {  // drain, oldest first; on error still wait before the buffers go away
        UnpackBatch &b = ub[cur ^ k];
        if (!b.live) continue;
        if (rc) { sq_unpack_wait(ctx, b.ticket); b.live = false; }
        else rc = finish(b);
    }
    for (int k = 0; k < 2; k++) { ub[k].comp.release(ctx); ub[k].out.release(ctx); }
    SQ_T("all chunks decoded");
    if (rc) { free(store); return rc; }

    struct Key { uint64_t a, b; bool operator==(const Key &o) const { return a == o.a && b == o.b; } };
    struct KeyHash { size_t operator()(const Key &k) const { return (size_t)(k.a ^ (k.b * 0x9E3779B97F4A7C15ULL)); } };
    std::unordered_map<Key, size_t, KeyHash> map;
    map.reserve(a.records.size() * 2);
    for (size_t k = 0; k < a.records.size(); k++) map[{get64(a.records[k].digest), get64(a.records[k].digest + 8)}] = k;

    // rebuild_files (reader.rs:316-413)
    std::string outdir(output_dir);
    mkdir(outdir.c_str(), 0777);
    std::atomic<int32_t> err{0};
    std::atomic<size_t> err_idx{0};
    parallel_for(man.size(), threads, [&](size_t fi) {
        const ManifestEntry &e = man[fi];
        std::string full = outdir + "/" + std::string((const char *)e.path, e.path_len);
        for (size_t p = outdir.size() + 1; p < full.size(); p++)
            if (full[p] == '/') { full[p] = 0; mkdir(full.c_str(), 0777); full[p] = '/'; }  // create_dir_all(parent)
        int fd = open(full.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
        if (fd < 0) { err = SQ_ERR_CREATE_FILE; err_idx = fi; return; }
        for (uint32_t c = 0; c < e.chunk_count; c++) {
            auto it = map.find({get64(e.hashes + (size_t)c * 16), get64(e.hashes + (size_t)c * 16 + 8)});
            if (it == map.end()) { err = SQ_ERR_MISSING_CHUNK; err_idx = fi; break; }  // reader.rs:397-401
            const Decoded &d = dec[it->second];
            size_t w = 0;
            while (w < d.len) {
                ssize_t r = write(fd, store + d.off + w, d.len - w);
                if (r <= 0) { err = SQ_ERR_IO; err_idx = fi; break; }
                w += (size_t)r;
            }
        }
        close(fd);
    });
    SQ_T("files rebuilt");
    free(store);
    if (err) {
        const ManifestEntry &e = man[err_idx];
        return sq_set_error(ctx, err, "%s `%.*s`", sq_strerror(err), (int)e.path_len, (const char *)e.path);
    }
    if (summary) {
        memset(summary, 0, sizeof *summary);
        summary->unique_chunks = a.nchunks; summary->total_original_size = total; summary->archive_size = a.size;
        summary->timestamp = a.timestamp; summary->file_count = a.file_count;
        summary->compression_ratio = total ? (double)a.size / (double)total * 100.0 : 0.0;
        snprintf(summary->version, sizeof summary->version, "%s", a.version);
        summary->seconds_total = now_s() - t0; summary->seconds_device = t_dev;
    }
    return SQ_OK;
}
