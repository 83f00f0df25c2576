// `squishrs` command line: same subcommands, flags and user-visible strings as the reference CLI
// (reference src/cmd/mod.rs:11-58, src/lib.rs:19-111, src/main.rs:5-10), over the C ABI.
#include <unistd.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <map>
#include <string>
#include <vector>

#include "squish_b200.h"

static double now_s() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }
static void tmark(const char *what, double t0) { if (getenv("SQ_TIMING")) fprintf(stderr, "[cli %8.3f] %s\n", now_s() - t0, what); }

static std::string format_bytes(uint64_t b) {  // byte-unit Decimal, "{:.2} {unit}" (cmd/mod.rs:161-165)
    const char *u[] = {"B", "KB", "MB", "GB", "TB", "PB", "EB"};
    double v = (double)b;
    int i = 0;
    while (v >= 1000.0 && i < 6) { v /= 1000.0; i++; }
    char buf[64];
    snprintf(buf, sizeof buf, "%.2f %s", v, u[i]);
    return buf;
}
static std::string with_commas(uint64_t v) {  // num_format Locale::en
    std::string s = std::to_string(v);
    for (int i = (int)s.size() - 3; i > 0; i -= 3) s.insert((size_t)i, ",");
    return s;
}
static std::string table(const std::string &title, const std::vector<std::pair<std::string, std::string>> &rows, bool two_col_title) {
    size_t w0 = 0, w1 = 0;
    for (auto &r : rows) { w0 = std::max(w0, r.first.size()); w1 = std::max(w1, r.second.size()); }
    std::string t0 = title, t1;
    if (two_col_title) { size_t bar = title.find('|'); t0 = title.substr(0, bar); t1 = title.substr(bar + 1); w0 = std::max(w0, t0.size()); w1 = std::max(w1, t1.size()); }
    else if (title.size() > w0 + w1 + 3) w1 = title.size() - w0 - 3;
    std::string sep = "+" + std::string(w0 + 2, '-') + "+" + std::string(w1 + 2, '-') + "+\n", out = sep;
    if (two_col_title) out += "| " + t0 + std::string(w0 - t0.size(), ' ') + " | " + t1 + std::string(w1 - t1.size(), ' ') + " |\n";
    else out += "| " + title + std::string(w0 + w1 + 3 - title.size(), ' ') + " |\n";
    out += sep;
    for (auto &r : rows) out += "| " + r.first + std::string(w0 - r.first.size(), ' ') + " | " + r.second + std::string(w1 - r.second.size(), ' ') + " |\n";
    return out + sep;
}
static void usage() {
    fprintf(stderr,
            "Compact, compress, and deduplicate files into a single archive\n\n"
            "Usage: squishrs [OPTIONS] <COMMAND>\n\nCommands:\n  pack    Pack a directory\n  list    List files in an archive\n"
            "  unpack  Extract archive contents\n\nOptions:\n  -j, --max-threads <MAX_THREADS>  [default: 25]\n"
            "      --device <N>                 CUDA device ordinal [default: 0]\n"
            "      --devices <K>                use K GPUs of this box, ordinals device .. device+K-1 [default: 1]\n");
}
static int fail(const char *what) {
    fprintf(stderr, "\033[31mError: %s\033[0m\n", what);  // main.rs:6-8
    return 1;
}

int main(int argc, char **argv) {
    int threads = 25, device = 0, devices = 1;  // cmd/mod.rs:16
    bool simple = false;
    std::string cmd, arg, output;
    bool have_out = false;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        if ((a == "-j" || a == "--max-threads") && i + 1 < argc) threads = atoi(argv[++i]);
        else if (a == "--device" && i + 1 < argc) device = atoi(argv[++i]);
        else if (a == "--devices" && i + 1 < argc) devices = atoi(argv[++i]);
        else if ((a == "-o" || a == "--output") && i + 1 < argc) { output = argv[++i]; have_out = true; }
        else if (a == "--simple") simple = true;
        else if (a == "-h" || a == "--help") { usage(); return 0; }
        else if (cmd.empty()) cmd = a;
        else if (arg.empty()) arg = a;
        else { usage(); return 2; }
    }
    if (cmd.empty() || arg.empty() || (cmd != "pack" && cmd != "list" && cmd != "unpack")) { usage(); return 2; }

    if (cmd == "list") {  // lib.rs:57-85
        sq_summary s;
        char *listing = nullptr;
        int32_t rc = sq_archive_list(arg.c_str(), &s, &listing);
        if (rc) return fail(sq_last_error(nullptr)[0] ? sq_last_error(nullptr) : sq_strerror(rc));
        std::vector<std::pair<uint64_t, std::string>> files;
        // "size path\n" records.  A stored path may itself contain '\n' (legal on Linux): a record ends at the first newline
        // that is followed by the end of the listing or by the next record's "digits space".
        for (char *p = listing; p && *p;) {
            char *sp = strchr(p, ' ');
            if (!sp) break;
            char *nl = sp;
            for (;;) {
                nl = strchr(nl + 1, '\n');
                if (!nl) { nl = sp + strlen(sp); break; }
                const char *q = nl + 1;
                if (!*q) break;
                const char *d = q;
                while (*d >= '0' && *d <= '9') d++;
                if (d > q && *d == ' ') break;
            }
            files.push_back({strtoull(p, nullptr, 10), std::string(sp + 1, nl)});
            p = *nl ? nl + 1 : nl;
        }
        sq_free(listing);
        if (simple) {
            printf("squish_size(bytes): %llu, original_size(bytes): %llu, compression ratio: %.2f%%, number_of_files: %zu, chunks_count: %llu\n",
                   (unsigned long long)s.archive_size, (unsigned long long)s.total_original_size, s.compression_ratio, files.size(),
                   (unsigned long long)s.unique_chunks);
            printf("%10s  File Path\n", "Size (Bytes)");
            printf("----------  --------------------\n");
            for (auto &f : files) printf("%10llu  %s\n", (unsigned long long)f.first, f.second.c_str());
        } else {  // build_list_summary_table (cmd/mod.rs:97-158)
            char date[64];
            time_t t = (time_t)s.timestamp;
            struct tm tmv;
            localtime_r(&t, &tmv);
            strftime(date, sizeof date, "%H:%M %d/%m/%Y", &tmv);  // header.rs:87-95
            char ratio[32];
            snprintf(ratio, sizeof ratio, "%.1f%%", s.compression_ratio);
            printf("\nSquash breakdown:\n%s", table("Squash Summary",
                                                    {{"Creation Date (UTC)", date}, {"Squish Version", s.version},
                                                     {"Compressed size", format_bytes(s.archive_size)},
                                                     {"Original size", format_bytes(s.total_original_size)}, {"Compression Ratio", ratio},
                                                     {"Number of files", with_commas(files.size())}, {"Number of chunks", with_commas(s.unique_chunks)}},
                                                    false).c_str());
            std::map<std::string, size_t> dirs;
            for (auto &f : files) dirs[f.second.substr(0, f.second.find('/'))]++;
            std::vector<std::pair<std::string, size_t>> dv(dirs.begin(), dirs.end());
            std::stable_sort(dv.begin(), dv.end(), [](auto &x, auto &y) { return x.second > y.second; });
            std::vector<std::pair<std::string, std::string>> rows;
            for (auto &d : dv) rows.push_back({d.first, with_commas(d.second)});
            printf("\nTop-level directory breakdown:\n%s\n", table("Directory|File Count", rows, true).c_str());
        }
        return 0;
    }

    const double t_start = now_s();
    // Driver initialisation touches every visible GPU (seconds on an 8-GPU box): show it only the devices this command uses.
    if (!getenv("CUDA_VISIBLE_DEVICES") && device >= 0 && devices >= 1) {
        std::string vis;
        for (int k = 0; k < devices; k++) vis += (k ? "," : "") + std::to_string(device + k);
        setenv("CUDA_VISIBLE_DEVICES", vis.c_str(), 1);
        device = 0;
    }
    sq_config cfg = {device, 0, 1ull << 22, 4096, getenv("SQ_DETERMINISTIC") ? SQ_FLAG_DETERMINISTIC : 0u};  // SQ_DETERMINISTIC=1: byte-reproducible archives
    sq_ctx *ctx = nullptr;
    int32_t rc = sq_create(&cfg, &ctx);
    tmark("sq_create done", t_start);
    if (rc) return fail(sq_last_error(nullptr));
    std::vector<sq_ctx *> ctxs{ctx};
    for (int k = 1; k < devices; k++) {  // one context per GPU, all owned by this process
        sq_config ck = cfg;
        ck.device = device + k;
        sq_ctx *c = nullptr;
        if ((rc = sq_create(&ck, &c))) return fail(sq_last_error(nullptr));
        ctxs.push_back(c);
    }
    if (cmd == "pack") {  // lib.rs:26-56
        if (!have_out) output = arg + ".squish";  // default uses the UNtrimmed input (lib.rs:31)
        sq_pack_report rep;
        rc = sq_archive_pack_multi(ctxs.data(), (uint32_t)ctxs.size(), arg.c_str(), output.c_str(), threads, &rep);
        if (rc) { std::string m = sq_last_error(ctx); sq_destroy(ctx); return fail(m.c_str()); }
        const char *shown = output.rfind("./", 0) == 0 ? output.c_str() + 2 : output.c_str();
        printf("\033[32mPacking complete!\033[0m\nCompressed to %s\n\033[34mFinal archive size\033[0m: %s\n", shown, format_bytes(rep.archive_size).c_str());
    } else {  // unpack, lib.rs:86-107
        if (!have_out) {
            output = arg;
            const std::string suf = ".squish";
            if (output.size() >= suf.size() && output.compare(output.size() - suf.size(), suf.size(), suf) == 0) output.resize(output.size() - suf.size());
        }
        sq_summary s;
        rc = sq_archive_unpack_multi(ctxs.data(), (uint32_t)ctxs.size(), arg.c_str(), output.c_str(), threads, &s);
        if (rc) { std::string m = sq_last_error(ctx); sq_destroy(ctx); return fail(m.c_str()); }
        printf("\033[32mUnpacking complete!\033[0m\n%s was unsquished into /%s\n", arg.c_str(), output.c_str());
    }
    tmark("command done", t_start);
    // The job is done and every output file is closed: leave without tearing the CUDA context down call by call (freeing the
    // scratch and destroying streams costs a few tenths of a second; the driver reclaims everything at process exit anyway).
    fflush(stdout);
    fflush(stderr);
    if (!getenv("SQ_CLEAN_EXIT")) _exit(0);
    sq_destroy(ctx);
    tmark("sq_destroy done", t_start);
    return 0;
}
