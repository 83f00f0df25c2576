/* ORACLE — test infrastructure only.  Command-line front end over squish_ref.c,
 * mirroring `squishrs [-j N] pack <in> [-o out] | list <a> [--simple] | unpack <a> [-o dir]`
 * (reference src/cmd/mod.rs:11-58, src/lib.rs:19-111).  Used as the timed CPU baseline. */
#include "squish_ref.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

int main(int argc, char **argv) {
    int threads = 25; /* cmd/mod.rs:16 */
    const char *cmd = NULL, *arg = NULL, *out = NULL;
    int simple = 0, pdec = 0;
    for (int i = 1; i < argc; i++) {
        if (!strcmp(argv[i], "-j") || !strcmp(argv[i], "--max-threads")) { if (++i < argc) threads = atoi(argv[i]); }
        else if (!strcmp(argv[i], "-o") || !strcmp(argv[i], "--output")) { if (++i < argc) out = argv[i]; }
        else if (!strcmp(argv[i], "--simple")) simple = 1;
        else if (!strcmp(argv[i], "--parallel-decode")) pdec = 1;
        else if (!cmd) cmd = argv[i];
        else if (!arg) arg = argv[i];
    }
    if (!cmd || !arg) { fprintf(stderr, "usage: refcpu [-j N] pack <dir> [-o out] | list <a> [--simple] | unpack <a> [-o dir]\n"); return 2; }
    char buf[4096];
    int rc;
    if (!strcmp(cmd, "pack")) {
        if (!out) { snprintf(buf, sizeof buf, "%s.squish", arg); out = buf; } /* lib.rs:31 */
        sqo_pack_stats st;
        rc = sqo_pack_dir(arg, out, threads, &st);
        if (!rc) printf("Packing complete!\nCompressed to %s\nFinal archive size: %llu bytes (%llu unique / %llu chunks, %.3f s)\n", out,
                        (unsigned long long)st.archive_size, (unsigned long long)st.unique_chunks, (unsigned long long)st.total_chunks, st.seconds);
    } else if (!strcmp(cmd, "list")) {
        sqo_summary s; char *paths = NULL;
        rc = sqo_list(arg, &s, &paths);
        if (!rc) {
            (void)simple;
            printf("squish_size(bytes): %llu, original_size(bytes): %llu, compression ratio: %.2f%%, number_of_files: %u, chunks_count: %llu\n",
                   (unsigned long long)s.archive_size, (unsigned long long)s.total_original_size, s.compression_ratio, s.file_count,
                   (unsigned long long)s.unique_chunks); /* lib.rs:67-75 */
            fputs(paths, stdout);
            sqo_free(paths);
        }
    } else if (!strcmp(cmd, "unpack")) {
        if (!out) { snprintf(buf, sizeof buf, "%s", arg); char *p = strstr(buf, ".squish"); if (p && !p[7]) *p = 0; out = buf; } /* lib.rs:88-93 */
        sqo_summary s;
        rc = sqo_unpack(arg, out, threads, pdec, &s);
        if (!rc) printf("Unpacking complete!\n%s was unsquished into /%s (decode %.3f s, rebuild %.3f s)\n", arg, out, s.decode_seconds, s.rebuild_seconds);
    } else { fprintf(stderr, "unknown command %s\n", cmd); return 2; }
    if (rc) { fprintf(stderr, "Error: %s\n", sqo_strerror(rc)); return 1; } /* main.rs:6-8 */
    return 0;
}
