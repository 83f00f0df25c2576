// Host build of the K4 decoder core (zstd_dec_core.h with a one-lane "warp"): test tool, not product path.
//   g++ -O2 -shared -fPIC -I squishrs_b200/csrc tests/harness/dec_model.cc -o tests/harness/libdecmodel.so
#include <stdlib.h>
#include <vector>
#include "zstd_dec_core.h"

extern "C" long dec_model_payload(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t cap) {
    std::vector<uint8_t> padded(n + 64, 0);  // the core may read up to 15 bytes past a stream (device buffers carry the same slack)
    memcpy(padded.data() + 16, src, n);
    zd::Tables *T = new zd::Tables();
    zd::Scratch *S = new zd::Scratch();
    std::vector<uint8_t> lit(Z_BLOCK_MAX + 64);
    memset(T, 0, sizeof *T);
    long r = (long)zd::decode_payload(padded.data() + 16, n, dst, cap, T, S, lit.data());
    delete T; delete S;
    return r;
}

// The two-pass (block-parallel) decoder on the host: scan, table snapshots, pass 1 block by block (literals placed, matches
// stored, repeat offsets symbolic), pass 2 in order.  Returns the decoded size, or -100 if the payload is not eligible or a pass
// hands it to the one-pass decoder (the product then runs decode_payload, which also produces any error).
extern "C" long dec_model_payload_two_pass(const uint8_t *src, uint32_t n, uint8_t *dst, uint32_t cap) {
    std::vector<uint8_t> padded(n + 64, 0);
    memcpy(padded.data() + 16, src, n);
    const uint8_t *p = padded.data() + 16;
    const uint32_t MAXB = 16, SEQ_CAP = Z_BLOCK_MAX / 6 + 8;
    std::vector<zd::BlockTask> tasks(MAXB);
    zd::FrameInfo fi;
    const uint32_t nb = zd::scan_frame(p, n, MAXB, tasks.data(), &fi);
    if (!nb) return -100;
    zd::Tables *T = new zd::Tables();
    zd::Scratch *S = new zd::Scratch();
    memset(T, 0, sizeof *T);
    std::vector<zd::Tables> snaps(nb);
    long r = -100;
    if (!fi.chained || zd::snapshot_frame_tables(p, tasks.data(), nb, T, S, snaps.data())) {
        std::vector<uint8_t> lits(Z_BLOCK_MAX + 64);
        std::vector<zd::StoredSeq> seqs((size_t)nb * SEQ_CAP);
        std::vector<zd::BlockState> states(nb);
        for (uint32_t b = nb; b-- > 0;)  // any order: the blocks do not depend on each other in pass 1 (here: last to first)
            if (tasks[b].type == 2) {
                if (fi.chained) *T = snaps[b];
                zd::decode_block_first_pass(p + tasks[b].src_off, tasks[b].size, fi.chained != 0, T, S, lits.data(), dst, tasks[b].out_start, cap, b == 0,
                                            seqs.data() + (size_t)b * SEQ_CAP, SEQ_CAP, &states[b]);
            }
        r = (long)zd::execute_frame_matches(p, tasks.data(), &fi, states.data(), seqs.data(), SEQ_CAP, dst, cap);
    }
    delete T; delete S;
    return r;
}
