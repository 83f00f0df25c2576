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
