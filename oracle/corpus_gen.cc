/* Test / bench infrastructure, not product: the synthetic corpus generator on the host, for the CPU reference
 * arm of bench.py.  The generator itself is the header both the product's device kernel and this file compile
 * (squishrs_b200/csrc/corpus.h: pure integer arithmetic, identical bytes on host and device); compiling it here keeps the
 * reference arm from loading the product library. */
#include <stdint.h>
#include <string.h>
#include "../squishrs_b200/csrc/corpus.h"

extern "C" int sqo_corpus_fill(void *h_out, uint64_t len, uint64_t seed, uint64_t payload_id, uint32_t klass) {
    uint8_t *out = (uint8_t *)h_out;
    alignas(8) uint8_t tmp[SQC_PAGE];
    for (uint64_t off = 0, page = 0; off < len; off += SQC_PAGE, page++) {
        uint32_t limit = len - off < SQC_PAGE ? (uint32_t)(len - off) : SQC_PAGE;
        sqc_fill_page(tmp, SQC_PAGE, seed, payload_id, klass, page);
        memcpy(out + off, tmp, limit);
    }
    return 0;
}
