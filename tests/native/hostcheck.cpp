// hostcheck.cpp — TEST INFRASTRUCTURE ONLY: compiles the host-callable halves of the device headers with g++ so that
// CPU tests (no GPU in the build container) can check them against the numpy restatement / the oracle bit for bit.
// Nothing here is part of the product path; libf2q.so never links it.
#include <stdint.h>
#include <string.h>

#include "../../2fast2q_b200/csrc/synth_gen.h"

extern "C" {

__attribute__((visibility("default")))
void hc_synth(const f2q_synth_spec* sp, const uint8_t* guides, uint64_t first, uint64_t n, uint8_t* out) {
    const uint64_t rec = 2ull * sp->read_len + 18;
    for (uint64_t k = 0; k < n; k++) f2q::synth_record(*sp, guides, first + k, out + k * rec);
}

}
