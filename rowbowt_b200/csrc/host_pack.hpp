// Host-side 2-bit packing of read bytes (see host_pack.cpp, rbg_pack_bytes in include/rowbowt_gpu.h).
#pragma once
#include <cstdint>

namespace rbg {

// Batch bytes [x0, x1) (x0 a multiple of 32; x1 a multiple of 32 or the end of the batch) -> packed[x0/32 ..]:
// the base at byte x at bits 2*(x&31) of packed[x>>5].  code_of: byte -> 0..3, 4 = terminator, -1 = no symbol of the
// index.  Bytes without a 2-bit code flag their read (1 = dead, 2 = exotic) with an atomic OR.  Returns the number
// of terminator bytes seen.
uint64_t pack_bytes_host(const int8_t* code_of, const uint8_t* bases, const uint64_t* offs, uint64_t n_reads,
                         uint64_t x0, uint64_t x1, uint64_t* packed, uint8_t* flags);

}  // namespace rbg
