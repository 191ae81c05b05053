// Host-side 2-bit packer behind rbg_pack_bytes (include/rowbowt_gpu.h): the stream pack_kernel (kernels.cu)
// produces on the device, made by the FASTQ parser threads instead so that 46 bytes per 150 bp read cross PCIe
// instead of 158.  Pure CPU code: AVX2 + BMI2 when the host has them (32 bases per
// iteration), a byte-wise table otherwise and for any 32-byte group that holds a byte without a 2-bit code.
#include "host_pack.hpp"

#include <immintrin.h>

#include <cstring>

namespace rbg {
namespace {

inline uint64_t owner_of(const uint64_t* offs, uint64_t n_reads, uint64_t x) {      // last read with offs[i] <= x
    uint64_t lo = 0, hi = n_reads;                                                    // invariant offs[lo] <= x < offs[hi]
    while (hi - lo > 1) {
        const uint64_t mid = (lo + hi) >> 1;
        if (offs[mid] <= x) lo = mid; else hi = mid;
    }
    return lo;
}

// one word, byte by byte; bytes at or beyond `limit` count as code 0 and flag nothing
inline uint64_t pack_word_scalar(const int8_t* code_of, const uint8_t* bases, const uint64_t* offs, uint64_t n_reads,
                                 uint64_t x0, uint64_t limit, uint8_t* flags, uint64_t& exotic) {
    uint64_t out = 0;
    for (uint64_t i = 0; i < 32 && x0 + i < limit; ++i) {
        const int code = code_of[bases[x0 + i]];
        if (code >= 0 && code < 4) {
            out |= (uint64_t) code << (2 * i);
        } else {
            const uint64_t r = owner_of(offs, n_reads, x0 + i);
            __atomic_fetch_or(flags + r, (uint8_t) (code == 4 ? 2 : 1), __ATOMIC_RELAXED);      // RBG_READ_EXOTIC : RBG_READ_DEAD
            if (code == 4) ++exotic;
        }
    }
    return out;
}

__attribute__((target("avx2,bmi2")))
uint64_t pack_avx2(const int8_t* code_of, const uint8_t* bases, const uint64_t* offs, uint64_t n_reads,
                   uint64_t x0, uint64_t x1, uint64_t* packed, uint8_t* flags) {
    uint64_t exotic = 0;
    const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T');
    const __m256i three = _mm256_set1_epi8(3);
    uint64_t x = x0;
    for (; x + 32 <= x1; x += 32) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(bases + x));
        const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, cA), _mm256_cmpeq_epi8(v, cC)),
                                           _mm256_or_si256(_mm256_cmpeq_epi8(v, cG), _mm256_cmpeq_epi8(v, cT)));
        if ((uint32_t) _mm256_movemask_epi8(ok) != 0xFFFFFFFFu) {
            packed[x >> 5] = pack_word_scalar(code_of, bases, offs, n_reads, x, x1, flags, exotic);
            continue;
        }
        // A=0x41 C=0x43 G=0x47 T=0x54: ((c >> 1) ^ (c >> 2)) & 3 = 0,1,2,3 (the bits a 16-bit shift drags in land above bit 1)
        const __m256i t = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(v, 1), _mm256_srli_epi16(v, 2)), three);
        const uint64_t m = 0x0303030303030303ull;
        const uint64_t w = _pext_u64((uint64_t) _mm256_extract_epi64(t, 0), m) | _pext_u64((uint64_t) _mm256_extract_epi64(t, 1), m) << 16 |
                           _pext_u64((uint64_t) _mm256_extract_epi64(t, 2), m) << 32 | _pext_u64((uint64_t) _mm256_extract_epi64(t, 3), m) << 48;
        packed[x >> 5] = w;
    }
    if (x < x1) packed[x >> 5] = pack_word_scalar(code_of, bases, offs, n_reads, x, x1, flags, exotic);
    return exotic;
}

uint64_t pack_scalar(const int8_t* code_of, const uint8_t* bases, const uint64_t* offs, uint64_t n_reads,
                     uint64_t x0, uint64_t x1, uint64_t* packed, uint8_t* flags) {
    uint64_t exotic = 0;
    for (uint64_t x = x0; x < x1; x += 32) packed[x >> 5] = pack_word_scalar(code_of, bases, offs, n_reads, x, x1, flags, exotic);
    return exotic;
}

}  // namespace

uint64_t pack_bytes_host(const int8_t* code_of, const uint8_t* bases, const uint64_t* offs, uint64_t n_reads,
                         uint64_t x0, uint64_t x1, uint64_t* packed, uint8_t* flags) {
    if (x1 <= x0) return 0;
    // the vector path hard-codes A,C,G,T -> 0..3: only when all four are symbols of this index
    const bool plain = code_of[(uint8_t) 'A'] == 0 && code_of[(uint8_t) 'C'] == 1 && code_of[(uint8_t) 'G'] == 2 && code_of[(uint8_t) 'T'] == 3;
    static const bool simd = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2");
    if (plain && simd) return pack_avx2(code_of, bases, offs, n_reads, x0, x1, packed, flags);
    return pack_scalar(code_of, bases, offs, n_reads, x0, x1, packed, flags);
}

}  // namespace rbg
