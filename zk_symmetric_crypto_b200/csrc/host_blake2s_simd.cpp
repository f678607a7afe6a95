// Host Blake2s-256 compression, row-parallel 128-bit SIMD (RFC 7693 section 3.2 with the four column / diagonal G functions of
// a round in the four lanes of one vector), runtime-dispatched: AVX-512VL (single-instruction rotates) > SSSE3 > scalar caller.
// The host hashes the Fiat-Shamir transcript (the 33,288 sampled values of a ChaCha proof are one 532 KB message: ~10 % of a
// product-size proof) and the two public-input hashes of the whole plaintext / ciphertext
// (/root/reference/stwo/src/chacha/bitwise/air_stream.rs:44-53; upstream core/channel/blake2s.rs mix_felts).
// Plain C++ translation unit (g++), so the vector intrinsics never pass through nvcc's front end.
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

namespace {

const uint32_t IV[8] = {0x6A09E667u, 0xBB67AE85u, 0x3C6EF372u, 0xA54FF53Au, 0x510E527Fu, 0x9B05688Cu, 0x1F83D9ABu, 0x5BE0CD19u};
struct RotAvx512 {
    template <int N>
    __attribute__((target("avx512f,avx512vl"), always_inline)) static inline __m128i ror(__m128i x) { return _mm_ror_epi32(x, N); }
};
struct RotSsse3 {
    template <int N>
    __attribute__((target("ssse3"), always_inline)) static inline __m128i ror(__m128i x) {
        if (N == 16) return _mm_shuffle_epi8(x, _mm_set_epi8(13, 12, 15, 14, 9, 8, 11, 10, 5, 4, 7, 6, 1, 0, 3, 2));
        if (N == 8) return _mm_shuffle_epi8(x, _mm_set_epi8(12, 15, 14, 13, 8, 11, 10, 9, 4, 7, 6, 5, 0, 3, 2, 1));
        return _mm_or_si128(_mm_srli_epi32(x, N), _mm_slli_epi32(x, 32 - N));
    }
};

// one round with compile-time message indices (the compiler turns the four message vectors into a few shuffles of the loaded block)
#define B2S_SIMD_ROUND(ROT, s0, s1, s2, s3, s4, s5, s6, s7, s8, s9, s10, s11, s12, s13, s14, s15)                         \
    {                                                                                                                     \
        __m128i mx = _mm_set_epi32((int)m[s6], (int)m[s4], (int)m[s2], (int)m[s0]);                                       \
        __m128i my = _mm_set_epi32((int)m[s7], (int)m[s5], (int)m[s3], (int)m[s1]);                                       \
        a = _mm_add_epi32(_mm_add_epi32(a, b), mx); d = ROT::template ror<16>(_mm_xor_si128(d, a));                       \
        c = _mm_add_epi32(c, d); b = ROT::template ror<12>(_mm_xor_si128(b, c));                                          \
        a = _mm_add_epi32(_mm_add_epi32(a, b), my); d = ROT::template ror<8>(_mm_xor_si128(d, a));                        \
        c = _mm_add_epi32(c, d); b = ROT::template ror<7>(_mm_xor_si128(b, c));                                           \
        b = _mm_shuffle_epi32(b, _MM_SHUFFLE(0, 3, 2, 1));                                                                \
        c = _mm_shuffle_epi32(c, _MM_SHUFFLE(1, 0, 3, 2));                                                                \
        d = _mm_shuffle_epi32(d, _MM_SHUFFLE(2, 1, 0, 3));                                                                \
        mx = _mm_set_epi32((int)m[s14], (int)m[s12], (int)m[s10], (int)m[s8]);                                            \
        my = _mm_set_epi32((int)m[s15], (int)m[s13], (int)m[s11], (int)m[s9]);                                            \
        a = _mm_add_epi32(_mm_add_epi32(a, b), mx); d = ROT::template ror<16>(_mm_xor_si128(d, a));                       \
        c = _mm_add_epi32(c, d); b = ROT::template ror<12>(_mm_xor_si128(b, c));                                          \
        a = _mm_add_epi32(_mm_add_epi32(a, b), my); d = ROT::template ror<8>(_mm_xor_si128(d, a));                        \
        c = _mm_add_epi32(c, d); b = ROT::template ror<7>(_mm_xor_si128(b, c));                                           \
        b = _mm_shuffle_epi32(b, _MM_SHUFFLE(2, 1, 0, 3));                                                                \
        c = _mm_shuffle_epi32(c, _MM_SHUFFLE(1, 0, 3, 2));                                                                \
        d = _mm_shuffle_epi32(d, _MM_SHUFFLE(0, 3, 2, 1));                                                                \
    }

#define B2S_SIMD_BODY(ROT)                                                                                               \
    __m128i ha = _mm_loadu_si128((const __m128i*)h), hb = _mm_loadu_si128((const __m128i*)(h + 4));                       \
    const __m128i iv0 = _mm_loadu_si128((const __m128i*)IV), iv1 = _mm_loadu_si128((const __m128i*)(IV + 4));             \
    for (size_t blk = 0; blk < nblocks; blk++) {                                                                          \
        uint32_t m[16];                                                                                                   \
        memcpy(m, data + 64 * blk, 64);                                                                                   \
        t += 64;                                                                                                          \
        __m128i a = ha, b = hb, c = iv0;                                                                                  \
        __m128i d = _mm_xor_si128(iv1, _mm_set_epi32(0, 0, (int)(uint32_t)(t >> 32), (int)(uint32_t)t));                  \
        B2S_SIMD_ROUND(ROT, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15)    \
        B2S_SIMD_ROUND(ROT, 14, 10, 4, 8, 9, 15, 13, 6, 1, 12, 0, 2, 11, 7, 5, 3)    \
        B2S_SIMD_ROUND(ROT, 11, 8, 12, 0, 5, 2, 15, 13, 10, 14, 3, 6, 7, 1, 9, 4)    \
        B2S_SIMD_ROUND(ROT, 7, 9, 3, 1, 13, 12, 11, 14, 2, 6, 5, 10, 4, 0, 15, 8)    \
        B2S_SIMD_ROUND(ROT, 9, 0, 5, 7, 2, 4, 10, 15, 14, 1, 11, 12, 6, 8, 3, 13)    \
        B2S_SIMD_ROUND(ROT, 2, 12, 6, 10, 0, 11, 8, 3, 4, 13, 7, 5, 15, 14, 1, 9)    \
        B2S_SIMD_ROUND(ROT, 12, 5, 1, 15, 14, 13, 4, 10, 0, 7, 6, 3, 9, 2, 8, 11)    \
        B2S_SIMD_ROUND(ROT, 13, 11, 7, 14, 12, 1, 3, 9, 5, 0, 15, 4, 8, 6, 2, 10)    \
        B2S_SIMD_ROUND(ROT, 6, 15, 14, 9, 11, 3, 0, 8, 12, 2, 13, 7, 1, 4, 10, 5)    \
        B2S_SIMD_ROUND(ROT, 10, 2, 8, 4, 7, 6, 1, 5, 15, 11, 9, 14, 3, 12, 13, 0)    \
        ha = _mm_xor_si128(ha, _mm_xor_si128(a, c));                                                                      \
        hb = _mm_xor_si128(hb, _mm_xor_si128(b, d));                                                                      \
    }                                                                                                                     \
    _mm_storeu_si128((__m128i*)h, ha);                                                                                    \
    _mm_storeu_si128((__m128i*)(h + 4), hb);

__attribute__((target("avx512f,avx512vl"))) void blocks_avx512(uint32_t h[8], const uint8_t* data, size_t nblocks, uint64_t t) {
    B2S_SIMD_BODY(RotAvx512)
}
__attribute__((target("ssse3"))) void blocks_ssse3(uint32_t h[8], const uint8_t* data, size_t nblocks, uint64_t t) {
    B2S_SIMD_BODY(RotSsse3)
}

}  // namespace

// Compresses `nblocks` full, NON-final 64-byte blocks into h; t = bytes absorbed before them.  Returns 0 when no SIMD path is
// available on this CPU (the caller then runs its scalar loop).
extern "C" int s2c_host_blake2s_blocks(uint32_t h[8], const uint8_t* data, size_t nblocks, uint64_t t) {
    static const int level = [] {
        __builtin_cpu_init();
        if (__builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512vl")) return 2;
        if (__builtin_cpu_supports("ssse3")) return 1;
        return 0;
    }();
    if (level == 2) blocks_avx512(h, data, nblocks, t);
    else if (level == 1) blocks_ssse3(h, data, nblocks, t);
    else return 0;
    return 1;
}
