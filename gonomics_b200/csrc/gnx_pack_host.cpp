// gnx_pack_host.cpp -- dnaTwoBit.NewTwoBit of uniform batches on the host (the staging pass of gnx_affine_batch and
// gnx_pack_twobit_host).  Reference layout: dna/dnaTwoBit/dnaTwoBit.go:28-42 (32 bases per uint64, the first base in
// bits 63:62, the tail word left-aligned).  Plain C++ (compiled by the host compiler, no CUDA): a scalar form and an
// AVX2 form picked at run time.
#include <cstdint>
#include <cstring>
#if defined(__x86_64__) || defined(_M_X64)
#include <immintrin.h>
#define GNX_X86 1
#endif

namespace {

// 32 bases -> one word.  Four bases at a time: t * 0x40100401 drops the four 2-bit fields of a little-endian 32-bit
// load into the product's top byte.
inline uint64_t word_scalar(const uint8_t *b, uint64_t &bad)
{
    uint64_t x[4];
    memcpy(x, b, 32);
    bad |= x[0] | x[1] | x[2] | x[3];
    uint64_t v = 0;
    for (int q = 0; q < 4; ++q) {
        const uint32_t lo = (uint32_t)x[q], hi = (uint32_t)(x[q] >> 32);
        v = (v << 16) | (uint64_t)(((lo * 0x40100401u) >> 24) << 8) | (uint64_t)((hi * 0x40100401u) >> 24);
    }
    return v;
}

inline uint64_t tail_word(const uint8_t *b, int64_t tail, uint64_t &bad)
{
    uint64_t v = 0;
    for (int64_t i = 0; i < tail; ++i) {
        bad |= b[i];
        v |= (uint64_t)(b[i] & 3u) << (62 - 2 * i);
    }
    return v;
}

bool pack_scalar(uint64_t *dst, const uint8_t *src, int64_t count, int64_t len, int64_t wlen)
{
    uint64_t bad = 0;
    const int64_t full = len / 32, tail = len - full * 32;
    for (int64_t p = 0; p < count; ++p) {
        const uint8_t *b = src + p * len;
        uint64_t *w = dst + p * wlen;
        for (int64_t k = 0; k < full; ++k, b += 32)
            w[k] = word_scalar(b, bad);
        if (tail)
            w[full] = tail_word(b, tail, bad);
    }
    return (bad & 0xfcfcfcfcfcfcfcfcull) == 0;
}

#ifdef GNX_X86
// 32 bases per step: maddubs (4 b0 + b1) -> 16 four-bit values, madd (16 p0 + p1) -> 8 bytes in 8 int32 lanes, one byte
// shuffle per 128-bit half puts them in big-endian order.
__attribute__((target("avx2"))) bool pack_avx2(uint64_t *dst, const uint8_t *src, int64_t count, int64_t len, int64_t wlen)
{
    const __m256i k41 = _mm256_set1_epi16(0x0104);      // bytes (4, 1): first base of a pair is the high one
    const __m256i k161 = _mm256_set1_epi32(0x00010010); // int16 (16, 1)
    const __m256i shuf = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1,
                                          -1, -1, -1, -1, -1);
    __m256i badv = _mm256_setzero_si256();
    uint64_t bad = 0;
    const int64_t full = len / 32, tail = len - full * 32;
    for (int64_t p = 0; p < count; ++p) {
        const uint8_t *b = src + p * len;
        uint64_t *w = dst + p * wlen;
        for (int64_t k = 0; k < full; ++k, b += 32) {
            const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(b));
            badv = _mm256_or_si256(badv, x);
            const __m256i y = _mm256_maddubs_epi16(x, k41);
            const __m256i z = _mm256_madd_epi16(y, k161);
            const __m256i s = _mm256_shuffle_epi8(z, shuf);
            const uint32_t hi = (uint32_t)_mm256_extract_epi32(s, 0), lo = (uint32_t)_mm256_extract_epi32(s, 4);
            w[k] = ((uint64_t)hi << 32) | lo;
        }
        if (tail)
            w[full] = tail_word(b, tail, bad);
    }
    const __m256i m = _mm256_and_si256(badv, _mm256_set1_epi8((char)0xfc));
    return _mm256_testz_si256(m, m) && (bad & 0xfcfcfcfcfcfcfcfcull) == 0;
}
#endif

} // namespace

// false if any base is >= 4 (the words are then unspecified)
extern "C" bool gnx_pack_range_host(uint64_t *dst, const uint8_t *src, int64_t count, int64_t len, int64_t wlen)
{
#ifdef GNX_X86
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    if (have_avx2)
        return pack_avx2(dst, src, count, len, wlen);
#endif
    return pack_scalar(dst, src, count, len, wlen);
}

extern "C" bool gnx_pack_range_host_scalar(uint64_t *dst, const uint8_t *src, int64_t count, int64_t len, int64_t wlen)
{
    return pack_scalar(dst, src, count, len, wlen);
}
