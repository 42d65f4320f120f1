// host_pack.cpp — encode() (src/lib/mod.rs:49-61) for a contiguous STREAM of symbols on the host, 32 per AVX2 step.
// When the rows of a batch lie back to back and L is a multiple of 8, the packed batch is one stream as well: byte j of it
// holds the masks of symbols 2j (low nibble) and 2j + 1 (high nibble), which is the width-4 BitEnc layout read as
// little-endian u32 words (bitenc.rs:311-322).  A/C/G/T/N in either case are translated by two byte shuffles keyed by the
// symbol's low nibble (one gives the mask, the other the letter it must be); a 32-byte step with any other byte in it —
// IUPAC codes, '.', 'U', junk — goes through the caller's 256-entry table instead, so the result is encode()'s for
// every byte value.  Plain C++ (no CUDA): compiled by the host compiler, selected at run time (fq::have_avx2()).
#include <cstddef>
#include <cstdint>

#if defined(__x86_64__) && defined(__GNUC__)
#include <immintrin.h>
#define FQ_X86_AVX2 1
#endif

namespace fq {

bool have_avx2() {
#ifdef FQ_X86_AVX2
    return __builtin_cpu_supports("avx2");
#else
    return false;
#endif
}

static inline void pack_scalar(const uint8_t* in, uint64_t n_bytes, uint8_t* out, const uint8_t* lut) {
    for (uint64_t j = 0; j + 1 < n_bytes; j += 2) out[j >> 1] = (uint8_t)(lut[in[j]] | (lut[in[j + 1]] << 4));
}

#ifdef FQ_X86_AVX2
__attribute__((target("avx2"))) void pack_stream_avx2(const uint8_t* in, uint64_t n_bytes, uint8_t* out, const uint8_t* lut) {
    const __m256i low_nibble = _mm256_set1_epi8(0x0F);
    const __m256i upper_case = _mm256_set1_epi8((char)0xDF);
    // keyed by the low nibble of the byte: 'A' 0x41, 'C' 0x43, 'T' 0x54, 'G' 0x47, 'N' 0x4E (and their lower-case forms)
    const __m256i mask_of = _mm256_broadcastsi128_si256(_mm_setr_epi8(0, 1, 0, 2, 8, 0, 0, 4, 0, 0, 0, 0, 0, 0, 15, 0));
    const __m256i letter_of = _mm256_broadcastsi128_si256(
        _mm_setr_epi8(-1, 'A', -1, 'C', 'T', -1, -1, 'G', -1, -1, -1, -1, -1, -1, 'N', -1));
    const __m256i weights = _mm256_set1_epi16(0x1001);  // even byte x 1 + odd byte x 16
    // the words are read next by a DMA engine, not by this core: when the destination is 16-byte aligned they are written
    // with non-temporal stores (no read-for-ownership of the destination lines: a third less host memory traffic per read)
    const bool nt = (reinterpret_cast<uintptr_t>(out) & 15u) == 0u;
    uint64_t i = 0;
    for (; i + 32 <= n_bytes; i += 32) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + i));
        const __m256i key = _mm256_and_si256(v, low_nibble);
        const __m256i letter = _mm256_shuffle_epi8(letter_of, key);
        const __m256i ok = _mm256_cmpeq_epi8(_mm256_and_si256(v, upper_case), letter);
        if (_mm256_movemask_epi8(ok) != -1) {  // something other than A/C/G/T/N here: the table knows every byte
            pack_scalar(in + i, 32, out + (i >> 1), lut);
            continue;
        }
        const __m256i mask = _mm256_shuffle_epi8(mask_of, key);
        const __m256i pairs = _mm256_maddubs_epi16(mask, weights);           // 16 x (lo | hi << 4) in 16-bit lanes
        const __m256i bytes = _mm256_packus_epi16(pairs, pairs);             // per 128-bit lane: its 8 bytes, twice
        const __m128i both = _mm256_castsi256_si128(_mm256_permute4x64_epi64(bytes, 0x08));  // lane 0's, then lane 1's
        if (nt)
            _mm_stream_si128(reinterpret_cast<__m128i*>(out + (i >> 1)), both);
        else
            _mm_storeu_si128(reinterpret_cast<__m128i*>(out + (i >> 1)), both);
    }
    if (nt) _mm_sfence();
    pack_scalar(in + i, n_bytes - i, out + (i >> 1), lut);
}
#endif

// n_bytes even; out gets n_bytes / 2 bytes; lut[b] = encode_byte(b)
void pack_stream(const uint8_t* in, uint64_t n_bytes, uint8_t* out, const uint8_t* lut) {
#ifdef FQ_X86_AVX2
    if (have_avx2()) {
        pack_stream_avx2(in, n_bytes, out, lut);
        return;
    }
#endif
    pack_scalar(in, n_bytes, out, lut);
}

}  // namespace fq
