// host_pack.cpp — host side of the *_host entry points: ASCII reads -> 2-bit codes before they cross PCIe.
//
// nuc2int (kmer.h:56-69) maps a base to (c >> 1) & 3 (A0 C1 T2 G3, either case) and rejects every other byte. A chunk of
// a read batch packed this way is a quarter of the bytes on the bus; the read kernels then copy the words instead of
// converting them (front.cuh, phase A). A chunk holding any byte nuc2int rejects is NOT packed: it travels as ASCII, so
// that the kernel applies the reference's rule (only bytes under a queried k-mer raise, blight.cpp:782 / kmer.h:68).
// Layout: base p of the text -> bits 30 - 2 (p & 15) .. of word p >> 4 (first base in the high bits), the layout of the
// kernels' shared-memory strips.
#include <immintrin.h>

#include <cstdint>
#include <cstring>

#include "host_pack.hpp"

namespace blight {

namespace {

// 32 bases -> two words; *bad accumulates bytes outside ACGTacgt
inline void pack32(const unsigned char* p, uint32_t* out, __m256i& bad) {
	const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p));
	const __m256i u = _mm256_and_si256(v, _mm256_set1_epi8((char)0xDF));  // fold case
	const __m256i ok = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(u, _mm256_set1_epi8('A')), _mm256_cmpeq_epi8(u, _mm256_set1_epi8('C'))),
	                                   _mm256_or_si256(_mm256_cmpeq_epi8(u, _mm256_set1_epi8('G')), _mm256_cmpeq_epi8(u, _mm256_set1_epi8('T'))));
	bad = _mm256_or_si256(bad, _mm256_xor_si256(ok, _mm256_set1_epi8((char)0xFF)));
	const __m256i c = _mm256_and_si256(_mm256_srli_epi16(v, 1), _mm256_set1_epi8(3));  // codes, one per byte
	const __m256i p2 = _mm256_maddubs_epi16(c, _mm256_set1_epi16(0x0104));             // 4 * even + odd: two bases per 16 bits
	const __m256i p4 = _mm256_madd_epi16(p2, _mm256_set1_epi32(0x00010010));           // 16 * even + odd: four bases per 32 bits
	// the four bytes of a word, first base's byte last (little-endian word, first base in the high bits)
	const __m256i sh = _mm256_setr_epi8(12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
	                                    12, 8, 4, 0, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
	const __m256i w = _mm256_shuffle_epi8(p4, sh);
	out[0] = (uint32_t)_mm256_extract_epi32(w, 0);
	out[1] = (uint32_t)_mm256_extract_epi32(w, 4);
}

}  // namespace

bool pack2_block(const char* text, uint64_t n_bases, uint32_t* words) {
	const unsigned char* p = reinterpret_cast<const unsigned char*>(text);
	__m256i bad = _mm256_setzero_si256();
	uint64_t i = 0;
	for (; i + 32 <= n_bases; i += 32) pack32(p + i, words + (i >> 4), bad);
	bool any_bad = !_mm256_testz_si256(bad, bad);
	if (i < n_bases) {  // tail: missing bases read as 'A'
		unsigned char tmp[32];
		std::memset(tmp, 'A', sizeof tmp);
		std::memcpy(tmp, p + i, size_t(n_bases - i));
		uint32_t w2[2];
		__m256i bad2 = _mm256_setzero_si256();
		pack32(tmp, w2, bad2);
		words[i >> 4] = w2[0];
		if (n_bases - i > 16) words[(i >> 4) + 1] = w2[1];
		any_bad = any_bad || !_mm256_testz_si256(bad2, bad2);
	}
	return !any_bad;
}

}  // namespace blight
