// kmer_math.hpp — k-mer algebra shared by the host builder and the sm_100a kernels.
//
// Every function states the reference definition it must agree with bit for bit
// (citations into the reference tree).  The formulations are ours:
// reverse-complement goes through bit reversal (one BREV on the GPU), and the minimizer
// ordering is carried as a biased 32-bit key whose inverse recovers the m-mer.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define BL_HD __host__ __device__ __forceinline__
#else
#define BL_HD inline
#endif

namespace blight {

// Nucleotide code (c>>1)&3 : A=0 C=1 T=2 G=3, lower case accepted (kmer.h:56-69).
// Returns 4 for any byte the reference would reject with std::domain_error.
BL_HD uint32_t nuc_code(unsigned char c) {
	const uint32_t u = c & 0xDFu;  // fold case
	const bool ok = (u == 'A') | (u == 'C') | (u == 'G') | (u == 'T');
	return ok ? ((c >> 1) & 3u) : 4u;
}

#if defined(__CUDACC__)
// Four ASCII bases at once (one 32-bit word of the text, first base in the low byte): their 2-bit codes packed first base
// high in the low 8 bits of the result, and in bits 8..11 one flag per base (bit 11 = first) for bytes nuc2int rejects.
// Byte-wise SIMD compares instead of 4 x (fold case, 4 compares): the front end's phase A drops from ~10 to ~3
// instructions per base.
__device__ __forceinline__ uint32_t nuc_code4(uint32_t w) {
	const uint32_t u = w & 0xDFDFDFDFu;  // fold case
	const uint32_t ok = __vcmpeq4(u, 0x41414141u) | __vcmpeq4(u, 0x43434343u) | __vcmpeq4(u, 0x47474747u) | __vcmpeq4(u, 0x54545454u);
	const uint32_t c = (w >> 1) & 0x03030303u;
	const uint32_t codes = (c * 0x40100401u) >> 24;                // b0 b1 b2 b3 -> one byte, first base in the high bits
	const uint32_t bad = ((~ok & 0x01010101u) * 0x08040201u) >> 24;  // one bit per byte, first base -> bit 3
	return codes | ((bad & 0xFu) << 8);
}
#endif

BL_HD uint64_t bitrev64(uint64_t x) {
#if defined(__CUDA_ARCH__)
	return __brevll(x);
#else
	x = __builtin_bswap64(x);
	x = ((x & 0x0f0f0f0f0f0f0f0full) << 4) | ((x >> 4) & 0x0f0f0f0f0f0f0f0full);
	x = ((x & 0x3333333333333333ull) << 2) | ((x >> 2) & 0x3333333333333333ull);
	x = ((x & 0x5555555555555555ull) << 1) | ((x >> 1) & 0x5555555555555555ull);
	return x;
#endif
}

BL_HD uint32_t bitrev32(uint32_t x) {
#if defined(__CUDA_ARCH__)
	return __brev(x);
#else
	x = __builtin_bswap32(x);
	x = ((x & 0x0f0f0f0fu) << 4) | ((x >> 4) & 0x0f0f0f0fu);
	x = ((x & 0x33333333u) << 2) | ((x >> 2) & 0x33333333u);
	x = ((x & 0x55555555u) << 1) | ((x >> 1) & 0x55555555u);
	return x;
#endif
}

BL_HD int popc32(uint32_t x) {
#if defined(__CUDA_ARCH__)
	return __popc(x);
#else
	return __builtin_popcount(x);
#endif
}

BL_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
	return __popcll(x);
#else
	return __builtin_popcountll(x);
#endif
}

// Reverse complement of an n-nucleotide word, first base in the high used field.
// Equals rcb(uint64_t,n) (kmer.h:218-232): complement = XOR 0b10 per base, order reversed.
// Bit reversal reverses base order but also swaps the two bits of each base; swap them back.
BL_HD uint64_t rc64(uint64_t x, unsigned n) {
	uint64_t r = bitrev64(x ^ 0xAAAAAAAAAAAAAAAAull);
	r = ((r & 0x5555555555555555ull) << 1) | ((r >> 1) & 0x5555555555555555ull);
	return r >> (64 - 2 * n);
}

// Same for 32-bit words: rcb(uint32_t,n) (kmer.h:236-251).
BL_HD uint32_t rc32(uint32_t x, unsigned n) {
	uint32_t r = bitrev32(x ^ 0xAAAAAAAAu);
	r = ((r & 0x55555555u) << 1) | ((r >> 1) & 0x55555555u);
	return r >> (32 - 2 * n);
}

// ParityCanonical::canonize (kmer.h:475-487), m odd: whichever of x / rc(x) has odd popcount, >> 1.
BL_HD uint32_t parity_canon(uint32_t mmer, unsigned m) {
	return ((popc32(mmer) & 1) ? mmer : rc32(mmer, m)) >> 1;
}

// revhash(uint32_t) (kmer.h:102-108); the reference compares the result as int32_t.
BL_HD uint32_t revhash32(uint32_t x) {
	x = ((x >> 16) ^ x) * 0x2c1b3c6du;
	x = ((x >> 16) ^ x) * 0x297a2d39u;
	x = ((x >> 16) ^ x);
	return x;
}

// Inverse of revhash32 (multiplicative inverses mod 2^32 of the two constants).
BL_HD uint32_t unrevhash32(uint32_t x) {
	x = ((x >> 16) ^ x) * 0x0cf0b109u;
	x = ((x >> 16) ^ x) * 0x64ea2d65u;
	x = ((x >> 16) ^ x);
	return x;
}

// Ordering key of an m-mer: unsigned order of the key == signed order of revhash (kmer.h:798-804).
BL_HD uint32_t mini_key(uint32_t canon_mmer) { return revhash32(canon_mmer) ^ 0x80000000u; }
BL_HD uint32_t mini_from_key(uint32_t key) { return unrevhash32(key ^ 0x80000000u); }

// SingleHashFunctor::hash_bis (bbhash.h:172-185).
BL_HD uint64_t hash_bis(uint64_t key, uint64_t seed) {
	uint64_t h = seed;
	h ^= (h << 7) ^ key * (h >> 3) ^ (~((h << 11) + (key ^ (h >> 5))));
	h = (~h) + (h << 21);
	h = h ^ (h >> 24);
	h = (h + (h << 3)) + (h << 8);
	h = h ^ (h >> 14);
	h = (h + (h << 2)) + (h << 4);
	h = h ^ (h >> 28);
	h = h + (h << 31);
	return h;
}

constexpr uint64_t kSeed0 = 0xAAAAAAAA55555555ull;  // bbhash.h:219-223
constexpr uint64_t kSeed1 = 0x33333333CCCCCCCCull;  // bbhash.h:225-229

// xorshift128* step used for levels >= 2 (bbhash.h:233-239): updates (s0,s1), returns the hash.
BL_HD uint64_t xs128_next(uint64_t& s0, uint64_t& s1) {
	uint64_t a = s0;
	const uint64_t b = s1;
	s0 = b;
	a ^= a << 23;
	s1 = a ^ b ^ (a >> 17) ^ (b >> 26);
	return s1 + b;
}

// fastmod64 (bbhash.h:660-662): high 64 bits of hash * domain.
BL_HD uint64_t mulhi64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
	return __umul64hi(a, b);
#else
	return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// Minimizer of a canonical k-mer, the definition (patched minimizer_naive, kmer.h:791-810):
// over the k-m+1 m-mers of `canon`, the parity-canonical value with the smallest signed revhash.
BL_HD uint32_t minimizer_of_kmer(uint64_t canon, unsigned k, unsigned m) {
	const uint32_t mask = (m == 16) ? 0xFFFFFFFFu : ((1u << (2 * m)) - 1u);
	uint32_t best = 0xFFFFFFFFu;
	for (unsigned i = 0; i + m <= k; i++) {
		const uint32_t key = mini_key(parity_canon((uint32_t)(canon >> (2 * i)) & mask, m));
		best = key < best ? key : best;
	}
	return mini_from_key(best);
}

}  // namespace blight
