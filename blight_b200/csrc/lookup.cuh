// lookup.cuh — the lookup core: one canonical k-mer -> identifier, on the device layout of device_index.hpp.
//
// Follows query_get_hash (blight.cpp:716-742) step for step; what changes is how many sectors each step reads.
//   1. bucket[minimizer]: empty -> -1                                   (blight.cpp:719)        1 x 16 B, L1/L2 resident
//   2. group = minimizer >> lb, its DevMphf                             (blight.cpp:722)        L1 resident
//   3. BBHash levels: hash, mulhi, test bit, rank                       (bbhash.h:561-577)      1 sector per level
//   4. position field, << b, as uint32                                  (blight.cpp:473-482)    1 sector
//   5. guard pos+k-1 < bucket length (first window only)                (blight.cpp:729)
//   6. scan 2^b windows of the bucket sequence for canon(window)==x     (blight.cpp:730-739)    1-2 sectors (b<=6)
//   7. id = rank + id_offset, else -1                                   (blight.cpp:736,741)
//
// L2 residency is steered per array: level bits (small, probed ~2x per k-mer) are loaded evict_last, position
// sectors (large, touched once) evict_first and not allocated in L1, sequences with the default policy.
#pragma once
#include <cuda_fp16.h>

#include <cstdint>

#include "device_index.hpp"
#include "kmer_math.hpp"

namespace blight {

// one level-bits sector: 7 words of bits + ones-before-this-chunk; kept in L2 as long as possible
__device__ __forceinline__ void ld_bits_sector(const uint32_t* p, uint32_t (&w)[8]) {
	asm volatile("ld.global.nc.L2::evict_last.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
	             : "l"(p));
}

// one position sector: touched once per query, so neither L1 nor L2 should keep it
__device__ __forceinline__ void ld_pos_sector(const uint32_t* p, uint32_t (&w)[8]) {
	asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
	             : "l"(p));
}

// one filter block: default L2 policy, not worth an L1 line
__device__ __forceinline__ void ld_filter_sector(const uint32_t* p, uint32_t (&w)[8]) {
	asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
	             : "l"(p));
}

// word j (0..7) of a sector held in registers
__device__ __forceinline__ uint32_t pick8(const uint32_t (&w)[8], uint32_t j) {
	const bool b0 = j & 1, b1 = j & 2, b2 = j & 4;
	const uint32_t s01 = b0 ? w[1] : w[0], s23 = b0 ? w[3] : w[2], s45 = b0 ? w[5] : w[4], s67 = b0 ? w[7] : w[6];
	const uint32_t lo = b1 ? s23 : s01, hi = b1 ? s67 : s45;
	return b2 ? hi : lo;
}

// word j (0..6) of a sector held in registers, without dynamic register indexing
__device__ __forceinline__ uint32_t pick7(const uint32_t (&w)[8], uint32_t j) {
	const bool b0 = j & 1, b1 = j & 2, b2 = j & 4;
	const uint32_t s01 = b0 ? w[1] : w[0], s23 = b0 ? w[3] : w[2], s45 = b0 ? w[5] : w[4];
	const uint32_t lo = b1 ? s23 : s01, hi = b1 ? w[6] : s45;
	return b2 ? hi : lo;
}

// Reference loop form of the 2^b-window scan (blight.cpp:730-739): slide one base at a time. Used for k < 8 or b < 3.
// canon(window) == x  <=>  window == x or window == rx, because x is canonical (x <= rx).
// Returns the index of a matching window, or -1.
static __device__ __noinline__ int32_t scan_windows_loop(const uint32_t* __restrict__ seq, uint64_t P, uint32_t k, uint32_t nwin,
                                                  uint64_t x, uint64_t rx) {
	const uint64_t wi = P >> 4;
	const uint32_t s = 2u * (uint32_t)(P & 15);
	const uint32_t* q = seq + wi;
	uint32_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
	uint32_t top_hi = __funnelshift_l(b, a, s), top_lo = __funnelshift_l(c, b, s);  // the 32 bases starting at P
	const uint32_t sh = 64 - 2 * k;                                                  // window = top >> sh
	q += 3;
	uint32_t prev = c, feed = 0;
	for (uint32_t j = 0; j < nwin; j++) {
		if ((j & 15) == 0) {
			const uint32_t nxt = __ldg(q++);
			feed = __funnelshift_l(nxt, prev, s);
			prev = nxt;
		}
		const uint64_t w = (((uint64_t)top_hi << 32) | top_lo) >> sh;
		if (w == x || w == rx) return (int32_t)j;
		top_hi = __funnelshift_l(top_lo, top_hi, 2);
		top_lo = __funnelshift_l(feed, top_lo, 2);
		feed <<= 2;
	}
	return -1;
}

// the k bases of the packed sequence starting at base P (three word loads)
__device__ __forceinline__ uint64_t window_at(const uint32_t* __restrict__ seq, uint64_t P, uint32_t k) {
	const uint32_t* q = seq + (P >> 4);
	const uint32_t s = 2u * (uint32_t)(P & 15);
	const uint32_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
	return ((((uint64_t)__funnelshift_l(b, a, s)) << 32) | __funnelshift_l(c, b, s)) >> (64 - 2 * k);
}

__device__ __forceinline__ bool window_matches(const uint32_t* __restrict__ seq, uint64_t P, uint32_t k, uint64_t x, uint64_t rx) {
	const uint64_t w = window_at(seq, P, k);
	return w == x || w == rx;
}

// The same scan, branch-free and two windows per instruction: the first 8 bases (16 bits) of every window are
// compared against the first 8 bases of x and of rx with half-word SIMD compares (HSET2); windows are handled in 8 residue
// classes (start offset mod 8), so that in the text shifted by the class offset every candidate prefix is half-word
// aligned. Survivors (the true match, plus ~2^-16 false candidates per window) are verified in full.
// Requires k >= 8 and nwin >= 8 (a power of two). Returns the index of a matching window, or -1.
__device__ __forceinline__ int32_t scan_windows(const uint32_t* __restrict__ seq, uint64_t P, uint32_t k, uint32_t nwin,
                                             uint64_t x, uint64_t rx) {
	const uint32_t* q = seq + (P >> 4);
	const uint32_t s = 2u * (uint32_t)(P & 15);
	// Half-word equality through the fp16x2 compare unit (one HSET2 per two windows and target). Bit 14 (the top
	// exponent bit) is cleared on both sides so no half is ever NaN/Inf: equal bits => equal floats, hence no false
	// negatives; +0/-0 and the dropped bit only add false candidates, which the full verification rejects.
	const uint32_t kNoNan = 0xBFFFBFFFu;
	const uint32_t X2u = ((uint32_t)(x >> (2 * k - 16)) * 0x00010001u) & kNoNan;
	const uint32_t R2u = ((uint32_t)(rx >> (2 * k - 16)) * 0x00010001u) & kNoNan;
	const __half2 X2 = *reinterpret_cast<const __half2*>(&X2u);
	const __half2 R2 = *reinterpret_cast<const __half2*>(&R2u);
	uint32_t wprev = __ldg(q);
	for (uint32_t base = 0; base < nwin; base += 64, q += 4) {
		// A[i] = bases [P + base + 16 i, +16): the text of this 64-window chunk, re-aligned to word boundaries
		const uint32_t nwords = nwin - base >= 64 ? 4 : (nwin - base + 15) / 16;  // words of window starts
		uint32_t A[5];
		#pragma unroll
		for (int i = 0; i < 5; i++) {
			const uint32_t wn = (i <= (int)nwords) ? __ldg(q + i + 1) : 0u;
			A[i] = __funnelshift_l(wn, wprev, s);
			wprev = wn;
		}
		wprev = __ldg(q + 4);
		uint32_t acc0 = 0, acc1 = 0;  // candidate windows: bit (16*(1-h) + r + 8*(i&1)) of acc[i>>1]  <->  window r + 8*(2i+h)
		#pragma unroll
		for (int i = 0; i < 4; i++) {
			if (i < (int)nwords) {
				#pragma unroll
				for (int r = 0; r < 8; r++) {
					const uint32_t Su = (r ? __funnelshift_l(A[i + 1], A[i], 2 * r) : A[i]) & kNoNan;
					const __half2 S = *reinterpret_cast<const __half2*>(&Su);
					const uint32_t c = __heq2_mask(S, X2) | __heq2_mask(S, R2);
					const uint32_t K = 0x00010001u << (r + 8 * (i & 1));
					if (i < 2) acc0 |= c & K; else acc1 |= c & K;
				}
			}
		}
		if (nwin - base == 8) acc0 &= 0xFFFF0000u;  // only the upper half-word of word 0 holds window starts
		#pragma unroll 1
		for (int half = 0; half < 2; half++) {
			uint32_t acc = half ? acc1 : acc0;
			while (acc) {
				const uint32_t bit = 31 - __clz(acc);  // earliest window first is not required: any match gives the same answer
				acc &= ~(1u << bit);
				const uint32_t h = bit >= 16 ? 0u : 1u, bb = bit & 15u;
				const uint32_t j = (bb & 7u) + 8u * (2u * (2u * half + (bb >> 3)) + h);
				if (window_matches(seq, P + base + j, k, x, rx)) return (int32_t)(base + j);
			}
		}
	}
	return -1;
}

// Descriptors of the bucket a k-mer routes to (all L1/L2 resident).
struct BucketRef {
	uint4 bd;                  // start_lo, start_hi, nuc, -
	const DevMphf* M;
	uint4 m0;                  // bits_sector_base, pos_sector_base
	const uint32_t* bits;
};

__device__ __forceinline__ BucketRef load_bucket(const DevIndexView& I, uint32_t mini) {
	BucketRef B;
	B.bd = __ldg(I.bucket + mini);
	B.M = I.mphf + (mini >> I.lb);
	B.m0 = __ldg(reinterpret_cast<const uint4*>(B.M));
	B.bits = I.bits + ((((uint64_t)B.m0.y << 32) | B.m0.x) << 3);
	return B;
}

// BBHash levels [lv_begin, lv_end) (bbhash.h:619-639): first level whose bit is set wins. (s0, s1) is the hasher state
// (h0, h1, then xorshift128*), `off` the first bit of level lv_begin; both are carried so the probe can be resumed.
// On a hit, w holds the sector and r the bit inside its 224-bit chunk.
template <bool SMALL>
__device__ __forceinline__ bool probe_levels(const BucketRef& B, uint64_t x, int lv_begin, int lv_end, uint64_t& s0, uint64_t& s1,
                                             uint64_t& off, uint32_t (&w)[8], uint32_t& r) {
	#pragma unroll 1
	for (int level = lv_begin; level < lv_end; level++) {
		uint64_t h;
		if (level == 0) h = s0 = hash_bis(x, kSeed0);
		else if (level == 1) h = s1 = hash_bis(x, kSeed1);
		else h = xs128_next(s0, s1);
		if (SMALL) {
			const uint32_t dom = __ldg(&B.M->dom32[level]);
			// fastmod64 (bbhash.h:660-662) with a 32-bit domain: hi64(h * dom)
			const uint32_t bit = (uint32_t)off + (uint32_t)(((uint64_t)(uint32_t)(h >> 32) * dom + __umulhi((uint32_t)h, dom)) >> 32);
			const uint32_t chunk = bit / kChunkBits;
			r = bit - chunk * kChunkBits;
			ld_bits_sector(B.bits + ((uint64_t)chunk << 3), w);
			off += dom;
		} else {
			const uint64_t dom = __ldg(&B.M->dom[level]);
			const uint64_t bit = off + __umul64hi(h, dom);
			const uint64_t chunk = bit / kChunkBits;
			r = (uint32_t)(bit - chunk * kChunkBits);
			ld_bits_sector(B.bits + (chunk << 3), w);
			off += dom;
		}
		if ((pick7(w, r >> 5) >> (r & 31)) & 1u) return true;
	}
	return false;
}

// rank of a key that hit bit r of the sector w (bitVector::rank, bbhash.h:467-480): ones before this chunk + ones
// below the bit inside the chunk
__device__ __forceinline__ uint32_t rank_in_sector(const uint32_t (&w)[8], uint32_t r) {
	const uint32_t j = r >> 5, bi = r & 31;
	uint32_t rank = w[7];
	#pragma unroll
	for (int i = 0; i < 7; i++) {
		const uint32_t below = (i < (int)j) ? 0xFFFFFFFFu : ((i == (int)j) ? ((1u << bi) - 1u) : 0u);
		rank += __popc(w[i] & below);
	}
	return rank;
}

// fallback map (bbhash.h:567-575), sorted by key: rank of x, or false
static __device__ __noinline__ bool fallback_rank(const DevIndexView& I, const uint4& m1, const uint4& m2, uint64_t x, uint32_t& rank) {
	const uint64_t fb_off = ((uint64_t)m1.w << 32) | m1.z;
	uint32_t lo = 0, hi = m2.x;
	while (lo < hi) {
		const uint32_t mid = (lo + hi) >> 1;
		if (__ldg(I.fb_keys + fb_off + mid) < x) lo = mid + 1; else hi = mid;
	}
	if (lo >= m2.x || __ldg(I.fb_keys + fb_off + lo) != x) return false;
	rank = (uint32_t)__ldg(I.fb_vals + fb_off + lo);
	return true;
}

// Everything after the level probe: rank (or fallback map), position, guard, window scan, id.
// T_out receives the absolute base position of the window that matched (when the result is >= 0).
__device__ __forceinline__ int64_t finish_lookup(const DevIndexView& I, const BucketRef& B, uint64_t x, bool hit,
                                                 const uint32_t (&w)[8], uint32_t r, uint64_t* T_out = nullptr) {
	const uint4 m1 = __ldg(reinterpret_cast<const uint4*>(B.M) + 1);  // id_offset, fb_off
	const uint4 m2 = __ldg(reinterpret_cast<const uint4*>(B.M) + 2);  // fb_count, nbits, fields_per_sector, fps_magic
	uint32_t rank;
	if (hit) rank = rank_in_sector(w, r);
	else if (!fallback_rank(I, m1, m2, x, rank)) return -1;
	// position field (blight.cpp:473-482): uint32 arithmetic, << b
	const uint32_t nbits = m2.y, fps = m2.z;
	uint32_t psec = __umulhi(rank, m2.w);  // floor(rank / fps) or one less
	uint32_t slot = rank - psec * fps;
	if (slot >= fps) { slot -= fps; psec++; }
	const uint32_t* ps = I.pos + (((((uint64_t)B.m0.w << 32) | B.m0.z) + psec) << 3);
	const uint32_t o = slot * nbits, ow = o >> 5;
	uint32_t pw[8];
	ld_pos_sector(ps, pw);
	uint32_t field = __funnelshift_r(pick8(pw, ow), pick8(pw, (ow + 1) & 7), o & 31);
	if (nbits < 32) field &= (1u << nbits) - 1u;
	const bool exact = I.flags & kFlagExactPos;
	const uint32_t low = exact ? field & ((1u << I.b) - 1u) : 0u;  // b <= 8 whenever the exact layout is on
	const uint32_t pos = exact ? field - low : field << I.b;
	if (!((uint64_t)pos + I.k - 1 < (uint64_t)B.bd.z)) return -1;
	const uint64_t P = (((uint64_t)B.bd.y << 32) | B.bd.x) + pos;
	const uint64_t rx = rc64(x, I.k);
	const uint64_t id = (uint64_t)rank + (((uint64_t)m1.y << 32) | m1.x);
	if (exact && window_matches(I.seq, P + low, I.k, x, rx)) {  // the k-mer's own window: one of the 2^b the reference scans
		if (T_out) *T_out = P + low;
		return (int64_t)id;
	}
	const int32_t j = (I.k >= 8 && I.b >= 3) ? scan_windows(I.seq, P, I.k, 1u << I.b, x, rx) : scan_windows_loop(I.seq, P, I.k, 1u << I.b, x, rx);
	if (j < 0) return -1;
	if (T_out) *T_out = P + (uint32_t)j;
	return (int64_t)id;
}

// ---- the negative filter (device_index.hpp: `filter`) ----------------------------------------------------------------
// block = hi32(h) scaled to the block count, bit i of the key = 5 bits of a second product, one per word of the block.
__device__ __forceinline__ uint64_t filter_hash(uint64_t x) {
	x ^= x >> 31; x *= 0x9E3779B97F4A7C15ull;
	x ^= x >> 29; x *= 0xBF58476D1CE4E5B9ull;
	x ^= x >> 32;
	return x;
}
__device__ __forceinline__ uint64_t filter_bits(uint64_t h) { return (h * 0xD6E8FEB86659FD93ull) >> 24; }  // 40 bits

// false: x is certainly not in V, the reference answers -1. true: unknown.
__device__ __forceinline__ bool filter_maybe(const DevIndexView& I, uint64_t x) {
	const uint64_t h = filter_hash(x);
	const uint32_t blk = __umulhi((uint32_t)(h >> 32), I.filter_blocks);
	const uint64_t g = filter_bits(h);
	uint32_t w[8];
	ld_filter_sector(I.filter + ((uint64_t)blk << 3), w);
	uint32_t ok = 1;
	#pragma unroll
	for (int i = 0; i < 8; i++) ok &= w[i] >> ((uint32_t)(g >> (5 * i)) & 31u);
	return ok & 1u;
}

__device__ __forceinline__ void filter_insert(uint32_t* filter, uint32_t filter_blocks, uint64_t x) {
	const uint64_t h = filter_hash(x);
	const uint32_t blk = __umulhi((uint32_t)(h >> 32), filter_blocks);
	const uint64_t g = filter_bits(h);
	uint32_t* p = filter + ((uint64_t)blk << 3);
	#pragma unroll
	for (int i = 0; i < 8; i++) atomicOr(p + i, 1u << ((uint32_t)(g >> (5 * i)) & 31u));
}

template <bool SMALL>
__device__ __forceinline__ int64_t lookup_one(const DevIndexView& I, uint64_t x, uint32_t mini, uint64_t* T_out = nullptr) {
	const BucketRef B = load_bucket(I, mini);
	if (B.bd.z == 0) return -1;
	uint64_t s0 = 0, s1 = 0, off = 0;
	uint32_t w[8];
	uint32_t r = 0;
	const bool hit = probe_levels<SMALL>(B, x, 0, kLevels, s0, s1, off, w, r);
	return finish_lookup(I, B, x, hit, w, r, T_out);
}

}  // namespace blight
