// lookup.cuh — the lookup core: one canonical k-mer -> identifier, on the device layout of device_index.hpp.
//
// Follows query_get_hash (blight.cpp:716-742) step for step; what changes is how many sectors each step reads.
//   1. bucket[minimizer]: empty -> -1                                   (blight.cpp:719)        1 x 16 B
//   2. group = minimizer >> lb, its DevMphf                             (blight.cpp:722)        L1-resident
//   3. BBHash levels: hash, mulhi, test bit, rank                       (bbhash.h:561-577)      1 sector per level
//   4. position field, << b, as uint32                                  (blight.cpp:473-482)    1 sector
//   5. guard pos+k-1 < bucket length (first window only)                (blight.cpp:729)
//   6. scan 2^b windows of the bucket sequence for canon(window)==x     (blight.cpp:730-739)    1-2 sectors (b<=6)
//   7. id = rank + id_offset, else -1                                   (blight.cpp:736,741)
#pragma once
#include <cstdint>

#include "device_index.hpp"
#include "kmer_math.hpp"

namespace blight {

__device__ __forceinline__ void ld_sector(const uint32_t* p, uint32_t (&w)[8]) {
	asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
	             : "l"(p));
}

// Scans `nwin` consecutive k-mer windows of the packed sequence starting at nucleotide P for x or its reverse
// complement rx. canon(window) == x  <=>  window == x or window == rx, because x is canonical (x <= rx).
__device__ __forceinline__ bool scan_windows(const uint32_t* __restrict__ seq, uint64_t P, uint32_t k, uint32_t nwin,
                                             uint64_t x, uint64_t rx) {
	const uint64_t wi = P >> 4;
	const uint32_t s = 2u * (uint32_t)(P & 15);
	const uint32_t* q = seq + wi;
	uint32_t a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
	// top = the 32 nucleotides starting at P; feed = the nucleotides after them
	uint32_t top_hi = __funnelshift_l(b, a, s), top_lo = __funnelshift_l(c, b, s);
	const uint32_t sh = 64 - 2 * k;  // window = top >> sh
	const uint32_t xl = (uint32_t)x, rl = (uint32_t)rx;
	q += 3;
	uint32_t prev = c;
	for (uint32_t j = 0; j < nwin; j += 16) {
		const uint32_t nxt = __ldg(q++);
		uint32_t feed = __funnelshift_l(nxt, prev, s);
		prev = nxt;
		#pragma unroll
		for (int t = 0; t < 16; t++) {
			const uint32_t wl = sh >= 32 ? (top_hi >> (sh - 32)) : __funnelshift_r(top_lo, top_hi, sh);
			if (wl == xl || wl == rl) {
				const uint64_t w = (((uint64_t)top_hi << 32) | top_lo) >> sh;
				if ((w == x || w == rx) && j + t < nwin) return true;
			}
			top_hi = __funnelshift_l(top_lo, top_hi, 2);
			top_lo = __funnelshift_l(feed, top_lo, 2);
			feed <<= 2;
		}
	}
	return false;
}

__device__ __forceinline__ int64_t lookup_one(const DevIndexView& I, uint64_t x, uint32_t mini) {
	const uint4 bd = __ldg(I.bucket + mini);
	if (bd.z == 0) return -1;
	const DevMphf* __restrict__ M = I.mphf + (mini >> I.lb);
	const uint4 m0 = __ldg(reinterpret_cast<const uint4*>(M));      // bits_sector_base, pos_sector_base
	const uint32_t* bits = I.bits + ((((uint64_t)m0.y << 32) | m0.x) << 3);

	// BBHash levels (bbhash.h:619-639): first level whose bit is set wins
	uint64_t s0 = 0, s1 = 0, off = 0, rank = ~0ull;
	#pragma unroll 1
	for (int level = 0; level < kLevels; level++) {
		uint64_t h;
		if (level == 0) h = s0 = hash_bis(x, kSeed0);
		else if (level == 1) h = s1 = hash_bis(x, kSeed1);
		else h = xs128_next(s0, s1);
		const uint64_t dom = __ldg(&M->dom[level]);
		const uint64_t bit = off + __umul64hi(h, dom);
		const uint64_t chunk = bit / kChunkBits;
		const uint32_t r = (uint32_t)(bit - chunk * kChunkBits);
		uint32_t w[8];
		ld_sector(bits + (chunk << 3), w);
		const uint32_t j = r >> 5, bi = r & 31;
		uint32_t hit = 0, cnt = w[7];
		#pragma unroll
		for (int i = 0; i < 7; i++) {
			const uint32_t below = (i < (int)j) ? 0xFFFFFFFFu : ((i == (int)j) ? ((1u << bi) - 1u) : 0u);
			cnt += __popc(w[i] & below);
			hit |= (i == (int)j) ? ((w[i] >> bi) & 1u) : 0u;
		}
		if (hit) { rank = cnt; break; }
		off += dom;
	}
	const uint4 m1 = __ldg(reinterpret_cast<const uint4*>(M) + 1);  // id_offset, fb_off
	const uint4 m2 = __ldg(reinterpret_cast<const uint4*>(M) + 2);  // fb_count, nbits, fields_per_sector, present
	if (rank == ~0ull) {
		// fallback map (bbhash.h:567-575), sorted by key
		const uint64_t fb_off = ((uint64_t)m1.w << 32) | m1.z;
		uint32_t lo = 0, hi = m2.x;
		while (lo < hi) {
			const uint32_t mid = (lo + hi) >> 1;
			if (__ldg(I.fb_keys + fb_off + mid) < x) lo = mid + 1; else hi = mid;
		}
		if (lo >= m2.x || __ldg(I.fb_keys + fb_off + lo) != x) return -1;
		rank = __ldg(I.fb_vals + fb_off + lo);
	}
	// position field (blight.cpp:473-482): uint32 arithmetic, << b
	const uint32_t nbits = m2.y, fps = m2.z;
	const uint64_t psec = rank / fps;
	const uint32_t slot = (uint32_t)(rank - psec * fps);
	const uint32_t* ps = I.pos + (((((uint64_t)m0.w << 32) | m0.z) + psec) << 3);
	const uint32_t o = slot * nbits, ow = o >> 5;
	const uint32_t p0 = __ldg(ps + ow), p1 = __ldg(ps + (ow < 7 ? ow + 1 : 7));
	uint32_t field = __funnelshift_r(p0, p1, o & 31);
	if (nbits < 32) field &= (1u << nbits) - 1u;
	const uint32_t pos = field << I.b;
	if (!((uint64_t)pos + I.k - 1 < (uint64_t)bd.z)) return -1;
	const uint64_t P = (((uint64_t)bd.y << 32) | bd.x) + pos;
	if (!scan_windows(I.seq, P, I.k, 1u << I.b, x, rc64(x, I.k))) return -1;
	return (int64_t)(rank + (((uint64_t)m1.y << 32) | m1.x));
}

}  // namespace blight
