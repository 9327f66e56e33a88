// front.cuh — the per-strip front end shared by the read kernels that work on super-k-mers (kernels.cu: k_reads_sk,
// part_kernels.cu: k_dispatch_runs): 2-bit pack, m-mer ordering keys, window minima, runs of equal minimizer.
//   nuc2int / str2num          kmer.h:56-98        -> phase A
//   minimizer_naive (patched)  kmer.h:791-810      -> phases B, C1 (one key per m-mer, window minimum over k-m+1 keys)
//   super-k-mer boundaries     kmer.h:629-693      -> C1 (runs of equal minimizer inside one read)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kmer_math.hpp"

namespace blight {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kStrip = 256;                          // k-mer start positions per warp strip
constexpr int kPerLane = kStrip / 32;
constexpr int kMaxW = 32;                            // k - m + 1 <= 31
constexpr int kStripWords = (kStrip + 32) / 16 + 2;  // packed words per strip (+halo k-1 <= 30, +1 for the funnel)
constexpr int kStripKeys = kStrip + kMaxW;
static_assert(kStripWords <= 32, "one lane packs one word");

// first read r in [0, n_reads) with off[r+1] > p, i.e. the read containing base position p (or the gap before it).
// Starts from the position a uniform read length would give (p0 = where read 0 of this array starts: the array may be a
// window of a larger batch) and brackets the answer exponentially: 2-3 loads for the usual near-uniform batches, a plain
// binary search in the worst case.
__device__ __noinline__ uint64_t find_read(const uint64_t* __restrict__ off, uint64_t n_reads, double reads_per_base, uint64_t p, uint64_t p0) {
	uint64_t g = (uint64_t)((double)(p > p0 ? p - p0 : 0) * reads_per_base);
	if (g > n_reads - 1) g = n_reads - 1;
	uint64_t lo, hi;  // invariant: answer in [lo, hi]
	if (__ldg(off + g + 1) > p) {
		hi = g; lo = g;
		uint64_t step = 1;
		while (lo > 0) {
			const uint64_t c = lo > step ? lo - step : 0;
			if (__ldg(off + c + 1) > p) { hi = c; lo = c; step <<= 1; }
			else { lo = c + 1; break; }
		}
	} else {
		lo = g + 1; hi = lo;
		uint64_t step = 1;
		while (hi < n_reads - 1 && !(__ldg(off + hi + 1) > p)) {
			lo = hi + 1;
			hi = hi + step < n_reads ? hi + step : n_reads - 1;
			step <<= 1;
		}
		if (hi > n_reads - 1) hi = n_reads - 1;
		if (lo > hi) lo = hi;
	}
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		if (__ldg(off + mid + 1) > p) hi = mid; else lo = mid + 1;
	}
	return lo;
}

constexpr int kMaxRuns = 64;                              // runs per strip handled through C2/C3
constexpr uint32_t kTagNone = 0xFF, kTagOverflow = 0xFE;  // s_runid: no k-mer here / run table full: plain lookup
constexpr int kKeySlots = kStripKeys + kStripKeys / 8 + 8;

__device__ __forceinline__ uint32_t kidx(uint32_t q) { return q + (q >> 3); }  // one pad word per 8 keys: a lane walks 8 consecutive keys

__device__ __forceinline__ uint64_t strip_kmer(const uint32_t* pack, uint32_t q, uint32_t k) {
	const uint32_t wi = q >> 4, s = 2u * (q & 15);
	const uint32_t a = pack[wi], b = pack[wi + 1], c = pack[wi + 2];
	return (((uint64_t)__funnelshift_l(b, a, s) << 32) | __funnelshift_l(c, b, s)) >> (64 - 2 * k);
}

// read bookkeeping of one lane while it walks consecutive positions
struct ReadCursor {
	uint64_t r, beg, next, end;
	__device__ __forceinline__ void load(const uint64_t* __restrict__ off, const uint64_t* __restrict__ endp) {
		beg = __ldg(off + r);
		next = __ldg(off + r + 1);
		end = endp ? __ldg(endp + r) : next;
	}
	// moves to the read containing p (or the gap before it); true if the read changed
	__device__ __forceinline__ bool seek(const uint64_t* __restrict__ off, const uint64_t* __restrict__ endp, uint64_t n_reads, uint64_t p) {
		if (p < next || r + 1 >= n_reads) return false;
		do { r++; next = __ldg(off + r + 1); } while (p >= next && r + 1 < n_reads);
		load(off, endp);
		return true;
	}
};

// cold paths, kept out of line so that the strip loop stays dense in the instruction cache
__device__ __noinline__ uint4 load16_slow(const char* __restrict__ p, uint32_t n) {  // n < 16 valid bytes, or p unaligned
	unsigned char ch[16];
	#pragma unroll 1
	for (uint32_t j = 0; j < 16; j++) ch[j] = j < n ? (unsigned char)p[j] : (unsigned char)'A';
	return *reinterpret_cast<uint4*>(ch);
}

__device__ __noinline__ uint32_t window_min_slow(const uint32_t* keys, uint32_t q, uint32_t w) {
	uint32_t best = 0xFFFFFFFFu;
	#pragma unroll 1
	for (uint32_t e = q; e < q + w; e++) best = min(best, keys[kidx(e)]);
	return best;
}

// the 64 bases starting at strip position q, as four packed words (first base in the high bits of .x)
__device__ __forceinline__ uint4 strip_bases64(const uint32_t* pack, uint32_t q) {
	const uint32_t wi = q >> 4, s = 2u * (q & 15);
	uint32_t v[5];
	#pragma unroll
	for (int i = 0; i < 5; i++) v[i] = (wi + i < (uint32_t)kStripWords) ? pack[wi + i] : 0u;
	return make_uint4(__funnelshift_l(v[1], v[0], s), __funnelshift_l(v[2], v[1], s), __funnelshift_l(v[3], v[2], s), __funnelshift_l(v[4], v[3], s));
}

// the 64 bases of the packed bucket sequences starting at base P (five word loads, one or two sectors)
__device__ __forceinline__ uint4 seq_bases64(const uint32_t* __restrict__ seq, uint64_t P) {
	const uint32_t* p = seq + (P >> 4);
	const uint32_t s = 2u * (uint32_t)(P & 15);
	const uint32_t a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4);
	return make_uint4(__funnelshift_l(b, a, s), __funnelshift_l(c, b, s), __funnelshift_l(d, c, s), __funnelshift_l(e, d, s));
}

// reverse complement of a 64-base block (base j of the result = complement of base 63 - j)
__device__ __forceinline__ uint4 rc_bases64(uint4 r) {
	auto rc16 = [](uint32_t w) {
		uint32_t t = __brev(w ^ 0xAAAAAAAAu);
		return ((t & 0x55555555u) << 1) | ((t >> 1) & 0x55555555u);
	};
	return make_uint4(rc16(r.w), rc16(r.z), rc16(r.y), rc16(r.x));
}

// 64 bases shifted towards the front by `n` bases (n <= 64), zero filled
__device__ __forceinline__ uint4 shl_bases64(uint4 r, uint32_t n) {
	uint64_t hi = ((uint64_t)r.x << 32) | r.y, lo = ((uint64_t)r.z << 32) | r.w;
	const uint32_t sh = 2 * n;
	if (sh >= 64) { hi = sh >= 128 ? 0 : lo << (sh - 64); lo = 0; }
	else if (sh) { hi = (hi << sh) | (lo >> (64 - sh)); lo <<= sh; }
	return make_uint4((uint32_t)(hi >> 32), (uint32_t)hi, (uint32_t)(lo >> 32), (uint32_t)lo);
}

// For two 64-base blocks: bit 63 - j of the result is set iff the k bases starting at base j differ somewhere
// (bases past the block count as equal). One lane compares a whole super-k-mer against the index text this way.
__device__ __forceinline__ uint64_t mismatch_windows64(uint4 a, uint4 b, uint32_t k) {
	auto comp = [](uint32_t d) {  // 16 bases, 2 bits each -> 16 bits, 1 = the base differs; first base -> bit 15
		uint32_t x = (d | (d >> 1)) & 0x55555555u;
		x = (x | (x >> 1)) & 0x33333333u;
		x = (x | (x >> 2)) & 0x0F0F0F0Fu;
		x = (x | (x >> 4)) & 0x00FF00FFu;
		return (x | (x >> 8)) & 0xFFFFu;
	};
	const uint64_t M = ((uint64_t)comp(a.x ^ b.x) << 48) | ((uint64_t)comp(a.y ^ b.y) << 32) | ((uint64_t)comp(a.z ^ b.z) << 16) | comp(a.w ^ b.w);
	uint64_t A = M;  // OR over a window of k bases, by doubling: bit p collects bits p, p-1, .., p-k+1
	#pragma unroll 1
	for (uint32_t win = 1; win < k;) {
		const uint32_t sh = min(win, k - win);
		A |= A << sh;
		win += sh;
	}
	return A;
}

// Shared-memory slice of one warp.
struct StripSmem {
	uint32_t* pack;    // [kStripWords]  2-bit codes, 16 per word, first base in the high bits
	uint32_t* bad;     // [kStripWords]  1 bit per base (bit 15-j of word i = base 16i+j): not ACGTacgt
	uint32_t* keys;    // [kKeySlots]    ordering key of the m-mer at strip position q, at kidx(q)
	uint16_t* run_q;   // [kMaxRuns]     strip position of the run's first k-mer
	uint32_t* run_key; // [kMaxRuns]     ordering key of the run's minimizer (mini_from_key gives the bucket)
	uint64_t* run_o;   // [kMaxRuns]     output slot of the run's first k-mer (WANT_O)
	uint64_t* runid8;  // [kStrip / 8]   run of every strip position, one byte each (kTagNone / kTagOverflow)
};

// Phases A, B, C1 of one strip (256 start positions from base t0). Returns the number of runs found (may exceed
// kMaxRuns: the surplus k-mers carry kTagOverflow); `invalid` counts k-mers holding a byte nuc2int rejects.
template <bool WANT_O>
__device__ __forceinline__ uint32_t strip_front(const StripSmem& S, uint32_t lane, uint32_t k, uint32_t m, const char* __restrict__ bases,
                                                const uint64_t* __restrict__ read_off, const uint64_t* __restrict__ read_end,
                                                const uint64_t* __restrict__ kmer_off, uint64_t n_reads, uint64_t total_bases,
                                                double reads_per_base, uint64_t guess_p0, bool aligned16, const uint32_t* __restrict__ packed,
                                                uint64_t t0, uint32_t& invalid) {
	uint32_t* pack = S.pack;
	uint32_t* bad = S.bad;
	uint32_t* keys = S.keys;
	const uint32_t w = k - m + 1;
	const uint32_t mmask = (1u << (2 * m)) - 1u;
	const uint64_t kones = (1ull << k) - 1;
	const uint32_t n_pos = (uint32_t)min((uint64_t)kStrip, total_bases - t0);
	const uint32_t n_load = (uint32_t)min((uint64_t)(kStrip + 32), total_bases - t0);
	__syncwarp();
	// A. pack (or, when the caller already holds 2-bit codes — the host packer of the *_host entry points — copy)
	if (lane < kStripWords && packed) {
		pack[lane] = lane * 16 < n_load ? __ldcs(packed + (t0 >> 4) + lane) : 0u;
		bad[lane] = 0;
	} else if (lane < kStripWords) {
		const uint32_t b0 = lane * 16;
		uint32_t word = 0, badw = 0;
		if (b0 < n_load) {
			const uint4 v = (aligned16 && b0 + 16 <= n_load) ? __ldcs(reinterpret_cast<const uint4*>(bases + t0 + b0))
			                                                  : load16_slow(bases + t0 + b0, n_load - b0);
			const uint32_t q4[4] = {nuc_code4(v.x), nuc_code4(v.y), nuc_code4(v.z), nuc_code4(v.w)};
			#pragma unroll
			for (int j = 0; j < 4; j++) {
				word = (word << 8) | (q4[j] & 0xFFu);
				badw = (badw << 4) | (q4[j] >> 8);
			}
		}
		pack[lane] = word;
		bad[lane] = badw;
	}
	ReadCursor rc_;
	rc_.r = 0;
	if (lane == 0) rc_.r = find_read(read_off, n_reads, reads_per_base, t0, guess_p0);
	rc_.r = __shfl_sync(0xffffffffu, rc_.r, 0);
	__syncwarp();
	// B. m-mer keys
	for (uint32_t q = lane; q < n_pos + w - 1; q += 32) {
		const uint32_t wi = q >> 4, s = 2u * (q & 15);
		const uint32_t v = __funnelshift_l(pack[wi + 1], pack[wi], s) >> (32 - 2 * m);
		keys[kidx(q)] = mini_key(parity_canon(v & mmask, m));
	}
	__syncwarp();

	// C1. runs of equal minimizer inside one read; lane L owns positions 8L .. 8L+7
	const uint32_t q0 = lane * 8;
	uint32_t kmin[8];  // window minimum (as key) of each owned position
	{
		const uint32_t* kp = keys + 9 * lane;  // kidx(q0 + e) = 9 L + e + (e >> 3)
		// w >= 8 (the launcher sends smaller windows to k_reads):
		// window j = keys [j, j + w) = (suffix of [j, 7)) + common [7, w) + (prefix of [w, w + j))
		uint32_t common = kp[7];
		for (uint32_t e = 8; e < w; e++) common = min(common, kp[e + (e >> 3)]);
		kmin[7] = common;
		uint32_t sfx = 0xFFFFFFFFu;
		#pragma unroll
		for (int j = 6; j >= 0; j--) { sfx = min(sfx, kp[j]); kmin[j] = min(common, sfx); }
		uint32_t pfx = 0xFFFFFFFFu;
		const uint32_t* kw = kp + w - 1;
		const uint32_t wl = (w - 1) & 7;  // kidx is not linear: position w - 1 + j sits at w - 1 + j + ((w - 1 + j) >> 3)
		#pragma unroll
		for (int j = 1; j < 8; j++) { pfx = min(pfx, kw[j + ((w - 1) >> 3) + ((wl + j) >> 3)]); kmin[j] = min(kmin[j], pfx); }
	}
	// which owned positions start a k-mer, and where reads change
	const uint64_t p0 = t0 + q0;
	while (rc_.r + 1 < n_reads && __ldg(read_off + rc_.r + 1) <= p0) rc_.r++;
	rc_.load(read_off, read_end);
	const uint32_t r_first = (uint32_t)rc_.r;
	uint32_t have = 0, newread = 0;
	{
		const uint32_t wi = q0 >> 4, o16 = q0 & 15;
		const uint64_t bb = ((uint64_t)bad[wi] << 48) | ((uint64_t)bad[wi + 1] << 32) | ((uint64_t)bad[wi + 2] << 16) | bad[wi + 3];
		#pragma unroll 1
		for (int j = 0; j < 8; j++) {
			const uint64_t p = p0 + j;
			if (j && rc_.seek(read_off, read_end, n_reads, p)) newread |= 1u << j;
			if (q0 + j < n_pos && p >= rc_.beg && p + k <= rc_.end) {
				// nuc2int rejects any byte outside ACGTacgt (kmer.h:56-69); only bases of queried k-mers are ever looked at
				if ((bb >> (64 - o16 - j - k)) & kones) invalid++;
				else have |= 1u << j;
			}
		}
	}
	const uint32_t r_last = (uint32_t)rc_.r;
	// run starts: a k-mer whose predecessor is missing, in another read, or has another minimizer
	uint32_t p_have = __shfl_up_sync(0xffffffffu, have >> 7, 1);
	const uint32_t p_key = __shfl_up_sync(0xffffffffu, kmin[7], 1);
	const uint32_t p_r = __shfl_up_sync(0xffffffffu, r_last, 1);
	if (lane == 0) p_have = 0;
	uint32_t bnd = 0;
	if ((have & 1u) && (!p_have || p_key != kmin[0] || p_r != r_first)) bnd = 1u;
	#pragma unroll
	for (int j = 1; j < 8; j++)
		if (((have >> j) & 1u) && (!((have >> (j - 1)) & 1u) || kmin[j] != kmin[j - 1] || ((newread >> j) & 1u))) bnd |= 1u << j;
	uint32_t incl = __popc(bnd);
	#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
		if (lane >= (uint32_t)o) incl += t;
	}
	const uint32_t n_runs = __shfl_sync(0xffffffffu, incl, 31);
	const uint32_t run_base = incl - __popc(bnd);
	{
		// run records (the minimizer is filled in by C2) and the run of every owned position
		ReadCursor c2;
		c2.r = 0; c2.beg = c2.next = c2.end = 0;
		if (WANT_O) {
			c2.r = rc_.r - (r_last - r_first);
			c2.load(read_off, read_end);
		}
		// the run's minimizer is known here (kmin[] stays in registers: static indices only) and is never recomputed
		#pragma unroll
		for (int j = 0; j < 8; j++) {
			const uint32_t id = run_base + __popc(bnd & ((1u << j) - 1u));
			if (((bnd >> j) & 1u) && id < (uint32_t)kMaxRuns) S.run_key[id] = kmin[j];
		}
		uint32_t id = run_base;
		#pragma unroll 1
		for (uint32_t bm = bnd; bm && id < (uint32_t)kMaxRuns; bm &= bm - 1, id++) {
			const uint32_t j = __ffs(bm) - 1;
			S.run_q[id] = (uint16_t)(q0 + j);
			if (WANT_O) {
				c2.seek(read_off, read_end, n_reads, p0 + j);
				S.run_o[id] = __ldg(kmer_off + c2.r) + (p0 + j - c2.beg);
			}
		}
		uint64_t tags = 0;
		#pragma unroll
		for (int j = 0; j < 8; j++) {
			uint32_t tag = kTagNone;
			if ((have >> j) & 1u) {
				const uint32_t rid = run_base + __popc(bnd & ((2u << j) - 1u)) - 1u;  // the first k-mer of a strip always starts a run
				tag = rid < (uint32_t)kMaxRuns ? rid : kTagOverflow;
			}
			tags |= (uint64_t)tag << (8 * j);
		}
		S.runid8[lane] = tags;
	}
	__syncwarp();
	return n_runs;
}

}  // namespace
// Next work item of a persistent warp. With a ticket counter (zeroed before the launch) the items past the first one per warp
// are handed out on demand — a sub-batch is only some 30-50 items per warp and their cost varies (one chunk's lookups walk
// more BBHash levels than another's), so a fixed stride left the SMs idling behind the slowest warps at every kernel's end.
__device__ __forceinline__ uint64_t next_item(unsigned long long* ticket, uint64_t cur, uint64_t n_warps, uint64_t first, uint32_t lane) {
	if (!ticket) return cur + n_warps;
	unsigned long long t = 0;
	if (lane == 0) t = atomicAdd(ticket, 1ull);
	t = __shfl_sync(0xffffffffu, t, 0);
	return first + n_warps + t;
}

}  // namespace blight
