// kernels.cu — sm_100a kernels of the batched query path and their launchers.
//
//   k_lookup_kmers   canonical k-mers (+ optional minimizers) -> ids          query_kmer_hash, blight.cpp:545-550
//   k_reads          ASCII reads -> 2-bit pack -> canonical k-mers + rolling  query_sequence_hash/bool,
//                    minimizers -> {pairs | ids | counts}, one lookup per     blight.cpp:554-591; kmer.h:791-810
//                    k-mer (the literal formulation)
//   k_reads_sk       the same path per super-k-mer: first k-mer of a run      + kmer.h:629-693 (super-k-mers)
//                    through the lookup, the others against one predicted
//                    window -> {ids | counts | abundance / colour / gather}   + the snippet applications' consumers
//   k_window_valid   upload-time pass: every window of every bucket through
//                    the lookup core -> valid / pos_id / filter / exact
//                    positions (device_index.hpp)
//
// k_reads is a persistent kernel of independent warps.  A warp takes strips of kStrip consecutive base positions of
// the input buffer (strip s -> warp s mod #warps), packs strip+halo bases 2 bits each into its private slice of
// shared memory (coalesced 16-byte streaming loads), computes the ordering key of every m-mer once, takes the
// window minimum over the k-m+1 keys of each k-mer (equal to minimizer_naive on the canonical k-mer, because the
// set of canonical m-mers of a k-mer is strand invariant and revhash is a bijection), and either stores
// (canon, minimizer) or goes straight into the lookup core, adjacent lanes holding adjacent k-mers so that the
// k-mers of one super-k-mer share their bucket descriptor, MPHF descriptor and sequence sectors in L1.  There is no
// block-level barrier anywhere: lookups have very uneven latency, and warps must not wait for each other.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdlib>
#include <mutex>

#include "front.cuh"
#include "kernels.hpp"
#include "lookup.cuh"

namespace blight {

std::atomic<uint64_t> g_launches{0};

namespace {

template <bool HAS_MINI, bool SMALL>
__global__ void __launch_bounds__(kThreads) k_lookup_kmers(DevIndexView I, const uint64_t* __restrict__ canon,
                                                           const uint32_t* __restrict__ mini, uint64_t n,
                                                           int64_t* __restrict__ ids) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const uint64_t x = __ldcs(canon + i);
		const uint32_t mn = HAS_MINI ? __ldcs(mini + i) : minimizer_of_kmer(x, I.k, I.m);
		__stcs(reinterpret_cast<long long*>(ids + i), (long long)lookup_one<SMALL>(I, x, mn));
	}
}

// ---- the upload-time passes over every window of the index text (device_index.hpp: valid / pos_id / filter / exact positions) ----
//
// k_window_answers   one thread per base position p (every window start the reference's scan can reach: the text plus the
//                    2^b windows past its end, which read as zero padding):
//                      x  = canonical k-mer spelled by the window at p
//                      own answer   lookup(x routed to p's OWN bucket)     -> valid[p], pos_id[p]   (what a query predicted
//                                   to sit at p receives: it has the bucket's minimizer, k_reads_sk C3)
//                      true answer  lookup(x routed to minimizer(x))       -> filter insert when found
//                    The two differ for windows that span two super-k-mers of a bucket ("junction" windows) and whose k-mer
//                    has another minimizer: the reference's 2^b scan (blight.cpp:732-739) never re-checks the bucket length,
//                    so a query routed to bucket B can match a window that starts in a FOLLOWING bucket. Every k-mer the
//                    reference answers "found" matched some window of the text, hence is inserted here: the filter has no
//                    false negatives with respect to the reference's answers.
//                    cand[p]: the own answer matched at p itself and x has the bucket's minimizer: p may own the low bits of
//                    its position field.
// k_claim_low        exact-position layout: among the candidate windows that hit the same position field (the key the MPHF
//                    was built on, plus alien junction k-mers that collide on its rank) ONE gets the field's low b bits:
//                    windows with a candidate neighbour first (keys inside a super-k-mer; aliens sit among windows that are
//                    mostly not found), then the smallest offset. Deterministic: atomicMin over (priority, offset).
// k_write_low        the winner ORs its offset into the (zero) low bits. Nothing reads `pos` in this kernel.
template <bool SMALL>
__global__ void __launch_bounds__(kThreads) k_window_answers(DevIndexView I, uint64_t n_buckets, uint64_t total_nuc, uint64_t n_scan,
                                                             uint32_t* __restrict__ valid, uint32_t* __restrict__ pos_id, uint32_t* __restrict__ cand,
                                                             uint32_t* __restrict__ filter, uint32_t filter_blocks) {
	const uint64_t n_round = (n_scan + 31) & ~31ull;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_round; p += stride) {
		bool v = false, c = false;
		if (p < n_scan) {
			const uint64_t wv = window_at(I.seq, p, I.k);
			const uint64_t rc = rc64(wv, I.k);
			const uint64_t x = wv < rc ? wv : rc;
			const uint32_t mnx = minimizer_of_kmer(x, I.k, I.m);
			int64_t id = -1;
			uint32_t own = 0xFFFFFFFFu;
			if (p < total_nuc) {
				// last bucket whose start is <= p: the non-empty bucket holding p (empty ones share their successor's start)
				uint64_t lo = 0, hi = n_buckets - 1;
				while (lo < hi) {
					const uint64_t mid = (lo + hi + 1) >> 1;
					const uint4 bd = __ldg(I.bucket + mid);
					if ((((uint64_t)bd.y << 32) | bd.x) <= p) lo = mid; else hi = mid - 1;
				}
				const uint4 bd = __ldg(I.bucket + lo);
				const uint64_t start = ((uint64_t)bd.y << 32) | bd.x;
				if (p >= start && p - start < bd.z) {
					own = (uint32_t)lo;
					uint64_t T = ~0ull;
					id = lookup_one<SMALL>(I, x, own, &T);
					v = id >= 0;
					c = v && mnx == own && T == p;
				}
				if (pos_id) pos_id[p] = v ? (uint32_t)((uint64_t)id - I.id_base) : 0xFFFFFFFFu;
			}
			if (filter) {
				const bool found = mnx == own ? v : lookup_one<SMALL>(I, x, mnx) >= 0;
				if (found) filter_insert(filter, filter_blocks, x);
			}
		}
		const uint32_t vw = __ballot_sync(0xffffffffu, v);
		const uint32_t cw = __ballot_sync(0xffffffffu, c);
		if ((threadIdx.x & 31) == 0 && p < ((total_nuc + 31) & ~31ull)) {
			valid[p >> 5] = vw;
			if (cand) cand[p >> 5] = cw;
		}
	}
}

// where the position field of local identifier `lid` lives: group by binary search over the id offsets
struct FieldRef { uint32_t* word; uint32_t shift, nbits; };
__device__ __forceinline__ FieldRef field_of(const DevIndexView& I, uint32_t* pos_rw, uint32_t n_groups, uint64_t gid) {
	uint32_t lo = 0, hi = n_groups - 1;
	while (lo < hi) {  // last present group whose id_offset <= gid (absent groups repeat their successor's offset and hold no key)
		const uint32_t mid = (lo + hi + 1) >> 1;
		if (I.mphf[mid].id_offset <= gid) lo = mid; else hi = mid - 1;
	}
	while (!I.mphf[lo].present && lo > 0) lo--;
	const DevMphf& M = I.mphf[lo];
	const uint32_t rank = (uint32_t)(gid - M.id_offset);
	const uint32_t psec = rank / M.fields_per_sector, slot = rank - psec * M.fields_per_sector;
	const uint32_t o = slot * M.nbits;
	return FieldRef{pos_rw + ((M.pos_sector_base + psec) << 3) + (o >> 5), o & 31, M.nbits};
}

// key of a candidate window: (no candidate neighbour) << 8 | offset inside its 2^b windows
__device__ __forceinline__ uint32_t claim_key(const DevIndexView& I, const uint32_t* __restrict__ cand, uint64_t p, uint64_t total_nuc) {
	const bool left = p > 0 && ((cand[(p - 1) >> 5] >> ((p - 1) & 31)) & 1u);
	const bool right = p + 1 < total_nuc && ((cand[(p + 1) >> 5] >> ((p + 1) & 31)) & 1u);
	return (left || right) ? 0u : 0x100u;
}

template <bool WRITE>
__global__ void __launch_bounds__(kThreads) k_claim_low(DevIndexView I, uint64_t n_buckets, uint64_t total_nuc, const uint32_t* __restrict__ cand,
                                                        const uint32_t* __restrict__ pos_id, uint32_t* __restrict__ claim, uint32_t* pos_rw,
                                                        uint32_t n_groups) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < total_nuc; p += stride) {
		if (!((cand[p >> 5] >> (p & 31)) & 1u)) continue;
		uint64_t lo = 0, hi = n_buckets - 1;
		while (lo < hi) {
			const uint64_t mid = (lo + hi + 1) >> 1;
			const uint4 bd = __ldg(I.bucket + mid);
			if ((((uint64_t)bd.y << 32) | bd.x) <= p) lo = mid; else hi = mid - 1;
		}
		const uint4 bd = __ldg(I.bucket + lo);
		const uint64_t start = ((uint64_t)bd.y << 32) | bd.x;
		const uint32_t j = (uint32_t)(p - start) & ((1u << I.b) - 1u);  // the field's window range starts at a multiple of 2^b
		const uint32_t key = claim_key(I, cand, p, total_nuc) | j;
		const uint32_t lid = pos_id[p];
		if (!WRITE) {
			atomicMin(claim + lid, key);
		} else if (j && claim[lid] == key) {
			const FieldRef f = field_of(I, pos_rw, n_groups, (uint64_t)lid + I.id_base);
			atomicOr(f.word, j << f.shift);
			if (f.shift + I.b > 32) atomicOr(f.word + 1, j >> (32 - f.shift));
		}
	}
}

enum ReadsMode { kEmitPairs = 0, kLookupIds = 1, kLookupCount = 2, kConsumeCount = 3, kConsumeColor = 4, kGatherTable = 5 };

// Where the identifier of a k-mer goes (k_reads_sk). Besides the id array of query_sequence_hash, the consumers the
// reference's applications put behind it (Abundance_De_Bruijn_graph_snippet.cpp:118-151, Colored_…:117-151), fused so
// that ids never leave the GPU: abundance[id]++, color[id * n_colors + c] = true, and the read-side gather abundance[id].
struct Sink {
	int64_t* ids;       // kLookupIds: ids[o] = id
	uint32_t* table;    // kConsumeCount: table[id] += 1; kConsumeColor: bit id * n_colors + color set; kGatherTable: read
	uint32_t* out32;    // kGatherTable: out32[o] = table[id], 0xFFFFFFFF when the k-mer is absent
	uint32_t n_colors, color;
};
template <int MODE> constexpr bool mode_wants_slot() { return MODE == kLookupIds || MODE == kGatherTable; }
template <int MODE> constexpr bool mode_wants_id() { return MODE != kLookupCount && MODE != kEmitPairs; }

template <int MODE>
__device__ __forceinline__ void emit(const Sink& K, int64_t id, uint64_t o) {
	if (MODE == kLookupIds) __stcs(reinterpret_cast<long long*>(K.ids + o), (long long)id);
	else if (MODE == kConsumeCount) { if (id >= 0) atomicAdd(K.table + id, 1u); }
	else if (MODE == kConsumeColor) {
		if (id >= 0) { const uint64_t bit = (uint64_t)id * K.n_colors + K.color; atomicOr(K.table + (bit >> 5), 1u << (bit & 31)); }
	} else if (MODE == kGatherTable) __stcs(K.out32 + o, id >= 0 ? __ldg(K.table + id) : 0xFFFFFFFFu);
}

template <int MODE, bool SMALL>
__global__ void __launch_bounds__(kThreads, 4) k_reads(DevIndexView I, uint32_t k, uint32_t m, const char* __restrict__ bases,
                                                    const uint64_t* __restrict__ read_off, const uint64_t* __restrict__ read_end,
                                                    const uint64_t* __restrict__ kmer_off, uint64_t n_reads, uint64_t total_bases,
                                                    uint64_t strip_lo, uint64_t strip_hi, bool aligned16, const uint32_t* __restrict__ packed,
                                                    double reads_per_base, uint64_t guess_p0, uint64_t* __restrict__ out_canon,
                                                    uint32_t* __restrict__ out_mini, int64_t* __restrict__ out_ids,
                                                    uint64_t* __restrict__ ctr, unsigned long long* ticket) {
	__shared__ uint32_t s_pack[kWarps][kStripWords];  // 2-bit codes, 16 per word, first base in the high bits
	__shared__ uint32_t s_bad[kWarps][kStripWords];   // 1 bit per base (bit 15-j of word i = base 16i+j): not ACGTacgt
	__shared__ uint32_t s_keys[kWarps][kStripKeys];   // ordering key of the m-mer starting at each strip position

	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t* pack = s_pack[wid];
	uint32_t* bad = s_bad[wid];
	uint32_t* keys = s_keys[wid];
	const uint32_t w = k - m + 1;
	const uint32_t mmask = (1u << (2 * m)) - 1u;
	const uint64_t warp_stride = (uint64_t)gridDim.x * kWarps;
	uint32_t found = 0, notfound = 0, invalid = 0;

	for (uint64_t strip = strip_lo + (uint64_t)blockIdx.x * kWarps + wid; strip < strip_hi; strip = next_item(ticket, strip, warp_stride, strip_lo, lane)) {
		const uint64_t t0 = strip * kStrip;
		const uint32_t n_pos = (uint32_t)min((uint64_t)kStrip, total_bases - t0);                        // positions owned by this strip
		const uint32_t n_load = (uint32_t)min((uint64_t)(kStrip + 32), total_bases - t0);  // owned + halo (k-1 <= 30)
		__syncwarp();
		// A. pack: lane i converts bases [16i, 16i+16) of the strip (or copies them when the caller holds 2-bit codes)
		if (lane < kStripWords && packed) {
			pack[lane] = lane * 16 < n_load ? __ldcs(packed + (t0 >> 4) + lane) : 0u;
			bad[lane] = 0;
		} else if (lane < kStripWords) {
			const uint32_t b0 = lane * 16;
			uint32_t word = 0, badw = 0;
			if (b0 < n_load) {
				const uint4 v = (aligned16 && b0 + 16 <= n_load) ? __ldcs(reinterpret_cast<const uint4*>(bases + t0 + b0))  // t0 % 256 == 0
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
		// read containing the first position of the strip (lane 0 searches, everybody starts from there)
		uint64_t r = 0;
		if (lane == 0) r = find_read(read_off, n_reads, reads_per_base, t0, guess_p0);
		r = __shfl_sync(0xffffffffu, r, 0);
		__syncwarp();
		// B. m-mer keys
		for (uint32_t q = lane; q < n_pos + w - 1; q += 32) {
			const uint32_t wi = q >> 4, s = 2u * (q & 15);
			const uint32_t v = __funnelshift_l(pack[wi + 1], pack[wi], s) >> (32 - 2 * m);
			keys[q] = mini_key(parity_canon(v & mmask, m));
		}
		__syncwarp();
		// C. one k-mer per lane: adjacent lanes, adjacent positions
		#pragma unroll 1
		for (uint32_t q = lane; q < n_pos; q += 32) {
			const uint64_t p = t0 + q;
			while (r + 1 < n_reads && __ldg(read_off + r + 1) <= p) r++;  // positions only grow: gallop forward
			const uint64_t rbeg = __ldg(read_off + r), rend = read_end ? __ldg(read_end + r) : __ldg(read_off + r + 1);
			if (!(p >= rbeg && p + k <= rend)) continue;  // no k-mer starts here (tail of a read, a gap, a read shorter than k)
			const uint32_t wi = q >> 4, s = 2u * (q & 15);
			// nuc2int rejects any byte outside ACGTacgt (kmer.h:56-69); only bases of queried k-mers are ever looked at
			const uint64_t bb = ((uint64_t)bad[wi] << 32) | ((uint64_t)bad[wi + 1] << 16) | bad[wi + 2];
			if ((bb >> (48 - (q & 15) - k)) & ((1ull << k) - 1)) { invalid++; continue; }
			uint32_t best = keys[q];
			for (uint32_t j = 1; j < w; j++) best = min(best, keys[q + j]);
			const uint32_t a = pack[wi], b = pack[wi + 1], c = pack[wi + 2];
			const uint64_t fwd = ((((uint64_t)__funnelshift_l(b, a, s)) << 32) | __funnelshift_l(c, b, s)) >> (64 - 2 * k);
			const uint64_t rc = rc64(fwd, k);
			const uint64_t x = fwd < rc ? fwd : rc;
			const uint32_t mn = mini_from_key(best);
			const uint64_t o = MODE != kLookupCount ? __ldg(kmer_off + r) + (p - rbeg) : 0;
			if (MODE == kEmitPairs) {
				__stcs(reinterpret_cast<unsigned long long*>(out_canon + o), (unsigned long long)x);
				__stcs(out_mini + o, mn);
			} else {
				const int64_t id = lookup_one<SMALL>(I, x, mn);
				if (MODE == kLookupIds) __stcs(reinterpret_cast<long long*>(out_ids + o), (long long)id);
				if (id >= 0) found++; else notfound++;
			}
		}
	}

	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		found += __shfl_xor_sync(0xffffffffu, found, o);
		notfound += __shfl_xor_sync(0xffffffffu, notfound, o);
		invalid += __shfl_xor_sync(0xffffffffu, invalid, o);
	}
	if (lane == 0) {
		if (found) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_FOUND], (unsigned long long)found);
		if (notfound) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_NOT_FOUND], (unsigned long long)notfound);
		if (found + notfound) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_QUERIES], (unsigned long long)found + notfound);
		if (invalid) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_INVALID], (unsigned long long)invalid);
	}
}

// ---- k_reads_sk: the same path, super-k-mer aware ------------------------------------------------------------------
//
// Consecutive k-mers of a read that share their minimizer (a super-k-mer, kmer.h:629-693) sit at consecutive positions
// of one bucket when they are in the graph. Per strip a warp therefore
//   C1  cuts the strip's k-mers into runs of equal minimizer inside one read (a lane owns 8 consecutive positions: the
//       window minimum is rolled over them, the read bookkeeping is amortised),
//   C2  sends the FIRST k-mer of every run through the whole lookup (one run per lane: a dense batch), keeping where it
//       matched (T), on which strand, and how far the bucket extends from there,
//   C3  checks for every other k-mer of a run the single window at T +- distance: on a match the answer is what the
//       reference answers for that window — the precomputed valid[] bit, or in id mode the pos_id[] entry
//       (device_index.hpp): no MPHF probe, no position read, no 2^b scan,
//   C4  sends what is left (k-mers covering a sequencing error, runs whose first k-mer is absent) through the negative
//       filter (one sector proves "-1" for the ones in no bucket) and the survivors through the whole lookup, compacted
//       into dense batches.
// C2 and C4 share ONE inlined copy of the lookup core (two turns of the same loop): the kernel must stay inside the
// 32 KB instruction cache, warps sit at unrelated program counters (the v6 profile lost 20 % of its issue slots to
// instruction fetch).
// Answers are identical to k_reads: a window that equals the query decides "found" exactly as the reference does.
// resident CTAs per SM the super-k-mer kernel is compiled for: 4 (64 registers). Measured: 5 (48 registers, 64 bytes of
// spills) 44.1 ms vs 41.3 ms per 1.2 G k-mers in counting mode.
#ifndef BLIGHT_SK_BLOCKS
#define BLIGHT_SK_BLOCKS 4
#endif

template <int MODE, bool SMALL>
__global__ void __launch_bounds__(kThreads, BLIGHT_SK_BLOCKS) k_reads_sk(DevIndexView I, uint32_t k, uint32_t m, const char* __restrict__ bases,
                                                       const uint64_t* __restrict__ read_off, const uint64_t* __restrict__ read_end,
                                                       const uint64_t* __restrict__ kmer_off, uint64_t n_reads, uint64_t total_bases,
                                                       uint64_t strip_lo, uint64_t strip_hi, bool aligned16, const uint32_t* __restrict__ packed,
                                                       double reads_per_base, uint64_t guess_p0, Sink K, uint64_t* __restrict__ ctr,
                                                       unsigned long long* ticket) {
	constexpr bool kSlot = mode_wants_slot<MODE>(), kId = mode_wants_id<MODE>();
	__shared__ uint32_t s_pack[kWarps][kStripWords];
	__shared__ uint32_t s_bad[kWarps][kStripWords];
	__shared__ uint32_t s_keys[kWarps][kKeySlots];   // key of position q at kidx(q)
	__shared__ uint64_t s_run_T[kWarps][kMaxRuns];   // where the run's first k-mer matched (absolute base position)
	__shared__ uint64_t s_run_o[kSlot ? kWarps : 1][kMaxRuns];  // output slot of the run's first k-mer
	__shared__ uint32_t s_run_mn[kWarps][kMaxRuns];  // minimizer of the run
	__shared__ uint32_t s_run_key[kWarps][kMaxRuns]; // its ordering key, as the front end found it
	__shared__ uint16_t s_run_q[kWarps][kMaxRuns];   // strip position of the run's first k-mer
	__shared__ uint16_t s_run_dmax[kWarps][kMaxRuns];// largest distance whose predicted window still starts inside the bucket
	__shared__ uint8_t s_run_flag[kWarps][kMaxRuns]; // bit0: first k-mer found, bit1: text and read on the same strand
	__shared__ uint64_t s_runid8[kWarps][kStrip / 8];// run of every strip position, one byte each
	__shared__ uint8_t s_resid[kWarps][kStrip];      // positions left for C4
	__shared__ uint64_t s_run_okv[kWarps][kMaxRuns]; // C3a: which k-mers of the run match their predicted window (low half) and those windows' valid bits

	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t* pack = s_pack[wid];
	uint32_t* keys = s_keys[wid];
	uint8_t* runid = reinterpret_cast<uint8_t*>(s_runid8[wid]);
	const uint32_t ow = kSlot ? wid : 0;
	const StripSmem S{pack, s_bad[wid], keys, s_run_q[wid], s_run_key[wid], s_run_o[ow], s_runid8[wid]};
	const uint32_t w = k - m + 1;
	const uint64_t warp_stride = (uint64_t)gridDim.x * kWarps;
	const uint32_t lt_mask = (1u << lane) - 1u;
	const bool filter_anchors = I.filter && (I.flags & kFlagFilterAnchors);
	uint32_t found = 0, notfound = 0, invalid = 0;

	for (uint64_t strip = strip_lo + (uint64_t)blockIdx.x * kWarps + wid; strip < strip_hi; strip = next_item(ticket, strip, warp_stride, strip_lo, lane)) {
		const uint64_t t0 = strip * kStrip;
		__syncwarp();
		const uint32_t n_runs = strip_front<kSlot>(S, lane, k, m, bases, read_off, read_end, kmer_off, n_reads, total_bases,
		                                                        reads_per_base, guess_p0, aligned16, packed, t0, invalid);
		uint32_t n_res = 0;
		if (n_runs > (uint32_t)kMaxRuns) {
			// more runs than the table holds (many tiny reads): those k-mers take the plain lookup in C4
			#pragma unroll 1
			for (int it = 0; it < kPerLane; it++) {
				const uint32_t q = it * 32 + lane;
				const bool ov = runid[q] == kTagOverflow;
				const uint32_t om = __ballot_sync(0xffffffffu, ov);
				if (ov) s_resid[wid][n_res + __popc(om & lt_mask)] = (uint8_t)q;
				n_res += __popc(om);
			}
		}
		const uint32_t n_anch = n_runs < (uint32_t)kMaxRuns ? n_runs : (uint32_t)kMaxRuns;

		#pragma unroll 1
		for (int phase = 0; phase < 2; phase++) {
			if (phase == 1) {
				// C3a. one lane per run: the run's text against the index text next to where its first k-mer matched, all of the
				// run's windows at once (64 bases each side; 2-bit XOR -> per-base mismatch bits -> OR over k consecutive bases).
				// Bit d of the low half: k-mer d of the run equals the window at T +- d; of the high half: that window's valid bit.
				#pragma unroll 1
				for (uint32_t base = 0; base < n_anch; base += 32) {
					const uint32_t id = base + lane;
					if (id < n_anch) {
						uint64_t okv = 0;
						const uint32_t flag = s_run_flag[wid][id];
						const uint32_t cnt = min(31u, (uint32_t)s_run_dmax[wid][id]);  // k-mers 1..cnt have their window inside the bucket
						if ((flag & 1) && cnt) {
							const bool same = flag & 2;
							const uint64_t Ta = s_run_T[wid][id];
							const uint64_t S0 = same ? Ta : Ta - cnt;  // first base of the index text that is compared
							uint4 R = strip_bases64(pack, s_run_q[wid][id]);
							if (!same) R = shl_bases64(rc_bases64(R), 64 - (cnt + k));  // the cnt + k bases of the run, reverse complemented
							const uint64_t A = mismatch_windows64(R, seq_bases64(I.seq, S0), k);
							// same strand: window j = k-mer j; opposite strand: window j = k-mer cnt - j
							uint32_t ok = same ? __brev((uint32_t)(~A >> 32)) : (uint32_t)(~A >> (63 - cnt));
							ok &= cnt == 31 ? 0xFFFFFFFEu : ((2u << cnt) - 2u);
							uint32_t v = 0;
							if (!kId) {
								const uint32_t* vp = I.valid + (S0 >> 5);
								const uint32_t u = __funnelshift_r(__ldg(vp), __ldg(vp + 1), (uint32_t)(S0 & 31));  // bit j = valid[S0 + j]
								v = same ? u : (__brev(u) >> (31 - cnt));
							}
							okv = ok | ((uint64_t)v << 32);
						}
						s_run_okv[wid][id] = okv;
					}
				}
				__syncwarp();
				// C3b. every position: answered by its run's masks, or left for the lookup
				#pragma unroll 1
				for (int it = 0; it < kPerLane; it++) {
					const uint32_t q = it * 32 + lane;
					const uint32_t id = runid[q];
					bool left = false;
					if (id < (uint32_t)kMaxRuns && q != s_run_q[wid][id]) {
						const uint32_t d = q - s_run_q[wid][id];
						const uint64_t okv = s_run_okv[wid][id];
						if (d < 32 && ((okv >> d) & 1)) {
							// the query equals this window, so the reference's answer is the window's own
							bool v;
							if (kId) {  // the launcher sends id queries here only when the table exists
								const uint64_t Ta = s_run_T[wid][id];
								const uint32_t pid = __ldg(I.pos_id + ((s_run_flag[wid][id] & 2) ? Ta + d : Ta - d));
								v = pid != 0xFFFFFFFFu;
								emit<MODE>(K, v ? (int64_t)(pid + I.id_base) : -1, kSlot ? s_run_o[ow][id] + d : 0);
							} else {
								v = (okv >> (32 + d)) & 1;
							}
							if (v) found++; else notfound++;
						} else {
							left = true;
						}
					}
					const uint32_t lm = __ballot_sync(0xffffffffu, left);
					if (left) s_resid[wid][n_res + __popc(lm & lt_mask)] = (uint8_t)q;
					n_res += __popc(lm);
				}
				__syncwarp();
				// C4a. the rest through the negative filter; what passes is compacted again (in place: a slot is rewritten
				// only after the batch that held it was read)
				if (I.filter) {
					uint32_t n_keep = 0;
					#pragma unroll 1
					for (uint32_t base = 0; base < n_res; base += 32) {
						const uint32_t i = base + lane;
						bool keep = false;
						uint32_t q = 0;
						if (i < n_res) {
							q = s_resid[wid][i];
							const uint64_t f = strip_kmer(pack, q, k), rc = rc64(f, k);
							keep = runid[q] == kTagOverflow || filter_maybe(I, f < rc ? f : rc);
							if (!keep) {
								notfound++;
								if (kSlot) {
									const uint32_t id = runid[q];
									emit<MODE>(K, -1, s_run_o[ow][id] + (q - s_run_q[wid][id]));
								}
							}
						}
						const uint32_t km = __ballot_sync(0xffffffffu, keep);
						__syncwarp();
						if (keep) s_resid[wid][n_keep + __popc(km & lt_mask)] = (uint8_t)q;
						n_keep += __popc(km);
					}
					n_res = n_keep;
					__syncwarp();
				}
			}
			// C2 (phase 0: the first k-mer of every run) / C4b (phase 1: what C3 and the filter left): the whole lookup
			const uint32_t n_items = phase == 0 ? n_anch : n_res;
			#pragma unroll 1
			for (uint32_t base = 0; base < n_items; base += 32) {
				const uint32_t i = base + lane;
				if (i < n_items) {
					uint32_t q, id;
					if (phase == 0) { id = i; q = s_run_q[wid][id]; }
					else { q = s_resid[wid][i]; id = runid[q]; }
					uint32_t mn;
					uint64_t o = 0;
					if (phase == 1 && id != kTagOverflow) {
						mn = s_run_mn[wid][id];
					} else if (phase == 0) {
						mn = mini_from_key(s_run_key[wid][id]);
						s_run_mn[wid][id] = mn;
					} else {
						mn = mini_from_key(window_min_slow(keys, q, w));
					}
					if (kSlot) {
						if (id != kTagOverflow) {
							o = s_run_o[ow][id] + (q - s_run_q[wid][id]);
						} else {
							const uint64_t r = find_read(read_off, n_reads, reads_per_base, t0 + q, guess_p0);
							o = __ldg(kmer_off + r) + (t0 + q - __ldg(read_off + r));
						}
					}
					const uint64_t f = strip_kmer(pack, q, k), rc = rc64(f, k);
					const uint64_t x = f < rc ? f : rc;
					uint64_t T = 0;
					const bool pass = !(phase == 0 && filter_anchors) || filter_maybe(I, x);
					const int64_t idr = pass ? lookup_one<SMALL>(I, x, mn, &T) : -1;
					if (idr >= 0) found++; else notfound++;
					if (kId) emit<MODE>(K, idr, o);
					if (phase == 0) {
						uint32_t flag = 0, dmax = 0;
						if (idr >= 0) {
							const bool same = window_at(I.seq, T, k) == f;
							flag = 1u | (same ? 2u : 0u);
							const uint4 bd = __ldg(I.bucket + mn);
							const uint64_t bstart = ((uint64_t)bd.y << 32) | bd.x;
							// the scan may match past the bucket's end (blight.cpp:732-739 never looks at the length again):
							// no prediction from there, valid[] is defined per bucket
							if (T - bstart < bd.z) dmax = (uint32_t)min(same ? bstart + bd.z - 1 - T : T - bstart, (uint64_t)0xFFFF);
						}
						s_run_T[wid][id] = T;
						s_run_flag[wid][id] = (uint8_t)flag;
						s_run_dmax[wid][id] = (uint16_t)dmax;
					}
				}
			}
			__syncwarp();
		}
	}

	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		found += __shfl_xor_sync(0xffffffffu, found, o);
		notfound += __shfl_xor_sync(0xffffffffu, notfound, o);
		invalid += __shfl_xor_sync(0xffffffffu, invalid, o);
	}
	if (lane == 0) {
		if (found) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_FOUND], (unsigned long long)found);
		if (notfound) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_NOT_FOUND], (unsigned long long)notfound);
		if (found + notfound) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_QUERIES], (unsigned long long)found + notfound);
		if (invalid) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_INVALID], (unsigned long long)invalid);
	}
}

int check(cudaError_t e) { return e == cudaSuccess ? 0 : BLIGHT_ERR_CUDA; }

int sm_count() {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	return sms;
}

// A zeroed 8-byte device counter for one launch on `stream`: persistent kernels hand their strips out on demand through it
// (front.cuh: next_item). Slots come from a per-device ring; a slot is reused 4096 launches later, long after its kernel ended.
// Null (allocation failed, or BLIGHT_TICKETS=0) = fixed stride.
constexpr int kTicketRing = 4096, kMaxDevices = 64;
unsigned long long* g_ticket_ring[kMaxDevices] = {};
std::atomic<uint32_t> g_ticket_next[kMaxDevices];
std::mutex g_ticket_mu;

unsigned long long* ticket_for_launch(cudaStream_t stream) {
	static const bool off = [] { const char* e = getenv("BLIGHT_TICKETS"); return e && e[0] == '0'; }();
	if (off) return nullptr;
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
	unsigned long long* ring = g_ticket_ring[dev];
	if (!ring) {
		std::lock_guard<std::mutex> lk(g_ticket_mu);
		ring = g_ticket_ring[dev];
		if (!ring) {
			if (cudaMalloc(reinterpret_cast<void**>(&ring), kTicketRing * 8) != cudaSuccess) { cudaGetLastError(); return nullptr; }
			g_ticket_ring[dev] = ring;
		}
	}
	unsigned long long* slot = ring + (g_ticket_next[dev].fetch_add(1, std::memory_order_relaxed) % kTicketRing);
	if (cudaMemsetAsync(slot, 0, 8, stream) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return slot;
}

template <class K>
int blocks_per_sm(K kernel) {
	int nb = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kThreads, 0) != cudaSuccess || nb < 1) nb = 1;
	return nb;
}

double rpb_of(const ReadBatch& B) { return B.rpb > 0 ? B.rpb : (double)B.n_reads / (double)(B.total_bases ? B.total_bases : 1); }

template <int MODE, bool SMALL>
void launch_reads_plain(const DevIndexView& v, uint32_t k, uint32_t m, const ReadBatch& B, uint64_t strip_lo, uint64_t strip_hi, bool al,
                        uint64_t* d_canon, uint32_t* d_mini, int64_t* d_ids, uint64_t* d_ctr, cudaStream_t stream) {
	static const int per_sm = blocks_per_sm(k_reads<MODE, SMALL>);
	const uint64_t n_strips = strip_hi - strip_lo;
	const uint64_t want = (n_strips + kWarps - 1) / kWarps;
	const uint64_t cap = (uint64_t)sm_count() * per_sm;  // persistent: one resident wave, warps stride over the strips
	const unsigned grid = (unsigned)(want < cap ? want : cap);
	k_reads<MODE, SMALL><<<grid, kThreads, 0, stream>>>(v, k, m, B.d_bases, B.d_read_off, B.d_read_end, B.d_kmer_off, B.n_reads, B.total_bases,
	                                                   strip_lo, strip_hi, al, B.d_packed, rpb_of(B), B.guess_p0, d_canon, d_mini, d_ids, d_ctr,
	                                                   grid == cap ? ticket_for_launch(stream) : nullptr);
}

// Which read kernel serves a mode. Measured on B200 (100 M-k-mer index, b=6): counting mode 1.88e10 k-mers/s with the
// super-k-mer kernel vs 1.33e10 plain; id mode 1.25e10 vs 1.31e10 (every k-mer still needs its MPHF rank, so the
// prediction saves less than its bookkeeping costs). BLIGHT_READS_KERNEL=plain|sk overrides (tuning / tests).
bool use_superkmer_kernel(bool want_ids, bool has_pos_id) {
	static const int forced = [] {
		const char* e = getenv("BLIGHT_READS_KERNEL");
		return !e ? 0 : (e[0] == 'p' ? 1 : (e[0] == 's' ? 2 : 0));
	}();
	if (want_ids && !has_pos_id) return false;  // ids without the table: every k-mer needs its MPHF rank anyway (1.25e10 vs 1.31e10 plain)
	if (forced) return forced == 2;
	return true;
}

template <int MODE, bool SMALL>
void launch_reads_sk(const DevIndexView& v, uint32_t k, uint32_t m, const ReadBatch& B, uint64_t strip_lo, uint64_t strip_hi, bool al,
                     const Sink& sink, uint64_t* d_ctr, cudaStream_t stream) {
	static const int per_sm = blocks_per_sm(k_reads_sk<MODE, SMALL>);
	const uint64_t n_strips = strip_hi - strip_lo;
	const uint64_t want = (n_strips + kWarps - 1) / kWarps;
	const uint64_t cap = (uint64_t)sm_count() * per_sm;
	const unsigned grid = (unsigned)(want < cap ? want : cap);
	k_reads_sk<MODE, SMALL><<<grid, kThreads, 0, stream>>>(v, k, m, B.d_bases, B.d_read_off, B.d_read_end, B.d_kmer_off, B.n_reads, B.total_bases,
	                                                      strip_lo, strip_hi, al, B.d_packed, rpb_of(B), B.guess_p0, sink, d_ctr,
	                                                      grid == cap ? ticket_for_launch(stream) : nullptr);  // fewer CTAs than fit: one strip per warp anyway
}

template <int MODE, bool SMALL>
void launch_reads_t(const DevIndexView& v, uint32_t k, uint32_t m, const ReadBatch& B, uint64_t strip_lo, uint64_t strip_hi, bool al,
                    uint64_t* d_canon, uint32_t* d_mini, int64_t* d_ids, uint64_t* d_ctr, cudaStream_t stream) {
	if (MODE != kEmitPairs && v.valid && k - m + 1 >= 8 && use_superkmer_kernel(MODE == kLookupIds, v.pos_id != nullptr)) {
		launch_reads_sk<MODE == kEmitPairs ? kLookupCount : MODE, SMALL>(v, k, m, B, strip_lo, strip_hi, al, Sink{d_ids, nullptr, nullptr, 0, 0}, d_ctr, stream);
		return;
	}
	launch_reads_plain<MODE, SMALL>(v, k, m, B, strip_lo, strip_hi, al, d_canon, d_mini, d_ids, d_ctr, stream);
}

}  // namespace

const char* g_last_cuda_error = "";

int launch_window_answers(const DevIndexView& I, uint64_t n_buckets, uint64_t total_nuc, uint32_t* d_valid, uint32_t* d_lid, uint32_t* d_cand,
                          uint32_t* d_filter, uint32_t filter_blocks, cudaStream_t stream) {
	if (total_nuc == 0) return 0;
	const uint64_t n_scan = total_nuc + (1ull << I.b);
	const uint64_t want = (n_scan + kThreads - 1) / kThreads;
	const uint64_t cap = (uint64_t)sm_count() * 8;
	const unsigned grid = (unsigned)(want < cap ? want : cap);
	if (I.small) k_window_answers<true><<<grid, kThreads, 0, stream>>>(I, n_buckets, total_nuc, n_scan, d_valid, d_lid, d_cand, d_filter, filter_blocks);
	else k_window_answers<false><<<grid, kThreads, 0, stream>>>(I, n_buckets, total_nuc, n_scan, d_valid, d_lid, d_cand, d_filter, filter_blocks);
	g_launches++;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) g_last_cuda_error = cudaGetErrorString(e);
	return check(e);
}

int launch_exact_positions(const DevIndexView& I, uint64_t n_buckets, uint64_t n_groups, uint64_t total_nuc, const uint32_t* d_cand,
                           const uint32_t* d_lid, uint32_t* d_claim, uint32_t* d_pos_rw, cudaStream_t stream) {
	if (total_nuc == 0) return 0;
	const uint64_t want = (total_nuc + kThreads - 1) / kThreads;
	const uint64_t cap = (uint64_t)sm_count() * 8;
	const unsigned grid = (unsigned)(want < cap ? want : cap);
	k_claim_low<false><<<grid, kThreads, 0, stream>>>(I, n_buckets, total_nuc, d_cand, d_lid, d_claim, d_pos_rw, (uint32_t)n_groups);
	k_claim_low<true><<<grid, kThreads, 0, stream>>>(I, n_buckets, total_nuc, d_cand, d_lid, d_claim, d_pos_rw, (uint32_t)n_groups);
	g_launches += 2;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) g_last_cuda_error = cudaGetErrorString(e);
	return check(e);
}

int launch_reads_sink(const DevIndexView& I, int kind, const ReadBatch& B, uint32_t* d_table, uint32_t n_colors, uint32_t color, uint32_t* d_out32,
                      uint64_t* d_ctr, cudaStream_t stream) {
	if (B.n_reads == 0 || B.total_bases == 0) return 0;
	const uint64_t strip_hi = (B.total_bases + kStrip - 1) / kStrip;
	const bool al = (reinterpret_cast<uintptr_t>(B.d_bases) & 15) == 0;
	const Sink sink{nullptr, d_table, d_out32, n_colors, color};
#define BL_SINK(MODE)                                                                                  \
	do {                                                                                                \
		if (I.small) launch_reads_sk<MODE, true>(I, I.k, I.m, B, 0, strip_hi, al, sink, d_ctr, stream); \
		else launch_reads_sk<MODE, false>(I, I.k, I.m, B, 0, strip_hi, al, sink, d_ctr, stream);        \
	} while (0)
	if (kind == 0) BL_SINK(kConsumeCount);
	else if (kind == 1) BL_SINK(kConsumeColor);
	else BL_SINK(kGatherTable);
#undef BL_SINK
	g_launches++;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) g_last_cuda_error = cudaGetErrorString(e);
	return check(e);
}

int launch_lookup_kmers(const DevIndexView& I, const uint64_t* d_canon, const uint32_t* d_mini, uint64_t n, int64_t* d_ids,
                        cudaStream_t stream) {
	if (n == 0) return 0;
	const uint64_t want = (n + kThreads - 1) / kThreads;
	const uint64_t cap = (uint64_t)sm_count() * 8;
	const unsigned grid = (unsigned)(want < cap ? want : cap);
	if (d_mini) {
		if (I.small) k_lookup_kmers<true, true><<<grid, kThreads, 0, stream>>>(I, d_canon, d_mini, n, d_ids);
		else k_lookup_kmers<true, false><<<grid, kThreads, 0, stream>>>(I, d_canon, d_mini, n, d_ids);
	} else {
		if (I.small) k_lookup_kmers<false, true><<<grid, kThreads, 0, stream>>>(I, d_canon, nullptr, n, d_ids);
		else k_lookup_kmers<false, false><<<grid, kThreads, 0, stream>>>(I, d_canon, nullptr, n, d_ids);
	}
	g_launches++;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) g_last_cuda_error = cudaGetErrorString(e);
	return check(e);
}

int launch_reads(const DevIndexView* I, uint32_t k, uint32_t m, const ReadBatch& B, uint64_t* d_canon, uint32_t* d_mini, int64_t* d_ids,
                 uint64_t* d_ctr, cudaStream_t stream, uint64_t pos_begin, uint64_t pos_end) {
	if (B.n_reads == 0 || B.total_bases == 0) return 0;
	if (pos_end > B.total_bases) pos_end = B.total_bases;
	if (pos_begin >= pos_end) return 0;
	if ((pos_begin % kStrip) != 0 || (pos_end < B.total_bases && (pos_end % kStrip) != 0)) return BLIGHT_ERR_INVALID_ARG;
	if ((B.d_bases == nullptr) == (B.d_packed == nullptr)) return BLIGHT_ERR_INVALID_ARG;  // exactly one representation
	const uint64_t strip_lo = pos_begin / kStrip, strip_hi = (pos_end + kStrip - 1) / kStrip;
	DevIndexView v{};
	if (I) v = *I;
	const bool al = (reinterpret_cast<uintptr_t>(B.d_bases) & 15) == 0;
#define BL_ARGS v, k, m, B, strip_lo, strip_hi, al, d_canon, d_mini, d_ids, d_ctr, stream
	if (!I) launch_reads_t<kEmitPairs, true>(BL_ARGS);
	else if (d_ids) { if (v.small) launch_reads_t<kLookupIds, true>(BL_ARGS); else launch_reads_t<kLookupIds, false>(BL_ARGS); }
	else { if (v.small) launch_reads_t<kLookupCount, true>(BL_ARGS); else launch_reads_t<kLookupCount, false>(BL_ARGS); }
#undef BL_ARGS
	g_launches++;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) g_last_cuda_error = cudaGetErrorString(e);
	return check(e);
}

}  // namespace blight
