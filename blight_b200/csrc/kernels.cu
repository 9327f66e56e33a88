// kernels.cu — sm_100a kernels of the batched query path and their launchers.
//
//   k_lookup_kmers   canonical k-mers (+ optional minimizers) -> ids          query_kmer_hash, blight.cpp:545-550
//   k_reads          ASCII reads -> 2-bit pack -> canonical k-mers + rolling  query_sequence_hash/bool,
//                    minimizers -> {pairs | ids | counts}                     blight.cpp:554-591; kmer.h:791-810
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

#include "kernels.hpp"
#include "lookup.cuh"

namespace blight {

std::atomic<uint64_t> g_launches{0};

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kStrip = 256;                          // k-mer start positions per warp strip
constexpr int kPerLane = kStrip / 32;
constexpr int kMaxW = 32;                            // k - m + 1 <= 31
constexpr int kStripWords = (kStrip + 32) / 16 + 2;  // packed words per strip (+halo k-1 <= 30, +1 for the funnel)
constexpr int kStripKeys = kStrip + kMaxW;
constexpr int kQueue = 64;                           // straggler slots per warp (at most 31 waiting + 32 new)
static_assert(kStripWords <= 32, "one lane packs one word");

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

enum ReadsMode { kEmitPairs = 0, kLookupIds = 1, kLookupCount = 2 };

// first read r in [0, n_reads) with off[r+1] > p, i.e. the read containing base position p (or the gap before it).
// Starts from the position a uniform read length would give and brackets the answer exponentially: 2-3 loads for
// the usual near-uniform batches, a plain binary search in the worst case.
__device__ __forceinline__ uint64_t find_read(const uint64_t* __restrict__ off, uint64_t n_reads, double reads_per_base, uint64_t p) {
	uint64_t g = (uint64_t)((double)p * reads_per_base);
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

// EAGER = BBHash levels probed in lock step before a k-mer is parked (kLevels: never park)
template <int MODE, bool SMALL, int EAGER>
__global__ void __launch_bounds__(kThreads, 4) k_reads(DevIndexView I, uint32_t k, uint32_t m, const char* __restrict__ bases,
                                                    const uint64_t* __restrict__ read_off, const uint64_t* __restrict__ read_end,
                                                    const uint64_t* __restrict__ kmer_off, uint64_t n_reads, uint64_t total_bases,
                                                    uint64_t strip_lo, uint64_t strip_hi, bool aligned16, uint64_t* __restrict__ out_canon,
                                                    uint32_t* __restrict__ out_mini, int64_t* __restrict__ out_ids,
                                                    uint64_t* __restrict__ ctr) {
	__shared__ uint32_t s_pack[kWarps][kStripWords];  // 2-bit codes, 16 per word, first base in the high bits
	__shared__ uint32_t s_bad[kWarps][kStripWords];   // 1 bit per base (bit 15-j of word i = base 16i+j): not ACGTacgt
	__shared__ uint32_t s_keys[kWarps][kStripKeys];   // ordering key of the m-mer starting at each strip position
	// Straggler queue of the warp: k-mers that no bit accommodated in BBHash levels 0..EAGER-1 (~22 % of a read
	// batch). A warp would otherwise iterate the level loop to its slowest lane with a handful of lanes active; such
	// k-mers are parked here with their hasher state and finished 32 at a time with every lane busy.
	constexpr bool kParks = MODE != kEmitPairs && EAGER < kLevels;
	__shared__ uint64_t q_x[kParks ? kWarps : 1][kParks ? kQueue : 1], q_s0[kParks ? kWarps : 1][kParks ? kQueue : 1];
	__shared__ uint64_t q_s1[kParks ? kWarps : 1][kParks ? kQueue : 1], q_off[kParks ? kWarps : 1][kParks ? kQueue : 1];
	__shared__ uint64_t q_o[kParks && MODE == kLookupIds ? kWarps : 1][kParks ? kQueue : 1];
	__shared__ uint32_t q_mn[kParks ? kWarps : 1][kParks ? kQueue : 1];

	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	uint32_t* pack = s_pack[wid];
	uint32_t* bad = s_bad[wid];
	uint32_t* keys = s_keys[wid];
	const uint32_t w = k - m + 1;
	const uint32_t mmask = (1u << (2 * m)) - 1u;
	const uint64_t n_strips = strip_hi;  // this launch handles strips [strip_lo, strip_hi) of the buffer
	const uint64_t warp_stride = (uint64_t)gridDim.x * kWarps;
	const double reads_per_base = (double)n_reads / (double)total_bases;
	uint32_t found = 0, notfound = 0, invalid = 0;
	uint32_t q_head = 0, q_count = 0;  // warp-uniform
	const uint32_t qw = kParks ? wid : 0;

	uint64_t strip = strip_lo + (uint64_t)blockIdx.x * kWarps + wid;
	uint64_t t0 = 0, r = 0;
	uint32_t n_pos = 0;
	int it = 0;
	for (;;) {
		const bool flush = strip >= n_strips;  // one extra turn after the last strip empties the queue
		if (!flush && it == 0) {
			t0 = strip * kStrip;
			n_pos = (uint32_t)min((uint64_t)kStrip, total_bases - t0);                        // positions owned by this strip
			const uint32_t n_load = (uint32_t)min((uint64_t)(kStrip + 32), total_bases - t0);  // owned + halo (k-1 <= 30)
			__syncwarp();
			// A. pack: lane i converts bases [16i, 16i+16) of the strip
			if (lane < kStripWords) {
				const uint32_t b0 = lane * 16;
				uint32_t word = 0, badw = 0;
				if (b0 < n_load) {
					unsigned char ch[16];
					if (aligned16 && b0 + 16 <= n_load) {
						const uint4 v = __ldcs(reinterpret_cast<const uint4*>(bases + t0 + b0));  // t0 % 256 == 0
						*reinterpret_cast<uint4*>(ch) = v;
					} else {
						#pragma unroll
						for (int j = 0; j < 16; j++) ch[j] = (b0 + j < n_load) ? (unsigned char)bases[t0 + b0 + j] : (unsigned char)'A';
					}
					#pragma unroll
					for (int j = 0; j < 16; j++) {
						const uint32_t c = nuc_code(ch[j]);
						badw = (badw << 1) | (c >> 2);
						word = (word << 2) | (c & 3u);
					}
				}
				pack[lane] = word;
				bad[lane] = badw;
			}
			// read containing the first position of the strip (lane 0 searches, everybody starts from there)
			if (lane == 0) r = find_read(read_off, n_reads, reads_per_base, t0);
			r = __shfl_sync(0xffffffffu, r, 0);
			__syncwarp();
			// B. m-mer keys
			for (uint32_t q = lane; q < n_pos + w - 1; q += 32) {
				const uint32_t wi = q >> 4, s = 2u * (q & 15);
				const uint32_t v = __funnelshift_l(pack[wi + 1], pack[wi], s) >> (32 - 2 * m);
				keys[q] = mini_key(parity_canon(v & mmask, m));
			}
			__syncwarp();
		}

		// C. one k-mer per lane: adjacent lanes, adjacent positions
		bool have = false;
		uint64_t hx = 0, ho = 0;
		uint32_t hmn = 0;
		const uint32_t q = it * 32 + lane;
		if (!flush && q < n_pos) {
			const uint64_t p = t0 + q;
			while (r + 1 < n_reads && __ldg(read_off + r + 1) <= p) r++;  // positions only grow: gallop forward
			const uint64_t rbeg = __ldg(read_off + r), rend = read_end ? __ldg(read_end + r) : __ldg(read_off + r + 1);
			if (p >= rbeg && p + k <= rend) {  // else no k-mer starts here (tail of a read, a gap, a read shorter than k)
				const uint32_t wi = q >> 4, s = 2u * (q & 15);
				// nuc2int rejects any byte outside ACGTacgt (kmer.h:56-69); only bases of queried k-mers are ever looked at
				const uint64_t bb = ((uint64_t)bad[wi] << 32) | ((uint64_t)bad[wi + 1] << 16) | bad[wi + 2];
				if ((bb >> (48 - (q & 15) - k)) & ((1ull << k) - 1)) {
					invalid++;
				} else {
					uint32_t best = keys[q];
					for (uint32_t j = 1; j < w; j++) best = min(best, keys[q + j]);
					const uint32_t a = pack[wi], b = pack[wi + 1], c = pack[wi + 2];
					const uint64_t top = ((uint64_t)__funnelshift_l(b, a, s) << 32) | __funnelshift_l(c, b, s);
					const uint64_t fwd = top >> (64 - 2 * k);
					const uint64_t rc = rc64(fwd, k);
					hx = fwd < rc ? fwd : rc;
					hmn = mini_from_key(best);
					if (MODE != kLookupCount) ho = __ldg(kmer_off + r) + (p - rbeg);
					have = true;
				}
			}
		}
		if (MODE == kEmitPairs) {
			if (have) {
				__stcs(reinterpret_cast<unsigned long long*>(out_canon + ho), (unsigned long long)hx);
				__stcs(out_mini + ho, hmn);
			}
		} else {
			// the strip's k-mers: bucket, then levels 0..EAGER-1 in lock step
			bool go = false, hit = false, park = false;
			BucketRef B;
			uint32_t sw[8];
			uint32_t sr = 0;
			uint64_t s0 = 0, s1 = 0, off = 0;
			if (have) {
				B = load_bucket(I, hmn);
				if (B.bd.z == 0) {  // empty bucket (blight.cpp:719)
					notfound++;
					if (MODE == kLookupIds) __stcs(reinterpret_cast<long long*>(out_ids + ho), -1ll);
				} else {
					hit = probe_levels<SMALL>(B, hx, 0, EAGER, s0, s1, off, sw, sr);
					go = hit || !kParks;
					park = !hit && kParks;
				}
			}
			const uint32_t pm = kParks ? __ballot_sync(0xffffffffu, park) : 0u;
			if (kParks && park) {
				const uint32_t slot = (q_head + q_count + __popc(pm & ((1u << lane) - 1u))) & (kQueue - 1);
				q_x[qw][slot] = hx; q_s0[qw][slot] = s0; q_s1[qw][slot] = s1; q_off[qw][slot] = off; q_mn[qw][slot] = hmn;
				if (MODE == kLookupIds) q_o[qw][slot] = ho;
			}
			q_count += __popc(pm);
			__syncwarp();
			// finish the eager hits, then whole warps of parked k-mers while there are enough of them
			for (;;) {
				if (go) {
					const int64_t id = finish_lookup(I, B, hx, hit, sw, sr);
					if (MODE == kLookupIds) __stcs(reinterpret_cast<long long*>(out_ids + ho), (long long)id);
					if (id >= 0) found++; else notfound++;
				}
				if (!kParks || q_count < (flush ? 1u : 32u)) break;
				const uint32_t take = q_count < 32 ? q_count : 32;
				go = lane < take;
				if (go) {
					const uint32_t slot = (q_head + lane) & (kQueue - 1);
					hx = q_x[qw][slot]; hmn = q_mn[qw][slot];
					s0 = q_s0[qw][slot]; s1 = q_s1[qw][slot]; off = q_off[qw][slot];
					if (MODE == kLookupIds) ho = q_o[qw][slot];
					B = load_bucket(I, hmn);
					hit = probe_levels<SMALL>(B, hx, EAGER, kLevels, s0, s1, off, sw, sr);
				}
				q_head = (q_head + take) & (kQueue - 1);
				q_count -= take;
				__syncwarp();
			}
		}
		if (flush) break;
		if (++it == kPerLane) { it = 0; strip += warp_stride; }
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

template <class K>
int blocks_per_sm(K kernel) {
	int nb = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kThreads, 0) != cudaSuccess || nb < 1) nb = 1;
	return nb;
}

int eager_levels() {
	static const int v = [] {
		const char* e = getenv("BLIGHT_EAGER_LEVELS");  // tuning knob: 2, 3 or 16 (never park)
		const int x = e ? atoi(e) : 16;
		return (x == 2 || x == 3) ? x : 16;
	}();
	return v;
}

template <int MODE, bool SMALL, int EAGER>
void launch_reads_e(const DevIndexView& v, uint32_t k, uint32_t m, const char* d_bases, const uint64_t* d_read_off,
                    const uint64_t* d_read_end, const uint64_t* d_kmer_off, uint64_t n_reads, uint64_t total_bases,
                    uint64_t strip_lo, uint64_t strip_hi, bool al, uint64_t* d_canon, uint32_t* d_mini, int64_t* d_ids, uint64_t* d_ctr,
                    cudaStream_t stream) {
	static const int per_sm = blocks_per_sm(k_reads<MODE, SMALL, EAGER>);
	const uint64_t n_strips = strip_hi - strip_lo;
	const uint64_t want = (n_strips + kWarps - 1) / kWarps;
	const uint64_t cap = (uint64_t)sm_count() * per_sm;  // persistent: one resident wave, warps stride over the strips
	const unsigned grid = (unsigned)(want < cap ? want : cap);
	k_reads<MODE, SMALL, EAGER><<<grid, kThreads, 0, stream>>>(v, k, m, d_bases, d_read_off, d_read_end, d_kmer_off, n_reads, total_bases,
	                                                          strip_lo, strip_hi, al, d_canon, d_mini, d_ids, d_ctr);
}

template <int MODE, bool SMALL>
void launch_reads_t(const DevIndexView& v, uint32_t k, uint32_t m, const char* d_bases, const uint64_t* d_read_off,
                    const uint64_t* d_read_end, const uint64_t* d_kmer_off, uint64_t n_reads, uint64_t total_bases,
                    uint64_t strip_lo, uint64_t strip_hi, bool al, uint64_t* d_canon, uint32_t* d_mini, int64_t* d_ids, uint64_t* d_ctr,
                    cudaStream_t stream) {
	const int e = MODE == kEmitPairs ? 16 : eager_levels();
	if (e == 2) launch_reads_e<MODE, SMALL, 2>(v, k, m, d_bases, d_read_off, d_read_end, d_kmer_off, n_reads, total_bases, strip_lo, strip_hi, al, d_canon, d_mini, d_ids, d_ctr, stream);
	else if (e == 3) launch_reads_e<MODE, SMALL, 3>(v, k, m, d_bases, d_read_off, d_read_end, d_kmer_off, n_reads, total_bases, strip_lo, strip_hi, al, d_canon, d_mini, d_ids, d_ctr, stream);
	else launch_reads_e<MODE, SMALL, kLevels>(v, k, m, d_bases, d_read_off, d_read_end, d_kmer_off, n_reads, total_bases, strip_lo, strip_hi, al, d_canon, d_mini, d_ids, d_ctr, stream);
}

}  // namespace

const char* g_last_cuda_error = "";

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

int launch_reads(const DevIndexView* I, uint32_t k, uint32_t m, const char* d_bases, const uint64_t* d_read_off,
                 const uint64_t* d_read_end, const uint64_t* d_kmer_off, uint64_t n_reads, uint64_t total_bases,
                 uint64_t* d_canon, uint32_t* d_mini, int64_t* d_ids, uint64_t* d_ctr, cudaStream_t stream, uint64_t pos_begin,
                 uint64_t pos_end) {
	if (n_reads == 0 || total_bases == 0) return 0;
	if (pos_end > total_bases) pos_end = total_bases;
	if (pos_begin >= pos_end) return 0;
	if ((pos_begin % kStrip) != 0 || (pos_end < total_bases && (pos_end % kStrip) != 0)) return BLIGHT_ERR_INVALID_ARG;
	const uint64_t strip_lo = pos_begin / kStrip, strip_hi = (pos_end + kStrip - 1) / kStrip;
	DevIndexView v{};
	if (I) v = *I;
	const bool al = (reinterpret_cast<uintptr_t>(d_bases) & 15) == 0;
#define BL_ARGS v, k, m, d_bases, d_read_off, d_read_end, d_kmer_off, n_reads, total_bases, strip_lo, strip_hi, al, d_canon, d_mini, d_ids, d_ctr, stream
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
