// part_kernels.cu — the bucket-partitioned multi-GPU query path (BASELINE configs[4]), fused with its
// exchange: super-k-mers travel to the GPU that owns their minimizer bucket and identifiers travel back as plain
// peer-memory stores over NVLink, inside the two kernels that produce them. No staging buffers, no separate
// all-to-all of payloads, no scatter pass.
//
//   k_dispatch_runs   (on the GPU holding the reads)  front end of k_reads_sk (front.cuh: pack, minimizer keys, runs
//                     of equal minimizer = super-k-mers, kmer.h:629-693) -> one 32-byte record per run {2-bit bases of
//                     the run, output slot, minimizer, length, source rank} STORED DIRECTLY into the owner's inbox
//                     (owner = rank whose MPHF-group range holds minimizer >> lb, the reference's MPHF selection,
//                     blight.cpp:722). Slots of a (source, owner) pair are reserved with a local atomic; only the
//                     record itself crosses NVLink (~2.7 B per k-mer instead of 12 B of (canon, minimizer)).
//   k_runs_lookup     (on the owner)  per warp 32 records: the first k-mer of every run through the whole lookup
//                     (lookup.cuh), every other k-mer of the run against the ONE window next to where the first one
//                     matched (answer = pos_id / valid of that window, device_index.hpp), the rest through the
//                     negative filter and the whole lookup; identifiers are STORED DIRECTLY into the source GPU's
//                     id buffer at the run's output slot. In counting mode nothing travels back but two counters.
//
// Ordering between GPUs is the caller's (blight_b200/dist.py): one tiny NCCL all-to-all of the per-pair record counts
// between the two kernels (it is also the barrier that makes the records visible), one all-reduce of the counters at
// the end of a batch (the barrier after which every id has landed).
#include <cuda_runtime.h>

#include <cstring>

#include "capi_common.hpp"
#include "front.cuh"
#include "kernels.hpp"
#include "lookup.cuh"

namespace blight {
namespace {

constexpr int kMaxRanks = BLIGHT_MAX_RANKS;
constexpr int kRecWords = 5;                  // 4 words of bases + one zero word for the funnel
constexpr uint32_t kMaxRecKmers = 33;         // k-mers per record: bounds the residual list of a warp
constexpr int kMaxPairs = 32 * (kMaxRecKmers - 1);

struct alignas(32) RunRec {
	uint32_t bases[4];  // n + k - 1 bases, 2 bits each, first base in the high bits of bases[0]
	uint64_t o;         // output slot of the run's first k-mer in the source's id buffer
	uint32_t mn;        // minimizer (bucket) of every k-mer of the run
	uint32_t n_src;     // bits 0-7: k-mers in the run, bits 8-15: source rank
};
static_assert(sizeof(RunRec) == 32, "one record = one sector");

struct Route {
	uint32_t world, rank, lb, pad;
	uint32_t cuts[kMaxRanks + 1];
	RunRec* inbox[kMaxRanks];  // this source's region in every owner's inbox (peer pointers)
	uint64_t cap;              // records per region
};

__device__ __forceinline__ uint32_t owner_of(const Route& R, uint32_t mini) {
	const uint32_t g = mini >> R.lb;
	uint32_t o = 0;
	#pragma unroll 1
	while (o + 1 < R.world && g >= R.cuts[o + 1]) o++;
	return o;
}

// the 64 bases starting at strip position q, as four packed words
__device__ __forceinline__ uint4 strip_bases64(const uint32_t* pack, uint32_t q) {
	const uint32_t wi = q >> 4, s = 2u * (q & 15);
	uint32_t v[5];
	#pragma unroll
	for (int i = 0; i < 5; i++) v[i] = (wi + i < (uint32_t)kStripWords) ? pack[wi + i] : 0u;
	return make_uint4(__funnelshift_l(v[1], v[0], s), __funnelshift_l(v[2], v[1], s), __funnelshift_l(v[3], v[2], s), __funnelshift_l(v[4], v[3], s));
}

template <bool WANT_O>
__global__ void __launch_bounds__(kThreads, 4) k_dispatch_runs(uint32_t k, uint32_t m, const char* __restrict__ bases,
                                                             const uint64_t* __restrict__ read_off, const uint64_t* __restrict__ read_end,
                                                             const uint64_t* __restrict__ kmer_off, uint64_t n_reads, uint64_t total_bases,
                                                             uint64_t strip_lo, uint64_t strip_hi, bool aligned16, Route R,
                                                             unsigned long long* __restrict__ counts, uint64_t* __restrict__ ctr,
                                                             uint32_t* __restrict__ err) {
	__shared__ uint32_t s_pack[kWarps][kStripWords];
	__shared__ uint32_t s_bad[kWarps][kStripWords];
	__shared__ uint32_t s_keys[kWarps][kKeySlots];
	__shared__ uint64_t s_run_o[WANT_O ? kWarps : 1][kMaxRuns];
	__shared__ uint16_t s_run_q[kWarps][kMaxRuns];
	__shared__ uint32_t s_run_n[kWarps][kMaxRuns];  // k-mers of the run inside this strip
	__shared__ uint64_t s_runid8[kWarps][kStrip / 8];

	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t ow = WANT_O ? wid : 0;
	const StripSmem S{s_pack[wid], s_bad[wid], s_keys[wid], s_run_q[wid], s_run_o[ow], s_runid8[wid]};
	const uint8_t* runid = reinterpret_cast<const uint8_t*>(s_runid8[wid]);
	const uint32_t w = k - m + 1;
	const uint32_t nmax = min(64u - k + 1u, kMaxRecKmers);  // k-mers one record carries (64 bases)
	const uint64_t warp_stride = (uint64_t)gridDim.x * kWarps;
	const double reads_per_base = (double)n_reads / (double)total_bases;
	const uint32_t lt_mask = (1u << lane) - 1u;
	uint32_t invalid = 0, queries = 0;

	for (uint64_t strip = strip_lo + (uint64_t)blockIdx.x * kWarps + wid; strip < strip_hi; strip += warp_stride) {
		const uint64_t t0 = strip * kStrip;
		__syncwarp();
		s_run_n[wid][lane] = 0;
		s_run_n[wid][lane + 32] = 0;
		const uint32_t n_runs = strip_front<WANT_O>(S, lane, k, m, bases, read_off, read_end, kmer_off, n_reads, total_bases, reads_per_base,
		                                            aligned16, t0, invalid);
		// length of every run: each lane adds up its own 8 positions
		{
			const uint64_t tags = s_runid8[wid][lane];
			uint32_t cur = kTagNone, cnt = 0;
			#pragma unroll
			for (int j = 0; j < 8; j++) {
				const uint32_t tag = (uint32_t)(tags >> (8 * j)) & 0xFFu;
				if (tag != cur) {
					if (cur < (uint32_t)kMaxRuns) atomicAdd(&s_run_n[wid][cur], cnt);
					cur = tag; cnt = 0;
				}
				cnt++;
				if (tag != kTagNone) queries++;
			}
			if (cur < (uint32_t)kMaxRuns) atomicAdd(&s_run_n[wid][cur], cnt);
		}
		__syncwarp();
		const uint32_t n_tab = n_runs < (uint32_t)kMaxRuns ? n_runs : (uint32_t)kMaxRuns;
		// one turn per 32 records: runs of the table (cut into pieces of nmax k-mers), then — only when the table overflowed —
		// the surplus k-mers one by one
		const uint32_t n_extra = n_runs > (uint32_t)kMaxRuns ? (uint32_t)kStrip : 0u;
		#pragma unroll 1
		for (uint32_t base = 0; base < n_tab + n_extra; base += 32) {
			const uint32_t i = base + lane;
			uint32_t q = 0, n = 0;
			uint64_t o = 0;
			if (i < n_tab) {
				q = s_run_q[wid][i];
				n = s_run_n[wid][i];
				if (WANT_O) o = s_run_o[ow][i];
			} else if (i >= n_tab && i - n_tab < (uint32_t)kStrip && n_extra) {
				const uint32_t qq = i - n_tab;
				if (runid[qq] == kTagOverflow) {
					q = qq; n = 1;
					if (WANT_O) {
						const uint64_t r = find_read(read_off, n_reads, reads_per_base, t0 + q);
						o = __ldg(kmer_off + r) + (t0 + q - __ldg(read_off + r));
					}
				}
			}
			uint32_t mn = 0, dst = 0;
			if (n) {
				mn = mini_from_key(window_min_slow(S.keys, q, w));
				dst = owner_of(R, mn);
			}
			#pragma unroll 1
			for (uint32_t off = 0; __any_sync(0xffffffffu, off < n); off += nmax) {
				const bool act = off < n;
				const uint32_t peers = __match_any_sync(0xffffffffu, act ? dst : 0xFFFFu);
				if (act) {
					const uint32_t leader = __ffs(peers) - 1;
					unsigned long long slot0 = 0;
					if (lane == leader) slot0 = atomicAdd(&counts[dst], (unsigned long long)__popc(peers));
					slot0 = __shfl_sync(peers, slot0, leader);
					const uint64_t slot = slot0 + __popc(peers & lt_mask);
					if (slot < R.cap) {
						const uint32_t nn = min(n - off, nmax);
						const uint4 b = strip_bases64(S.pack, q + off);
						uint4* dstp = reinterpret_cast<uint4*>(R.inbox[dst] + slot);
						const uint64_t oo = o + off;
						dstp[0] = b;
						dstp[1] = make_uint4((uint32_t)oo, (uint32_t)(oo >> 32), mn, nn | (R.rank << 8));
					} else {
						atomicOr(err, 1u);
					}
				}
			}
		}
	}
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		invalid += __shfl_xor_sync(0xffffffffu, invalid, o);
		queries += __shfl_xor_sync(0xffffffffu, queries, o);
	}
	if (lane == 0) {
		if (queries) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_QUERIES], (unsigned long long)queries);
		if (invalid) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_INVALID], (unsigned long long)invalid);
	}
}

struct OwnerArgs {
	uint32_t world, pad;
	const RunRec* region[kMaxRanks];  // records received from every source
	int64_t* out[kMaxRanks];          // id buffer of every source (peer pointers), unused in counting mode
};


// k-mer at offset d of a record's bases
__device__ __forceinline__ uint64_t rec_kmer(const uint32_t* W, uint32_t d, uint32_t k) {
	const uint32_t wi = d >> 4, s = 2u * (d & 15);
	const uint32_t a = W[wi], b = W[wi + 1], c = W[wi + 2];
	return (((uint64_t)__funnelshift_l(b, a, s) << 32) | __funnelshift_l(c, b, s)) >> (64 - 2 * k);
}

template <bool WANT_IDS, bool SMALL>
__global__ void __launch_bounds__(kThreads, 4) k_runs_lookup(DevIndexView I, OwnerArgs A, const unsigned long long* __restrict__ counts,
                                                           uint64_t* __restrict__ ctr) {
	__shared__ unsigned long long s_pref[kMaxRanks + 1];
	__shared__ uint32_t s_w[kWarps][32][kRecWords + 1];  // +1: odd stride
	__shared__ uint64_t s_T[kWarps][32];
	__shared__ uint64_t s_o[WANT_IDS ? kWarps : 1][32];
	__shared__ uint32_t s_mn[kWarps][32];
	__shared__ uint16_t s_dmax[kWarps][32];
	__shared__ uint16_t s_incl[kWarps][32];  // inclusive prefix of (n - 1): k-mers after the first, flattened over the 32 runs
	__shared__ uint8_t s_n[kWarps][32], s_flag[kWarps][32], s_src[kWarps][32];
	__shared__ uint16_t s_res[kWarps][kMaxPairs];  // (run << 8) | d of the k-mers left for the whole lookup

	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t ow = WANT_IDS ? wid : 0;
	const uint32_t k = I.k;
	const uint32_t lt_mask = (1u << lane) - 1u;
	const bool filter_anchors = I.filter && (I.flags & kFlagFilterAnchors);
	if (threadIdx.x == 0) {
		unsigned long long acc = 0;
		for (uint32_t s = 0; s < A.world; s++) { s_pref[s] = acc; acc += counts[s]; }
		s_pref[A.world] = acc;
	}
	__syncthreads();
	const uint64_t total = s_pref[A.world];
	const uint64_t n_chunks = (total + 31) / 32;
	const uint64_t warp_stride = (uint64_t)gridDim.x * kWarps;
	uint32_t found = 0, notfound = 0;

	for (uint64_t chunk = (uint64_t)blockIdx.x * kWarps + wid; chunk < n_chunks; chunk += warp_stride) {
		const uint64_t g = chunk * 32 + lane;
		__syncwarp();
		// the warp's 32 records into shared memory
		uint32_t n = 0;
		if (g < total) {
			uint32_t s = 0;
			while (s + 1 < A.world && g >= s_pref[s + 1]) s++;
			const uint4* rp = reinterpret_cast<const uint4*>(A.region[s] + (g - s_pref[s]));
			const uint4 b = __ldcs(rp), h = __ldcs(rp + 1);
			uint32_t* W = s_w[wid][lane];
			W[0] = b.x; W[1] = b.y; W[2] = b.z; W[3] = b.w; W[4] = 0;
			if (WANT_IDS) s_o[ow][lane] = ((uint64_t)h.y << 32) | h.x;
			s_mn[wid][lane] = h.z;
			n = h.w & 0xFFu;
			s_src[wid][lane] = (uint8_t)((h.w >> 8) & 0xFFu);
		}
		s_n[wid][lane] = (uint8_t)n;
		uint32_t incl = n ? n - 1 : 0;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= (uint32_t)o) incl += t;
		}
		s_incl[wid][lane] = (uint16_t)incl;
		const uint32_t n_pairs = __shfl_sync(0xffffffffu, incl, 31);
		const uint32_t n_first = (uint32_t)min((uint64_t)32, total - chunk * 32);
		__syncwarp();

		uint32_t n_res = 0;
		#pragma unroll 1
		for (int phase = 0; phase < 2; phase++) {
			if (phase == 1) {
				// every other k-mer of a run: the one window next to where the first one matched
				#pragma unroll 1
				for (uint32_t base = 0; base < n_pairs; base += 32) {
					const uint32_t i = base + lane;
					bool left = false;
					uint32_t run = 0, d = 0;
					if (i < n_pairs) {
						// first run whose inclusive prefix exceeds i
						uint32_t lo = 0, hi = 31;
						#pragma unroll
						for (int s = 0; s < 5; s++) {
							const uint32_t mid = (lo + hi) >> 1;
							if (s_incl[wid][mid] > i) hi = mid; else lo = mid + 1;
						}
						run = lo;
						d = i - (run ? s_incl[wid][run - 1] : 0u) + 1;
						left = true;
						const uint32_t flag = s_flag[wid][run];
						if ((flag & 1) && d <= s_dmax[wid][run]) {
							const bool same = flag & 2;
							const uint64_t Ta = s_T[wid][run];
							const uint64_t Tp = same ? Ta + d : Ta - d;
							const uint64_t f = rec_kmer(s_w[wid][run], d, k), rc = rc64(f, k);
							if (window_at(I.seq, Tp, k) == (same ? f : rc)) {
								left = false;
								bool v;
								int64_t idr = -1;
								if (WANT_IDS) {
									const uint32_t pid = __ldg(I.pos_id + Tp);
									v = pid != 0xFFFFFFFFu;
									if (v) idr = (int64_t)pid;
								} else {
									v = (__ldg(I.valid + (Tp >> 5)) >> (Tp & 31)) & 1u;
								}
								if (v) found++; else notfound++;
								if (WANT_IDS) A.out[s_src[wid][run]][s_o[ow][run] + d] = idr;
							}
						}
					}
					const uint32_t lm = __ballot_sync(0xffffffffu, left);
					if (left) s_res[wid][n_res + __popc(lm & lt_mask)] = (uint16_t)((run << 8) | d);
					n_res += __popc(lm);
				}
				__syncwarp();
				if (I.filter) {
					uint32_t n_keep = 0;
					#pragma unroll 1
					for (uint32_t base = 0; base < n_res; base += 32) {
						const uint32_t i = base + lane;
						bool keep = false;
						uint32_t it = 0;
						if (i < n_res) {
							it = s_res[wid][i];
							const uint32_t run = it >> 8, d = it & 0xFFu;
							const uint64_t f = rec_kmer(s_w[wid][run], d, k), rc = rc64(f, k);
							keep = filter_maybe(I, f < rc ? f : rc);
							if (!keep) {
								notfound++;
								if (WANT_IDS) A.out[s_src[wid][run]][s_o[ow][run] + d] = -1ll;
							}
						}
						const uint32_t km = __ballot_sync(0xffffffffu, keep);
						__syncwarp();
						if (keep) s_res[wid][n_keep + __popc(km & lt_mask)] = (uint16_t)it;
						n_keep += __popc(km);
					}
					n_res = n_keep;
					__syncwarp();
				}
			}
			// phase 0: the first k-mer of every run; phase 1: what the prediction and the filter left. One copy of the lookup.
			const uint32_t n_items = phase == 0 ? n_first : n_res;
			#pragma unroll 1
			for (uint32_t base = 0; base < n_items; base += 32) {
				const uint32_t i = base + lane;
				if (i < n_items) {
					uint32_t run, d;
					if (phase == 0) { run = i; d = 0; }
					else { const uint32_t it = s_res[wid][i]; run = it >> 8; d = it & 0xFFu; }
					const uint32_t mn = s_mn[wid][run];
					const uint64_t f = rec_kmer(s_w[wid][run], d, k), rc = rc64(f, k);
					const uint64_t x = f < rc ? f : rc;
					uint64_t T = 0;
					const bool pass = !(phase == 0 && filter_anchors) || filter_maybe(I, x);
					const int64_t idr = pass ? lookup_one<SMALL>(I, x, mn, &T) : -1;
					if (idr >= 0) found++; else notfound++;
					if (WANT_IDS) A.out[s_src[wid][run]][s_o[ow][run] + d] = idr;
					if (phase == 0) {
						uint32_t flag = 0, dmax = 0;
						if (idr >= 0) {
							const bool same = window_at(I.seq, T, k) == f;
							flag = 1u | (same ? 2u : 0u);
							const uint4 bd = __ldg(I.bucket + mn);
							const uint64_t bstart = ((uint64_t)bd.y << 32) | bd.x;
							if (T - bstart < bd.z) dmax = (uint32_t)min(same ? bstart + bd.z - 1 - T : T - bstart, (uint64_t)0xFFFF);
						}
						s_T[wid][run] = T;
						s_flag[wid][run] = (uint8_t)flag;
						s_dmax[wid][run] = (uint16_t)dmax;
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
	}
	if (lane == 0) {
		if (found) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_FOUND], (unsigned long long)found);
		if (notfound) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_NOT_FOUND], (unsigned long long)notfound);
	}
}

int sm_count_() {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	return sms;
}

template <class K>
int per_sm(K kernel) {
	int nb = 0;
	if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, kThreads, 0) != cudaSuccess || nb < 1) nb = 1;
	return nb;
}

int finish(const char* what) {
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) return fail(BL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
	return BL_OK;
}

}  // namespace
}  // namespace blight

using namespace blight;

extern "C" {

int blight_part_dispatch(uint32_t k, uint32_t m, const char* d_bases, const uint64_t* d_read_off, const uint64_t* d_kmer_off,
                         uint64_t n_reads, uint64_t total_bases, uint64_t pos_begin, uint64_t pos_end, const blight_part_route* route,
                         uint64_t* d_counts, uint64_t* d_ctr, uint32_t* d_err, void* stream) {
	if (!route || !d_counts || !d_ctr || !d_err || (n_reads && (!d_bases || !d_read_off))) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (route->world == 0 || route->world > (uint32_t)kMaxRanks || route->rank >= route->world) return fail(BL_ERR_INVALID_ARG, "bad world / rank");
	if (k < 8 || k > 32 || m >= k || k - m + 1 < 8 || k - m + 1 > 32) return fail(BL_ERR_INVALID_ARG, "partition mode needs 8 <= k <= 32 and 8 <= k-m+1 <= 32");
	if (n_reads == 0 || total_bases == 0) return BL_OK;
	if (pos_end > total_bases) pos_end = total_bases;
	if (pos_begin >= pos_end) return BL_OK;
	if ((pos_begin % kStrip) != 0 || (pos_end < total_bases && (pos_end % kStrip) != 0)) return fail(BL_ERR_INVALID_ARG, "sub-batch bounds must be multiples of 256");
	Route R{};
	R.world = route->world; R.rank = route->rank; R.lb = route->lb; R.cap = route->cap;
	for (uint32_t i = 0; i <= route->world; i++) R.cuts[i] = route->cuts[i];
	for (uint32_t i = 0; i < route->world; i++) {
		if (!route->inbox[i]) return fail(BL_ERR_INVALID_ARG, "null inbox pointer");
		R.inbox[i] = static_cast<RunRec*>(route->inbox[i]);
	}
	const uint64_t strip_lo = pos_begin / kStrip, strip_hi = (pos_end + kStrip - 1) / kStrip;
	const bool al = (reinterpret_cast<uintptr_t>(d_bases) & 15) == 0;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	const uint64_t want = (strip_hi - strip_lo + kWarps - 1) / kWarps;
	if (d_kmer_off) {
		static const int nb = per_sm(k_dispatch_runs<true>);
		const uint64_t cap = (uint64_t)sm_count_() * nb;
		k_dispatch_runs<true><<<(unsigned)(want < cap ? want : cap), kThreads, 0, st>>>(k, m, d_bases, d_read_off, nullptr, d_kmer_off, n_reads, total_bases,
			strip_lo, strip_hi, al, R, reinterpret_cast<unsigned long long*>(d_counts), d_ctr, d_err);
	} else {
		static const int nb = per_sm(k_dispatch_runs<false>);
		const uint64_t cap = (uint64_t)sm_count_() * nb;
		k_dispatch_runs<false><<<(unsigned)(want < cap ? want : cap), kThreads, 0, st>>>(k, m, d_bases, d_read_off, nullptr, nullptr, n_reads, total_bases,
			strip_lo, strip_hi, al, R, reinterpret_cast<unsigned long long*>(d_counts), d_ctr, d_err);
	}
	g_launches++;
	return finish("k_dispatch_runs");
}

int blight_part_lookup(const blight_index* idx, uint32_t world, const void* const* regions, const uint64_t* d_counts, int64_t* const* out,
                       uint64_t max_records, uint64_t* d_ctr, void* stream) {
	if (!idx || !regions || !d_counts || !d_ctr) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (world == 0 || world > (uint32_t)kMaxRanks) return fail(BL_ERR_INVALID_ARG, "bad world");
	if (!idx->v.valid) return fail(BL_ERR_INVALID_ARG, "index has no valid-window bitmap");
	if (out && !idx->v.pos_id) return fail(BL_ERR_INVALID_ARG, "id mode of the partitioned path needs the position->id table (BLIGHT_POS_ID)");
	if (idx->v.k < 8) return fail(BL_ERR_INVALID_ARG, "partition mode needs k >= 8");
	OwnerArgs A{};
	A.world = world;
	for (uint32_t i = 0; i < world; i++) {
		A.region[i] = static_cast<const RunRec*>(regions[i]);
		A.out[i] = out ? out[i] : nullptr;
	}
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	const uint64_t want = std::max<uint64_t>(1, ((max_records + 31) / 32 + kWarps - 1) / kWarps);
	const unsigned long long* cnt = reinterpret_cast<const unsigned long long*>(d_counts);
#define BL_LAUNCH(IDS, SM)                                                                              \
	do {                                                                                                 \
		static const int nb = per_sm(k_runs_lookup<IDS, SM>);                                            \
		const uint64_t cap = (uint64_t)sm_count_() * nb;                                                 \
		k_runs_lookup<IDS, SM><<<(unsigned)(want < cap ? want : cap), kThreads, 0, st>>>(idx->v, A, cnt, d_ctr); \
	} while (0)
	if (out) { if (idx->v.small) BL_LAUNCH(true, true); else BL_LAUNCH(true, false); }
	else { if (idx->v.small) BL_LAUNCH(false, true); else BL_LAUNCH(false, false); }
#undef BL_LAUNCH
	g_launches++;
	return finish("k_runs_lookup");
}

int blight_peer_alloc(uint64_t bytes, void** d_ptr, unsigned char* handle64) {
	if (!d_ptr || !handle64 || bytes == 0) return fail(BL_ERR_INVALID_ARG, "bad argument");
	void* p = nullptr;
	cudaError_t e = cudaMalloc(&p, bytes);
	if (e != cudaSuccess) return fail(BL_ERR_NOMEM, std::string("cudaMalloc(peer buffer): ") + cudaGetErrorString(e));
	cudaIpcMemHandle_t h;
	e = cudaIpcGetMemHandle(&h, p);
	if (e != cudaSuccess) { cudaFree(p); return fail(BL_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
	static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
	std::memcpy(handle64, &h, 64);
	*d_ptr = p;
	return BL_OK;
}

int blight_peer_open(const unsigned char* handle64, void** d_ptr) {
	if (!handle64 || !d_ptr) return fail(BL_ERR_INVALID_ARG, "null argument");
	cudaIpcMemHandle_t h;
	std::memcpy(&h, handle64, 64);
	void* p = nullptr;
	cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
	if (e != cudaSuccess) return fail(BL_ERR_CUDA, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e));
	*d_ptr = p;
	return BL_OK;
}

int blight_peer_close(void* d_ptr) {
	if (!d_ptr) return BL_OK;
	cudaError_t e = cudaIpcCloseMemHandle(d_ptr);
	return e == cudaSuccess ? BL_OK : fail(BL_ERR_CUDA, std::string("cudaIpcCloseMemHandle: ") + cudaGetErrorString(e));
}

int blight_peer_free(void* d_ptr) {
	if (!d_ptr) return BL_OK;
	cudaError_t e = cudaFree(d_ptr);
	return e == cudaSuccess ? BL_OK : fail(BL_ERR_CUDA, std::string("cudaFree: ") + cudaGetErrorString(e));
}

}  // extern "C"
