// part_kernels.cu — the bucket-partitioned multi-GPU query path (BASELINE configs[4]), fused with its
// exchange: super-k-mers travel to the GPU that owns their minimizer bucket and identifiers travel back as plain
// peer-memory stores over NVLink, issued by the kernels that produce them. No NCCL payload collectives, no staging in HBM.
//
//   k_dispatch_runs   (on the GPU holding the reads)  front end of k_reads_sk (front.cuh: pack, minimizer keys, runs
//                     of equal minimizer = super-k-mers, kmer.h:629-693) -> one 32-byte record per run {2-bit bases of
//                     the run, k-mer offset in the (source, owner) stream, minimizer, length, source rank} STORED
//                     DIRECTLY into the owner's inbox (owner = rank whose MPHF-group range holds minimizer >> lb, the
//                     reference's MPHF selection, blight.cpp:722). Record slot and k-mer offset of a (source, owner) pair
//                     are reserved together with ONE local 64-bit atomic, so k-mer offsets are consecutive in slot order.
//                     Only the record crosses NVLink (~2.7 B per k-mer instead of 12 B of (canon, minimizer)); where
//                     the run's ids must finally go stays at the source, in a side table.
//   k_runs_lookup     (on the owner)  per warp 32 consecutive records of one source: the first k-mer of every run
//                     through the whole lookup (lookup.cuh), every other k-mer of the run against the ONE window next
//                     to where the first one matched (answer = pos_id / valid of that window, device_index.hpp), the
//                     rest through the negative filter and the whole lookup. The warp's ids are collected in shared memory
//                     and return one of three ways (part_session.cu): int64 ids STORED DIRECTLY into the source GPU's id
//                     array, run by run (default); or as one contiguous, sector-aligned stream of 32-bit ids into the
//                     source's return region (or the owner's own memory: pull), widened into read order by
//                     k_scatter_runs. In counting mode nothing travels back but two counters.
//                     Owners take their sources ROUND-ROBIN, starting at their right-hand neighbour: taking them one after
//                     the other made all owners (they run in step) return ids to the SAME GPU at any moment.
//   k_scatter_runs    (back on the source, stream / pull return)  return streams -> int64 ids in read order, through the
//                     side table.
//
// Both persistent kernels hand their work items out on demand (front.cuh: next_item) when the caller passes a zeroed ticket
// counter. Ordering between GPUs is the session's (part_session.cu): device-side flags in peer memory, no collective call on
// the data path; the round-1 Python pipeline (blight_b200/dist.py, BLIGHT_PART_PIPELINE=legacy: one tiny NCCL all-to-all of
// the per-pair counters per sub-batch) still drives the same kernels through the C ABI entry points below.
#include <cuda_runtime.h>

#include <cstring>

#include "capi_common.hpp"
#include "front.cuh"
#include "kernels.hpp"
#include "lookup.cuh"

namespace blight {
thread_local int g_part_blocks_per_sm = 0;
namespace {

constexpr int kMaxRanks = BLIGHT_MAX_RANKS;
constexpr int kRecWords = 5;                  // 4 words of bases + one zero word for the funnel
constexpr uint32_t kMaxRecKmers = 25;         // k-mers per record (k - m + 1 of the usual shapes): bounds a warp's staging
constexpr int kMaxIds = 32 * kMaxRecKmers;
constexpr int kResCap = 128;                  // entries of a warp's work list (drained whenever fewer than 32 slots are left)
constexpr unsigned long long kKmerBits = 40;  // packed pair counter: slots << 40 | k-mers
constexpr unsigned long long kKmerMask = (1ull << kKmerBits) - 1;
constexpr uint32_t kIdAbsent = 0xFFFFFFFFu;

struct alignas(32) RunRec {
	uint32_t bases[4];  // n + k - 1 bases, 2 bits each, first base in the high bits of bases[0]
	uint32_t ko;        // stream return: k-mer offset of the run's first k-mer in the (source, owner) stream of this sub-batch
	                    // direct return: low half of the slot of the run's first k-mer in the source's id array
	uint32_t mn;        // minimizer (bucket) of every k-mer of the run
	uint32_t n_src;     // bits 0-7: k-mers in the run, bits 8-15: source rank
	uint32_t o_hi;      // direct return: high half of the slot
};
static_assert(sizeof(RunRec) == 32, "one record = one sector");

struct Route {
	uint32_t world, rank, lb, pad;
	uint32_t cuts[kMaxRanks + 1];
	RunRec* inbox[kMaxRanks];  // this source's region in every owner's inbox (peer pointers)
	uint64_t cap;              // records per region
	uint64_t kcap;             // k-mers per return region
	uint4* side;               // local: {o_lo, o_hi, ko, n} of record `slot` bound for owner d at [d * cap + slot]; null = counting / direct
	uint32_t direct;           // ids return straight into this source's id array (the record carries the slot), no side table
};

__device__ __forceinline__ uint32_t owner_of(const Route& R, uint32_t mini) {
	const uint32_t g = mini >> R.lb;
	uint32_t o = 0;
	#pragma unroll 1
	while (o + 1 < R.world && g >= R.cuts[o + 1]) o++;
	return o;
}

template <bool WANT_O>
__global__ void __launch_bounds__(kThreads, 4) k_dispatch_runs(uint32_t k, uint32_t m, const char* __restrict__ bases,
                                                             const uint64_t* __restrict__ read_off, const uint64_t* __restrict__ read_end,
                                                             const uint64_t* __restrict__ kmer_off, uint64_t n_reads, uint64_t total_bases,
                                                             uint64_t strip_lo, uint64_t strip_hi, bool aligned16, const uint32_t* __restrict__ packed,
                                                             double reads_per_base, uint64_t guess_p0, Route R,
                                                             unsigned long long* __restrict__ counts, uint64_t* __restrict__ ctr,
                                                             uint32_t* __restrict__ err, unsigned long long* ticket) {
	__shared__ uint32_t s_pack[kWarps][kStripWords];
	__shared__ uint32_t s_bad[kWarps][kStripWords];
	__shared__ uint32_t s_keys[kWarps][kKeySlots];
	__shared__ uint64_t s_run_o[WANT_O ? kWarps : 1][kMaxRuns];
	__shared__ uint16_t s_run_q[kWarps][kMaxRuns];
	__shared__ uint32_t s_run_key[kWarps][kMaxRuns];
	__shared__ uint32_t s_run_n[kWarps][kMaxRuns];  // k-mers of the run inside this strip
	__shared__ uint64_t s_runid8[kWarps][kStrip / 8];

	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t ow = WANT_O ? wid : 0;
	const StripSmem S{s_pack[wid], s_bad[wid], s_keys[wid], s_run_q[wid], s_run_key[wid], s_run_o[ow], s_runid8[wid]};
	const uint8_t* runid = reinterpret_cast<const uint8_t*>(s_runid8[wid]);
	const uint32_t w = k - m + 1;
	const uint32_t nmax = min(64u - k + 1u, kMaxRecKmers);  // k-mers one record carries (64 bases)
	const uint64_t warp_stride = (uint64_t)gridDim.x * kWarps;
	const uint32_t lt_mask = (1u << lane) - 1u;
	uint32_t invalid = 0, queries = 0;

	for (uint64_t strip = strip_lo + (uint64_t)blockIdx.x * kWarps + wid; strip < strip_hi; strip = next_item(ticket, strip, warp_stride, strip_lo, lane)) {
		const uint64_t t0 = strip * kStrip;
		__syncwarp();
		s_run_n[wid][lane] = 0;
		s_run_n[wid][lane + 32] = 0;
		const uint32_t n_runs = strip_front<WANT_O>(S, lane, k, m, bases, read_off, read_end, kmer_off, n_reads, total_bases, reads_per_base,
		                                            guess_p0, aligned16, packed, t0, invalid);
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
						const uint64_t r = find_read(read_off, n_reads, reads_per_base, t0 + q, guess_p0);
						o = __ldg(kmer_off + r) + (t0 + q - __ldg(read_off + r));
					}
				}
			}
			uint32_t mn = 0, dst = 0;
			if (n) {
				mn = mini_from_key(i < n_tab ? s_run_key[wid][i] : window_min_slow(S.keys, q, w));
				dst = owner_of(R, mn);
			}
			#pragma unroll 1
			for (uint32_t off = 0; __any_sync(0xffffffffu, off < n); off += nmax) {
				const bool act = off < n;
				const uint32_t nn = act ? min(n - off, nmax) : 0u;
				// lanes bound for the same owner: exclusive prefix and total of their k-mer counts, one warp scan per owner present
				uint32_t peers = 0, kpre = 0, ktot = 0;
				uint32_t todo = __ballot_sync(0xffffffffu, act);
				#pragma unroll 1
				while (todo) {
					const uint32_t d0 = __shfl_sync(0xffffffffu, dst, __ffs(todo) - 1);
					const bool mine = act && dst == d0;
					uint32_t inc = mine ? nn : 0u;
					#pragma unroll
					for (int sh = 1; sh < 32; sh <<= 1) {
						const uint32_t t = __shfl_up_sync(0xffffffffu, inc, sh);
						if (lane >= (uint32_t)sh) inc += t;
					}
					const uint32_t tot = __shfl_sync(0xffffffffu, inc, 31);
					const uint32_t grp = __ballot_sync(0xffffffffu, mine);
					if (mine) { peers = grp; kpre = inc - nn; ktot = tot; }
					todo &= ~grp;
				}
				if (act) {
					const uint32_t leader = __ffs(peers) - 1;
					unsigned long long r0 = 0;
					if (lane == leader) r0 = atomicAdd(&counts[dst], ((unsigned long long)__popc(peers) << kKmerBits) | ktot);
					r0 = __shfl_sync(peers, r0, leader);
					const uint64_t slot = (r0 >> kKmerBits) + __popc(peers & lt_mask);
					const uint64_t ko = (r0 & kKmerMask) + kpre;
					if (slot < R.cap && ko + nn <= R.kcap) {
						const uint4 b = strip_bases64(S.pack, q + off);
						uint4* dstp = reinterpret_cast<uint4*>(R.inbox[dst] + slot);
						dstp[0] = b;
						const uint64_t oo = o + off;
						if (WANT_O && R.direct) {
							dstp[1] = make_uint4((uint32_t)oo, mn, nn | (R.rank << 8), (uint32_t)(oo >> 32));
						} else {
							dstp[1] = make_uint4((uint32_t)ko, mn, nn | (R.rank << 8), 0u);
							if (WANT_O) R.side[(uint64_t)dst * R.cap + slot] = make_uint4((uint32_t)oo, (uint32_t)(oo >> 32), (uint32_t)ko, nn);
						}
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
	uint32_t world, rank;             // rank: this owner's position among the sources (only sets where its round-robin starts)
	uint64_t cap, kcap;               // records per inbox region, ids per return region
	const RunRec* region[kMaxRanks];  // records received from every source
	uint32_t* ret[kMaxRanks];         // stream return: this owner's return region at every source (peer pointers)
	int64_t* out[kMaxRanks];          // direct return: every source's id array (peer pointers) ...
	uint64_t out_cap[kMaxRanks];      // ... and its length
	uint32_t direct;
};

// k-mer at offset d of a record's bases
__device__ __forceinline__ uint64_t rec_kmer(const uint32_t* W, uint32_t d, uint32_t k) {
	const uint32_t wi = d >> 4, s = 2u * (d & 15);
	const uint32_t a = W[wi], b = W[wi + 1], c = W[wi + 2];
	return (((uint64_t)__funnelshift_l(b, a, s) << 32) | __funnelshift_l(c, b, s)) >> (64 - 2 * k);
}

// run holding flattened index i, given the inclusive prefix of the runs' sizes (32 entries)
__device__ __forceinline__ uint32_t run_of(const uint16_t* incl, uint32_t i) {
	uint32_t lo = 0, hi = 31;
	#pragma unroll
	for (int s = 0; s < 5; s++) {
		const uint32_t mid = (lo + hi) >> 1;
		if (incl[mid] > i) hi = mid; else lo = mid + 1;
	}
	return lo;
}

template <bool WANT_IDS, bool SMALL>
__global__ void __launch_bounds__(kThreads, 4) k_runs_lookup(DevIndexView I, OwnerArgs A, const unsigned long long* __restrict__ counts,
                                                           uint64_t* __restrict__ ctr, unsigned long long* ticket) {
	__shared__ unsigned long long s_cnt[kMaxRanks];  // records received from source s
	__shared__ unsigned long long s_rounds;          // chunks (32 records of one source) of the fullest region
	__shared__ uint32_t s_w[kWarps][32][kRecWords + 1];  // +1: odd stride
	__shared__ uint64_t s_T[kWarps][32];
	__shared__ uint32_t s_mn[kWarps][32];
	__shared__ uint16_t s_dmax[kWarps][32];
	__shared__ uint16_t s_incl[kWarps][32];  // inclusive prefix of n: the ids of the warp's runs, flattened, live at [incl[r-1], incl[r])
	__shared__ uint8_t s_flag[kWarps][32];
	__shared__ uint64_t s_okv[kWarps][32];
	__shared__ uint16_t s_res[kWarps][kResCap];                 // work list: flattened indices waiting for the whole lookup
	__shared__ uint32_t s_ids[WANT_IDS ? kWarps : 1][kMaxIds];  // the warp's answers, in the order they travel back
	__shared__ uint64_t s_o[WANT_IDS ? kWarps : 1][32];         // direct return: slot of every run's first k-mer at the source

	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint32_t iw = WANT_IDS ? wid : 0;
	const uint32_t k = I.k;
	const uint32_t lt_mask = (1u << lane) - 1u;
	const bool filter_anchors = (I.flags & kFlagFilterAnchors) != 0;
	if (threadIdx.x == 0) {
		unsigned long long most = 0;
		for (uint32_t s = 0; s < A.world; s++) {
			// a source that ran out of room dropped the records past `cap` (and raised its error flag: the batch is answered
			// again through another path); never read past the region
			const unsigned long long c = min(counts[s] >> kKmerBits, (unsigned long long)A.cap);
			s_cnt[s] = c;
			most = max(most, (c + 31) / 32);
		}
		s_rounds = most;
	}
	__syncthreads();
	// Chunk order: round-robin over the sources, starting from this owner's right-hand neighbour. Owners run in step (they leave
	// the same barrier together); taking the sources one after the other made all of them return ids to the SAME GPU at any
	// moment — 8 senders into one NVLink port (measured: lookups 1.23 ms on the rank that happened to run ahead, 1.61 ms on
	// the last one). Interleaved, every owner's stores are spread evenly over all sources all the time.
	const uint64_t n_chunks = s_rounds * A.world;
	const uint64_t warp_stride = (uint64_t)gridDim.x * kWarps;
	const uint32_t mn_limit = 1u << (2 * I.m - 1);
	uint32_t found = 0, notfound = 0;

	for (uint64_t chunk = (uint64_t)blockIdx.x * kWarps + wid; chunk < n_chunks; chunk = next_item(ticket, chunk, warp_stride, 0, lane)) {
		const uint64_t round = chunk / A.world;
		const uint32_t src = (uint32_t)((chunk - round * A.world + A.rank + 1) % A.world);
		const uint64_t rec0 = round * 32;
		if (rec0 >= s_cnt[src]) continue;
		const uint32_t n_first = (uint32_t)min((unsigned long long)32, s_cnt[src] - rec0);
		__syncwarp();
		// the warp's 32 records into shared memory
		uint32_t n = 0, ko = 0;
		if (lane < n_first) {
			const uint4* rp = reinterpret_cast<const uint4*>(A.region[src] + rec0 + lane);
			const uint4 b = __ldcs(rp), h = __ldcs(rp + 1);
			uint32_t* W = s_w[wid][lane];
			W[0] = b.x; W[1] = b.y; W[2] = b.z; W[3] = b.w; W[4] = 0;
			ko = h.x;
			if (WANT_IDS) s_o[iw][lane] = ((uint64_t)h.w << 32) | h.x;
			s_mn[wid][lane] = h.y < mn_limit ? h.y : 0u;  // (a slot the source dropped holds an older record or zeros: stay in bounds)
			n = min(h.z & 0xFFu, kMaxRecKmers);
		}
		uint32_t incl = n;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= (uint32_t)o) incl += t;
		}
		s_incl[wid][lane] = (uint16_t)incl;
		const uint32_t n_ids = __shfl_sync(0xffffffffu, incl, 31);
		const uint32_t ko0 = __shfl_sync(0xffffffffu, ko, 0);  // k-mer offsets are consecutive in slot order
		__syncwarp();

		// Work list of flattened k-mer indices waiting for the whole lookup: first the first k-mer of every run (d == 0), then,
		// turn by turn, what the one-window prediction left. ONE copy of filter + lookup drains it (instruction cache).
		uint32_t n_res = n_first, base = 0;
		if (lane < n_first) s_res[wid][lane] = (uint16_t)(incl - n);
		bool anchors = true;
		__syncwarp();
		#pragma unroll 1
		do {
			// drain: negative filter (anchors only if asked to), then the lookup, compacted in place
			if (I.filter && (!anchors || filter_anchors)) {
				uint32_t n_keep = 0;
				#pragma unroll 1
				for (uint32_t b0 = 0; b0 < n_res; b0 += 32) {
					const uint32_t j = b0 + lane;
					bool keep = false;
					uint32_t i = 0;
					if (j < n_res) {
						i = s_res[wid][j];
						const uint32_t run = run_of(s_incl[wid], i);
						const uint32_t d = i - (run ? s_incl[wid][run - 1] : 0u);
						const uint64_t f = rec_kmer(s_w[wid][run], d, k), rc = rc64(f, k);
						keep = filter_maybe(I, f < rc ? f : rc);
						if (!keep) {
							notfound++;
							if (WANT_IDS) s_ids[iw][i] = kIdAbsent;
							if (d == 0) s_flag[wid][run] = 0;
						}
					}
					const uint32_t km = __ballot_sync(0xffffffffu, keep);
					__syncwarp();
					if (keep) s_res[wid][n_keep + __popc(km & lt_mask)] = (uint16_t)i;
					n_keep += __popc(km);
				}
				n_res = n_keep;
				__syncwarp();
			}
			#pragma unroll 1
			for (uint32_t b0 = 0; b0 < n_res; b0 += 32) {
				const uint32_t j = b0 + lane;
				if (j < n_res) {
					const uint32_t i = s_res[wid][j];
					const uint32_t run = run_of(s_incl[wid], i);
					const uint32_t d = i - (run ? s_incl[wid][run - 1] : 0u);
					const uint32_t mn = s_mn[wid][run];
					const uint64_t f = rec_kmer(s_w[wid][run], d, k), rc = rc64(f, k);
					uint64_t T = 0;
					const int64_t idr = lookup_one<SMALL>(I, f < rc ? f : rc, mn, &T);
					if (idr >= 0) found++; else notfound++;
					if (WANT_IDS) s_ids[iw][i] = idr >= 0 ? (uint32_t)((uint64_t)idr - I.id_base) : kIdAbsent;  // slice-local: fits 32 bits (the table exists)
					if (d == 0) {
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
			n_res = 0;
			__syncwarp();
			if (anchors) {
				// one lane per run: the run's text against the index text next to where its first k-mer matched, all of its windows
				// at once (front.cuh: mismatch_windows64). Low half of okv: k-mer d equals the window at T +- d; high half: that
				// window's valid bit.
				uint64_t okv = 0;
				if (lane < n_first) {
					const uint32_t flag = s_flag[wid][lane];
					const uint32_t cnt = min(31u, (uint32_t)s_dmax[wid][lane]);
					if ((flag & 1) && cnt) {
						const bool same = flag & 2;
						const uint64_t Ta = s_T[wid][lane];
						const uint64_t S0 = same ? Ta : Ta - cnt;
						const uint32_t* W = s_w[wid][lane];
						uint4 R = make_uint4(W[0], W[1], W[2], W[3]);
						if (!same) R = shl_bases64(rc_bases64(R), 64 - (cnt + k));
						const uint64_t Am = mismatch_windows64(R, seq_bases64(I.seq, S0), k);
						uint32_t ok = same ? __brev((uint32_t)(~Am >> 32)) : (uint32_t)(~Am >> (63 - cnt));
						ok &= cnt == 31 ? 0xFFFFFFFEu : ((2u << cnt) - 2u);
						uint32_t v = 0;
						if (!WANT_IDS) {
							const uint32_t* vp = I.valid + (S0 >> 5);
							const uint32_t u = __funnelshift_r(__ldg(vp), __ldg(vp + 1), (uint32_t)(S0 & 31));
							v = same ? u : (__brev(u) >> (31 - cnt));
						}
						okv = ok | ((uint64_t)v << 32);
					}
				}
				s_okv[wid][lane] = okv;
				anchors = false;
				__syncwarp();
			}
			// fill: every other k-mer of a run is answered by its run's masks, or joins the work list
			#pragma unroll 1
			while (base < n_ids && n_res + 32 <= (uint32_t)kResCap) {
				const uint32_t i = base + lane;
				bool left = false;
				if (i < n_ids) {
					const uint32_t run = run_of(s_incl[wid], i);
					const uint32_t d = i - (run ? s_incl[wid][run - 1] : 0u);
					if (d) {
						const uint64_t okv = s_okv[wid][run];
						if (d < 32 && ((okv >> d) & 1)) {
							bool v;
							if (WANT_IDS) {
								const uint64_t Ta = s_T[wid][run];
								const uint32_t pid = __ldg(I.pos_id + ((s_flag[wid][run] & 2) ? Ta + d : Ta - d));
								v = pid != kIdAbsent;
								s_ids[iw][i] = pid;
							} else {
								v = (okv >> (32 + d)) & 1;
							}
							if (v) found++; else notfound++;
						} else {
							left = true;
						}
					}
				}
				const uint32_t lm = __ballot_sync(0xffffffffu, left);
				if (left) s_res[wid][n_res + __popc(lm & lt_mask)] = (uint16_t)i;
				n_res += __popc(lm);
				base += 32;
			}
			__syncwarp();
		} while (n_res > 0 || base < n_ids);
		if (WANT_IDS && A.direct) {
			// the warp's answers straight into the source's id array (NVLink): the ids of a run are consecutive int64 slots, so
			// lanes holding one run store one contiguous segment
			int64_t* dst = A.out[src];
			const uint64_t lim = A.out_cap[src];
			const long long id0 = (long long)I.id_base;
			for (uint32_t t = lane; t < n_ids; t += 32) {
				const uint32_t run = run_of(s_incl[wid], t);
				const uint64_t o = s_o[iw][run] + (t - (run ? s_incl[wid][run - 1] : 0u));
				const uint32_t v = s_ids[iw][t];
				if (o < lim) __stcs(reinterpret_cast<long long*>(dst + o), v == kIdAbsent ? -1ll : id0 + (long long)v);
			}
		} else if (WANT_IDS) {
			// the warp's answers back to the source: one contiguous stream of 32-bit ids over NVLink, 16 bytes per lane and
			// store (512 bytes per warp instruction) between a scalar head and tail that bring the stream to 16-byte alignment —
			// remote stores are bound by the number of requests in flight, not by their bytes
			uint32_t* dst = A.ret[src] + ko0;
			if ((uint64_t)ko0 + n_ids <= A.kcap) {
				const uint32_t* sid = s_ids[iw];
				const uint32_t head = min(n_ids, (4u - (ko0 & 3u)) & 3u);
				if (lane < head) dst[lane] = sid[lane];
				const uint32_t n4 = (n_ids - head) >> 2;
				for (uint32_t g = lane; g < n4; g += 32) {
					const uint32_t t = head + 4 * g;
					*reinterpret_cast<uint4*>(dst + t) = make_uint4(sid[t], sid[t + 1], sid[t + 2], sid[t + 3]);
				}
				const uint32_t done = head + 4 * n4;
				if (done + lane < n_ids) dst[done + lane] = sid[done + lane];
			}
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

// Source side, after the owners answered: return streams -> int64 ids in read order. A warp takes 32 consecutive
// records of one owner from the side table; their ids are consecutive in that owner's return region.
struct ScatterSrc {
	const uint32_t* ret[kMaxRanks];  // owner d's ids for this source: a local region the owner pushed into, or (pull) the owner's own memory
	uint64_t id_base[kMaxRanks];     // first identifier of every owner's slice: owners return slice-local 32-bit ids
	uint32_t rank;                   // this source's position among the owners (where its round-robin starts)
};

__device__ __forceinline__ uint4 ld_stream16(const uint32_t* p) {
	uint4 v;
	asm volatile("ld.global.cs.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
	return v;
}

// The warp's id stream is brought into shared memory with 16-byte loads that are ALL issued before the first one is used
// (pull return: they cross NVLink, a warp keeps up to 3.2 KB in flight), then leaves as coalesced int64 stores.
__global__ void __launch_bounds__(kThreads) k_scatter_runs(const uint4* __restrict__ side, uint64_t cap, const unsigned long long* __restrict__ counts,
                                                           ScatterSrc A, uint64_t kcap, uint32_t world, int64_t* __restrict__ out) {
	__shared__ unsigned long long s_cnt[kMaxRanks];
	__shared__ unsigned long long s_rounds;
	__shared__ uint64_t s_o[kWarps][32];
	__shared__ uint16_t s_incl[kWarps][32];
	__shared__ uint32_t s_v[kWarps][kMaxIds];
	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (threadIdx.x == 0) {
		unsigned long long most = 0;
		for (uint32_t d = 0; d < world; d++) {
			const unsigned long long c = min(counts[d] >> kKmerBits, (unsigned long long)cap);
			s_cnt[d] = c;
			most = max(most, (c + 31) / 32);
		}
		s_rounds = most;
	}
	__syncthreads();
	// round-robin over the owners, as in k_runs_lookup (pull return: no two sources fetch from the same owner in step)
	const uint64_t n_chunks = s_rounds * world;
	const uint64_t warp_stride = (uint64_t)gridDim.x * kWarps;
	constexpr int kVec = (kMaxIds / 4 + 31) / 32;  // 16-byte loads per lane that cover a warp's longest stream
	for (uint64_t chunk = (uint64_t)blockIdx.x * kWarps + wid; chunk < n_chunks; chunk += warp_stride) {
		const uint64_t round = chunk / world;
		const uint32_t d = (uint32_t)((chunk - round * world + A.rank + 1) % world);
		const uint64_t rec0 = round * 32;
		if (rec0 >= s_cnt[d]) continue;
		const uint32_t n_first = (uint32_t)min((unsigned long long)32, s_cnt[d] - rec0);
		__syncwarp();
		uint32_t n = 0, ko = 0;
		if (lane < n_first) {
			const uint4 e = __ldcs(side + (uint64_t)d * cap + rec0 + lane);
			s_o[wid][lane] = ((uint64_t)e.y << 32) | e.x;
			ko = e.z;
			n = min(e.w, kMaxRecKmers);
		}
		uint32_t incl = n;
		#pragma unroll
		for (int o = 1; o < 32; o <<= 1) {
			const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
			if (lane >= (uint32_t)o) incl += t;
		}
		s_incl[wid][lane] = (uint16_t)incl;
		const uint32_t n_ids = __shfl_sync(0xffffffffu, incl, 31);
		const uint32_t ko0 = __shfl_sync(0xffffffffu, ko, 0);
		if ((uint64_t)ko0 + n_ids > kcap) continue;  // the owner dropped what did not fit its return region (overflow flag raised there)
		const uint32_t* srcp = A.ret[d] + ko0;
		uint32_t* sv = s_v[wid];
		const uint32_t head = min(n_ids, (4u - (ko0 & 3u)) & 3u);
		const uint32_t n4 = (n_ids - head) >> 2;
		const uint32_t done = head + 4 * n4;
		uint4 v[kVec];
		#pragma unroll
		for (int j = 0; j < kVec; j++) {
			const uint32_t g = lane + 32u * j;
			if (g < n4) v[j] = ld_stream16(srcp + head + 4 * g);
		}
		uint32_t vh = 0, vt = 0;
		if (lane < head) vh = __ldcs(srcp + lane);
		if (done + lane < n_ids) vt = __ldcs(srcp + done + lane);
		#pragma unroll
		for (int j = 0; j < kVec; j++) {
			const uint32_t g = lane + 32u * j;
			if (g < n4) {
				const uint32_t t = head + 4 * g;
				sv[t] = v[j].x; sv[t + 1] = v[j].y; sv[t + 2] = v[j].z; sv[t + 3] = v[j].w;
			}
		}
		if (lane < head) sv[lane] = vh;
		if (done + lane < n_ids) sv[done + lane] = vt;
		__syncwarp();
		const long long id0 = (long long)A.id_base[d];
		for (uint32_t i = lane; i < n_ids; i += 32) {
			const uint32_t run = run_of(s_incl[wid], i);
			const uint32_t dd = i - (run ? s_incl[wid][run - 1] : 0u);
			const uint32_t x = sv[i];
			__stcs(reinterpret_cast<long long*>(out + s_o[wid][run] + dd), x == kIdAbsent ? -1ll : id0 + (long long)x);
		}
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

// resident CTAs per SM a launch may take: what fits, unless the caller shares the SMs between two kernels (part_session.cu)
int limit_per_sm(int fits) { return g_part_blocks_per_sm > 0 && g_part_blocks_per_sm < fits ? g_part_blocks_per_sm : fits; }

struct DeviceGuardLite {
	int prev = -1;
	explicit DeviceGuardLite(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
	~DeviceGuardLite() { if (prev >= 0) cudaSetDevice(prev); }
};

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
	ReadBatch B;
	B.d_bases = d_bases; B.d_read_off = d_read_off; B.d_kmer_off = d_kmer_off; B.n_reads = n_reads; B.total_bases = total_bases;
	return part_dispatch_batch(k, m, B, pos_begin, pos_end, route, d_counts, d_ctr, d_err, stream);
}

}  // extern "C"

namespace blight {
// Loads every kernel of the partitioned path now. With lazy module loading the first launch of a kernel may have to wait
// for the device to drain — which never happens while another rank's wait kernel spins on a flag this rank has yet to
// publish. Sessions call this when they are created (nothing is in flight then).
int part_kernels_preload() {
	cudaFuncAttributes a;
	cudaError_t e = cudaFuncGetAttributes(&a, k_dispatch_runs<true>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_dispatch_runs<false>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_runs_lookup<true, true>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_runs_lookup<true, false>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_runs_lookup<false, true>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_runs_lookup<false, false>);
	if (e == cudaSuccess) e = cudaFuncGetAttributes(&a, k_scatter_runs);
	(void)per_sm(k_dispatch_runs<true>); (void)per_sm(k_dispatch_runs<false>);
	return e == cudaSuccess ? BL_OK : fail(BL_ERR_CUDA, std::string("cannot load the partition kernels: ") + cudaGetErrorString(e));
}

int part_dispatch_batch(uint32_t k, uint32_t m, const ReadBatch& B, uint64_t pos_begin, uint64_t pos_end, const blight_part_route* route,
                        uint64_t* d_counts, uint64_t* d_ctr, uint32_t* d_err, void* stream, uint64_t* d_ticket) {
	unsigned long long* tk = reinterpret_cast<unsigned long long*>(d_ticket);
	const uint64_t n_reads = B.n_reads, total_bases = B.total_bases;
	if (!route || !d_counts || !d_ctr || !d_err || (n_reads && ((!B.d_bases && !B.d_packed) || !B.d_read_off))) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (route->world == 0 || route->world > (uint32_t)kMaxRanks || route->rank >= route->world) return fail(BL_ERR_INVALID_ARG, "bad world / rank");
	if (k < 8 || k > 32 || m >= k || k - m + 1 < 8 || k - m + 1 > 32) return fail(BL_ERR_INVALID_ARG, "partition mode needs 8 <= k <= 32 and 8 <= k-m+1 <= 32");
	if (route->cap == 0 || route->cap >= (1ull << 24) || route->kcap >= (1ull << 32)) return fail(BL_ERR_INVALID_ARG, "region capacities: cap < 2^24 records, kcap < 2^32 k-mers");
	if (n_reads == 0 || total_bases == 0) return BL_OK;
	if (pos_end > total_bases) pos_end = total_bases;
	if (pos_begin >= pos_end) return BL_OK;
	if ((pos_begin % kStrip) != 0 || (pos_end < total_bases && (pos_end % kStrip) != 0)) return fail(BL_ERR_INVALID_ARG, "sub-batch bounds must be multiples of 256");
	Route R{};
	R.world = route->world; R.rank = route->rank; R.lb = route->lb; R.cap = route->cap; R.kcap = route->kcap;
	R.side = static_cast<uint4*>(route->side);
	R.direct = (B.d_kmer_off && !route->side) ? 1u : 0u;  // id mode without a side table: the records carry the output slots
	for (uint32_t i = 0; i <= route->world; i++) R.cuts[i] = route->cuts[i];
	for (uint32_t i = 0; i < route->world; i++) {
		if (!route->inbox[i]) return fail(BL_ERR_INVALID_ARG, "null inbox pointer");
		R.inbox[i] = static_cast<RunRec*>(route->inbox[i]);
	}
	const uint64_t strip_lo = pos_begin / kStrip, strip_hi = (pos_end + kStrip - 1) / kStrip;
	const bool al = (reinterpret_cast<uintptr_t>(B.d_bases) & 15) == 0;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	const uint64_t want = (strip_hi - strip_lo + kWarps - 1) / kWarps;
	const double rpb = B.rpb > 0 ? B.rpb : (double)n_reads / (double)total_bases;
	if (B.d_kmer_off) {
		static const int nb = per_sm(k_dispatch_runs<true>);
		const uint64_t cap = (uint64_t)sm_count_() * limit_per_sm(nb);
		k_dispatch_runs<true><<<(unsigned)(want < cap ? want : cap), kThreads, 0, st>>>(k, m, B.d_bases, B.d_read_off, B.d_read_end, B.d_kmer_off, n_reads,
			total_bases, strip_lo, strip_hi, al, B.d_packed, rpb, B.guess_p0, R, reinterpret_cast<unsigned long long*>(d_counts), d_ctr, d_err, tk);
	} else {
		static const int nb = per_sm(k_dispatch_runs<false>);
		const uint64_t cap = (uint64_t)sm_count_() * limit_per_sm(nb);
		k_dispatch_runs<false><<<(unsigned)(want < cap ? want : cap), kThreads, 0, st>>>(k, m, B.d_bases, B.d_read_off, B.d_read_end, nullptr, n_reads,
			total_bases, strip_lo, strip_hi, al, B.d_packed, rpb, B.guess_p0, R, reinterpret_cast<unsigned long long*>(d_counts), d_ctr, d_err, tk);
	}
	g_launches++;
	return finish("k_dispatch_runs");
}

// ret[d] = where owner d's 32-bit ids for this source are read from: world device pointers (local regions, or peer
// pointers into the owners' own memory for the pull return path)
int part_scatter_from(const void* d_side, uint64_t cap, const uint64_t* d_counts, const void* const* ret, uint64_t kcap, uint32_t world,
                      uint32_t rank, uint64_t max_records, const uint64_t* id_bases, int64_t* d_ids, void* stream) {
	if (!d_side || !d_counts || !ret || !d_ids) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (world == 0 || world > (uint32_t)kMaxRanks) return fail(BL_ERR_INVALID_ARG, "bad world");
	const uint64_t want = std::max<uint64_t>(1, ((max_records + 31) / 32 + world + kWarps - 1) / kWarps);
	static const int nb = per_sm(k_scatter_runs);
	const uint64_t capb = (uint64_t)sm_count_() * nb;
	ScatterSrc A{};
	A.rank = rank < world ? rank : 0;
	for (uint32_t i = 0; i < world; i++) {
		if (!ret[i]) return fail(BL_ERR_INVALID_ARG, "null return region");
		A.ret[i] = static_cast<const uint32_t*>(ret[i]);
		A.id_base[i] = id_bases ? id_bases[i] : 0;
	}
	k_scatter_runs<<<(unsigned)(want < capb ? want : capb), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
		static_cast<const uint4*>(d_side), cap, reinterpret_cast<const unsigned long long*>(d_counts), A, kcap, world, d_ids);
	g_launches++;
	return finish("k_scatter_runs");
}
}  // namespace blight

extern "C" {

int blight_part_lookup(const blight_index* idx, uint32_t world, const void* const* regions, const uint64_t* d_counts, void* const* ret,
                       uint64_t cap, uint64_t kcap, uint64_t* d_ctr, void* stream) {
	return blight_part_lookup_direct(idx, world, regions, d_counts, ret, nullptr, nullptr, cap, kcap, d_ctr, stream);
}

int blight_part_lookup_direct(const blight_index* idx, uint32_t world, const void* const* regions, const uint64_t* d_counts, void* const* ret,
                              void* const* out_ids, const uint64_t* out_caps, uint64_t cap, uint64_t kcap, uint64_t* d_ctr, void* stream) {
	return part_lookup_from(idx, world, 0, regions, d_counts, ret, out_ids, out_caps, cap, kcap, d_ctr, stream);
}

}  // extern "C"

int blight::part_lookup_from(const blight_index* idx, uint32_t world, uint32_t rank, const void* const* regions, const uint64_t* d_counts,
                             void* const* ret, void* const* out_ids, const uint64_t* out_caps, uint64_t cap, uint64_t kcap, uint64_t* d_ctr,
                             void* stream, uint64_t* d_ticket) {
	unsigned long long* tk = reinterpret_cast<unsigned long long*>(d_ticket);
	if (out_ids && (ret || !out_caps)) return fail(BL_ERR_INVALID_ARG, "direct return: pass out_ids + out_caps and no return regions");
	const uint64_t max_records = (uint64_t)world * cap;
	if (!idx || !regions || !d_counts || !d_ctr) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (world == 0 || world > (uint32_t)kMaxRanks) return fail(BL_ERR_INVALID_ARG, "bad world");
	if (!idx->v.valid) return fail(BL_ERR_INVALID_ARG, "index has no valid-window bitmap");
	if ((ret || out_ids) && !idx->v.pos_id) return fail(BL_ERR_INVALID_ARG, "id mode of the partitioned path needs the position->id table (BLIGHT_POS_ID)");
	if (idx->v.k < 8) return fail(BL_ERR_INVALID_ARG, "partition mode needs k >= 8");
	DeviceGuardLite guard(idx->device);
	OwnerArgs A{};
	A.world = world;
	A.rank = rank < world ? rank : 0;
	A.cap = cap;
	A.kcap = kcap;
	for (uint32_t i = 0; i < world; i++) {
		A.region[i] = static_cast<const RunRec*>(regions[i]);
		A.ret[i] = ret ? static_cast<uint32_t*>(ret[i]) : nullptr;
		A.out[i] = out_ids ? static_cast<int64_t*>(out_ids[i]) : nullptr;
		A.out_cap[i] = out_ids ? out_caps[i] : 0;
	}
	A.direct = out_ids ? 1u : 0u;
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	const uint64_t want = std::max<uint64_t>(1, ((max_records + 31) / 32 + world + kWarps - 1) / kWarps);
	const unsigned long long* cnt = reinterpret_cast<const unsigned long long*>(d_counts);
#define BL_LAUNCH(IDS, SM)                                                                              \
	do {                                                                                                 \
		static const int nb = per_sm(k_runs_lookup<IDS, SM>);                                            \
		const uint64_t cap = (uint64_t)sm_count_() * limit_per_sm(nb);                                   \
		k_runs_lookup<IDS, SM><<<(unsigned)(want < cap ? want : cap), kThreads, 0, st>>>(idx->v, A, cnt, d_ctr, tk); \
	} while (0)
	if (ret || out_ids) { if (idx->v.small) BL_LAUNCH(true, true); else BL_LAUNCH(true, false); }
	else { if (idx->v.small) BL_LAUNCH(false, true); else BL_LAUNCH(false, false); }
#undef BL_LAUNCH
	g_launches++;
	return finish("k_runs_lookup");
}

extern "C" {

int blight_part_scatter(const void* d_side, uint64_t cap, const uint64_t* d_counts, const void* d_ret, uint64_t kcap, uint32_t world,
                        uint64_t max_records, const uint64_t* id_bases, int64_t* d_ids, void* stream) {
	if (!d_ret) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (world == 0 || world > (uint32_t)kMaxRanks) return fail(BL_ERR_INVALID_ARG, "bad world");
	const void* regions[kMaxRanks];
	for (uint32_t i = 0; i < world; i++) regions[i] = static_cast<const uint32_t*>(d_ret) + (uint64_t)i * kcap;
	return part_scatter_from(d_side, cap, d_counts, regions, kcap, world, 0, max_records, id_bases, d_ids, stream);
}

int blight_peer_alloc(uint64_t bytes, void** d_ptr, unsigned char* handle64) {
	if (!d_ptr || !handle64 || bytes == 0) return fail(BL_ERR_INVALID_ARG, "bad argument");
	void* p = nullptr;
	cudaError_t e = cudaMalloc(&p, bytes);
	if (e != cudaSuccess) return fail(BL_ERR_NOMEM, std::string("cudaMalloc(peer buffer): ") + cudaGetErrorString(e));
	e = cudaMemset(p, 0, bytes);  // an inbox slot that was never written must still parse as a (harmless) record
	if (e != cudaSuccess) { cudaFree(p); return fail(BL_ERR_CUDA, std::string("cudaMemset(peer buffer): ") + cudaGetErrorString(e)); }
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
