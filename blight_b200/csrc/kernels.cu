// kernels.cu — sm_100a kernels of the batched query path and their launchers.
//
//   k_lookup_kmers   canonical k-mers (+ optional minimizers) -> ids          query_kmer_hash, blight.cpp:545-550
//   k_reads          ASCII reads -> 2-bit pack -> canonical k-mers + rolling  query_sequence_hash/bool,
//                    minimizers -> {pairs | ids | counts}                     blight.cpp:554-591; kmer.h:791-810
//
// k_reads works on tiles of the concatenated base stream.  One CTA packs TILE+halo bases into shared memory
// (coalesced 16-byte loads), computes the ordering key of every m-mer once, takes the window minimum over the
// k-m+1 keys of each k-mer (equal to minimizer_naive on the canonical k-mer, because the set of canonical m-mers
// of a k-mer is strand invariant and revhash is a bijection), and either stores (canon, minimizer) or goes
// straight into the lookup core with adjacent lanes holding adjacent k-mers, so that the k-mers of one super-k-mer
// share their bucket descriptor, MPHF descriptor and sequence sectors in L1.
#include <cuda_runtime.h>

#include <atomic>

#include "kernels.hpp"
#include "lookup.cuh"

namespace blight {

std::atomic<uint64_t> g_launches{0};

namespace {

constexpr int kThreads = 256;
constexpr int kTile = 2048;            // base positions per CTA
constexpr int kPerThread = kTile / kThreads;
constexpr int kMaxW = 32;              // k - m + 1 <= 31
constexpr int kPackWords = (kTile + 32 + 15) / 16 + 2;

__device__ __forceinline__ void block_add(uint64_t* ctr, uint32_t found, uint32_t notfound, uint32_t invalid) {
	__shared__ uint32_t acc[3];
	if (threadIdx.x < 3) acc[threadIdx.x] = 0;
	__syncthreads();
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) {
		found += __shfl_xor_sync(0xffffffffu, found, o);
		notfound += __shfl_xor_sync(0xffffffffu, notfound, o);
		invalid += __shfl_xor_sync(0xffffffffu, invalid, o);
	}
	if ((threadIdx.x & 31) == 0) {
		if (found) atomicAdd(&acc[0], found);
		if (notfound) atomicAdd(&acc[1], notfound);
		if (invalid) atomicAdd(&acc[2], invalid);
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		if (acc[0]) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_FOUND], (unsigned long long)acc[0]);
		if (acc[1]) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_NOT_FOUND], (unsigned long long)acc[1]);
		if (acc[0] + acc[1]) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_QUERIES], (unsigned long long)(acc[0] + acc[1]));
		if (acc[2]) atomicAdd((unsigned long long*)&ctr[BLIGHT_CTR_INVALID], (unsigned long long)acc[2]);
	}
}

template <bool HAS_MINI>
__global__ void __launch_bounds__(kThreads) k_lookup_kmers(DevIndexView I, const uint64_t* __restrict__ canon,
                                                           const uint32_t* __restrict__ mini, uint64_t n,
                                                           int64_t* __restrict__ ids) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const uint64_t x = __ldg(canon + i);
		const uint32_t mn = HAS_MINI ? __ldg(mini + i) : minimizer_of_kmer(x, I.k, I.m);
		ids[i] = lookup_one(I, x, mn);
	}
}

enum ReadsMode { kEmitPairs = 0, kLookupIds = 1, kLookupCount = 2 };

// first read r in [lo, hi] with off[r+1] > p, i.e. the read containing base position p
__device__ __forceinline__ uint64_t find_read(const uint64_t* __restrict__ off, uint64_t lo, uint64_t hi, uint64_t p) {
	while (lo < hi) {
		const uint64_t mid = (lo + hi) >> 1;
		if (__ldg(off + mid + 1) > p) hi = mid; else lo = mid + 1;
	}
	return lo;
}

template <int MODE>
__global__ void __launch_bounds__(kThreads) k_reads(DevIndexView I, uint32_t k, uint32_t m, const char* __restrict__ bases,
                                                    const uint64_t* __restrict__ read_off, const uint64_t* __restrict__ read_end,
                                                    const uint64_t* __restrict__ kmer_off, uint64_t n_reads, uint64_t total_bases,
                                                    bool aligned16, uint64_t* __restrict__ out_canon,
                                                    uint32_t* __restrict__ out_mini, int64_t* __restrict__ out_ids,
                                                    uint64_t* __restrict__ ctr) {
	__shared__ uint32_t pack[kPackWords];           // 2-bit codes, 16 per word, first base in the high bits
	__shared__ uint32_t bad[kPackWords];            // 1 bit per base (bit 15-j of word i = base 16i+j): not ACGTacgt
	__shared__ uint32_t keys[kTile + kMaxW];        // ordering key of the m-mer starting at each tile position
	__shared__ uint64_t s_rlo, s_rhi;

	const uint64_t t0 = (uint64_t)blockIdx.x * kTile;
	const uint32_t w = k - m + 1;
	const uint32_t n_pos = (uint32_t)min((uint64_t)kTile, total_bases - t0);         // positions owned by this tile
	const uint32_t n_load = (uint32_t)min((uint64_t)(kTile + 32), total_bases - t0); // owned + halo (k-1 <= 30)
	// A. pack: thread i converts bases [16i, 16i+16) of the tile
	for (uint32_t i = threadIdx.x; i < kPackWords; i += kThreads) {
		const uint32_t b0 = i * 16;
		uint32_t word = 0, badw = 0;
		if (b0 < n_load) {
			unsigned char ch[16];
			if (aligned16 && b0 + 16 <= n_load) {
				const uint4 v = __ldg(reinterpret_cast<const uint4*>(bases + t0 + b0));  // t0 % 2048 == 0
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
		pack[i] = word;
		bad[i] = badw;
	}
	if (threadIdx.x == 0) {
		// reads overlapping this tile (for the per-position search below)
		const uint64_t lo = find_read(read_off, 0, n_reads - 1, t0);
		const uint64_t last = t0 + n_pos - 1;
		s_rlo = lo;
		s_rhi = find_read(read_off, lo, n_reads - 1, last);
	}
	__syncthreads();

	// B. m-mer keys
	const uint32_t mmask = (1u << (2 * m)) - 1u;
	for (uint32_t q = threadIdx.x; q < n_pos + w - 1 && q < kTile + kMaxW; q += kThreads) {
		const uint32_t wi = q >> 4, s = 2u * (q & 15);
		const uint32_t v = __funnelshift_l(pack[wi + 1], pack[wi], s) >> (32 - 2 * m);
		keys[q] = mini_key(parity_canon(v & mmask, m));
	}
	__syncthreads();

	// C. k-mers: lane-contiguous positions
	const uint64_t rlo = s_rlo, rhi = s_rhi;
	uint32_t found = 0, notfound = 0, invalid = 0;
	#pragma unroll 1
	for (int it = 0; it < kPerThread; it++) {
		const uint32_t q = it * kThreads + threadIdx.x;
		if (q >= n_pos) break;
		const uint64_t p = t0 + q;
		const uint64_t r = find_read(read_off, rlo, rhi, p);
		const uint64_t rbeg = __ldg(read_off + r), rend = read_end ? __ldg(read_end + r) : __ldg(read_off + r + 1);
		if (p < rbeg || p + k > rend) continue;  // no k-mer starts here (tail of a read, a gap, or a read shorter than k)
		const uint32_t wi = q >> 4, s = 2u * (q & 15);
		{
			// nuc2int rejects any byte outside ACGTacgt (kmer.h:56-69); only bases of queried k-mers are ever looked at
			const uint64_t bb = ((uint64_t)bad[wi] << 32) | ((uint64_t)bad[wi + 1] << 16) | bad[wi + 2];
			if ((bb >> (48 - (q & 15) - k)) & ((1ull << k) - 1)) { invalid++; continue; }
		}
		uint32_t best = keys[q];
		for (uint32_t j = 1; j < w; j++) best = min(best, keys[q + j]);
		const uint32_t a = pack[wi], b = pack[wi + 1], c = pack[wi + 2];
		const uint64_t top = ((uint64_t)__funnelshift_l(b, a, s) << 32) | __funnelshift_l(c, b, s);
		const uint64_t fwd = top >> (64 - 2 * k);
		const uint64_t rc = rc64(fwd, k);
		const uint64_t x = fwd < rc ? fwd : rc;
		const uint32_t mn = mini_from_key(best);
		const uint64_t o = __ldg(kmer_off + r) + (p - rbeg);
		if (MODE == kEmitPairs) {
			out_canon[o] = x;
			out_mini[o] = mn;
		} else {
			const int64_t id = lookup_one(I, x, mn);
			if (MODE == kLookupIds) out_ids[o] = id;
			if (id >= 0) found++; else notfound++;
		}
	}
	if (MODE == kEmitPairs) { found = 0; notfound = 0; }
	block_add(ctr, found, notfound, invalid);
}

int check(cudaError_t e) { return e == cudaSuccess ? 0 : BLIGHT_ERR_CUDA; }

}  // namespace

const char* g_last_cuda_error = "";

int launch_lookup_kmers(const DevIndexView& I, const uint64_t* d_canon, const uint32_t* d_mini, uint64_t n, int64_t* d_ids,
                        cudaStream_t stream) {
	if (n == 0) return 0;
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const uint64_t want = (n + kThreads - 1) / kThreads;
	const unsigned grid = (unsigned)(want < (uint64_t)sms * 64 ? want : (uint64_t)sms * 64);
	if (d_mini) k_lookup_kmers<true><<<grid, kThreads, 0, stream>>>(I, d_canon, d_mini, n, d_ids);
	else k_lookup_kmers<false><<<grid, kThreads, 0, stream>>>(I, d_canon, nullptr, n, d_ids);
	g_launches++;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) g_last_cuda_error = cudaGetErrorString(e);
	return check(e);
}

int launch_reads(const DevIndexView* I, uint32_t k, uint32_t m, const char* d_bases, const uint64_t* d_read_off,
                 const uint64_t* d_read_end, const uint64_t* d_kmer_off, uint64_t n_reads, uint64_t total_bases,
                 uint64_t* d_canon, uint32_t* d_mini, int64_t* d_ids, uint64_t* d_ctr, cudaStream_t stream) {
	if (n_reads == 0 || total_bases == 0) return 0;
	const uint64_t tiles = (total_bases + kTile - 1) / kTile;
	if (tiles > 0x7FFFFFFFull) return BLIGHT_ERR_INVALID_ARG;
	DevIndexView v{};
	if (I) v = *I;
	const bool al = (reinterpret_cast<uintptr_t>(d_bases) & 15) == 0;
	if (!I) k_reads<kEmitPairs><<<(unsigned)tiles, kThreads, 0, stream>>>(v, k, m, d_bases, d_read_off, d_read_end, d_kmer_off, n_reads, total_bases, al, d_canon, d_mini, nullptr, d_ctr);
	else if (d_ids) k_reads<kLookupIds><<<(unsigned)tiles, kThreads, 0, stream>>>(v, k, m, d_bases, d_read_off, d_read_end, d_kmer_off, n_reads, total_bases, al, nullptr, nullptr, d_ids, d_ctr);
	else k_reads<kLookupCount><<<(unsigned)tiles, kThreads, 0, stream>>>(v, k, m, d_bases, d_read_off, d_read_end, d_kmer_off, n_reads, total_bases, al, nullptr, nullptr, nullptr, d_ctr);
	g_launches++;
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess) g_last_cuda_error = cudaGetErrorString(e);
	return check(e);
}

}  // namespace blight
