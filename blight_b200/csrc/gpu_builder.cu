// gpu_builder.cu — construct_index on the GPU (kmer_Set_Light::construct_index, blight.cpp:108-125).
//
// Same result as the host builder (builder.cpp), i.e. bit for bit what the reference builds with cores=1 — the flat image
// is compared word for word in the tests — but organised for the device: no work items, no per-group loops, every step is a
// data-parallel pass over ALL sequences / super-k-mers / k-mers at once.
//
//   1. chop      the sequences are laid end to end in a virtual coordinate v; per position: 2-bit code + bases left in its
//                sequence, m-mer ordering key (kmer.h:791-810 with fix P1), window minimum over k-m+1 keys = minimizer of
//                the k-mer at v; a super-k-mer starts where the minimizer changes or a sequence begins (kmer.h:640-693)
//   2. order     stable LSD radix sort of the super-k-mers by minimizer (hand-written: warp-private digit histograms, one
//                scan, stable warp-level multisplit), which is the order the reference appends them to their buckets in
//                (blight.cpp:236-247, 311-351); prefix sums give every super-k-mer its nucleotide offset and key range
//   3. text      the bucket sequences, one 64-bit word of the vector<bool> image per thread (blight.cpp:311-324)
//   4. BBHash    all MPHF groups level by level: test-and-set of the level bit, a second bit array for collisions, keys whose
//                bit collided go on to the next level (bbhash.h:668-707 — the result depends only on the key SET, so the
//                bit arrays equal the reference's); leftovers after 16 levels go to the host for the fallback map
//   5. ranks     popcount per 16 words + one scan (bbhash.h:447-465)
//   6. positions field[rank(k-mer)] = offset in bucket >> b (blight.cpp:486-519), atomic OR into the bit-packed slab
//
// Level domains are the reference's double-precision recipe and stay on the host (bbhash.h:591-614). The input must be a
// k-mer SET (as BCALM unitigs are): a k-mer present twice is undefined here ("later writes win" in the reference).
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "capi_common.hpp"
#include "flat_index.hpp"
#include "kernels.hpp"
#include "kmer_math.hpp"

namespace blight {
namespace {

constexpr int kT = 256;
constexpr int kItems = 16;                 // elements per thread of a scan tile
constexpr uint64_t kTile = uint64_t(kT) * kItems;

int cu_fail(cudaError_t e, const char* what) { return fail(BL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); }
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cu_fail(e__, #call); } while (0)

unsigned grid_for(uint64_t n, uint64_t per_block = kT) { return (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n + per_block - 1) / per_block, 1u << 30)); }

// Every device buffer of one build, carved out of a few large slabs: a build needs ~40 buffers of up to a GB, and one
// cudaMalloc / cudaFree per buffer cost it tens to hundreds of milliseconds of driver time (measured 5 - 350 ms per phase;
// the stream-ordered pool was worse: 0.8 - 1.2 s while the pool grew). hint() announces what the next phase will ask for.
struct Arena {
	struct Slab { char* p; size_t cap, used; };
	std::vector<Slab> slabs;
	size_t next_hint = 0;
	void hint(size_t bytes) { next_hint = bytes + bytes / 16 + (1u << 20); }
	template <class T> int alloc(T** p, uint64_t n, bool zero = false) {
		const size_t bytes = (std::max<size_t>(size_t(n) * sizeof(T), 256) + 255) & ~size_t(255);
		if (slabs.empty() || slabs.back().used + bytes > slabs.back().cap) {
			const size_t cap = std::max(bytes, next_hint);
			void* q = nullptr;
			cudaError_t e = cudaMalloc(&q, cap);
			if (e != cudaSuccess) return fail(BL_ERR_NOMEM, std::string("cudaMalloc(GPU builder): ") + cudaGetErrorString(e));
			slabs.push_back(Slab{static_cast<char*>(q), cap, 0});
			next_hint = 0;
		}
		Slab& S = slabs.back();
		void* q = S.p + S.used;
		S.used += bytes;
		if (zero) { cudaError_t e = cudaMemsetAsync(q, 0, bytes, 0); if (e != cudaSuccess) return cu_fail(e, "cudaMemset"); }
		*p = static_cast<T*>(q);
		return BL_OK;
	}
	void release(void*) {}  // slabs go together
	~Arena() { for (Slab& S : slabs) cudaFree(S.p); }
};

// Pageable host memory <-> device through two pinned bounce buffers, the host side of every chunk copied by all cores while
// the DMA engine moves the previous one (a plain cudaMemcpy on pageable memory ran at 3.5 - 4.5 GB/s: 100 ms for the image).
struct Bounce {
	static constexpr size_t kChunk = 64u << 20;
	char* buf[2] = {nullptr, nullptr};
	cudaEvent_t ev[2] = {nullptr, nullptr};
	cudaStream_t st = nullptr;
	int threads = 1;
	bool ok = false;
	Bounce() {
		ok = cudaHostAlloc(reinterpret_cast<void**>(&buf[0]), kChunk, cudaHostAllocDefault) == cudaSuccess &&
		     cudaHostAlloc(reinterpret_cast<void**>(&buf[1]), kChunk, cudaHostAllocDefault) == cudaSuccess &&
		     cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) == cudaSuccess && cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) == cudaSuccess &&
		     cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
		if (!ok) cudaGetLastError();
		threads = std::max(1, std::min(16, (int)std::thread::hardware_concurrency()));
	}
	~Bounce() {
		for (char* b : buf) if (b) cudaFreeHost(b);
		for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
		if (st) cudaStreamDestroy(st);
	}
	void host_copy(char* dst, const char* src, size_t n) const {  // own threads, not OpenMP (a second runtime may live in the process)
		const int T = n < (4u << 20) ? 1 : threads;
		std::vector<std::thread> th;
		for (int t = 1; t < T; t++) th.emplace_back([=] { std::memcpy(dst + n * t / T, src + n * t / T, n * (t + 1) / T - n * t / T); });
		std::memcpy(dst, src, n / T);
		for (auto& x : th) x.join();
	}
	// both wait for everything queued on the legacy stream first (the data they move was produced / is consumed there)
	int h2d(void* dst, const void* src, size_t bytes) {
		if (!ok) { CU(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice)); return BL_OK; }
		CU(cudaStreamSynchronize(0));
		for (size_t o = 0, i = 0; o < bytes; o += kChunk, i++) {
			const size_t n = std::min(kChunk, bytes - o);
			if (i >= 2) CU(cudaEventSynchronize(ev[i & 1]));
			host_copy(buf[i & 1], static_cast<const char*>(src) + o, n);
			CU(cudaMemcpyAsync(static_cast<char*>(dst) + o, buf[i & 1], n, cudaMemcpyHostToDevice, st));
			CU(cudaEventRecord(ev[i & 1], st));
		}
		CU(cudaStreamSynchronize(st));
		return BL_OK;
	}
	int d2h(void* dst, const void* src, size_t bytes) {
		if (!ok) { CU(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost)); return BL_OK; }
		CU(cudaStreamSynchronize(0));
		const size_t n_chunks = (bytes + kChunk - 1) / kChunk;
		for (size_t i = 0; i <= n_chunks; i++) {
			if (i < n_chunks) {  // chunk i on its way while chunk i - 1 is copied out of its buffer below
				const size_t o = i * kChunk;
				CU(cudaMemcpyAsync(buf[i & 1], static_cast<const char*>(src) + o, std::min(kChunk, bytes - o), cudaMemcpyDeviceToHost, st));
				CU(cudaEventRecord(ev[i & 1], st));
			}
			if (i > 0) {
				const size_t o = (i - 1) * kChunk;
				CU(cudaEventSynchronize(ev[(i - 1) & 1]));
				host_copy(static_cast<char*>(dst) + o, buf[(i - 1) & 1], std::min(kChunk, bytes - o));
			}
		}
		return BL_OK;
	}
};

// ---- exclusive scan: out[i] = sum of in[0, i), block sums in bsum ----------------------------------------------------
template <class In>
__global__ void __launch_bounds__(kT) k_scan_reduce(const In* __restrict__ in, uint64_t n, uint64_t* __restrict__ bsum) {
	__shared__ uint64_t s[kT / 32];
	const uint64_t base = (uint64_t)blockIdx.x * kTile + (uint64_t)threadIdx.x * kItems;
	uint64_t a = 0;
	#pragma unroll
	for (int j = 0; j < kItems; j++) if (base + j < n) a += (uint64_t)in[base + j];
	#pragma unroll
	for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
	if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = a;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint64_t t = 0;
		for (int w = 0; w < kT / 32; w++) t += s[w];
		bsum[blockIdx.x] = t;
	}
}

// single block: exclusive scan of bsum[0, nb) in place, total to bsum[nb]
__global__ void __launch_bounds__(1024) k_scan_blocks(uint64_t* bsum, uint64_t nb) {
	__shared__ uint64_t s[1024];
	const uint64_t per = (nb + 1023) / 1024;
	const uint64_t lo = min(nb, (uint64_t)threadIdx.x * per), hi = min(nb, lo + per);
	uint64_t a = 0;
	for (uint64_t i = lo; i < hi; i++) a += bsum[i];
	s[threadIdx.x] = a;
	__syncthreads();
	for (int o = 1; o < 1024; o <<= 1) {
		const uint64_t t = threadIdx.x >= (unsigned)o ? s[threadIdx.x - o] : 0;
		__syncthreads();
		s[threadIdx.x] += t;
		__syncthreads();
	}
	uint64_t run = s[threadIdx.x] - a;
	for (uint64_t i = lo; i < hi; i++) { const uint64_t v = bsum[i]; bsum[i] = run; run += v; }
	if (threadIdx.x == 1023) bsum[nb] = s[1023];
}

template <class In>
__global__ void __launch_bounds__(kT) k_scan_apply(const In* __restrict__ in, uint64_t n, const uint64_t* __restrict__ bsum, uint64_t* __restrict__ out) {
	__shared__ uint64_t s[kT];
	const uint64_t base = (uint64_t)blockIdx.x * kTile + (uint64_t)threadIdx.x * kItems;
	uint64_t v[kItems], a = 0;
	#pragma unroll
	for (int j = 0; j < kItems; j++) { v[j] = base + j < n ? (uint64_t)in[base + j] : 0; a += v[j]; }
	s[threadIdx.x] = a;
	__syncthreads();
	for (int o = 1; o < kT; o <<= 1) {
		const uint64_t t = threadIdx.x >= (unsigned)o ? s[threadIdx.x - o] : 0;
		__syncthreads();
		s[threadIdx.x] += t;
		__syncthreads();
	}
	uint64_t run = bsum[blockIdx.x] + s[threadIdx.x] - a;
	#pragma unroll
	for (int j = 0; j < kItems; j++) { if (base + j < n) out[base + j] = run; run += v[j]; }
}

struct Scan {
	uint64_t* bsum = nullptr;
	uint64_t cap = 0;
};

template <class In>
int exclusive_scan(Arena& A, Scan& S, const In* d_in, uint64_t n, uint64_t* d_out, uint64_t* total) {
	const uint64_t nb = std::max<uint64_t>(1, (n + kTile - 1) / kTile);
	if (S.cap < nb + 1) {
		int rc = A.alloc(&S.bsum, nb + 1 + nb / 4);
		if (rc != BL_OK) return rc;
		S.cap = nb + 1 + nb / 4;
	}
	k_scan_reduce<In><<<(unsigned)nb, kT>>>(d_in, n, S.bsum);
	k_scan_blocks<<<1, 1024>>>(S.bsum, nb);
	k_scan_apply<In><<<(unsigned)nb, kT>>>(d_in, n, S.bsum, d_out);
	g_launches += 3;
	CU(cudaGetLastError());
	if (total) CU(cudaMemcpy(total, S.bsum + nb, 8, cudaMemcpyDeviceToHost));
	return BL_OK;
}

// ---- 1. chop -------------------------------------------------------------------------------------------------------------
// view of v: last s with vstart[s] <= v
__device__ __forceinline__ uint64_t view_of(const uint64_t* __restrict__ vstart, uint64_t n_views, uint64_t v) {
	uint64_t lo = 0, hi = n_views - 1;
	while (lo < hi) {
		const uint64_t mid = (lo + hi + 1) >> 1;
		if (__ldg(vstart + mid) <= v) lo = mid; else hi = mid - 1;
	}
	return lo;
}

__global__ void __launch_bounds__(kT) k_codes(const char* __restrict__ text, const uint64_t* __restrict__ starts, const uint64_t* __restrict__ vstart,
                                              uint64_t n_views, uint64_t total_v, uint8_t* __restrict__ codes, uint8_t* __restrict__ rem, uint32_t* err) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total_v; v += stride) {
		const uint64_t s = view_of(vstart, n_views, v);
		const uint64_t off = v - __ldg(vstart + s), left = __ldg(vstart + s + 1) - v;
		const uint32_t c = nuc_code((unsigned char)text[__ldg(starts + s) + off]);
		if (c > 3) atomicOr(err, 1u);
		codes[v] = (uint8_t)(c & 3u);
		rem[v] = (uint8_t)(left > 255 ? 255 : left);
	}
}

__global__ void __launch_bounds__(kT) k_mkeys(const uint8_t* __restrict__ codes, const uint8_t* __restrict__ rem, uint64_t total_v, uint32_t m,
                                              uint32_t* __restrict__ key) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total_v; v += stride) {
		uint32_t kk = 0xFFFFFFFFu;
		if (rem[v] >= m) {
			uint32_t x = 0;
			for (uint32_t j = 0; j < m; j++) x = (x << 2) | codes[v + j];
			kk = mini_key(parity_canon(x, m));
		}
		key[v] = kk;
	}
}

// minimizer (as ordering key) of the k-mer at v, 0xFFFFFFFF where no k-mer starts; flag: a super-k-mer starts at v
__global__ void __launch_bounds__(kT) k_kmin(const uint32_t* __restrict__ key, const uint8_t* __restrict__ rem, uint64_t total_v, uint32_t k, uint32_t w,
                                             uint32_t* __restrict__ kmin) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total_v; v += stride) {
		uint32_t best = 0xFFFFFFFFu;
		if (rem[v] >= k)
			for (uint32_t j = 0; j < w; j++) best = min(best, __ldg(key + v + j));
		kmin[v] = best;
	}
}

__global__ void __launch_bounds__(kT) k_flags(const uint32_t* __restrict__ kmin, const uint8_t* __restrict__ rem, uint64_t total_v, uint32_t k,
                                              uint8_t* __restrict__ flag) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total_v; v += stride) {
		uint8_t f = 0;
		if (rem[v] >= k) f = (v == 0 || rem[v - 1] == 1 || kmin[v] != kmin[v - 1]) ? 1 : 0;
		flag[v] = f;
	}
}

__global__ void __launch_bounds__(kT) k_sk_records(const uint8_t* __restrict__ flag, const uint64_t* __restrict__ skidx, const uint32_t* __restrict__ kmin,
                                                   uint64_t total_v, uint64_t* __restrict__ sk_v, uint32_t* __restrict__ sk_mini) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; v < total_v; v += stride)
		if (flag[v]) { const uint64_t j = skidx[v]; sk_v[j] = v; sk_mini[j] = mini_from_key(kmin[v]); }
}

// k-mers of every super-k-mer: up to the next one, or to the last k-mer of its sequence; bucket totals on the way
__global__ void __launch_bounds__(kT) k_sk_len(const uint64_t* __restrict__ sk_v, const uint32_t* __restrict__ sk_mini, uint64_t n_sk,
                                               const uint64_t* __restrict__ vstart, uint64_t n_views, uint32_t k, uint32_t* __restrict__ sk_nk,
                                               unsigned long long* __restrict__ bnuc, unsigned long long* __restrict__ bkm) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_sk; j += stride) {
		const uint64_t v = sk_v[j];
		const uint64_t s = view_of(vstart, n_views, v);
		const uint64_t kend = __ldg(vstart + s + 1) - k + 1;  // one past the last k-mer start of the sequence
		const uint64_t next = j + 1 < n_sk ? sk_v[j + 1] : ~0ull;
		const uint32_t nk = (uint32_t)(min(next, kend) - v);
		sk_nk[j] = nk;
		atomicAdd(bnuc + sk_mini[j], (unsigned long long)nk + k - 1);
		atomicAdd(bkm + sk_mini[j], (unsigned long long)nk);
	}
}

// ---- 2. stable LSD radix sort of (key, value) pairs, 8 bits per pass; a warp owns a tile of kRadixTile elements ---------------
constexpr int kRadixTile = 2048;

__global__ void __launch_bounds__(kT) k_radix_hist(const uint32_t* __restrict__ keys, uint64_t n, uint32_t shift, uint64_t n_tiles, uint32_t* __restrict__ ghist) {
	__shared__ uint32_t hist[kT / 32][256];
	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint64_t tile = (uint64_t)blockIdx.x * (kT / 32) + wid;
	for (int d = lane; d < 256; d += 32) hist[wid][d] = 0;
	__syncwarp();
	if (tile < n_tiles) {
		const uint64_t lo = tile * kRadixTile, hi = min(n, lo + kRadixTile);
		for (uint64_t i = lo + lane; i < hi; i += 32) atomicAdd(&hist[wid][(keys[i] >> shift) & 255u], 1u);
		__syncwarp();
		for (int d = lane; d < 256; d += 32) ghist[(uint64_t)d * n_tiles + tile] = hist[wid][d];
	}
}

__global__ void __launch_bounds__(kT) k_radix_scatter(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, uint64_t n, uint32_t shift,
                                                      uint64_t n_tiles, const uint64_t* __restrict__ goff, uint32_t* __restrict__ keys_out,
                                                      uint32_t* __restrict__ vals_out) {
	__shared__ uint32_t base[kT / 32][256];
	const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const uint64_t tile = (uint64_t)blockIdx.x * (kT / 32) + wid;
	if (tile >= n_tiles) return;
	for (int d = lane; d < 256; d += 32) base[wid][d] = (uint32_t)goff[(uint64_t)d * n_tiles + tile];
	__syncwarp();
	const uint64_t lo = tile * kRadixTile, hi = min(n, lo + kRadixTile);
	const uint32_t lt = (1u << lane) - 1u;
	for (uint64_t i0 = lo; i0 < hi; i0 += 32) {
		const uint64_t i = i0 + lane;
		const bool act = i < hi;
		const uint32_t kk = act ? keys[i] : 0u, vv = act ? vals[i] : 0u;
		const uint32_t d = act ? ((kk >> shift) & 255u) : 256u + lane;  // idle lanes match nobody
		const uint32_t peers = __match_any_sync(0xffffffffu, d);
		uint32_t b = 0;
		if (act) b = base[wid][d];
		__syncwarp();
		if (act && lane == (uint32_t)(__ffs(peers) - 1)) base[wid][d] = b + __popc(peers);
		__syncwarp();
		if (act) {
			const uint32_t o = b + __popc(peers & lt);  // lanes hold consecutive elements: order inside a digit is kept
			keys_out[o] = kk;
			vals_out[o] = vv;
		}
	}
}

__global__ void __launch_bounds__(kT) k_iota(uint32_t* a, uint64_t n) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) a[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(kT) k_gather_len(const uint32_t* __restrict__ perm, const uint32_t* __restrict__ sk_nk, uint64_t n_sk, uint32_t k,
                                                   uint32_t* __restrict__ len_sorted, uint32_t* __restrict__ nk_sorted) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_sk; i += stride) {
		const uint32_t nk = sk_nk[perm[i]];
		nk_sorted[i] = nk;
		len_sorted[i] = nk + k - 1;
	}
}

// ---- 3. bucket sequences: one 64-bit word (32 nucleotides) of the vector<bool> image per thread ----------------------------
__global__ void __launch_bounds__(kT) k_seq_words(const uint64_t* __restrict__ dest, uint64_t n_sk, const uint32_t* __restrict__ perm,
                                                  const uint64_t* __restrict__ sk_v, const uint8_t* __restrict__ codes, uint64_t total_nuc,
                                                  uint64_t n_words, uint64_t* __restrict__ seq) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += stride) {
		uint64_t p = w << 5;
		const uint64_t pend = min(total_nuc, p + 32);
		// super-k-mer holding p: last i with dest[i] <= p
		uint64_t lo = 0, hi = n_sk - 1;
		while (lo < hi) {
			const uint64_t mid = (lo + hi + 1) >> 1;
			if (__ldg(dest + mid) <= p) lo = mid; else hi = mid - 1;
		}
		uint64_t i = lo, d0 = __ldg(dest + i), dn = i + 1 < n_sk ? __ldg(dest + i + 1) : ~0ull, src = __ldg(sk_v + perm[i]);
		uint64_t v = 0;
		for (; p < pend; p++) {
			if (p >= dn) { i++; d0 = dn; dn = i + 1 < n_sk ? __ldg(dest + i + 1) : ~0ull; src = __ldg(sk_v + perm[i]); }
			const uint64_t c = codes[src + (p - d0)];
			v |= (((c >> 1) & 1) | ((c & 1) << 1)) << (2 * (p & 31));  // nucleotide p -> bit 2p = code >> 1, bit 2p+1 = code & 1
		}
		seq[w] = v;
	}
}

// ---- canonical k-mers in bucket order, with their MPHF group and offset in the bucket ---------------------------------------
__global__ void __launch_bounds__(kT) k_keys(const uint32_t* __restrict__ perm, const uint64_t* __restrict__ sk_v, const uint32_t* __restrict__ mini_sorted,
                                             const uint32_t* __restrict__ nk_sorted, const uint64_t* __restrict__ dest, const uint64_t* __restrict__ keybase,
                                             const uint64_t* __restrict__ bucket_start, const uint8_t* __restrict__ codes, uint64_t n_sk, uint32_t k, uint32_t lb,
                                             uint64_t* __restrict__ keys, uint32_t* __restrict__ kgroup, uint32_t* __restrict__ koff) {
	const uint64_t kmask = (1ull << (2 * k)) - 1;
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_sk; i += stride) {
		const uint64_t src = sk_v[perm[i]];
		const uint32_t nk = nk_sorted[i], mn = mini_sorted[i];
		const uint64_t kb = keybase[i];
		const uint32_t off0 = (uint32_t)(dest[i] - bucket_start[mn]);
		uint64_t fwd = 0, rev = 0;
		for (uint32_t j = 0; j < nk + k - 1; j++) {
			const uint64_t c = codes[src + j];
			fwd = ((fwd << 2) | c) & kmask;
			rev = (rev >> 2) | ((c ^ 2) << (2 * k - 2));
			if (j + 1 >= k) {
				const uint64_t t = kb + (j + 1 - k);
				keys[t] = fwd < rev ? fwd : rev;
				kgroup[t] = mn >> lb;
				koff[t] = off0 + (j + 1 - k);
			}
		}
	}
}

// ---- 4. BBHash levels ----------------------------------------------------------------------------------------------------
struct GroupDev {
	uint64_t bits_base;           // first bit of the group in the global bit array (multiple of 64)
	uint64_t level_off[kLevels];  // first bit of every level inside the group
	uint64_t dom[kLevels];
	uint64_t pos_start;
	uint64_t block_base;          // first 16-word block of the group in the global block numbering
	uint64_t n_words;
	uint32_t nbits, pad;
};

__device__ __forceinline__ uint64_t level_hash(uint64_t key, int level) {
	uint64_t s0 = hash_bis(key, kSeed0);
	if (level == 0) return s0;
	uint64_t s1 = hash_bis(key, kSeed1);
	if (level == 1) return s1;
	uint64_t h = 0;
	for (int l = 2; l <= level; l++) h = xs128_next(s0, s1);
	return h;
}

__device__ __forceinline__ uint64_t level_bit(const GroupDev& G, uint64_t key, int level) {
	return G.bits_base + G.level_off[level] + __umul64hi(level_hash(key, level), G.dom[level]);
}

__global__ void __launch_bounds__(kT) k_level_place(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ kgroup, const uint64_t* __restrict__ list,
                                                    uint64_t n_items, int level, const GroupDev* __restrict__ groups, uint32_t* bits, uint32_t* coll,
                                                    uint32_t* __restrict__ gcoll) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items; i += stride) {
		const uint64_t ki = list ? list[i] : i;
		const uint32_t g = kgroup[ki];
		const uint64_t bit = level_bit(groups[g], keys[ki], level);
		const uint32_t msk = 1u << (bit & 31);
		if (atomicOr(bits + (bit >> 5), msk) & msk) {  // somebody was here first: nobody keeps this bit (bbhash.h:668-707)
			atomicOr(coll + (bit >> 5), msk);
			gcoll[g] = 1;
		}
	}
}

__global__ void __launch_bounds__(kT) k_level_sift(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ kgroup, const uint64_t* __restrict__ list,
                                                   uint64_t n_items, int level, const GroupDev* __restrict__ groups, const uint32_t* __restrict__ coll,
                                                   uint64_t* __restrict__ next, unsigned long long* __restrict__ n_next, uint64_t* __restrict__ final_bit) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const uint64_t n_round = (n_items + 31) & ~31ull;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += stride) {
		bool left = false;
		uint64_t ki = 0;
		if (i < n_items) {
			ki = list ? list[i] : i;
			const uint64_t bit = level_bit(groups[kgroup[ki]], keys[ki], level);
			left = (coll[bit >> 5] >> (bit & 31)) & 1u;
			if (!left) final_bit[ki] = bit;
		}
		const uint32_t lm = __ballot_sync(0xffffffffu, left);
		if (lm) {
			const uint32_t lane = threadIdx.x & 31;
			unsigned long long b = 0;
			if (lane == 0) b = atomicAdd(n_next, (unsigned long long)__popc(lm));
			b = __shfl_sync(0xffffffffu, b, 0);
			if (left) next[b + __popc(lm & ((1u << lane) - 1u))] = ki;
		}
	}
}

__global__ void __launch_bounds__(kT) k_level_clear(uint32_t* __restrict__ bits, uint32_t* __restrict__ coll, uint64_t n_words32) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words32; i += stride) {
		const uint32_t c = coll[i];
		if (c) { bits[i] &= ~c; coll[i] = 0; }
	}
}

// a group is finished at the first level none of its keys collided at (bbhash.h:709-728: the ranks stop there)
__global__ void __launch_bounds__(kT) k_level_groups(uint32_t* __restrict__ gcoll, uint32_t* __restrict__ flevel, uint64_t n_groups, int level) {
	const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= n_groups) return;
	if (flevel[g] == 0xFFFFFFFFu && !gcoll[g]) flevel[g] = (uint32_t)level;
	gcoll[g] = 0;
}

// ---- 5. ranks -------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) k_block_pop(const GroupDev* __restrict__ groups, const uint64_t* __restrict__ gblock_first, uint64_t n_groups,
                                                  uint64_t n_blocks, const uint64_t* __restrict__ bits64, uint32_t* __restrict__ blockpop) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t gb = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gb < n_blocks; gb += stride) {
		uint64_t lo = 0, hi = n_groups - 1;  // last group whose first block is <= gb (empty groups repeat the next one's)
		while (lo < hi) {
			const uint64_t mid = (lo + hi + 1) >> 1;
			if (__ldg(gblock_first + mid) <= gb) lo = mid; else hi = mid - 1;
		}
		const GroupDev& G = groups[lo];
		const uint64_t w0 = (gb - G.block_base) * 16, w1 = min(G.n_words, w0 + 16);
		const uint64_t* p = bits64 + (G.bits_base >> 6);
		uint32_t c = 0;
		for (uint64_t w = w0; w < w1; w++) c += (uint32_t)__popcll(p[w]);
		blockpop[gb] = c;
	}
}

// ---- 6. positions ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kT) k_positions(const uint32_t* __restrict__ kgroup, const uint32_t* __restrict__ koff, const uint64_t* __restrict__ final_bit,
                                                  uint64_t n_keys, const GroupDev* __restrict__ groups, const uint64_t* __restrict__ bits64,
                                                  const uint64_t* __restrict__ bscan, uint32_t b, unsigned long long* __restrict__ pos) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t ki = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; ki < n_keys; ki += stride) {
		const uint64_t fb = final_bit[ki];
		const uint32_t off = koff[ki];
		if (fb == ~0ull || off == 0) continue;  // fallback keys are patched on the host; offset 0 is never written (blight.cpp:486-519)
		const GroupDev& G = groups[kgroup[ki]];
		const uint64_t rel = fb - G.bits_base, wi = rel >> 6;
		const uint64_t* p = bits64 + (G.bits_base >> 6);
		uint64_t rank = bscan[G.block_base + (wi >> 4)] - bscan[G.block_base];
		for (uint64_t x = wi & ~15ull; x < wi; x++) rank += (uint64_t)__popcll(p[x]);
		rank += (uint64_t)__popcll(p[wi] & ((1ull << (rel & 63)) - 1));
		const uint32_t nb = G.nbits;
		uint64_t v = ((uint64_t)off >> b) & (nb >= 64 ? ~0ull : ((1ull << nb) - 1));
		uint64_t bitpos = G.pos_start + rank * nb;
		uint32_t left = nb;
		while (left) {
			const uint32_t sh = (uint32_t)(bitpos & 63), take = min(left, 64u - sh);
			atomicOr(pos + (bitpos >> 6), (unsigned long long)(v << sh));
			v = take >= 64 ? 0 : v >> take;
			bitpos += take;
			left -= take;
		}
	}
}

__global__ void __launch_bounds__(kT) k_ranks_out(const GroupDev* __restrict__ groups, const uint64_t* __restrict__ gblock_first, uint64_t n_groups,
                                                  uint64_t n_blocks, const uint64_t* __restrict__ bscan, uint64_t* __restrict__ ranks) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t gb = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; gb < n_blocks; gb += stride) {
		uint64_t lo = 0, hi = n_groups - 1;
		while (lo < hi) {
			const uint64_t mid = (lo + hi + 1) >> 1;
			if (__ldg(gblock_first + mid) <= gb) lo = mid; else hi = mid - 1;
		}
		ranks[gb] = bscan[gb] - bscan[groups[lo].block_base];
	}
}

// Level domains, the reference's double-precision recipe (bbhash.h:591-614) with gamma = 2 (blight.h:60).
__attribute__((optimize("fp-contract=off"))) void level_domains_host(uint64_t nelem, uint64_t dom[kLevels]) {
	const double gamma = 2.0;
	double proba_collision = 1.0 - pow(((gamma * (double)nelem - 1) / (gamma * (double)nelem)), (double)(nelem - 1));
	size_t hash_domain = (size_t)(ceil(double(nelem) * gamma));
	for (unsigned ii = 0; ii < (unsigned)kLevels; ii++) {
		dom[ii] = (((uint64_t)(hash_domain * pow(proba_collision, (double)ii)) + 63) / 64) * 64;
		if (dom[ii] == 0) dom[ii] = 64;
	}
}

void host_write_field(std::vector<uint64_t>& pos, uint64_t bitpos, unsigned nbits, uint64_t val) {
	unsigned left = nbits;
	while (left) {
		const uint64_t wi = bitpos >> 6;
		const unsigned sh = unsigned(bitpos & 63), take = std::min<unsigned>(left, 64 - sh);
		const uint64_t fm = ((take >= 64) ? ~0ull : ((1ull << take) - 1)) << sh;
		pos[wi] = (pos[wi] & ~fm) | ((val << sh) & fm);
		val = take >= 64 ? 0 : val >> take;
		bitpos += take;
		left -= take;
	}
}

}  // namespace

// sequences = views [starts[i], starts[i] + lens[i]) of text[0, text_len) (host memory); may overlap
int build_flat_index_gpu(const char* text, uint64_t text_len, const std::vector<uint64_t>& starts_in, const std::vector<uint64_t>& lens_in,
                         const BuildParams& P, int device, FlatIndex& F, std::string* err, double* seconds_device) {
	auto set_err = [&](const std::string& s) { if (err) *err = s; };
	int rc = check_params(P, err);
	if (rc != BL_OK) return rc;
	const unsigned k = P.k, m = P.m, b = P.b;
	if (k < m || k - m + 1 > 32) { set_err("k - m + 1 must not exceed 32"); return BL_ERR_INVALID_ARG; }
	int n_dev = 0;
	if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { set_err("no CUDA device available (this library has no CPU fallback)"); return BL_ERR_NO_DEVICE; }
	if (device < 0 || device >= n_dev) { set_err("device ordinal out of range"); return BL_ERR_INVALID_ARG; }
	int prev = -1;
	cudaGetDevice(&prev);
	cudaSetDevice(device);
	struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev};

	F = FlatIndex();
	FlatHeader& H = F.h;
	std::memcpy(H.magic, "BLFLAT01", 8);
	H.k = k; H.m = m; H.n_log2 = P.n_log2; H.s_log2 = P.s_log2; H.b = b;
	H.n_buckets = 1ull << (2 * m - 1);
	H.n_mphf = 1ull << P.n_log2;
	const unsigned lb = F.lb();

	// sequences shorter than k are skipped (undefined in the reference, kmer.h:705)
	std::vector<uint64_t> starts, vstart(1, 0);
	for (size_t i = 0; i < starts_in.size(); i++)
		if (lens_in[i] >= k) { starts.push_back(starts_in[i]); vstart.push_back(vstart.back() + lens_in[i]); }
	const uint64_t n_views = starts.size(), total_v = vstart.back();
	auto finish_empty = [&]() {
		F.bucket_start.assign(H.n_buckets, 0);
		F.bucket_nuc.assign(H.n_buckets, 0);
		F.mphf.assign(H.n_mphf, MphfRec{});
		uint64_t tp = 0;
		for (auto& r : F.mphf) { r.nbits = 1; r.pos_start = tp; tp += 8; }
		H.positions_bits = tp; H.pos_words = (tp + 63) / 64;
		F.pos.assign(H.pos_words, 0);
		return flat_validate(F, err);
	};
	if (n_views == 0) return finish_empty();

	// BLIGHT_BUILD_DEBUG=1: wall clock of every phase on stderr (each mark synchronises the device)
	const bool dbg = [] { const char* e = getenv("BLIGHT_BUILD_DEBUG"); return e && atoi(e); }();
	double t_mark = 0;
	auto mark = [&](const char* what) {
		if (!dbg) return;
		cudaDeviceSynchronize();
		const double t = std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
		if (t_mark > 0) fprintf(stderr, "[blight build] %-28s %8.2f ms\n", what, 1e3 * (t - t_mark));
		t_mark = t;
	};
	mark("start");
	Arena A;
	Bounce X;
	Scan S;
	cudaEvent_t ev0, ev1;
	CU(cudaEventCreate(&ev0));
	CU(cudaEventCreate(&ev1));
	struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evg{ev0, ev1};
	CU(cudaEventRecord(ev0, 0));

	char* d_text; uint64_t *d_starts, *d_vstart; uint8_t *d_codes, *d_rem, *d_flag; uint32_t *d_key, *d_kmin, *d_err; uint64_t* d_skidx;
	A.hint(text_len + n_views * 16 + total_v * 19 + (total_v / kTile + 2) * 10 + 4096);
#define AL(p, n, ...) do { rc = A.alloc(&p, n, ##__VA_ARGS__); if (rc != BL_OK) { set_err(g_last_error); return rc; } } while (0)
	AL(d_text, text_len + 64);
	AL(d_starts, n_views);
	AL(d_vstart, n_views + 1);
	AL(d_codes, total_v + 64, true);
	AL(d_rem, total_v + 64, true);
	AL(d_key, total_v + 64);
	AL(d_kmin, total_v);
	AL(d_flag, total_v);
	AL(d_skidx, total_v);
	AL(d_err, 1, true);
	if ((rc = X.h2d(d_text, text, text_len)) != BL_OK) { set_err(g_last_error); return rc; }
	CU(cudaMemcpy(d_starts, starts.data(), n_views * 8, cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d_vstart, vstart.data(), (n_views + 1) * 8, cudaMemcpyHostToDevice));
	CU(cudaMemset(d_key, 0xFF, (total_v + 64) * 4));

	mark("alloc + H2D");
	// 1. chop
	const unsigned gv = grid_for(total_v, kT * 4);
	k_codes<<<gv, kT>>>(d_text, d_starts, d_vstart, n_views, total_v, d_codes, d_rem, d_err);
	k_mkeys<<<gv, kT>>>(d_codes, d_rem, total_v, m, d_key);
	k_kmin<<<gv, kT>>>(d_key, d_rem, total_v, k, k - m + 1, d_kmin);
	k_flags<<<gv, kT>>>(d_kmin, d_rem, total_v, k, d_flag);
	g_launches += 4;
	CU(cudaGetLastError());
	mark("chop kernels");
	uint32_t h_err = 0;
	CU(cudaMemcpy(&h_err, d_err, 4, cudaMemcpyDeviceToHost));
	if (h_err) { set_err("Invalid char in DNA"); return BL_ERR_INVALID_BASE; }
	A.release(d_text); A.release(d_key);
	uint64_t n_sk = 0;
	if ((rc = exclusive_scan(A, S, d_flag, total_v, d_skidx, &n_sk)) != BL_OK) { set_err(g_last_error); return rc; }
	if (n_sk >= 0xFFFFFFFFull) { set_err("more than 2^32 super-k-mers"); return BL_ERR_INVALID_ARG; }
	H.number_super_kmer = n_sk;

	uint64_t* d_sk_v; uint32_t *d_sk_mini, *d_sk_nk; unsigned long long *d_bnuc, *d_bkm;
	A.hint(n_sk * 56 + H.n_buckets * 24 + ((n_sk + kRadixTile - 1) / kRadixTile) * 256 * 13 + H.n_mphf * (sizeof(GroupDev) + 16) + 65536);
	AL(d_sk_v, n_sk);
	AL(d_sk_mini, n_sk);
	AL(d_sk_nk, n_sk);
	AL(d_bnuc, H.n_buckets, true);
	AL(d_bkm, H.n_buckets, true);
	const unsigned gs = grid_for(n_sk);
	k_sk_records<<<gv, kT>>>(d_flag, d_skidx, d_kmin, total_v, d_sk_v, d_sk_mini);
	k_sk_len<<<gs, kT>>>(d_sk_v, d_sk_mini, n_sk, d_vstart, n_views, k, d_sk_nk, d_bnuc, d_bkm);
	g_launches += 2;
	CU(cudaGetLastError());
	A.release(d_kmin); A.release(d_flag); A.release(d_skidx);

	mark("super-k-mer records");
	// bucket table
	std::vector<uint64_t> bnuc(H.n_buckets), bkm(H.n_buckets);
	CU(cudaMemcpy(bnuc.data(), d_bnuc, H.n_buckets * 8, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(bkm.data(), d_bkm, H.n_buckets * 8, cudaMemcpyDeviceToHost));
	F.bucket_start.resize(H.n_buckets);
	F.bucket_nuc.resize(H.n_buckets);
	uint64_t acc = 0;
	for (uint64_t i = 0; i < H.n_buckets; i++) {
		if (bnuc[i] > 0xFFFFFFFFull) { set_err("a minimizer bucket exceeds 2^32 nucleotides (blight.h:33); use a larger m"); return BL_ERR_INVALID_ARG; }
		F.bucket_start[i] = acc;
		F.bucket_nuc[i] = uint32_t(bnuc[i]);
		acc += bnuc[i];
		H.number_kmer += bkm[i];
	}
	H.total_nuc = acc;
	H.seq_words = (acc * 2 + 63) / 64;
	const uint64_t N = H.number_kmer;
	uint64_t* d_bucket_start;
	AL(d_bucket_start, H.n_buckets);
	CU(cudaMemcpy(d_bucket_start, F.bucket_start.data(), H.n_buckets * 8, cudaMemcpyHostToDevice));

	mark("bucket table (D2H + host)");
	// 2. order: stable radix sort by minimizer, 8 bits per pass
	uint32_t *d_ka, *d_kb, *d_va, *d_vb, *d_ghist; uint64_t* d_goff;
	const uint64_t n_tiles = (n_sk + kRadixTile - 1) / kRadixTile;
	AL(d_ka, n_sk); AL(d_kb, n_sk); AL(d_va, n_sk); AL(d_vb, n_sk);
	AL(d_ghist, 256 * n_tiles);
	AL(d_goff, 256 * n_tiles);
	CU(cudaMemcpy(d_ka, d_sk_mini, n_sk * 4, cudaMemcpyDeviceToDevice));
	k_iota<<<gs, kT>>>(d_va, n_sk);
	g_launches++;
	const unsigned gt = grid_for(n_tiles, kT / 32);
	for (unsigned shift = 0; shift < 2 * m - 1; shift += 8) {
		k_radix_hist<<<gt, kT>>>(d_ka, n_sk, shift, n_tiles, d_ghist);
		g_launches++;
		if ((rc = exclusive_scan(A, S, d_ghist, 256 * n_tiles, d_goff, nullptr)) != BL_OK) { set_err(g_last_error); return rc; }
		k_radix_scatter<<<gt, kT>>>(d_ka, d_va, n_sk, shift, n_tiles, d_goff, d_kb, d_vb);
		g_launches++;
		std::swap(d_ka, d_kb);
		std::swap(d_va, d_vb);
	}
	CU(cudaGetLastError());
	uint32_t* d_mini_sorted = d_ka;
	uint32_t* d_perm = d_va;
	uint32_t *d_len_sorted, *d_nk_sorted; uint64_t *d_dest, *d_keybase;
	AL(d_len_sorted, n_sk); AL(d_nk_sorted, n_sk); AL(d_dest, n_sk); AL(d_keybase, n_sk);
	k_gather_len<<<gs, kT>>>(d_perm, d_sk_nk, n_sk, k, d_len_sorted, d_nk_sorted);
	g_launches++;
	if ((rc = exclusive_scan(A, S, d_len_sorted, n_sk, d_dest, nullptr)) != BL_OK) { set_err(g_last_error); return rc; }
	if ((rc = exclusive_scan(A, S, d_nk_sorted, n_sk, d_keybase, nullptr)) != BL_OK) { set_err(g_last_error); return rc; }

	mark("radix sort + prefix sums");
	// MPHF group descriptors (blight.cpp:280-306) and level domains
	F.mphf.assign(H.n_mphf, MphfRec{});
	std::vector<GroupDev> groups(H.n_mphf);
	std::vector<uint64_t> gblock_first(H.n_mphf + 1, 0);
	uint64_t total_bits = 0, total_blocks = 0;
	{
		uint64_t total_pos = 0, id_base = 0;
		for (uint64_t g = 0; g < H.n_mphf; g++) {
			uint64_t nkeys = 0; uint32_t maxb = 0;
			for (uint64_t bc = g << lb; bc < ((g + 1) << lb); bc++) { nkeys += bkm[bc]; maxb = std::max(maxb, F.bucket_nuc[bc]); }
			int nb = (maxb == 0 ? 0 : 32 - __builtin_clz(maxb)) - int(b);
			if (nb < 1) nb = 1;
			MphfRec& r = F.mphf[g];
			r.nbits = uint32_t(nb); r.pos_start = total_pos; r.nelem = nkeys; r.id_offset = id_base; r.present = nkeys ? 1 : 0;
			total_pos += uint64_t(nb) * nkeys + 8;
			id_base += nkeys;
			if (nkeys >= (1ull << 32)) { set_err("an MPHF group holds 2^32 or more k-mers; use a larger n"); return BL_ERR_INVALID_ARG; }
			GroupDev& G = groups[g];
			std::memset(&G, 0, sizeof G);
			G.bits_base = total_bits; G.pos_start = r.pos_start; G.nbits = r.nbits; G.block_base = total_blocks;
			gblock_first[g] = total_blocks;
			if (r.present) {
				level_domains_host(nkeys, r.dom);
				uint64_t off = 0;
				for (int l = 0; l < kLevels; l++) { G.dom[l] = r.dom[l]; G.level_off[l] = off; off += r.dom[l]; }
				G.n_words = off / 64;
				r.bits_word_off = total_bits / 64;
				r.bits_nwords = G.n_words;
				total_bits += off;
				total_blocks += (G.n_words + 15) / 16;
			}
		}
		gblock_first[H.n_mphf] = total_blocks;
		H.positions_bits = total_pos;
		H.pos_words = (total_pos + 63) / 64;
	}
	GroupDev* d_groups; uint64_t* d_gblock_first;
	AL(d_groups, H.n_mphf);
	AL(d_gblock_first, H.n_mphf + 1);
	CU(cudaMemcpy(d_groups, groups.data(), H.n_mphf * sizeof(GroupDev), cudaMemcpyHostToDevice));
	CU(cudaMemcpy(d_gblock_first, gblock_first.data(), (H.n_mphf + 1) * 8, cudaMemcpyHostToDevice));

	mark("group descriptors (host)");
	// 3. bucket sequences
	uint64_t* d_seq;
	A.hint(H.seq_words * 8 + N * 40 + total_bits / 4 + H.pos_words * 8 + total_blocks * 20 + H.n_mphf * 8 + 65536);
	AL(d_seq, H.seq_words + 1);
	if (H.seq_words) {
		k_seq_words<<<grid_for(H.seq_words), kT>>>(d_dest, n_sk, d_perm, d_sk_v, d_codes, H.total_nuc, H.seq_words, d_seq);
		g_launches++;
	}

	// keys
	uint64_t *d_keys, *d_final; uint32_t *d_kgroup, *d_koff;
	AL(d_keys, N); AL(d_final, N); AL(d_kgroup, N); AL(d_koff, N);
	CU(cudaMemset(d_final, 0xFF, N * 8));
	k_keys<<<gs, kT>>>(d_perm, d_sk_v, d_mini_sorted, d_nk_sorted, d_dest, d_keybase, d_bucket_start, d_codes, n_sk, k, lb, d_keys, d_kgroup, d_koff);
	g_launches++;
	CU(cudaGetLastError());

	mark("bucket text + keys");
	// 4. BBHash, all groups level by level
	const uint64_t n_words32 = total_bits / 32;
	uint32_t *d_bits, *d_coll, *d_gcoll, *d_flevel; uint64_t *d_la, *d_lb; unsigned long long* d_nnext;
	AL(d_bits, n_words32 + 2, true);
	AL(d_coll, n_words32 + 2, true);
	AL(d_gcoll, H.n_mphf, true);
	AL(d_flevel, H.n_mphf);
	CU(cudaMemset(d_flevel, 0xFF, H.n_mphf * 4));
	AL(d_nnext, 1, true);
	d_la = d_lb = nullptr;
	uint64_t n_items = N;
	const uint64_t* list = nullptr;
	std::vector<uint64_t> leftovers;
	for (int level = 0; level < kLevels && n_items; level++) {
		const unsigned gi = grid_for(n_items, kT * 2);
		k_level_place<<<gi, kT>>>(d_keys, d_kgroup, list, n_items, level, d_groups, d_bits, d_coll, d_gcoll);
		if (!d_la) AL(d_la, n_items);  // the first sift keeps at most every key
		if (list == d_la && !d_lb) AL(d_lb, n_items);  // survivors of level 1: at most what level 0 left
		uint64_t* next = (list == d_la) ? d_lb : d_la;
		CU(cudaMemset(d_nnext, 0, 8));
		k_level_sift<<<gi, kT>>>(d_keys, d_kgroup, list, n_items, level, d_groups, d_coll, next, d_nnext, d_final);
		k_level_clear<<<grid_for(n_words32, kT * 4), kT>>>(d_bits, d_coll, n_words32);
		k_level_groups<<<grid_for(H.n_mphf), kT>>>(d_gcoll, d_flevel, H.n_mphf, level);
		g_launches += 4;
		CU(cudaGetLastError());
		unsigned long long nn = 0;
		CU(cudaMemcpy(&nn, d_nnext, 8, cudaMemcpyDeviceToHost));
		list = next;
		n_items = nn;
	}
	if (n_items) {  // keys no level accommodated: the fallback map, on the host (bbhash.h:709-728)
		leftovers.resize(n_items);
		CU(cudaMemcpy(leftovers.data(), list, n_items * 8, cudaMemcpyDeviceToHost));
		std::sort(leftovers.begin(), leftovers.end());  // key order of the reference's iteration
	}

	mark("BBHash levels");
	// 5. ranks
	uint32_t* d_blockpop; uint64_t *d_bscan, *d_ranks;
	AL(d_blockpop, total_blocks + 1);
	AL(d_bscan, total_blocks + 1);
	AL(d_ranks, total_blocks + 1);
	const uint64_t* d_bits64 = reinterpret_cast<const uint64_t*>(d_bits);
	if (total_blocks) {
		k_block_pop<<<grid_for(total_blocks), kT>>>(d_groups, d_gblock_first, H.n_mphf, total_blocks, d_bits64, d_blockpop);
		g_launches++;
		if ((rc = exclusive_scan(A, S, d_blockpop, total_blocks, d_bscan, nullptr)) != BL_OK) { set_err(g_last_error); return rc; }
		k_ranks_out<<<grid_for(total_blocks), kT>>>(d_groups, d_gblock_first, H.n_mphf, total_blocks, d_bscan, d_ranks);
		g_launches++;
	}

	mark("ranks");
	// 6. positions
	unsigned long long* d_pos;
	AL(d_pos, H.pos_words + 1, true);
	k_positions<<<grid_for(N, kT * 2), kT>>>(d_kgroup, d_koff, d_final, N, d_groups, d_bits64, d_bscan, b, d_pos);
	g_launches++;
	CU(cudaGetLastError());
	CU(cudaEventRecord(ev1, 0));

	mark("positions");
	// export
	F.seq.assign(H.seq_words, 0);
	F.pos.assign(H.pos_words, 0);
	F.bits.assign(total_bits / 64, 0);
	std::vector<uint64_t> ranks_all(total_blocks), blockpop_last;
	std::vector<uint32_t> flevel(H.n_mphf);
	if (H.seq_words && (rc = X.d2h(F.seq.data(), d_seq, H.seq_words * 8)) != BL_OK) { set_err(g_last_error); return rc; }
	if ((rc = X.d2h(F.pos.data(), d_pos, H.pos_words * 8)) != BL_OK) { set_err(g_last_error); return rc; }
	if (total_bits && (rc = X.d2h(F.bits.data(), d_bits, total_bits / 8)) != BL_OK) { set_err(g_last_error); return rc; }
	if (total_blocks) CU(cudaMemcpy(ranks_all.data(), d_ranks, total_blocks * 8, cudaMemcpyDeviceToHost));
	CU(cudaMemcpy(flevel.data(), d_flevel, H.n_mphf * 4, cudaMemcpyDeviceToHost));
	if (seconds_device) { float ms = 0; cudaEventElapsedTime(&ms, ev0, ev1); *seconds_device = ms * 1e-3; }

	mark("export D2H");
	// leftovers by group (they are sorted by key index, groups are contiguous key ranges)
	std::vector<uint64_t> key_begin(H.n_mphf + 1, 0);
	for (uint64_t g = 0; g < H.n_mphf; g++) key_begin[g + 1] = key_begin[g] + F.mphf[g].nelem;
	size_t li = 0;
	for (uint64_t g = 0; g < H.n_mphf; g++) {
		MphfRec& R = F.mphf[g];
		if (!R.present) continue;
		const GroupDev& G = groups[g];
		const bool finished = flevel[g] != 0xFFFFFFFFu;
		const uint64_t upto = finished ? G.level_off[flevel[g]] + G.dom[flevel[g]] : G.n_words * 64;
		const uint64_t max_idx = (upto + 63) / 64;
		R.ranks_off = F.ranks.size();
		R.nranks = (max_idx + 15) / 16;
		F.ranks.insert(F.ranks.end(), ranks_all.begin() + G.block_base, ranks_all.begin() + G.block_base + R.nranks);
		R.fb_off = F.fb_keys.size();
		R.fb_count = 0;
		if (li < leftovers.size() && leftovers[li] < key_begin[g + 1]) {
			// ones in the whole bit array of the group = the rank the first leftover gets
			uint64_t cur = 0;
			for (uint64_t w = 0; w < G.n_words; w++) cur += (uint64_t)__builtin_popcountll(F.bits[R.bits_word_off + w]);
			// `fm[key] = cur++` in key order (bbhash.h:709-728): a k-mer present twice keeps the LAST rank handed to it, and every
			// occurrence then writes its offset to that rank's field, later writes winning (blight.cpp:486-519)
			std::unordered_map<uint64_t, uint64_t> fm;
			std::vector<std::pair<uint64_t, uint32_t>> seen;  // (key, offset in bucket) of every leftover, in order
			for (; li < leftovers.size() && leftovers[li] < key_begin[g + 1]; li++) {
				uint64_t key = 0; uint32_t off = 0;
				CU(cudaMemcpy(&key, d_keys + leftovers[li], 8, cudaMemcpyDeviceToHost));
				CU(cudaMemcpy(&off, d_koff + leftovers[li], 4, cudaMemcpyDeviceToHost));
				fm[key] = cur++;
				seen.emplace_back(key, off);
			}
			std::vector<std::pair<uint64_t, uint64_t>> fb(fm.begin(), fm.end());
			std::sort(fb.begin(), fb.end());
			for (auto& kv : fb) { F.fb_keys.push_back(kv.first); F.fb_vals.push_back(kv.second); }
			R.fb_count = fb.size();
			for (auto& ko : seen)
				if (ko.second) host_write_field(F.pos, R.pos_start + fm[ko.first] * R.nbits, R.nbits, ((uint64_t)ko.second >> b) & (R.nbits >= 64 ? ~0ull : ((1ull << R.nbits) - 1)));
		}
	}
	H.bits_words_total = F.bits.size();
	H.ranks_total = F.ranks.size();
	H.fallback_total = F.fb_keys.size();
#undef AL
	return flat_validate(F, err);
}

}  // namespace blight

using namespace blight;

extern "C" {

int blight_flat_build_gpu(const char* bases, const uint64_t* starts, const uint64_t* lengths, uint64_t n_seqs, uint32_t k, uint32_t m,
                          uint32_t n_log2, uint32_t s_log2, uint32_t b, int device, blight_flat** out, double* device_seconds) {
	if (!out || (n_seqs && (!bases || !starts || !lengths))) return fail(BL_ERR_INVALID_ARG, "null argument");
	BuildParams p; p.k = k; p.m = m; p.n_log2 = n_log2; p.s_log2 = s_log2; p.b = b;
	// the range of `bases` the sequences cover travels to the device once; the views are relative to it
	uint64_t lo = ~0ull, hi = 0;
	for (uint64_t i = 0; i < n_seqs; i++) { lo = std::min(lo, starts[i]); hi = std::max(hi, starts[i] + lengths[i]); }
	if (n_seqs == 0) lo = hi = 0;
	std::vector<uint64_t> st(n_seqs), ln(lengths, lengths + n_seqs);
	for (uint64_t i = 0; i < n_seqs; i++) st[i] = starts[i] - lo;
	blight_flat* f = new blight_flat();
	std::string err;
	int rc = build_flat_index_gpu(bases + lo, hi - lo, st, ln, p, device, f->f, &err, device_seconds);
	if (rc != BL_OK) { delete f; return fail(rc, err); }
	*out = f;
	return BL_OK;
}

int blight_flat_build_file_gpu(const char* unitig_path, uint32_t k, uint32_t m, uint32_t n_log2, uint32_t s_log2, uint32_t b, int device,
                               blight_flat** out) {
	if (!out || !unitig_path) return fail(BL_ERR_INVALID_ARG, "null argument");
	BuildParams p; p.k = k; p.m = m; p.n_log2 = n_log2; p.s_log2 = s_log2; p.b = b;
	std::string err;
	int rc = check_params(p, &err);
	if (rc != BL_OK) return fail(rc, err);
	std::string storage;
	std::vector<SeqView> seqs;
	rc = read_fasta_records(unitig_path, storage, seqs, &err);
	if (rc != BL_OK) return fail(rc, err);
	std::vector<uint64_t> st(seqs.size()), ln(seqs.size());
	for (size_t i = 0; i < seqs.size(); i++) { st[i] = uint64_t(seqs[i].p - storage.data()); ln[i] = seqs[i].len; }
	blight_flat* f = new blight_flat();
	rc = build_flat_index_gpu(storage.data(), storage.size(), st, ln, p, device, f->f, &err, nullptr);
	if (rc != BL_OK) { delete f; return fail(rc, err); }
	*out = f;
	return BL_OK;
}

}  // extern "C"
