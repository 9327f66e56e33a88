// dist_kernels.cu — kernels either side of the NCCL all-to-all of the bucket-partitioned multi-GPU mode
// (SURVEY.md §8e): bin (canon, minimizer) pairs by the rank that owns their MPHF group, and scatter the ids that
// come back into query order.  Owner of a k-mer = the rank whose contiguous group range holds
// minimizer >> lb (the same `minimizer / number_bucket_per_mphf` the reference uses to pick the MPHF, blight.cpp:722).
#include <cuda_runtime.h>

#include "capi_common.hpp"
#include "kernels.hpp"

namespace blight {
namespace {

constexpr int kMaxRanks = 64;

__device__ __forceinline__ uint32_t owner_of(uint32_t mini, uint32_t lb, const uint32_t* cuts, uint32_t world) {
	const uint32_t g = mini >> lb;
	uint32_t o = 0;
	while (o + 1 < world && g >= cuts[o + 1]) o++;  // world <= 8 in practice; cuts live in shared memory
	return o;
}

__global__ void __launch_bounds__(256) k_owner_count(const uint32_t* __restrict__ mini, uint64_t n, const uint32_t* __restrict__ group_cuts,
                                                     uint32_t world, uint32_t lb, unsigned long long* __restrict__ counts) {
	__shared__ uint32_t cuts[kMaxRanks + 1];
	__shared__ unsigned long long local[kMaxRanks];
	if (threadIdx.x <= world) cuts[threadIdx.x] = group_cuts[threadIdx.x];
	if (threadIdx.x < world) local[threadIdx.x] = 0;
	__syncthreads();
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const uint32_t o = owner_of(__ldcs(mini + i), lb, cuts, world);
		const uint32_t peers = __match_any_sync(__activemask(), o);
		if ((threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&local[o], (unsigned long long)__popc(peers));
	}
	__syncthreads();
	if (threadIdx.x < world && local[threadIdx.x]) atomicAdd(&counts[threadIdx.x], local[threadIdx.x]);
}

// cursors[o] must hold the exclusive prefix of counts when this starts; it ends at the inclusive prefix.
__global__ void __launch_bounds__(256) k_owner_scatter(const uint64_t* __restrict__ canon, const uint32_t* __restrict__ mini, uint64_t n,
                                                       const uint32_t* __restrict__ group_cuts, uint32_t world, uint32_t lb,
                                                       unsigned long long* __restrict__ cursors, uint64_t* __restrict__ send_canon,
                                                       uint32_t* __restrict__ send_mini, uint64_t* __restrict__ send_src) {
	__shared__ uint32_t cuts[kMaxRanks + 1];
	if (threadIdx.x <= world) cuts[threadIdx.x] = group_cuts[threadIdx.x];
	__syncthreads();
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	const uint32_t lane = threadIdx.x & 31;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
		const uint32_t mn = __ldcs(mini + i);
		const uint32_t o = owner_of(mn, lb, cuts, world);
		// warp-aggregated reservation: one atomic per owner present in the warp, lanes keep their relative order
		const uint32_t peers = __match_any_sync(__activemask(), o);
		const uint32_t leader = __ffs(peers) - 1;
		unsigned long long base = 0;
		if (lane == leader) base = atomicAdd(&cursors[o], (unsigned long long)__popc(peers));
		base = __shfl_sync(peers, base, leader);
		const uint64_t dst = base + __popc(peers & ((1u << lane) - 1u));
		send_canon[dst] = __ldcs(canon + i);
		send_mini[dst] = mn;
		send_src[dst] = i;
	}
}

__global__ void __launch_bounds__(256) k_scatter_ids(const int64_t* __restrict__ ids_back, const uint64_t* __restrict__ src, uint64_t n,
                                                     int64_t* __restrict__ out) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[__ldcs(src + i)] = __ldcs(ids_back + i);
}

unsigned grid_for(uint64_t n) {
	int dev = 0, sms = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
	const uint64_t want = (n + 255) / 256, cap = (uint64_t)sms * 8;
	return (unsigned)(want < cap ? (want ? want : 1) : cap);
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

int blight_owner_count(const uint32_t* d_mini, uint64_t n, const uint32_t* d_group_cuts, uint32_t world, uint32_t lb,
                       uint64_t* d_counts, void* stream) {
	if (world == 0 || world > (uint32_t)kMaxRanks || !d_group_cuts || !d_counts || (n && !d_mini)) return fail(BL_ERR_INVALID_ARG, "bad argument");
	if (n == 0) return BL_OK;
	k_owner_count<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_mini, n, d_group_cuts, world, lb,
	                                                                        reinterpret_cast<unsigned long long*>(d_counts));
	g_launches++;
	return finish("k_owner_count");
}

int blight_owner_scatter(const uint64_t* d_canon, const uint32_t* d_mini, uint64_t n, const uint32_t* d_group_cuts,
                         uint32_t world, uint32_t lb, uint64_t* d_cursors, uint64_t* d_send_canon, uint32_t* d_send_mini,
                         uint64_t* d_send_src, void* stream) {
	if (world == 0 || world > (uint32_t)kMaxRanks || !d_group_cuts || !d_cursors) return fail(BL_ERR_INVALID_ARG, "bad argument");
	if (n == 0) return BL_OK;
	if (!d_canon || !d_mini || !d_send_canon || !d_send_mini || !d_send_src) return fail(BL_ERR_INVALID_ARG, "null argument");
	k_owner_scatter<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_canon, d_mini, n, d_group_cuts, world, lb,
	                                                                          reinterpret_cast<unsigned long long*>(d_cursors),
	                                                                          d_send_canon, d_send_mini, d_send_src);
	g_launches++;
	return finish("k_owner_scatter");
}

int blight_scatter_ids(const int64_t* d_ids_back, const uint64_t* d_src, uint64_t n, int64_t* d_out, void* stream) {
	if (n == 0) return BL_OK;
	if (!d_ids_back || !d_src || !d_out) return fail(BL_ERR_INVALID_ARG, "null argument");
	k_scatter_ids<<<grid_for(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(d_ids_back, d_src, n, d_out);
	g_launches++;
	return finish("k_scatter_ids");
}

}  // extern "C"
