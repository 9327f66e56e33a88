// tools/gather_bench.cu — measures the practical ceiling the lookup core runs against: random 32-byte-sector reads
// over an array larger than L2 (SURVEY.md §8d asks for this next to the sequential HBM peak).
//   independent : every thread issues ILP independent random sector loads per iteration
//   chained     : every thread walks CHAIN dependent random sector loads per item (like MPHF -> position -> sequence)
// Usage: gather_bench [array MiB = 320] [items per launch = 2^28]
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
	x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
	return x;
}

__device__ __forceinline__ uint32_t ld_sector_sum(const uint32_t* p) {
	uint32_t a, b, c, d, e, f, g, h;
	asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(a), "=r"(b), "=r"(c), "=r"(d), "=r"(e), "=r"(f), "=r"(g), "=r"(h) : "l"(p));
	return a ^ b ^ c ^ d ^ e ^ f ^ g ^ h;
}

template <int ILP>
__global__ void __launch_bounds__(256) k_indep(const uint32_t* __restrict__ arr, uint64_t n_sectors, uint64_t items, uint32_t* out) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * ILP;
	uint32_t acc = 0;
	for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * ILP; i < items; i += stride) {
		uint32_t v[ILP];
		#pragma unroll
		for (int j = 0; j < ILP; j++) v[j] = ld_sector_sum(arr + (mix(i + j) % n_sectors) * 8);
		#pragma unroll
		for (int j = 0; j < ILP; j++) acc ^= v[j];
	}
	if (acc == 0x12345678u) out[0] = acc;
}

// WIDE consecutive sectors per random access (64 B / 128 B granules): does the memory system charge per sector or per granule?
template <int WIDE>
__global__ void __launch_bounds__(256) k_wide(const uint32_t* __restrict__ arr, uint64_t n_sectors, uint64_t items, uint32_t* out) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	uint32_t acc = 0;
	const uint64_t n_gran = n_sectors / WIDE;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += stride) {
		const uint32_t* p = arr + (mix(i) % n_gran) * 8 * WIDE;
		uint32_t v[WIDE];
		#pragma unroll
		for (int j = 0; j < WIDE; j++) v[j] = ld_sector_sum(p + 8 * j);
		#pragma unroll
		for (int j = 0; j < WIDE; j++) acc ^= v[j];
	}
	if (acc == 0x12345678u) out[0] = acc;
}

template <int CHAIN>
__global__ void __launch_bounds__(256) k_chain(const uint32_t* __restrict__ arr, uint64_t n_sectors, uint64_t items, uint32_t* out) {
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
	uint32_t acc = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < items; i += stride) {
		uint64_t x = i;
		#pragma unroll
		for (int c = 0; c < CHAIN; c++) {
			const uint32_t v = ld_sector_sum(arr + (mix(x) % n_sectors) * 8);
			x = x * 0x9E3779B97F4A7C15ull + v + c;
		}
		acc ^= (uint32_t)x;
	}
	if (acc == 0x12345678u) out[0] = acc;
}

template <class F>
float time_ms(F f, int reps) {
	cudaEvent_t a, b;
	cudaEventCreate(&a); cudaEventCreate(&b);
	f();
	cudaDeviceSynchronize();
	cudaEventRecord(a);
	for (int i = 0; i < reps; i++) f();
	cudaEventRecord(b);
	cudaEventSynchronize(b);
	float ms = 0;
	cudaEventElapsedTime(&ms, a, b);
	return ms / reps;
}

int main(int argc, char** argv) {
	const uint64_t mib = argc > 1 ? strtoull(argv[1], nullptr, 10) : 320;
	const uint64_t items = argc > 2 ? strtoull(argv[2], nullptr, 10) : (1ull << 28);
	const uint64_t n_sectors = mib * 1024 * 1024 / 32;
	const int gran = argc > 3 ? atoi(argv[3]) : 0;  // cudaLimitMaxL2FetchGranularity (0 = leave the default)
	if (gran) {
		cudaError_t ge = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)gran);
		if (ge != cudaSuccess) printf("setlimit failed: %s\n", cudaGetErrorString(ge));
	}
	size_t gran_now = 0;
	cudaDeviceGetLimit(&gran_now, cudaLimitMaxL2FetchGranularity);
	uint32_t *arr, *out;
	if (cudaMalloc(&arr, n_sectors * 32) != cudaSuccess || cudaMalloc(&out, 4) != cudaSuccess) { printf("alloc failed\n"); return 1; }
	cudaMemset(arr, 1, n_sectors * 32);
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	printf("{\"array_mib\": %llu, \"items\": %llu, \"sms\": %d, \"l2_fetch_granularity\": %zu", (unsigned long long)mib, (unsigned long long)items, sms, gran_now);
	for (int occ : {8}) {
		const int grid = sms * occ;
		float ms;
		ms = time_ms([&] { k_indep<1><<<grid, 256>>>(arr, n_sectors, items, out); }, 3);
		printf(", \"indep1_occ%d_Gsect_s\": %.2f", occ, items / ms / 1e6);
		ms = time_ms([&] { k_indep<4><<<grid, 256>>>(arr, n_sectors, items, out); }, 3);
		printf(", \"indep4_occ%d_Gsect_s\": %.2f", occ, items / ms / 1e6);
		ms = time_ms([&] { k_indep<8><<<grid, 256>>>(arr, n_sectors, items, out); }, 3);
		printf(", \"indep8_occ%d_Gsect_s\": %.2f", occ, items / ms / 1e6);
		ms = time_ms([&] { k_wide<2><<<grid, 256>>>(arr, n_sectors, items / 2, out); }, 3);
		printf(", \"wide64B_occ%d_Gaccess_s\": %.2f", occ, (items / 2) / ms / 1e6);
		ms = time_ms([&] { k_wide<4><<<grid, 256>>>(arr, n_sectors, items / 4, out); }, 3);
		printf(", \"wide128B_occ%d_Gaccess_s\": %.2f", occ, (items / 4) / ms / 1e6);
		ms = time_ms([&] { k_chain<3><<<grid, 256>>>(arr, n_sectors, items / 4, out); }, 3);
		printf(", \"chain3_occ%d_Gsect_s\": %.2f", occ, 3.0 * (items / 4) / ms / 1e6);
	}
	printf("}\n");
	cudaError_t e = cudaDeviceSynchronize();
	if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 1; }
	return 0;
}
