// capi_device.cu — the device half of the C ABI (include/blight_b200.h): re-layout + upload of the flat index,
// device-buffer query entry points, and the host-buffer (end to end) entry points.
#include <cuda_runtime.h>
#include <omp.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "capi_common.hpp"
#include "device_index.hpp"
#include "kernels.hpp"
#include "kmer_math.hpp"

using namespace blight;

namespace {

int cuda_fail(cudaError_t e, const char* what) {
	return fail(BL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU(call)                                          \
	do {                                                  \
		cudaError_t e__ = (call);                         \
		if (e__ != cudaSuccess) return cuda_fail(e__, #call); \
	} while (0)

struct DeviceGuard {
	int prev = -1;
	explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
	~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <class T>
int upload(const std::vector<T>& v, void** d, uint64_t* bytes_acc) {
	const size_t bytes = std::max<size_t>(v.size() * sizeof(T), 256);
	CU(cudaMalloc(d, bytes));
	CU(cudaMemset(*d, 0, bytes));
	if (!v.empty()) CU(cudaMemcpy(*d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
	*bytes_acc += bytes;
	return BL_OK;
}

int ensure_device(int device) {
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) return fail(BL_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU fallback)");
	if (device < 0 || device >= n) return fail(BL_ERR_INVALID_ARG, "device ordinal out of range");
	return BL_OK;
}

}  // namespace

extern "C" {

int blight_index_upload(const blight_flat* ff, int device, blight_index** out) {
	return blight_index_upload_opts(ff, device, nullptr, out);
}

int blight_index_upload_opts(const blight_flat* ff, int device, const blight_upload_options* opts, blight_index** out) {
	if (!ff || !out) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (opts && opts->struct_size != sizeof(blight_upload_options)) return fail(BL_ERR_INVALID_ARG, "blight_upload_options.struct_size");
	// -1 = library default; the environment knobs of DESIGN.md (tests, experiments) only replace defaults
	auto knob = [](int32_t asked, const char* env, int32_t dflt) {
		if (asked >= 0) return asked;
		const char* e = getenv(env);
		return e ? (int32_t)atoi(e) : dflt;
	};
	const int32_t o_pos_id = knob(opts ? opts->pos_id : -1, "BLIGHT_POS_ID", 1);
	const int32_t o_filter_bits = knob(opts ? opts->filter_bits : -1, "BLIGHT_FILTER_BITS", 20);
	const int32_t o_exact = knob(opts ? opts->exact_pos : -1, "BLIGHT_EXACT_POS", 1);
	const int32_t o_anchors = knob(opts ? opts->filter_anchors : -1, "BLIGHT_FILTER_ANCHORS", 1);
	int rc = ensure_device(device);
	if (rc != BL_OK) return rc;
	const FlatIndex& F = ff->f;
	const FlatHeader& H = F.h;
	std::string err;
	rc = flat_validate(F, &err);
	if (rc != BL_OK) return fail(rc, err);
	for (const MphfRec& r : F.mphf)
		if (r.nelem >= (1ull << 32)) return fail(BL_ERR_INVALID_ARG, "an MPHF group holds 2^32 or more k-mers; use a larger n");

	// ---- re-layout on the host ----
	std::vector<uint4> bucket(H.n_buckets);
	for (uint64_t i = 0; i < H.n_buckets; i++)
		bucket[i] = make_uint4((uint32_t)F.bucket_start[i], (uint32_t)(F.bucket_start[i] >> 32), F.bucket_nuc[i], 0u);

	std::vector<DevMphf> mphf(H.n_mphf);
	uint64_t bits_sectors = 0, pos_sectors = 0;
	uint32_t small = 1;
	// exact-position layout (device_index.hpp): fields b bits wider, low bits filled in by the upload pass
	uint64_t n_keys = 0, max_id = 0, min_id = ~0ull;
	for (const MphfRec& r : F.mphf)
		if (r.present) { n_keys += r.nelem; max_id = std::max<uint64_t>(max_id, r.id_offset + r.nelem); min_id = std::min<uint64_t>(min_id, r.id_offset); }
	if (n_keys == 0) min_id = max_id = 0;
	const uint64_t n_local = max_id - min_id;  // identifiers of this index (or slice) span [min_id, max_id)
	const bool lid_fits = n_local < 0xFFFFFFFFull && H.total_nuc > 0;
	bool exact = o_exact != 0 && H.b > 0 && H.b <= 8 && lid_fits;
	for (const MphfRec& r : F.mphf) if (r.present && (r.nbits ? r.nbits : 1) + H.b > 32) exact = false;
	const uint32_t xb = exact ? H.b : 0;
	for (uint64_t g = 0; g < H.n_mphf; g++) {
		const MphfRec& r = F.mphf[g];
		DevMphf& d = mphf[g];
		std::memset(&d, 0, sizeof d);
		d.bits_sector_base = bits_sectors;
		d.pos_sector_base = pos_sectors;
		d.id_offset = r.id_offset;
		d.fb_off = r.fb_off;
		d.fb_count = (uint32_t)r.fb_count;
		d.nbits = (r.nbits ? r.nbits : 1) + xb;
		d.fields_per_sector = 256 / d.nbits;
		d.fps_magic = (uint32_t)((1ull << 32) / d.fields_per_sector);
		d.present = r.present;
		for (int l = 0; l < kLevels; l++) { d.dom[l] = r.present ? r.dom[l] : 64; d.dom32[l] = (uint32_t)d.dom[l]; }
		if (r.present && r.bits_nwords * 64 >= (1ull << 32)) small = 0;
		if (r.present) {
			bits_sectors += (r.bits_nwords * 64 + kChunkBits - 1) / kChunkBits;
			pos_sectors += (r.nelem + d.fields_per_sector - 1) / d.fields_per_sector;
		}
	}
	std::vector<uint32_t> bits((bits_sectors + 1) * 8, 0), pos((pos_sectors + 1) * 8, 0);
	// launchers such as torchrun export OMP_NUM_THREADS=1; the re-layout of a large index should not crawl because of it
	const int relayout_threads = std::max(omp_get_max_threads(), std::min(8, (int)std::thread::hardware_concurrency()));
	#pragma omp parallel for schedule(dynamic, 1) num_threads(relayout_threads)
	for (uint64_t g = 0; g < H.n_mphf; g++) {
		const MphfRec& r = F.mphf[g];
		if (!r.present) continue;
		const DevMphf& d = mphf[g];
		// level bits: 224-bit chunks + running popcount
		const uint32_t* src = reinterpret_cast<const uint32_t*>(F.bits.data() + r.bits_word_off);
		const uint64_t n32 = r.bits_nwords * 2;
		const uint64_t chunks = (r.bits_nwords * 64 + kChunkBits - 1) / kChunkBits;
		uint32_t ones = 0;
		uint32_t* dst = bits.data() + d.bits_sector_base * 8;
		for (uint64_t c = 0; c < chunks; c++, dst += 8) {
			dst[7] = ones;
			for (int j = 0; j < 7; j++) {
				const uint64_t si = c * 7 + j;
				const uint32_t wv = si < n32 ? src[si] : 0u;
				dst[j] = wv;
				ones += (uint32_t)popc32(wv);
			}
		}
		// positions: fields_per_sector fields per 32-byte sector, LSB-first inside the sector
		uint32_t* pd = pos.data() + d.pos_sector_base * 8;
		const uint32_t nb = d.nbits, fps = d.fields_per_sector;
		const uint32_t nb_src = r.nbits ? r.nbits : 1;  // width in the flat image
		for (uint64_t rk = 0; rk < r.nelem; rk++) {
			const uint64_t bitpos = r.pos_start + rk * nb_src;
			const uint64_t w0 = F.pos[bitpos >> 6];
			const uint64_t w1 = ((bitpos >> 6) + 1 < F.pos.size()) ? F.pos[(bitpos >> 6) + 1] : 0;
			const unsigned sh = unsigned(bitpos & 63);
			uint64_t v = sh ? ((w0 >> sh) | (w1 << (64 - sh))) : w0;
			v &= (nb_src >= 64) ? ~0ull : ((1ull << nb_src) - 1);
			v <<= xb;  // exact layout: the low b bits start at zero
			const uint64_t sec = rk / fps;
			const uint32_t o = uint32_t(rk % fps) * nb;
			uint32_t* sp = pd + sec * 8;
			sp[o >> 5] |= uint32_t(v << (o & 31));
			if ((o & 31) + nb > 32) sp[(o >> 5) + 1] |= uint32_t(v >> (32 - (o & 31)));
		}
	}
	// sequences: bit reversal of each 32-bit group turns the vector<bool> image (nucleotide p at bits 2p,2p+1 with the
	// code's high bit first) into 16 codes per word, first base in the high bits
	const uint64_t seq_pad = ((1ull << H.b) + H.k) / 16 + 8;
	std::vector<uint32_t> seq(H.seq_words * 2 + seq_pad, 0);
	{
		const uint32_t* src = reinterpret_cast<const uint32_t*>(F.seq.data());
		const int64_t n32 = (int64_t)H.seq_words * 2;
		#pragma omp parallel for schedule(static) num_threads(relayout_threads)
		for (int64_t i = 0; i < n32; i++) seq[i] = bitrev32(src[i]);
	}

	// ---- upload ----
	DeviceGuard guard(device);
	blight_index* idx = new blight_index();
	idx->device = device;
	uint64_t bytes = 0;
	rc = upload(bucket, &idx->d_bucket, &bytes);
	if (rc == BL_OK) rc = upload(mphf, &idx->d_mphf, &bytes);
	if (rc == BL_OK) rc = upload(bits, &idx->d_bits, &bytes);
	if (rc == BL_OK) rc = upload(pos, &idx->d_pos, &bytes);
	if (rc == BL_OK) rc = upload(seq, &idx->d_seq, &bytes);
	if (rc == BL_OK) rc = upload(F.fb_keys, &idx->d_fbk, &bytes);
	if (rc == BL_OK) rc = upload(F.fb_vals, &idx->d_fbv, &bytes);
	if (rc != BL_OK) { blight_index_free(idx); return rc; }
	cudaStream_t st;
	cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
	if (e != cudaSuccess) { blight_index_free(idx); return cuda_fail(e, "cudaStreamCreate"); }
	idx->host_stream = st;
	cudaStream_t cs;
	cudaEvent_t e1, e2;
	if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&e1, cudaEventDisableTiming) != cudaSuccess ||
	    cudaEventCreateWithFlags(&e2, cudaEventDisableTiming) != cudaSuccess) {
		blight_index_free(idx);
		return fail(BL_ERR_CUDA, "cannot create the copy stream of the host entry points");
	}
	idx->copy_stream = cs; idx->ev_copy = e1; idx->ev_ws = e2;
	idx->host_mutex = new std::mutex();
	DevIndexView& v = idx->v;
	v.bucket = static_cast<const uint4*>(idx->d_bucket);
	v.mphf = static_cast<const DevMphf*>(idx->d_mphf);
	v.bits = static_cast<const uint32_t*>(idx->d_bits);
	v.pos = static_cast<const uint32_t*>(idx->d_pos);
	v.seq = static_cast<const uint32_t*>(idx->d_seq);
	v.fb_keys = static_cast<const uint64_t*>(idx->d_fbk);
	v.fb_vals = static_cast<const uint64_t*>(idx->d_fbv);
	v.k = H.k; v.m = H.m; v.b = H.b; v.lb = F.lb();
	v.kmask = (1ull << (2 * H.k)) - 1;
	if (const char* e = getenv("BLIGHT_FORCE_WIDE")) { if (atoi(e)) small = 0; }  // test knob: the 64-bit bit arithmetic of huge MPHF groups
	v.small = small;
	if (exact) v.flags |= kFlagExactPos;
	{
		// Derived tables (device_index.hpp): per-position "answered found" bitmap, per-position identifier table, negative filter,
		// exact positions — the lookup core run once over every window of the index text. Only the bitmap is required: when
		// HBM is short the optional tables are left out (blight_info.layout says what is there) and the kernels take the
		// paths that do without them.
		const size_t vbytes = ((size_t)(H.total_nuc + 31) / 32 + 1) * 4;
		cudaError_t ve = cudaMalloc(&idx->d_valid, vbytes);
		if (ve == cudaSuccess) ve = cudaMemset(idx->d_valid, 0, vbytes);
		if (ve != cudaSuccess) { blight_index_free(idx); return cuda_fail(ve, "cudaMalloc(valid bitmap)"); }
		v.id_base = min_id;  // the per-position table holds id - id_base (32 bits)
		const bool want_pid = o_pos_id != 0 && lid_fits;
		size_t pbytes = 0, fbytes = 0;
		uint32_t fblocks = 0;
		auto soft_fail = [](cudaError_t e) { if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return true; } return false; };
		if (want_pid) {
			pbytes = ((size_t)H.total_nuc + 32) * 4;
			ve = cudaMalloc(&idx->d_pos_id, pbytes);
			if (ve != cudaSuccess) {
				if (!soft_fail(ve)) { blight_index_free(idx); return cuda_fail(ve, "cudaMalloc(position -> id table)"); }
				idx->d_pos_id = nullptr; pbytes = 0;
			}
		}
		if (o_filter_bits > 0 && n_keys) {
			const uint64_t nb = std::min<uint64_t>((n_keys * (uint64_t)o_filter_bits + 255) / 256 + 1, 0xFFFFFFFFull);
			fblocks = (uint32_t)nb;
			fbytes = (size_t)nb * 32;
			ve = cudaMalloc(&idx->d_filter, fbytes);
			if (ve == cudaSuccess) ve = cudaMemset(idx->d_filter, 0, fbytes);
			if (ve != cudaSuccess) {
				if (!soft_fail(ve)) { blight_index_free(idx); return cuda_fail(ve, "cudaMalloc(filter)"); }
				cudaFree(idx->d_filter); idx->d_filter = nullptr; fbytes = 0; fblocks = 0;
			}
		}
		// scratch of the exact-position passes: candidate bitmap, one claim word per key, and (without the identifier table)
		// a per-position identifier array that lives only during the upload
		void *d_cand = nullptr, *d_claim = nullptr, *d_lid_tmp = nullptr;
		if (exact) {
			bool ok = cudaMalloc(&d_cand, vbytes) == cudaSuccess && cudaMalloc(&d_claim, (size_t)(n_local + 1) * 4) == cudaSuccess;
			if (ok && !idx->d_pos_id) ok = cudaMalloc(&d_lid_tmp, ((size_t)H.total_nuc + 32) * 4) == cudaSuccess;
			if (ok) ok = cudaMemset(d_cand, 0, vbytes) == cudaSuccess && cudaMemset(d_claim, 0xFF, (size_t)(n_local + 1) * 4) == cudaSuccess;
			if (!ok) {
				// no room for the scratch: keep the (already widened) fields with zero low bits — every lookup then starts its scan at
				// the reference's own truncated position, which is exact as well
				cudaGetLastError();
				cudaFree(d_cand); cudaFree(d_claim); cudaFree(d_lid_tmp);
				d_cand = d_claim = d_lid_tmp = nullptr;
			}
		}
		uint32_t* d_lid = idx->d_pos_id ? static_cast<uint32_t*>(idx->d_pos_id) : static_cast<uint32_t*>(d_lid_tmp);
		int vrc = launch_window_answers(v, H.n_buckets, H.total_nuc, static_cast<uint32_t*>(idx->d_valid), d_lid, static_cast<uint32_t*>(d_cand),
		                                static_cast<uint32_t*>(idx->d_filter), fblocks, nullptr);
		if (vrc == BL_OK && exact && d_cand)
			vrc = launch_exact_positions(v, H.n_buckets, H.n_mphf, H.total_nuc, static_cast<const uint32_t*>(d_cand), d_lid,
			                             static_cast<uint32_t*>(d_claim), static_cast<uint32_t*>(idx->d_pos), nullptr);
		ve = cudaDeviceSynchronize();
		cudaFree(d_cand); cudaFree(d_claim); cudaFree(d_lid_tmp);
		if (vrc != BL_OK || ve != cudaSuccess) {
			blight_index_free(idx);
			return fail(BL_ERR_CUDA, std::string("upload passes failed: ") + (ve != cudaSuccess ? cudaGetErrorString(ve) : g_last_cuda_error));
		}
		v.pos_id = static_cast<const uint32_t*>(idx->d_pos_id);
		v.filter = static_cast<const uint32_t*>(idx->d_filter);
		v.filter_blocks = fblocks;
		if (o_anchors) v.flags |= kFlagFilterAnchors;  // anchors through the filter too: measured 51.1 vs 52.0 ms (counting), 56.8 vs 57.9 ms (ids)
		bytes += pbytes + fbytes;
		v.valid = static_cast<const uint32_t*>(idx->d_valid);
		bytes += vbytes;
	}
	fill_info(F, &idx->info);
	idx->info.device_bytes = bytes;
	idx->info.layout = (idx->v.pos_id ? BLIGHT_LAYOUT_POS_ID : 0u) | (idx->v.filter ? BLIGHT_LAYOUT_FILTER : 0u) |
	                   ((idx->v.flags & kFlagExactPos) ? BLIGHT_LAYOUT_EXACT_POS : 0u);
	*out = idx;
	return BL_OK;
}

void blight_index_free(blight_index* idx) {
	if (!idx) return;
	DeviceGuard guard(idx->device);
	cudaFree(idx->d_bucket); cudaFree(idx->d_mphf); cudaFree(idx->d_bits); cudaFree(idx->d_pos); cudaFree(idx->d_seq);
	cudaFree(idx->d_fbk); cudaFree(idx->d_fbv); cudaFree(idx->d_valid); cudaFree(idx->d_pos_id); cudaFree(idx->d_filter);
	for (void* w : idx->ws) cudaFree(w);
	if (idx->stream_ctx) stream_ctx_free(idx->stream_ctx);
	if (idx->host_pool) host_pool_free(idx->host_pool);
	if (idx->host_stream) cudaStreamDestroy(static_cast<cudaStream_t>(idx->host_stream));
	if (idx->copy_stream) cudaStreamDestroy(static_cast<cudaStream_t>(idx->copy_stream));
	if (idx->ev_copy) cudaEventDestroy(static_cast<cudaEvent_t>(idx->ev_copy));
	if (idx->ev_ws) cudaEventDestroy(static_cast<cudaEvent_t>(idx->ev_ws));
	delete static_cast<std::mutex*>(idx->host_mutex);
	delete idx;
}

int blight_index_info(const blight_index* idx, blight_info* out) {
	if (!idx || !out) return fail(BL_ERR_INVALID_ARG, "null argument");
	*out = idx->info;
	return BL_OK;
}

// ---- device-buffer entry points ---------------------------------------------------------------------------

int blight_query_kmers(const blight_index* idx, const uint64_t* d_canon, uint64_t n, int64_t* d_ids, void* stream) {
	if (!idx || (n && (!d_canon || !d_ids))) return fail(BL_ERR_INVALID_ARG, "null argument");
	DeviceGuard guard(idx->device);
	int rc = launch_lookup_kmers(idx->v, d_canon, nullptr, n, d_ids, static_cast<cudaStream_t>(stream));
	return rc == BL_OK ? rc : fail(rc, std::string("kernel launch failed: ") + g_last_cuda_error);
}

int blight_query_kmers_mini(const blight_index* idx, const uint64_t* d_canon, const uint32_t* d_mini, uint64_t n,
                            int64_t* d_ids, void* stream) {
	if (!idx || (n && (!d_canon || !d_mini || !d_ids))) return fail(BL_ERR_INVALID_ARG, "null argument");
	DeviceGuard guard(idx->device);
	int rc = launch_lookup_kmers(idx->v, d_canon, d_mini, n, d_ids, static_cast<cudaStream_t>(stream));
	return rc == BL_OK ? rc : fail(rc, std::string("kernel launch failed: ") + g_last_cuda_error);
}

int blight_reads_to_kmers(uint32_t k, uint32_t m, const char* d_bases, const uint64_t* d_read_off,
                          const uint64_t* d_kmer_off, uint64_t n_reads, uint64_t total_bases, uint64_t* d_canon,
                          uint32_t* d_mini, uint64_t* d_ctr, void* stream) {
	if (n_reads && (!d_bases || !d_read_off || !d_kmer_off || !d_canon || !d_mini || !d_ctr)) return fail(BL_ERR_INVALID_ARG, "null argument");
	BuildParams p; p.k = k; p.m = m; p.n_log2 = 0; p.s_log2 = 0; p.b = 0;
	std::string err;
	int rc = check_params(p, &err);
	if (rc != BL_OK) return fail(rc, err);
	ReadBatch B;
	B.d_bases = d_bases; B.d_read_off = d_read_off; B.d_kmer_off = d_kmer_off; B.n_reads = n_reads; B.total_bases = total_bases;
	rc = launch_reads(nullptr, k, m, B, d_canon, d_mini, nullptr, d_ctr, static_cast<cudaStream_t>(stream));
	return rc == BL_OK ? rc : fail(rc, std::string("kernel launch failed: ") + g_last_cuda_error);
}

int blight_query_reads(const blight_index* idx, const char* d_bases, const uint64_t* d_read_off,
                       const uint64_t* d_kmer_off, uint64_t n_reads, uint64_t total_bases, uint64_t total_kmers,
                       int64_t* d_ids, uint64_t* d_ctr, void* stream) {
	(void)total_kmers;
	if (!idx || (n_reads && (!d_bases || !d_read_off || !d_ctr || (d_ids && !d_kmer_off)))) return fail(BL_ERR_INVALID_ARG, "null argument");
	DeviceGuard guard(idx->device);
	ReadBatch B;
	B.d_bases = d_bases; B.d_read_off = d_read_off; B.d_kmer_off = d_kmer_off; B.n_reads = n_reads; B.total_bases = total_bases;
	int rc = launch_reads(&idx->v, idx->v.k, idx->v.m, B, nullptr, nullptr, d_ids, d_ctr, static_cast<cudaStream_t>(stream));
	return rc == BL_OK ? rc : fail(rc, std::string("kernel launch failed: ") + g_last_cuda_error);
}

int blight_query_reads_packed(const blight_index* idx, const uint32_t* d_packed, const uint64_t* d_read_off, const uint64_t* d_kmer_off,
                              uint64_t n_reads, uint64_t total_bases, int64_t* d_ids, uint64_t* d_ctr, void* stream) {
	if (!idx || (n_reads && (!d_packed || !d_read_off || !d_ctr || (d_ids && !d_kmer_off)))) return fail(BL_ERR_INVALID_ARG, "null argument");
	DeviceGuard guard(idx->device);
	ReadBatch B;
	B.d_packed = d_packed; B.d_read_off = d_read_off; B.d_kmer_off = d_kmer_off; B.n_reads = n_reads; B.total_bases = total_bases;
	int rc = launch_reads(&idx->v, idx->v.k, idx->v.m, B, nullptr, nullptr, d_ids, d_ctr, static_cast<cudaStream_t>(stream));
	return rc == BL_OK ? rc : fail(rc, std::string("kernel launch failed: ") + g_last_cuda_error);
}

int blight_consume_reads(const blight_index* idx, const char* d_bases, const uint64_t* d_read_off, uint64_t n_reads, uint64_t total_bases,
                         int kind, uint32_t* d_table, uint32_t n_colors, uint32_t color, uint64_t* d_ctr, void* stream) {
	if (!idx || !d_table || !d_ctr || (n_reads && (!d_bases || !d_read_off))) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (kind != BLIGHT_CONSUME_COUNT && kind != BLIGHT_CONSUME_COLOR) return fail(BL_ERR_INVALID_ARG, "unknown consumer");
	if (kind == BLIGHT_CONSUME_COLOR && (n_colors == 0 || color >= n_colors)) return fail(BL_ERR_INVALID_ARG, "color out of range");
	if (!idx->v.pos_id || idx->v.k - idx->v.m + 1 < 8) return fail(BL_ERR_INVALID_ARG, "the fused consumers need the position->id table (N < 2^32-1, BLIGHT_POS_ID) and k-m+1 >= 8");
	DeviceGuard guard(idx->device);
	ReadBatch B;
	B.d_bases = d_bases; B.d_read_off = d_read_off; B.n_reads = n_reads; B.total_bases = total_bases;
	int rc = launch_reads_sink(idx->v, kind, B, d_table, n_colors, color, nullptr, d_ctr, static_cast<cudaStream_t>(stream));
	return rc == BL_OK ? rc : fail(rc, std::string("kernel launch failed: ") + g_last_cuda_error);
}

int blight_gather_reads(const blight_index* idx, const char* d_bases, const uint64_t* d_read_off, const uint64_t* d_kmer_off, uint64_t n_reads,
                        uint64_t total_bases, const uint32_t* d_table, uint32_t* d_out, uint64_t* d_ctr, void* stream) {
	if (!idx || !d_table || !d_out || !d_ctr || (n_reads && (!d_bases || !d_read_off || !d_kmer_off))) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (!idx->v.pos_id || idx->v.k - idx->v.m + 1 < 8) return fail(BL_ERR_INVALID_ARG, "the fused consumers need the position->id table (N < 2^32-1, BLIGHT_POS_ID) and k-m+1 >= 8");
	DeviceGuard guard(idx->device);
	ReadBatch B;
	B.d_bases = d_bases; B.d_read_off = d_read_off; B.d_kmer_off = d_kmer_off; B.n_reads = n_reads; B.total_bases = total_bases;
	int rc = launch_reads_sink(idx->v, 2, B, const_cast<uint32_t*>(d_table), 0, 0, d_out, d_ctr, static_cast<cudaStream_t>(stream));
	return rc == BL_OK ? rc : fail(rc, std::string("kernel launch failed: ") + g_last_cuda_error);
}

uint64_t blight_launch_count(void) { return g_launches.load(); }

}  // extern "C"
