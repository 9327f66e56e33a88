// comm.cu — the multi-GPU layer behind the C ABI for ONE process driving several devices of a box (SURVEY.md §8b:
// "blightgpu_comm_init(n_gpus, mode)"), so that a kmer_Set_Light drop-in reaches all GPUs without any launcher:
//
//   BLIGHT_COMM_REPLICA     the whole index on every device, a batch's reads cut into one contiguous share per device
//                           (BASELINE configs[3]); no exchange on the data path, counters summed on the host
//   BLIGHT_COMM_PARTITION   the 2^n MPHF groups cut into contiguous ranges balanced by k-mer count, one slice per device
//                           (BASELINE configs[4]); every device runs the front end on its share of the reads and the
//                           super-k-mers travel to the owner of their minimizer bucket as peer-memory stores
//                           (part_session.cu, connect_local: the devices enable peer access instead of CUDA IPC)
//
// One host thread per device for the duration of a call; the only synchronisation between devices on the data path is the
// device-side flags of the sessions. The process-per-GPU form of the same two modes (torchrun) is blight_b200/dist.py.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "capi_common.hpp"
#include "device_index.hpp"
#include "kernels.hpp"

using namespace blight;

namespace {

constexpr int kMaxRanks = BLIGHT_MAX_RANKS;

int cu_fail(cudaError_t e, const char* what) { return fail(BL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); }
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cu_fail(e__, #call); } while (0)

// grow-only device buffers of one rank of the partitioned mode
struct RankWs {
	void* p[4] = {nullptr, nullptr, nullptr, nullptr};  // text, offsets (beg | end | koff), ctr, spare
	size_t cap[4] = {0, 0, 0, 0};
	cudaStream_t st = nullptr;
	int reserve(int slot, size_t bytes) {
		if (cap[slot] >= bytes) return BL_OK;
		if (p[slot]) cudaFree(p[slot]);
		p[slot] = nullptr; cap[slot] = 0;
		const size_t c = bytes + bytes / 8 + 4096;
		cudaError_t e = cudaMalloc(&p[slot], c);
		if (e != cudaSuccess) return cu_fail(e, "cudaMalloc(comm workspace)");
		cap[slot] = c;
		return BL_OK;
	}
};

// runs fn(g) for every rank on its own host thread; first error wins (its message is re-raised on the calling thread)
template <class F>
int for_each_rank(uint32_t n, F fn) {
	std::vector<int> rcs(n, BL_OK);
	std::vector<std::string> errs(n);
	std::vector<std::thread> th;
	for (uint32_t g = 0; g < n; g++)
		th.emplace_back([&, g] {
			rcs[g] = fn(g);
			if (rcs[g] != BL_OK) errs[g] = g_last_error;
		});
	for (auto& t : th) t.join();
	for (uint32_t g = 0; g < n; g++)
		if (rcs[g] != BL_OK) return fail(rcs[g], errs[g]);
	return BL_OK;
}

}  // namespace

struct blight_comm {
	uint32_t n = 0;
	int mode = BLIGHT_COMM_REPLICA;
	uint32_t k = 0, m = 0, lb = 0;
	int devices[kMaxRanks] = {};
	blight_index* idx[kMaxRanks] = {};
	// partition mode
	uint32_t cuts[kMaxRanks + 1] = {};
	blight_part_session* sess[kMaxRanks] = {};
	uint64_t sub = 0, cap = 0, ids_cap = 0;
	RankWs ws[kMaxRanks];
	blight_info info{};  // of the whole index
	std::mutex call;     // partitioned calls are collective over the devices: one at a time
	void* stream_host = nullptr;  // pinned buffers of the streaming file_query (stream_query.cu), created on first use
	std::mutex file_call;         // one streaming file_query at a time (they share those buffers)
};

namespace {

void free_sessions(blight_comm* c) {
	for (uint32_t g = 0; g < c->n; g++) {
		if (c->sess[g]) { cudaSetDevice(c->devices[g]); cudaDeviceSynchronize(); }
	}
	for (uint32_t g = 0; g < c->n; g++) { blight_part_session_free(c->sess[g]); c->sess[g] = nullptr; }
}

int make_sessions(blight_comm* c, uint64_t sub, uint64_t cap, uint64_t ids_cap) {
	free_sessions(c);
	c->sub = sub; c->cap = cap; c->ids_cap = ids_cap;
	for (uint32_t g = 0; g < c->n; g++) {
		blight_part_config cfg{};
		cfg.world = c->n; cfg.rank = g; cfg.lb = c->lb; cfg.sub_positions = sub; cfg.cap = cap; cfg.ids_capacity = ids_cap;
		cfg.ret_kmers = c->n <= 2 || cap >= sub ? sub : uint64_t(sub * 2.5 / c->n);  // cap >= sub: the retry after an overflow, sized for the worst case
		for (uint32_t i = 0; i <= c->n; i++) cfg.cuts[i] = c->cuts[i];
		int rc = blight_part_session_create(c->idx[g], &cfg, &c->sess[g]);
		if (rc != BL_OK) return rc;
	}
	for (uint32_t g = 0; g < c->n; g++)
		for (uint32_t q = 0; q < c->n; q++)
			if (q != g) {
				int rc = blight_part_session_connect_local(c->sess[g], q, c->sess[q]);
				if (rc != BL_OK) return rc;
			}
	return BL_OK;
}

// reads [r_lo, r_hi) of every rank: contiguous shares balanced by bases
void split_reads(const uint64_t* beg, uint64_t n_reads, uint32_t world, uint64_t* r_cut) {
	const uint64_t b0 = beg[0], total = beg[n_reads] - b0;
	r_cut[0] = 0;
	for (uint32_t g = 1; g < world; g++) {
		const uint64_t target = b0 + total / world * g;
		uint64_t r = uint64_t(std::lower_bound(beg, beg + n_reads + 1, target) - beg);
		r_cut[g] = std::max(r_cut[g - 1], std::min(r, n_reads));
	}
	r_cut[world] = n_reads;
}

// One collective batch of the partitioned mode. text / beg / end describe the records (end may be null); koff_local[g] the
// exclusive prefix of k-mer counts of rank g's records (only when ids are wanted), ids_out[g] where its ids go.
int partition_batch(blight_comm* c, const char* text, const uint64_t* beg, const uint64_t* end, const uint64_t* r_cut,
                    const std::vector<std::vector<uint64_t>>* koff_local, int64_t* const* ids_out, uint64_t* ctr) {
	const uint32_t W = c->n;
	const bool want_ids = ids_out != nullptr;
	uint64_t max_len = 0, max_kmers = 0;
	for (uint32_t g = 0; g < W; g++) {
		if (r_cut[g + 1] > r_cut[g]) max_len = std::max(max_len, (end ? end[r_cut[g + 1] - 1] : beg[r_cut[g + 1]]) - beg[r_cut[g]]);
		if (want_ids) max_kmers = std::max<uint64_t>(max_kmers, (*koff_local)[g].back());
	}
	if (!c->sess[0] || (want_ids && c->ids_cap < max_kmers)) {
		const double rpp = std::min(0.25, std::max(0.03, 0.5 / W));
		const uint64_t max_cap = (1ull << 24) - 1;
		uint64_t sub = std::min<uint64_t>(64ull << 20, uint64_t(max_cap / rpp)) / kReadsStrip * kReadsStrip;  // 8 B200s, 480 M k-mers per GPU: ids 19.5 / 18.4 / 18.0 ms at 32 / 64 / 128 M
		if (c->sub) sub = c->sub;
		uint64_t cap = c->cap ? c->cap : std::min<uint64_t>(std::max<uint64_t>(1024, uint64_t(sub * rpp)), max_cap);
		if (const char* e = getenv("BLIGHT_PART_CAP")) { const uint64_t v = strtoull(e, nullptr, 10); if (v && !c->cap) cap = std::min(v, max_cap); }  // test knob
		int rc = make_sessions(c, sub, cap, want_ids ? max_kmers + max_kmers / 8 + 1 : c->ids_cap);
		if (rc != BL_OK) return rc;
	}
	const uint64_t n_sub = std::max<uint64_t>(1, blight_part_session_sub_batches(c->sess[0], max_len, want_ids ? 1 : 0));
	std::vector<uint64_t> ctrs((size_t)W * BLIGHT_N_CTR, 0);
	std::vector<uint32_t> flags(W, 0);
	// Phase 1, every rank: allocations only. cudaMalloc waits for the device to drain (and, with peer access, touches the
	// peers): it must never run while another rank's wait kernel is spinning on a flag this rank has yet to publish.
	int rc = for_each_rank(W, [&](uint32_t g) -> int {
		cudaSetDevice(c->devices[g]);
		RankWs& w = c->ws[g];
		if (!w.st) CU(cudaStreamCreateWithFlags(&w.st, cudaStreamNonBlocking));
		const uint64_t r0 = r_cut[g], r1 = r_cut[g + 1], cnt = r1 - r0;
		const uint64_t len = cnt ? (end ? end[r1 - 1] : beg[r1]) - beg[r0] : 0;
		int rc2;
		if ((rc2 = w.reserve(0, len + 64)) != BL_OK) return rc2;
		if ((rc2 = w.reserve(1, (cnt + 1) * 8 * 3 + 64)) != BL_OK) return rc2;
		return w.reserve(2, BLIGHT_N_CTR * 8);
	});
	if (rc != BL_OK) return rc;
	// Phase 2: copies, the collective pipeline, results. Nothing below allocates.
	rc = for_each_rank(W, [&](uint32_t g) -> int {
		cudaSetDevice(c->devices[g]);
		RankWs& w = c->ws[g];
		const uint64_t r0 = r_cut[g], r1 = r_cut[g + 1], cnt = r1 - r0;
		const uint64_t t0 = beg[r0], len = cnt ? (end ? end[r1 - 1] : beg[r1]) - t0 : 0;
		int rc2;
		uint64_t* d_beg = static_cast<uint64_t*>(w.p[1]);
		uint64_t* d_end = d_beg + cnt + 1;
		uint64_t* d_koff = d_end + cnt + 1;
		uint64_t* d_ctr = static_cast<uint64_t*>(w.p[2]);
		CU(cudaMemsetAsync(d_ctr, 0, BLIGHT_N_CTR * 8, w.st));
		ReadBatch B;
		std::vector<uint64_t> hb, he;
		if (cnt) {
			hb.resize(cnt + 1);
			for (uint64_t i = 0; i <= cnt; i++) hb[i] = beg[r0 + i] - t0;
			if (end) hb[cnt] = len;  // the entry after the last record: the end of this rank's text
			CU(cudaMemcpyAsync(w.p[0], text + t0, len, cudaMemcpyHostToDevice, w.st));
			CU(cudaMemcpyAsync(d_beg, hb.data(), (cnt + 1) * 8, cudaMemcpyHostToDevice, w.st));
			if (end) {
				he.resize(cnt);
				for (uint64_t i = 0; i < cnt; i++) he[i] = end[r0 + i] - t0;
				CU(cudaMemcpyAsync(d_end, he.data(), cnt * 8, cudaMemcpyHostToDevice, w.st));
			}
			if (want_ids) CU(cudaMemcpyAsync(d_koff, (*koff_local)[g].data(), (cnt + 1) * 8, cudaMemcpyHostToDevice, w.st));
			B.d_bases = static_cast<const char*>(w.p[0]); B.d_read_off = d_beg; B.d_read_end = end ? d_end : nullptr;
			B.d_kmer_off = want_ids ? d_koff : nullptr;
			B.n_reads = cnt; B.total_bases = len;
		}
		if (!cnt && want_ids) B.d_kmer_off = d_koff;  // the mode is collective: an empty rank still runs the id pipeline
		rc2 = part_session_query_batch(c->sess[g], B, n_sub, d_ctr, w.st);
		if (rc2 != BL_OK) return rc2;
		if (want_ids && cnt && (*koff_local)[g].back())
			CU(cudaMemcpyAsync(ids_out[g], blight_part_session_ids(c->sess[g]), (*koff_local)[g].back() * 8, cudaMemcpyDeviceToHost, w.st));
		CU(cudaMemcpyAsync(&ctrs[(size_t)g * BLIGHT_N_CTR], d_ctr, BLIGHT_N_CTR * 8, cudaMemcpyDeviceToHost, w.st));
		rc2 = blight_part_session_status(c->sess[g], &flags[g], 1, w.st);  // synchronises the stream
		return rc2;
	});
	if (rc != BL_OK) return rc;
	uint32_t fl = 0;
	for (uint32_t g = 0; g < W; g++) {
		fl |= flags[g];
		for (int i = 0; i < BLIGHT_N_CTR; i++) ctr[i] += ctrs[(size_t)g * BLIGHT_N_CTR + i];
	}
	if (fl & BLIGHT_PART_TIMEOUT) return fail(BL_ERR_CUDA, "partitioned query: a device's flag never arrived");
	if (fl & BLIGHT_PART_OVERFLOW) return BLIGHT_PART_OVERFLOW;  // positive: the caller retries with roomier inboxes
	return BL_OK;
}

int partition_records(blight_comm* c, const char* text, const uint64_t* beg, const uint64_t* end, uint64_t n_reads, int64_t* ids_out,
                      uint64_t* ctr) {
	std::lock_guard<std::mutex> lock(c->call);
	const uint32_t W = c->n;
	uint64_t r_cut[kMaxRanks + 1];
	split_reads(beg, n_reads, W, r_cut);
	std::vector<std::vector<uint64_t>> koff(ids_out ? W : 0);
	int64_t* outs[kMaxRanks] = {};
	if (ids_out) {
		uint64_t base = 0;
		for (uint32_t g = 0; g < W; g++) {
			const uint64_t r0 = r_cut[g], cnt = r_cut[g + 1] - r0;
			koff[g].assign(cnt + 1, 0);
			for (uint64_t i = 0; i < cnt; i++) {
				const uint64_t l = (end ? end[r0 + i] : beg[r0 + i + 1]) - beg[r0 + i];
				koff[g][i + 1] = koff[g][i] + (l >= c->k ? l - c->k + 1 : 0);
			}
			outs[g] = ids_out + base;
			base += koff[g].back();
		}
	}
	for (int attempt = 0; attempt < 2; attempt++) {
		uint64_t tmp[BLIGHT_N_CTR] = {0, 0, 0, 0};
		int rc = partition_batch(c, text, beg, end, r_cut, ids_out ? &koff : nullptr, ids_out ? outs : nullptr, tmp);
		if (rc == (int)BLIGHT_PART_OVERFLOW && attempt == 0) {
			// far more super-k-mers per base than a read batch has: one record per position can never overflow
			const uint64_t sub = std::min<uint64_t>(c->sub, 8ull << 20);
			rc = make_sessions(c, sub, sub, c->ids_cap);
			if (rc != BL_OK) return rc;
			continue;
		}
		if (rc != BL_OK) return rc > 0 ? fail(BL_ERR_NOMEM, "partitioned query: inbox overflow") : rc;
		for (int i = 0; i < BLIGHT_N_CTR; i++) ctr[i] = tmp[i];
		break;
	}
	if (ctr[BLIGHT_CTR_INVALID]) return fail(BL_ERR_INVALID_BASE, "Invalid char in DNA");
	return BL_OK;
}

int replica_records(blight_comm* c, const char* text, const uint64_t* beg, const uint64_t* end, uint64_t n_reads, int64_t* ids_out,
                    uint64_t* ctr) {
	const uint32_t W = c->n;
	uint64_t r_cut[kMaxRanks + 1];
	split_reads(beg, n_reads, W, r_cut);
	// where each share's ids start: k-mers of the reads before it
	std::vector<uint64_t> id_base(W + 1, 0);
	std::vector<std::vector<uint64_t>> koff(ids_out ? W : 0);
	if (ids_out)
		for (uint32_t g = 0; g < W; g++) {
			const uint64_t r0 = r_cut[g], cnt = r_cut[g + 1] - r0;
			koff[g].assign(cnt + 1, 0);
			for (uint64_t i = 0; i < cnt; i++) {
				const uint64_t l = (end ? end[r0 + i] : beg[r0 + i + 1]) - beg[r0 + i];
				koff[g][i + 1] = koff[g][i] + (l >= c->k ? l - c->k + 1 : 0);
			}
			id_base[g + 1] = id_base[g] + koff[g].back();
		}
	std::vector<uint64_t> ctrs((size_t)W * BLIGHT_N_CTR, 0);
	int rc = for_each_rank(W, [&](uint32_t g) -> int {
		const uint64_t r0 = r_cut[g], r1 = r_cut[g + 1], cnt = r1 - r0;
		if (!cnt) return BL_OK;
		const uint64_t t0 = beg[r0], len = (end ? end[r1 - 1] : beg[r1]) - t0;
		std::vector<uint64_t> hb(cnt + 1), he;
		for (uint64_t i = 0; i <= cnt; i++) hb[i] = beg[r0 + i] - t0;
		if (end) {
			hb[cnt] = len;
			he.resize(cnt);
			for (uint64_t i = 0; i < cnt; i++) he[i] = end[r0 + i] - t0;
		}
		return host_query_records(c->idx[g], text + t0, len, hb.data(), end ? he.data() : nullptr, cnt, ids_out ? koff[g].data() : nullptr,
		                          ids_out ? ids_out + id_base[g] : nullptr, ids_out ? koff[g].back() : 0, &ctrs[(size_t)g * BLIGHT_N_CTR], end == nullptr);
	});
	// counters first: an invalid base on one device must not hide the others' counts from the caller's error path
	for (uint32_t g = 0; g < W; g++)
		for (int i = 0; i < BLIGHT_N_CTR; i++) ctr[i] += ctrs[(size_t)g * BLIGHT_N_CTR + i];
	return rc;
}

int comm_records(blight_comm* c, const char* text, const uint64_t* beg, const uint64_t* end, uint64_t n_reads, int64_t* ids_out, uint64_t* ctr) {
	std::memset(ctr, 0, sizeof(uint64_t) * BLIGHT_N_CTR);
	if (n_reads == 0) return BL_OK;
	return c->mode == BLIGHT_COMM_PARTITION ? partition_records(c, text, beg, end, n_reads, ids_out, ctr)
	                                        : replica_records(c, text, beg, end, n_reads, ids_out, ctr);
}

}  // namespace

extern "C" {

int blight_comm_init(const blight_flat* ff, const int* devices, uint32_t n_gpus, int mode, const blight_upload_options* opts, blight_comm** out) {
	if (!ff || !out || !devices) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (n_gpus == 0 || n_gpus > (uint32_t)kMaxRanks) return fail(BL_ERR_INVALID_ARG, "1 .. 16 devices");
	if (mode != BLIGHT_COMM_REPLICA && mode != BLIGHT_COMM_PARTITION) return fail(BL_ERR_INVALID_ARG, "unknown mode");
	const FlatIndex& F = ff->f;
	std::unique_ptr<blight_comm> c(new blight_comm());
	c->n = n_gpus; c->mode = mode; c->k = F.h.k; c->m = F.h.m; c->lb = F.lb();
	for (uint32_t g = 0; g < n_gpus; g++) c->devices[g] = devices[g];
	fill_info(F, &c->info);
	int rc = BL_OK;
	if (mode == BLIGHT_COMM_PARTITION) {
		if (F.h.k < 8 || F.h.k - F.h.m + 1 < 8) return fail(BL_ERR_INVALID_ARG, "partition mode needs k >= 8 and k-m+1 >= 8");
		if (F.h.n_mphf < n_gpus) return fail(BL_ERR_INVALID_ARG, "fewer MPHF groups than devices: use n >= log2(devices)");
		// cut after the group at which the running k-mer count first reaches g/world of the total, at least one group per device
		const uint64_t n = F.h.n_mphf;
		std::vector<double> csum(n);
		double acc = 0;
		for (uint64_t i = 0; i < n; i++) { acc += (double)F.mphf[i].nelem; csum[i] = acc; }
		c->cuts[0] = 0;
		for (uint32_t g = 1; g < n_gpus; g++) {
			uint64_t cut = uint64_t(std::lower_bound(csum.begin(), csum.end(), acc * g / n_gpus) - csum.begin()) + 1;
			cut = std::max<uint64_t>(cut, c->cuts[g - 1] + 1);
			cut = std::min<uint64_t>(cut, n - (n_gpus - g));
			c->cuts[g] = (uint32_t)cut;
		}
		c->cuts[n_gpus] = (uint32_t)n;
	}
	blight_comm* cp = c.get();
	rc = for_each_rank(n_gpus, [&](uint32_t g) -> int {
		if (mode == BLIGHT_COMM_REPLICA) return blight_index_upload_opts(ff, cp->devices[g], opts, &cp->idx[g]);
		blight_flat slice;
		std::string err;
		int r = flat_slice(F, cp->cuts[g], cp->cuts[g + 1], slice.f, &err);
		if (r != BL_OK) return fail(r, err);
		return blight_index_upload_opts(&slice, cp->devices[g], opts, &cp->idx[g]);
	});
	if (rc != BL_OK) { const std::string msg = g_last_error; blight_comm_free(c.release()); return fail(rc, msg); }
	if (mode == BLIGHT_COMM_PARTITION)
		for (uint32_t g = 0; g < n_gpus; g++)
			if (!cp->idx[g]->v.pos_id) { blight_comm_free(c.release()); return fail(BL_ERR_INVALID_ARG, "partition mode needs the position->id table on every slice"); }
	*out = c.release();
	return BL_OK;
}

void blight_comm_free(blight_comm* c) {
	if (!c) return;
	int prev = -1;
	cudaGetDevice(&prev);
	free_sessions(c);
	for (uint32_t g = 0; g < c->n; g++) {
		cudaSetDevice(c->devices[g]);
		for (void* p : c->ws[g].p) cudaFree(p);
		if (c->ws[g].st) cudaStreamDestroy(c->ws[g].st);
		blight_index_free(c->idx[g]);
	}
	if (c->stream_host) stream_host_free(c->stream_host);
	if (prev >= 0) cudaSetDevice(prev);
	delete c;
}

int blight_comm_describe(const blight_comm* c, blight_comm_info* out) {
	if (!c || !out) return fail(BL_ERR_INVALID_ARG, "null argument");
	std::memset(out, 0, sizeof *out);
	out->n_gpus = c->n; out->mode = (uint32_t)c->mode; out->whole = c->info;
	for (uint32_t g = 0; g < c->n; g++) {
		out->devices[g] = c->devices[g];
		out->device_bytes[g] = c->idx[g]->info.device_bytes;
		out->kmers[g] = c->idx[g]->info.number_kmer;
	}
	for (uint32_t g = 0; g <= c->n; g++) out->cuts[g] = c->cuts[g];
	return BL_OK;
}

int blight_comm_query_reads_host(blight_comm* c, const char* bases, const uint64_t* read_off, uint64_t n_reads, int64_t* ids_out, uint64_t* ctr) {
	if (!c || !ctr || (n_reads && (!bases || !read_off))) return fail(BL_ERR_INVALID_ARG, "null argument");
	return comm_records(c, bases, read_off, nullptr, n_reads, ids_out, ctr);
}

int blight_comm_query_fasta_host(blight_comm* c, const char* text, uint64_t len, uint64_t* ctr) {
	if (!c || !ctr || (len && !text)) return fail(BL_ERR_INVALID_ARG, "null argument");
	std::vector<SeqView> recs;
	split_fasta_records(text, len, recs);
	std::vector<uint64_t> beg(recs.size() + 1), end(recs.size());
	for (size_t i = 0; i < recs.size(); i++) { beg[i] = uint64_t(recs[i].p - text); end[i] = beg[i] + recs[i].len; }
	beg[recs.size()] = len;
	return comm_records(c, text, beg.data(), end.data(), end.size(), nullptr, ctr);
}

int blight_comm_query_file_host(blight_comm* c, const char* path, uint64_t* ctr) {
	if (!c || !ctr || !path) return fail(BL_ERR_INVALID_ARG, "null argument");
	// file_query(path) over several GPUs: the streaming reader of the one-GPU path (stream_query.cu: parallel pread into pinned
	// buffers, record cut on the host cores), each batch of records shared out over the devices. The file is never held in
	// memory; from tmpfs the reader (~27 GB/s of FASTA) is the bound, whatever the number of GPUs.
	std::lock_guard<std::mutex> lk(c->file_call);
	std::memset(ctr, 0, sizeof(uint64_t) * BLIGHT_N_CTR);
	return stream_fasta_chunks(path, &c->stream_host, [&](const char* text, uint64_t, const uint64_t* beg, const uint64_t* end, uint64_t n_rec) -> int {
		uint64_t part[BLIGHT_N_CTR] = {};
		const int rc = comm_records(c, text, beg, end, n_rec, nullptr, part);
		for (int i = 0; i < BLIGHT_N_CTR; i++) ctr[i] += part[i];
		return rc;
	});
}

int blight_comm_query_sequence_host(blight_comm* c, const char* seq, uint64_t len, int64_t* ids_out, uint64_t* n_out) {
	if (!c || !n_out || (len && !seq)) return fail(BL_ERR_INVALID_ARG, "null argument");
	*n_out = len >= c->k ? len - c->k + 1 : 0;
	if (*n_out == 0) return BL_OK;
	if (!ids_out) return fail(BL_ERR_INVALID_ARG, "null argument");
	uint64_t off[2] = {0, len}, ctr[BLIGHT_N_CTR];
	return comm_records(c, seq, off, nullptr, 1, ids_out, ctr);
}

}  // extern "C"
