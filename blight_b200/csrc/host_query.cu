// host_query.cu — the host-buffer (end to end) entry points of the C ABI: H2D, kernels and D2H inside the call.
//
// A batch in host memory is cut into chunks of base positions; every chunk is one copy and one launch of the read kernel
// over the k-mers that start inside it, so copies overlap lookups. Two producers feed the GPU from the two ENDS of the batch
// (kernels of different chunks are independent: counters are atomics, ids land at absolute slots):
//
//   front  (the calling thread)   chunk 0, 1, 2, ...  as ASCII, straight from the caller's buffer, at most two copies queued
//   back   (a second thread)      chunk n-1, n-2, ... packed to 2 bits per base by the host cores first (host_pack.cpp)
//
// Both enqueue their copies on ONE copy stream: with a stream each, the DMA engine served the front's never-empty queue and
// let the back's copies wait until the very end (measured: 4 of 47 chunks packed, 25 ms of the back thread spent waiting
// for its first staging buffer). In one stream a packed copy waits for at most the two raw copies queued before it.
//
// They meet wherever the PCIe link and the packer cores balance: on a box where one GPU has a whole x16 link and little
// else to do with its cores, most bytes travel raw; on a box where 8 GPUs share links and cores, the mix shifts by itself.
// A chunk holding a byte nuc2int rejects (kmer.h:68) is never packed — it travels as ASCII so that the kernel applies the
// reference's rule (only bytes under a queried k-mer raise). Each launch gets the WINDOW of the read-offset arrays that
// overlaps its chunk, copied with the chunk.
//
// Contexts (streams, events, pinned staging, device workspaces) are pooled per index: concurrent calls from several host
// threads (the reference's callers query from OpenMP loops, Abundance_De_Bruijn_graph_snippet.cpp:122-125) each take their
// own context instead of serialising on one stream.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "capi_common.hpp"
#include "device_index.hpp"
#include "host_pack.hpp"
#include "kernels.hpp"

using namespace blight;

namespace blight {
std::atomic<uint64_t> g_h2d_bytes{0}, g_d2h_bytes{0}, g_packed_bases{0}, g_pack_ns{0};
}

namespace {

constexpr int kSlots = 3;      // pinned staging buffers of the packer
constexpr int kWs = 8;         // device workspaces: 0 text, 1 beg, 2 end, 3 ctr, 4 koff, 5 ids, 6 packed, 7 canon
constexpr uint64_t kHalo = kReadsStrip;

int cuda_fail(cudaError_t e, const char* what) { return fail(BL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); }
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cuda_fail(e__, #call); } while (0)

struct DeviceGuard {
	int prev = -1;
	explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
	~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// The packer's own worker threads (not OpenMP: the process may hold a second OpenMP runtime — PyTorch ships one — and a
// team forked from a short-lived thread is neither cheap nor reliably wide). Workers pull block numbers from a shared counter.
class PackPool {
public:
	explicit PackPool(int n_threads) {
		for (int i = 1; i < n_threads; i++) workers_.emplace_back([this] { loop(); });
	}
	~PackPool() {
		{ std::lock_guard<std::mutex> l(m_); stop_ = true; gen_++; }
		cv_.notify_all();
		for (auto& t : workers_) t.join();
	}
	// packs text[0, n_bases) into words; false if a byte outside ACGTacgt was seen. The caller works too.
	bool run(const char* text, uint64_t n_bases, uint32_t* words) {
		text_ = text; n_bases_ = n_bases; words_ = words;
		n_blocks_ = int64_t((n_bases + kBlock - 1) / kBlock);
		next_.store(0); bad_.store(0);
		{ std::lock_guard<std::mutex> l(m_); pending_ = (int)workers_.size(); gen_++; }
		cv_.notify_all();
		work();
		std::unique_lock<std::mutex> l(m_);
		done_.wait(l, [&] { return pending_ == 0; });
		return bad_.load() == 0;
	}
	int threads() const { return (int)workers_.size() + 1; }

private:
	static constexpr uint64_t kBlock = 1u << 16;  // bases per block: a multiple of 16, every block starts on a word
	void work() {
		for (;;) {
			const int64_t b = next_.fetch_add(1);
			if (b >= n_blocks_) return;
			const uint64_t o = uint64_t(b) * kBlock;
			if (!pack2_block(text_ + o, std::min<uint64_t>(kBlock, n_bases_ - o), words_ + (o >> 4))) bad_.fetch_add(1);
		}
	}
	void loop() {
		uint64_t seen = 0;
		for (;;) {
			{
				std::unique_lock<std::mutex> l(m_);
				cv_.wait(l, [&] { return gen_ != seen; });
				seen = gen_;
				if (stop_) return;
			}
			work();
			{ std::lock_guard<std::mutex> l(m_); if (--pending_ == 0) done_.notify_one(); }
		}
	}
	std::vector<std::thread> workers_;
	std::mutex m_;
	std::condition_variable cv_, done_;
	uint64_t gen_ = 0;
	int pending_ = 0;
	bool stop_ = false;
	const char* text_ = nullptr; uint64_t n_bases_ = 0; uint32_t* words_ = nullptr; int64_t n_blocks_ = 0;
	std::atomic<int64_t> next_{0};
	std::atomic<int> bad_{0};
};

struct HostCtx {
	std::unique_ptr<PackPool> pool;
	int pack_skip = 0;  // calls left without the packer (it did not pay last time it ran, see host_query_records)
	cudaStream_t st_f = nullptr, st_b = nullptr, cs_raw = nullptr, cs_pk = nullptr;
	cudaEvent_t ev_f[2] = {nullptr, nullptr}, ev_slot[kSlots] = {nullptr, nullptr, nullptr}, ev_join = nullptr;
	uint32_t* stage[kSlots] = {nullptr, nullptr, nullptr};
	size_t stage_cap = 0;  // bytes each
	void* ws[kWs] = {};
	size_t ws_cap[kWs] = {};
	int init() {
		for (cudaStream_t* s : {&st_f, &st_b, &cs_raw, &cs_pk}) CU(cudaStreamCreateWithFlags(s, cudaStreamNonBlocking));
		for (cudaEvent_t* e : {&ev_f[0], &ev_f[1], &ev_slot[0], &ev_slot[1], &ev_slot[2], &ev_join}) CU(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
		return BL_OK;
	}
	~HostCtx() {
		for (cudaStream_t s : {st_f, st_b, cs_raw, cs_pk}) if (s) cudaStreamDestroy(s);
		for (cudaEvent_t e : {ev_f[0], ev_f[1], ev_slot[0], ev_slot[1], ev_slot[2], ev_join}) if (e) cudaEventDestroy(e);
		for (uint32_t* p : stage) if (p) cudaFreeHost(p);
		for (void* p : ws) if (p) cudaFree(p);
	}
	// grow-only device scratch
	int reserve(int slot, size_t bytes, void** out) {
		if (ws_cap[slot] < bytes) {
			if (ws[slot]) cudaFree(ws[slot]);
			ws[slot] = nullptr; ws_cap[slot] = 0;
			const size_t cap = bytes + bytes / 8 + 4096;
			cudaError_t e = cudaMalloc(&ws[slot], cap);
			if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(workspace)");
			ws_cap[slot] = cap;
		}
		*out = ws[slot];
		return BL_OK;
	}
	int reserve_stage(size_t bytes) {
		if (stage_cap >= bytes) return BL_OK;
		for (uint32_t*& p : stage) { if (p) cudaFreeHost(p); p = nullptr; }
		stage_cap = 0;
		for (uint32_t*& p : stage) CU(cudaHostAlloc(reinterpret_cast<void**>(&p), bytes, cudaHostAllocDefault));
		stage_cap = bytes;
		return BL_OK;
	}
};

struct HostPool {
	std::mutex m;
	std::vector<HostCtx*> idle, all;
	~HostPool() { for (HostCtx* c : all) delete c; }
};

std::mutex g_pool_create;

// one context for the duration of a call
struct Lease {
	HostPool* pool = nullptr;
	HostCtx* c = nullptr;
	int acquire(const blight_index* idx) {
		blight_index* mi = const_cast<blight_index*>(idx);
		{
			std::lock_guard<std::mutex> l(g_pool_create);
			if (!mi->host_pool) mi->host_pool = new HostPool();
		}
		pool = static_cast<HostPool*>(mi->host_pool);
		{
			std::lock_guard<std::mutex> l(pool->m);
			if (!pool->idle.empty()) { c = pool->idle.back(); pool->idle.pop_back(); return BL_OK; }
		}
		HostCtx* n = new HostCtx();
		const int rc = n->init();
		if (rc != BL_OK) { delete n; return rc; }
		std::lock_guard<std::mutex> l(pool->m);
		pool->all.push_back(n);
		c = n;
		return BL_OK;
	}
	~Lease() {
		if (c) { std::lock_guard<std::mutex> l(pool->m); pool->idle.push_back(c); }
	}
};

// host threads the packer may use: the cores of the box shared between its GPUs (a process per GPU, or a thread per GPU
// of one process, each packs for its own device); BLIGHT_HOST_THREADS overrides
int pack_threads() {
	static const int n = [] {
		if (const char* e = getenv("BLIGHT_HOST_THREADS")) { const int v = atoi(e); if (v > 0) return v; }
		int devs = 1;
		if (cudaGetDeviceCount(&devs) != cudaSuccess || devs < 1) devs = 1;
		const int hw = (int)std::thread::hardware_concurrency();
		return std::max(1, std::min(32, hw / devs));
	}();
	return n;
}

bool pack_enabled() {
	static const bool on = [] { const char* e = getenv("BLIGHT_HOST_PACK"); return !(e && atoi(e) == 0); }();
	return on;
}

struct Window { uint64_t r_lo, r_hi; };  // reads overlapping a chunk: entries [r_lo, r_hi] of beg

Window window_of(const uint64_t* beg, uint64_t n, uint64_t c0, uint64_t upto) {
	uint64_t lo = uint64_t(std::upper_bound(beg, beg + n + 1, c0) - beg);
	lo = lo ? lo - 1 : 0;                                                          // the read holding c0 (or the gap before it)
	uint64_t hi = uint64_t(std::lower_bound(beg + lo, beg + n + 1, upto) - beg);  // first read starting at or past `upto`
	if (hi > n) hi = n;
	if (hi <= lo) hi = std::min(n, lo + 1);
	return Window{lo, hi};
}

// BLIGHT_HOST_DEBUG=1: one line per call on stderr — chunks and host time of either producer
struct ProducerStats { int chunks = 0; double t_sync = 0, t_pack = 0, t_enqueue = 0; };
double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct Plan {
	ProducerStats sf, sb;
	const blight_index* idx;
	HostCtx* C;
	const char* text; uint64_t len;
	const uint64_t *beg, *end, *koff; uint64_t n;
	char* d_text; uint32_t* d_packed; uint64_t *d_beg, *d_end, *d_koff, *d_ctr; int64_t* d_ids;
	std::vector<uint64_t> cut;  // chunk i = [cut[i], cut[i+1])
	std::mutex m;
	int64_t front = 0, back = 0;  // unclaimed chunks are [front, back)
	int claim_front() { std::lock_guard<std::mutex> l(m); return front < back ? (int)front++ : -1; }
	int claim_back() { std::lock_guard<std::mutex> l(m); return back > front ? (int)--back : -1; }
};

// copies the offset window of chunk c on `cs` and launches its kernel on `st` once `ev` (recorded on cs by the caller after
// the text copy and these copies) has fired
int window_and_launch(Plan& P, int c, bool packed, cudaStream_t cs, cudaEvent_t ev, cudaStream_t st) {
	const uint64_t c0 = P.cut[c], c1 = P.cut[c + 1];
	const uint64_t upto = std::min(P.len, c1 + kHalo);
	const Window w = window_of(P.beg, P.n, c0, upto);
	const uint64_t cnt = w.r_hi - w.r_lo;
	CU(cudaMemcpyAsync(P.d_beg + w.r_lo, P.beg + w.r_lo, (cnt + 1) * 8, cudaMemcpyHostToDevice, cs));
	uint64_t bytes = (cnt + 1) * 8;
	if (P.end) { CU(cudaMemcpyAsync(P.d_end + w.r_lo, P.end + w.r_lo, cnt * 8, cudaMemcpyHostToDevice, cs)); bytes += cnt * 8; }
	if (P.d_ids) { CU(cudaMemcpyAsync(P.d_koff + w.r_lo, P.koff + w.r_lo, (cnt + 1) * 8, cudaMemcpyHostToDevice, cs)); bytes += (cnt + 1) * 8; }
	g_h2d_bytes += bytes;
	CU(cudaEventRecord(ev, cs));
	CU(cudaStreamWaitEvent(st, ev, 0));
	ReadBatch B;
	if (packed) B.d_packed = P.d_packed; else B.d_bases = P.d_text;
	B.d_read_off = P.d_beg + w.r_lo;
	B.d_read_end = P.end ? P.d_end + w.r_lo : nullptr;
	B.d_kmer_off = P.d_ids ? P.d_koff + w.r_lo : nullptr;
	B.n_reads = cnt;
	B.total_bases = P.len;
	B.guess_p0 = P.beg[w.r_lo];
	const uint64_t span = P.beg[w.r_hi] - P.beg[w.r_lo];
	B.rpb = span ? (double)cnt / (double)span : 0.0;
	const int rc = launch_reads(&P.idx->v, P.idx->v.k, P.idx->v.m, B, nullptr, nullptr, P.d_ids, P.d_ctr, st, c0, c1);
	if (rc != BL_OK) return fail(rc, std::string("kernel launch failed: ") + g_last_cuda_error);
	return BL_OK;
}

int front_producer(Plan& P) {
	HostCtx& C = *P.C;
	for (int nf = 0;; nf++) {
		const int c = P.claim_front();
		if (c < 0) return BL_OK;
		cudaEvent_t ev = C.ev_f[nf & 1];
		double t0 = now_s();
		if (nf >= 2) CU(cudaEventSynchronize(ev));  // at most two raw copies queued: the back producer gets its share of the link
		P.sf.t_sync += now_s() - t0;
		t0 = now_s();
		const uint64_t c0 = P.cut[c], upto = std::min(P.len, P.cut[c + 1] + kHalo);
		CU(cudaMemcpyAsync(P.d_text + c0, P.text + c0, upto - c0, cudaMemcpyHostToDevice, C.cs_raw));
		g_h2d_bytes += upto - c0;
		const int rc = window_and_launch(P, c, false, C.cs_raw, ev, C.st_f);
		if (rc != BL_OK) return rc;
		P.sf.t_enqueue += now_s() - t0;
		P.sf.chunks++;
	}
}

int back_producer(Plan& P) {
	HostCtx& C = *P.C;
	for (int nb = 0;; nb++) {
		const int c = P.claim_back();
		if (c < 0) return BL_OK;
		const int s = nb % kSlots;
		double ts = now_s();
		if (nb >= kSlots) CU(cudaEventSynchronize(C.ev_slot[s]));  // the copy that last read this staging buffer has left
		P.sb.t_sync += now_s() - ts;
		const uint64_t c0 = P.cut[c], upto = std::min(P.len, P.cut[c + 1] + kHalo);
		const uint64_t n_bases = upto - c0, n_words = (n_bases + 15) / 16;
		const auto t0 = std::chrono::steady_clock::now();
		const bool bad = !C.pool->run(P.text + c0, n_bases, C.stage[s]);
		g_pack_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
		P.sb.t_pack += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		ts = now_s();
		if (!bad) {
			g_packed_bases += n_bases;
			CU(cudaMemcpyAsync(P.d_packed + (c0 >> 4), C.stage[s], n_words * 4, cudaMemcpyHostToDevice, C.cs_raw));
			g_h2d_bytes += n_words * 4;
		} else {
			CU(cudaMemcpyAsync(P.d_text + c0, P.text + c0, n_bases, cudaMemcpyHostToDevice, C.cs_raw));
			g_h2d_bytes += n_bases;
		}
		const int rc = window_and_launch(P, c, !bad, C.cs_raw, C.ev_slot[s], C.st_b);
		if (rc != BL_OK) return rc;
		P.sb.t_enqueue += now_s() - ts;
		P.sb.chunks++;
	}
}

}  // namespace

namespace blight {

void host_pool_free(void* pool) { delete static_cast<HostPool*>(pool); }

int host_query_records(const blight_index* idx, const char* text, uint64_t len, const uint64_t* beg, const uint64_t* end, uint64_t n,
                       const uint64_t* koff, int64_t* ids_out, uint64_t total_kmers, uint64_t* ctr, bool allow_pack) {
	std::memset(ctr, 0, sizeof(uint64_t) * BLIGHT_N_CTR);
	if (n == 0 || len == 0) return BL_OK;
	DeviceGuard guard(idx->device);
	Lease lease;
	int rc = lease.acquire(idx);
	if (rc != BL_OK) return rc;
	HostCtx& C = *lease.c;

	uint64_t chunk = 32ull << 20;  // a multiple of kReadsStrip
	if (const char* e = getenv("BLIGHT_HOST_CHUNK_KB")) {  // tuning / test knob
		const uint64_t kb = strtoull(e, nullptr, 10);
		if (kb) chunk = ((kb << 10) + kReadsStrip - 1) / kReadsStrip * kReadsStrip;
	}
	const int threads = pack_threads();
	bool pack = allow_pack && pack_enabled() && len >= 4 * chunk;
	if (pack && C.pack_skip > 0) { C.pack_skip--; pack = false; }

	Plan P{};
	P.idx = idx; P.C = &C; P.text = text; P.len = len; P.beg = beg; P.end = end; P.koff = koff; P.n = n;
	void* p = nullptr;
	if ((rc = C.reserve(0, len + 64, &p)) != BL_OK) return rc;
	P.d_text = static_cast<char*>(p);
	if ((rc = C.reserve(1, (n + 1) * 8, &p)) != BL_OK) return rc;
	P.d_beg = static_cast<uint64_t*>(p);
	if (end) { if ((rc = C.reserve(2, n * 8, &p)) != BL_OK) return rc; P.d_end = static_cast<uint64_t*>(p); }
	if ((rc = C.reserve(3, BLIGHT_N_CTR * 8, &p)) != BL_OK) return rc;
	P.d_ctr = static_cast<uint64_t*>(p);
	if (ids_out) {
		if ((rc = C.reserve(4, (n + 1) * 8, &p)) != BL_OK) return rc;
		P.d_koff = static_cast<uint64_t*>(p);
		if ((rc = C.reserve(5, std::max<uint64_t>(total_kmers, 1) * 8, &p)) != BL_OK) return rc;
		P.d_ids = static_cast<int64_t*>(p);
	}
	if (pack) {
		if ((rc = C.reserve(6, (len + 15) / 16 * 4 + 256, &p)) != BL_OK) return rc;
		P.d_packed = static_cast<uint32_t*>(p);
		if ((rc = C.reserve_stage((chunk + kHalo + 15) / 16 * 4 + 64)) != BL_OK) return rc;
		if (!C.pool || C.pool->threads() != threads) C.pool.reset(new PackPool(threads));
	}
	// chunks: the first ones small (4 MB, doubling), so the first kernel starts after 0.1 ms of copy instead of a millisecond
	P.cut.push_back(0);
	for (uint64_t step = std::min<uint64_t>(chunk, 4ull << 20); P.cut.back() < len; step = std::min(chunk, step * 2))
		P.cut.push_back(std::min(len, P.cut.back() + step));
	P.front = 0;
	P.back = (int64_t)P.cut.size() - 1;

	CU(cudaMemsetAsync(P.d_ctr, 0, BLIGHT_N_CTR * 8, C.st_f));
	CU(cudaEventRecord(C.ev_join, C.st_f));
	CU(cudaStreamWaitEvent(C.st_b, C.ev_join, 0));  // the counters are zero before either stream adds to them

	int rc_back = BL_OK;
	std::string err_back;
	std::thread back;
	if (pack) {
		back = std::thread([&] {
			cudaSetDevice(idx->device);
			rc_back = back_producer(P);
			if (rc_back != BL_OK) err_back = g_last_error;
		});
	}
	rc = front_producer(P);
	if (back.joinable()) back.join();
	if (pack && P.sf.chunks && P.sb.chunks) {
		// Did the packer pay? It spends host memory bandwidth (every base is read by a core AND its packed form by the DMA
		// engine) to save link bandwidth. Where the host side is what limits the box — 8 GPUs pulling from one memory system —
		// its chunks arrive slower than plain copies would: then leave it out for a while (measured on the 8-GPU box: 73 ms
		// per step with it, 67 ms without; on the one-GPU box 30 ms with it, 36 ms without).
		uint64_t fb = 0, bb = 0;
		for (int64_t c = 0; c < P.front; c++) fb += P.cut[c + 1] - P.cut[c];
		for (int64_t c = P.back; c + 1 < (int64_t)P.cut.size(); c++) bb += P.cut[c + 1] - P.cut[c];
		if (bb * 4 < fb * 3) C.pack_skip = 15;
	}
	if (const char* e = getenv("BLIGHT_HOST_DEBUG")) {
		if (atoi(e))
			fprintf(stderr, "[blight host] %zu chunks: front %d (sync %.2f ms, enqueue %.2f ms)  back %d (sync %.2f ms, pack %.2f ms, enqueue %.2f ms)  threads %d\n",
			        P.cut.size() - 1, P.sf.chunks, 1e3 * P.sf.t_sync, 1e3 * P.sf.t_enqueue, P.sb.chunks, 1e3 * P.sb.t_sync, 1e3 * P.sb.t_pack,
			        1e3 * P.sb.t_enqueue, threads);
	}
	if (rc == BL_OK && rc_back != BL_OK) rc = fail(rc_back, err_back);
	if (rc != BL_OK) { cudaStreamSynchronize(C.st_f); cudaStreamSynchronize(C.st_b); cudaStreamSynchronize(C.cs_raw); cudaStreamSynchronize(C.cs_raw); return rc; }
	CU(cudaEventRecord(C.ev_join, C.st_b));
	CU(cudaStreamWaitEvent(C.st_f, C.ev_join, 0));
	if (ids_out && total_kmers) { CU(cudaMemcpyAsync(ids_out, P.d_ids, total_kmers * 8, cudaMemcpyDeviceToHost, C.st_f)); g_d2h_bytes += total_kmers * 8; }
	CU(cudaMemcpyAsync(ctr, P.d_ctr, BLIGHT_N_CTR * 8, cudaMemcpyDeviceToHost, C.st_f));
	g_d2h_bytes += BLIGHT_N_CTR * 8;
	CU(cudaStreamSynchronize(C.st_f));
	if (ctr[BLIGHT_CTR_INVALID]) return fail(BL_ERR_INVALID_BASE, "Invalid char in DNA");
	return BL_OK;
}

}  // namespace blight

extern "C" {

int blight_query_fasta_host(const blight_index* idx, const char* text, uint64_t len, uint64_t* ctr) {
	if (!idx || !ctr || (len && !text)) return fail(BL_ERR_INVALID_ARG, "null argument");
	std::vector<SeqView> recs;
	split_fasta_records(text, len, recs);
	std::vector<uint64_t> beg(recs.size() + 1), end(recs.size());
	for (size_t i = 0; i < recs.size(); i++) { beg[i] = uint64_t(recs[i].p - text); end[i] = beg[i] + recs[i].len; }
	beg[recs.size()] = len;
	return host_query_records(idx, text, len, beg.data(), end.data(), end.size(), nullptr, nullptr, 0, ctr, false);
}

int blight_query_file_host(const blight_index* idx, const char* path, uint64_t* ctr) {
	if (!idx || !ctr || !path) return fail(BL_ERR_INVALID_ARG, "null argument");
	{
		const char* e = getenv("BLIGHT_FILE_QUERY");  // "whole": read the file into memory first (tests compare the two)
		if (!e || e[0] != 'w') return stream_file_query(idx, path, ctr);
	}
	std::string storage, err;
	std::vector<SeqView> recs;
	int rc = read_fasta_records(path, storage, recs, &err);
	if (rc != BL_OK) return fail(rc, err);
	std::vector<uint64_t> beg(recs.size() + 1), end(recs.size());
	for (size_t i = 0; i < recs.size(); i++) { beg[i] = uint64_t(recs[i].p - storage.data()); end[i] = beg[i] + recs[i].len; }
	beg[recs.size()] = storage.size();
	return host_query_records(idx, storage.data(), storage.size(), beg.data(), end.data(), end.size(), nullptr, nullptr, 0, ctr, false);
}

int blight_query_reads_host(const blight_index* idx, const char* bases, const uint64_t* read_off, uint64_t n_reads,
                            int64_t* ids_out, uint64_t* ctr) {
	if (!idx || !ctr || (n_reads && (!bases || !read_off))) return fail(BL_ERR_INVALID_ARG, "null argument");
	std::memset(ctr, 0, sizeof(uint64_t) * BLIGHT_N_CTR);
	if (n_reads == 0) return BL_OK;
	const uint32_t k = idx->v.k;
	const uint64_t base0 = read_off[0];
	std::vector<uint64_t> rebased, koff;
	const uint64_t* beg = read_off;
	if (base0 != 0) {
		rebased.assign(read_off, read_off + n_reads + 1);
		for (auto& v : rebased) v -= base0;
		beg = rebased.data();
	}
	if (ids_out) {
		koff.assign(n_reads + 1, 0);
		for (uint64_t r = 0; r < n_reads; r++) {
			const uint64_t l = read_off[r + 1] - read_off[r];
			koff[r + 1] = koff[r] + (l >= k ? l - k + 1 : 0);
		}
	}
	return host_query_records(idx, bases + base0, read_off[n_reads] - base0, beg, nullptr, n_reads, ids_out ? koff.data() : nullptr, ids_out,
	                          ids_out ? koff[n_reads] : 0, ctr, true);
}

int blight_query_sequence_host(const blight_index* idx, const char* seq, uint64_t len, int64_t* ids_out, uint64_t* n_out) {
	if (!idx || !n_out || (len && !seq)) return fail(BL_ERR_INVALID_ARG, "null argument");
	const uint32_t k = idx->v.k;
	*n_out = len >= k ? len - k + 1 : 0;
	if (*n_out == 0) return BL_OK;  // query.size() < k: empty result (blight.cpp:577-579)
	if (!ids_out) return fail(BL_ERR_INVALID_ARG, "null argument");
	uint64_t off[2] = {0, len}, ctr[BLIGHT_N_CTR];
	return blight_query_reads_host(idx, seq, off, 1, ids_out, ctr);
}

int blight_query_sequence_bool_host(const blight_index* idx, const char* seq, uint64_t len, uint64_t* found, uint64_t* not_found) {
	if (!idx || !found || !not_found || (len && !seq)) return fail(BL_ERR_INVALID_ARG, "null argument");
	*found = *not_found = 0;
	if (len < idx->v.k) return BL_OK;  // blight.cpp:557-559
	uint64_t off[2] = {0, len}, ctr[BLIGHT_N_CTR];
	const int rc = blight_query_reads_host(idx, seq, off, 1, nullptr, ctr);
	*found = ctr[BLIGHT_CTR_FOUND];
	*not_found = ctr[BLIGHT_CTR_NOT_FOUND];
	return rc;
}

int blight_query_kmers_host(const blight_index* idx, const uint64_t* canon, uint64_t n, int64_t* ids_out) {
	if (!idx || (n && (!canon || !ids_out))) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (n == 0) return BL_OK;
	DeviceGuard guard(idx->device);
	Lease lease;
	int rc = lease.acquire(idx);
	if (rc != BL_OK) return rc;
	HostCtx& C = *lease.c;
	void *d_canon = nullptr, *d_ids = nullptr;
	if ((rc = C.reserve(7, n * 8, &d_canon)) != BL_OK) return rc;
	if ((rc = C.reserve(5, n * 8, &d_ids)) != BL_OK) return rc;
	CU(cudaMemcpyAsync(d_canon, canon, n * 8, cudaMemcpyHostToDevice, C.st_f));
	rc = launch_lookup_kmers(idx->v, static_cast<const uint64_t*>(d_canon), nullptr, n, static_cast<int64_t*>(d_ids), C.st_f);
	if (rc != BL_OK) return fail(rc, std::string("kernel launch failed: ") + g_last_cuda_error);
	CU(cudaMemcpyAsync(ids_out, d_ids, n * 8, cudaMemcpyDeviceToHost, C.st_f));
	CU(cudaStreamSynchronize(C.st_f));
	g_h2d_bytes += n * 8;
	g_d2h_bytes += n * 8;
	return BL_OK;
}

void blight_transfer_bytes(uint64_t* h2d, uint64_t* d2h) {
	if (h2d) *h2d = g_h2d_bytes.load();
	if (d2h) *d2h = g_d2h_bytes.load();
}

void blight_host_pack_stats(uint64_t* packed_bases, uint64_t* pack_ns, uint32_t* threads) {
	if (packed_bases) *packed_bases = g_packed_bases.load();
	if (pack_ns) *pack_ns = g_pack_ns.load();
	if (threads) *threads = (uint32_t)pack_threads();
}

}  // extern "C"
