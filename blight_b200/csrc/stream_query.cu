// stream_query.cu — streaming front end of file_query (kmer_Set_Light::file_query, blight.cpp:746-799).
//
// The reference reads the query file with getline() under an OpenMP critical section, 512 records at a time
// (blight.cpp:751-775). Here a reader thread fills pinned 64 MB buffers (gzread: plain or gzip, like zstr::ifstream,
// zstr.hpp:136-209), the calling thread cuts each buffer into 2-line records with all host cores, and the GPU works
// two buffers behind: H2D of chunk i+1 on the copy stream overlaps the read kernel of chunk i, the file is never held
// in memory as a whole.
//
// Record pairing is the reference's (blight.cpp:760-772, restated in flat_index.cpp: split_fasta_records): every
// iteration consumes exactly two lines — a header line (whatever it holds) and the line after it; the second line is a
// query sequence unless it, or the header, is empty. So records are the line pairs (2j, 2j+1) counted from the start of
// the file, which makes the cut parallel: find the newlines, pair them up. A chunk ends at an even line boundary; the
// unfinished tail is carried in front of the next buffer.
#include <cuda_runtime.h>
#include <omp.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <condition_variable>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include "capi_common.hpp"
#include "device_index.hpp"
#include "kernels.hpp"

namespace blight {
namespace {

constexpr size_t kChunkMax = 64ull << 20;  // bytes read per buffer (BLIGHT_STREAM_CHUNK_KB lowers it: tests)
constexpr size_t kHead = 8ull << 20;    // room in front of a buffer for the unfinished record of the previous one
constexpr int kBufs = 3;

struct Filled { int buf; size_t len; bool eof; bool io_error; };

template <class T>
class Channel {
	std::mutex m_;
	std::condition_variable cv_;
	std::deque<T> q_;
public:
	void push(T v) { { std::lock_guard<std::mutex> l(m_); q_.push_back(v); } cv_.notify_one(); }
	T pop() { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return !q_.empty(); }); T v = q_.front(); q_.pop_front(); return v; }
};

}  // namespace

// Pinned and device buffers of the streaming path, created on first use and kept with the index.
struct StreamCtx {
	char* text[kBufs] = {nullptr, nullptr, nullptr};
	cudaEvent_t copied[kBufs] = {nullptr, nullptr, nullptr};
	uint64_t* h_off[2] = {nullptr, nullptr};  // pinned: beg[0..n], then end[0..n-1]
	size_t h_off_cap[2] = {0, 0};
	char* d_text[2] = {nullptr, nullptr};
	uint64_t* d_off[2] = {nullptr, nullptr};
	size_t d_off_cap[2] = {0, 0};
	uint64_t* d_ctr = nullptr;
	cudaEvent_t done[2] = {nullptr, nullptr};
	cudaEvent_t ev_copy = nullptr;
	~StreamCtx() {
		for (int i = 0; i < kBufs; i++) { if (text[i]) cudaFreeHost(text[i]); if (copied[i]) cudaEventDestroy(copied[i]); }
		for (int s = 0; s < 2; s++) {
			if (h_off[s]) cudaFreeHost(h_off[s]);
			if (d_text[s]) cudaFree(d_text[s]);
			if (d_off[s]) cudaFree(d_off[s]);
			if (done[s]) cudaEventDestroy(done[s]);
		}
		if (d_ctr) cudaFree(d_ctr);
		if (ev_copy) cudaEventDestroy(ev_copy);
	}
};

void stream_ctx_free(void* p) { delete static_cast<StreamCtx*>(p); }

namespace {

int cu_fail(cudaError_t e, const char* what) { return fail(BL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); }
#define SCU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cu_fail(e__, #call); } while (0)

int ctx_init(StreamCtx& c) {
	for (int i = 0; i < kBufs; i++) {
		SCU(cudaHostAlloc(reinterpret_cast<void**>(&c.text[i]), kHead + kChunkMax + 64, cudaHostAllocDefault));
		SCU(cudaEventCreateWithFlags(&c.copied[i], cudaEventDisableTiming));
	}
	for (int s = 0; s < 2; s++) {
		SCU(cudaMalloc(reinterpret_cast<void**>(&c.d_text[s]), kHead + kChunkMax + 64));
		SCU(cudaEventCreateWithFlags(&c.done[s], cudaEventDisableTiming));
	}
	SCU(cudaMalloc(reinterpret_cast<void**>(&c.d_ctr), BLIGHT_N_CTR * 8));
	SCU(cudaEventCreateWithFlags(&c.ev_copy, cudaEventDisableTiming));
	return BL_OK;
}

int reserve_offsets(StreamCtx& c, int s, size_t n_rec) {
	const size_t need = 2 * n_rec + 2;
	if (c.h_off_cap[s] < need) {
		if (c.h_off[s]) cudaFreeHost(c.h_off[s]);
		c.h_off[s] = nullptr; c.h_off_cap[s] = 0;
		const size_t cap = need + need / 4 + 1024;
		SCU(cudaHostAlloc(reinterpret_cast<void**>(&c.h_off[s]), cap * 8, cudaHostAllocDefault));
		c.h_off_cap[s] = cap;
	}
	if (c.d_off_cap[s] < need) {
		if (c.d_off[s]) cudaFree(c.d_off[s]);
		c.d_off[s] = nullptr; c.d_off_cap[s] = 0;
		const size_t cap = need + need / 4 + 1024;
		SCU(cudaMalloc(reinterpret_cast<void**>(&c.d_off[s]), cap * 8));
		c.d_off_cap[s] = cap;
	}
	return BL_OK;
}

}  // namespace

// file_query(path) without holding the file in memory. ctr[BLIGHT_N_CTR] as blight_query_fasta_host.
int stream_file_query(const blight_index* idx, const char* path, uint64_t* ctr) {
	std::memset(ctr, 0, sizeof(uint64_t) * BLIGHT_N_CTR);
	gzFile gz = gzopen(path, "rb");
	if (!gz) return fail(BL_ERR_IO, std::string("Problem with files opening: ") + path);  // blight.cpp:188-189
	gzbuffer(gz, 1 << 20);
	int prev_dev = -1;
	cudaGetDevice(&prev_dev);
	if (prev_dev != idx->device) cudaSetDevice(idx->device);
	struct Restore { int d, cur; ~Restore() { if (d >= 0 && d != cur) cudaSetDevice(d); } } restore{prev_dev, idx->device};
	std::lock_guard<std::mutex> lock(*static_cast<std::mutex*>(idx->host_mutex));
	blight_index* mi = const_cast<blight_index*>(idx);
	if (!mi->stream_ctx) {
		StreamCtx* c = new StreamCtx();
		const int rc = ctx_init(*c);
		if (rc != BL_OK) { delete c; gzclose(gz); return rc; }
		mi->stream_ctx = c;
	}
	StreamCtx& C = *static_cast<StreamCtx*>(mi->stream_ctx);
	cudaStream_t st = static_cast<cudaStream_t>(idx->host_stream), cs = static_cast<cudaStream_t>(idx->copy_stream);

	size_t kChunk = kChunkMax;
	if (const char* e = getenv("BLIGHT_STREAM_CHUNK_KB")) {
		const size_t kb = strtoull(e, nullptr, 10);
		if (kb) kChunk = std::min(kChunkMax, kb << 10);
	}
	Channel<int> free_bufs;
	Channel<Filled> filled;
	for (int i = 0; i < kBufs; i++) free_bufs.push(i);
	// a plain file is read with parallel pread() straight into the pinned buffer (zlib's pass-through copies at ~5 GB/s
	// on one thread); gzip goes through gzread
	int fd = -1;
	uint64_t file_size = 0, file_off = 0;
	if (gzdirect(gz)) {
		fd = open(path, O_RDONLY);
		struct stat sb;
		if (fd >= 0 && fstat(fd, &sb) == 0 && S_ISREG(sb.st_mode)) file_size = uint64_t(sb.st_size);
		else { if (fd >= 0) close(fd); fd = -1; }
	}
	const int read_threads = std::max(1, std::min(8, omp_get_max_threads() / 2));
	std::thread reader([&] {
		for (;;) {
			const int b = free_bufs.pop();
			if (b < 0) return;
			size_t got = 0;
			bool eof = false, bad = false;
			if (fd >= 0) {
				const size_t want = size_t(std::min<uint64_t>(kChunk, file_size - file_off));
				char* dst = C.text[b] + kHead;
				int failed = 0;
				#pragma omp parallel for num_threads(read_threads) schedule(static) reduction(+ : failed)
				for (int t = 0; t < read_threads; t++) {
					size_t lo = want * t / read_threads;
					const size_t hi = want * (t + 1) / read_threads;
					while (lo < hi) {
						const ssize_t r = pread(fd, dst + lo, hi - lo, off_t(file_off + lo));
						if (r <= 0) { failed++; break; }
						lo += size_t(r);
					}
				}
				bad = failed != 0;
				got = want;
				file_off += want;
				eof = file_off >= file_size;
				filled.push(Filled{b, got, eof, bad});
				if (eof || bad) return;
				continue;
			}
			while (got < kChunk) {
				const int r = gzread(gz, C.text[b] + kHead + got, unsigned(std::min<size_t>(kChunk - got, 1u << 30)));
				if (r < 0) { bad = true; break; }
				if (r == 0) { eof = true; break; }
				got += size_t(r);
			}
			filled.push(Filled{b, got, eof, bad});
			if (eof || bad) return;
		}
	});
	auto stop_reader = [&] { free_bufs.push(-1); reader.join(); gzclose(gz); if (fd >= 0) close(fd); };

	int rc = BL_OK;
	cudaError_t ce = cudaMemsetAsync(C.d_ctr, 0, BLIGHT_N_CTR * 8, st);
	if (ce != cudaSuccess) { stop_reader(); return cu_fail(ce, "cudaMemsetAsync"); }
	std::vector<char> carry;
	std::vector<uint64_t> nl;
	int in_flight = -1;  // buffer whose H2D may still be running
	for (uint64_t i = 0;; i++) {
		const Filled f = filled.pop();
		if (f.io_error) { rc = fail(BL_ERR_IO, std::string("read error: ") + path); break; }
		if (carry.size() > kHead) { rc = fail(BL_ERR_FORMAT, "a FASTA record is longer than the streaming buffer (8 MB)"); break; }
		char* base = C.text[f.buf] + kHead - carry.size();
		if (!carry.empty()) std::memcpy(base, carry.data(), carry.size());
		const size_t len = carry.size() + f.len;
		const int s = int(i & 1);
		// the offsets of chunk i-2 must have left the pinned staging area, its kernel must be done with d_text[s]
		if ((ce = cudaEventSynchronize(C.done[s])) != cudaSuccess) { rc = cu_fail(ce, "cudaEventSynchronize"); break; }
		const size_t n_pairs = fasta_chunk_lines(base, len, f.eof, nl);
		if ((rc = reserve_offsets(C, s, n_pairs)) != BL_OK) break;
		// beg[0..n_rec], then end[0..n_rec) right behind it: one H2D copy
		uint64_t* beg = C.h_off[s];
		uint64_t* end_tmp = C.h_off[s] + n_pairs + 1;
		const ChunkCut cut = fasta_chunk_records(len, f.eof, nl, beg, end_tmp);
		const size_t n_rec = cut.n_rec, consumed = cut.consumed;
		beg[n_rec] = len;
		std::memmove(C.h_off[s] + n_rec + 1, end_tmp, n_rec * 8);
		if (!f.eof) carry.assign(base + consumed, base + len); else carry.clear();
		if (n_rec) {
			// H2D on the copy stream (after the kernel that last read this device buffer), kernel on the query stream
			if ((ce = cudaStreamWaitEvent(cs, C.done[s], 0)) != cudaSuccess) { rc = cu_fail(ce, "cudaStreamWaitEvent"); break; }
			if ((ce = cudaMemcpyAsync(C.d_text[s], base, consumed, cudaMemcpyHostToDevice, cs)) != cudaSuccess) { rc = cu_fail(ce, "cudaMemcpyAsync(text)"); break; }
			if ((ce = cudaMemcpyAsync(C.d_off[s], C.h_off[s], (2 * n_rec + 1) * 8, cudaMemcpyHostToDevice, cs)) != cudaSuccess) { rc = cu_fail(ce, "cudaMemcpyAsync(offsets)"); break; }
			cudaEventRecord(C.copied[f.buf], cs);
			cudaEventRecord(C.ev_copy, cs);
			cudaStreamWaitEvent(st, C.ev_copy, 0);
			ReadBatch B;
			B.d_bases = C.d_text[s]; B.d_read_off = C.d_off[s]; B.d_read_end = C.d_off[s] + n_rec + 1; B.n_reads = n_rec; B.total_bases = consumed;
			rc = launch_reads(&idx->v, idx->v.k, idx->v.m, B, nullptr, nullptr, nullptr, C.d_ctr, st);
			if (rc != BL_OK) { rc = fail(rc, std::string("kernel launch failed: ") + g_last_cuda_error); break; }
			cudaEventRecord(C.done[s], st);
		} else {
			cudaEventRecord(C.copied[f.buf], cs);
		}
		// hand the previous buffer back to the reader once its copy has left host memory
		if (in_flight >= 0) {
			cudaEventSynchronize(C.copied[in_flight]);
			free_bufs.push(in_flight);
		}
		in_flight = f.buf;
		if (f.eof) break;
	}
	if (rc == BL_OK) {
		ce = cudaMemcpyAsync(ctr, C.d_ctr, BLIGHT_N_CTR * 8, cudaMemcpyDeviceToHost, st);
		if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
		if (ce != cudaSuccess) rc = cu_fail(ce, "file_query: final synchronize");
	}
	cudaStreamSynchronize(cs);
	cudaStreamSynchronize(st);
	stop_reader();
	if (rc == BL_OK && ctr[BLIGHT_CTR_INVALID]) return fail(BL_ERR_INVALID_BASE, "Invalid char in DNA");
	return rc;
}

}  // namespace blight
