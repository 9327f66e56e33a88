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
//
// The reader (ChunkReader) is shared by the one-GPU path below and by the several-GPU path of comm.cu (stream_fasta_chunks):
// parallel pread() straight into pinned buffers, each reader thread scanning the bytes it just read for newlines while they
// are still in its core's cache.
#include <cuda_runtime.h>
#include <omp.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <functional>
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

using clk = std::chrono::steady_clock;
double since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

template <class T>
class Channel {
	std::mutex m_;
	std::condition_variable cv_;
	std::deque<T> q_;
public:
	void push(T v) { { std::lock_guard<std::mutex> l(m_); q_.push_back(v); } cv_.notify_one(); }
	T pop() { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [&] { return !q_.empty(); }); T v = q_.front(); q_.pop_front(); return v; }
};

// newline offsets of a buffer's new data, one list per reader thread (found right after its pread, while the bytes are still
// in that core's cache: a separate scan of the 64 MB buffer ran at 26 GB/s whatever the team size — a DRAM pass). Kept from
// call to call so that the lists are allocated once.
struct PartsStore {
	std::vector<std::vector<uint64_t>> parts[kBufs];
	std::vector<size_t> n[kBufs];
};

size_t chunk_bytes() {
	size_t c = kChunkMax;
	if (const char* e = getenv("BLIGHT_STREAM_CHUNK_KB")) {
		const size_t kb = strtoull(e, nullptr, 10);
		if (kb) c = std::min(kChunkMax, kb << 10);
	}
	return c;
}

// Fills the buffers bufs[0..kBufs) (each kHead + kChunkMax + 64 bytes; the data starts at kHead) with consecutive pieces of
// the file on a thread of its own. A plain file is read with parallel pread() (zlib's pass-through copies at ~5 GB/s on one
// thread); gzip goes through gzread.
class ChunkReader {
public:
	struct Chunk { int buf; size_t len; bool eof, io_error, scanned; };
	// The host cores are shared out between the reader's pread team and the record cut's team (both OpenMP): together no more
	// threads than cores, or either team waits at its barriers for members the other one pushed off a core (measured with one
	// 16-thread cut team beside 8 readers: 3.3 ms of cut per 64 MB chunk, then the slowest stage of the whole path).
	ChunkReader(char* const* bufs, PartsStore* store) : bufs_(bufs), P_(store), chunk_(chunk_bytes()) {
		const int host_threads = std::max(2, std::max(omp_get_max_threads(), std::min(8, (int)std::thread::hardware_concurrency())));
		read_threads = std::max(1, std::min(24, host_threads * 3 / 4));
		cut_team = std::max(1, host_threads - read_threads);
	}
	~ChunkReader() { stop(); }
	int open(const char* path) {
		gz_ = gzopen(path, "rb");
		if (!gz_) return fail(BL_ERR_IO, std::string("Problem with files opening: ") + path);  // blight.cpp:188-189
		gzbuffer(gz_, 1 << 20);
		if (gzdirect(gz_)) {
			fd_ = ::open(path, O_RDONLY);
			struct stat sb;
			if (fd_ >= 0 && fstat(fd_, &sb) == 0 && S_ISREG(sb.st_mode)) file_size_ = uint64_t(sb.st_size);
			else { if (fd_ >= 0) close(fd_); fd_ = -1; }
		}
		for (int i = 0; i < kBufs; i++) free_.push(i);
		thread_ = std::thread([this] { run(); });
		started_ = true;
		return BL_OK;
	}
	Chunk next() { return filled_.pop(); }
	void release(int buf) { free_.push(buf); }
	const std::vector<uint64_t>* parts(int buf) const { return P_->parts[buf].data(); }
	const size_t* part_n(int buf) const { return P_->n[buf].data(); }
	int n_parts(int buf) const { return (int)P_->parts[buf].size(); }
	void stop() {
		if (started_) { free_.push(-1); thread_.join(); started_ = false; }
		if (gz_) { gzclose(gz_); gz_ = nullptr; }
		if (fd_ >= 0) { close(fd_); fd_ = -1; }
	}
	int read_threads = 1, cut_team = 1;
	double t_busy = 0, t_wait = 0;  // ms the reader spent reading / waiting for a free buffer

private:
	void run() {
		uint64_t file_off = 0, n_read = 0;
		for (;;) {
			const clk::time_point tw = clk::now();
			const int b = free_.pop();
			t_wait += since(tw);
			if (b < 0) return;
			const clk::time_point tb = clk::now();
			size_t got = 0;
			bool eof = false, bad = false;
			char* dst = bufs_[b] + kHead;
			if (fd_ >= 0) {
				// the first buffers are small (8 MB, doubling): the GPU starts on the file a millisecond after the call, not after a
				// whole 64 MB buffer has been read
				const size_t ramp = std::min<size_t>(chunk_, (size_t(8) << 20) << std::min<uint64_t>(n_read, 8));
				n_read++;
				const size_t want = size_t(std::min<uint64_t>(ramp, file_size_ - file_off));
				int failed = 0;
				if ((int)P_->parts[b].size() != read_threads) { P_->parts[b].assign(read_threads, {}); P_->n[b].assign(read_threads, 0); }
				const int T = read_threads;
				#pragma omp parallel for num_threads(T) schedule(static) reduction(+ : failed)
				for (int t = 0; t < T; t++) {
					size_t lo = want * t / T;
					const size_t hi = want * (t + 1) / T;
					size_t n = 0;
					while (lo < hi) {
						const ssize_t r = pread(fd_, dst + lo, std::min<size_t>(hi - lo, 1u << 20), off_t(file_off + lo));
						if (r <= 0) { failed++; break; }
						scan_newlines(dst, lo, lo + size_t(r), P_->parts[b][t], n);
						lo += size_t(r);
					}
					P_->n[b][t] = n;
				}
				bad = failed != 0;
				got = want;
				file_off += want;
				eof = file_off >= file_size_;
				t_busy += since(tb);
				filled_.push(Chunk{b, got, eof, bad, true});
			} else {
				while (got < chunk_) {
					const int r = gzread(gz_, dst + got, unsigned(std::min<size_t>(chunk_ - got, 1u << 30)));
					if (r < 0) { bad = true; break; }
					if (r == 0) { eof = true; break; }
					got += size_t(r);
				}
				t_busy += since(tb);
				filled_.push(Chunk{b, got, eof, bad, false});
			}
			if (eof || bad) return;
		}
	}
	char* const* bufs_;
	PartsStore* P_;
	size_t chunk_;
	gzFile gz_ = nullptr;
	int fd_ = -1;
	uint64_t file_size_ = 0;
	Channel<int> free_;
	Channel<Chunk> filled_;
	std::thread thread_;
	bool started_ = false;
};

// line ends of a chunk = carried head + new data, from the reader's own lists when it scanned the data itself
size_t chunk_lines(const ChunkReader& R, const ChunkReader::Chunk& f, const char* base, size_t head_len, size_t len, std::vector<uint64_t>& nl) {
	return f.scanned ? fasta_chunk_lines_merge(base, head_len, len, f.eof, R.parts(f.buf), R.part_n(f.buf), R.n_parts(f.buf), nl, R.cut_team)
	                 : fasta_chunk_lines(base, len, f.eof, nl, R.cut_team);
}

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
	PartsStore parts;
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

// Host side alone (several GPUs, comm.cu): pinned buffers any device of the process may copy from, and the reader's lists.
struct StreamHost {
	char* text[kBufs] = {nullptr, nullptr, nullptr};
	PartsStore parts;
	~StreamHost() { for (int i = 0; i < kBufs; i++) if (text[i]) cudaFreeHost(text[i]); }
};

void stream_host_free(void* p) { delete static_cast<StreamHost*>(p); }

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

// The file as a sequence of record batches in pinned host memory, for a consumer that is done with a batch when it returns
// (comm.cu: the records of a batch are shared out over the devices). *host: created on first use, kept by the caller
// (stream_host_free). on_batch(text, len, beg, end, n_rec): records [beg[i], end[i]) of text[0, len).
int stream_fasta_chunks(const char* path, void** host, const std::function<int(const char*, uint64_t, const uint64_t*, const uint64_t*, uint64_t)>& on_batch) {
	if (!*host) {
		StreamHost* h = new StreamHost();
		for (int i = 0; i < kBufs; i++) {
			cudaError_t e = cudaHostAlloc(reinterpret_cast<void**>(&h->text[i]), kHead + kChunkMax + 64, cudaHostAllocPortable);
			if (e != cudaSuccess) { delete h; return cu_fail(e, "cudaHostAlloc(streaming buffers)"); }
		}
		*host = h;
	}
	StreamHost& H = *static_cast<StreamHost*>(*host);
	ChunkReader R(H.text, &H.parts);
	int rc = R.open(path);
	if (rc != BL_OK) return rc;
	std::vector<char> carry;
	std::vector<uint64_t> nl, beg, end;
	for (;;) {
		const ChunkReader::Chunk f = R.next();
		if (f.io_error) return fail(BL_ERR_IO, std::string("read error: ") + path);
		if (carry.size() > kHead) return fail(BL_ERR_FORMAT, "a FASTA record is longer than the streaming buffer (8 MB)");
		char* base = H.text[f.buf] + kHead - carry.size();
		if (!carry.empty()) std::memcpy(base, carry.data(), carry.size());
		const size_t len = carry.size() + f.len;
		const size_t n_pairs = chunk_lines(R, f, base, carry.size(), len, nl);
		beg.resize(n_pairs + 1); end.resize(n_pairs + 1);
		const ChunkCut cut = fasta_chunk_records(len, f.eof, nl, beg.data(), end.data(), R.cut_team);
		beg[cut.n_rec] = cut.consumed;
		if (!f.eof) carry.assign(base + cut.consumed, base + len); else carry.clear();
		if (cut.n_rec) {
			rc = on_batch(base, cut.consumed, beg.data(), end.data(), cut.n_rec);
			if (rc != BL_OK) return rc;
		}
		R.release(f.buf);
		if (f.eof) break;
	}
	return BL_OK;
}

// file_query(path) without holding the file in memory. ctr[BLIGHT_N_CTR] as blight_query_fasta_host.
int stream_file_query(const blight_index* idx, const char* path, uint64_t* ctr) {
	std::memset(ctr, 0, sizeof(uint64_t) * BLIGHT_N_CTR);
	int prev_dev = -1;
	cudaGetDevice(&prev_dev);
	if (prev_dev != idx->device) cudaSetDevice(idx->device);
	struct Restore { int d, cur; ~Restore() { if (d >= 0 && d != cur) cudaSetDevice(d); } } restore{prev_dev, idx->device};
	std::lock_guard<std::mutex> lock(*static_cast<std::mutex*>(idx->host_mutex));
	blight_index* mi = const_cast<blight_index*>(idx);
	if (!mi->stream_ctx) {
		StreamCtx* c = new StreamCtx();
		const int rc = ctx_init(*c);
		if (rc != BL_OK) { delete c; return rc; }
		mi->stream_ctx = c;
	}
	StreamCtx& C = *static_cast<StreamCtx*>(mi->stream_ctx);
	cudaStream_t st = static_cast<cudaStream_t>(idx->host_stream), cs = static_cast<cudaStream_t>(idx->copy_stream);
	// BLIGHT_STREAM_TRACE=1 (diagnostic): where the wall clock of the call went, per stage, on stderr
	const bool trace = getenv("BLIGHT_STREAM_TRACE") != nullptr;
	double t_wait_filled = 0, t_wait_done = 0, t_lines = 0, t_records = 0, t_carry = 0, t_enqueue = 0, t_wait_copied = 0;
	const clk::time_point t_call = clk::now();
	ChunkReader R(C.text, &C.parts);
	int rc = R.open(path);
	if (rc != BL_OK) return rc;
	cudaError_t ce = cudaMemsetAsync(C.d_ctr, 0, BLIGHT_N_CTR * 8, st);
	if (ce != cudaSuccess) return cu_fail(ce, "cudaMemsetAsync");
	std::vector<char> carry;
	std::vector<uint64_t> nl;
	int in_flight = -1;  // buffer whose H2D may still be running
	for (uint64_t i = 0;; i++) {
		clk::time_point tp = clk::now();
		const ChunkReader::Chunk f = R.next();
		t_wait_filled += since(tp);
		if (f.io_error) { rc = fail(BL_ERR_IO, std::string("read error: ") + path); break; }
		if (carry.size() > kHead) { rc = fail(BL_ERR_FORMAT, "a FASTA record is longer than the streaming buffer (8 MB)"); break; }
		tp = clk::now();
		char* base = C.text[f.buf] + kHead - carry.size();
		if (!carry.empty()) std::memcpy(base, carry.data(), carry.size());
		t_carry += since(tp);
		const size_t len = carry.size() + f.len;
		const int s = int(i & 1);
		// the offsets of chunk i-2 must have left the pinned staging area, its kernel must be done with d_text[s]
		tp = clk::now();
		if ((ce = cudaEventSynchronize(C.done[s])) != cudaSuccess) { rc = cu_fail(ce, "cudaEventSynchronize"); break; }
		t_wait_done += since(tp);
		tp = clk::now();
		const size_t n_pairs = chunk_lines(R, f, base, carry.size(), len, nl);
		t_lines += since(tp);
		tp = clk::now();
		if ((rc = reserve_offsets(C, s, n_pairs)) != BL_OK) break;
		// beg[0..n_rec], then end[0..n_rec) right behind it: one H2D copy
		uint64_t* beg = C.h_off[s];
		uint64_t* end_tmp = C.h_off[s] + n_pairs + 1;
		const ChunkCut cut = fasta_chunk_records(len, f.eof, nl, beg, end_tmp, R.cut_team);
		const size_t n_rec = cut.n_rec, consumed = cut.consumed;
		beg[n_rec] = len;
		std::memmove(C.h_off[s] + n_rec + 1, end_tmp, n_rec * 8);
		if (!f.eof) carry.assign(base + consumed, base + len); else carry.clear();
		t_records += since(tp);
		tp = clk::now();
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
		t_enqueue += since(tp);
		tp = clk::now();
		// hand the previous buffer back to the reader once its copy has left host memory
		if (in_flight >= 0) {
			cudaEventSynchronize(C.copied[in_flight]);
			R.release(in_flight);
		}
		t_wait_copied += since(tp);
		in_flight = f.buf;
		if (f.eof) break;
	}
	const double t_loop = since(t_call);
	if (rc == BL_OK) {
		ce = cudaMemcpyAsync(ctr, C.d_ctr, BLIGHT_N_CTR * 8, cudaMemcpyDeviceToHost, st);
		if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
		if (ce != cudaSuccess) rc = cu_fail(ce, "file_query: final synchronize");
	}
	cudaStreamSynchronize(cs);
	cudaStreamSynchronize(st);
	R.stop();
	if (trace)
		fprintf(stderr, "{\"stream_trace_ms\": {\"total\": %.2f, \"loop\": %.2f, \"reader_busy\": %.2f, \"reader_waits_for_buffer\": %.2f, \"main_waits_for_reader\": %.2f, "
		        "\"main_waits_for_kernel\": %.2f, \"lines\": %.2f, \"records\": %.2f, \"carry\": %.2f, \"enqueue\": %.2f, \"main_waits_for_copy\": %.2f, \"read_threads\": %d, \"cut_threads\": %d}}\n",
		        since(t_call), t_loop, R.t_busy, R.t_wait, t_wait_filled, t_wait_done, t_lines, t_records, t_carry, t_enqueue, t_wait_copied, R.read_threads, R.cut_team);
	if (rc == BL_OK && ctr[BLIGHT_CTR_INVALID]) return fail(BL_ERR_INVALID_BASE, "Invalid char in DNA");
	return rc;
}

}  // namespace blight
