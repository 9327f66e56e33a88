// part_session.cu — one rank of the bucket-partitioned query path (BASELINE configs[4]) as an object behind
// the C ABI: its peer-visible buffers, the connection to the other ranks' buffers, and the per-batch pipeline.
//
// The kernels are those of part_kernels.cu (k_dispatch_runs at the GPU holding the reads, k_runs_lookup at the owner of the
// minimizer bucket; records and identifiers cross NVLink as peer-memory stores issued by the kernels themselves). What this
// file adds is the ordering BETWEEN GPUs, without any collective library call on the data path:
//
//   k_publish_counts   (source, after its dispatch kernel)  stores, into every owner's mailbox, how many records it left in
//                      that owner's inbox, fences at system scope, then stores the sub-batch's sequence number
//   k_wait_counts      (owner, before its lookup kernel)    spins until every source's sequence number has arrived and
//                      copies the counts next to the lookup kernel's arguments
//
// Order on every rank, sub-batch i using inbox half i & 1 (blight_part_config.order):
//   serial    D(i)  P(i)  W(i)  L(i)                          one stream
//   ahead     D(i+1)  L(i)  P(i+1)  W(i+1)                    one stream: the wait never sees the dispatch skew
//   overlap   D(i+1) on the caller's stream BESIDE L(i) on a second one, each limited to half of every SM's CTA slots (the
//             front end is arithmetic, the lookups are memory latency: together they fill the SM as the one-GPU kernel does);
//             then P(i+1)  W(i+1)
// A source rewrites half b in D(i+2), which follows its own W(i+1); W(i+1) needs every rank's P(i+1), which that rank
// issued after its L(i): nobody is still reading the half. No deadlock: every wait depends only on kernels that precede
// the matching publish in the publisher's own streams.
// Identifiers return one of two ways (blight_part_config.return_path):
//   stream    an owner's warp stores its answers as ONE contiguous run of 32-bit slice-local ids into its return
//             region at the source; the source widens them into its int64 id array in read order through a local side table
//             (k_scatter_runs), one sub-batch behind, on its own stream. S(i) may start once W(i+1) has passed: every owner
//             published P(i+1) after its L(i).
//   pull      as stream, but the owner's warp stores its run of ids into its OWN memory (region [half][source]) and the
//             source's scatter pass fetches them over NVLink with 16-byte loads, all in flight before the first is used.
//             The lookup kernel then issues no remote store at all: on 8 B200s the pushed streams cost the lookups ~5 ms per
//             480 M k-mers (remote stores hold the issuing warps' memory pipeline; the loads of the scatter pass stall
//             nobody but the scatter pass, which runs beside the next sub-batch's lookups).
//   direct    (default) the owner stores int64 ids straight into the source's id array, run by run: no second pass, no side
//             table, no return regions. The stores are ~100-byte segments scattered over a multi-GB array; what made them
//             slow in the first measurements (38.8 ms per batch of 480 M k-mers on 8 B200s) was not their size but that all
//             owners, running in step, stored to the SAME source at any moment. With the round-robin chunk order of
//             k_runs_lookup: 18.0 ms (stream 18.9, pull 22.3 before the on-demand work distribution), counting 14.3.
// The ranks are processes (torchrun: buffers exchanged as CUDA IPC handles) or devices of one process
// (blight_comm, comm.cu: peer access).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "capi_common.hpp"
#include "device_index.hpp"
#include "kernels.hpp"

using namespace blight;

namespace {

constexpr int kMaxRanks = BLIGHT_MAX_RANKS;
constexpr int kMailSlots = 3;  // two inbox halves + the end-of-batch fence

// Written by the peers, read by the owner. One per rank, inside its mailbox allocation.
struct Mailbox {
	unsigned long long count[kMailSlots][kMaxRanks];  // packed (records << 40 | k-mers) source s left for this owner
	unsigned long long seq[kMailSlots][kMaxRanks];    // sequence number of the sub-batch those counts belong to
};

struct MailPtrs { Mailbox* m[kMaxRanks]; };

__global__ void k_publish_counts(MailPtrs peers, const unsigned long long* __restrict__ counts, uint32_t world, uint32_t rank, uint32_t slot,
                                 unsigned long long seq) {
	const uint32_t d = threadIdx.x;
	if (d >= world) return;
	Mailbox* mb = peers.m[d];
	mb->count[slot][rank] = counts ? counts[d] : 0ull;
	__threadfence_system();  // the records of the dispatch kernel (earlier in this stream) and the count, before the flag
	asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(&mb->seq[slot][rank]), "l"(seq) : "memory");
}

__global__ void k_wait_counts(const Mailbox* mb, uint32_t world, uint32_t slot, unsigned long long seq, unsigned long long* __restrict__ rcv,
                              uint32_t* __restrict__ err, long long spin_limit) {
	const uint32_t s = threadIdx.x;
	if (s >= world) return;
	const long long t0 = clock64();
	unsigned long long got;
	for (;;) {
		asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(got) : "l"(&mb->seq[slot][s]) : "memory");
		if (got >= seq) break;
		if (clock64() - t0 > spin_limit) { atomicOr(err, 2u); if (rcv) rcv[s] = 0; return; }  // a peer never arrived: give up loudly
		__nanosleep(200);
	}
	if (rcv) rcv[s] = mb->count[slot][s];
	if (rcv && s == 0) rcv[kMaxRanks] = 0;  // the lookup kernel's ticket counter
}

// Where a batch of `total` base positions is cut into sub-batches: pieces of `full`. (Measured and rejected: pieces that shrink
// geometrically towards the end of the batch so that less of the last scatter pass is left exposed — 18.27 vs 18.10 ms per
// 480 M k-mers in loop-back at 192 M positions per sub-batch; the extra barriers cost more than the exposure.)
std::vector<uint64_t> sub_batch_cuts(uint64_t total, uint64_t full) {
	std::vector<uint64_t> cuts{0};
	for (uint64_t at = 0; at < total;) {
		at += std::min(full, total - at);
		cuts.push_back(at);
	}
	return cuts;
}

int cu_fail(cudaError_t e, const char* what) { return fail(BL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e)); }
#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) return cu_fail(e__, #call); } while (0)

struct Guard {
	int prev = -1;
	explicit Guard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
	~Guard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

struct blight_part_session {
	const blight_index* idx = nullptr;
	int device = 0;
	blight_part_config cfg{};
	uint64_t region_bytes = 0;
	uint64_t kcap = 0;  // ids per (owner, sub-batch) return region
	// this rank's peer-visible buffers
	void* inbox = nullptr;     // [2][world sources][cap] records
	Mailbox* mail = nullptr;
	int64_t* ids = nullptr;    // this rank's id array, ids_capacity entries (direct return: the owners write into it)
	uint32_t* ret = nullptr;   // stream return: [2][world owners][sub_positions] 32-bit ids, written by the owners
	uint4* side = nullptr;     // stream return, local: [2][world owners][cap] where each record's ids go
	// the other ranks' buffers as seen from here
	void* p_inbox[kMaxRanks] = {};
	Mailbox* p_mail[kMaxRanks] = {};
	int64_t* p_ids[kMaxRanks] = {};
	uint32_t* p_ret[kMaxRanks] = {};
	uint64_t p_ids_cap[kMaxRanks] = {};
	uint64_t id_base[kMaxRanks] = {};  // first identifier of every rank's slice
	bool ipc_opened[kMaxRanks] = {};
	uint32_t connected = 0;
	// local scratch
	unsigned long long* counts[2] = {nullptr, nullptr};  // as a source: packed counters per owner
	unsigned long long* rcv[2] = {nullptr, nullptr};     // as an owner: counters received per source
	uint32_t* err = nullptr;
	unsigned long long seq = 0;  // sub-batches issued so far (same on every rank: the calls are collective)
	uint32_t order = BLIGHT_PART_ORDER_SERIAL;
	bool stream_ret = true;  // 32-bit id streams + scatter pass (stream and pull)
	bool pull = false;       // ... fetched by the source from the owner's memory instead of pushed by the owner
	bool tickets = true;     // dispatch / lookup kernels hand their work items out on demand (BLIGHT_PART_TICKETS=0: fixed stride)
	int split_lookup = 2, split_dispatch = 2;  // overlap order: resident CTAs per SM of either kernel while both run
	cudaStream_t side_st = nullptr;  // overlap order: the lookups' stream
	cudaStream_t scat_st = nullptr;  // stream return: the scatter pass
	cudaEvent_t ev_main = nullptr, ev_side = nullptr, ev_pub = nullptr, ev_scat[2] = {nullptr, nullptr};
};

extern "C" {

int blight_part_session_create(const blight_index* idx, const blight_part_config* cfg, blight_part_session** out) {
	if (!idx || !cfg || !out) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (cfg->world == 0 || cfg->world > (uint32_t)kMaxRanks || cfg->rank >= cfg->world) return fail(BL_ERR_INVALID_ARG, "bad world / rank");
	if (cfg->cap == 0 || cfg->cap >= (1ull << 24)) return fail(BL_ERR_INVALID_ARG, "cap: 1 .. 2^24-1 records per region");
	if (cfg->sub_positions == 0 || (cfg->sub_positions % kReadsStrip) != 0 || cfg->sub_positions >= (1ull << 32))
		return fail(BL_ERR_INVALID_ARG, "sub_positions: a multiple of 256 below 2^32");
	for (uint32_t i = 0; i < cfg->world; i++)
		if (cfg->cuts[i] > cfg->cuts[i + 1]) return fail(BL_ERR_INVALID_ARG, "cuts must ascend");
	if (cfg->ids_capacity && !idx->v.pos_id) return fail(BL_ERR_INVALID_ARG, "the id mode of the partitioned path needs the position->id table");
	Guard g(idx->device);
	{
		int rc = part_kernels_preload();
		cudaFuncAttributes fa;
		if (rc == BL_OK && (cudaFuncGetAttributes(&fa, k_publish_counts) != cudaSuccess || cudaFuncGetAttributes(&fa, k_wait_counts) != cudaSuccess))
			rc = fail(BL_ERR_CUDA, "cannot load the flag kernels");
		if (rc != BL_OK) return rc;
	}
	blight_part_session* s = new blight_part_session();
	s->idx = idx; s->device = idx->device; s->cfg = *cfg;
	s->region_bytes = cfg->cap * BLIGHT_RUN_RECORD_BYTES;
	s->kcap = cfg->ret_kmers && cfg->ret_kmers < cfg->sub_positions ? cfg->ret_kmers : cfg->sub_positions;
	s->kcap = (s->kcap + 255) & ~uint64_t(255);  // region bases stay 16-byte aligned for the vector stores of the return stream
	s->order = cfg->order;
	if (s->order == BLIGHT_PART_ORDER_DEFAULT) {
		s->order = BLIGHT_PART_ORDER_SERIAL;
		if (const char* e = getenv("BLIGHT_PART_ORDER")) s->order = e[0] == 'a' ? BLIGHT_PART_ORDER_AHEAD : (e[0] == 'o' ? BLIGHT_PART_ORDER_OVERLAP : BLIGHT_PART_ORDER_SERIAL);
	}
	if (s->order > BLIGHT_PART_ORDER_OVERLAP) { delete s; return fail(BL_ERR_INVALID_ARG, "unknown order"); }
	if (const char* e = getenv("BLIGHT_PART_SPLIT")) {  // "L,D": CTAs per SM of the lookup and the dispatch kernel in the overlap order
		int a = 0, bb = 0;
		if (sscanf(e, "%d,%d", &a, &bb) == 2 && a > 0 && bb > 0) { s->split_lookup = a; s->split_dispatch = bb; }
	}
	uint32_t rp = cfg->return_path;
	if (rp == BLIGHT_PART_RETURN_DEFAULT) {
		rp = BLIGHT_PART_RETURN_DIRECT;
		if (const char* e = getenv("BLIGHT_PART_RETURN"))
			rp = e[0] == 'd' ? BLIGHT_PART_RETURN_DIRECT : (e[0] == 'p' ? BLIGHT_PART_RETURN_PULL : BLIGHT_PART_RETURN_STREAM);
	}
	if (rp != BLIGHT_PART_RETURN_STREAM && rp != BLIGHT_PART_RETURN_DIRECT && rp != BLIGHT_PART_RETURN_PULL) { delete s; return fail(BL_ERR_INVALID_ARG, "unknown return path"); }
	if (const char* e = getenv("BLIGHT_PART_TICKETS")) s->tickets = e[0] != '0';
	s->stream_ret = rp != BLIGHT_PART_RETURN_DIRECT;
	s->pull = rp == BLIGHT_PART_RETURN_PULL;
	const size_t inbox_bytes = (size_t)2 * cfg->world * s->region_bytes;
	cudaError_t e = cudaMalloc(&s->inbox, inbox_bytes);
	if (e == cudaSuccess) e = cudaMemset(s->inbox, 0, inbox_bytes);  // a slot never written must still parse as a (harmless) record
	if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->mail), sizeof(Mailbox));
	if (e == cudaSuccess) e = cudaMemset(s->mail, 0, sizeof(Mailbox));
	if (e == cudaSuccess && cfg->ids_capacity) e = cudaMalloc(reinterpret_cast<void**>(&s->ids), (size_t)cfg->ids_capacity * 8);
	if (e == cudaSuccess && cfg->ids_capacity && s->stream_ret) {
		e = cudaMalloc(reinterpret_cast<void**>(&s->ret), (size_t)2 * cfg->world * s->kcap * 4);
		if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->side), (size_t)2 * cfg->world * cfg->cap * sizeof(uint4));
	}
	for (int b = 0; b < 2 && e == cudaSuccess; b++) {
		// one slot more than ranks: the ticket counter of the dispatch (counts) and of the lookup kernel (rcv) of this half
		e = cudaMalloc(reinterpret_cast<void**>(&s->counts[b]), (kMaxRanks + 1) * 8);
		if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->rcv[b]), (kMaxRanks + 1) * 8);
		if (e == cudaSuccess) e = cudaMemset(s->rcv[b], 0, (kMaxRanks + 1) * 8);
	}
	if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&s->err), 4);
	if (e == cudaSuccess) e = cudaMemset(s->err, 0, 4);
	// extra streams only for who needs them: several ranks on ONE device (tests) share its few hardware queues, and a
	// spinning wait kernel queued in front of another rank's dispatch would never see its flag
	if (e == cudaSuccess && s->order == BLIGHT_PART_ORDER_OVERLAP) e = cudaStreamCreateWithFlags(&s->side_st, cudaStreamNonBlocking);
	if (e == cudaSuccess && s->stream_ret && cfg->ids_capacity) e = cudaStreamCreateWithFlags(&s->scat_st, cudaStreamNonBlocking);
	for (cudaEvent_t* ev : {&s->ev_main, &s->ev_side, &s->ev_pub, &s->ev_scat[0], &s->ev_scat[1]})
		if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
	if (e != cudaSuccess) { blight_part_session_free(s); return cu_fail(e, "partition session buffers"); }
	// a rank is its own peer
	const uint32_t r = cfg->rank;
	s->p_inbox[r] = s->inbox; s->p_mail[r] = s->mail; s->p_ids[r] = s->ids; s->p_ids_cap[r] = cfg->ids_capacity; s->p_ret[r] = s->ret;
	s->id_base[r] = idx->info.id_base;
	s->connected = 1u << r;
	*out = s;
	return BL_OK;
}

void blight_part_session_free(blight_part_session* s) {
	if (!s) return;
	Guard g(s->device);
	cudaDeviceSynchronize();
	for (uint32_t r = 0; r < (uint32_t)kMaxRanks; r++)
		if (s->ipc_opened[r]) {
			cudaIpcCloseMemHandle(s->p_inbox[r]);
			cudaIpcCloseMemHandle(s->p_mail[r]);
			if (s->p_ids[r]) cudaIpcCloseMemHandle(s->p_ids[r]);
			if (s->p_ret[r]) cudaIpcCloseMemHandle(s->p_ret[r]);
		}
	cudaFree(s->inbox); cudaFree(s->mail); cudaFree(s->ids); cudaFree(s->ret); cudaFree(s->side);
	for (int b = 0; b < 2; b++) { cudaFree(s->counts[b]); cudaFree(s->rcv[b]); }
	cudaFree(s->err);
	if (s->side_st) cudaStreamDestroy(s->side_st);
	if (s->scat_st) cudaStreamDestroy(s->scat_st);
	for (cudaEvent_t ev : {s->ev_main, s->ev_side, s->ev_pub, s->ev_scat[0], s->ev_scat[1]}) if (ev) cudaEventDestroy(ev);
	delete s;
}

int blight_part_session_handles(const blight_part_session* s, unsigned char* handles256) {
	unsigned char* handles192 = handles256;
	if (!s || !handles256) return fail(BL_ERR_INVALID_ARG, "null argument");
	Guard g(s->device);
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
	cudaIpcMemHandle_t h;
	std::memset(handles256, 0, 256);
	CU(cudaIpcGetMemHandle(&h, s->inbox));
	std::memcpy(handles192, &h, 64);
	CU(cudaIpcGetMemHandle(&h, s->mail));
	std::memcpy(handles192 + 64, &h, 64);
	if (s->ids) {
		CU(cudaIpcGetMemHandle(&h, s->ids));
		std::memcpy(handles192 + 128, &h, 64);
	}
	if (s->ret) {
		CU(cudaIpcGetMemHandle(&h, s->ret));
		std::memcpy(handles256 + 192, &h, 64);
	}
	return BL_OK;
}

int blight_part_session_connect_ipc(blight_part_session* s, uint32_t peer, const unsigned char* handles256, uint64_t peer_ids_capacity,
                                    uint64_t peer_id_base) {
	const unsigned char* handles192 = handles256;
	if (!s || !handles256) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (peer >= s->cfg.world || peer == s->cfg.rank) return fail(BL_ERR_INVALID_ARG, "peer rank");
	Guard g(s->device);
	cudaIpcMemHandle_t h;
	void* p = nullptr;
	std::memcpy(&h, handles192, 64);
	CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	s->p_inbox[peer] = p;
	std::memcpy(&h, handles192 + 64, 64);
	CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	s->p_mail[peer] = static_cast<Mailbox*>(p);
	if (peer_ids_capacity) {
		std::memcpy(&h, handles192 + 128, 64);
		CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
		s->p_ids[peer] = static_cast<int64_t*>(p);
		s->p_ids_cap[peer] = peer_ids_capacity;
		if (s->stream_ret) {
			std::memcpy(&h, handles256 + 192, 64);
			CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
			s->p_ret[peer] = static_cast<uint32_t*>(p);
		}
	}
	s->id_base[peer] = peer_id_base;
	s->ipc_opened[peer] = true;
	s->connected |= 1u << peer;
	return BL_OK;
}

int blight_part_session_connect_local(blight_part_session* s, uint32_t peer, const blight_part_session* other) {
	if (!s || !other) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (peer >= s->cfg.world || peer == s->cfg.rank || other->cfg.rank != peer) return fail(BL_ERR_INVALID_ARG, "peer rank");
	if (other->device != s->device) {
		Guard g(s->device);
		int can = 0;
		CU(cudaDeviceCanAccessPeer(&can, s->device, other->device));
		if (!can) return fail(BL_ERR_CUDA, "no peer access between the two devices");
		cudaError_t e = cudaDeviceEnablePeerAccess(other->device, 0);
		if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
		else if (e != cudaSuccess) return cu_fail(e, "cudaDeviceEnablePeerAccess");
	}
	s->p_inbox[peer] = other->inbox;
	s->p_mail[peer] = other->mail;
	s->p_ids[peer] = other->ids;
	s->p_ids_cap[peer] = other->cfg.ids_capacity;
	s->p_ret[peer] = other->ret;
	s->id_base[peer] = other->idx->info.id_base;
	if (other->stream_ret != s->stream_ret || other->pull != s->pull) return fail(BL_ERR_INVALID_ARG, "the ranks disagree on the return path");
	s->connected |= 1u << peer;
	return BL_OK;
}

void* blight_part_session_ids(const blight_part_session* s) { return s ? s->ids : nullptr; }

uint64_t blight_part_session_sub_batches(const blight_part_session* s, uint64_t total_bases, int want_ids) {
	if (!s) return 0;
	(void)want_ids;
	return sub_batch_cuts(total_bases, s->cfg.sub_positions).size() - 1;
}

int blight_part_session_query(blight_part_session* s, const char* d_bases, const uint64_t* d_read_off, const uint64_t* d_kmer_off,
                              uint64_t n_reads, uint64_t total_bases, uint64_t n_sub, uint64_t* d_ctr, void* stream) {
	ReadBatch B;
	B.d_bases = d_bases; B.d_read_off = d_read_off; B.d_kmer_off = d_kmer_off; B.n_reads = n_reads; B.total_bases = total_bases;
	return part_session_query_batch(s, B, n_sub, d_ctr, stream);
}

}  // extern "C"

int blight::part_session_query_batch(blight_part_session* s, const ReadBatch& RB, uint64_t n_sub, uint64_t* d_ctr, void* stream) {
	const char* d_bases = RB.d_bases;
	const uint64_t* d_read_off = RB.d_read_off;
	const uint64_t* d_kmer_off = RB.d_kmer_off;
	const uint64_t n_reads = RB.n_reads, total_bases = RB.total_bases;
	if (!s || !d_ctr || (n_reads && ((!d_bases && !RB.d_packed) || !d_read_off))) return fail(BL_ERR_INVALID_ARG, "null argument");
	const blight_part_config& c = s->cfg;
	if (s->connected != (c.world == 32 ? 0xFFFFFFFFu : ((1u << c.world) - 1u))) return fail(BL_ERR_INVALID_ARG, "not every peer is connected");
	const bool want_ids = d_kmer_off != nullptr;
	if (want_ids && !s->ids) return fail(BL_ERR_INVALID_ARG, "the session was created without an id array (ids_capacity)");
	const std::vector<uint64_t> sub_cut = sub_batch_cuts(total_bases, c.sub_positions);
	if (sub_cut.size() - 1 > n_sub) return fail(BL_ERR_INVALID_ARG, "n_sub sub-batches do not cover the reads");
	Guard g(s->device);
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	const uint32_t world = c.world;
	const long long spin_limit = 60ll * 2000000000ll;  // ~60 s of SM clock

	const bool stream_ret = want_ids && s->stream_ret;
	blight_part_route routes[2];
	const void* regions[2][kMaxRanks];
	void* ret_at[2][kMaxRanks];          // stream return: where this owner's lookups store the ids of source d (at d; pull: here)
	const void* ret_from[2][kMaxRanks];  // ... and where this source's scatter pass reads the ids of owner d (here; pull: at d)
	for (int b = 0; b < 2; b++) {
		blight_part_route& rt = routes[b];
		std::memset(&rt, 0, sizeof rt);
		rt.world = world; rt.rank = c.rank; rt.lb = c.lb; rt.cap = c.cap; rt.kcap = s->kcap;
		rt.side = stream_ret ? static_cast<void*>(s->side + (size_t)b * world * c.cap) : nullptr;
		for (uint32_t i = 0; i <= world; i++) rt.cuts[i] = c.cuts[i];
		for (uint32_t d = 0; d < world; d++) {
			rt.inbox[d] = static_cast<char*>(s->p_inbox[d]) + ((size_t)b * world + c.rank) * s->region_bytes;  // [half b][source = me] at owner d
			regions[b][d] = static_cast<const char*>(s->inbox) + ((size_t)b * world + d) * s->region_bytes;   // [half b][source d] here
			const size_t mine_at_d = ((size_t)b * world + c.rank) * s->kcap, d_here = ((size_t)b * world + d) * s->kcap;
			// push: [half b][owner = me] at source d, read back from [half b][owner d] here; pull: [half b][source d] here,
			// fetched from [half b][source = me] at owner d
			ret_at[b][d] = stream_ret ? static_cast<void*>(s->pull ? s->ret + d_here : s->p_ret[d] + mine_at_d) : nullptr;
			ret_from[b][d] = stream_ret ? static_cast<const void*>(s->pull ? s->p_ret[d] + mine_at_d : s->ret + d_here) : nullptr;
		}
	}
	MailPtrs mp{};
	void* out_ids[kMaxRanks];
	for (uint32_t d = 0; d < world; d++) { mp.m[d] = s->p_mail[d]; out_ids[d] = s->p_ids[d]; }
	bool scat_pending[2] = {false, false};
	// BLIGHT_PART_TRACE=1 (diagnostic): timestamps of every step of the batch, printed per rank when the batch has drained
	const char* trace_env = getenv("BLIGHT_PART_TRACE");
	const bool trace_on = trace_env && trace_env[0] == '1';
	struct Mark { char tag; uint64_t i; cudaEvent_t ev; };
	std::vector<Mark> marks;
	auto mark = [&](char tag, uint64_t i, cudaStream_t on) {
		if (!trace_on) return;
		cudaEvent_t ev;
		if (cudaEventCreate(&ev) != cudaSuccess) return;
		cudaEventRecord(ev, on);
		marks.push_back({tag, i, ev});
	};
	mark('0', 0, st);

	auto dispatch = [&](uint64_t i) -> int {
		const int b = (int)(i & 1);
		if (scat_pending[b]) {  // the scatter of sub-batch i-2 still reads this half's counters, side table and return regions
			CU(cudaStreamWaitEvent(st, s->ev_scat[b], 0));
			scat_pending[b] = false;
		}
		CU(cudaMemsetAsync(s->counts[b], 0, (kMaxRanks + 1) * 8, st));
		if (i + 1 < sub_cut.size() && n_reads) {
			int rc = part_dispatch_batch(s->idx->v.k, s->idx->v.m, RB, sub_cut[i], sub_cut[i + 1], &routes[b],
			                             reinterpret_cast<uint64_t*>(s->counts[b]), d_ctr, s->err, st,
			                             s->tickets ? reinterpret_cast<uint64_t*>(s->counts[b] + kMaxRanks) : nullptr);
			if (rc != BL_OK) return rc;
		}
		mark('D', i, st);
		return BL_OK;
	};
	auto publish = [&](uint64_t i) -> int {
		k_publish_counts<<<1, kMaxRanks, 0, st>>>(mp, s->counts[i & 1], world, c.rank, (uint32_t)(i & 1), s->seq + i + 1);
		g_launches++;
		CU(cudaGetLastError());
		return BL_OK;
	};
	auto wait = [&](uint64_t i) -> int {
		k_wait_counts<<<1, kMaxRanks, 0, st>>>(s->mail, world, (uint32_t)(i & 1), s->seq + i + 1, s->rcv[i & 1], s->err, spin_limit);
		g_launches++;
		CU(cudaGetLastError());
		mark('W', i, st);
		return BL_OK;
	};
	auto lookup_on = [&](uint64_t i, cudaStream_t on) -> int {
		const int b = (int)(i & 1);
		const int rc = part_lookup_from(s->idx, world, c.rank, regions[b], reinterpret_cast<const uint64_t*>(s->rcv[b]), stream_ret ? ret_at[b] : nullptr,
		                                         want_ids && !stream_ret ? out_ids : nullptr, want_ids && !stream_ret ? s->p_ids_cap : nullptr, c.cap,
		                                         s->kcap, d_ctr, on, s->tickets ? reinterpret_cast<uint64_t*>(s->rcv[b] + kMaxRanks) : nullptr);
		mark('L', i, on);
		return rc;
	};
	auto lookup = [&](uint64_t i) -> int { return lookup_on(i, st); };
	// stream return: sub-batch i's ids into the id array, on the scatter stream, once the main stream has passed a point
	// after which every owner is known to have finished L(i) (the wait of sub-batch i+1, or the end-of-batch fence)
	auto scatter = [&](uint64_t i) -> int {
		if (!stream_ret) return BL_OK;
		const int b = (int)(i & 1);
		CU(cudaEventRecord(s->ev_pub, st));
		CU(cudaStreamWaitEvent(s->scat_st, s->ev_pub, 0));
		mark('s', i, s->scat_st);
		int rc = part_scatter_from(s->side + (size_t)b * world * c.cap, c.cap, reinterpret_cast<const uint64_t*>(s->counts[b]), ret_from[b], s->kcap,
		                           world, c.rank, (uint64_t)world * c.cap, s->id_base, s->ids, s->scat_st);
		if (rc != BL_OK) return rc;
		mark('S', i, s->scat_st);
		CU(cudaEventRecord(s->ev_scat[b], s->scat_st));
		scat_pending[b] = true;
		return BL_OK;
	};
	int rc = BL_OK;
	struct GridLimit { ~GridLimit() { g_part_blocks_per_sm = 0; } } grid_limit;  // whatever path leaves this function
#define STEP(x) do { rc = (x); if (rc != BL_OK) return rc; } while (0)
	if (s->order == BLIGHT_PART_ORDER_SERIAL) {
		for (uint64_t i = 0; i < n_sub; i++) {
			STEP(dispatch(i)); STEP(publish(i)); STEP(wait(i));
			if (i) STEP(scatter(i - 1));
			STEP(lookup(i));
		}
	} else if (s->order == BLIGHT_PART_ORDER_AHEAD && n_sub) {
		STEP(dispatch(0)); STEP(publish(0)); STEP(wait(0));
		for (uint64_t i = 0; i < n_sub; i++) {
			if (i) STEP(scatter(i - 1));
			if (i + 1 < n_sub) STEP(dispatch(i + 1));
			STEP(lookup(i));
			if (i + 1 < n_sub) { STEP(publish(i + 1)); STEP(wait(i + 1)); }
		}
	} else if (n_sub) {
		// overlap: the caller's stream carries D / P / W, the session's second stream the lookups; while both kinds of kernel
		// are in flight each takes a share of an SM's CTA slots, the first dispatch and the last lookup take all of them
		STEP(dispatch(0)); STEP(publish(0)); STEP(wait(0));
		for (uint64_t i = 0; i < n_sub; i++) {
			const bool both = i + 1 < n_sub;
			if (i) STEP(scatter(i - 1));
			CU(cudaEventRecord(s->ev_main, st));              // W(i) done (and, for i = 0, whatever the caller queued before)
			CU(cudaStreamWaitEvent(s->side_st, s->ev_main, 0));
			g_part_blocks_per_sm = both ? s->split_lookup : 0;
			STEP(lookup_on(i, s->side_st));
			CU(cudaEventRecord(s->ev_side, s->side_st));
			g_part_blocks_per_sm = both ? s->split_dispatch : 0;
			if (both) STEP(dispatch(i + 1));
			g_part_blocks_per_sm = 0;
			CU(cudaStreamWaitEvent(st, s->ev_side, 0));       // P(i+1) tells the peers this rank is done READING half i & 1 too
			if (both) { STEP(publish(i + 1)); STEP(wait(i + 1)); }
		}
	}
	// end-of-batch fence: once every rank's flag is here, every owner has finished its lookups, so the identifiers it stored
	// into this rank's id array have landed
	k_publish_counts<<<1, kMaxRanks, 0, st>>>(mp, nullptr, world, c.rank, 2u, s->seq + n_sub + 1);
	k_wait_counts<<<1, kMaxRanks, 0, st>>>(s->mail, world, 2u, s->seq + n_sub + 1, nullptr, s->err, spin_limit);
	g_launches += 2;
	CU(cudaGetLastError());
	s->seq += n_sub + 1;
	if (n_sub) STEP(scatter(n_sub - 1));
	for (int b = 0; b < 2; b++)
		if (scat_pending[b]) CU(cudaStreamWaitEvent(st, s->ev_scat[b], 0));
#undef STEP
	if (trace_on) {
		mark('E', n_sub, st);
		cudaStreamSynchronize(st);
		std::string line = "{\"part_trace\": {\"rank\": " + std::to_string(c.rank) + ", \"ids\": " + (want_ids ? "true" : "false") + ", \"t_ms\": [";
		for (size_t j = 0; j < marks.size(); j++) {
			float ms = 0;
			cudaEventElapsedTime(&ms, marks[0].ev, marks[j].ev);
			char buf[64];
			snprintf(buf, sizeof buf, "%s[\"%c%llu\", %.3f]", j ? ", " : "", marks[j].tag, (unsigned long long)marks[j].i, ms);
			line += buf;
		}
		line += "]}}\n";
		fputs(line.c_str(), stderr);
		for (auto& mk : marks) cudaEventDestroy(mk.ev);
	}
	return BL_OK;
}

extern "C" {

int blight_part_session_status(blight_part_session* s, uint32_t* flags_out, int reset, void* stream) {
	if (!s || !flags_out) return fail(BL_ERR_INVALID_ARG, "null argument");
	Guard g(s->device);
	cudaStream_t st = static_cast<cudaStream_t>(stream);
	CU(cudaMemcpyAsync(flags_out, s->err, 4, cudaMemcpyDeviceToHost, st));
	if (reset) CU(cudaMemsetAsync(s->err, 0, 4, st));
	CU(cudaStreamSynchronize(st));
	return BL_OK;
}

}  // extern "C"
