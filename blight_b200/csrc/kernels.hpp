// kernels.hpp — launchers of the sm_100a kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>

#include "device_index.hpp"
#include "errors.hpp"

namespace blight {

extern std::atomic<uint64_t> g_launches;
extern const char* g_last_cuda_error;

// Upload-time pass over every window of the index text (kernels.cu: k_window_answers): the per-position "answered found"
// bitmap d_valid, the per-position local identifier d_lid (id - I.id_base, 0xFFFFFFFF = -1; may be null), the candidate
// bitmap d_cand of the exact-position layout (may be null) and the negative filter d_filter (may be null).
int launch_window_answers(const DevIndexView& I, uint64_t n_buckets, uint64_t total_nuc, uint32_t* d_valid, uint32_t* d_lid, uint32_t* d_cand,
                          uint32_t* d_filter, uint32_t filter_blocks, cudaStream_t stream);
// Exact-position layout: one candidate window per position field (d_claim: one u32 per local identifier, preset to
// 0xFFFFFFFF) ORs where it sits inside its 2^b windows into the field's low b bits (d_pos_rw = the position sectors).
int launch_exact_positions(const DevIndexView& I, uint64_t n_buckets, uint64_t n_groups, uint64_t total_nuc, const uint32_t* d_cand,
                           const uint32_t* d_lid, uint32_t* d_claim, uint32_t* d_pos_rw, cudaStream_t stream);

// ids[i] = lookup(canon[i]); d_mini may be null (the minimizer is then computed from the k-mer).
int launch_lookup_kmers(const DevIndexView& I, const uint64_t* d_canon, const uint32_t* d_mini, uint64_t n, int64_t* d_ids,
                        cudaStream_t stream);

// A batch of reads on the device. The text is either ASCII (d_bases) or 2-bit codes (d_packed: base p in bits
// 30 - 2 (p & 15) .. of word p >> 4, A0 C1 T2 G3 — what the host packer of the *_host entry points produces); exactly one
// is non-null. Read r starts at d_read_off[r] and ends at d_read_end[r] (or d_read_off[r+1] when d_read_end is null);
// d_read_off has n_reads+1 ascending entries; offsets are base positions of the text either way. The ASCII buffer may
// hold anything between reads (FASTA headers, newlines). The offset arrays may be a WINDOW of a larger batch (the reads
// overlapping the positions a launch handles): guess_p0 = where their first read starts, rpb = reads per base there
// (0: n_reads / total_bases), which only steer the first guess of the kernels' read search.
struct ReadBatch {
	const char* d_bases = nullptr;
	const uint32_t* d_packed = nullptr;
	const uint64_t* d_read_off = nullptr;
	const uint64_t* d_read_end = nullptr;
	const uint64_t* d_kmer_off = nullptr;
	uint64_t n_reads = 0, total_bases = 0;
	double rpb = 0;
	uint64_t guess_p0 = 0;
};

// The read kernels over a batch:
//   I == null          : front end only, (canon, minimizer) pairs to d_canon / d_mini at d_kmer_off[r] + position
//   I != null, d_ids   : ids to d_ids at d_kmer_off[r] + position, counters accumulated into d_ctr
//   I != null, !d_ids  : counters only (d_kmer_off unused, may be null)
// Only k-mers starting in [pos_begin, pos_end) are handled (the text must be valid up to pos_end + k - 1), which lets
// a host batch be copied and queried chunk by chunk.
int launch_reads(const DevIndexView* I, uint32_t k, uint32_t m, const ReadBatch& B, uint64_t* d_canon, uint32_t* d_mini, int64_t* d_ids,
                 uint64_t* d_ctr, cudaStream_t stream, uint64_t pos_begin = 0, uint64_t pos_end = ~0ull);

// The super-k-mer read kernel with an id consumer fused behind the lookup (needs I.pos_id and k - m + 1 >= 8):
// kind 0: d_table[id] += 1; kind 1: bit id * n_colors + color of d_table set; kind 2: d_out32[d_kmer_off[r] + pos] = d_table[id]
// (0xFFFFFFFF when absent). Counters are accumulated into d_ctr as in launch_reads.
int launch_reads_sink(const DevIndexView& I, int kind, const ReadBatch& B, uint32_t* d_table, uint32_t n_colors, uint32_t color, uint32_t* d_out32,
                      uint64_t* d_ctr, cudaStream_t stream);

// part_kernels.cu: front end + dispatch of the k-mers starting in [pos_begin, pos_end) of a batch (blight_part_dispatch on a ReadBatch)
int part_dispatch_batch(uint32_t k, uint32_t m, const ReadBatch& B, uint64_t pos_begin, uint64_t pos_end, const ::blight_part_route* route,
                        uint64_t* d_counts, uint64_t* d_ctr, uint32_t* d_err, void* stream, uint64_t* d_ticket = nullptr);

// blight_part_scatter with one pointer per owner: ret[d] = where owner d's 32-bit ids for this source are read from (a local
// region the owner pushed into, or a peer pointer into the owner's own memory: the pull return path of part_session.cu)
int part_scatter_from(const void* d_side, uint64_t cap, const uint64_t* d_counts, const void* const* ret, uint64_t kcap, uint32_t world,
                      uint32_t rank, uint64_t max_records, const uint64_t* id_bases, int64_t* d_ids, void* stream);
// blight_part_lookup_direct with the owner's rank (where its round-robin over the sources starts)
int part_lookup_from(const ::blight_index* idx, uint32_t world, uint32_t rank, const void* const* regions, const uint64_t* d_counts,
                     void* const* ret, void* const* out_ids, const uint64_t* out_caps, uint64_t cap, uint64_t kcap, uint64_t* d_ctr, void* stream,
                     uint64_t* d_ticket = nullptr);  // d_ticket: zeroed device counter, work handed out on demand (null: fixed stride)
int part_kernels_preload();
// resident CTAs per SM the next dispatch / lookup launches of the calling thread may take (0: all that fit)
extern thread_local int g_part_blocks_per_sm;
// part_session.cu: blight_part_session_query on a ReadBatch (records with explicit ends, packed text)
int part_session_query_batch(::blight_part_session* s, const ReadBatch& B, uint64_t n_sub, uint64_t* d_ctr, void* stream);

// Start positions are handled in strips of this many bases; pos_begin of a partial launch must be a multiple of it.
constexpr uint64_t kReadsStrip = 256;

}  // namespace blight
