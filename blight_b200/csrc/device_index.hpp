// device_index.hpp — the HBM layout of a Blight index on a B200 and the view kernels receive.
//
// The flat image (flat_index.hpp) keeps the reference's bit conventions; the device layout below is ours and is
// chosen so that every dependent step of a lookup touches exactly one 32-byte sector:
//
//   bucket table   one uint4 per minimizer bucket {start_lo, start_hi, nuc, 0}          (blight.h:29-34)
//   MPHF table     one 256-byte DevMphf per group: sector bases, id offset, positions geometry, dom[16]
//                                                                                        (blight.h:39-44, bbhash.h:777)
//   level bits     per group, all 16 BBHash levels back to back, cut into 224-bit chunks; each chunk is stored
//                  as one 32-byte sector = 7 x u32 of bits + 1 x u32 "ones before this chunk" (group-relative).
//                  A level probe and the rank that follows (bbhash.h:404-408, 467-480) are ONE sector read
//                  instead of the reference's bit word + rank sample + up to 15 more words.
//   positions      per group, fields of `nbits` bits packed floor(256/nbits) to a 32-byte sector, never
//                  straddling a sector (the reference packs them back to back, blight.cpp:464-482).
//                  With kFlagExactPos (default; BLIGHT_EXACT_POS=0 switches it off) a field is b bits wider and holds the
//                  position of the k-mer's own window: the high bits are the reference's field (position >> b), the low b
//                  bits — which the reference gives up and recovers by scanning 2^b windows (blight.cpp:729-740) — are
//                  filled in at upload from the pass that looks every window up (one owner per field, kernels.cu:
//                  k_claim_low). A lookup then checks that ONE window first
//                  and only scans the 2^b windows from (field with the low bits cleared) when it does not match, which is
//                  what the reference does from the start; the answer is the same either way.
//   sequences      2-bit codes (A0 C1 T2 G3), 16 per u32, FIRST base in the HIGH bits, so a k-mer window is a
//                  funnel shift of adjacent words (the reference stores one bit per vector<bool> slot,
//                  blight.cpp:317-318); zero padded past the end for the 2^b-window scan (blight.cpp:732-739).
//   fallback       sorted (key, rank) arrays for keys no level accommodated (bbhash.h:567-575).
//   valid          1 bit per base position p of the sequences: does the reference answer "found" when queried with the
//                  k-mer spelled by the window starting at p, routed to p's own bucket? Computed once at upload by running
//                  the lookup core itself on every window. It lets a query whose neighbour in the read was found at
//                  position T check the window at T+-1 and, on a match, skip the position read and the 2^b scan while
//                  still answering exactly as the reference would (including its junction-window false positives).
//   pos_id         (optional, fewer than 2^32-1 k-mers in this index or slice) 1 x u32 per base position p: the identifier the
//                  reference returns for the k-mer spelled by the window at p, minus id_base (the first identifier of the
//                  slice); 0xFFFFFFFF when it answers -1. Computed by the same pass as `valid`. The
//                  k-mers of a super-k-mer sit at consecutive positions, so the ids of a run are ONE or two sectors
//                  instead of (levels + rank) sectors per k-mer.
//   filter         register-blocked Bloom filter over V = {canonical k-mers x spelled by ANY window of the text (the 2^b windows
//                  past its end included) for which the reference answers "found" when x is routed to ITS OWN minimizer's
//                  bucket}. Every k-mer the reference answers "found" equals some window of the text — also one that starts
//                  past the end of the bucket it was routed to, because the 2^b scan never re-checks the bucket length
//                  (blight.cpp:732-739) — so it is in V. One 32-byte sector per key, one probe bit in each of its 8 words. No
//                  false negatives, so a miss proves "-1" with a single sector read; a hit (true or false positive) takes
//                  the whole lookup.
#pragma once
#include <cstdint>

#include <vector_types.h>

#include "errors.hpp"
#include "flat_index.hpp"

namespace blight {

constexpr uint32_t kChunkBits = 224;  // level bits per 32-byte sector
constexpr uint32_t kFlagFilterAnchors = 1u;  // first k-mers of runs go through the filter too
constexpr uint32_t kFlagExactPos = 2u;       // position fields hold the exact window (see `positions` above)

struct alignas(64) DevMphf {
	uint64_t bits_sector_base;  // first sector of this group's level bits (index into DevIndexView::bits, in sectors)
	uint64_t pos_sector_base;   // first sector of this group's positions
	uint64_t id_offset;         // exclusive prefix of k-mer counts: id = rank + id_offset (blight.cpp:736)
	uint64_t fb_off;            // slice of the fallback arrays
	uint32_t fb_count;
	uint32_t nbits;             // position field width
	uint32_t fields_per_sector; // floor(256 / nbits)
	uint32_t fps_magic;         // floor(2^32 / fields_per_sector): rank / fps = umulhi(rank, magic) (+1 fix-up)
	uint32_t present;
	uint32_t pad[3];
	uint64_t dom[kLevels];      // level domains
	uint32_t dom32[kLevels];    // the same as 32-bit values when the whole group has fewer than 2^32 level bits
};
static_assert(sizeof(DevMphf) == 256, "DevMphf is 256 bytes");

struct DevIndexView {
	const uint4* bucket;
	const DevMphf* mphf;
	const uint32_t* bits;  // 8 words per sector
	const uint32_t* pos;   // 8 words per sector
	const uint32_t* seq;
	const uint64_t* fb_keys;
	const uint64_t* fb_vals;
	const uint32_t* valid;  // 1 bit per base position p (bit p&31 of word p>>5): see above
	const uint32_t* pos_id; // identifier per base position (0xFFFFFFFF: -1), or null
	const uint32_t* filter; // 8 words per block, or null
	uint32_t filter_blocks;
	uint32_t flags;         // kFlagFilterAnchors
	uint64_t kmask;
	uint64_t id_base;       // pos_id holds id - id_base: the identifiers of a slice (partition mode) start here
	uint32_t k, m, b, lb;
	uint32_t small;  // every MPHF group has fewer than 2^32 level bits: 32-bit bit arithmetic in the probe
};

}  // namespace blight

// Opaque handle of the C ABI.
struct blight_index {
	int device = -1;
	blight::DevIndexView v{};
	blight_info info{};
	void* d_bucket = nullptr; void* d_mphf = nullptr; void* d_bits = nullptr; void* d_pos = nullptr; void* d_seq = nullptr;
	void* d_fbk = nullptr; void* d_fbv = nullptr; void* d_valid = nullptr; void* d_pos_id = nullptr; void* d_filter = nullptr;
	void* host_stream = nullptr;  // internal stream of the *_host entry points
	void* copy_stream = nullptr;  // H2D chunks of a host batch, overlapped with the kernels on host_stream
	void* ev_copy = nullptr;
	void* ev_ws = nullptr;
	void* host_mutex = nullptr;
	void* stream_ctx = nullptr;   // pinned / device buffers of the streaming file_query (stream_query.cu), created on first use
	void* host_pool = nullptr;    // contexts (streams, staging, workspaces) of the host-buffer entry points (host_query.cu)
	void* ws[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // grow-only scratch of the *_host entry points
	size_t ws_cap[6] = {0, 0, 0, 0, 0, 0};
};
