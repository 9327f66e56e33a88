/* blight_b200.h — C ABI of the B200-native batched k-mer query path of Blight.
 *
 * The reference has no FFI: its boundary is the public C++ class kmer_Set_Light (blight.h:15-136).
 * This header is what a binding for that class calls into; the C++ mirror of the class that sits on
 * top of it is blight_b200/csrc/kmer_set_light.hpp.  Each entry point cites the reference interface
 * it replaces.  Conventions: plain pointers and sizes, `int` status (0 = BLIGHT_OK, negative = error,
 * text via blight_last_error()), caller owns every buffer it passes, the library owns the objects it
 * returns until the matching *_free.  "d_" parameters are DEVICE pointers on the index's device and the
 * call is asynchronous on `stream` (a cudaStream_t passed as void*, NULL = legacy default stream);
 * "h_" / unprefixed buffers are host memory and those calls return when the result is in the buffer.
 * Concurrent queries on one index from several host threads are allowed: device-buffer calls when each uses its
 * own stream, *_host calls always (each takes a context of its own from a pool kept with the index).
 *
 * There is no CPU fallback behind this ABI: every query entry point runs the sm_100a kernels or
 * returns BLIGHT_ERR_NO_DEVICE / BLIGHT_ERR_CUDA.
 */
#ifndef BLIGHT_B200_H
#define BLIGHT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLIGHT_OK 0
#define BLIGHT_ERR_INVALID_ARG (-1)  /* std::invalid_argument of the constructor, blight.h:75-92 */
#define BLIGHT_ERR_IO (-2)           /* std::runtime_error("Problem with files opening"), blight.cpp:188-189 */
#define BLIGHT_ERR_INVALID_BASE (-3) /* std::domain_error("Invalid char in DNA"), kmer.h:68 */
#define BLIGHT_ERR_CUDA (-4)
#define BLIGHT_ERR_NO_DEVICE (-5)
#define BLIGHT_ERR_FORMAT (-6)
#define BLIGHT_ERR_NOMEM (-7)

typedef struct blight_flat blight_flat;   /* host-side flat index image (BLFLAT01) */
typedef struct blight_index blight_index; /* device-resident index */

/* Index of the counters written by the read-query entry points. */
#define BLIGHT_CTR_FOUND 0     /* "Good kmer", TP, blight.cpp:784-785,793 */
#define BLIGHT_CTR_NOT_FOUND 1 /* "Erroneous kmers", FP, blight.cpp:786-787,794 */
#define BLIGHT_CTR_QUERIES 2   /* number_query, blight.cpp:687-688,795 */
#define BLIGHT_CTR_INVALID 3   /* queried k-mers holding a byte nuc2int rejects (kmer.h:68); non-zero => BLIGHT_ERR_INVALID_BASE */
#define BLIGHT_N_CTR 4

typedef struct blight_info {
	uint32_t k, m, n_log2, s_log2, b;
	uint32_t layout;            /* blight_index only: BLIGHT_LAYOUT_* bits of the derived tables actually resident */
	uint64_t n_buckets;         /* 2^(2m-1), blight.h:70 */
	uint64_t n_mphf;            /* 2^n, blight.h:68 */
	uint64_t number_kmer;       /* kmer_Set_Light::number_kmer, blight.h:52 */
	uint64_t number_super_kmer; /* kmer_Set_Light::number_super_kmer, blight.h:53 */
	uint64_t total_nuc;         /* bucketSeq.size()/2 */
	uint64_t positions_bits;    /* positions.size() */
	uint64_t mphf_bits;         /* sum of the BBHash level bit arrays */
	uint64_t fallback_keys;     /* entries of the BBHash fallback maps */
	uint64_t largest_mphf;      /* kmer_Set_Light::largest_MPHF, blight.h:54 */
	uint64_t largest_bucket;    /* kmer_Set_Light::largest_bucket_nuc_all, blight.h:59 */
	uint64_t device_bytes;      /* HBM bytes held by a blight_index (0 for a blight_flat) */
	uint64_t id_base;           /* smallest identifier this index can return: 0 for a whole index, the slice's first for blight_flat_slice */
} blight_info;

const char* blight_version(void);
/* Message of the last error raised on the calling thread ("" if none). */
const char* blight_last_error(void);

/* ---- host side: construction and (de)serialisation of the flat index ---------------------------------- */

/* kmer_Set_Light(k,m,n,s,cores,b) parameter validation, blight.h:62-96. */
int blight_check_params(uint32_t k, uint32_t m, uint32_t n_log2, uint32_t s_log2, uint32_t b);

/* kmer_Set_Light::construct_index(file), blight.h:134 / blight.cpp:108-125: builds, on the host, bit for bit the
 * index the reference builds with cores=1 (2-line FASTA records, plain or gzip). `threads` is ours (0 = all). */
int blight_flat_build_file(const char* unitig_path, uint32_t k, uint32_t m, uint32_t n_log2, uint32_t s_log2, uint32_t b,
                           uint32_t threads, blight_flat** out);
/* Same, from sequences already in memory: sequence i = bases[offsets[i] .. offsets[i+1]). */
int blight_flat_build_seqs(const char* bases, const uint64_t* offsets, uint64_t n_seqs, uint32_t k, uint32_t m,
                           uint32_t n_log2, uint32_t s_log2, uint32_t b, uint32_t threads, blight_flat** out);
/* Same, sequence i = bases[starts[i] .. starts[i]+lengths[i]) — spans may overlap (unitigs cut from one genome). */
int blight_flat_build_spans(const char* bases, const uint64_t* starts, const uint64_t* lengths, uint64_t n_seqs,
                            uint32_t k, uint32_t m, uint32_t n_log2, uint32_t s_log2, uint32_t b, uint32_t threads,
                            blight_flat** out);
/* construct_index ON THE GPU: the same flat image, word for word, as the host builders above (and hence as
 * the reference with cores=1), built by data-parallel passes on `device` (csrc/gpu_builder.cu). Sequence i =
 * bases[starts[i] .. starts[i]+lengths[i]) in HOST memory (spans may overlap). The input must be a k-mer SET, as BCALM unitigs
 * are. *device_seconds (may be NULL) receives the time between the first H2D copy and the last kernel. */
int blight_flat_build_gpu(const char* bases, const uint64_t* starts, const uint64_t* lengths, uint64_t n_seqs, uint32_t k, uint32_t m,
                          uint32_t n_log2, uint32_t s_log2, uint32_t b, int device, blight_flat** out, double* device_seconds);
int blight_flat_build_file_gpu(const char* unitig_path, uint32_t k, uint32_t m, uint32_t n_log2, uint32_t s_log2, uint32_t b, int device,
                               blight_flat** out);
/* The reference has no index persistence (only mphf::save/load, bbhash.h:731-775); this is ours. */
int blight_flat_save(const blight_flat* f, const char* path);
int blight_flat_load(const char* path, blight_flat** out);
void blight_flat_free(blight_flat* f);
int blight_flat_info(const blight_flat* f, blight_info* out);
/* 0 if a and b hold identical indices, 1 if they differ (description via blight_last_error), <0 on error. */
int blight_flat_compare(const blight_flat* a, const blight_flat* b);
/* Keeps only MPHF groups [g_begin, g_end) (minimizer-bucket partition for multi-GPU); ids stay global. */
int blight_flat_slice(const blight_flat* f, uint64_t g_begin, uint64_t g_end, blight_flat** out);
/* k-mer count of every MPHF group (n_mphf entries), for balancing a partition. */
int blight_flat_group_sizes(const blight_flat* f, uint64_t* sizes_out);

/* ---- device side ------------------------------------------------------------------------------------ */

/* Re-lays the flat image out for the GPU and uploads it to `device`, with the default (fastest) set of derived tables. */
int blight_index_upload(const blight_flat* f, int device, blight_index** out);

/* The HBM footprint is a choice: the reference's own arrays re-laid out cost ~28 bits per k-mer; the
 * derived tables that make the read kernels 2.5x faster cost ~130 more. Each field: 1 = on, 0 = off, -1 = library default. */
#define BLIGHT_LAYOUT_POS_ID 1u    /* position -> identifier table: 32 bits per base of index text (~106 bits per k-mer) */
#define BLIGHT_LAYOUT_FILTER 2u    /* negative filter: filter_bits per k-mer */
#define BLIGHT_LAYOUT_EXACT_POS 4u /* position fields widened by b bits: the 2^b-window scan becomes one window */
typedef struct blight_upload_options {
	uint32_t struct_size;   /* = sizeof(blight_upload_options) */
	int32_t pos_id;         /* default 1 (dropped by itself when the index or slice holds 2^32-1 k-mers or more, or HBM is short) */
	int32_t filter_bits;    /* bits per k-mer, 0 = no filter; default 20 */
	int32_t exact_pos;      /* default 1 */
	int32_t filter_anchors; /* first k-mers of super-k-mers go through the filter too; default 1 */
} blight_upload_options;
/* opts == NULL: all defaults. "Compact" = {sizeof, 0, 0, 0, 0}: only the reference's arrays (+1 bit per base). */
int blight_index_upload_opts(const blight_flat* f, int device, const blight_upload_options* opts, blight_index** out);
void blight_index_free(blight_index* idx);
int blight_index_info(const blight_index* idx, blight_info* out);

/* kmer_Set_Light::query_kmer_hash(canon), blight.h:131 / blight.cpp:545-550, batched: d_canon[i] must already be
 * canonical; d_ids[i] = identifier in [0, N) or -1. */
int blight_query_kmers(const blight_index* idx, const uint64_t* d_canon, uint64_t n, int64_t* d_ids, void* stream);

/* Same with the minimizer bucket supplied by the caller (query_get_hash(canon, minimizer), blight.cpp:716-742);
 * used by the bucket-partitioned multi-GPU path, where the owner receives (canon, minimizer) pairs. */
int blight_query_kmers_mini(const blight_index* idx, const uint64_t* d_canon, const uint32_t* d_mini, uint64_t n,
                            int64_t* d_ids, void* stream);

/* Front end only: canonical k-mers and minimizers of every k-mer of every read, in read order then position
 * (query_sequence_hash's order, blight.cpp:575-591; minimizer_naive, kmer.h:791-810). Read r is
 * d_bases[d_read_off[r] .. d_read_off[r+1]); its k-mers land at d_kmer_off[r].. (d_kmer_off = exclusive prefix of
 * max(0, len-k+1), n_reads+1 entries). d_ctr[BLIGHT_CTR_INVALID] counts k-mers holding a rejected byte. Needs no index: pass k, m. */
int blight_reads_to_kmers(uint32_t k, uint32_t m, const char* d_bases, const uint64_t* d_read_off,
                          const uint64_t* d_kmer_off, uint64_t n_reads, uint64_t total_bases, uint64_t* d_canon,
                          uint32_t* d_mini, uint64_t* d_ctr, void* stream);

/* kmer_Set_Light::query_sequence_hash / query_sequence_bool over a batch of reads (blight.h:129,133), the body of
 * file_query's loop (blight.cpp:780-789). d_ids may be NULL (bool mode: only counters). d_ctr has BLIGHT_N_CTR
 * entries and is ACCUMULATED into (zero it first). Reads shorter than k contribute nothing (blight.cpp:557-559). */
int blight_query_reads(const blight_index* idx, const char* d_bases, const uint64_t* d_read_off,
                       const uint64_t* d_kmer_off, uint64_t n_reads, uint64_t total_bases, uint64_t total_kmers,
                       int64_t* d_ids, uint64_t* d_ctr, void* stream);

/* The same on reads the caller already holds as 2-bit codes: base p of the batch in bits 30 - 2 (p & 15) .. of word p >> 4,
 * code (c >> 1) & 3 as nuc2int (kmer.h:56-69: A0 C1 T2 G3), first base in the high bits; offsets are base positions as above. */
int blight_query_reads_packed(const blight_index* idx, const uint32_t* d_packed, const uint64_t* d_read_off,
                              const uint64_t* d_kmer_off, uint64_t n_reads, uint64_t total_bases, int64_t* d_ids,
                              uint64_t* d_ctr, void* stream);

/* ---- id consumers fused behind the lookup: what the reference's applications do with the ids of
 * query_sequence_hash, done on the GPU so that ids never leave it. Need the position->id table (N < 2^32-1). ------- */
#define BLIGHT_CONSUME_COUNT 0 /* d_table[id] += 1 (uint32): `abundance[kmer_ids[i]]++`, Abundance_De_Bruijn_graph_snippet.cpp:132-142
                                  (the snippet's uint8 counter is this value mod 256) */
#define BLIGHT_CONSUME_COLOR 1 /* bit id * n_colors + color of d_table set: `color[kmer_ids[i]*color_number+i_file]=true`,
                                  Colored_De_Bruijn_graph_snippet.cpp:131-141 (vector<bool> bit order: bit j of word j/32) */
int blight_consume_reads(const blight_index* idx, const char* d_bases, const uint64_t* d_read_off, uint64_t n_reads,
                         uint64_t total_bases, int kind, uint32_t* d_table, uint32_t n_colors, uint32_t color, uint64_t* d_ctr,
                         void* stream);
/* The query side of the same applications (…snippet.cpp:176-192): d_out[d_kmer_off[r] + pos] = d_table[id] for every
 * k-mer of every read, 0xFFFFFFFF for k-mers the index does not hold. */
int blight_gather_reads(const blight_index* idx, const char* d_bases, const uint64_t* d_read_off, const uint64_t* d_kmer_off,
                        uint64_t n_reads, uint64_t total_bases, const uint32_t* d_table, uint32_t* d_out, uint64_t* d_ctr,
                        void* stream);

/* ---- host-buffer entry points (end to end: H2D, kernels, D2H inside the call) ------------------------- */

/* kmer_Set_Light::file_query on a text buffer holding 2-line FASTA records (blight.cpp:746-799): same record
 * pairing, same skip rule; ctr[BLIGHT_N_CTR] receives Good / Erroneous / Query performed / invalid bytes. */
int blight_query_fasta_host(const blight_index* idx, const char* text, uint64_t len, uint64_t* ctr);
/* file_query(path): plain or gzip file. */
int blight_query_file_host(const blight_index* idx, const char* path, uint64_t* ctr);
/* query_sequence_hash(seq): ids_out needs max(0, len-k+1) slots; *n_out receives the count. */
int blight_query_sequence_host(const blight_index* idx, const char* seq, uint64_t len, int64_t* ids_out,
                               uint64_t* n_out);
/* query_sequence_bool(seq), blight.h:129 / blight.cpp:554-571: (Good, Erroneous) of one sequence, counted on the device. */
int blight_query_sequence_bool_host(const blight_index* idx, const char* seq, uint64_t len, uint64_t* found, uint64_t* not_found);
/* Batched form over host reads (no separators; offsets as above); ids_out may be NULL. Large batches cross PCIe partly as
 * ASCII and partly 2-bit packed by the host cores (csrc/host_query.cu); BLIGHT_HOST_PACK=0 / BLIGHT_HOST_THREADS=n tune it. */
int blight_query_reads_host(const blight_index* idx, const char* bases, const uint64_t* read_off, uint64_t n_reads,
                            int64_t* ids_out, uint64_t* ctr);
/* query_kmer_hash over a host array of canonical k-mers. */
int blight_query_kmers_host(const blight_index* idx, const uint64_t* canon, uint64_t n, int64_t* ids_out);

/* ---- bucket-partitioned multi-GPU mode: the kernels either side of the all-to-all (no reference counterpart; the
 * owner rule is the reference's MPHF selection minimizer / number_bucket_per_mphf, blight.cpp:722) ---------------- */

/* d_counts[o] += number of k-mers whose MPHF group (minimizer >> lb) lies in [d_group_cuts[o], d_group_cuts[o+1]). */
int blight_owner_count(const uint32_t* d_mini, uint64_t n, const uint32_t* d_group_cuts, uint32_t world, uint32_t lb,
                       uint64_t* d_counts, void* stream);
/* Packs (canon, minimizer, source index) by owner. d_cursors[o] must hold the exclusive prefix of the counts. */
int blight_owner_scatter(const uint64_t* d_canon, const uint32_t* d_mini, uint64_t n, const uint32_t* d_group_cuts,
                         uint32_t world, uint32_t lb, uint64_t* d_cursors, uint64_t* d_send_canon, uint32_t* d_send_mini,
                         uint64_t* d_send_src, void* stream);
/* d_out[d_src[i]] = d_ids_back[i]: ids returned by the owners back into query order. */
int blight_scatter_ids(const int64_t* d_ids_back, const uint64_t* d_src, uint64_t n, int64_t* d_out, void* stream);

/* ---- the same mode with the exchange fused into the kernels (peer-memory stores over NVLink; no reference
 * counterpart). The source GPU cuts its reads into super-k-mers (runs of equal minimizer, kmer.h:629-693) and stores
 * one 32-byte record per run straight into the owner's inbox; the owner looks the run up and stores the identifiers
 * straight into the source's id buffer. Between the two kernels the caller exchanges the per-pair record counts
 * (one tiny all-to-all, which is also the barrier). ------------------------------------------------------------- */

#define BLIGHT_MAX_RANKS 16
#define BLIGHT_RUN_RECORD_BYTES 32

typedef struct blight_part_route {
	uint32_t world, rank;                /* ranks of the partition, this (source) rank */
	uint32_t lb;                         /* log2(buckets per MPHF group): group = minimizer >> lb (blight.cpp:722) */
	uint32_t reserved;
	uint32_t cuts[BLIGHT_MAX_RANKS + 1]; /* rank r owns MPHF groups [cuts[r], cuts[r+1]) */
	void* inbox[BLIGHT_MAX_RANKS];       /* DEVICE pointers: this source's region in every owner's inbox (peer memory) */
	uint64_t cap;                        /* records per region (< 2^24) */
	uint64_t kcap;                       /* k-mers per return region (< 2^32) */
	void* side;                          /* DEVICE, local: world * cap entries of 16 bytes (where each run's ids go); NULL in counting mode */
} blight_part_route;

/* Front end + dispatch of the k-mers starting in [pos_begin, pos_end) of a read batch (bounds: multiples of 256, or
 * the end). d_kmer_off == NULL: counting mode. d_counts[world] is ACCUMULATED into: per owner, records stored << 40 |
 * k-mers stored (zero it per sub-batch); d_ctr gets BLIGHT_CTR_QUERIES / BLIGHT_CTR_INVALID; *d_err |= 1 if a region
 * overflowed (those records are dropped: the caller must retry with smaller sub-batches or another path). */
int blight_part_dispatch(uint32_t k, uint32_t m, const char* d_bases, const uint64_t* d_read_off, const uint64_t* d_kmer_off,
                         uint64_t n_reads, uint64_t total_bases, uint64_t pos_begin, uint64_t pos_end, const blight_part_route* route,
                         uint64_t* d_counts, uint64_t* d_ctr, uint32_t* d_err, void* stream);
/* Owner side: regions[s] = records received from source s (device pointer), d_counts[s] = the packed counter source s
 * accumulated for this owner (DEVICE array: no host round trip), ret[s] = this owner's return region at source s
 * (peer pointer, 32-bit ids, 0xFFFFFFFF = -1; ret == NULL: counting mode). cap / kcap: records per inbox region and ids
 * per return region, as in the route (reads and writes never leave them). d_ctr gets BLIGHT_CTR_FOUND / BLIGHT_CTR_NOT_FOUND (accumulated). */
int blight_part_lookup(const blight_index* idx, uint32_t world, const void* const* regions, const uint64_t* d_counts,
                       void* const* ret, uint64_t cap, uint64_t kcap, uint64_t* d_ctr, void* stream);
/* Back on the source, once every owner has answered: return regions (d_ret: world regions of kcap 32-bit ids, region d
 * written by owner d) -> int64 ids at the slots query_sequence_hash would fill (blight.cpp:575-591), through the side
 * table and the counters the dispatch left. Owners return slice-local 32-bit ids: id_bases[d] (HOST array, world entries,
 * blight_info.id_base of owner d's index; NULL = all zero) is added back here. */
int blight_part_scatter(const void* d_side, uint64_t cap, const uint64_t* d_counts, const void* d_ret, uint64_t kcap, uint32_t world,
                        uint64_t max_records, const uint64_t* id_bases, int64_t* d_ids, void* stream);
/* blight_part_lookup with the identifiers returned DIRECTLY: out_ids[s] = source s's int64 id array (peer pointer),
 * out_caps[s] its length; the records must come from a dispatch whose route had side == NULL and d_kmer_off != NULL (they
 * then carry the slot of each run's first k-mer in that array). ret must be NULL. No scatter pass follows. */
int blight_part_lookup_direct(const blight_index* idx, uint32_t world, const void* const* regions, const uint64_t* d_counts,
                              void* const* ret, void* const* out_ids, const uint64_t* out_caps, uint64_t cap, uint64_t kcap,
                              uint64_t* d_ctr, void* stream);

/* ---- one rank of the partitioned path as an object: buffers, peers, and the per-batch pipeline, with the ordering between
 * GPUs done by device-side flags in peer memory (no collective call on the data path). The ranks are processes (exchange the
 * handles, connect_ipc) or devices of one process (connect_local). ------------------------------------------------------ */
typedef struct blight_part_session blight_part_session;
typedef struct blight_part_config {
	uint32_t world, rank;
	uint32_t lb;                         /* log2(buckets per MPHF group) */
	uint32_t order;                      /* BLIGHT_PART_ORDER_*: how the kernels of consecutive sub-batches are ordered */
	uint32_t cuts[BLIGHT_MAX_RANKS + 1]; /* rank r owns MPHF groups [cuts[r], cuts[r+1]) */
	uint64_t sub_positions;              /* base positions per sub-batch (multiple of 256, < 2^32) */
	uint64_t cap;                        /* records per (source, owner) inbox region and sub-batch (< 2^24) */
	uint64_t ids_capacity;               /* entries of this rank's id array (0: counting mode only) */
	uint32_t return_path;                /* BLIGHT_PART_RETURN_*: how identifiers travel back to the source */
	uint32_t reserved;
	uint64_t ret_kmers;                  /* stream return: ids per (owner, sub-batch) region; 0 = sub_positions (can never overflow);
	                                        smaller saves memory, a sub-batch sending one owner more raises BLIGHT_PART_OVERFLOW */
} blight_part_config;
#define BLIGHT_PART_RETURN_DEFAULT 0u /* what BLIGHT_PART_RETURN says (stream | pull | direct), else direct */
#define BLIGHT_PART_RETURN_STREAM 1u  /* contiguous 32-bit id streams per owner warp + a scatter pass at the source */
#define BLIGHT_PART_RETURN_DIRECT 2u  /* int64 ids stored by the owner straight into the source's id array */
#define BLIGHT_PART_RETURN_PULL 3u    /* as STREAM, but the streams stay in the owner's memory and the source's scatter pass fetches them */
#define BLIGHT_PART_ORDER_DEFAULT 0u /* what BLIGHT_PART_ORDER says (serial | ahead | overlap), else the library's choice */
#define BLIGHT_PART_ORDER_SERIAL 1u  /* dispatch(i), lookup(i), dispatch(i+1), ... on one stream */
#define BLIGHT_PART_ORDER_AHEAD 2u   /* dispatch(i+1) before lookup(i) on one stream: the wait for the peers never sees dispatch skew */
#define BLIGHT_PART_ORDER_OVERLAP 3u /* dispatch(i+1) BESIDE lookup(i) on two streams, each on half of every SM's CTA slots */
#define BLIGHT_PART_OVERFLOW 1u /* status flag: an inbox region was too small, records were dropped (answer the batch another way) */
#define BLIGHT_PART_TIMEOUT 2u  /* status flag: a peer's flag never arrived */
int blight_part_session_create(const blight_index* local_slice, const blight_part_config* cfg, blight_part_session** out);
void blight_part_session_free(blight_part_session* s);
/* 4 x 64 bytes: CUDA IPC handles of this rank's inbox, mailbox, id array and return regions (zeros when there is none). */
int blight_part_session_handles(const blight_part_session* s, unsigned char* handles256);
/* peer_id_base: blight_info.id_base of the peer's slice (owners answer with slice-local 32-bit ids on the stream path). */
int blight_part_session_connect_ipc(blight_part_session* s, uint32_t peer, const unsigned char* handles256, uint64_t peer_ids_capacity,
                                    uint64_t peer_id_base);
int blight_part_session_connect_local(blight_part_session* s, uint32_t peer, const blight_part_session* other);
/* DEVICE pointer of this rank's id array: after a query (and a synchronisation of its stream) slot d_kmer_off[r] + pos holds
 * the identifier query_sequence_hash would return for k-mer pos of read r (blight.cpp:575-591). */
void* blight_part_session_ids(const blight_part_session* s);
/* Sub-batches this rank cuts a batch of total_bases base positions into: ceil(total_bases / sub_positions). The n_sub of a
 * query is the maximum of this over the ranks. */
uint64_t blight_part_session_sub_batches(const blight_part_session* s, uint64_t total_bases, int want_ids);
/* One batch of reads held by THIS rank, collective: every rank calls it with the same n_sub (>= ceil(total_bases /
 * sub_positions) of every rank) and the same mode (d_kmer_off NULL everywhere = counting). Asynchronous on `stream`.
 * d_ctr accumulates QUERIES / INVALID for this rank's reads and FOUND / NOT_FOUND for the k-mers this rank OWNS: sum the
 * counters over the ranks. When the stream has drained, every identifier of this rank's reads is in its id array. */
int blight_part_session_query(blight_part_session* s, const char* d_bases, const uint64_t* d_read_off, const uint64_t* d_kmer_off,
                              uint64_t n_reads, uint64_t total_bases, uint64_t n_sub, uint64_t* d_ctr, void* stream);
/* BLIGHT_PART_* flags raised since the last reset (synchronises `stream`). */
int blight_part_session_status(blight_part_session* s, uint32_t* flags_out, int reset, void* stream);

/* Device buffers other processes of the box can map (CUDA IPC): alloc + 64-byte handle here, open there. */
int blight_peer_alloc(uint64_t bytes, void** d_ptr, unsigned char* handle64);
int blight_peer_open(const unsigned char* handle64, void** d_ptr);
int blight_peer_close(void* d_ptr);
int blight_peer_free(void* d_ptr);

/* ---- several GPUs of one box from ONE process: what a kmer_Set_Light drop-in holds instead of a single
 * blight_index when it is given more than one device. REPLICA: the whole index per device, a batch's reads cut into one
 * share per device (BASELINE configs[3]). PARTITION: MPHF groups cut into one contiguous range per device (balanced by
 * k-mer count; needs 2^n_log2 >= devices), super-k-mers routed to the owner of their minimizer bucket and identifiers
 * returned over NVLink by the kernels themselves (BASELINE configs[4]). Results are those of a single device holding the
 * whole index, identifier for identifier. A device may be listed more than once (tests on a one-GPU box). ------------- */
typedef struct blight_comm blight_comm;
#define BLIGHT_COMM_REPLICA 0
#define BLIGHT_COMM_PARTITION 1
typedef struct blight_comm_info {
	uint32_t n_gpus, mode;
	int32_t devices[BLIGHT_MAX_RANKS];
	uint64_t device_bytes[BLIGHT_MAX_RANKS]; /* HBM held per device */
	uint64_t kmers[BLIGHT_MAX_RANKS];        /* k-mers indexed per device (PARTITION: the slice; REPLICA: all) */
	uint32_t cuts[BLIGHT_MAX_RANKS + 1];     /* PARTITION: device g owns MPHF groups [cuts[g], cuts[g+1]) */
	blight_info whole;                       /* the index as a whole */
} blight_comm_info;
int blight_comm_init(const blight_flat* f, const int* devices, uint32_t n_gpus, int mode, const blight_upload_options* opts,
                     blight_comm** out);
void blight_comm_free(blight_comm* c);
int blight_comm_describe(const blight_comm* c, blight_comm_info* out);
/* The host-buffer entry points of a single index, spread over the devices (same arguments, same results). */
int blight_comm_query_reads_host(blight_comm* c, const char* bases, const uint64_t* read_off, uint64_t n_reads, int64_t* ids_out,
                                 uint64_t* ctr);
int blight_comm_query_fasta_host(blight_comm* c, const char* text, uint64_t len, uint64_t* ctr);
int blight_comm_query_file_host(blight_comm* c, const char* path, uint64_t* ctr);
int blight_comm_query_sequence_host(blight_comm* c, const char* seq, uint64_t len, int64_t* ids_out, uint64_t* n_out);

/* Test hook: the record cut of the streaming file_query (chunks of chunk_bytes, the unfinished tail carried over)
 * applied to a text in memory; records = sequence lines [beg, end) as file offsets, the reference's pairing rules
 * (blight.cpp:760-772). *n_out = records found (may exceed cap; only the first cap are stored). */
int blight_fasta_cut_stream(const char* text, uint64_t len, uint64_t chunk_bytes, uint64_t* beg_out, uint64_t* end_out,
                            uint64_t cap, uint64_t* n_out);
/* The same with the newline search done the way the streaming reader does it: the new bytes of every chunk scanned in
 * reader_slices consecutive slices (one per reader thread), the lists merged behind the carried tail (0: one scan per chunk). */
int blight_fasta_cut_stream_parts(const char* text, uint64_t len, uint64_t chunk_bytes, uint32_t reader_slices, uint64_t* beg_out,
                                  uint64_t* end_out, uint64_t cap, uint64_t* n_out);

/* Bytes the host-buffer entry points copied host->device and device->host in the calling process since load. */
void blight_transfer_bytes(uint64_t* h2d, uint64_t* d2h);
/* The host packer of blight_query_reads_host: bases it packed to 2 bits, wall time it spent packing (ns), threads it uses. */
void blight_host_pack_stats(uint64_t* packed_bases, uint64_t* pack_ns, uint32_t* threads);
/* Number of kernel launches issued by this library in the calling process (all threads) since load. */
uint64_t blight_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BLIGHT_B200_H */
