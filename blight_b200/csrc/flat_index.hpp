// flat_index.hpp — host-side flat image of a Blight index ("BLFLAT01").
//
// This is the interchange point of the drop-in: it holds, as plain arrays, exactly the state
// the reference keeps in kmer_Set_Light (blight.h:15-60) after construct_index — bucket table,
// bit-packed bucket sequences, bit-packed positions, and per-MPHF BBHash state (bbhash.h:777-786)
// — in the reference's own bit conventions.  It is produced either by the
// product's host builder (builder.cpp) or by an exporter attached to a reference object, is
// saved/loaded as one little-endian blob, and is re-laid-out for the GPU by device_index.cu.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace blight {

constexpr int kLevels = 16;  // boomphf::mphf<_nb_levels=16> (bbhash.h:512)

struct MphfRec {
	uint64_t id_offset;      // info_mphf::mphf_size after construction = exclusive prefix of k-mer counts (blight.cpp:298-301)
	uint64_t pos_start;      // info_mphf::start, bit offset into positions
	uint64_t nelem;          // keys in this MPHF
	uint64_t bits_word_off;  // offset (u64 words) of this MPHF's level bit array in FlatIndex::bits
	uint64_t bits_nwords;
	uint64_t ranks_off;      // offset into FlatIndex::ranks (one sample per 16 words, bbhash.h:447-465)
	uint64_t nranks;
	uint64_t fb_off;         // fallback map slice (sorted by key) in fb_keys/fb_vals
	uint64_t fb_count;
	uint32_t nbits;          // info_mphf::bit_to_encode
	uint32_t present;        // kmer_MPHF != nullptr
	uint64_t dom[kLevels];   // _hash_domains
};
static_assert(sizeof(MphfRec) == 9 * 8 + 8 + kLevels * 8, "MphfRec is part of the blob format");

struct FlatHeader {
	char magic[8];  // "BLFLAT01"
	uint32_t k, m, n_log2, s_log2, b, r0, r1, r2;
	uint64_t n_buckets, n_mphf, number_kmer, number_super_kmer, total_nuc, positions_bits;
	uint64_t seq_words, pos_words, bits_words_total, ranks_total, fallback_total;
};
static_assert(sizeof(FlatHeader) == 8 + 32 + 11 * 8, "FlatHeader is part of the blob format");

struct FlatIndex {
	FlatHeader h{};
	std::vector<uint64_t> bucket_start;  // nucleotide offset of each bucket in seq
	std::vector<uint32_t> bucket_nuc;    // bucket length in nucleotides (nuc_minimizer)
	std::vector<MphfRec> mphf;
	std::vector<uint64_t> seq;   // vector<bool> image of bucketSeq: nucleotide p -> bit 2p = code>>1, bit 2p+1 = code&1
	std::vector<uint64_t> pos;   // vector<bool> image of positions (LSB-first fields)
	std::vector<uint64_t> bits;  // concatenated MPHF level bit arrays
	std::vector<uint64_t> ranks; // concatenated rank samples
	std::vector<uint64_t> fb_keys, fb_vals;

	unsigned lb() const { return 2 * h.m - 1 - h.n_log2; }  // log2(buckets per MPHF)
};

// 0 on success; negative error codes otherwise (message in *err if given).
int flat_save(const FlatIndex& f, const std::string& path, std::string* err = nullptr);
int flat_load(const std::string& path, FlatIndex& f, std::string* err = nullptr);
// Structural validation (sizes, offsets in range). 0 if consistent.
int flat_validate(const FlatIndex& f, std::string* err = nullptr);

struct BuildParams {
	unsigned k = 31, m = 9, n_log2 = 17, s_log2 = 6, b = 6;  // bench_blight.cpp:41-45 defaults
	unsigned threads = 1;
};

// One input sequence (a unitig) as a view into caller memory.
struct SeqView { const char* p; uint64_t len; };

// Validates like the kmer_Set_Light constructor (blight.h:75-92). Returns 0 or an error code.
int check_params(const BuildParams& p, std::string* err = nullptr);

// Builds the index the reference's construct_index (blight.cpp:108-125, cores=1) would build on the same
// sequences in the same order, bit for bit.  Sequences shorter than k are skipped (undefined in the reference).
int build_flat_index(const std::vector<SeqView>& seqs, const BuildParams& p, FlatIndex& out, std::string* err = nullptr);

// Reads a 2-line-record FASTA (plain or gzip) with the reference's record pairing (blight.cpp:212-229) into
// `storage` and returns views on the sequence lines.
int read_fasta_records(const std::string& path, std::string& storage, std::vector<SeqView>& seqs, std::string* err = nullptr);
// Same pairing rule on an in-memory text buffer.
void split_fasta_records(const char* text, uint64_t len, std::vector<SeqView>& seqs);

// The same pairing on one chunk of a stream (stream_query.cu). Every iteration of the reference's loop consumes exactly
// two lines, so records are the line pairs (2j, 2j+1) counted from the start of the file: the cut is a parallel newline
// search plus one pass over the pairs. text[0, len) must start at an even line (the carry of the previous chunk in
// front); only complete pairs are consumed unless `eof`, where the last line needs no terminator.
// Two steps so that the caller can size its arrays: fasta_chunk_lines finds the line ends (all host threads) and returns
// the number of line pairs, an upper bound of the records; fasta_chunk_records fills beg[0..n_rec) / end[0..n_rec) with
// the sequence lines (offsets into text).
struct ChunkCut { size_t n_rec, consumed; };
// newline offsets (from `text`) of text[lo, hi) written to out[n...], n advanced; out grows as needed and is never shrunk
void scan_newlines(const char* text, size_t lo, size_t hi, std::vector<uint64_t>& out, size_t& n);
// fasta_chunk_lines when the new data (text + head_len .. text + len) was already scanned in n_parts consecutive slices:
// parts[i][0..counts[i]) = newline offsets from the start of the NEW data; the carried head text[0, head_len) is scanned here
size_t fasta_chunk_lines_merge(const char* text, size_t head_len, size_t len, bool eof, const std::vector<uint64_t>* parts, const size_t* counts,
                               int n_parts, std::vector<uint64_t>& nl, int threads = 0);
// threads: size of the OpenMP team (0: all host threads)
size_t fasta_chunk_lines(const char* text, size_t len, bool eof, std::vector<uint64_t>& nl, int threads = 0);
ChunkCut fasta_chunk_records(size_t len, bool eof, const std::vector<uint64_t>& nl, uint64_t* beg, uint64_t* end, int threads = 0);

}  // namespace blight
