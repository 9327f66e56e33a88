// capi_host.cpp — the host half of the C ABI (include/blight_b200.h): flat index construction,
// persistence, comparison and partition slicing.  No CUDA in this file.
#include "capi_common.hpp"

#include <omp.h>
#include <cstring>

using namespace blight;

namespace blight {
thread_local std::string g_last_error;
int fail(int code, const std::string& msg) { g_last_error = msg; return code; }
}  // namespace blight

extern "C" {

const char* blight_version(void) { return "blight_b200 0.1 (sm_100a)"; }
const char* blight_last_error(void) { return g_last_error.c_str(); }

int blight_check_params(uint32_t k, uint32_t m, uint32_t n_log2, uint32_t s_log2, uint32_t b) {
	BuildParams p; p.k = k; p.m = m; p.n_log2 = n_log2; p.s_log2 = s_log2; p.b = b;
	std::string err;
	int rc = check_params(p, &err);
	return rc == BL_OK ? rc : fail(rc, err);
}

int blight_flat_build_seqs(const char* bases, const uint64_t* offsets, uint64_t n_seqs, uint32_t k, uint32_t m,
                           uint32_t n_log2, uint32_t s_log2, uint32_t b, uint32_t threads, blight_flat** out) {
	if (!out || (n_seqs && (!bases || !offsets))) return fail(BL_ERR_INVALID_ARG, "null argument");
	BuildParams p; p.k = k; p.m = m; p.n_log2 = n_log2; p.s_log2 = s_log2; p.b = b; p.threads = threads;
	std::vector<SeqView> seqs(n_seqs);
	for (uint64_t i = 0; i < n_seqs; i++) seqs[i] = SeqView{bases + offsets[i], offsets[i + 1] - offsets[i]};
	blight_flat* f = new blight_flat();
	std::string err;
	int rc = build_flat_index(seqs, p, f->f, &err);
	if (rc != BL_OK) { delete f; return fail(rc, err); }
	*out = f;
	return BL_OK;
}

int blight_flat_build_spans(const char* bases, const uint64_t* starts, const uint64_t* lengths, uint64_t n_seqs, uint32_t k,
                            uint32_t m, uint32_t n_log2, uint32_t s_log2, uint32_t b, uint32_t threads, blight_flat** out) {
	if (!out || (n_seqs && (!bases || !starts || !lengths))) return fail(BL_ERR_INVALID_ARG, "null argument");
	BuildParams p; p.k = k; p.m = m; p.n_log2 = n_log2; p.s_log2 = s_log2; p.b = b; p.threads = threads;
	std::vector<SeqView> seqs(n_seqs);
	for (uint64_t i = 0; i < n_seqs; i++) seqs[i] = SeqView{bases + starts[i], lengths[i]};
	blight_flat* f = new blight_flat();
	std::string err;
	int rc = build_flat_index(seqs, p, f->f, &err);
	if (rc != BL_OK) { delete f; return fail(rc, err); }
	*out = f;
	return BL_OK;
}

int blight_flat_build_file(const char* unitig_path, uint32_t k, uint32_t m, uint32_t n_log2, uint32_t s_log2, uint32_t b,
                           uint32_t threads, blight_flat** out) {
	if (!out || !unitig_path) return fail(BL_ERR_INVALID_ARG, "null argument");
	BuildParams p; p.k = k; p.m = m; p.n_log2 = n_log2; p.s_log2 = s_log2; p.b = b; p.threads = threads;
	std::string err;
	int rc = check_params(p, &err);
	if (rc != BL_OK) return fail(rc, err);
	std::string storage;
	std::vector<SeqView> seqs;
	rc = read_fasta_records(unitig_path, storage, seqs, &err);
	if (rc != BL_OK) return fail(rc, err);
	blight_flat* f = new blight_flat();
	rc = build_flat_index(seqs, p, f->f, &err);
	if (rc != BL_OK) { delete f; return fail(rc, err); }
	*out = f;
	return BL_OK;
}

int blight_flat_save(const blight_flat* f, const char* path) {
	if (!f || !path) return fail(BL_ERR_INVALID_ARG, "null argument");
	std::string err;
	int rc = flat_save(f->f, path, &err);
	return rc == BL_OK ? rc : fail(rc, err);
}

int blight_flat_load(const char* path, blight_flat** out) {
	if (!out || !path) return fail(BL_ERR_INVALID_ARG, "null argument");
	blight_flat* f = new blight_flat();
	std::string err;
	int rc = flat_load(path, f->f, &err);
	if (rc != BL_OK) { delete f; return fail(rc, err); }
	*out = f;
	return BL_OK;
}

void blight_flat_free(blight_flat* f) { delete f; }

int blight_flat_info(const blight_flat* f, blight_info* out) {
	if (!f || !out) return fail(BL_ERR_INVALID_ARG, "null argument");
	fill_info(f->f, out);
	return BL_OK;
}

int blight_flat_compare(const blight_flat* a, const blight_flat* b) {
	if (!a || !b) return fail(BL_ERR_INVALID_ARG, "null argument");
	const FlatIndex& A = a->f; const FlatIndex& B = b->f;
	auto diff = [&](const char* what) { g_last_error = std::string("flat indices differ in ") + what; return 1; };
	if (std::memcmp(&A.h, &B.h, sizeof A.h) != 0) return diff("header");
	if (A.bucket_start != B.bucket_start) return diff("bucket_start");
	if (A.bucket_nuc != B.bucket_nuc) return diff("bucket_nuc");
	if (A.mphf.size() != B.mphf.size()) return diff("mphf count");
	for (size_t i = 0; i < A.mphf.size(); i++)
		if (std::memcmp(&A.mphf[i], &B.mphf[i], sizeof(MphfRec)) != 0) return diff(("mphf record " + std::to_string(i)).c_str());
	if (A.seq != B.seq) return diff("bucket sequences");
	if (A.bits != B.bits) return diff("mphf bit arrays");
	if (A.ranks != B.ranks) return diff("mphf ranks");
	if (A.fb_keys != B.fb_keys || A.fb_vals != B.fb_vals) return diff("mphf fallback");
	if (A.pos != B.pos) return diff("positions");
	return 0;
}

int blight_flat_group_sizes(const blight_flat* f, uint64_t* sizes_out) {
	if (!f || !sizes_out) return fail(BL_ERR_INVALID_ARG, "null argument");
	for (size_t g = 0; g < f->f.mphf.size(); g++) sizes_out[g] = f->f.mphf[g].nelem;
	return BL_OK;
}

int blight_flat_slice(const blight_flat* f, uint64_t g_begin, uint64_t g_end, blight_flat** out) {
	if (!f || !out) return fail(BL_ERR_INVALID_ARG, "null argument");
	if (g_begin > g_end || g_end > f->f.h.n_mphf) return fail(BL_ERR_INVALID_ARG, "group range out of bounds");
	blight_flat* r = new blight_flat();
	std::string err;
	int rc = flat_slice(f->f, g_begin, g_end, r->f, &err);
	if (rc != BL_OK) { delete r; return fail(rc, err); }
	*out = r;
	return BL_OK;
}

int blight_fasta_cut_stream(const char* text, uint64_t len, uint64_t chunk_bytes, uint64_t* beg_out, uint64_t* end_out, uint64_t cap,
                            uint64_t* n_out) {
	return blight_fasta_cut_stream_parts(text, len, chunk_bytes, 0, beg_out, end_out, cap, n_out);
}

int blight_fasta_cut_stream_parts(const char* text, uint64_t len, uint64_t chunk_bytes, uint32_t reader_slices, uint64_t* beg_out,
                                  uint64_t* end_out, uint64_t cap, uint64_t* n_out) {
	if (!n_out || (len && !text) || chunk_bytes == 0 || reader_slices > 4096) return fail(BL_ERR_INVALID_ARG, "bad argument");
	// the loop of stream_file_query (stream_query.cu) on a text in memory: chunks of chunk_bytes, unfinished tail carried.
	// reader_slices > 0: as the streaming reader does it — the NEW bytes of every chunk are scanned for newlines in that many
	// consecutive slices (offsets from the start of the new data), the carried head is scanned at the merge.
	std::vector<char> buf;
	std::vector<uint64_t> nl, beg, end;
	std::vector<std::vector<uint64_t>> parts(reader_slices);
	std::vector<size_t> part_n(reader_slices);
	uint64_t n = 0, file_off = 0, buf_file_off = 0;
	for (;;) {
		const uint64_t got = std::min<uint64_t>(chunk_bytes, len - file_off);
		const size_t head = buf.size();
		buf.insert(buf.end(), text + file_off, text + file_off + got);
		file_off += got;
		const bool eof = file_off >= len;
		size_t n_pairs;
		if (reader_slices) {
			for (uint32_t t = 0; t < reader_slices; t++) {
				size_t cnt = 0;
				scan_newlines(buf.data() + head, got * t / reader_slices, got * (t + 1) / reader_slices, parts[t], cnt);
				part_n[t] = cnt;
			}
			n_pairs = fasta_chunk_lines_merge(buf.data(), head, buf.size(), eof, parts.data(), part_n.data(), (int)reader_slices, nl);
		} else {
			n_pairs = fasta_chunk_lines(buf.data(), buf.size(), eof, nl);
		}
		beg.resize(n_pairs + 1); end.resize(n_pairs + 1);
		const ChunkCut cut = fasta_chunk_records(buf.size(), eof, nl, beg.data(), end.data());
		for (size_t i = 0; i < cut.n_rec; i++, n++)
			if (n < cap && beg_out && end_out) { beg_out[n] = buf_file_off + beg[i]; end_out[n] = buf_file_off + end[i]; }
		buf.erase(buf.begin(), buf.begin() + cut.consumed);
		buf_file_off += cut.consumed;
		if (eof) break;
	}
	*n_out = n;
	return BL_OK;
}

}  // extern "C"
