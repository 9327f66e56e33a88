// flat_index.cpp — BLFLAT01 blob save / load / validate, FASTA record splitting.
#include "flat_index.hpp"
#include "errors.hpp"
#include "capi_common.hpp"

#include <immintrin.h>
#include <omp.h>
#include <zlib.h>
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <thread>

namespace blight {

namespace {
void set_err(std::string* err, const std::string& s) { if (err) *err = s; }

template <class T>
bool write_vec(std::ofstream& os, const std::vector<T>& v, size_t pad_to = 8) {
	if (!v.empty()) os.write(reinterpret_cast<const char*>(v.data()), std::streamsize(v.size() * sizeof(T)));
	size_t bytes = v.size() * sizeof(T);
	static const char zeros[8] = {0};
	if (bytes % pad_to) os.write(zeros, std::streamsize(pad_to - bytes % pad_to));
	return os.good();
}

template <class T>
bool read_vec(std::ifstream& is, std::vector<T>& v, uint64_t n, size_t pad_to = 8) {
	v.resize(n);
	if (n) is.read(reinterpret_cast<char*>(v.data()), std::streamsize(n * sizeof(T)));
	size_t bytes = n * sizeof(T);
	if (bytes % pad_to) is.ignore(std::streamsize(pad_to - bytes % pad_to));
	return is.good();
}
}  // namespace

int flat_save(const FlatIndex& f, const std::string& path, std::string* err) {
	std::ofstream os(path, std::ios::binary);
	if (!os) { set_err(err, "cannot open " + path + " for writing"); return BL_ERR_IO; }
	os.write(reinterpret_cast<const char*>(&f.h), sizeof f.h);
	bool ok = write_vec(os, f.bucket_start) && write_vec(os, f.bucket_nuc) && write_vec(os, f.mphf) && write_vec(os, f.seq) &&
	          write_vec(os, f.pos) && write_vec(os, f.bits) && write_vec(os, f.ranks) && write_vec(os, f.fb_keys) &&
	          write_vec(os, f.fb_vals);
	os.flush();
	if (!ok || !os.good()) { set_err(err, "write failed: " + path); return BL_ERR_IO; }
	return BL_OK;
}

int flat_load(const std::string& path, FlatIndex& f, std::string* err) {
	std::ifstream is(path, std::ios::binary);
	if (!is) { set_err(err, "cannot open " + path); return BL_ERR_IO; }
	is.read(reinterpret_cast<char*>(&f.h), sizeof f.h);
	if (!is || std::memcmp(f.h.magic, "BLFLAT01", 8) != 0) { set_err(err, "not a BLFLAT01 blob: " + path); return BL_ERR_FORMAT; }
	const FlatHeader& h = f.h;
	if (h.m == 0 || h.m > 15 || h.n_buckets != (1ull << (2 * h.m - 1)) || h.n_log2 > 2 * h.m - 1 || h.n_mphf != (1ull << h.n_log2)) {
		set_err(err, "inconsistent header");
		return BL_ERR_FORMAT;
	}
	bool ok = read_vec(is, f.bucket_start, h.n_buckets) && read_vec(is, f.bucket_nuc, h.n_buckets) && read_vec(is, f.mphf, h.n_mphf) &&
	          read_vec(is, f.seq, h.seq_words) && read_vec(is, f.pos, h.pos_words) && read_vec(is, f.bits, h.bits_words_total) &&
	          read_vec(is, f.ranks, h.ranks_total) && read_vec(is, f.fb_keys, h.fallback_total) &&
	          read_vec(is, f.fb_vals, h.fallback_total);
	if (!ok) { set_err(err, "truncated blob: " + path); return BL_ERR_FORMAT; }
	return flat_validate(f, err);
}

int flat_validate(const FlatIndex& f, std::string* err) {
	// A blob may come from anywhere (load_index, the slices of the partitioned mode): everything the device code later uses
	// as an index or an extent is checked here, with arithmetic that cannot wrap.
	const FlatHeader& h = f.h;
	auto bad = [&](const char* what) { set_err(err, std::string("flat index invalid: ") + what); return BL_ERR_FORMAT; };
	auto mul_ok = [](uint64_t a, uint64_t b, uint64_t* out) { return !__builtin_mul_overflow(a, b, out); };
	auto add_ok = [](uint64_t a, uint64_t b, uint64_t* out) { return !__builtin_add_overflow(a, b, out); };
	if (h.k == 0 || h.k > 31) return bad("k");
	if ((h.m & 1) == 0 || h.m > 15 || h.m > h.k) return bad("m");
	if (h.n_log2 > 2 * h.m - 1) return bad("n");
	if (h.b > 24) return bad("b");  // check_params' limit
	if (f.bucket_start.size() != h.n_buckets || f.bucket_nuc.size() != h.n_buckets) return bad("bucket table size");
	if (f.mphf.size() != h.n_mphf) return bad("mphf table size");
	uint64_t t = 0;
	if (!mul_ok(h.total_nuc, 2, &t) || !add_ok(t, 63, &t) || f.seq.size() != h.seq_words || h.seq_words != t / 64) return bad("seq words");
	if (!add_ok(h.positions_bits, 63, &t) || f.pos.size() != h.pos_words || h.pos_words != t / 64) return bad("pos words");
	if (f.bits.size() != h.bits_words_total || f.ranks.size() != h.ranks_total) return bad("mphf arrays");
	if (f.fb_keys.size() != h.fallback_total || f.fb_vals.size() != h.fallback_total) return bad("fallback arrays");
	// buckets tile the text in order (the upload pass finds a position's bucket by binary search over the starts)
	for (uint64_t i = 0; i < h.n_buckets; i++) {
		uint64_t e = 0;
		if (!add_ok(f.bucket_start[i], f.bucket_nuc[i], &e) || e > h.total_nuc) return bad("bucket extent");
		if (i + 1 < h.n_buckets && f.bucket_start[i + 1] != e) return bad("bucket starts are not the running sum of the bucket lengths");
	}
	for (const MphfRec& r : f.mphf) {
		if (r.nbits == 0 || r.nbits > 32) return bad("mphf nbits");
		if (!r.present) continue;
		uint64_t e = 0;
		if (!add_ok(r.bits_word_off, r.bits_nwords, &e) || e > h.bits_words_total) return bad("mphf bits extent");
		if (!add_ok(r.ranks_off, r.nranks, &e) || e > h.ranks_total) return bad("mphf ranks extent");
		if (!add_ok(r.fb_off, r.fb_count, &e) || e > h.fallback_total) return bad("mphf fallback extent");
		uint64_t tot = 0;
		for (int l = 0; l < kLevels; l++) {
			if (r.dom[l] == 0 || (r.dom[l] & 63)) return bad("mphf level domain");
			if (!add_ok(tot, r.dom[l], &tot)) return bad("mphf level domain");
		}
		uint64_t nb = 0;
		if (!mul_ok(r.bits_nwords, 64, &nb) || tot != nb) return bad("mphf level domains do not sum to the bit array size");
		uint64_t pe = 0;
		if (!mul_ok(r.nelem, (uint64_t)r.nbits, &pe) || !add_ok(pe, r.pos_start, &pe) || pe > h.positions_bits) return bad("mphf positions extent");
		// a rank is (ones before a set bit) or a fallback value: both must index the group's position fields
		uint64_t ones = 0;
		for (uint64_t w = 0; w < r.bits_nwords; w++) ones += (uint64_t)__builtin_popcountll(f.bits[r.bits_word_off + w]);
		if (ones > r.nelem) return bad("more level bits set than keys");
		for (uint64_t j = 0; j < r.fb_count; j++)
			if (f.fb_vals[r.fb_off + j] >= r.nelem) return bad("fallback rank out of range");
	}
	return BL_OK;
}

// getline()-pairing of the reference (blight.cpp:212-229 / 760-775): line A is skipped as the header whatever it
// holds; an empty A swallows the next line too; an empty sequence line drops the record.
void split_fasta_records(const char* text, uint64_t len, std::vector<SeqView>& seqs) {
	uint64_t p = 0;
	bool eof = false;
	auto getline = [&](const char*& s, uint64_t& n) {
		if (p >= len) { s = text + len; n = 0; eof = true; return; }
		const char* nl = static_cast<const char*>(std::memchr(text + p, '\n', len - p));
		s = text + p;
		if (nl) { n = uint64_t(nl - s); p += n + 1; }
		else { n = len - p; p = len; eof = true; }
	};
	while (!eof) {
		const char* s; uint64_t n;
		getline(s, n);
		if (n == 0) { getline(s, n); continue; }
		getline(s, n);
		if (n == 0) continue;
		seqs.push_back(SeqView{s, n});
	}
}

namespace {
int cut_threads() { return std::max(omp_get_max_threads(), std::min(8, (int)std::thread::hardware_concurrency())); }  // OMP_NUM_THREADS=1 launchers
}  // namespace

void scan_newlines(const char* text, size_t lo, size_t hi, std::vector<uint64_t>& out, size_t& n) {
	size_t p = lo;
	const __m256i nlv = _mm256_set1_epi8('\n');
	for (; p + 32 <= hi; p += 32) {
		uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256(reinterpret_cast<const __m256i*>(text + p)), nlv));
		if (!m) continue;
		if (out.size() < n + 32) out.resize(std::max<size_t>(2 * out.size(), n + 4096));  // room for a block's worth, checked once per block
		uint64_t* o = out.data();
		do { o[n++] = p + (uint64_t)__builtin_ctz(m); m &= m - 1; } while (m);
	}
	if (out.size() < n + 32) out.resize(n + 4096);
	for (; p < hi; p++) if (text[p] == '\n') out[n++] = p;
}

size_t fasta_chunk_lines(const char* text, size_t len, bool eof, std::vector<uint64_t>& nl, int threads) {
	// Positions of every newline, all host threads: each scans its share 32 bytes at a time into a list it keeps from call to
	// call (a fresh vector per chunk cost more in page faults than the scan), then the lists are copied side by side into nl.
	// (First version: memchr per line + a serial merge, 3.3 ms per 64 MB chunk — the slowest stage of the streaming file_query.)
	// `threads`: the team's size when the caller runs other threads beside it (the streaming reader): a team larger than the
	// cores left over spends its time at the region's barriers waiting for descheduled members
	const int T = threads > 0 ? threads : cut_threads();
	std::vector<size_t> cnt(T + 1, 0);
	#pragma omp parallel num_threads(T)
	{
		static thread_local std::vector<uint64_t> mine;
		const int t = omp_get_thread_num(), nt = omp_get_num_threads();
		size_t n = 0;
		for (int tt = t; tt < T; tt += nt) {  // (a runtime may hand out fewer threads than asked for)
			scan_newlines(text, len * tt / T, len * (tt + 1) / T, mine, n);
			cnt[tt + 1] = n;  // cumulative over this thread's shares; un-cumulated below when nt < T
		}
		#pragma omp barrier
		#pragma omp single
		{
			if (nt < T) {
				// shares of one thread sit back to back in its list: cnt[tt+1] holds the list size after share tt
				std::vector<size_t> own(T + 1, 0);
				for (int tt = 0; tt < T; tt++) own[tt + 1] = cnt[tt + 1] - (tt >= nt ? cnt[tt + 1 - nt] : 0);
				cnt = own;
			}
			for (int tt = 0; tt < T; tt++) cnt[tt + 1] += cnt[tt];
			nl.resize(cnt[T] + 1);
		}
		size_t off = 0;
		for (int tt = t; tt < T; tt += nt) {
			const size_t c = cnt[tt + 1] - cnt[tt];
			if (c) std::memcpy(nl.data() + cnt[tt], mine.data() + off, c * 8);
			off += c;
		}
	}
	nl.resize(cnt[T]);
	if (eof && len > 0 && text[len - 1] != '\n') nl.push_back(len);  // the last line needs no terminator
	return nl.size() / 2;
}

size_t fasta_chunk_lines_merge(const char* text, size_t head_len, size_t len, bool eof, const std::vector<uint64_t>* parts, const size_t* counts,
                               int n_parts, std::vector<uint64_t>& nl, int threads) {
	// the reader scanned the new data slice by slice while it was still in its cores' caches (positions relative to the start
	// of the new data); only the carried head is scanned here, and the lists are copied side by side, shifted by the head
	static thread_local std::vector<uint64_t> head;
	size_t nh = 0;
	scan_newlines(text, 0, head_len, head, nh);
	std::vector<size_t> off(n_parts + 1, nh);
	for (int i = 0; i < n_parts; i++) off[i + 1] = off[i] + counts[i];
	nl.resize(off[n_parts] + 1);
	if (nh) std::memcpy(nl.data(), head.data(), nh * 8);
	const int T = threads > 0 ? threads : cut_threads();
	#pragma omp parallel for num_threads(T) schedule(dynamic, 1)
	for (int i = 0; i < n_parts; i++) {
		const uint64_t* src = parts[i].data();
		uint64_t* dst = nl.data() + off[i];
		for (size_t j = 0; j < counts[i]; j++) dst[j] = src[j] + head_len;
	}
	nl.resize(off[n_parts]);
	if (eof && len > 0 && text[len - 1] != '\n') nl.push_back(len);
	return nl.size() / 2;
}

ChunkCut fasta_chunk_records(size_t len, bool eof, const std::vector<uint64_t>& nl, uint64_t* beg, uint64_t* end, int threads) {
	const size_t n_pairs = nl.size() / 2;
	ChunkCut c{0, eof ? len : (n_pairs ? size_t(nl[2 * n_pairs - 1]) + 1 : 0)};
	// record j = line pair (2j, 2j+1); an empty header swallows its line, an empty sequence drops the record. Two parallel
	// passes: count the records of each share, then write them at the share's offset.
	auto record = [&](size_t j, uint64_t& ss, uint64_t& se) {
		const uint64_t hs = j ? nl[2 * j - 1] + 1 : 0, he = nl[2 * j];
		ss = he + 1; se = nl[2 * j + 1];
		return he > hs && se > ss;
	};
	const int T = n_pairs < 4096 ? 1 : (threads > 0 ? threads : cut_threads());
	std::vector<size_t> cnt(T + 1, 0);
	#pragma omp parallel for num_threads(T) schedule(static, 1)
	for (int t = 0; t < T; t++) {
		size_t n = 0;
		uint64_t ss, se;
		for (size_t j = n_pairs * t / T; j < n_pairs * (t + 1) / T; j++) n += record(j, ss, se);
		cnt[t + 1] = n;
	}
	for (int t = 0; t < T; t++) cnt[t + 1] += cnt[t];
	#pragma omp parallel for num_threads(T) schedule(static, 1)
	for (int t = 0; t < T; t++) {
		size_t o = cnt[t];
		uint64_t ss, se;
		for (size_t j = n_pairs * t / T; j < n_pairs * (t + 1) / T; j++)
			if (record(j, ss, se)) { beg[o] = ss; end[o] = se; o++; }
	}
	c.n_rec = cnt[T];
	return c;
}

int read_fasta_records(const std::string& path, std::string& storage, std::vector<SeqView>& seqs, std::string* err) {
	gzFile gz = gzopen(path.c_str(), "rb");  // transparent for plain files, like zstr::ifstream (zstr.hpp:153-167)
	if (!gz) { set_err(err, "Problem with files opening: " + path); return BL_ERR_IO; }
	gzbuffer(gz, 1 << 20);
	storage.clear();
	std::vector<char> buf(1 << 24);
	for (;;) {
		int got = gzread(gz, buf.data(), unsigned(buf.size()));
		if (got < 0) { gzclose(gz); set_err(err, "read error: " + path); return BL_ERR_IO; }
		if (got == 0) break;
		storage.append(buf.data(), size_t(got));
	}
	gzclose(gz);
	split_fasta_records(storage.data(), storage.size(), seqs);
	return BL_OK;
}


void fill_info(const FlatIndex& f, blight_info* out) {
	std::memset(out, 0, sizeof *out);
	const FlatHeader& h = f.h;
	out->k = h.k; out->m = h.m; out->n_log2 = h.n_log2; out->s_log2 = h.s_log2; out->b = h.b;
	out->n_buckets = h.n_buckets;
	out->n_mphf = h.n_mphf;
	out->number_kmer = h.number_kmer;
	out->number_super_kmer = h.number_super_kmer;
	out->total_nuc = h.total_nuc;
	out->positions_bits = h.positions_bits;
	out->mphf_bits = h.bits_words_total * 64;
	out->fallback_keys = h.fallback_total;
	for (const MphfRec& r : f.mphf) out->largest_mphf = r.nelem > out->largest_mphf ? r.nelem : out->largest_mphf;
	for (uint32_t n : f.bucket_nuc) out->largest_bucket = n > out->largest_bucket ? n : out->largest_bucket;
	uint64_t base = ~0ull;
	for (const MphfRec& r : f.mphf) if (r.present && r.id_offset < base) base = r.id_offset;
	out->id_base = base == ~0ull ? 0 : base;
}

namespace {
// copies `nbits` bits from src starting at bit `sbit` to dst starting at bit `dbit` (dst pre-zeroed there)
void copy_bits(const std::vector<uint64_t>& src, uint64_t sbit, std::vector<uint64_t>& dst, uint64_t dbit, uint64_t nbits) {
	while (nbits) {
		const unsigned so = unsigned(sbit & 63), dof = unsigned(dbit & 63);
		unsigned take = 64 - (so > dof ? so : dof);
		if (take > nbits) take = unsigned(nbits);
		const uint64_t mask = take == 64 ? ~0ull : ((1ull << take) - 1);
		const uint64_t v = (src[sbit >> 6] >> so) & mask;
		dst[dbit >> 6] |= v << dof;
		sbit += take; dbit += take; nbits -= take;
	}
}
}  // namespace

int flat_slice(const FlatIndex& f, uint64_t g0, uint64_t g1, FlatIndex& o, std::string* err) {
	const FlatHeader& h = f.h;
	const unsigned lb = f.lb();
	o = FlatIndex();
	o.h = h;
	o.bucket_start.assign(h.n_buckets, 0);
	o.bucket_nuc.assign(h.n_buckets, 0);
	o.mphf.assign(h.n_mphf, MphfRec{});
	const uint64_t b0 = g0 << lb, b1 = g1 << lb;
	// buckets of a group range are contiguous in seq
	const uint64_t nuc0 = b0 < h.n_buckets ? f.bucket_start[b0] : h.total_nuc;
	const uint64_t nuc1 = b1 < h.n_buckets ? f.bucket_start[b1] : h.total_nuc;
	// the 2^b-window scan may run past the end of a bucket into what follows (blight.cpp:729-739): keep that tail
	// so a slice answers exactly like the whole index
	const uint64_t tail = std::min<uint64_t>(h.total_nuc - nuc1, (1ull << h.b) + h.k);
	o.h.total_nuc = nuc1 - nuc0 + tail;
	o.h.seq_words = (o.h.total_nuc * 2 + 63) / 64;
	o.seq.assign(o.h.seq_words, 0);
	copy_bits(f.seq, nuc0 * 2, o.seq, 0, o.h.total_nuc * 2);
	for (uint64_t b = b0; b < b1; b++) { o.bucket_start[b] = f.bucket_start[b] - nuc0; o.bucket_nuc[b] = f.bucket_nuc[b]; }
	for (uint64_t b = 0; b < h.n_buckets; b++) if (b < b0) o.bucket_start[b] = 0; else if (b >= b1) o.bucket_start[b] = nuc1 - nuc0;
	// positions slabs of the range are contiguous too
	const uint64_t p0 = g0 < h.n_mphf ? f.mphf[g0].pos_start : h.positions_bits;
	const uint64_t p1 = g1 < h.n_mphf ? f.mphf[g1].pos_start : h.positions_bits;
	o.h.positions_bits = p1 - p0;
	o.h.pos_words = (o.h.positions_bits + 63) / 64;
	o.pos.assign(o.h.pos_words, 0);
	copy_bits(f.pos, p0, o.pos, 0, o.h.positions_bits);
	o.h.number_kmer = 0;
	for (uint64_t g = 0; g < h.n_mphf; g++) {
		MphfRec& r = o.mphf[g];
		if (g < g0 || g >= g1) {
			r = MphfRec{};
			r.nbits = f.mphf[g].nbits;
			r.id_offset = f.mphf[g].id_offset;
			r.pos_start = g < g0 ? 0 : o.h.positions_bits;
			continue;
		}
		r = f.mphf[g];
		r.pos_start -= p0;
		o.h.number_kmer += r.nelem;
		if (!r.present) continue;
		r.bits_word_off = o.bits.size();
		o.bits.insert(o.bits.end(), f.bits.begin() + f.mphf[g].bits_word_off, f.bits.begin() + f.mphf[g].bits_word_off + r.bits_nwords);
		r.ranks_off = o.ranks.size();
		o.ranks.insert(o.ranks.end(), f.ranks.begin() + f.mphf[g].ranks_off, f.ranks.begin() + f.mphf[g].ranks_off + r.nranks);
		r.fb_off = o.fb_keys.size();
		o.fb_keys.insert(o.fb_keys.end(), f.fb_keys.begin() + f.mphf[g].fb_off, f.fb_keys.begin() + f.mphf[g].fb_off + r.fb_count);
		o.fb_vals.insert(o.fb_vals.end(), f.fb_vals.begin() + f.mphf[g].fb_off, f.fb_vals.begin() + f.mphf[g].fb_off + r.fb_count);
	}
	o.h.bits_words_total = o.bits.size();
	o.h.ranks_total = o.ranks.size();
	o.h.fallback_total = o.fb_keys.size();
	return flat_validate(o, err);
}

}  // namespace blight
