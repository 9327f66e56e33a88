// oracle/ref_harness.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A thin extern "C" shim around the UNMODIFIED reference sources, which are compiled
// where they lie (a scratch copy carrying the two documented one-line fixes P1/P2, see
// oracle/build_ref.sh and SURVEY.md F3).  The reference translation unit is pulled in
// by #include so that its `inline` helpers (get_kmer, update_kmer, query_get_pos_unitig;
// blight.cpp:387,406,686) are visible; the build uses -fno-access-control so the private
// state of boomphf::mphf (bbhash.h:777-786) can be read and written by the exporter /
// importer below without touching the reference.
//
// What it offers to tests/ and to bench.py's cpu_baseline / --impl reference legs:
//   * construct an index with the reference's own construct_index (blight.cpp:108)
//   * per-k-mer / per-sequence queries through the reference's own query functions
//   * export of the reference's in-memory index to the flat "BLFLAT01" blob
//   * import of such a blob back into a reference object (so the reference query code
//     can be timed on an index produced by the product's fast host builder)
//   * an OpenMP loop over query_sequence_bool on pre-loaded reads (the in-memory variant
//     of file_query, blight.cpp:746-799, without its racy TP/FP totals)
#include "blight.cpp"

#include <unistd.h>
#include <cstring>
#include <cstdio>
#include <fstream>
#include <algorithm>
#include <sstream>

namespace {

struct Handle {
	kmer_Set_Light* ksl;
	unsigned k, m, n, s, b;
};

struct CoutSilencer {
	std::streambuf* old;
	std::ostringstream sink;
	explicit CoutSilencer(bool on) : old(nullptr) { if(on) old = std::cout.rdbuf(sink.rdbuf()); }
	~CoutSilencer() { if(old) std::cout.rdbuf(old); }
};

uint64_t* vb_words(std::vector<bool>& v) { return v.begin()._M_p; }

struct MphfRec {
	uint64_t id_offset, pos_start, nelem, bits_word_off, bits_nwords, ranks_off, nranks, fb_off, fb_count;
	uint32_t nbits, present;
	uint64_t dom[16];
};
static_assert(sizeof(MphfRec) == 9*8 + 8 + 16*8, "MphfRec layout");

struct Header {
	char magic[8];
	uint32_t k, m, n_log2, s_log2, b, r0, r1, r2;
	uint64_t n_buckets, n_mphf, number_kmer, number_super_kmer, total_nuc, positions_bits;
	uint64_t seq_words, pos_words, bits_words_total, ranks_total, fallback_total;
};
static_assert(sizeof(Header) == 8 + 32 + 11*8, "Header layout");

} // namespace

extern "C" {

void* blref_create(unsigned k, unsigned m, unsigned n, unsigned s, unsigned cores, unsigned b) {
	try {
		Handle* h = new Handle{nullptr, k, m, n, s, b};
		h->ksl = new kmer_Set_Light(k, m, n, s, cores, b);
		return h;
	} catch(const std::exception& e) {
		return nullptr;
	}
}

void blref_destroy(void* hv) {
	Handle* h = static_cast<Handle*>(hv);
	if(!h) return;
	delete h->ksl;
	delete h;
}

// construct_index writes and deletes "_out<i>" in the CWD (blight.cpp:132,352): run it in workdir.
int blref_construct(void* hv, const char* unitig_path, const char* workdir, int quiet) {
	Handle* h = static_cast<Handle*>(hv);
	char cwd[4096];
	if(!getcwd(cwd, sizeof cwd)) return -2;
	if(chdir(workdir) != 0) return -3;
	int rc = 0;
	try {
		CoutSilencer sil(quiet != 0);
		h->ksl->construct_index(unitig_path);
	} catch(const std::exception& e) {
		std::fprintf(stderr, "blref_construct: %s\n", e.what());
		rc = -1;
	}
	if(chdir(cwd) != 0) return -4;
	return rc;
}

uint64_t blref_number_kmer(void* hv) { return static_cast<Handle*>(hv)->ksl->number_kmer; }
uint64_t blref_number_super_kmer(void* hv) { return static_cast<Handle*>(hv)->ksl->number_super_kmer; }
uint64_t blref_number_query(void* hv) { return static_cast<Handle*>(hv)->ksl->number_query; }
uint64_t blref_largest_mphf(void* hv) { return static_cast<Handle*>(hv)->ksl->largest_MPHF; }
uint64_t blref_largest_bucket(void* hv) { return static_cast<Handle*>(hv)->ksl->largest_bucket_nuc_all; }

uint32_t blref_minimizer(uint64_t canon, unsigned k, unsigned m) { return minimizer_naive(canon, k, m); }

// returns number of ids written, -1 on invalid base (nuc2int throws, kmer.h:68)
int64_t blref_query_sequence_hash(void* hv, const char* seq, uint64_t len, int64_t* out, uint64_t cap) {
	Handle* h = static_cast<Handle*>(hv);
	try {
		std::vector<int64_t> r = h->ksl->query_sequence_hash(std::string(seq, len));
		if(r.size() > cap) return -2;
		std::copy(r.begin(), r.end(), out);
		return int64_t(r.size());
	} catch(const std::domain_error&) {
		return -1;
	}
}

int blref_query_sequence_bool(void* hv, const char* seq, uint64_t len, uint32_t* found, uint32_t* not_found) {
	Handle* h = static_cast<Handle*>(hv);
	try {
		auto p = h->ksl->query_sequence_bool(std::string(seq, len));
		*found = p.first; *not_found = p.second;
		return 0;
	} catch(const std::domain_error&) {
		return -1;
	}
}

void blref_query_kmers_hash(void* hv, const uint64_t* canon, uint64_t n, int64_t* out, int threads) {
	Handle* h = static_cast<Handle*>(hv);
	#pragma omp parallel for num_threads(threads) schedule(static, 4096)
	for(uint64_t i = 0; i < n; i++) out[i] = h->ksl->query_kmer_hash(canon[i]);
}

void blref_query_kmers_bool(void* hv, const uint64_t* canon, uint64_t n, uint8_t* out, int threads) {
	Handle* h = static_cast<Handle*>(hv);
	#pragma omp parallel for num_threads(threads) schedule(static, 4096)
	for(uint64_t i = 0; i < n; i++) out[i] = h->ksl->query_kmer_bool(canon[i]) ? 1 : 0;
}

// In-memory batched query: reads r = bases[offs[r] .. offs[r+1]) (no separators), the same
// skip rule as file_query (size >= k, blight.cpp:782).  ids_out may be null (bool mode);
// otherwise ids for read r land at ids_out[kmer_offs[r]..].  Returns seconds spent.
double blref_query_reads(void* hv, const char* bases, const uint64_t* offs, uint64_t n_reads, int threads,
                         int64_t* ids_out, const uint64_t* kmer_offs, uint64_t* found, uint64_t* not_found) {
	Handle* h = static_cast<Handle*>(hv);
	uint64_t tp = 0, fp = 0;
	const unsigned k = h->k;
	double t0 = omp_get_wtime();
	#pragma omp parallel for num_threads(threads) schedule(dynamic, 512) reduction(+:tp,fp)
	for(uint64_t r = 0; r < n_reads; r++) {
		const uint64_t len = offs[r+1] - offs[r];
		if(len < k) continue;
		std::string q(bases + offs[r], len);
		if(ids_out) {
			std::vector<int64_t> v = h->ksl->query_sequence_hash(q);
			int64_t* dst = ids_out + kmer_offs[r];
			for(size_t i = 0; i < v.size(); i++) { dst[i] = v[i]; if(v[i] >= 0) tp++; else fp++; }
		} else {
			auto p = h->ksl->query_sequence_bool(q);
			tp += p.first; fp += p.second;
		}
	}
	double t1 = omp_get_wtime();
	*found = tp; *not_found = fp;
	return t1 - t0;
}

// The reference's own file_query (prints its recap to stdout). Uses the `cores` given at create.
int blref_file_query(void* hv, const char* path, int quiet) {
	Handle* h = static_cast<Handle*>(hv);
	try {
		CoutSilencer sil(quiet != 0);
		h->ksl->file_query(path);
		return 0;
	} catch(const std::exception& e) {
		std::fprintf(stderr, "blref_file_query: %s\n", e.what());
		return -1;
	}
}

// ---- export: reference object -> BLFLAT01 blob ------------------------------------------------
int blref_export(void* hv, const char* path) {
	Handle* h = static_cast<Handle*>(hv);
	kmer_Set_Light& K = *h->ksl;
	Header hd;
	std::memset(&hd, 0, sizeof hd);
	std::memcpy(hd.magic, "BLFLAT01", 8);
	hd.k = h->k; hd.m = h->m; hd.n_log2 = h->n; hd.s_log2 = h->s; hd.b = h->b;
	hd.n_buckets = K.minimizer_number.value();
	hd.n_mphf = K.mphf_number.value();
	hd.number_kmer = K.number_kmer;
	hd.number_super_kmer = K.number_super_kmer;
	hd.total_nuc = K.bucketSeq.size() / 2;
	hd.positions_bits = K.positions.size();
	hd.seq_words = (K.bucketSeq.size() + 63) / 64;
	hd.pos_words = (K.positions.size() + 63) / 64;

	std::vector<MphfRec> recs(hd.n_mphf);
	std::vector<std::pair<uint64_t,uint64_t>> fb_all;
	for(uint64_t i = 0; i < hd.n_mphf; i++) {
		auto& info = K.all_mphf[i];
		MphfRec& r = recs[i];
		std::memset(&r, 0, sizeof r);
		r.id_offset = info.mphf_size;
		r.pos_start = info.start;
		r.nbits = info.bit_to_encode;
		r.present = info.kmer_MPHF ? 1 : 0;
		if(info.kmer_MPHF) {
			auto& M = *info.kmer_MPHF;
			r.nelem = M._nelem;
			for(int l = 0; l < 16; l++) r.dom[l] = M._hash_domains[l];
			r.bits_word_off = hd.bits_words_total;
			r.bits_nwords = M.bitset._nwords;
			r.ranks_off = hd.ranks_total;
			r.nranks = M.bitset._nranks;
			r.fb_off = fb_all.size();
			r.fb_count = M._final_hash.size();
			hd.bits_words_total += r.bits_nwords;
			hd.ranks_total += r.nranks;
			std::vector<std::pair<uint64_t,uint64_t>> fb(M._final_hash.begin(), M._final_hash.end());
			std::sort(fb.begin(), fb.end());
			fb_all.insert(fb_all.end(), fb.begin(), fb.end());
		}
	}
	hd.fallback_total = fb_all.size();

	std::ofstream os(path, std::ios::binary);
	if(!os) return -1;
	auto wr = [&](const void* p, size_t n) { os.write(static_cast<const char*>(p), std::streamsize(n)); };
	wr(&hd, sizeof hd);
	{
		std::vector<uint64_t> st(hd.n_buckets);
		std::vector<uint32_t> nu(hd.n_buckets + (hd.n_buckets & 1));
		for(uint64_t i = 0; i < hd.n_buckets; i++) { st[i] = K.all_buckets[i].start; nu[i] = K.all_buckets[i].nuc_minimizer; }
		wr(st.data(), st.size()*8);
		wr(nu.data(), nu.size()*4);
	}
	wr(recs.data(), recs.size()*sizeof(MphfRec));
	// bits past size() in the last word of a vector<bool> are indeterminate: the blob defines them as zero
	auto wr_bits = [&](std::vector<bool>& v, uint64_t nwords) {
		if(!nwords) return;
		wr(vb_words(v), (nwords-1)*8);
		uint64_t last = vb_words(v)[nwords-1];
		if(v.size() % 64) last &= (uint64_t(1) << (v.size() % 64)) - 1;
		wr(&last, 8);
	};
	wr_bits(K.bucketSeq, hd.seq_words);
	wr_bits(K.positions, hd.pos_words);
	for(uint64_t i = 0; i < hd.n_mphf; i++)
		if(K.all_mphf[i].kmer_MPHF) wr(K.all_mphf[i].kmer_MPHF->bitset._bitArray.get(), recs[i].bits_nwords*8);
	for(uint64_t i = 0; i < hd.n_mphf; i++)
		if(K.all_mphf[i].kmer_MPHF && recs[i].nranks) wr(K.all_mphf[i].kmer_MPHF->bitset._ranks.get(), recs[i].nranks*8);
	for(auto& p : fb_all) wr(&p.first, 8);
	for(auto& p : fb_all) wr(&p.second, 8);
	os.flush();
	return os.good() ? 0 : -2;
}

// ---- import: BLFLAT01 blob -> reference object (created with the blob's own parameters) --------
void* blref_import(const char* path, unsigned cores) {
	std::ifstream is(path, std::ios::binary);
	if(!is) return nullptr;
	Header hd;
	is.read(reinterpret_cast<char*>(&hd), sizeof hd);
	if(!is || std::memcmp(hd.magic, "BLFLAT01", 8) != 0) return nullptr;
	Handle* h = static_cast<Handle*>(blref_create(hd.k, hd.m, hd.n_log2, hd.s_log2, cores, hd.b));
	if(!h) return nullptr;
	kmer_Set_Light& K = *h->ksl;
	if(K.minimizer_number.value() != hd.n_buckets || K.mphf_number.value() != hd.n_mphf) { blref_destroy(h); return nullptr; }
	auto rd = [&](void* p, size_t n) { is.read(static_cast<char*>(p), std::streamsize(n)); };
	{
		std::vector<uint64_t> st(hd.n_buckets);
		std::vector<uint32_t> nu(hd.n_buckets + (hd.n_buckets & 1));
		rd(st.data(), st.size()*8);
		rd(nu.data(), nu.size()*4);
		for(uint64_t i = 0; i < hd.n_buckets; i++) {
			K.all_buckets[i].start = st[i];
			K.all_buckets[i].current_pos = st[i] + nu[i];
			K.all_buckets[i].nuc_minimizer = nu[i];
		}
	}
	std::vector<MphfRec> recs(hd.n_mphf);
	rd(recs.data(), recs.size()*sizeof(MphfRec));
	K.bucketSeq.resize(hd.total_nuc*2);
	if(hd.seq_words) rd(vb_words(K.bucketSeq), hd.seq_words*8);
	K.positions.resize(hd.positions_bits);
	if(hd.pos_words) rd(vb_words(K.positions), hd.pos_words*8);
	K.number_kmer = hd.number_kmer;
	K.number_super_kmer = hd.number_super_kmer;
	K.positions_total_size = hd.positions_bits;
	using MPHF = kmer_Set_Light::MPHF;
	for(uint64_t i = 0; i < hd.n_mphf; i++) {
		auto& info = K.all_mphf[i];
		info.mphf_size = recs[i].id_offset;
		info.start = recs[i].pos_start;
		info.bit_to_encode = recs[i].nbits;
		if(recs[i].present) {
			info.kmer_MPHF = std::unique_ptr<MPHF>(new MPHF());
			MPHF& M = *info.kmer_MPHF;
			M._nelem = recs[i].nelem;
			M._gamma = 2.0;
			for(int l = 0; l < 16; l++) M._hash_domains[l] = recs[i].dom[l];
			M.bitset = boomphf::bitVector(recs[i].bits_nwords*64);
			rd(M.bitset._bitArray.get(), recs[i].bits_nwords*8);
		}
	}
	for(uint64_t i = 0; i < hd.n_mphf; i++) {
		if(!recs[i].present) continue;
		auto& bs = K.all_mphf[i].kmer_MPHF->bitset;
		bs._nranks = recs[i].nranks;
		bs._ranks = std::unique_ptr<uint64_t[]>(new uint64_t[recs[i].nranks ? recs[i].nranks : 1]());
		if(recs[i].nranks) rd(bs._ranks.get(), recs[i].nranks*8);
	}
	std::vector<uint64_t> fk(hd.fallback_total), fv(hd.fallback_total);
	if(hd.fallback_total) { rd(fk.data(), fk.size()*8); rd(fv.data(), fv.size()*8); }
	for(uint64_t i = 0; i < hd.n_mphf; i++) {
		if(!recs[i].present) continue;
		auto& M = *K.all_mphf[i].kmer_MPHF;
		for(uint64_t j = 0; j < recs[i].fb_count; j++) M._final_hash[fk[recs[i].fb_off + j]] = fv[recs[i].fb_off + j];
	}
	if(!is) { blref_destroy(h); return nullptr; }
	return h;
}

int blref_max_threads() { return omp_get_max_threads(); }

} // extern "C"
