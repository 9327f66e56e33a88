// builder.cpp — host construction of the flat index.
//
// Produces, bit for bit, what the reference's construct_index (blight.cpp:108-125, run with cores=1,
// fixes P1+P2 of SURVEY.md F3 applied) leaves in memory, but is organised differently: there are no
// temporary super-bucket files and no validity bitmap; sequences are chopped into super-k-mers in
// parallel work items (block prefix/suffix window minimum instead of a monotone deque), a stable
// counting sort by minimizer gives every super-k-mer its final nucleotide offset, and MPHF groups are
// built and their positions filled independently in parallel.
//
// Semantics that must match the reference:
//   super-k-mers      maximal runs of k-mers with equal minimizer VALUE (kmer.h:640-693), emitted in
//                     file order then position; appended to bucket[minimizer] in that order
//                     (blight.cpp:236-247, 311-324, 335-351)
//   bucket table      start = prefix sum of nucleotides in bucket order (blight.cpp:285-290)
//   per MPHF group    nbits = max(1, ceil(log2(max bucket + 1)) - b), positions slab = nbits*keys + 8 bits,
//                     id offset = exclusive prefix of key counts (blight.cpp:291-303)
//   BBHash, gamma=2   level domains by the double-precision formula (bbhash.h:591-614), a bit survives a
//                     level iff exactly one pending key hashes to it (bbhash.h:668-707), ranks sampled
//                     every 16 words (bbhash.h:447-465), leftovers to the fallback map with consecutive
//                     ranks in key order (bbhash.h:709-728)
//   positions         field[rank(kmer)] = offset_in_bucket >> b for every k-mer but the bucket's first
//                     (blight.cpp:486-519), later writes win
#include "flat_index.hpp"
#include "errors.hpp"
#include "kmer_math.hpp"

#include <omp.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <unordered_map>

namespace blight {

namespace {

void set_err(std::string* err, const std::string& s) { if (err) *err = s; }

struct SuperK {
	const char* p;  // first base in the caller's text
	uint32_t len;   // nucleotides (k-mers = len-k+1)
	uint32_t mini;  // bucket
};

// ---- chopping -------------------------------------------------------------------------------------------

struct WorkItem { uint32_t seq; uint64_t kbeg, kend; };  // k-mers [kbeg,kend) of sequence seq

struct ChopScratch {
	std::vector<uint32_t> key, pre, suf;
};

// Emits the runs of equal window-minimum over k-mers [kbeg,kend) of s. Returns false on an invalid base.
bool chop_item(const char* s, uint64_t kbeg, uint64_t kend, unsigned k, unsigned m, ChopScratch& sc, std::vector<SuperK>& out) {
	const unsigned w = k - m + 1;
	const uint64_t nk = kend - kbeg;
	const uint64_t nm = nk + w - 1;       // m-mers needed
	const char* base = s + kbeg;          // first base of the first k-mer
	sc.key.resize(nm);
	const uint32_t mask = (1u << (2 * m)) - 1u;
	uint32_t fwd = 0, rev = 0;
	for (uint64_t i = 0; i < nm + m - 1; i++) {
		const uint32_t c = nuc_code((unsigned char)base[i]);
		if (c > 3) return false;
		fwd = ((fwd << 2) | c) & mask;
		rev = (rev >> 2) | ((c ^ 2u) << (2 * m - 2));
		if (i + 1 >= m) sc.key[i + 1 - m] = mini_key(((popc32(fwd) & 1) ? fwd : rev) >> 1);
	}
	// window minimum of width w by block prefix / suffix minima
	sc.pre.resize(nm);
	sc.suf.resize(nm);
	for (uint64_t b0 = 0; b0 < nm; b0 += w) {
		const uint64_t b1 = std::min<uint64_t>(b0 + w, nm);
		uint32_t a = 0xFFFFFFFFu;
		for (uint64_t i = b0; i < b1; i++) { a = std::min(a, sc.key[i]); sc.pre[i] = a; }
		a = 0xFFFFFFFFu;
		for (uint64_t i = b1; i-- > b0;) { a = std::min(a, sc.key[i]); sc.suf[i] = a; }
	}
	uint64_t run = 0;
	uint32_t run_key = std::min(sc.suf[0], sc.pre[w - 1]);
	for (uint64_t i = 1; i < nk; i++) {
		const uint32_t cur = std::min(sc.suf[i], sc.pre[i + w - 1]);
		if (cur != run_key) {
			out.push_back(SuperK{base + run, uint32_t(i - run + k - 1), mini_from_key(run_key)});
			run = i;
			run_key = cur;
		}
	}
	out.push_back(SuperK{base + run, uint32_t(nk - run + k - 1), mini_from_key(run_key)});
	return true;
}

// ---- BBHash ---------------------------------------------------------------------------------------------

// Level domains, the reference's double-precision recipe (bbhash.h:591-614) with gamma = 2 (blight.h:60).
__attribute__((optimize("fp-contract=off"))) void level_domains(uint64_t nelem, uint64_t dom[kLevels]) {
	const double gamma = 2.0;
	double proba_collision = 1.0 - pow(((gamma * (double)nelem - 1) / (gamma * (double)nelem)), (double)(nelem - 1));
	size_t hash_domain = (size_t)(ceil(double(nelem) * gamma));
	for (unsigned ii = 0; ii < (unsigned)kLevels; ii++) {
		dom[ii] = (((uint64_t)(hash_domain * pow(proba_collision, (double)ii)) + 63) / 64) * 64;
		if (dom[ii] == 0) dom[ii] = 64;
	}
}

struct MphfBuilt {
	uint64_t dom[kLevels];
	std::vector<uint64_t> bits, ranks;
	std::vector<std::pair<uint64_t, uint64_t>> fb;  // sorted by key
};

inline bool bit_get(const std::vector<uint64_t>& v, uint64_t i) { return (v[i >> 6] >> (i & 63)) & 1; }

struct Pending { uint64_t key, s0, s1, bit; };

void build_mphf(const std::vector<uint64_t>& keys, MphfBuilt& M) {
	const uint64_t n = keys.size();
	level_domains(n, M.dom);
	uint64_t total = 0;
	for (int l = 0; l < kLevels; l++) total += M.dom[l];
	M.bits.assign(total / 64, 0);
	std::vector<uint64_t> coll(M.dom[0] / 64);
	std::vector<Pending> pend, next;
	uint64_t level_off = 0;
	bool finished = false;
	for (int level = 0; level < kLevels && !finished; level++) {
		const uint64_t dom = M.dom[level];
		std::fill(coll.begin(), coll.begin() + dom / 64, 0);
		bool collided = false;
		auto place = [&](uint64_t local) {
			uint64_t& wd = M.bits[(level_off + local) >> 6];
			const uint64_t msk = 1ull << (local & 63);  // level_off is a multiple of 64
			if (wd & msk) { coll[local >> 6] |= msk; collided = true; }
			else wd |= msk;
		};
		if (level == 0) {
			for (uint64_t i = 0; i < n; i++) place(mulhi64(hash_bis(keys[i], kSeed0), dom));
		} else {
			for (Pending& p : pend) {
				uint64_t h;
				if (level == 1) h = p.s1 = hash_bis(p.key, kSeed1);
				else h = xs128_next(p.s0, p.s1);
				p.bit = mulhi64(h, dom);
				place(p.bit);
			}
		}
		if (collided)
			for (uint64_t wi = 0; wi < dom / 64; wi++) M.bits[(level_off >> 6) + wi] &= ~coll[wi];
		if (!collided) {
			level_off += dom;
			finished = true;
			break;
		}
		// keys that did not keep a bit at this level go on to the next one (stable order)
		next.clear();
		if (level == 0) {
			for (uint64_t i = 0; i < n; i++) {
				const uint64_t s0 = hash_bis(keys[i], kSeed0);
				if (!bit_get(M.bits, level_off + mulhi64(s0, dom))) next.push_back(Pending{keys[i], s0, 0, 0});
			}
		} else {
			for (const Pending& p : pend)
				if (!bit_get(M.bits, level_off + p.bit)) next.push_back(p);
		}
		pend.swap(next);
		level_off += dom;
	}
	// ranks: one sample per 16 words over the levels in use (all of them if the fallback is needed)
	const uint64_t upto = finished ? level_off : total;
	const uint64_t max_idx = (upto + 63) / 64;
	M.ranks.assign((max_idx + 15) / 16, 0);
	uint64_t cur = 0;
	for (uint64_t ii = 0, r = 0; ii < max_idx; ii++) {
		if ((ii & 15) == 0) M.ranks[r++] = cur;
		cur += (uint64_t)popc64(M.bits[ii]);
	}
	M.fb.clear();
	if (!finished) {
		std::unordered_map<uint64_t, uint64_t> fm;
		for (const Pending& p : pend) fm[p.key] = cur++;
		M.fb.assign(fm.begin(), fm.end());
		std::sort(M.fb.begin(), M.fb.end());
	}
}

// mphf::lookup (bbhash.h:561-577) on the built arrays; ~0 if the key is in no level and not in the fallback.
uint64_t mphf_rank(const MphfBuilt& M, uint64_t key) {
	uint64_t s0 = 0, s1 = 0, off = 0;
	for (int level = 0; level < kLevels; level++) {
		uint64_t h;
		if (level == 0) h = s0 = hash_bis(key, kSeed0);
		else if (level == 1) h = s1 = hash_bis(key, kSeed1);
		else h = xs128_next(s0, s1);
		const uint64_t bit = off + mulhi64(h, M.dom[level]);
		if (bit_get(M.bits, bit)) {
			const uint64_t wi = bit >> 6;
			uint64_t r = M.ranks[wi >> 4];
			for (uint64_t x = wi & ~15ull; x < wi; x++) r += (uint64_t)popc64(M.bits[x]);
			return r + (uint64_t)popc64(M.bits[wi] & ((1ull << (bit & 63)) - 1));
		}
		off += M.dom[level];
	}
	auto it = std::lower_bound(M.fb.begin(), M.fb.end(), std::make_pair(key, uint64_t(0)));
	if (it != M.fb.end() && it->first == key) return it->second;
	return ~0ull;
}

inline void or_word(uint64_t* w, uint64_t v, bool shared) {
	if (shared) __atomic_fetch_or(w, v, __ATOMIC_RELAXED);
	else *w |= v;
}

}  // namespace

int check_params(const BuildParams& p, std::string* err) {
	// kmer_Set_Light constructor (blight.h:75-92) plus the limits its members imply (Pow2(2k) needs 2k < 64, kmer.h:27)
	if (p.k == 0 || p.k > 31) { set_err(err, "kmer size too large"); return BL_ERR_INVALID_ARG; }
	if ((p.m & 1) == 0) { set_err(err, "minimizer_length must be odd"); return BL_ERR_INVALID_ARG; }
	if (p.m > 15) { set_err(err, "minimizer_length size too large"); return BL_ERR_INVALID_ARG; }
	if (p.m > p.k) { set_err(err, "minimizer_length must not be larger than k"); return BL_ERR_INVALID_ARG; }
	if (p.n_log2 > 2 * p.m - 1) { set_err(err, "log2_mphfs_number must not be larger than 2*minimizer_length - 1"); return BL_ERR_INVALID_ARG; }
	if (p.s_log2 > p.n_log2) { set_err(err, "log2_superbuckets_number must not be larger than log2_mphfs_number"); return BL_ERR_INVALID_ARG; }
	if (p.b > 24) { set_err(err, "bits_to_save too large"); return BL_ERR_INVALID_ARG; }
	return BL_OK;
}

int build_flat_index(const std::vector<SeqView>& seqs, const BuildParams& P, FlatIndex& F, std::string* err) {
	int rc = check_params(P, err);
	if (rc != BL_OK) return rc;
	const unsigned k = P.k, m = P.m, b = P.b;
	const int threads = P.threads ? int(P.threads) : omp_get_max_threads();

	F = FlatIndex();
	FlatHeader& H = F.h;
	std::memcpy(H.magic, "BLFLAT01", 8);
	H.k = k; H.m = m; H.n_log2 = P.n_log2; H.s_log2 = P.s_log2; H.b = b;
	H.n_buckets = 1ull << (2 * m - 1);
	H.n_mphf = 1ull << P.n_log2;
	const unsigned lb = F.lb();

	// 1. work items: slices of at most kItem k-mers, in file order
	const uint64_t kItem = 1ull << 20;
	std::vector<WorkItem> items;
	for (size_t s = 0; s < seqs.size(); s++) {
		if (seqs[s].len < k) continue;  // undefined in the reference (kmer.h:705); skipped here
		const uint64_t nk = seqs[s].len - k + 1;
		for (uint64_t a = 0; a < nk; a += kItem) items.push_back(WorkItem{uint32_t(s), a, std::min(nk, a + kItem)});
	}
	std::vector<std::vector<SuperK>> item_out(items.size());
	bool bad_base = false;
	#pragma omp parallel num_threads(threads)
	{
		ChopScratch sc;
		#pragma omp for schedule(dynamic, 1)
		for (size_t i = 0; i < items.size(); i++) {
			if (!chop_item(seqs[items[i].seq].p, items[i].kbeg, items[i].kend, k, m, sc, item_out[i])) {
				#pragma omp atomic write
				bad_base = true;
			}
		}
	}
	if (bad_base) { set_err(err, "Invalid char in DNA"); return BL_ERR_INVALID_BASE; }

	// 2. stitch items of one sequence (a run may continue across an item boundary), count per bucket
	std::vector<SuperK> sk;
	{
		size_t tot = 0;
		for (auto& v : item_out) tot += v.size();
		sk.reserve(tot);
		for (size_t i = 0; i < items.size(); i++) {
			auto& v = item_out[i];
			size_t j0 = 0;
			if (i > 0 && items[i].seq == items[i - 1].seq && !sk.empty() && !v.empty() && sk.back().mini == v[0].mini) {
				sk.back().len += v[0].len - (k - 1);
				j0 = 1;
			}
			sk.insert(sk.end(), v.begin() + j0, v.end());
			std::vector<SuperK>().swap(v);
		}
	}
	H.number_super_kmer = sk.size();
	std::vector<uint64_t> bnuc(H.n_buckets, 0), bkm(H.n_buckets, 0);
	for (const SuperK& s : sk) { bnuc[s.mini] += s.len; bkm[s.mini] += s.len - (k - 1); }
	F.bucket_start.resize(H.n_buckets);
	F.bucket_nuc.resize(H.n_buckets);
	uint64_t acc = 0;
	for (uint64_t i = 0; i < H.n_buckets; i++) {
		if (bnuc[i] > 0xFFFFFFFFull) { set_err(err, "a minimizer bucket exceeds 2^32 nucleotides (blight.h:33); use a larger m"); return BL_ERR_INVALID_ARG; }
		F.bucket_start[i] = acc;
		F.bucket_nuc[i] = uint32_t(bnuc[i]);
		acc += bnuc[i];
		H.number_kmer += bkm[i];
	}
	H.total_nuc = acc;
	H.seq_words = (acc * 2 + 63) / 64;

	// 3. MPHF group descriptors (blight.cpp:280-306)
	F.mphf.assign(H.n_mphf, MphfRec{});
	{
		uint64_t total_pos = 0, id_base = 0;
		for (uint64_t g = 0; g < H.n_mphf; g++) {
			uint64_t keys = 0; uint32_t maxb = 0;
			for (uint64_t bc = g << lb; bc < ((g + 1) << lb); bc++) { keys += bkm[bc]; maxb = std::max(maxb, F.bucket_nuc[bc]); }
			int nb = (maxb == 0 ? 0 : 32 - __builtin_clz(maxb)) - int(b);  // ceil(log2(maxb+1)) == bit length
			if (nb < 1) nb = 1;
			MphfRec& r = F.mphf[g];
			r.nbits = uint32_t(nb);
			r.pos_start = total_pos;
			r.nelem = keys;
			r.id_offset = id_base;
			r.present = keys ? 1 : 0;
			total_pos += uint64_t(nb) * keys + 8;
			id_base += keys;
		}
		H.positions_bits = total_pos;
		H.pos_words = (total_pos + 63) / 64;
	}

	// 4. stable counting sort by bucket -> every super-k-mer gets its final nucleotide offset
	std::vector<SuperK> sorted(sk.size());
	std::vector<uint64_t> dest(sk.size() + 1);
	{
		std::vector<uint64_t> cursor(H.n_buckets + 1, 0);
		for (const SuperK& s : sk) cursor[s.mini + 1]++;
		for (uint64_t i = 0; i < H.n_buckets; i++) cursor[i + 1] += cursor[i];
		for (const SuperK& s : sk) sorted[cursor[s.mini]++] = s;
		std::vector<SuperK>().swap(sk);
		uint64_t a = 0;
		for (size_t i = 0; i < sorted.size(); i++) { dest[i] = a; a += sorted[i].len; }
		dest[sorted.size()] = a;
	}
	// first super-k-mer of every MPHF group in `sorted`
	std::vector<uint64_t> gfirst(H.n_mphf + 1, 0);
	{
		size_t i = 0;
		for (uint64_t g = 0; g < H.n_mphf; g++) {
			while (i < sorted.size() && (sorted[i].mini >> lb) < g) i++;
			gfirst[g] = i;
		}
		gfirst[H.n_mphf] = sorted.size();
	}

	// 5. pack the bucket sequences (vector<bool> image: nucleotide p -> bit 2p = code>>1, bit 2p+1 = code&1)
	F.seq.assign(H.seq_words, 0);
	#pragma omp parallel for num_threads(threads) schedule(dynamic, 4096)
	for (size_t i = 0; i < sorted.size(); i++) {
		const SuperK& s = sorted[i];
		uint64_t p = dest[i];
		const uint64_t pend = p + s.len;
		const char* c = s.p;
		const uint64_t first_w = p >> 5, last_w = (pend - 1) >> 5;
		while (p < pend) {
			const uint64_t wi = p >> 5;
			const uint64_t stop = std::min(pend, (wi + 1) << 5);
			uint64_t v = 0;
			for (; p < stop; p++, c++) {
				const uint64_t code = nuc_code((unsigned char)*c);
				v |= (((code >> 1) & 1) | ((code & 1) << 1)) << (2 * (p & 31));
			}
			or_word(&F.seq[wi], v, wi == first_w || wi == last_w);
		}
	}

	// 6. per group: BBHash over the canonical k-mers in bucket order, then positions
	F.pos.assign(H.pos_words, 0);
	std::vector<MphfBuilt> built(H.n_mphf);
	const uint64_t kmask = (k == 32) ? ~0ull : ((1ull << (2 * k)) - 1);
	#pragma omp parallel num_threads(threads)
	{
		std::vector<uint64_t> keys;
		#pragma omp for schedule(dynamic, 1)
		for (uint64_t g = 0; g < H.n_mphf; g++) {
			MphfRec& R = F.mphf[g];
			if (!R.present) continue;
			keys.clear();
			keys.reserve(R.nelem);
			for (size_t i = gfirst[g]; i < gfirst[g + 1]; i++) {
				const SuperK& s = sorted[i];
				uint64_t fwd = 0, rev = 0;
				for (uint32_t j = 0; j < s.len; j++) {
					const uint64_t c = nuc_code((unsigned char)s.p[j]);
					fwd = ((fwd << 2) | c) & kmask;
					rev = (rev >> 2) | ((c ^ 2) << (2 * k - 2));
					if (j + 1 >= k) keys.push_back(std::min(fwd, rev));
				}
			}
			MphfBuilt& M = built[g];
			build_mphf(keys, M);
			// positions (blight.cpp:486-519): every k-mer except the one at offset 0 of its bucket
			const uint64_t slab_first = R.pos_start >> 6, slab_last = (R.pos_start + R.nbits * R.nelem + 7) >> 6;
			size_t ki = 0;
			for (size_t i = gfirst[g]; i < gfirst[g + 1]; i++) {
				const SuperK& s = sorted[i];
				const uint64_t off0 = dest[i] - F.bucket_start[s.mini];
				for (uint32_t j = 0; j + k <= s.len; j++, ki++) {
					const uint64_t off = off0 + j;
					if (off == 0) continue;
					const uint64_t rank = mphf_rank(M, keys[ki]);
					const uint64_t val = (off >> b) & ((R.nbits >= 64) ? ~0ull : ((1ull << R.nbits) - 1));
					uint64_t bitpos = R.pos_start + rank * R.nbits;
					unsigned left = R.nbits;
					uint64_t v = val;
					while (left) {
						const uint64_t wi = bitpos >> 6;
						const unsigned sh = unsigned(bitpos & 63);
						const unsigned take = std::min<unsigned>(left, 64 - sh);
						const uint64_t fm = ((take >= 64) ? ~0ull : ((1ull << take) - 1)) << sh;
						const uint64_t fv = (v << sh) & fm;
						if (wi == slab_first || wi == slab_last) {
							__atomic_fetch_and(&F.pos[wi], ~fm, __ATOMIC_RELAXED);
							__atomic_fetch_or(&F.pos[wi], fv, __ATOMIC_RELAXED);
						} else {
							F.pos[wi] = (F.pos[wi] & ~fm) | fv;
						}
						v >>= take; bitpos += take; left -= take;
					}
				}
			}
		}
	}

	// 7. concatenate the MPHF arrays
	for (uint64_t g = 0; g < H.n_mphf; g++) {
		MphfRec& R = F.mphf[g];
		if (!R.present) continue;
		MphfBuilt& M = built[g];
		std::memcpy(R.dom, M.dom, sizeof R.dom);
		R.bits_word_off = F.bits.size();
		R.bits_nwords = M.bits.size();
		R.ranks_off = F.ranks.size();
		R.nranks = M.ranks.size();
		R.fb_off = F.fb_keys.size();
		R.fb_count = M.fb.size();
		F.bits.insert(F.bits.end(), M.bits.begin(), M.bits.end());
		F.ranks.insert(F.ranks.end(), M.ranks.begin(), M.ranks.end());
		for (auto& kv : M.fb) { F.fb_keys.push_back(kv.first); F.fb_vals.push_back(kv.second); }
		MphfBuilt().bits.swap(M.bits);
	}
	H.bits_words_total = F.bits.size();
	H.ranks_total = F.ranks.size();
	H.fallback_total = F.fb_keys.size();
	return flat_validate(F, err);
}

}  // namespace blight
