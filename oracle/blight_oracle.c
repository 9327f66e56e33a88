/* oracle/blight_oracle.c — TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C CPU restatement of the reference's batched k-mer query path, working on the flat BLFLAT01 image of
 * the reference's own data layout.  It exists so that the CUDA path can be checked on machines where the
 * reference itself is not present (the GPU box): only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it, and only as the checker.  The product never links it.
 *
 * Pinning: the reference ships no tests or golden vectors; this restatement is pinned against
 * outputs of the reference itself (oracle/_ref, built by oracle/build_ref.sh from /root/reference + the two
 * documented one-line fixes), on the lambda sample in five index shapes, on random absent k-mers and on
 * error-bearing simulated reads (tests/test_oracle.py), and against the vectors committed under tests/golden/.
 *
 * Every function names the reference lines it follows.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BLO_LEVELS 16

typedef struct {
	uint64_t id_offset, pos_start, nelem, bits_word_off, bits_nwords, ranks_off, nranks, fb_off, fb_count;
	uint32_t nbits, present;
	uint64_t dom[BLO_LEVELS];
} blo_mphf;

typedef struct {
	char magic[8];
	uint32_t k, m, n_log2, s_log2, b, r0, r1, r2;
	uint64_t n_buckets, n_mphf, number_kmer, number_super_kmer, total_nuc, positions_bits;
	uint64_t seq_words, pos_words, bits_words_total, ranks_total, fallback_total;
} blo_header;

typedef struct {
	blo_header h;
	uint64_t* bucket_start;
	uint32_t* bucket_nuc;
	blo_mphf* mphf;
	uint64_t* seq; /* padded with zero words: the reference reads past the end (blight.cpp:732-739), defined as zeros here */
	uint64_t* pos;
	uint64_t* bits;
	uint64_t* ranks;
	uint64_t* fb_keys;
	uint64_t* fb_vals;
} blo_index;

static void* rd_array(FILE* f, size_t elem, uint64_t n, uint64_t extra) {
	size_t bytes = (size_t)(elem * n);
	void* p = calloc((size_t)(elem * (n + extra)) + 8, 1);
	if (!p) return NULL;
	if (bytes && fread(p, 1, bytes, f) != bytes) { free(p); return NULL; }
	if (bytes % 8) fseek(f, (long)(8 - bytes % 8), SEEK_CUR);
	return p;
}

void blo_free(blo_index* x) {
	if (!x) return;
	free(x->bucket_start); free(x->bucket_nuc); free(x->mphf); free(x->seq); free(x->pos);
	free(x->bits); free(x->ranks); free(x->fb_keys); free(x->fb_vals); free(x);
}

blo_index* blo_load(const char* path) {
	FILE* f = fopen(path, "rb");
	if (!f) return NULL;
	blo_index* x = (blo_index*)calloc(1, sizeof *x);
	if (fread(&x->h, sizeof x->h, 1, f) != 1 || memcmp(x->h.magic, "BLFLAT01", 8) != 0) { fclose(f); free(x); return NULL; }
	const blo_header* h = &x->h;
	uint64_t pad_words = (((uint64_t)1 << h->b) + h->k) / 32 + 2;
	x->bucket_start = (uint64_t*)rd_array(f, 8, h->n_buckets, 0);
	x->bucket_nuc = (uint32_t*)rd_array(f, 4, h->n_buckets, 0);
	x->mphf = (blo_mphf*)rd_array(f, sizeof(blo_mphf), h->n_mphf, 0);
	x->seq = (uint64_t*)rd_array(f, 8, h->seq_words, pad_words);
	x->pos = (uint64_t*)rd_array(f, 8, h->pos_words, 1);
	x->bits = (uint64_t*)rd_array(f, 8, h->bits_words_total, 0);
	x->ranks = (uint64_t*)rd_array(f, 8, h->ranks_total, 0);
	x->fb_keys = (uint64_t*)rd_array(f, 8, h->fallback_total, 0);
	x->fb_vals = (uint64_t*)rd_array(f, 8, h->fallback_total, 0);
	fclose(f);
	if (!x->bucket_start || !x->bucket_nuc || !x->mphf || !x->seq || !x->pos || !x->bits || !x->ranks || !x->fb_keys || !x->fb_vals) {
		blo_free(x);
		return NULL;
	}
	return x;
}

uint32_t blo_k(const blo_index* x) { return x->h.k; }
uint32_t blo_m(const blo_index* x) { return x->h.m; }
uint64_t blo_number_kmer(const blo_index* x) { return x->h.number_kmer; }

/* nuc2int, kmer.h:56-69: (c>>1)&3 for ACGTacgt, anything else is an error (-1 here, std::domain_error there) */
static int nuc2int(unsigned char c) {
	uint8_t d = (uint8_t)(c - 65);
	if (d <= 51 && (0x0008004500080045ull & (1ull << d))) return (c >> 1) & 3;
	return -1;
}

/* rcb(uint64_t, n), kmer.h:218-232 */
static uint64_t rcb64(uint64_t in, unsigned n) {
	uint64_t res = __builtin_bswap64(in ^ 0xaaaaaaaaaaaaaaaaull);
	const uint64_t c1 = 0x0f0f0f0f0f0f0f0full, c2 = 0x3333333333333333ull;
	res = ((res & c1) << 4) | ((res & (c1 << 4)) >> 4);
	res = ((res & c2) << 2) | ((res & (c2 << 2)) >> 2);
	return res >> (64 - 2 * n);
}

/* rcb(uint32_t, n), kmer.h:236-251 */
static uint32_t rcb32(uint32_t in, unsigned n) {
	uint32_t res = __builtin_bswap32(in ^ 0xaaaaaaaau);
	const uint32_t c1 = 0x0f0f0f0fu, c2 = 0x33333333u;
	res = ((res & c1) << 4) | ((res & (c1 << 4)) >> 4);
	res = ((res & c2) << 2) | ((res & (c2 << 2)) >> 2);
	return res >> (32 - 2 * n);
}

/* revhash(uint32_t) -> int32_t, kmer.h:102-108 */
static int32_t revhash(uint32_t x) {
	x = ((x >> 16) ^ x) * 0x2c1b3c6du;
	x = ((x >> 16) ^ x) * 0x297a2d39u;
	x = ((x >> 16) ^ x);
	return (int32_t)x;
}

/* ParityCanonical::canonize, kmer.h:480-482 */
static uint32_t parity_canonize(uint32_t x, unsigned m) {
	return ((__builtin_popcount(x) & 1) ? x : rcb32(x, m)) >> 1;
}

/* minimizer_naive with fix P1 (canonize(mmer, m)), kmer.h:791-810 */
uint32_t blo_minimizer(uint64_t seq, unsigned k, unsigned m) {
	const uint32_t mask = (uint32_t)(((uint64_t)1 << (2 * m)) - 1);
	uint32_t mmer = (uint32_t)seq & mask;
	mmer = parity_canonize(mmer, m);
	uint32_t mini = mmer;
	int32_t hash_mini = revhash(mini);
	for (unsigned i = 1; i <= k - m; i++) {
		seq >>= 2;
		mmer = (uint32_t)seq & mask;
		mmer = parity_canonize(mmer, m);
		int32_t hash = revhash(mmer);
		if (hash_mini > hash) { mini = mmer; hash_mini = hash; }
	}
	return mini;
}

/* SingleHashFunctor::hash_bis, bbhash.h:172-185 */
static uint64_t hash_bis(uint64_t key, uint64_t seed) {
	uint64_t hash = seed;
	hash ^= (hash << 7) ^ key * (hash >> 3) ^ (~((hash << 11) + (key ^ (hash >> 5))));
	hash = (~hash) + (hash << 21);
	hash = hash ^ (hash >> 24);
	hash = (hash + (hash << 3)) + (hash << 8);
	hash = hash ^ (hash >> 14);
	hash = (hash + (hash << 2)) + (hash << 4);
	hash = hash ^ (hash >> 28);
	hash = hash + (hash << 31);
	return hash;
}

/* mphf::lookup, bbhash.h:561-577, with level_finder_tmp (617-639), XorshiftHashFunctors::iter (219-250),
 * fastmod64 (660-662), bitVector::operator[] (404-408) and bitVector::rank (467-480). ~0 = not in the set. */
static uint64_t mphf_lookup(const blo_index* x, const blo_mphf* M, uint64_t key) {
	const uint64_t* bits = x->bits + M->bits_word_off;
	const uint64_t* ranks = x->ranks + M->ranks_off;
	uint64_t s[2] = {0, 0};
	uint64_t bit_begin = 0;
	for (unsigned level = 0; level < BLO_LEVELS; level++) {
		uint64_t hash;
		if (level == 0) hash = s[0] = hash_bis(key, 0xAAAAAAAA55555555ull);
		else if (level == 1) hash = s[1] = hash_bis(key, 0x33333333CCCCCCCCull);
		else {
			uint64_t s1 = s[0];
			const uint64_t s0 = s[1];
			s[0] = s0;
			s1 ^= s1 << 23;
			hash = (s[1] = (s1 ^ s0 ^ (s1 >> 17) ^ (s0 >> 26))) + s0;
		}
		const uint64_t dom = M->dom[level];
		const uint64_t bit = (uint64_t)(((unsigned __int128)hash * (unsigned __int128)dom) >> 64) + bit_begin;
		if ((bits[bit >> 6] >> (bit & 63)) & 1) {
			const uint64_t word_idx = bit / 64, word_offset = bit % 64, block = word_idx / 16;
			uint64_t r = ranks[block];
			for (uint64_t w = block * 16; w < word_idx; ++w) r += (uint64_t)__builtin_popcountll(bits[w]);
			r += (uint64_t)__builtin_popcountll(bits[word_idx] & (((uint64_t)1 << word_offset) - 1));
			return r;
		}
		bit_begin += dom;
	}
	/* _final_hash.find, bbhash.h:567-575; the blob keeps each map sorted by key */
	uint64_t lo = 0, hi = M->fb_count;
	const uint64_t* fk = x->fb_keys + M->fb_off;
	while (lo < hi) {
		uint64_t mid = (lo + hi) / 2;
		if (fk[mid] < key) lo = mid + 1; else hi = mid;
	}
	if (lo < M->fb_count && fk[lo] == key) return x->fb_vals[M->fb_off + lo];
	return ~(uint64_t)0;
}

/* vector<bool> bit i of the packed arrays */
static inline unsigned vb(const uint64_t* w, uint64_t i) { return (unsigned)((w[i >> 6] >> (i & 63)) & 1); }

/* bool_to_int, blight.cpp:473-482 (uint32 result, times 2^b) */
static uint32_t bool_to_int(const blo_index* x, unsigned nbits, uint64_t pos, uint64_t start) {
	uint32_t res = 0, acc = 1;
	for (uint64_t i = 0; i < nbits; ++i, acc <<= 1)
		if (vb(x->pos, i + pos * nbits + start)) res |= acc;
	return res << x->h.b;
}

/* get_kmer, blight.cpp:387-396 */
static uint64_t get_kmer(const blo_index* x, uint64_t mini, uint64_t pos) {
	uint64_t res = 0;
	uint64_t bit = (x->bucket_start[mini] + pos) * 2;
	const uint64_t bitlast = bit + 2 * x->h.k;
	for (; bit < bitlast; bit += 2) { res <<= 2; res |= (uint64_t)(vb(x->seq, bit) * 2 | vb(x->seq, bit + 1)); }
	return res;
}

/* update_kmer / update_kmer_local, blight.cpp:406-417 */
static uint64_t update_kmer(const blo_index* x, uint64_t pos, uint64_t mini, uint64_t input) {
	const uint64_t bit0 = (x->bucket_start[mini] + pos) * 2;
	input <<= 2;
	input |= (uint64_t)(vb(x->seq, bit0) * 2 | vb(x->seq, bit0 + 1));
	return input & (((uint64_t)1 << (2 * x->h.k)) - 1);
}

/* query_get_hash, blight.cpp:716-742 */
int64_t blo_query_get_hash(const blo_index* x, uint64_t canon, uint32_t minimizer) {
	const unsigned k = x->h.k;
	if (x->bucket_nuc[minimizer] == 0) return -1;
	const blo_mphf* M = &x->mphf[minimizer >> (2 * x->h.m - 1 - x->h.n_log2)];
	if (!M->present) return -1; /* "Empty MPHF for non empty bucket" cannot happen in a well-formed index */
	const uint64_t hash = mphf_lookup(x, M, canon);
	if (hash == ~(uint64_t)0) return -1;
	const uint64_t pos = bool_to_int(x, M->nbits, hash, M->pos_start);
	if ((pos + k - 1) < x->bucket_nuc[minimizer]) {
		uint64_t seqR = get_kmer(x, minimizer, pos);
		const uint64_t n_check = (uint64_t)1 << x->h.b;
		for (uint64_t j = pos; j < pos + n_check; ++j) {
			const uint64_t rc = rcb64(seqR, k);
			const uint64_t canonR = seqR <= rc ? seqR : rc;
			if (canon == canonR) return (int64_t)(hash + M->id_offset);
			seqR = update_kmer(x, j + k, minimizer, seqR);
		}
	}
	return -1;
}

/* query_kmer_hash, blight.cpp:545-550 */
int64_t blo_query_kmer_hash(const blo_index* x, uint64_t canon) {
	return blo_query_get_hash(x, canon, blo_minimizer(canon, x->h.k, x->h.m));
}

void blo_query_kmers_hash(const blo_index* x, const uint64_t* canon, uint64_t n, int64_t* out) {
	for (uint64_t i = 0; i < n; i++) out[i] = blo_query_kmer_hash(x, canon[i]);
}

/* query_sequence_hash, blight.cpp:575-591 (str2num kmer.h:90-98, updateK / updateRCK blight.cpp:78-98, min_k 86-91).
 * Returns the number of ids written, 0 for len < k, -1 on an invalid base. canon_out / mini_out are optional. */
int64_t blo_query_sequence_hash(const blo_index* x, const char* q, uint64_t len, int64_t* out, uint64_t* canon_out, uint32_t* mini_out) {
	const unsigned k = x->h.k;
	if (len < k) return 0;
	const uint64_t mask = ((uint64_t)1 << (2 * k)) - 1;
	uint64_t seq = 0;
	for (unsigned i = 0; i < k; i++) {
		int c = nuc2int((unsigned char)q[i]);
		if (c < 0) return -1;
		seq = (seq << 2) | (uint64_t)c;
	}
	uint64_t rc = rcb64(seq, k);
	uint64_t n = 0;
	for (uint64_t i = 0;; ++i) {
		const uint64_t canon = seq <= rc ? seq : rc;
		if (canon_out) canon_out[n] = canon;
		if (mini_out) mini_out[n] = blo_minimizer(canon, k, x->h.m);
		if (out) out[n] = blo_query_kmer_hash(x, canon);
		n++;
		if (i + k >= len) break;
		int c = nuc2int((unsigned char)q[i + k]);
		if (c < 0) return -1;
		seq = ((seq << 2) | (uint64_t)c) & mask;
		rc = (rc >> 2) | ((uint64_t)(c ^ 2) << (2 * k - 2));
	}
	return (int64_t)n;
}

/* Body of file_query's loop over pre-split reads (blight.cpp:780-789): reads shorter than k are skipped.
 * ids_out (optional) receives read r's ids at kmer_offs[r]. ctr[0]=found ctr[1]=not found ctr[2]=queries.
 * Returns 0, or -1 on an invalid base. */
int blo_query_reads_mt(const blo_index* x, const char* bases, const uint64_t* offs, uint64_t n_reads, int64_t* ids_out,
                       const uint64_t* kmer_offs, uint64_t* ctr, int threads);

int blo_query_reads(const blo_index* x, const char* bases, const uint64_t* offs, uint64_t n_reads, int64_t* ids_out,
                    const uint64_t* kmer_offs, uint64_t* ctr) {
	return blo_query_reads_mt(x, bases, offs, n_reads, ids_out, kmer_offs, ctr, 1);
}

/* The same over `threads` host threads (reads are independent, blight.cpp:776-790 does the same with OpenMP): lets the
 * large-configuration parity tests check tens of millions of k-mers in seconds. */
int blo_query_reads_mt(const blo_index* x, const char* bases, const uint64_t* offs, uint64_t n_reads, int64_t* ids_out,
                       const uint64_t* kmer_offs, uint64_t* ctr, int threads) {
	const unsigned k = x->h.k;
	uint64_t maxlen = 0;
	for (uint64_t r = 0; r < n_reads; r++) if (offs[r + 1] - offs[r] > maxlen) maxlen = offs[r + 1] - offs[r];
	uint64_t found = 0, notfound = 0, queries = 0;
	int bad = 0;
	if (threads < 1) threads = 1;
	#pragma omp parallel num_threads(threads) reduction(+ : found, notfound, queries, bad)
	{
		int64_t* tmp = (int64_t*)malloc(sizeof(int64_t) * (size_t)(maxlen + 1));
		#pragma omp for schedule(dynamic, 256)
		for (int64_t r = 0; r < (int64_t)n_reads; r++) {
			const uint64_t len = offs[r + 1] - offs[r];
			if (len < k) continue;
			int64_t n = blo_query_sequence_hash(x, bases + offs[r], len, tmp, NULL, NULL);
			if (n < 0) { bad++; continue; }
			for (int64_t i = 0; i < n; i++) {
				if (tmp[i] >= 0) found++; else notfound++;
				if (ids_out) ids_out[kmer_offs[r] + (uint64_t)i] = tmp[i];
			}
			queries += (uint64_t)n;
		}
		free(tmp);
	}
	ctr[0] = found; ctr[1] = notfound; ctr[2] = queries;
	return bad ? -1 : 0;
}
