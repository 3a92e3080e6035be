/*
 * btl_oracle.c -- CPU restatement (plain C) of the btl_bloomfilter k-mer hot path.
 * TEST INFRASTRUCTURE ONLY; see btl_oracle.h.  Parity is pinned by tests/test_oracle.py
 * (known-answer vectors from the reference + the compiled reference itself, oracle/_ref).
 *
 * The arithmetic is restated from its definition (seeds, split-rotate, min, multiply-mix)
 * rather than from the reference's pre-rotated lookup tables: R^n is computed by applying
 * the split-rotate n times (or in closed form), so the msTab31l/msTab33r tables of
 * nthash.hpp:230-347 are validated, not reproduced.
 */
#include "btl_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* nthash.hpp:183-193 */
#define MULTI_SHIFT 27
static const uint64_t MULTI_SEED = 0x90b45d39fb6da1faULL;
static const uint64_t SEED_A = 0x3c8bfbb395c60474ULL;
static const uint64_t SEED_C = 0x3193c18562a02b4cULL;
static const uint64_t SEED_G = 0x20323ed082572324ULL;
static const uint64_t SEED_T = 0x295549f54be24456ULL;
#define CP_OFF 0x07 /* nthash.hpp:180 */

/* seedTab, nthash.hpp:195-228: A/a, C/c, G/g, T/t/U/u and -- as a by-product of the
 * "& cpOff" complement trick -- the raw bytes 1(T) 3(G) 4(A) 5(A) 7(C); all else 0. */
uint64_t ora_seed(unsigned char c)
{
	switch (c) {
	case 'A': case 'a': case 4: case 5: return SEED_A;
	case 'C': case 'c': case 7: return SEED_C;
	case 'G': case 'g': case 3: return SEED_G;
	case 'T': case 't': case 'U': case 'u': case 1: return SEED_T;
	default: return 0;
	}
}

/* rol1 then swap bit 0 with bit 33 (nthash.hpp:350-352,377-380): the low 33 bits and the
 * high 31 bits each rotate left by one. */
uint64_t ora_srol(uint64_t v)
{
	uint64_t r = (v << 1) | (v >> 63);
	uint64_t x = (r ^ (r >> 33)) & 1;
	return r ^ (x | (x << 33));
}

/* ror1 then swap bit 32 with bit 63 (nthash.hpp:360-362,383-386): inverse of ora_srol. */
uint64_t ora_sror(uint64_t v)
{
	uint64_t r = (v >> 1) | (v << 63);
	uint64_t x = ((r >> 32) ^ (r >> 63)) & 1;
	return r ^ ((x << 32) | (x << 63));
}

uint64_t ora_srol_n(uint64_t v, unsigned n)
{
	/* closed form: rot33(lo33, n%33) | rot31(hi31, n%31)  (nthash.hpp:364-374) */
	uint64_t lo = v & 0x1FFFFFFFFULL, hi = v >> 33;
	unsigned a = n % 33, b = n % 31;
	if (a) lo = ((lo << a) | (lo >> (33 - a))) & 0x1FFFFFFFFULL;
	if (b) hi = ((hi << b) | (hi >> (31 - b))) & 0x7FFFFFFFULL;
	return lo | (hi << 33);
}

uint64_t ora_multi_mult(unsigned i, unsigned k)
{
	return (uint64_t)i ^ ((uint64_t)k * MULTI_SEED); /* nthash.hpp:686: i ^ k * multiSeed */
}

static void multi_hash(uint64_t b, unsigned k, unsigned h, uint64_t *hv)
{
	hv[0] = b;
	for (unsigned i = 1; i < h; i++) { /* nthash.hpp:684-690 */
		uint64_t t = b * ora_multi_mult(i, k);
		t ^= t >> MULTI_SHIFT;
		hv[i] = t;
	}
}

static int base_fr(const char *s, unsigned k, uint64_t *fh, uint64_t *rh, unsigned *locN)
{
	uint64_t f = 0, r = 0;
	*locN = 0;
	for (int i = (int)k - 1; i >= 0; i--) { /* nthash.hpp:671-683 */
		if (ora_seed((unsigned char)s[i]) == 0) {
			*locN = (unsigned)i;
			*fh = f;
			*rh = r;
			return 0;
		}
		f = ora_srol(f) ^ ora_seed((unsigned char)s[k - 1 - i]);
		r = ora_srol(r) ^ ora_seed((unsigned char)s[i] & CP_OFF);
	}
	*fh = f;
	*rh = r;
	return 1;
}

int ora_ntmc64_base(const char *s, unsigned k, unsigned h, uint64_t *fh, uint64_t *rh,
                    unsigned *locN, uint64_t *hv)
{
	if (!base_fr(s, k, fh, rh, locN))
		return 0;
	multi_hash(*rh < *fh ? *rh : *fh, k, h, hv);
	return 1;
}

static void roll_fr(unsigned char out, unsigned char in, unsigned k, uint64_t *fh, uint64_t *rh)
{
	/* NTF64 nthash.hpp:442-448 */
	*fh = ora_srol(*fh) ^ ora_seed(in) ^ ora_srol_n(ora_seed(out), k);
	/* NTR64 nthash.hpp:451-457 */
	*rh = ora_sror(*rh ^ ora_srol_n(ora_seed(in & CP_OFF), k) ^ ora_seed(out & CP_OFF));
}

void ora_ntmc64_roll(unsigned char out, unsigned char in, unsigned k, unsigned h, uint64_t *fh,
                     uint64_t *rh, uint64_t *hv)
{
	roll_fr(out, in, k, fh, rh);
	multi_hash(*rh < *fh ? *rh : *fh, k, h, hv);
}

/* ---------------- ntHashIterator ---------------- */
static void nt_init(ora_nt_iter *it)
{
	if (it->k > it->len) { /* ntHashIterator.hpp:61-64 */
		it->pos = ORA_END;
		return;
	}
	unsigned locN = 0;
	size_t last = it->len - it->k + 1;
	while (it->pos < last &&
	       !ora_ntmc64_base(it->seq + it->pos, it->k, it->h, &it->fh, &it->rh, &locN, it->hv))
		it->pos += locN + 1;
	if (it->pos >= last)
		it->pos = ORA_END;
}

void ora_nt_iter_init(ora_nt_iter *it, const char *seq, size_t len, unsigned h, unsigned k)
{
	it->seq = seq;
	it->len = len;
	it->h = h;
	it->k = k;
	it->pos = 0;
	it->fh = it->rh = 0;
	nt_init(it);
}

void ora_nt_iter_next(ora_nt_iter *it)
{
	++it->pos; /* ntHashIterator.hpp:75-85 */
	if (it->pos >= it->len - it->k + 1) {
		it->pos = ORA_END;
		return;
	}
	if (ora_seed((unsigned char)it->seq[it->pos + it->k - 1]) == 0) {
		it->pos += it->k;
		nt_init(it);
	} else {
		ora_ntmc64_roll((unsigned char)it->seq[it->pos - 1],
		                (unsigned char)it->seq[it->pos - 1 + it->k], it->k, it->h, &it->fh,
		                &it->rh, it->hv);
	}
}

/* ---------------- stHashIterator ---------------- */
int ora_seedset_parse(ora_seedset *ss, const char *const *seeds, unsigned n_seeds, unsigned h2,
                      unsigned k)
{
	memset(ss, 0, sizeof *ss);
	if (n_seeds == 0 || n_seeds > ORA_MAX_SEEDS || h2 == 0 || n_seeds * h2 > ORA_MAX_HASH)
		return -1;
	ss->n_seeds = n_seeds;
	ss->h2 = h2;
	ss->k = k;
	for (unsigned j = 0; j < n_seeds; j++) {
		size_t L = strlen(seeds[j]);
		if (L != k) { /* positions index kmerSeq[0..k): longer seeds would read past the k-mer */
			ora_seedset_free(ss);
			return -1;
		}
		ss->dc[j] = (unsigned *)malloc(sizeof(unsigned) * (L ? L : 1));
		for (size_t p = 0; p < L; p++)
			if (seeds[j][p] != '1') /* stHashIterator.hpp:28 */
				ss->dc[j][ss->n_dc[j]++] = (unsigned)p;
	}
	return 0;
}

void ora_seedset_free(ora_seedset *ss)
{
	for (unsigned j = 0; j < ORA_MAX_SEEDS; j++) {
		free(ss->dc[j]);
		ss->dc[j] = NULL;
	}
}

/* per-seed masking + extra hashes, nthash.hpp:839-853 (base) == :864-877 (roll) */
static void st_finish(const char *kmer, const ora_seedset *ss, uint64_t fh, uint64_t rh,
                      uint64_t *hv, uint8_t *stn)
{
	unsigned k = ss->k, m2 = ss->h2;
	for (unsigned j = 0; j < ss->n_seeds; j++) {
		uint64_t fs = fh, rs = rh;
		for (unsigned t = 0; t < ss->n_dc[j]; t++) {
			unsigned p = ss->dc[j][t];
			unsigned char c = (unsigned char)kmer[p];
			fs ^= ora_srol_n(ora_seed(c), k - 1 - p);
			rs ^= ora_srol_n(ora_seed(c & CP_OFF), p);
		}
		uint8_t s = rs < fs;
		uint64_t b = s ? rs : fs;
		hv[j * m2] = b;
		stn[j * m2] = s;
		for (unsigned j2 = 1; j2 < m2; j2++) {
			uint64_t t = b * ora_multi_mult(j2, k);
			t ^= t >> MULTI_SHIFT;
			hv[j * m2 + j2] = t;
			stn[j * m2 + j2] = s;
		}
	}
}

static void st_init(ora_st_iter *it)
{
	unsigned k = it->ss->k;
	if (k > it->len) {
		it->pos = ORA_END;
		return;
	}
	unsigned locN = 0;
	size_t last = it->len - k + 1;
	while (it->pos < last) {
		if (base_fr(it->seq + it->pos, k, &it->fh, &it->rh, &locN)) {
			st_finish(it->seq + it->pos, it->ss, it->fh, it->rh, it->hv, it->stn);
			break;
		}
		it->pos += locN + 1;
	}
	if (it->pos >= last)
		it->pos = ORA_END;
}

void ora_st_iter_init(ora_st_iter *it, const char *seq, size_t len, const ora_seedset *ss)
{
	it->seq = seq;
	it->len = len;
	it->ss = ss;
	it->pos = 0;
	it->fh = it->rh = 0;
	st_init(it);
}

void ora_st_iter_next(ora_st_iter *it)
{
	unsigned k = it->ss->k;
	++it->pos;
	if (it->pos >= it->len - k + 1) {
		it->pos = ORA_END;
		return;
	}
	if (ora_seed((unsigned char)it->seq[it->pos + k - 1]) == 0) {
		it->pos += k;
		st_init(it);
	} else {
		roll_fr((unsigned char)it->seq[it->pos - 1], (unsigned char)it->seq[it->pos - 1 + k], k,
		        &it->fh, &it->rh);
		st_finish(it->seq + it->pos, it->ss, it->fh, it->rh, it->hv, it->stn);
	}
}

/* ---------------- BloomFilter ---------------- */
void ora_bf_insert(uint8_t *filter, uint64_t m, unsigned h, const uint64_t *hv)
{
	for (unsigned i = 0; i < h; i++) {
		uint64_t n = hv[i] % m;
		__sync_or_and_fetch(&filter[n / 8], (uint8_t)(1u << (n % 8)));
	}
}

int ora_bf_contains(const uint8_t *filter, uint64_t m, unsigned h, const uint64_t *hv)
{
	for (unsigned i = 0; i < h; i++) {
		uint64_t n = hv[i] % m;
		if (!(filter[n / 8] & (1u << (n % 8))))
			return 0;
	}
	return 1;
}

int ora_bf_insert_and_check(uint8_t *filter, uint64_t m, unsigned h, const uint64_t *hv)
{
	int found = 1;
	for (unsigned i = 0; i < h; i++) {
		uint64_t n = hv[i] % m;
		uint8_t old = __sync_fetch_and_or(&filter[n / 8], (uint8_t)(1u << (n % 8)));
		found &= (old >> (n % 8)) & 1;
	}
	return found;
}

uint64_t ora_bf_popcount(const uint8_t *filter, uint64_t m)
{
	uint64_t pop = 0;
	for (uint64_t i = 0; i < (m + 7) / 8; i++)
		pop += (uint64_t)__builtin_popcount(filter[i]);
	return pop;
}

/* ---------------- CountingBloomFilter<uint8_t> ---------------- */
uint8_t ora_cbf_mincount(const uint8_t *cnt, uint64_t m, unsigned h, const uint64_t *hv)
{
	uint8_t mn = cnt[hv[0] % m];
	for (unsigned i = 1; i < h; i++) {
		uint8_t v = cnt[hv[i] % m];
		if (v < mn)
			mn = v;
	}
	return mn;
}

/* incrementMin, single-threaded: every counter equal to the minimum goes to min+1 unless the
 * minimum is already 255.  A slot addressed twice by the same k-mer is bumped once (the second
 * CAS sees min+1 and fails), CountingBloomFilter.hpp:134-162. */
void ora_cbf_insert(uint8_t *cnt, uint64_t m, unsigned h, const uint64_t *hv)
{
	uint8_t mn = ora_cbf_mincount(cnt, m, h, hv);
	uint8_t nv = (uint8_t)(mn + 1);
	if (mn > nv)
		return;
	for (unsigned i = 0; i < h; i++) {
		uint64_t p = hv[i] % m;
		if (cnt[p] == mn)
			cnt[p] = nv;
	}
}

void ora_cbf_increment_all(uint8_t *cnt, uint64_t m, unsigned h, const uint64_t *hv)
{
	for (unsigned i = 0; i < h; i++) {
		uint64_t p = hv[i] % m;
		if (cnt[p] != 255)
			cnt[p]++;
	}
}

uint64_t ora_cbf_popcount(const uint8_t *cnt, uint64_t m)
{
	uint64_t c = 0;
	for (uint64_t i = 0; i < m; i++)
		c += cnt[i] != 0;
	return c;
}

uint64_t ora_cbf_filtered_popcount(const uint8_t *cnt, uint64_t m, unsigned thr)
{
	uint64_t c = 0;
	for (uint64_t i = 0; i < m; i++)
		c += cnt[i] >= thr;
	return c;
}

/* ---------------- headers ---------------- */
/* cpptoml writes a double with showpoint + 17 significant digits and then squeezes a leading
 * exponent zero ("e0"->"e", "e-0"->"e-"), cpptoml.h:3477-3494. */
static void toml_double(char *out, size_t cap, double v)
{
	char tmp[64];
	snprintf(tmp, sizeof tmp, "%#.17g", v);
	char *p = strstr(tmp, "e0");
	if (p)
		memmove(p + 1, p + 2, strlen(p + 2) + 1);
	p = strstr(tmp, "e-0");
	if (p)
		memmove(p + 2, p + 3, strlen(p + 3) + 1);
	snprintf(out, cap, "%s", tmp);
}

/* Key order is the iteration order of the std::unordered_map cpptoml keeps the table in
 * (cpptoml.h:48-51) for the insertion sequence of BloomFilter.hpp:273-279 under libstdc++;
 * probed from the reference and re-checked by tests against oracle/_ref. */
int ora_bf_header(char *buf, size_t cap, uint64_t size_bits, uint64_t size_bytes, unsigned h,
                  unsigned k, double dFPR, uint64_t nEntry, uint64_t tEntry)
{
	char d[64];
	toml_double(d, sizeof d, dFPR);
	return snprintf(buf, cap,
	                "[BTLBloomFilter_v1]\n\tnEntry = %lld\n\tdFPR = %s\n\tEntry = %lld\n"
	                "\tBloomFilterSizeInBytes = %lld\n\tBloomFilterSize = %lld\n\tHashNum = %lld\n"
	                "\tKmerSize = %lld\n[HeaderEnd]\n",
	                (long long)nEntry, d, (long long)tEntry, (long long)size_bytes,
	                (long long)size_bits, (long long)h, (long long)k);
}

int ora_cbf_header(char *buf, size_t cap, uint64_t size, uint64_t size_bytes, unsigned h,
                   unsigned k, unsigned bits_per_counter)
{
	return snprintf(buf, cap,
	                "[BTLCountingBloomFilter_v1]\n\tBloomFilterSize = %lld\n\tHashNum = %lld\n"
	                "\tKmerSize = %lld\n\tBloomFilterSizeInBytes = %lld\n\tBitsPerCounter = %lld\n"
	                "[HeaderEnd]\n",
	                (long long)size, (long long)h, (long long)k, (long long)size_bytes,
	                (long long)bits_per_counter);
}

/* ---------------- flat batch drivers ---------------- */
static inline void set_bit(uint8_t *bits, uint64_t p)
{
	bits[p >> 3] |= (uint8_t)(1u << (p & 7));
}

uint64_t ora_hash_seqs(unsigned h, unsigned k, const char *bases, const uint64_t *off,
                       uint64_t n_seqs, uint64_t *hashes, uint8_t *valid_bits)
{
	uint64_t n = 0;
	ora_nt_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_nt_iter_init(&it, bases + off[s], off[s + 1] - off[s], h, k);
		for (; it.pos != ORA_END; ora_nt_iter_next(&it)) {
			uint64_t p = off[s] + it.pos;
			if (hashes)
				memcpy(hashes + p * h, it.hv, sizeof(uint64_t) * h);
			if (valid_bits)
				set_bit(valid_bits, p);
			n++;
		}
	}
	return n;
}

uint64_t ora_bf_insert_seqs(uint8_t *filter, uint64_t m, unsigned h, unsigned k, const char *bases,
                            const uint64_t *off, uint64_t n_seqs)
{
	uint64_t n = 0;
	ora_nt_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_nt_iter_init(&it, bases + off[s], off[s + 1] - off[s], h, k);
		for (; it.pos != ORA_END; ora_nt_iter_next(&it)) {
			ora_bf_insert(filter, m, h, it.hv);
			n++;
		}
	}
	return n;
}

uint64_t ora_bf_contains_seqs(const uint8_t *filter, uint64_t m, unsigned h, unsigned k,
                              const char *bases, const uint64_t *off, uint64_t n_seqs,
                              uint8_t *hit_bits, uint8_t *valid_bits, uint64_t *n_hits)
{
	uint64_t n = 0, hits = 0;
	ora_nt_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_nt_iter_init(&it, bases + off[s], off[s + 1] - off[s], h, k);
		for (; it.pos != ORA_END; ora_nt_iter_next(&it)) {
			uint64_t p = off[s] + it.pos;
			if (valid_bits)
				set_bit(valid_bits, p);
			if (ora_bf_contains(filter, m, h, it.hv)) {
				if (hit_bits)
					set_bit(hit_bits, p);
				hits++;
			}
			n++;
		}
	}
	if (n_hits)
		*n_hits = hits;
	return n;
}

uint64_t ora_bf_insert_and_check_seqs(uint8_t *filter, uint64_t m, unsigned h, unsigned k,
                                      const char *bases, const uint64_t *off, uint64_t n_seqs,
                                      uint8_t *found_bits, uint8_t *valid_bits)
{
	uint64_t n = 0;
	ora_nt_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_nt_iter_init(&it, bases + off[s], off[s + 1] - off[s], h, k);
		for (; it.pos != ORA_END; ora_nt_iter_next(&it)) {
			uint64_t p = off[s] + it.pos;
			if (valid_bits)
				set_bit(valid_bits, p);
			if (ora_bf_insert_and_check(filter, m, h, it.hv) && found_bits)
				set_bit(found_bits, p);
			n++;
		}
	}
	return n;
}

/* MIBFConstructSupport::insertBVColli (MIBFConstructSupport.hpp:55-74) over a batch: 64-bit words of an
 * sdsl::bit_vector of m bits; *colli = k-mers all of whose h bits were already set; returns the k-mers.
 * (insertBV, :76-87, is the same without the count.) */
uint64_t ora_mibf_insert_bv_seqs(uint64_t *words, uint64_t m, unsigned h, unsigned k, const char *bases,
                                 const uint64_t *off, uint64_t n_seqs, uint64_t *colli)
{
	uint64_t n = 0, count = 0;
	ora_nt_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_nt_iter_init(&it, bases + off[s], off[s + 1] - off[s], h, k);
		for (; it.pos != ORA_END; ora_nt_iter_next(&it)) {
			unsigned colliCount = 0;
			for (unsigned i = 0; i < h; i++) {
				uint64_t pos = it.hv[i] % m;
				uint64_t *dataIndex = words + (pos >> 6);
				uint64_t bitMaskValue = (uint64_t)1 << (pos & 0x3F);
				colliCount += (unsigned)(__sync_fetch_and_or(dataIndex, bitMaskValue) >> (pos & 0x3F) & 1);
			}
			if (colliCount == h)
				count++;
			n++;
		}
	}
	if (colli)
		*colli = count;
	return n;
}

uint64_t ora_cbf_insert_seqs(uint8_t *cnt, uint64_t m, unsigned h, unsigned k, const char *bases,
                             const uint64_t *off, uint64_t n_seqs)
{
	uint64_t n = 0;
	ora_nt_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_nt_iter_init(&it, bases + off[s], off[s + 1] - off[s], h, k);
		for (; it.pos != ORA_END; ora_nt_iter_next(&it)) {
			ora_cbf_insert(cnt, m, h, it.hv);
			n++;
		}
	}
	return n;
}

uint64_t ora_cbf_increment_all_seqs(uint8_t *cnt, uint64_t m, unsigned h, unsigned k,
                                    const char *bases, const uint64_t *off, uint64_t n_seqs)
{
	uint64_t n = 0;
	ora_nt_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_nt_iter_init(&it, bases + off[s], off[s + 1] - off[s], h, k);
		for (; it.pos != ORA_END; ora_nt_iter_next(&it)) {
			ora_cbf_increment_all(cnt, m, h, it.hv);
			n++;
		}
	}
	return n;
}

uint64_t ora_cbf_mincount_seqs(const uint8_t *cnt, uint64_t m, unsigned h, unsigned k,
                               const char *bases, const uint64_t *off, uint64_t n_seqs,
                               uint8_t *counts, uint8_t *valid_bits)
{
	uint64_t n = 0;
	ora_nt_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_nt_iter_init(&it, bases + off[s], off[s + 1] - off[s], h, k);
		for (; it.pos != ORA_END; ora_nt_iter_next(&it)) {
			uint64_t p = off[s] + it.pos;
			if (valid_bits)
				set_bit(valid_bits, p);
			if (counts)
				counts[p] = ora_cbf_mincount(cnt, m, h, it.hv);
			n++;
		}
	}
	return n;
}

uint64_t ora_cbf_contains_seqs(const uint8_t *cnt, uint64_t m, unsigned h, unsigned k,
                               unsigned threshold, const char *bases, const uint64_t *off,
                               uint64_t n_seqs, uint8_t *hit_bits, uint8_t *valid_bits,
                               uint64_t *n_hits)
{
	uint64_t n = 0, hits = 0;
	ora_nt_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_nt_iter_init(&it, bases + off[s], off[s + 1] - off[s], h, k);
		for (; it.pos != ORA_END; ora_nt_iter_next(&it)) {
			uint64_t p = off[s] + it.pos;
			if (valid_bits)
				set_bit(valid_bits, p);
			/* CountingBloomFilter.hpp:190-196 */
			if (ora_cbf_mincount(cnt, m, h, it.hv) >= threshold) {
				if (hit_bits)
					set_bit(hit_bits, p);
				hits++;
			}
			n++;
		}
	}
	if (n_hits)
		*n_hits = hits;
	return n;
}

/* ---- spaced seeds ---- */
uint64_t ora_st_hash_seqs(const ora_seedset *ss, const char *bases, const uint64_t *off,
                          uint64_t n_seqs, uint64_t *hashes, uint8_t *strands, uint8_t *valid_bits)
{
	uint64_t n = 0;
	unsigned H = ss->n_seeds * ss->h2;
	ora_st_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_st_iter_init(&it, bases + off[s], off[s + 1] - off[s], ss);
		for (; it.pos != ORA_END; ora_st_iter_next(&it)) {
			uint64_t p = off[s] + it.pos;
			if (hashes)
				memcpy(hashes + p * H, it.hv, sizeof(uint64_t) * H);
			if (strands)
				memcpy(strands + p * H, it.stn, H);
			if (valid_bits)
				set_bit(valid_bits, p);
			n++;
		}
	}
	return n;
}

uint64_t ora_st_bf_insert_seqs(uint8_t *filter, uint64_t m, const ora_seedset *ss,
                               const char *bases, const uint64_t *off, uint64_t n_seqs)
{
	uint64_t n = 0;
	unsigned H = ss->n_seeds * ss->h2;
	ora_st_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_st_iter_init(&it, bases + off[s], off[s + 1] - off[s], ss);
		for (; it.pos != ORA_END; ora_st_iter_next(&it)) {
			ora_bf_insert(filter, m, H, it.hv);
			n++;
		}
	}
	return n;
}

uint64_t ora_st_bf_contains_seqs(const uint8_t *filter, uint64_t m, const ora_seedset *ss,
                                 const char *bases, const uint64_t *off, uint64_t n_seqs,
                                 uint8_t *hit_bits, uint8_t *valid_bits, uint64_t *n_hits)
{
	uint64_t n = 0, hits = 0;
	unsigned H = ss->n_seeds * ss->h2;
	ora_st_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_st_iter_init(&it, bases + off[s], off[s + 1] - off[s], ss);
		for (; it.pos != ORA_END; ora_st_iter_next(&it)) {
			uint64_t p = off[s] + it.pos;
			if (valid_bits)
				set_bit(valid_bits, p);
			if (ora_bf_contains(filter, m, H, it.hv)) {
				if (hit_bits)
					set_bit(hit_bits, p);
				hits++;
			}
			n++;
		}
	}
	if (n_hits)
		*n_hits = hits;
	return n;
}

uint64_t ora_st_cbf_insert_seqs(uint8_t *cnt, uint64_t m, const ora_seedset *ss, const char *bases,
                                const uint64_t *off, uint64_t n_seqs)
{
	uint64_t n = 0;
	unsigned H = ss->n_seeds * ss->h2;
	ora_st_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_st_iter_init(&it, bases + off[s], off[s + 1] - off[s], ss);
		for (; it.pos != ORA_END; ora_st_iter_next(&it)) {
			ora_cbf_insert(cnt, m, H, it.hv);
			n++;
		}
	}
	return n;
}

uint64_t ora_st_cbf_mincount_seqs(const uint8_t *cnt, uint64_t m, const ora_seedset *ss,
                                  const char *bases, const uint64_t *off, uint64_t n_seqs,
                                  uint8_t *counts, uint8_t *valid_bits)
{
	uint64_t n = 0;
	unsigned H = ss->n_seeds * ss->h2;
	ora_st_iter it;
	for (uint64_t s = 0; s < n_seqs; s++) {
		ora_st_iter_init(&it, bases + off[s], off[s + 1] - off[s], ss);
		for (; it.pos != ORA_END; ora_st_iter_next(&it)) {
			uint64_t p = off[s] + it.pos;
			if (valid_bits)
				set_bit(valid_bits, p);
			if (counts)
				counts[p] = ora_cbf_mincount(cnt, m, H, it.hv);
			n++;
		}
	}
	return n;
}

/* ---------------- synthetic inputs ---------------- */
uint64_t ora_splitmix64(uint64_t x)
{
	uint64_t z = x + 0x9e3779b97f4a7c15ULL;
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
	return z ^ (z >> 31);
}

static inline char synth_base(uint64_t i, uint64_t seed)
{
	return "ACGT"[(ora_splitmix64(seed ^ (i >> 5)) >> (2 * (i & 31))) & 3];
}

void ora_synth_genome(char *out, uint64_t start, uint64_t n, uint64_t seed)
{
	for (uint64_t i = 0; i < n; i++)
		out[i] = synth_base(start + i, seed);
}

void ora_synth_reads(char *out, uint64_t first_read, uint64_t n_reads, unsigned read_len,
                     uint64_t g_start, uint64_t g_len, uint64_t genome_seed, uint64_t read_seed)
{
	for (uint64_t r = 0; r < n_reads; r++) {
		uint64_t st = g_start + ora_splitmix64(read_seed + first_read + r) % (g_len - read_len);
		for (unsigned j = 0; j < read_len; j++)
			out[r * read_len + j] = synth_base(st + j, genome_seed);
	}
}

/* ---------------- OpenMP timing leg (pattern of Tests/AdHoc/ParallelFilter.cpp:104-122) ---------------- */
int ora_max_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

double ora_bench_bf(uint8_t *filter, uint64_t m, unsigned h, unsigned k, const char *bases,
                    const uint64_t *off, uint64_t n_seqs, int do_insert, int threads,
                    uint64_t *n_kmers, uint64_t *n_hits)
{
	uint64_t n = 0, hits = 0;
	struct timespec t0, t1;
	clock_gettime(CLOCK_MONOTONIC, &t0);
#ifdef _OPENMP
	if (threads > 0)
		omp_set_num_threads(threads);
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : n, hits)
#endif
	for (int64_t s = 0; s < (int64_t)n_seqs; s++) {
		ora_nt_iter it;
		ora_nt_iter_init(&it, bases + off[s], off[s + 1] - off[s], h, k);
		for (; it.pos != ORA_END; ora_nt_iter_next(&it)) {
			if (do_insert)
				ora_bf_insert(filter, m, h, it.hv);
			else
				hits += (uint64_t)ora_bf_contains(filter, m, h, it.hv);
			n++;
		}
	}
	clock_gettime(CLOCK_MONOTONIC, &t1);
	(void)threads;
	if (n_kmers)
		*n_kmers = n;
	if (n_hits)
		*n_hits = hits;
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}
