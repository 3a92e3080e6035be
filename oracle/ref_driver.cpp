/*
 * ref_driver.cpp -- thin extern "C" driver around the UNMODIFIED reference headers.
 *
 * TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile with -I$(REF) (the read-only upstream
 * tree, /root/reference) into oracle/_ref/libbtlref.so; no reference source is copied into this
 * repository.  It exists to (1) pin the plain-C restatement in btl_oracle.c against the real
 * reference, (2) generate tests/golden/ (oracle/make_golden.py), (3) serve as the CPU baseline
 * ("kind": "reference") of bench.py.  The product path never loads it.
 *
 * Every loop here is the reference's own usage pattern: README.md:30-57 (ntHashIterator +
 * insert/contains), README.md:86-113 (counting filter), BloomFilterUtil.h:10-17 (insertSeq),
 * Tests/AdHoc/ParallelFilter.cpp:104-122 (OpenMP over reads).
 */
#include "KmerBloomFilter.hpp"
#include "CountingBloomFilter.hpp"
#include "vendor/ntHashIterator.hpp"
#include "vendor/stHashIterator.hpp"

#include <chrono>
#include <cstdint>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {
struct RefBF : public BloomFilter
{
	RefBF(size_t bits, unsigned h, unsigned k)
	  : BloomFilter(bits, h, k)
	{}
	explicit RefBF(const std::string& path)
	  : BloomFilter(path)
	{}
	uint8_t* data() { return m_filter; }
	void setdFPR(double v) { m_dFPR = v; }
};
typedef CountingBloomFilter<uint8_t> RefCBF;

inline void
set_bit(uint8_t* bits, uint64_t p)
{
	bits[p >> 3] |= (uint8_t)(1u << (p & 7));
}

std::vector<std::vector<unsigned>>
parse(const char* const* seeds, unsigned n)
{
	std::vector<std::string> s;
	for (unsigned i = 0; i < n; i++)
		s.push_back(seeds[i]);
	return stHashIterator::parseSeed(s);
}
} // namespace

extern "C" {

/* ---------- raw iterator output ---------- */
uint64_t
ref_hash_seqs(unsigned h, unsigned k, const char* bases, const uint64_t* off, uint64_t n_seqs,
              uint64_t* hashes, uint8_t* valid_bits)
{
	uint64_t n = 0;
	for (uint64_t s = 0; s < n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		ntHashIterator itr(seq, h, k);
		while (itr != itr.end()) {
			uint64_t p = off[s] + itr.pos();
			if (hashes)
				for (unsigned i = 0; i < h; i++)
					hashes[p * h + i] = (*itr)[i];
			if (valid_bits)
				set_bit(valid_bits, p);
			++n;
			++itr;
		}
	}
	return n;
}

uint64_t
ref_st_hash_seqs(const char* const* seeds, unsigned n_seeds, unsigned h2, unsigned k,
                 const char* bases, const uint64_t* off, uint64_t n_seqs, uint64_t* hashes,
                 uint8_t* strands, uint8_t* valid_bits)
{
	auto ss = parse(seeds, n_seeds);
	unsigned H = n_seeds * h2;
	uint64_t n = 0;
	for (uint64_t s = 0; s < n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		stHashIterator itr(seq, ss, n_seeds, h2, k);
		while (itr != itr.end()) {
			uint64_t p = off[s] + itr.pos();
			for (unsigned i = 0; i < H; i++) {
				if (hashes)
					hashes[p * H + i] = (*itr)[i];
				if (strands)
					strands[p * H + i] = itr.strandArray()[i];
			}
			if (valid_bits)
				set_bit(valid_bits, p);
			++n;
			++itr;
		}
	}
	return n;
}

/* ---------- BloomFilter ---------- */
void*
ref_bf_new(uint64_t bits, unsigned h, unsigned k)
{
	return new RefBF(bits, h, k);
}
void*
ref_bf_load(const char* path)
{
	return new RefBF(std::string(path));
}
void
ref_bf_free(void* f)
{
	delete (RefBF*)f;
}
uint8_t*
ref_bf_data(void* f)
{
	return ((RefBF*)f)->data();
}
uint64_t
ref_bf_size_bits(void* f)
{
	return ((RefBF*)f)->getFilterSize();
}
uint64_t
ref_bf_size_bytes(void* f)
{
	return ((RefBF*)f)->sizeInBytes();
}
unsigned
ref_bf_hash_num(void* f)
{
	return ((RefBF*)f)->getHashNum();
}
unsigned
ref_bf_kmer_size(void* f)
{
	return ((RefBF*)f)->getKmerSize();
}
uint64_t
ref_bf_pop(void* f)
{
	return ((RefBF*)f)->getPop();
}
double
ref_bf_fpr(void* f)
{
	return ((RefBF*)f)->getFPR();
}
void
ref_bf_set_meta(void* f, double dFPR, uint64_t nEntry, uint64_t tEntry)
{
	((RefBF*)f)->setdFPR(dFPR);
	((RefBF*)f)->setnEntry(nEntry);
	((RefBF*)f)->settEntry(tEntry);
}
void
ref_bf_get_meta(void* f, uint64_t* nEntry, uint64_t* tEntry)
{
	*nEntry = ((RefBF*)f)->getnEntry();
	*tEntry = ((RefBF*)f)->gettEntry();
}
void
ref_bf_store(void* f, const char* path)
{
	((RefBF*)f)->storeFilter(path);
}

uint64_t
ref_bf_insert_seqs(void* f, const char* bases, const uint64_t* off, uint64_t n_seqs)
{
	RefBF& bloom = *(RefBF*)f;
	unsigned h = bloom.getHashNum(), k = bloom.getKmerSize();
	uint64_t n = 0;
	for (uint64_t s = 0; s < n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		ntHashIterator itr(seq, h, k);
		while (itr != itr.end()) {
			bloom.insert(*itr);
			++n;
			++itr;
		}
	}
	return n;
}

uint64_t
ref_bf_contains_seqs(void* f, const char* bases, const uint64_t* off, uint64_t n_seqs,
                     uint8_t* hit_bits, uint8_t* valid_bits, uint64_t* n_hits)
{
	RefBF& bloom = *(RefBF*)f;
	unsigned h = bloom.getHashNum(), k = bloom.getKmerSize();
	uint64_t n = 0, hits = 0;
	for (uint64_t s = 0; s < n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		ntHashIterator itr(seq, h, k);
		while (itr != itr.end()) {
			uint64_t p = off[s] + itr.pos();
			if (valid_bits)
				set_bit(valid_bits, p);
			if (bloom.contains(*itr)) {
				if (hit_bits)
					set_bit(hit_bits, p);
				++hits;
			}
			++n;
			++itr;
		}
	}
	if (n_hits)
		*n_hits = hits;
	return n;
}

uint64_t
ref_bf_insert_and_check_seqs(void* f, const char* bases, const uint64_t* off, uint64_t n_seqs,
                             uint8_t* found_bits, uint8_t* valid_bits)
{
	RefBF& bloom = *(RefBF*)f;
	unsigned h = bloom.getHashNum(), k = bloom.getKmerSize();
	uint64_t n = 0;
	for (uint64_t s = 0; s < n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		ntHashIterator itr(seq, h, k);
		while (itr != itr.end()) {
			uint64_t p = off[s] + itr.pos();
			if (valid_bits)
				set_bit(valid_bits, p);
			if (bloom.insertAndCheck(*itr) && found_bits)
				set_bit(found_bits, p);
			++n;
			++itr;
		}
	}
	return n;
}

uint64_t
ref_st_bf_insert_seqs(void* f, const char* const* seeds, unsigned n_seeds, unsigned h2,
                      const char* bases, const uint64_t* off, uint64_t n_seqs)
{
	RefBF& bloom = *(RefBF*)f;
	auto ss = parse(seeds, n_seeds);
	unsigned k = bloom.getKmerSize();
	uint64_t n = 0;
	for (uint64_t s = 0; s < n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		stHashIterator itr(seq, ss, n_seeds, h2, k);
		while (itr != itr.end()) {
			bloom.insert(*itr);
			++n;
			++itr;
		}
	}
	return n;
}

uint64_t
ref_st_bf_contains_seqs(void* f, const char* const* seeds, unsigned n_seeds, unsigned h2,
                        const char* bases, const uint64_t* off, uint64_t n_seqs, uint8_t* hit_bits,
                        uint8_t* valid_bits, uint64_t* n_hits)
{
	RefBF& bloom = *(RefBF*)f;
	auto ss = parse(seeds, n_seeds);
	unsigned k = bloom.getKmerSize();
	uint64_t n = 0, hits = 0;
	for (uint64_t s = 0; s < n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		stHashIterator itr(seq, ss, n_seeds, h2, k);
		while (itr != itr.end()) {
			uint64_t p = off[s] + itr.pos();
			if (valid_bits)
				set_bit(valid_bits, p);
			if (bloom.contains(*itr)) {
				if (hit_bits)
					set_bit(hit_bits, p);
				++hits;
			}
			++n;
			++itr;
		}
	}
	if (n_hits)
		*n_hits = hits;
	return n;
}

/* ---------- KmerBloomFilter (KmerBloomFilter.hpp:47-74): one k-mer as text, table-driven NTC64/NTE64 ---------- */
struct RefKBF : public KmerBloomFilter
{
	RefKBF(size_t bits, unsigned h, unsigned k)
	  : KmerBloomFilter(bits, h, k)
	{}
	uint8_t* data() { return m_filter; }
};
void*
ref_kbf_new(uint64_t bits, unsigned h, unsigned k)
{
	return new RefKBF(bits, h, k);
}
void
ref_kbf_free(void* f)
{
	delete (RefKBF*)f;
}
uint8_t*
ref_kbf_data(void* f)
{
	return ((RefKBF*)f)->data();
}
void
ref_kbf_insert(void* f, const char* kmer)
{
	((RefKBF*)f)->insert(kmer);
}
int
ref_kbf_contains(void* f, const char* kmer)
{
	return ((RefKBF*)f)->contains(kmer) ? 1 : 0;
}

/* ---------- CountingBloomFilter<uint8_t> ---------- */
void*
ref_cbf_new(uint64_t bytes, unsigned h, unsigned k, unsigned thr)
{
	return new RefCBF(bytes, h, k, thr);
}
void*
ref_cbf_load(const char* path, unsigned thr)
{
	return new RefCBF(std::string(path), thr);
}
void
ref_cbf_free(void* f)
{
	delete (RefCBF*)f;
}
uint64_t
ref_cbf_size(void* f)
{
	return ((RefCBF*)f)->size();
}
uint64_t
ref_cbf_size_bytes(void* f)
{
	return ((RefCBF*)f)->sizeInBytes();
}
unsigned
ref_cbf_hash_num(void* f)
{
	return ((RefCBF*)f)->getHashNum();
}
unsigned
ref_cbf_kmer_size(void* f)
{
	return ((RefCBF*)f)->getKmerSize();
}
uint64_t
ref_cbf_popcount(void* f)
{
	return ((RefCBF*)f)->popCount();
}
uint64_t
ref_cbf_filtered_popcount(void* f)
{
	return ((RefCBF*)f)->filtered_popcount();
}
void
ref_cbf_dump(void* f, uint8_t* out)
{
	RefCBF& c = *(RefCBF*)f;
	for (size_t i = 0; i < c.size(); i++)
		out[i] = c[i];
}
void
ref_cbf_store(void* f, const char* path)
{
	((RefCBF*)f)->storeFilter(path);
}

uint64_t
ref_cbf_insert_seqs(void* f, const char* bases, const uint64_t* off, uint64_t n_seqs)
{
	RefCBF& c = *(RefCBF*)f;
	unsigned h = c.getHashNum(), k = c.getKmerSize();
	uint64_t n = 0;
	for (uint64_t s = 0; s < n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		ntHashIterator itr(seq, h, k);
		while (itr != itr.end()) {
			c.insert(*itr);
			++n;
			++itr;
		}
	}
	return n;
}

uint64_t
ref_cbf_increment_all_seqs(void* f, const char* bases, const uint64_t* off, uint64_t n_seqs)
{
	RefCBF& c = *(RefCBF*)f;
	unsigned h = c.getHashNum(), k = c.getKmerSize();
	uint64_t n = 0;
	for (uint64_t s = 0; s < n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		ntHashIterator itr(seq, h, k);
		while (itr != itr.end()) {
			c.incrementAll(*itr);
			++n;
			++itr;
		}
	}
	return n;
}

uint64_t
ref_cbf_mincount_seqs(void* f, const char* bases, const uint64_t* off, uint64_t n_seqs,
                      uint8_t* counts, uint8_t* valid_bits)
{
	RefCBF& c = *(RefCBF*)f;
	unsigned h = c.getHashNum(), k = c.getKmerSize();
	uint64_t n = 0;
	for (uint64_t s = 0; s < n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		ntHashIterator itr(seq, h, k);
		while (itr != itr.end()) {
			uint64_t p = off[s] + itr.pos();
			if (valid_bits)
				set_bit(valid_bits, p);
			if (counts)
				counts[p] = c.minCount(*itr);
			++n;
			++itr;
		}
	}
	return n;
}

uint64_t
ref_cbf_contains_seqs(void* f, const char* bases, const uint64_t* off, uint64_t n_seqs,
                      uint8_t* hit_bits, uint8_t* valid_bits, uint64_t* n_hits)
{
	RefCBF& c = *(RefCBF*)f;
	unsigned h = c.getHashNum(), k = c.getKmerSize();
	uint64_t n = 0, hits = 0;
	for (uint64_t s = 0; s < n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		ntHashIterator itr(seq, h, k);
		while (itr != itr.end()) {
			uint64_t p = off[s] + itr.pos();
			if (valid_bits)
				set_bit(valid_bits, p);
			if (c.contains(*itr)) {
				if (hit_bits)
					set_bit(hit_bits, p);
				++hits;
			}
			++n;
			++itr;
		}
	}
	if (n_hits)
		*n_hits = hits;
	return n;
}

uint64_t
ref_st_cbf_insert_seqs(void* f, const char* const* seeds, unsigned n_seeds, unsigned h2,
                       const char* bases, const uint64_t* off, uint64_t n_seqs)
{
	RefCBF& c = *(RefCBF*)f;
	auto ss = parse(seeds, n_seeds);
	unsigned k = c.getKmerSize();
	uint64_t n = 0;
	for (uint64_t s = 0; s < n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		stHashIterator itr(seq, ss, n_seeds, h2, k);
		while (itr != itr.end()) {
			c.insert(*itr);
			++n;
			++itr;
		}
	}
	return n;
}

uint64_t
ref_st_cbf_mincount_seqs(void* f, const char* const* seeds, unsigned n_seeds, unsigned h2,
                         const char* bases, const uint64_t* off, uint64_t n_seqs, uint8_t* counts,
                         uint8_t* valid_bits)
{
	RefCBF& c = *(RefCBF*)f;
	auto ss = parse(seeds, n_seeds);
	unsigned k = c.getKmerSize();
	uint64_t n = 0;
	for (uint64_t s = 0; s < n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		stHashIterator itr(seq, ss, n_seeds, h2, k);
		while (itr != itr.end()) {
			uint64_t p = off[s] + itr.pos();
			if (valid_bits)
				set_bit(valid_bits, p);
			if (counts)
				counts[p] = c.minCount(*itr);
			++n;
			++itr;
		}
	}
	return n;
}

/* ---------- OpenMP timing legs (Tests/AdHoc/ParallelFilter.cpp:104-122 pattern) ---------- */
int
ref_max_threads(void)
{
#ifdef _OPENMP
	return omp_get_max_threads();
#else
	return 1;
#endif
}

/* mode: 0 = BF contains, 1 = BF insert.  Returns seconds. */
double
ref_bench_bf(void* f, const char* bases, const uint64_t* off, uint64_t n_seqs, int do_insert,
             int threads, uint64_t* n_kmers, uint64_t* n_hits)
{
	RefBF& bloom = *(RefBF*)f;
	unsigned h = bloom.getHashNum(), k = bloom.getKmerSize();
	uint64_t n = 0, hits = 0;
#ifdef _OPENMP
	if (threads > 0)
		omp_set_num_threads(threads);
#endif
	(void)threads;
	auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : n, hits)
	for (int64_t s = 0; s < (int64_t)n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		ntHashIterator itr(seq, h, k);
		while (itr != itr.end()) {
			if (do_insert)
				bloom.insert(*itr);
			else
				hits += bloom.contains(*itr);
			++n;
			++itr;
		}
	}
	auto t1 = std::chrono::steady_clock::now();
	if (n_kmers)
		*n_kmers = n;
	if (n_hits)
		*n_hits = hits;
	return std::chrono::duration<double>(t1 - t0).count();
}

double
ref_bench_cbf(void* f, const char* bases, const uint64_t* off, uint64_t n_seqs, int do_insert,
              int threads, uint64_t* n_kmers, uint64_t* n_hits)
{
	RefCBF& c = *(RefCBF*)f;
	unsigned h = c.getHashNum(), k = c.getKmerSize();
	uint64_t n = 0, hits = 0;
#ifdef _OPENMP
	if (threads > 0)
		omp_set_num_threads(threads);
#endif
	(void)threads;
	auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : n, hits)
	for (int64_t s = 0; s < (int64_t)n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		ntHashIterator itr(seq, h, k);
		while (itr != itr.end()) {
			if (do_insert)
				c.insert(*itr);
			else
				hits += c.contains(*itr);
			++n;
			++itr;
		}
	}
	auto t1 = std::chrono::steady_clock::now();
	if (n_kmers)
		*n_kmers = n;
	if (n_hits)
		*n_hits = hits;
	return std::chrono::duration<double>(t1 - t0).count();
}

/* Spaced seeds (stHashIterator.hpp:53-57 + BloomFilter insert/contains), same OpenMP pattern. */
double
ref_bench_st_bf(void* f, const char* const* seeds, unsigned n_seeds, unsigned h2, const char* bases,
                const uint64_t* off, uint64_t n_seqs, int do_insert, int threads, uint64_t* n_kmers,
                uint64_t* n_hits)
{
	RefBF& bloom = *(RefBF*)f;
	auto ss = parse(seeds, n_seeds);
	unsigned k = bloom.getKmerSize();
	uint64_t n = 0, hits = 0;
#ifdef _OPENMP
	if (threads > 0)
		omp_set_num_threads(threads);
#endif
	(void)threads;
	auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : n, hits)
	for (int64_t s = 0; s < (int64_t)n_seqs; s++) {
		std::string seq(bases + off[s], off[s + 1] - off[s]);
		stHashIterator itr(seq, ss, n_seeds, h2, k);
		while (itr != itr.end()) {
			if (do_insert)
				bloom.insert(*itr);
			else
				hits += bloom.contains(*itr);
			++n;
			++itr;
		}
	}
	auto t1 = std::chrono::steady_clock::now();
	if (n_kmers)
		*n_kmers = n;
	if (n_hits)
		*n_hits = hits;
	return std::chrono::duration<double>(t1 - t0).count();
}

} // extern "C"
