// btlbf/KmerBloomFilter.hpp -- drop-in for the reference's KmerBloomFilter (KmerBloomFilter.hpp:17-74), the
// class its SWIG module exports as "BloomFilter" (swig/BloomFilter.i:17): BloomFilter + insert/contains of one
// k-mer given as text.  The k-mer is hashed on the GPU by the same fused kernel as the batched path (a batch of
// one sequence of kmerSize bases), i.e. with the iterator-consistent canonical ntHash.  For kmerSize % 4 != 0
// that equals the reference's table-driven NTC64(kmer,k)/NTE64 values; for kmerSize % 4 == 0 the reference's
// own result is undefined behaviour (shift by 64, nthash.hpp:356,389) and disagrees with its iterator, so the
// iterator value is used.  k-mers containing a non-ACGTU byte are ignored by insert and never contained.
#ifndef BTLBF_KMERBLOOMFILTER_HPP
#define BTLBF_KMERBLOOMFILTER_HPP

#include <cstring>

#include "BloomFilter.hpp"

class KmerBloomFilter : public BloomFilter
{
  public:
	KmerBloomFilter() = default;
	KmerBloomFilter(size_t filterSize, unsigned hashNum, unsigned kmerSize, int device = 0)
	  : BloomFilter(filterSize, hashNum, kmerSize, device)
	{}
	explicit KmerBloomFilter(const std::string& filterFilePath, int device = 0)
	  : BloomFilter(filterFilePath, device)
	{}

	using BloomFilter::contains;
	using BloomFilter::insert;

	bool contains(const char* kmer) const // KmerBloomFilter.hpp:47-61
	{
		uint64_t off[2] = { 0, m_kmerSize };
		uint8_t hit[4] = { 0, 0, 0, 0 };
		uint64_t nk = 0, nh = 0;
		btlbf::check(btlbf_contains_seqs(m_f, kmer, off, 1, hit, nullptr, &nk, &nh), "contains");
		return nh == 1;
	}
	void insert(const char* kmer) // KmerBloomFilter.hpp:63-74
	{
		uint64_t off[2] = { 0, m_kmerSize };
		insertSeqs(kmer, off, 1);
	}
	// many k-mers of kmerSize bases each, concatenated
	uint64_t insertKmers(const char* kmers, uint64_t n)
	{
		std::vector<uint64_t> off(n + 1);
		for (uint64_t i = 0; i <= n; i++)
			off[i] = i * m_kmerSize;
		return insertSeqs(kmers, off.data(), n);
	}
};

#endif
