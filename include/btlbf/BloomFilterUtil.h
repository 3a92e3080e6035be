// btlbf/BloomFilterUtil.h -- the reference's sequence-level helper (BloomFilterUtil.h:10-17) over the GPU path.
#ifndef BTLBF_BLOOMFILTERUTIL_H
#define BTLBF_BLOOMFILTERUTIL_H

#include <string>

#include "BloomFilter.hpp"

// Loads every k-mer of seq into the filter (ntHashIterator + insert, fused on the GPU).
// hashNum / kmerSize must be the filter's own, as in every caller of the reference.
inline void
insertSeq(BloomFilter& bloom, const std::string& seq, unsigned hashNum, unsigned kmerSize)
{
	if (hashNum != bloom.getHashNum() || kmerSize != bloom.getKmerSize()) {
		std::cerr << "ERROR: insertSeq: hashNum/kmerSize differ from the filter's" << std::endl;
		exit(EXIT_FAILURE);
	}
	uint64_t off[2] = { 0, seq.size() };
	bloom.insertSeqs(seq.data(), off, 1);
}

// BloomFilterUtil.h:29-33
inline double
calcApproxFPR(size_t size, size_t numEntr, unsigned hashFunctNum)
{
	return pow(1.0 - pow(1.0 - 1.0 / double(size), double(numEntr) * hashFunctNum), double(hashFunctNum));
}

// BloomFilterUtil.h:39-46
inline double
calcRedunancyFPR(size_t size, size_t numEntr, unsigned hashFunctNum)
{
	double total = log(calcApproxFPR(size, 1, hashFunctNum));
	for (size_t i = 2; i < numEntr; ++i)
		total = log(exp(total) + calcApproxFPR(size, i, hashFunctNum));
	return exp(total) / numEntr;
}

#endif
