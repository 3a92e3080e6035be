/* SWIG interface for the GPU-backed classes (the counterpart of the reference's swig/BloomFilter.i).
 * `swig` is not installed in the build image, so this file is shipped but not generated/tested here:
 *     swig -Wall -c++ -perl5 BloomFilter.i
 *     g++ -std=c++11 -fPIC -c BloomFilter_wrap.cxx -I../include $(perl -MExtUtils::Embed -e ccopts)
 *     g++ -shared BloomFilter_wrap.o -o BloomFilter.so -L../btl_bloomfilter_b200 -lbtlbf_cuda
 * As in the reference, KmerBloomFilter is exported under the name BloomFilter. */
%module BloomFilter
%include "std_string.i"
%include "stdint.i"
%include "std_vector.i"
namespace std {
   %template(SizetVector) vector<size_t>;
   %template(StringVector) vector<string>;
}

%{
#include "btlbf/KmerBloomFilter.hpp"
#include "btlbf/BloomFilterUtil.h"
%}

%rename(BloomFilter) KmerBloomFilter;

using namespace std;

class KmerBloomFilter {
public:
        KmerBloomFilter();
        ~KmerBloomFilter();
        KmerBloomFilter(uint64_t filterSize, unsigned hashNum, unsigned kmerSize);
        KmerBloomFilter(const string &filterFilePath);

        void insert(vector<uint64_t> const &precomputed);
        void insert(const char* kmer);

        bool contains(vector<uint64_t> const &values);
        bool contains(const char* kmer);

        /* new: batched GPU path (ntHashIterator + insert/contains fused, one call per batch of sequences) */
        uint64_t insertSeqs(vector<string> const &seqs);

        void storeFilter(string const &filterFilePath);
        uint64_t getPop();
        unsigned getHashNum();
        unsigned getKmerSize();
        uint64_t getFilterSize();
};

void insertSeq(KmerBloomFilter &bloom, const string& seq, unsigned numHashes, unsigned k);
