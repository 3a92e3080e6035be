// btlbf/CountingBloomFilter.hpp -- drop-in C++ host class for the reference's CountingBloomFilter<T>
// (CountingBloomFilter.hpp) with T = uint8_t: counters in B200 HBM, operations as CUDA kernels behind
// include/btlbf.h.  insert (incrementMin) is order-dependent; the batched insertSeqs reproduces the
// reference's single-threaded, read-order, position-order result exactly.
#ifndef BTLBF_COUNTINGBLOOMFILTER_HPP
#define BTLBF_COUNTINGBLOOMFILTER_HPP

#include <cmath>
#include <ostream>
#include <string>
#include <vector>

#include "Device.hpp"

template<typename T>
class CountingBloomFilter
{
	static_assert(sizeof(T) == 1, "the GPU path implements CountingBloomFilter<uint8_t> only");

  public:
	CountingBloomFilter() = default;
	// sizeInBytes is rounded up to a multiple of 8.  CountingBloomFilter.hpp:29-50
	CountingBloomFilter(size_t sizeInBytes, unsigned hashNum, unsigned kmerSize, unsigned countThreshold, int device = 0)
	  : m_ctx(btlbf::defaultContext(device))
	  , m_countThreshold(countThreshold)
	{
		size_t remainder = sizeInBytes % 8;
		if (remainder != 0)
			sizeInBytes += 8 - remainder;
		btlbf::check(btlbf_filter_create(m_ctx, BTLBF_COUNTING8, sizeInBytes, hashNum, kmerSize, countThreshold, &m_f),
		             "allocating the filter");
		refreshInfo();
	}
	CountingBloomFilter(const std::string& path, unsigned countThreshold, int device = 0) // :260-266
	  : m_ctx(btlbf::defaultContext(device))
	  , m_countThreshold(countThreshold)
	{
		loadFilter(path);
	}
	CountingBloomFilter(const CountingBloomFilter&) = delete;
	CountingBloomFilter& operator=(const CountingBloomFilter&) = delete;
	~CountingBloomFilter() { btlbf_filter_destroy(m_f); }

	T operator[](size_t i) // :51 (reads the whole array back: for tests, not for hot loops)
	{
		std::vector<uint8_t> host(m_sizeInBytes);
		btlbf::check(btlbf_filter_download(m_f, host.data(), host.size()), "reading the filter back");
		return (T)host[i];
	}

	// ---- per-k-mer interface (hashes: anything indexable with m_hashNum values, e.g. *ntHashIterator)
	template<typename U>
	T minCount(const U& hashes) const // :52-64
	{
		uint64_t hv[64];
		uint8_t c = 0;
		btlbf::check(btlbf_mincount_hashes(m_f, gather(hashes, hv), 1, &c), "minCount");
		return (T)c;
	}
	template<typename U>
	bool contains(const U& hashes) const // :190-196
	{
		return minCount(hashes) >= m_countThreshold;
	}
	template<typename U>
	void insert(const U& hashes) // :198-204
	{
		incrementMin(hashes);
	}
	template<typename U>
	bool insertAndCheck(const U& hashes) // :206-214
	{
		uint64_t hv[64];
		uint8_t found = 0;
		btlbf::check(btlbf_insert_hashes(m_f, gather(hashes, hv), 1, &found), "insertAndCheck");
		return found != 0;
	}
	template<typename U>
	void incrementMin(const U& hashes) // :134-162
	{
		uint64_t hv[64];
		btlbf::check(btlbf_insert_hashes(m_f, gather(hashes, hv), 1, nullptr), "incrementMin");
	}
	template<typename U>
	void incrementAll(const U& hashes) // :164-183
	{
		uint64_t hv[64];
		btlbf::check(btlbf_increment_all_hashes(m_f, gather(hashes, hv), 1), "incrementAll");
	}

	// ---- batched entry points (README.md:86-113 for a whole batch)
	uint64_t insertSeqs(const btlbf::SeqBatch& b)
	{
		uint64_t n = 0;
		btlbf::check(btlbf_insert_seqs(m_f, b.bases.data(), b.offsets.data(), b.size(), &n), "insertSeqs");
		return n;
	}
	uint64_t insertSeqs(const std::vector<std::string>& seqs) { return insertSeqs(btlbf::SeqBatch(seqs)); }
	uint64_t insertSeqs(const btlbf::PackedSeqBatch& b) // 2-bit packed input: same counters as the ASCII batch
	{
		uint64_t n = 0;
		btlbf::check(btlbf_insert_seqs_packed(m_f, b.codes.data(), b.invalidPlane(), b.offsets.data(), b.size(), &n), "insertSeqs");
		return n;
	}
	btlbf::SeqHits containsSeqs(const btlbf::PackedSeqBatch& b) const
	{
		btlbf::SeqHits r;
		r.hitBits.assign(btlbf::bitBytes(b.nBases()), 0);
		r.validBits.assign(btlbf::bitBytes(b.nBases()), 0);
		btlbf::check(btlbf_contains_seqs_packed(m_f, b.codes.data(), b.invalidPlane(), b.offsets.data(), b.size(),
		                                        r.hitBits.data(), r.validBits.data(), &r.nKmers, &r.nHits),
		             "containsSeqs");
		return r;
	}
	btlbf::SeqHits containsSeqs(const btlbf::SeqBatch& b) const
	{
		btlbf::SeqHits r;
		uint64_t n = b.bases.size();
		r.hitBits.assign(btlbf::bitBytes(n), 0);
		r.validBits.assign(btlbf::bitBytes(n), 0);
		btlbf::check(btlbf_contains_seqs(m_f, b.bases.data(), b.offsets.data(), b.size(), r.hitBits.data(),
		                                 r.validBits.data(), &r.nKmers, &r.nHits),
		             "containsSeqs");
		return r;
	}
	btlbf::SeqHits containsSeqs(const std::vector<std::string>& seqs) const { return containsSeqs(btlbf::SeqBatch(seqs)); }
	// minCount of every window: counts[p] (0 for invalid windows)
	std::vector<T> minCountSeqs(const btlbf::SeqBatch& b, btlbf::SeqHits* validOut = nullptr) const
	{
		uint64_t n = b.bases.size();
		std::vector<T> counts(n, 0);
		std::vector<uint8_t> valid(btlbf::bitBytes(n), 0);
		uint64_t nk = 0;
		btlbf::check(btlbf_mincount_seqs(m_f, b.bases.data(), b.offsets.data(), b.size(),
		                                 reinterpret_cast<uint8_t*>(counts.data()), valid.data(), &nk),
		             "minCountSeqs");
		if (validOut) {
			validOut->validBits.swap(valid);
			validOut->nKmers = nk;
		}
		return counts;
	}
	uint64_t incrementAllSeqs(const btlbf::SeqBatch& b)
	{
		uint64_t n = 0;
		btlbf::check(btlbf_increment_all_seqs(m_f, b.bases.data(), b.offsets.data(), b.size(), &n), "incrementAllSeqs");
		return n;
	}
	void setSeeds(const std::vector<std::string>& seeds, unsigned h2 = 1)
	{
		std::vector<const char*> p;
		for (const auto& s : seeds)
			p.push_back(s.c_str());
		btlbf::check(btlbf_filter_set_seeds(m_f, p.data(), (unsigned)p.size(), h2), "setSeeds");
	}

	// ---- accessors (:65-75)
	unsigned getKmerSize() const { return m_kmerSize; }
	unsigned getHashNum() const { return m_hashNum; }
	unsigned threshold() const { return m_countThreshold; }
	size_t size() const { return m_size; }
	size_t sizeInBytes() const { return m_sizeInBytes; }
	size_t popCount() const // :216-228
	{
		uint64_t n = 0;
		btlbf::check(btlbf_filter_popcount(m_f, &n), "popCount");
		return (size_t)n;
	}
	size_t filtered_popcount() const // :230-242
	{
		uint64_t n = 0;
		btlbf::check(btlbf_filter_count_ge(m_f, m_countThreshold, &n), "filtered_popcount");
		return (size_t)n;
	}
	double FPR() const { return std::pow((double)popCount() / (double)m_size, m_hashNum); }                   // :244-250
	double filtered_FPR() const { return std::pow((double)filtered_popcount() / (double)m_size, m_hashNum); } // :252-258

	// ---- file layout (BTLCountingBloomFilter_v1)
	void loadFilter(const std::string& path) // :268-281 (+ loadHeader :283-329)
	{
		if (!m_ctx)
			m_ctx = btlbf::defaultContext(0);
		btlbf_filter* f = nullptr;
		btlbf::check(btlbf_filter_load(m_ctx, path.c_str(), BTLBF_COUNTING8, m_countThreshold, &f, nullptr, nullptr, nullptr),
		             path.c_str());
		btlbf_filter_destroy(m_f);
		m_f = f;
		refreshInfo();
	}
	void storeHeader(std::ostream& out) const // :344-367
	{
		char buf[1024];
		size_t len = 0;
		btlbf::check(btlbf_format_header(BTLBF_COUNTING8, m_size, m_sizeInBytes, m_hashNum, m_kmerSize, 0, 0, 0, buf,
		                                 sizeof buf, &len),
		             "storeHeader");
		out.write(buf, (std::streamsize)len);
	}
	void storeFilter(const std::string& path) const // :331-342
	{
		std::cerr << "Writing a " << m_sizeInBytes << " byte filter to " << path << " on disk.\n";
		btlbf::check(btlbf_filter_store(m_f, path.c_str(), 0, 0, 0), path.c_str());
	}
	friend std::ostream& operator<<(std::ostream& out, const CountingBloomFilter& bloom) // :370-379
	{
		bloom.storeHeader(out);
		std::vector<uint8_t> host(bloom.m_sizeInBytes);
		btlbf::check(btlbf_filter_download(bloom.m_f, host.data(), host.size()), "reading the filter back");
		out.write(reinterpret_cast<const char*>(host.data()), (std::streamsize)host.size());
		return out;
	}

	btlbf_filter* handle() const { return m_f; }

  private:
	template<typename U>
	const uint64_t* gather(const U& hashes, uint64_t* tmp) const
	{
		for (unsigned i = 0; i < m_hashNum && i < 64; ++i)
			tmp[i] = hashes[i];
		return tmp;
	}
	void refreshInfo()
	{
		uint64_t size = 0, bytes = 0;
		btlbf::check(btlbf_filter_info(m_f, nullptr, &size, &bytes, &m_hashNum, &m_kmerSize, nullptr), "filter info");
		m_size = size;
		m_sizeInBytes = bytes;
	}

	btlbf_ctx* m_ctx = nullptr;
	btlbf_filter* m_f = nullptr;
	size_t m_size = 0;
	size_t m_sizeInBytes = 0;
	unsigned m_hashNum = 0;
	unsigned m_kmerSize = 0;
	unsigned m_countThreshold = 0;
};

#endif
