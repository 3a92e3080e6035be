// btlbf/BloomFilter.hpp -- drop-in C++ host class for the reference's BloomFilter (BloomFilter.hpp),
// whose bit array lives in B200 HBM and whose operations run as CUDA kernels behind include/btlbf.h.
//
// Same public names and signatures as the reference class (file:line of the member it replaces is
// given at each method), the BTLBloomFilter_v1 file layout byte for byte, and the new batched entry
// points insertSeqs / containsSeqs / insertAndCheckSeqs that fuse ntHashIterator with insert/contains.
// Header-only; link with -lbtlbf_cuda.  Not copyable (as in the reference, BloomFilter.hpp:384).
#ifndef BTLBF_BLOOMFILTER_HPP
#define BTLBF_BLOOMFILTER_HPP

#include <cassert>
#include <cmath>
#include <fstream>
#include <istream>
#include <ostream>
#include <string>
#include <vector>

#include "Device.hpp"

class BloomFilter
{
  public:
	BloomFilter() = default; // BloomFilter.hpp:46-56

	// De novo filter of filterSize bits (multiple of 8).  BloomFilter.hpp:66-78
	BloomFilter(size_t filterSize, unsigned hashNum, unsigned kmerSize, int device = 0)
	  : m_ctx(btlbf::defaultContext(device))
	{
		create(filterSize, hashNum, kmerSize);
	}

	// Sized from the expected number of elements and an FPR; hashNum == 0 picks the optimum.  :85-104
	BloomFilter(size_t expectedElemNum, double fpr, unsigned hashNum, unsigned kmerSize, int device = 0)
	  : m_ctx(btlbf::defaultContext(device))
	  , m_dFPR(fpr)
	{
		if (hashNum == 0)
			hashNum = calcOptiHashNum(fpr);
		create(calcOptimalSize(expectedElemNum, fpr, hashNum), hashNum, kmerSize);
	}

	// From a BTLBloomFilter_v1 file.  :106-110
	explicit BloomFilter(const std::string& filterFilePath, int device = 0)
	  : m_ctx(btlbf::defaultContext(device))
	{
		loadFilter(filterFilePath);
	}

	BloomFilter(const BloomFilter&) = delete;
	BloomFilter& operator=(const BloomFilter&) = delete;

	~BloomFilter() { btlbf_filter_destroy(m_f); }

	void loadFilter(const std::string& filterFilePath) // :112-121 (+ loadHeader :123-166)
	{
		if (!m_ctx)
			m_ctx = btlbf::defaultContext(0);
		btlbf_filter* f = nullptr;
		btlbf::check(
		    btlbf_filter_load(m_ctx, filterFilePath.c_str(), BTLBF_BLOOM, 0, &f, &m_dFPR, &m_nEntry, &m_tEntry),
		    filterFilePath.c_str());
		btlbf_filter_destroy(m_f);
		m_f = f;
		refreshInfo();
	}

	// Reads the header ("[BTLBloomFilter_v1]" ... "[HeaderEnd]") from the stream, which is left at the first byte
	// of the raw array, sets the members from it and allocates a zeroed filter of that size (initSize).  :118-166
	void loadHeader(std::istream& file)
	{
		if (!m_ctx)
			m_ctx = btlbf::defaultContext(0);
		std::string text, line;
		bool headerEnd = false;
		while (std::getline(file, line)) {
			text.append(line + "\n");
			if (line == "[HeaderEnd]") {
				headerEnd = true;
				break;
			}
			if (text.size() == line.size() + 1 && line != "[BTLBloomFilter_v1]")
				break; // wrong magic: no need to read on
		}
		(void)headerEnd; // the parser reports a missing header end like the reference does
		uint64_t size = 0, bytes = 0;
		unsigned h = 0, k = 0;
		btlbf::check(btlbf_parse_header(BTLBF_BLOOM, text.data(), text.size(), &size, &bytes, &h, &k, &m_dFPR, &m_nEntry,
		                                &m_tEntry, nullptr),
		             "loadHeader");
		btlbf_filter_destroy(m_f);
		m_f = nullptr;
		create(size, h, k);
	}

	// ---- per-k-mer interface: the caller supplies the m_hashNum hash values (e.g. *ntHashIterator)
	void insert(std::vector<uint64_t> const& precomputed) { insert(checked(precomputed)); } // :171-180
	void insert(const uint64_t precomputed[])                                               // :185-194
	{
		btlbf::check(btlbf_insert_hashes(m_f, precomputed, 1, nullptr), "insert");
	}
	bool insertAndCheck(const uint64_t precomputed[]) // :200-214
	{
		uint8_t found = 0;
		btlbf::check(btlbf_insert_hashes(m_f, precomputed, 1, &found), "insertAndCheck");
		return found != 0;
	}
	bool insertAndCheck(std::vector<uint64_t> const& precomputed) { return insertAndCheck(checked(precomputed)); }
	bool contains(std::vector<uint64_t> const& precomputed) const { return contains(checked(precomputed)); } // :237-247
	bool contains(const uint64_t precomputed[]) const                                                        // :252-262
	{
		uint8_t hit = 0;
		btlbf::check(btlbf_contains_hashes(m_f, precomputed, 1, &hit), "contains");
		return hit != 0;
	}
	// many k-mers per call: hashes[i*m_hashNum + j]
	void insert(const uint64_t* hashes, uint64_t nKmers)
	{
		btlbf::check(btlbf_insert_hashes(m_f, hashes, nKmers, nullptr), "insert");
	}
	void contains(const uint64_t* hashes, uint64_t nKmers, uint8_t* hit) const
	{
		btlbf::check(btlbf_contains_hashes(m_f, hashes, nKmers, hit), "contains");
	}

	// ---- batched entry points: ntHashIterator over every sequence fused with the filter operation.
	// Replace the loop of README.md:30-43 / BloomFilterUtil.h:10-17 (insertSeq) for a whole batch.
	uint64_t insertSeqs(const btlbf::SeqBatch& b) { return insertSeqs(b.bases.data(), b.offsets.data(), b.size()); }
	uint64_t insertSeqs(const std::vector<std::string>& seqs) { return insertSeqs(btlbf::SeqBatch(seqs)); }
	uint64_t insertSeqs(const char* bases, const uint64_t* offsets, uint64_t nSeqs)
	{
		uint64_t n = 0;
		btlbf::check(btlbf_insert_seqs(m_f, bases, offsets, nSeqs, &n), "insertSeqs");
		return n;
	}
	uint64_t insertSeqs(const btlbf::PackedSeqBatch& b) // 2-bit packed input: same filter bytes as the ASCII batch
	{
		uint64_t n = 0;
		btlbf::check(btlbf_insert_seqs_packed(m_f, b.codes.data(), b.invalidPlane(), b.offsets.data(), b.size(), &n), "insertSeqs");
		return n;
	}
	// The query twin (README.md:46-57): one hit bit and one valid bit per window.
	// FASTA / FASTQ files: the record loop of swig/writeBloom_rolling.cpp:19-59 (read a record, insertSeq) as one
	// call; threads = 0 picks the number of parser threads.  Returns the k-mers inserted / found.
	uint64_t insertFile(const std::string& path, int threads = 0, uint64_t* nSeqs = nullptr)
	{
		uint64_t n = 0;
		btlbf::check(btlbf_insert_file(m_f, path.c_str(), threads, nSeqs, &n), path.c_str());
		return n;
	}
	uint64_t queryFile(const std::string& path, uint64_t* nKmers = nullptr, int threads = 0, uint64_t* nSeqs = nullptr) const
	{
		uint64_t hits = 0;
		btlbf::check(btlbf_query_file(m_f, path.c_str(), threads, nSeqs, nKmers, &hits), path.c_str());
		return hits;
	}

	btlbf::SeqHits containsSeqs(const btlbf::SeqBatch& b) const
	{
		return containsSeqs(b.bases.data(), b.offsets.data(), b.size());
	}
	btlbf::SeqHits containsSeqs(const std::vector<std::string>& seqs) const { return containsSeqs(btlbf::SeqBatch(seqs)); }
	btlbf::SeqHits containsSeqs(const char* bases, const uint64_t* offsets, uint64_t nSeqs) const
	{
		btlbf::SeqHits r;
		uint64_t n = nSeqs ? offsets[nSeqs] : 0;
		r.hitBits.assign(btlbf::bitBytes(n), 0);
		r.validBits.assign(btlbf::bitBytes(n), 0);
		btlbf::check(
		    btlbf_contains_seqs(m_f, bases, offsets, nSeqs, r.hitBits.data(), r.validBits.data(), &r.nKmers, &r.nHits),
		    "containsSeqs");
		return r;
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
	// insertAndCheck of every k-mer in the reference's order: hit bit = "was already present".
	btlbf::SeqHits insertAndCheckSeqs(const btlbf::SeqBatch& b)
	{
		btlbf::SeqHits r;
		uint64_t n = b.bases.size();
		r.hitBits.assign(btlbf::bitBytes(n), 0);
		r.validBits.assign(btlbf::bitBytes(n), 0);
		btlbf::check(btlbf_insert_and_check_seqs(m_f, b.bases.data(), b.offsets.data(), b.size(), r.hitBits.data(),
		                                         r.validBits.data(), &r.nKmers),
		             "insertAndCheckSeqs");
		for (uint8_t v : r.hitBits)
			r.nHits += (uint64_t)__builtin_popcount(v);
		return r;
	}
	// Spaced seeds (stHashIterator): strings of kmerSize chars, '1' = care; hashNum == seeds.size()*h2.
	void setSeeds(const std::vector<std::string>& seeds, unsigned h2 = 1)
	{
		std::vector<const char*> p;
		for (const auto& s : seeds)
			p.push_back(s.c_str());
		btlbf::check(btlbf_filter_set_seeds(m_f, p.data(), (unsigned)p.size(), h2), "setSeeds");
	}

	// ---- file layout
	void writeHeader(std::ostream& out) const // :264-288
	{
		char buf[1024];
		size_t len = 0;
		btlbf::check(btlbf_format_header(BTLBF_BLOOM, m_size, m_sizeInBytes, m_hashNum, m_kmerSize, m_dFPR, m_nEntry,
		                                 m_tEntry, buf, sizeof buf, &len),
		             "writeHeader");
		out.write(buf, (std::streamsize)len);
	}
	friend std::ostream& operator<<(std::ostream& out, const BloomFilter& bloom) // :291-297
	{
		bloom.writeHeader(out);
		std::vector<uint8_t> host(bloom.m_sizeInBytes);
		btlbf::check(btlbf_filter_download(bloom.m_f, host.data(), host.size()), "reading the filter back");
		out.write(reinterpret_cast<const char*>(host.data()), (std::streamsize)host.size());
		return out;
	}
	void storeFilter(const std::string& filterFilePath) const // :304-314
	{
		std::cerr << "Writing a " << m_sizeInBytes << " byte filter to " << filterFilePath << " on disk.\n";
		btlbf::check(btlbf_filter_store(m_f, filterFilePath.c_str(), m_dFPR, m_nEntry, m_tEntry), filterFilePath.c_str());
	}

	// ---- statistics
	uint64_t getPop() const // :316-323
	{
		uint64_t n = 0;
		btlbf::check(btlbf_filter_popcount(m_f, &n), "getPop");
		return n;
	}
	unsigned getHashNum() const { return m_hashNum; }
	unsigned getKmerSize() const { return m_kmerSize; }
	double getRedudancyFPR() // :333-341
	{
		assert(m_nEntry > 0);
		double total = log(calcFPR_numInserted(1));
		for (uint64_t i = 2; i < m_nEntry; ++i)
			total = log(exp(total) + calcFPR_numInserted(i));
		return exp(total) / m_nEntry;
	}
	double getFPR() // :346-350
	{
		m_FPR = pow(double(getPop()) / double(m_size), double(m_hashNum));
		return m_FPR;
	}
	double getFPRPrecompute() const { return m_FPR; }
	double getFPR_numEle() const // :363-367
	{
		assert(m_nEntry > 0);
		return calcFPR_numInserted(m_nEntry);
	}
	uint64_t getnEntry() { return m_nEntry; }
	uint64_t gettEntry() { return m_tEntry; }
	void setnEntry(uint64_t value) { m_nEntry = value; }
	void settEntry(uint64_t value) { m_tEntry = value; }
	uint64_t getFilterSize() const { return m_size; }
	uint64_t sizeInBytes() const { return m_sizeInBytes; }

	btlbf_filter* handle() const { return m_f; } // for the C ABI's device-resident calls

  protected:
	void create(size_t filterSize, unsigned hashNum, unsigned kmerSize)
	{
		if (filterSize % 8 != 0) { // initSize, :389-394
			std::cerr << "ERROR: Filter Size \"" << filterSize << "\" is not a multiple of 8." << std::endl;
			exit(1);
		}
		btlbf::check(btlbf_filter_create(m_ctx, BTLBF_BLOOM, filterSize, hashNum, kmerSize, 0, &m_f), "allocating the filter");
		refreshInfo();
	}
	void refreshInfo()
	{
		uint64_t size = 0, bytes = 0;
		btlbf::check(btlbf_filter_info(m_f, nullptr, &size, &bytes, &m_hashNum, &m_kmerSize, nullptr), "filter info");
		m_size = size;
		m_sizeInBytes = bytes;
	}
	const uint64_t* checked(std::vector<uint64_t> const& v) const
	{
		if (v.size() < m_hashNum)
			(void)v.at(m_hashNum - 1); // throws std::out_of_range like the reference's .at(i)
		return v.data();
	}
	// :406-413 (multiple of 64)
	static size_t calcOptimalSize(size_t entries, double fpr, unsigned hashNum)
	{
		size_t non64ApproxVal = size_t(-double(entries) * double(hashNum) / log(1.0 - pow(fpr, double(1 / double(hashNum)))));
		return non64ApproxVal + (64 - non64ApproxVal % 64);
	}
	static unsigned calcOptiHashNum(double fpr) { return unsigned(-log(fpr) / log(2)); } // :419
	double calcFPR_numInserted(size_t numEntr) const                                    // :425-429
	{
		return pow(1.0 - pow(1.0 - 1.0 / double(m_size), double(numEntr) * m_hashNum), double(m_hashNum));
	}

	btlbf_ctx* m_ctx = nullptr;
	btlbf_filter* m_f = nullptr;
	size_t m_size = 0;
	size_t m_sizeInBytes = 0;
	unsigned m_hashNum = 0;
	unsigned m_kmerSize = 0;
	double m_dFPR = 0;
	uint64_t m_nEntry = 0;
	uint64_t m_tEntry = 0;
	double m_FPR = 0;
};

#endif
