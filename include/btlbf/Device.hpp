// btlbf/Device.hpp -- shared plumbing of the C++ host classes: one btlbf_ctx per GPU, error convention,
// flat batches.  Header-only, over the C ABI of include/btlbf.h (link with -lbtlbf_cuda).
#ifndef BTLBF_DEVICE_HPP
#define BTLBF_DEVICE_HPP

#include <cstdint>
#include <cstdlib>
#include <iostream>
#include <mutex>
#include <string>
#include <vector>

#include "../btlbf.h"

namespace btlbf {

// The reference reports errors by printing to cerr and calling exit (BloomFilter.hpp:124-129,391-394;
// vendor/IOUtil.h:14-22).  The C ABI never exits; this layer keeps the reference's convention.
inline void
die(const char* what)
{
	std::cerr << "ERROR: " << what << ": " << btlbf_last_error() << std::endl;
	exit(EXIT_FAILURE);
}

inline void
check(int rc, const char* what)
{
	if (rc != BTLBF_OK)
		die(what);
}

// process-wide default context of a device (created on first use, never destroyed before exit)
inline btlbf_ctx*
defaultContext(int device = 0)
{
	static btlbf_ctx* ctxs[64] = { nullptr };
	static std::mutex mu; // filters may be constructed from several host threads
	std::lock_guard<std::mutex> lock(mu);
	if (device < 0 || device >= 64) {
		std::cerr << "ERROR: bad device index " << device << std::endl;
		exit(EXIT_FAILURE);
	}
	if (!ctxs[device])
		check(btlbf_ctx_create(device, &ctxs[device]), "creating the GPU context");
	return ctxs[device];
}

// Flat batch of sequences: all bases concatenated + n+1 offsets (the layout the C ABI takes).
struct SeqBatch
{
	std::string bases;
	std::vector<uint64_t> offsets;

	SeqBatch()
	  : offsets(1, 0)
	{}
	explicit SeqBatch(const std::vector<std::string>& seqs)
	  : offsets(1, 0)
	{
		size_t total = 0;
		for (const auto& s : seqs)
			total += s.size();
		bases.reserve(total);
		offsets.reserve(seqs.size() + 1);
		for (const auto& s : seqs)
			add(s);
	}
	void add(const std::string& s)
	{
		bases.append(s);
		offsets.push_back(bases.size());
	}
	uint64_t size() const { return offsets.size() - 1; }
};

// The same batch as 2 bits per base (include/btlbf.h, "2-bit packed input"): codes A0 C1 G2 T3 in the order of
// vendor/nthash.hpp:51, four bases per byte, plus one invalid bit per base (empty when every base is one of
// ACGTUacgtu).  A quarter to three eighths of the bytes that cross PCIe; the results are those of the ASCII batch.
struct PackedSeqBatch
{
	std::vector<uint8_t> codes, invalid;
	std::vector<uint64_t> offsets;

	PackedSeqBatch()
	  : offsets(1, 0)
	{}
	explicit PackedSeqBatch(const SeqBatch& b, int threads = 0)
	  : codes((b.bases.size() + 3) / 4)
	  , invalid((b.bases.size() + 7) / 8)
	  , offsets(b.offsets)
	{
		uint64_t bad = 0;
		check(btlbf_pack_seqs(b.bases.data(), b.bases.size(), codes.data(), invalid.data(), threads, &bad), "packing the batch");
		if (bad == 0)
			invalid.clear();
	}
	uint64_t size() const { return offsets.size() - 1; }
	uint64_t nBases() const { return offsets.back(); }
	const uint8_t* invalidPlane() const { return invalid.empty() ? nullptr : invalid.data(); }
};

// Per-window results of a batched query.  Window p = the k-mer starting at flat base position p
// (sequence offset + ntHashIterator::pos()).
struct SeqHits
{
	std::vector<uint8_t> hitBits;   // bit p: the k-mer is in the filter
	std::vector<uint8_t> validBits; // bit p: window p is a k-mer the reference's iterator visits
	uint64_t nKmers = 0;
	uint64_t nHits = 0;
	bool hit(uint64_t p) const { return (hitBits[p >> 3] >> (p & 7)) & 1; }
	bool valid(uint64_t p) const { return (validBits[p >> 3] >> (p & 7)) & 1; }
};

inline size_t
bitBytes(uint64_t n)
{
	return (size_t)((n + 31) / 32 * 4);
}

} // namespace btlbf
#endif
