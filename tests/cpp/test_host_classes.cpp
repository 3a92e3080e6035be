// test_host_classes.cpp -- the reference's unit scenarios (Tests/Unit/BloomFilterTests.cpp:53-146,
// Tests/Unit/CountingBloomFilterTests.cpp:54-246) and README usage, written against the GPU-backed
// drop-in classes in include/btlbf/.  Hash values come from the test oracle (plain-C restatement),
// which plays the role of the reference's ntHashIterator here.  Run by tests/test_cpp_host.py on a GPU.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <iterator>
#include <sstream>
#include <string>
#include <vector>

#include "btlbf/BloomFilter.hpp"
#include "btlbf/BloomFilterUtil.h"
#include "btlbf/CountingBloomFilter.hpp"
#include "btlbf/KmerBloomFilter.hpp"

extern "C" {
#include "../../oracle/btl_oracle.h"
}

#define CHECK(cond)                                                                    \
	do {                                                                               \
		if (!(cond)) {                                                                 \
			fprintf(stderr, "CHECK failed at %s:%d: %s\n", __FILE__, __LINE__, #cond); \
			exit(1);                                                                   \
		}                                                                              \
	} while (0)

static size_t
fileSize(const std::string& p)
{
	std::ifstream f(p, std::ios::binary | std::ios::ate);
	return (size_t)f.tellg();
}

static void
bloomScenario(const std::string& tmp)
{
	const size_t filterSize = 1000000000;
	const unsigned numHashes = 5, k = 4;
	const std::string seq = "ACGTAC";
	BloomFilter filter(filterSize, numHashes, k);
	ora_nt_iter it;
	for (ora_nt_iter_init(&it, seq.data(), seq.size(), numHashes, k); it.pos != ORA_END; ora_nt_iter_next(&it))
		filter.insert(it.hv);
	for (ora_nt_iter_init(&it, seq.data(), seq.size(), numHashes, k); it.pos != ORA_END; ora_nt_iter_next(&it))
		CHECK(filter.contains(it.hv));
	std::vector<uint64_t> v(it.hv, it.hv + numHashes);
	ora_nt_iter_init(&it, seq.data(), seq.size(), numHashes, k);
	std::vector<uint64_t> first(it.hv, it.hv + numHashes);
	CHECK(filter.contains(first));
	CHECK(filter.insertAndCheck(first));
	const std::string path = tmp + "/unit.bf";
	filter.storeFilter(path);
	// "[HeaderEnd]" present and the body is exactly sizeInBytes
	std::ifstream in(path, std::ios::binary);
	std::string line;
	bool end = false;
	size_t headerBytes = 0;
	while (std::getline(in, line)) {
		headerBytes += line.size() + 1;
		if (line == "[HeaderEnd]") {
			end = true;
			break;
		}
	}
	CHECK(end);
	CHECK(fileSize(path) - headerBytes == filter.sizeInBytes());
	CHECK(filter.sizeInBytes() == filterSize / 8);
	BloomFilter filter2(path);
	CHECK(filter2.getFilterSize() == filterSize && filter2.getHashNum() == numHashes && filter2.getKmerSize() == k);
	for (ora_nt_iter_init(&it, seq.data(), seq.size(), numHashes, k); it.pos != ORA_END; ora_nt_iter_next(&it))
		CHECK(filter2.contains(it.hv));
	CHECK(filter2.getPop() == filter.getPop());
	// operator<< writes the same bytes as storeFilter
	BloomFilter small(1024, 3, 5);
	insertSeq(small, "TAGAATCACCCAAAGA", 3, 5);
	std::ostringstream os;
	os << small;
	small.storeFilter(tmp + "/small.bf");
	std::ifstream sf(tmp + "/small.bf", std::ios::binary);
	std::stringstream ss;
	ss << sf.rdbuf();
	CHECK(os.str() == ss.str());
	// batched path == per-k-mer path
	BloomFilter a(8 * 1237, 4, 7), b(8 * 1237, 4, 7);
	std::vector<std::string> seqs = { "TAGAATCACCCAAAGANNTAGGACCA", "", "acgtagctagcattgGGATCGATTTAGC", "ACG" };
	uint64_t n = a.insertSeqs(seqs);
	uint64_t m = 0;
	for (const auto& s : seqs)
		for (ora_nt_iter_init(&it, s.data(), s.size(), 4, 7); it.pos != ORA_END; ora_nt_iter_next(&it)) {
			b.insert(it.hv);
			m++;
		}
	CHECK(n == m);
	std::ostringstream oa, ob;
	oa << a;
	ob << b;
	CHECK(oa.str() == ob.str());
	{
		// the same sequences through a FASTA file (multi-line records): insertFile == insertSeqs
		std::ofstream fa(tmp + "/seqs.fa");
		for (size_t i = 0; i < seqs.size(); i++) {
			fa << ">s" << i << " record\n";
			for (size_t j = 0; j < seqs[i].size(); j += 10)
				fa << seqs[i].substr(j, 10) << "\n";
		}
		fa.close();
		BloomFilter c(8 * 1237, 4, 7);
		uint64_t nSeqs = 0;
		CHECK(c.insertFile(tmp + "/seqs.fa", 0, &nSeqs) == n);
		CHECK(nSeqs == seqs.size());
		std::ostringstream oc;
		oc << c;
		CHECK(oc.str() == oa.str());
		uint64_t nk = 0;
		CHECK(c.queryFile(tmp + "/seqs.fa", &nk) == n && nk == n);
	}
	btlbf::SeqHits h = a.containsSeqs(seqs);
	CHECK(h.nKmers == n && h.nHits == n);
	btlbf::SeqBatch batch(seqs);
	for (size_t s = 0; s < seqs.size(); s++)
		for (ora_nt_iter_init(&it, seqs[s].data(), seqs[s].size(), 4, 7); it.pos != ORA_END; ora_nt_iter_next(&it))
			CHECK(h.valid(batch.offsets[s] + it.pos) && h.hit(batch.offsets[s] + it.pos));
	{
		// the same batch as 2 bits per base: same filter bytes, same per-window results
		btlbf::PackedSeqBatch pk(batch);
		CHECK(!pk.invalid.empty() && pk.codes.size() == (batch.bases.size() + 3) / 4);
		BloomFilter p(8 * 1237, 4, 7);
		CHECK(p.insertSeqs(pk) == n);
		std::ostringstream op;
		op << p;
		CHECK(op.str() == oa.str());
		btlbf::SeqHits hp = p.containsSeqs(pk);
		CHECK(hp.nKmers == n && hp.nHits == n && hp.hitBits == h.hitBits && hp.validBits == h.validBits);
	}
	// KmerBloomFilter: text k-mers, consistent with the iterator path (swig/test.pl scenario: 20-mers)
	KmerBloomFilter kb(8 * 4099, 4, 21), kb2(8 * 4099, 4, 21);
	const char* kmers[4] = { "ATCGGGTCATCAACCAATATA", "ATCGGGTCATCAACCAATATT", "ATCGGGTCATCAACCAATAAA", "ATCGGGTCATCAACCAATAGG" };
	for (int i = 0; i < 4; i++) {
		kb.insert(kmers[i]);
		ora_nt_iter_init(&it, kmers[i], 21, 4, 21);
		kb2.insert(it.hv);
	}
	std::ostringstream ka, kb_;
	ka << kb;
	kb_ << kb2;
	CHECK(ka.str() == kb_.str());
	for (int i = 0; i < 4; i++)
		CHECK(kb.contains(kmers[i]));
	CHECK(!kb.contains("ATCGGGTCATCAACCAATACC"));
	insertSeq(kb, "ATCGGGTCATCAACCAATACCGG", 4, 21);
	CHECK(kb.contains("ATCGGGTCATCAACCAATACC"));
	printf("bloom scenario ok (%llu k-mers)\n", (unsigned long long)n);
}

static void
countingScenario(const std::string& tmp)
{
	const size_t expectedSize = 100001;
	const unsigned numHashes = 5, k = 8, threshold = 1;
	const std::string seq = "ACGTACACTGGACTGAGTCT";
	CountingBloomFilter<uint8_t> filter(expectedSize, numHashes, k, threshold);
	CHECK(filter.sizeInBytes() == 100008 && filter.size() == filter.sizeInBytes());
	ora_nt_iter it;
	for (ora_nt_iter_init(&it, seq.data(), seq.size(), numHashes, k); it.pos != ORA_END; ora_nt_iter_next(&it))
		filter.insert(it.hv);
	for (ora_nt_iter_init(&it, seq.data(), seq.size(), numHashes, k); it.pos != ORA_END; ora_nt_iter_next(&it)) {
		CHECK(filter.contains(it.hv));
		CHECK(filter.minCount(it.hv) >= 1);
	}
	// a sequence that was not inserted is absent (fixed instead of srand(time(0)))
	const std::string other = "GGCATTAGCCGATATTTCAGGCAATCGGCTAAATTTCCGGAATCGCGCTATAAGCTTTCAG";
	for (ora_nt_iter_init(&it, other.data(), other.size(), numHashes, k); it.pos != ORA_END; ora_nt_iter_next(&it))
		CHECK(!filter.contains(it.hv));
	const std::string path = tmp + "/unit.cbf";
	filter.storeFilter(path);
	CountingBloomFilter<uint8_t> filter2(path, threshold);
	CHECK(filter2.size() == filter.size() && filter2.sizeInBytes() == filter.sizeInBytes());
	CHECK(filter2.getHashNum() == numHashes && filter2.getKmerSize() == k);
	CHECK(filter2.popCount() == filter.popCount());
	for (ora_nt_iter_init(&it, seq.data(), seq.size(), numHashes, k); it.pos != ORA_END; ora_nt_iter_next(&it))
		CHECK(filter2.contains(it.hv));
	// batched insert == per-k-mer insert, including repeats (order-dependent updates)
	CountingBloomFilter<uint8_t> a(512, 4, 11, 2), b(512, 4, 11, 2);
	std::vector<std::string> seqs = { std::string(200, 'A'), "ACACACACACACACACACACACACACAC", "ACGTTGCATGCATGCCGATGCATGCAGT",
		                              "ACGTTGCATGCATGCCGATGCATGCAGT" };
	uint64_t n = a.insertSeqs(seqs);
	uint64_t m = 0;
	for (const auto& s : seqs)
		for (ora_nt_iter_init(&it, s.data(), s.size(), 4, 11); it.pos != ORA_END; ora_nt_iter_next(&it)) {
			b.insert(it.hv);
			m++;
		}
	CHECK(n == m);
	std::ostringstream oa, ob;
	oa << a;
	ob << b;
	CHECK(oa.str() == ob.str());
	CHECK(a.filtered_popcount() == b.filtered_popcount());
	// saturation: 300 increments stop at 255
	ora_nt_iter_init(&it, seq.data(), seq.size(), numHashes, k);
	for (int i = 0; i < 300; i++)
		filter.incrementAll(it.hv);
	CHECK(filter.minCount(it.hv) == 255);
	CHECK(filter.insertAndCheck(it.hv));
	CHECK(filter.minCount(it.hv) == 255);
	printf("counting scenario ok (%llu k-mers)\n", (unsigned long long)n);
}

// BloomFilter::loadHeader(std::istream&) (BloomFilter.hpp:118-166): members from the header, a zeroed filter of
// that size, the stream left at the first byte of the raw array.
static void
loadHeaderScenario(const std::string& tmp)
{
	BloomFilter a(8 * 5003, 3, 9);
	insertSeq(a, "TAGAATCACCCAAAGATTTACCAGGATACCA", 3, 9);
	a.setnEntry(23);
	a.settEntry(24);
	a.storeFilter(tmp + "/hdr.bf");
	std::ifstream in(tmp + "/hdr.bf", std::ios::binary);
	BloomFilter b;
	b.loadHeader(in);
	CHECK(b.getFilterSize() == a.getFilterSize() && b.sizeInBytes() == a.sizeInBytes());
	CHECK(b.getHashNum() == 3 && b.getKmerSize() == 9 && b.getnEntry() == 23 && b.gettEntry() == 24);
	CHECK(b.getPop() == 0);
	std::vector<char> body((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
	CHECK(body.size() == a.sizeInBytes());
	std::ostringstream oa;
	oa << a;
	CHECK(oa.str().size() > body.size() && oa.str().compare(oa.str().size() - body.size(), body.size(), body.data(), body.size()) == 0);
	printf("loadHeader scenario ok\n");
}

// The reference's parallel build (Tests/AdHoc/ParallelFilter.cpp:104-122): OpenMP threads share one filter and call
// the per-k-mer insert (README.md:30-43) concurrently.  Same bytes as the batched build of the same sequence.
static void
threadedLoopScenario()
{
	const unsigned k = 25, h = 4;
	const size_t bits = 1 << 23, n = 1000000, piece = 65536;
	std::string genome(n, 'A');
	ora_synth_genome(&genome[0], 0, n, 42);
	BloomFilter shared(bits, h, k), batched(bits, h, k);
	std::vector<std::string> whole(1, genome);
	const uint64_t expect = batched.insertSeqs(whole);
	uint64_t total = 0;
	auto t0 = std::chrono::steady_clock::now();
	const long pieces = (long)((n + piece - 1) / piece);
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
	for (long p = 0; p < pieces; p++) {
		size_t lo = (size_t)p * piece, hi = lo + piece + k - 1;
		if (hi > n)
			hi = n;
		ora_nt_iter it;
		for (ora_nt_iter_init(&it, genome.data() + lo, hi - lo, h, k); it.pos != ORA_END; ora_nt_iter_next(&it)) {
			shared.insert(it.hv);
			total++;
		}
	}
	const uint64_t pop = shared.getPop(); // applies whatever is still queued
	const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	CHECK(total == expect);
	CHECK(pop == batched.getPop());
	std::ostringstream oa, ob;
	oa << shared;
	ob << batched;
	CHECK(oa.str() == ob.str());
	// the query twin (README.md:46-57), one k-mer per call (a device round trip each): a sample of the k-mers
	ora_nt_iter it;
	size_t seen = 0;
	for (ora_nt_iter_init(&it, genome.data(), 3000, h, k); it.pos != ORA_END; ora_nt_iter_next(&it), seen++)
		CHECK(shared.contains(it.hv));
	CHECK(seen == 3000 - k + 1);
	printf("threaded per-k-mer insert loop ok: %llu k-mers, %.2f Mk-mer/s\n", (unsigned long long)total, total / sec / 1e6);
	CHECK(total / sec > 1e6);
}

int
main(int argc, char** argv)
{
	std::string tmp = argc > 1 ? argv[1] : "/tmp";
	bloomScenario(tmp);
	countingScenario(tmp);
	loadHeaderScenario(tmp);
	threadedLoopScenario();
	printf("ALL OK\n");
	return 0;
}
