// pack.cu -- host-side packer in front of the 2-bit-packed entry points (host code only).
//
// Callers that keep reads as 2 bits per base (codes A0 C1 G2 T3, the order of vendor/nthash.hpp:51 convertTab) hand
// them to btlbf_*_seqs_packed as they are; callers with ASCII use btlbf_pack_seqs once (e.g. in their parser threads)
// and then move 0.25 - 0.375 bytes per base over PCIe instead of 1.  Validity follows the reference's seedTab
// (vendor/nthash.hpp:189-228): A C G T U and their lower-case forms hash, every other byte breaks the k-mers that
// contain it -- it gets code 0 and a set bit in the invalid plane.  The five raw bytes 1 3 4 5 7 that the reference
// also hashes (with their own complement rule) have no 2-bit form: the packer refuses them, the ASCII entry points
// handle them.
#include "../../include/btlbf.h"

#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

extern "C" int btlbf_set_error(int code, const char* msg); // capi.cu

namespace {

struct PackTab
{
	uint8_t t[256]; // 0..3 code, 4 invalid, 5 raw byte without a packed form
	PackTab()
	{
		memset(t, 4, sizeof t);
		t['A'] = t['a'] = 0;
		t['C'] = t['c'] = 1;
		t['G'] = t['g'] = 2;
		t['T'] = t['t'] = t['U'] = t['u'] = 3;
		t[1] = t[3] = t[4] = t[5] = t[7] = 5;
	}
};
const PackTab kTab;

// bases [lo, hi) with lo a multiple of 8 (whole output bytes of both planes belong to one worker)
void pack_range(const uint8_t* bases, uint64_t lo, uint64_t hi, uint8_t* codes, uint8_t* invalid, uint64_t* n_invalid,
                uint64_t* first_raw)
{
	uint64_t bad = 0, raw = ~0ull;
	for (uint64_t i = lo; i < hi; i += 8) {
		const uint64_t n = hi - i < 8 ? hi - i : 8;
		uint32_t c = 0, v = 0;
		for (uint64_t j = 0; j < n; j++) {
			const uint8_t x = kTab.t[bases[i + j]];
			c |= (uint32_t)(x & 3u) << (2 * j);
			if (x >= 4) {
				c &= ~(3u << (2 * j));
				v |= 1u << j;
				bad++;
				if (x == 5 && raw == ~0ull)
					raw = i + j;
			}
		}
		codes[i >> 2] = (uint8_t)c;
		if (n > 4)
			codes[(i >> 2) + 1] = (uint8_t)(c >> 8);
		if (invalid)
			invalid[i >> 3] = (uint8_t)v;
	}
	*n_invalid = bad;
	*first_raw = raw;
}

} // namespace

extern "C" int btlbf_pack_seqs(const char* bases, uint64_t n_bases, uint8_t* codes, uint8_t* invalid, int threads,
                               uint64_t* n_invalid)
{
	if (n_invalid)
		*n_invalid = 0;
	if (n_bases == 0)
		return BTLBF_OK;
	if (!bases || !codes)
		return btlbf_set_error(BTLBF_ERR_ARG, "null argument");
	unsigned nt = threads > 0 ? (unsigned)threads : std::thread::hardware_concurrency();
	if (nt < 1)
		nt = 1;
	if (nt > 64)
		nt = 64;
	if (n_bases < ((uint64_t)1 << 20))
		nt = 1;
	uint64_t per = ((n_bases + nt - 1) / nt + 7) / 8 * 8;
	std::vector<uint64_t> bad(nt, 0), raw(nt, ~0ull);
	std::vector<std::thread> pool;
	for (unsigned t = 0; t < nt; t++) {
		const uint64_t lo = (uint64_t)t * per, hi = lo + per < n_bases ? lo + per : n_bases;
		if (lo >= n_bases)
			break;
		if (nt == 1) {
			pack_range((const uint8_t*)bases, lo, hi, codes, invalid, &bad[t], &raw[t]);
			break;
		}
		pool.emplace_back(pack_range, (const uint8_t*)bases, lo, hi, codes, invalid, &bad[t], &raw[t]);
	}
	for (auto& th : pool)
		th.join();
	uint64_t total = 0, first = ~0ull;
	for (unsigned t = 0; t < nt; t++) {
		total += bad[t];
		if (raw[t] < first)
			first = raw[t];
	}
	if (n_invalid)
		*n_invalid = total;
	if (first != ~0ull) {
		char msg[160];
		snprintf(msg, sizeof msg, "byte value %u at position %llu hashes in the reference but has no 2-bit form: "
		         "use the ASCII entry points for this input", (unsigned)(uint8_t)bases[first], (unsigned long long)first);
		return btlbf_set_error(BTLBF_ERR_ARG, msg);
	}
	if (total && !invalid)
		return btlbf_set_error(BTLBF_ERR_ARG, "the input holds bytes that are not bases: an invalid plane is required");
	return BTLBF_OK;
}
