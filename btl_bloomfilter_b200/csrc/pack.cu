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
#if defined(__x86_64__) && !defined(__CUDA_ARCH__)
#include <immintrin.h>
#define BTL_PACK_AVX2 1
#endif

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

// bases [lo, hi) with lo a multiple of 8 (whole output bytes of both planes belong to one worker), byte by byte
void pack_range_scalar(const uint8_t* bases, uint64_t lo, uint64_t hi, uint8_t* codes, uint8_t* invalid, uint64_t* n_invalid,
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
	*n_invalid += bad;
	if (raw < *first_raw)
		*first_raw = raw;
}

#if defined(BTL_PACK_AVX2)
// 32 bases per step.  Case is folded with & 0xDF; the code comes from a 16-entry shuffle on the low nibble
// (A 0x41 -> 0, C 0x43 -> 1, G 0x47 -> 2, T 0x54 / U 0x55 -> 3), validity from five byte compares; the 2-bit codes of four
// neighbouring bytes are folded into one byte with two multiply-adds.  Blocks that hold a byte below 8 (the raw values
// 1 3 4 5 7 must be reported, not packed) go through the scalar loop.
__attribute__((target("avx2"))) void pack_range_avx2(const uint8_t* bases, uint64_t lo, uint64_t hi, uint8_t* codes, uint8_t* invalid,
                                                      uint64_t* n_invalid, uint64_t* first_raw)
{
	const __m256i fold = _mm256_set1_epi8((char)0xDF), nib = _mm256_set1_epi8(0x0F);
	const __m256i lut = _mm256_setr_epi8(0, 0, 0, 1, 3, 3, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 3, 3, 0, 2, 0, 0, 0, 0, 0, 0, 0, 0);
	const __m256i cA = _mm256_set1_epi8('A'), cC = _mm256_set1_epi8('C'), cG = _mm256_set1_epi8('G'), cT = _mm256_set1_epi8('T'),
	              cU = _mm256_set1_epi8('U'), eight = _mm256_set1_epi8(8);
	const __m256i m1 = _mm256_set1_epi16(0x0401), m2 = _mm256_set1_epi32(0x00100001);
	const __m256i pick = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1,
	                                      -1, -1, -1, -1, -1);
	uint64_t bad = 0;
	uint64_t i = lo;
	for (; i + 32 <= hi; i += 32) {
		const __m256i x = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(bases + i));
		// any byte in 0..7?  (unsigned x < 8  <=>  min(x, 8) != 8 ... use max: max(x, 8) == 8 means x <= 8; test x < 8 via subs)
		const __m256i small = _mm256_cmpeq_epi8(_mm256_subs_epu8(eight, x), _mm256_setzero_si256()); // 0xFF where x >= 8
		if (_mm256_movemask_epi8(small) != -1) {
			pack_range_scalar(bases, i, i + 32, codes, invalid, &bad, first_raw);
			continue;
		}
		const __m256i up = _mm256_and_si256(x, fold);
		__m256i ok = _mm256_or_si256(_mm256_cmpeq_epi8(up, cA), _mm256_cmpeq_epi8(up, cC));
		ok = _mm256_or_si256(ok, _mm256_or_si256(_mm256_cmpeq_epi8(up, cG), _mm256_cmpeq_epi8(up, cT)));
		ok = _mm256_or_si256(ok, _mm256_cmpeq_epi8(up, cU));
		const uint32_t inv = ~(uint32_t)_mm256_movemask_epi8(ok);
		__m256i c = _mm256_shuffle_epi8(lut, _mm256_and_si256(up, nib));
		c = _mm256_and_si256(c, ok); // an invalid base carries code 0
		const __m256i p16 = _mm256_maddubs_epi16(c, m1);  // c0 + 4 c1 per 16-bit lane
		const __m256i p32 = _mm256_madd_epi16(p16, m2);   // (c0 + 4 c1) + 16 (c2 + 4 c3) per 32-bit lane: one packed byte
		const __m256i by = _mm256_shuffle_epi8(p32, pick); // the four packed bytes of each 128-bit half, gathered
		const uint32_t lo4 = (uint32_t)_mm256_extract_epi32(by, 0), hi4 = (uint32_t)_mm256_extract_epi32(by, 4);
		memcpy(codes + (i >> 2), &lo4, 4);
		memcpy(codes + (i >> 2) + 4, &hi4, 4);
		if (invalid)
			memcpy(invalid + (i >> 3), &inv, 4);
		bad += (uint64_t)__builtin_popcount(inv);
	}
	*n_invalid += bad;
	if (i < hi)
		pack_range_scalar(bases, i, hi, codes, invalid, n_invalid, first_raw);
}
#endif

void pack_range(const uint8_t* bases, uint64_t lo, uint64_t hi, uint8_t* codes, uint8_t* invalid, uint64_t* n_invalid,
                uint64_t* first_raw)
{
	*n_invalid = 0;
	*first_raw = ~0ull;
#if defined(BTL_PACK_AVX2)
	static const bool avx2 = __builtin_cpu_supports("avx2");
	if (avx2) {
		pack_range_avx2(bases, lo, hi, codes, invalid, n_invalid, first_raw);
		return;
	}
#endif
	pack_range_scalar(bases, lo, hi, codes, invalid, n_invalid, first_raw);
}

} // namespace

// The packer without the error convention: n_invalid = bytes that are not bases, first_raw = position of the first raw
// byte 1 3 4 5 7 (which has no packed form) or ~0.  Used by btlbf_pack_seqs and by capi.cu's host-side packing of the
// ASCII entry points (context option "host_pack").
void btl_pack_bases(const char* bases, uint64_t n_bases, uint8_t* codes, uint8_t* invalid, int threads, uint64_t* n_invalid,
                    uint64_t* first_raw)
{
	*n_invalid = 0;
	*first_raw = ~0ull;
	if (n_bases == 0)
		return;
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
	for (unsigned t = 0; t < nt; t++) {
		*n_invalid += bad[t];
		if (raw[t] < *first_raw)
			*first_raw = raw[t];
	}
}

extern "C" int btlbf_pack_seqs(const char* bases, uint64_t n_bases, uint8_t* codes, uint8_t* invalid, int threads,
                               uint64_t* n_invalid)
{
	if (n_invalid)
		*n_invalid = 0;
	if (n_bases == 0)
		return BTLBF_OK;
	if (!bases || !codes)
		return btlbf_set_error(BTLBF_ERR_ARG, "null argument");
	uint64_t total = 0, first = ~0ull;
	btl_pack_bases(bases, n_bases, codes, invalid, threads, &total, &first);
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
