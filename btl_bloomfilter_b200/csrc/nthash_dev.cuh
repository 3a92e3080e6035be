// nthash_dev.cuh -- ntHash arithmetic for the sm_100a kernels (host+device so the same code is
// unit-tested on the CPU against the oracle by tests/test_host_arith.py).
//
// Reference semantics (paths relative to the upstream btl_bloomfilter tree):
//   seeds / validity      vendor/nthash.hpp:189-228 (seedTab), :180 (cpOff complement trick)
//   split-rotate R, R^-1  vendor/nthash.hpp:350-352,377-380 and :360-362,383-386
//   R^n                   vendor/nthash.hpp:230-347 (msTab31l | msTab33r) == rot33(lo) | rot31(hi)
//   multi-hash mixing     vendor/nthash.hpp:183-186,684-690
//   bit / counter index   BloomFilter.hpp:188-191, CountingBloomFilter.hpp:56-58 (hash % m_size)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BTL_HD __host__ __device__ __forceinline__
#else
#define BTL_HD inline
#endif

namespace btl {

constexpr uint64_t kSeedA = 0x3c8bfbb395c60474ULL;
constexpr uint64_t kSeedC = 0x3193c18562a02b4cULL;
constexpr uint64_t kSeedG = 0x20323ed082572324ULL;
constexpr uint64_t kSeedT = 0x295549f54be24456ULL;
constexpr uint64_t kMultiSeed = 0x90b45d39fb6da1faULL;
constexpr int kMultiShift = 27;

// Base classes used on the device: bits 0-1 = 2-bit code (A0 C1 G2 T3, complement = 3-code),
// bit 2 = "self-complementary" (the raw bytes 1,3,4,5,7 whose complement seed, seedTab[c & 7],
// is their own seed), bit 3 = invalid (seedTab[c] == seedN).
constexpr uint8_t kClsSelf = 4;
constexpr uint8_t kClsBad = 8;

BTL_HD uint8_t base_class(unsigned c)
{
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': case 'U': case 'u': return 3;
	case 4: case 5: return 0 | kClsSelf;
	case 7: return 1 | kClsSelf;
	case 3: return 2 | kClsSelf;
	case 1: return 3 | kClsSelf;
	default: return kClsBad;
	}
}

BTL_HD uint64_t code_seed(unsigned code)
{
	return code == 0 ? kSeedA : code == 1 ? kSeedC : code == 2 ? kSeedG : kSeedT;
}
// forward-strand seed of a class (0 for invalid)
BTL_HD uint64_t class_fseed(unsigned cls)
{
	return (cls & kClsBad) ? 0 : code_seed(cls & 3);
}
// reverse-complement-strand seed of a class: seedTab[c & cpOff]
BTL_HD uint64_t class_rseed(unsigned cls)
{
	return (cls & kClsBad) ? 0 : code_seed((cls & kClsSelf) ? (cls & 3) : 3 - (cls & 3));
}

// R: rotate the low 33 bits and the high 31 bits left by one, independently.
BTL_HD uint64_t srol(uint64_t v)
{
	uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
	uint32_t nlo = (lo << 1) | (hi & 1u);
	uint32_t nhi = ((hi << 1) & ~3u) | ((hi >> 30) & 2u) | (lo >> 31);
	return ((uint64_t)nhi << 32) | nlo;
}

// R^-1
BTL_HD uint64_t sror(uint64_t v)
{
	uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
	uint32_t nlo = (lo >> 1) | (hi << 31);
	uint32_t nhi = ((hi >> 1) & 0x7ffffffeu) | (lo & 1u) | ((hi & 2u) << 30);
	return ((uint64_t)nhi << 32) | nlo;
}

// R^n for arbitrary n
BTL_HD uint64_t srol_n(uint64_t v, unsigned n)
{
	uint64_t lo = v & 0x1FFFFFFFFULL, hi = v >> 33;
	unsigned a = n % 33u, b = n % 31u;
	lo = ((lo << a) | (lo >> (33u - a))) & 0x1FFFFFFFFULL;
	hi = ((hi << b) | (hi >> (31u - b))) & 0x7FFFFFFFULL;
	return lo | (hi << 33);
}

// multiplier of extra hash i >= 1 for k-mer size k: (i ^ k * multiSeed)
BTL_HD uint64_t multi_mult(unsigned i, unsigned k)
{
	return (uint64_t)i ^ ((uint64_t)k * kMultiSeed);
}

BTL_HD uint64_t multi_mix(uint64_t b, uint64_t mult)
{
	uint64_t t = b * mult;
	return t ^ (t >> kMultiShift);
}

// Exact x % m for a launch-invariant m.  pow2 != 0: m is a power of two and mask == m-1.
// Otherwise magic == floor(2^64 / m) (== UINT64_MAX / m because m does not divide 2^64):
// q' = mulhi(x, magic) is floor(x/m) or floor(x/m)-1, so one conditional subtraction finishes.
struct FastMod
{
	uint64_t m;
	uint64_t magic; // or mask when pow2
	uint32_t pow2;
};

BTL_HD FastMod make_fastmod(uint64_t m)
{
	FastMod f;
	f.m = m;
	f.pow2 = (m & (m - 1)) == 0;
	f.magic = f.pow2 ? m - 1 : 0xffffffffffffffffULL / m;
	return f;
}

BTL_HD uint64_t mulhi64(uint64_t a, uint64_t b)
{
#if defined(__CUDA_ARCH__)
	return __umul64hi(a, b);
#else
	return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

template<bool POW2>
BTL_HD uint64_t fastmod(uint64_t x, const FastMod& f)
{
	if (POW2)
		return x & f.magic;
	if ((f.magic >> 32) == 0) {
		// m > 2^32 (every filter beyond 512 MiB): the reciprocal and the quotient fit in 32 bits, so the
		// 64x64 high product and q*m shrink to three 32x32->64 multiplies
		const uint32_t mg = (uint32_t)f.magic;
		uint64_t lo = (uint64_t)(uint32_t)x * mg;
		uint64_t hi = (uint64_t)(uint32_t)(x >> 32) * mg + (lo >> 32);
		uint32_t q = (uint32_t)(hi >> 32);
		uint64_t qm = (uint64_t)q * (uint32_t)f.m + ((uint64_t)(q * (uint32_t)(f.m >> 32)) << 32);
		uint64_t r = x - qm;
		return r >= f.m ? r - f.m : r;
	}
	uint64_t q = mulhi64(x, f.magic);
	uint64_t r = x - q * f.m;
	return r >= f.m ? r - f.m : r;
}

// splitmix64 and the synthetic genome; mirrored by oracle/btl_oracle.c
BTL_HD uint64_t splitmix64(uint64_t x)
{
	uint64_t z = x + 0x9e3779b97f4a7c15ULL;
	z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
	z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
	return z ^ (z >> 31);
}

BTL_HD unsigned synth_code(uint64_t i, uint64_t seed)
{
	return (unsigned)(splitmix64(seed ^ (i >> 5)) >> (2 * (i & 31))) & 3u;
}

} // namespace btl
