// tile_core.cuh -- the body of the fused sequence kernels, written as per-thread phase functions.
//
// One CTA (kTPB threads) processes one tile of kTile consecutive window start positions of the flat
// base stream:
//   phase A  stage the tile's ASCII bytes (+ k-1 halo) into shared memory with 16-byte loads,
//            find the first sequence boundary of the tile (binary search in the offsets array)
//   phase B  classify the bytes (seedTab semantics) in place, pack them into 2-bit codes + an
//            invalid-bit plane, mark sequence starts in a start-bit plane
//   phase C  every thread rolls the canonical ntHash over its kWPT consecutive windows: O(k) seeding
//            once, then one split-rotate step per window from register-resident 2-bit streams; each
//            valid window is extended to its h hashes and handed to the fused filter operation
//            (atomicOr bit set / word gather + test / counter gather + min / ...).
// The phases are separated by __syncthreads() in kernels.cu.  The functions are host+device and take
// the thread index as an argument so that tests/emu/ can run the very same source on the CPU
// (test infrastructure; the product only ever runs them inside the CUDA kernels).
//
// Reference semantics reproduced (upstream paths): vendor/ntHashIterator.hpp:59-86 (which windows
// are visited), vendor/nthash.hpp:667-692,581-590 (NTMC64), :820-878 (NTMSM64),
// BloomFilter.hpp:185-194,252-262, CountingBloomFilter.hpp:53-64,134-183,190-196.
#pragma once
#include "kernels.cuh"

namespace btl {

// ---------------------------------------------------------------- memory-operation shims
#if defined(__CUDA_ARCH__)
BTL_HD uint32_t mem_atomic_or(uint32_t* p, uint32_t v) { return atomicOr(p, v); }
BTL_HD void mem_red_or(uint32_t* p, uint32_t v) { atomicOr(p, v); } // result unused -> RED.OR
BTL_HD uint32_t mem_atomic_cas(uint32_t* p, uint32_t c, uint32_t v) { return atomicCAS(p, c, v); }
BTL_HD void mem_add64(uint64_t* p, uint64_t v) { atomicAdd((unsigned long long*)p, (unsigned long long)v); }
BTL_HD uint32_t mem_atomic_inc(uint32_t* p) { return atomicAdd(p, 1u); }
BTL_HD uint64_t mem_atomic_min64(uint64_t* p, uint64_t v)
{
	return atomicMin((unsigned long long*)p, (unsigned long long)v);
}
BTL_HD uint32_t smem_atomic_inc(uint32_t* p) { return atomicAdd(p, 1u); }
BTL_HD uint32_t ld_ro(const uint32_t* p) { return __ldg(p); }      // read-only filter gathers
BTL_HD uint8_t ld_ro(const uint8_t* p) { return __ldg(p); }
BTL_HD uint8_t ld_cg(const uint8_t* p) { return __ldcg(p); }        // L2-coherent counter reads
BTL_HD uint32_t ld_cg(const uint32_t* p) { return __ldcg(p); }
BTL_HD uint64_t ld_cg(const uint64_t* p) { return __ldcg((const unsigned long long*)p); }
#else
BTL_HD uint32_t mem_atomic_or(uint32_t* p, uint32_t v) { return __sync_fetch_and_or(p, v); }
BTL_HD void mem_red_or(uint32_t* p, uint32_t v) { __sync_fetch_and_or(p, v); }
BTL_HD uint32_t mem_atomic_cas(uint32_t* p, uint32_t c, uint32_t v) { return __sync_val_compare_and_swap(p, c, v); }
BTL_HD void mem_add64(uint64_t* p, uint64_t v) { __sync_fetch_and_add(p, v); }
BTL_HD uint32_t mem_atomic_inc(uint32_t* p) { return __sync_fetch_and_add(p, 1u); }
BTL_HD uint64_t mem_atomic_min64(uint64_t* p, uint64_t v)
{
	uint64_t old = *p;
	if (v < old)
		*p = v;
	return old;
}
BTL_HD uint32_t smem_atomic_inc(uint32_t* p) { return (*p)++; }
BTL_HD uint32_t ld_ro(const uint32_t* p) { return *p; }
BTL_HD uint8_t ld_ro(const uint8_t* p) { return *p; }
BTL_HD uint8_t ld_cg(const uint8_t* p) { return *(const volatile uint8_t*)p; }
BTL_HD uint32_t ld_cg(const uint32_t* p) { return *(const volatile uint32_t*)p; }
BTL_HD uint64_t ld_cg(const uint64_t* p) { return *(const volatile uint64_t*)p; }
#endif

// L2 residency hints for the ordered updates: the reservation sketch (re-read by every pass of a batch) is
// marked evict_last, the counters that stream through once per batch evict_first, so that the stream does
// not push the sketch out of the 126 MB L2.
#if defined(__CUDA_ARCH__)
BTL_HD uint64_t l2_keep()
{
	uint64_t p;
	asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}
BTL_HD uint64_t l2_stream()
{
	uint64_t p;
	asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	return p;
}
BTL_HD uint32_t ld_keep(const uint32_t* a)
{
	uint32_t v;
	asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(l2_keep()) : "memory");
	return v;
}
BTL_HD uint32_t atomic_or_keep(uint32_t* a, uint32_t v)
{
	uint32_t old;
	asm volatile("atom.global.or.L2::cache_hint.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(a), "r"(v), "l"(l2_keep()) : "memory");
	return old;
}
BTL_HD void red_or_keep(uint32_t* a, uint32_t v)
{
	asm volatile("red.global.or.L2::cache_hint.b32 [%0], %1, %2;" ::"l"(a), "r"(v), "l"(l2_keep()) : "memory");
}
BTL_HD uint32_t ld_stream(const uint8_t* a)
{
	uint32_t v;
	asm volatile("ld.global.cg.L2::cache_hint.u8 %0, [%1], %2;" : "=r"(v) : "l"(a), "l"(l2_stream()) : "memory");
	return v;
}
BTL_HD void st_stream(uint8_t* a, uint32_t v)
{
	asm volatile("st.global.L2::cache_hint.u8 [%0], %1, %2;" ::"l"(a), "r"(v), "l"(l2_stream()) : "memory");
}
#else
BTL_HD uint32_t ld_keep(const uint32_t* a) { return *(const volatile uint32_t*)a; }
BTL_HD uint32_t atomic_or_keep(uint32_t* a, uint32_t v) { return __sync_fetch_and_or(a, v); }
BTL_HD void red_or_keep(uint32_t* a, uint32_t v) { __sync_fetch_and_or(a, v); }
BTL_HD uint32_t ld_stream(const uint8_t* a) { return *(const volatile uint8_t*)a; }
BTL_HD void st_stream(uint8_t* a, uint32_t v) { *(volatile uint8_t*)a = (uint8_t)v; }
#endif

BTL_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) // (hi:lo >> sh) low word, sh in [0,31]
{
#if defined(__CUDA_ARCH__)
	return __funnelshift_r(lo, hi, sh);
#else
	return sh ? (lo >> sh) | (hi << (32 - sh)) : lo;
#endif
}

// ---------------------------------------------------------------- shared-memory carve-up
struct TileSmem
{
	uint8_t* tile;     // nb bytes: ASCII, then base classes in place
	uint32_t* codes;   // 2-bit codes, 16 bases per word
	uint32_t* badw;    // invalid-bit plane, 32 bases per word
	uint32_t* startw;  // sequence-start plane, 32 bases per word
	uint64_t* gtab;    // g_f[16] g_fk[16] g_r[16] g_rk[16]
	uint4* seedtab;    // [seed_rows(k)][4]: base i of a window with code c adds (x,y) = R^(k-2-i)(g_f[c]) to the forward hash
	                   // of the window's first k-1 bases and (z,w) = R^(i+1)(g_r[c]) to the reverse one (see seed_packed)
	uint64_t* sttab;   // spaced: TF[k][8], TR[k][8]
	uint8_t* lut;      // byte -> class
	uint64_t* scratch; // [0] first sequence index of the tile, [1] exotic flag, [2..] reductions
	uint32_t* cursors; // binned build: per-partition append cursor of this CTA (persists across its tiles)
	uint32_t writer;   // binned build: index of this CTA's private sub-buckets
	uint32_t nb;       // staged bytes (multiple of 32)
};

// rows of the seeding table: one per base of the k-1 bases that precede a thread's first rolled-in base, when those
// fit the thread's aligned 32-base register stream (k <= 33); larger k seed step by step
BTL_HD uint32_t seed_rows(uint32_t k)
{
	return k >= 2u && k <= 33u ? k - 1u : 0u;
}

BTL_HD uint32_t tile_bytes(uint32_t k, uint32_t tile = (uint32_t)kTile)
{
	return (tile + k - 1 + 31u) / 32u * 32u + 32u;
}

// tile: windows staged per CTA pass (kTile for the kTPB-thread kernels, kSortTile for the sort-bin kernel)
BTL_HD size_t tile_smem_bytes(uint32_t k, bool spaced, uint32_t nbins = 0, uint32_t tile = (uint32_t)kTile)
{
	uint32_t nb = tile_bytes(k, tile);
	size_t s = nb;                       // tile
	s += (nb / 16 + 4) * 4;              // codes
	s += (nb / 32 + 4) * 4 * 2;          // badw, startw
	s += 64 * 8;                         // gtab
	s += (size_t)seed_rows(k) * 64;      // seedtab
	s += spaced ? (size_t)k * 16 * 8 : 0; // sttab
	s += 256;                            // lut
	s += 16 * 8;                         // scratch
	s += ((size_t)nbins * 4 + 15) / 16 * 16; // cursors
	return (s + 15) / 16 * 16;
}

BTL_HD TileSmem carve_smem(uint8_t* raw, uint32_t k, bool spaced, uint32_t nbins = 0, uint32_t tile = (uint32_t)kTile)
{
	TileSmem sm;
	sm.nb = tile_bytes(k, tile);
	uint8_t* p = raw;
	sm.gtab = (uint64_t*)p;    p += 64 * 8;
	sm.seedtab = (uint4*)p;    p += (size_t)seed_rows(k) * 64;
	sm.scratch = (uint64_t*)p; p += 16 * 8;
	sm.sttab = (uint64_t*)p;   p += spaced ? (size_t)k * 16 * 8 : 0;
	sm.cursors = (uint32_t*)p; p += ((size_t)nbins * 4 + 15) / 16 * 16;
	sm.writer = 0;
	sm.tile = p;               p += sm.nb; // nb is a multiple of 32 -> keeps 16-byte alignment
	sm.codes = (uint32_t*)p;   p += (sm.nb / 16 + 4) * 4;
	sm.badw = (uint32_t*)p;    p += (sm.nb / 32 + 4) * 4;
	sm.startw = (uint32_t*)p;  p += (sm.nb / 32 + 4) * 4;
	sm.lut = p;
	return sm;
}

// ---------------------------------------------------------------- phase A: stage the tile
BTL_HD void tile_phase_a(const SeqParams& P, const TileSmem& sm, uint64_t t0, int tid, int ntid)
{
	// tables
	for (int i = tid; i < 64; i += ntid)
		sm.gtab[i] = i < 16 ? P.g_f[i] : i < 32 ? P.g_fk[i - 16] : i < 48 ? P.g_r[i - 32] : P.g_rk[i - 48];
	for (int i = tid; i < 256; i += ntid)
		sm.lut[i] = base_class((unsigned)i);
	for (uint32_t i = tid; i < seed_rows(P.k) * 4u; i += ntid) {
		const uint32_t j = i >> 2, c = i & 3u;
		const uint64_t f = srol_n(P.g_f[c], P.k - 2u - j), r = srol_n(P.g_r[c], j + 1u);
		sm.seedtab[i] = make_uint4((uint32_t)f, (uint32_t)(f >> 32), (uint32_t)r, (uint32_t)(r >> 32));
	}
	if (P.n_seeds)
		for (uint32_t i = tid; i < P.k * 16; i += ntid)
			sm.sttab[i] = P.st_tab[i];
	for (uint32_t i = tid; i < sm.nb / 32 + 4; i += ntid)
		sm.startw[i] = 0;
	if (tid == 0)
		sm.scratch[1] = 0;

	uint64_t avail = P.n_bases > t0 ? P.n_bases - t0 : 0;
	if (P.packed) {
		// 2-bit packed input: the code words and the invalid plane are the caller's, copied as they are (t0 is a
		// multiple of 32, so both start on a word; the buffers are padded to whole 16 bytes).  Bases beyond the end
		// of the chunk are marked invalid here; phase B then has nothing to classify.
		const uint32_t* csrc = reinterpret_cast<const uint32_t*>(P.bases + (t0 >> 2));
		const uint32_t* vsrc = P.invalid ? reinterpret_cast<const uint32_t*>(P.invalid + (t0 >> 3)) : nullptr;
		for (uint32_t v = tid; v < sm.nb / 16; v += ntid)
			sm.codes[v] = (uint64_t)v * 16 < avail ? csrc[v] : 0u;
		for (uint32_t g = tid; g < sm.nb / 32; g += ntid) {
			const uint64_t p = (uint64_t)g * 32;
			uint32_t bad = (vsrc && p < avail) ? vsrc[g] : 0u;
			if (p >= avail)
				bad = 0xffffffffu;
			else if (avail - p < 32)
				bad |= 0xffffffffu << (uint32_t)(avail - p);
			sm.badw[g] = bad;
		}
	}
	// ASCII bytes [t0, t0+nb) of the chunk, zero (= invalid) beyond its end
	const uint8_t* src = P.bases + t0;
	bool aligned = (((uintptr_t)src) & 15u) == 0;
	uint32_t nvec = P.packed ? 0u : sm.nb / 16;
	for (uint32_t v = tid; v < nvec; v += ntid) {
		uint64_t off = (uint64_t)v * 16;
		uint4 val;
		if (aligned && off + 16 <= avail) {
			val = *reinterpret_cast<const uint4*>(src + off);
		} else {
			uint32_t w[4] = { 0, 0, 0, 0 };
			for (int b = 0; b < 16; b++)
				if (off + b < avail)
					w[b >> 2] |= (uint32_t)src[off + b] << (8 * (b & 3));
			val.x = w[0]; val.y = w[1]; val.z = w[2]; val.w = w[3];
		}
		*reinterpret_cast<uint4*>(sm.tile + off) = val;
	}

	// first sequence offset >= flat start of the tile (lower bound over offsets[0..n_seqs])
	if (tid == 0) {
		uint64_t gstart = P.base0 + t0;
		uint64_t lo = 0, hi = P.n_seqs + 1;
		while (lo < hi) {
			uint64_t mid = (lo + hi) >> 1;
			if (P.offsets[mid] < gstart)
				lo = mid + 1;
			else
				hi = mid;
		}
		sm.scratch[0] = lo;
	}
}

// ---------------------------------------------------------------- phase B: classify + pack
BTL_HD void tile_phase_b(const SeqParams& P, const TileSmem& sm, uint64_t t0, int tid, int ntid)
{
	uint32_t ngroups = sm.nb / 32;
	uint32_t* tile32 = reinterpret_cast<uint32_t*>(sm.tile);
	bool exotic = false;
	for (uint32_t g = tid; g < (P.packed ? 0u : ngroups); g += ntid) {
		uint32_t clo = 0, chi = 0, bad = 0;
#pragma unroll
		for (int w = 0; w < 8; w++) {
			uint32_t x = tile32[g * 8 + w];
			uint32_t y = 0;
#pragma unroll
			for (int b = 0; b < 4; b++) {
				uint32_t cls = sm.lut[(x >> (8 * b)) & 255u];
				y |= cls << (8 * b);
				int idx = w * 4 + b; // base index within the group
				if (idx < 16)
					clo |= (cls & 3u) << (2 * idx);
				else
					chi |= (cls & 3u) << (2 * (idx - 16));
				bad |= ((cls >> 3) & 1u) << idx;
				exotic |= (cls & kClsSelf) != 0;
			}
			tile32[g * 8 + w] = y;
		}
		sm.codes[2 * g] = clo;
		sm.codes[2 * g + 1] = chi;
		sm.badw[g] = bad;
	}
	if (tid < 4) { // padding words read by the register streams
		sm.codes[2 * ngroups + tid] = 0;
		sm.badw[ngroups + tid] = 0xffffffffu;
	}
	if (exotic)
		sm.scratch[1] = 1;

	// sequence starts inside [gstart, gstart + nb)
	uint64_t gstart = P.base0 + t0, gend = gstart + sm.nb;
	for (uint64_t i = sm.scratch[0] + (uint64_t)tid; i <= P.n_seqs; i += (uint64_t)ntid) {
		uint64_t o = P.offsets[i];
		if (o >= gend)
			break;
		uint32_t rel = (uint32_t)(o - gstart);
#if defined(__CUDA_ARCH__)
		atomicOr(&sm.startw[rel >> 5], 1u << (rel & 31));
#else
		sm.startw[rel >> 5] |= 1u << (rel & 31);
#endif
	}
}

// ---------------------------------------------------------------- per-window hash expansion
// Calls fn(i, hash_i, strand) for the h hashes of one valid window in reference order; stops early
// when fn returns false.  w = tile-local window index (spaced seeds re-read the window's classes).
template<bool SPACED, class Fn>
BTL_HD void for_each_hash(const SeqParams& P, const TileSmem& sm, uint32_t w, uint64_t F, uint64_t RC, Fn&& fn)
{
	if (!SPACED) {
		bool st = RC < F;
		uint64_t b = st ? RC : F;
		if (!fn(0u, b, st))
			return;
		for (uint32_t i = 1; i < P.h; i++)
			if (!fn(i, multi_mix(b, P.mult[i]), st))
				return;
	} else {
		const uint64_t* TF = sm.sttab;
		const uint64_t* TR = sm.sttab + (size_t)P.k * 8;
		// The class of a don't-care base: from the 2-bit plane when the tile has no exotic bytes (lanes are 32
		// bases apart: 2-way bank conflicts there against 8-way on the byte plane -- these reads are what bounds
		// the spaced-seed kernels), else from the byte plane (bit 2 = self-complementary raw byte).
		const bool packed = !P.force_generic && sm.scratch[1] == 0;
		for (uint32_t j = 0; j < P.n_seeds; j++) {
			uint64_t fs = F, rs = RC;
			for (uint32_t t = P.st_dc_off[j]; t < P.st_dc_off[j + 1]; t++) {
				uint32_t pos = P.st_dc[t];
				uint32_t q = w + pos;
				uint32_t c = packed ? (sm.codes[q >> 4] >> (2 * (q & 15))) & 3u : sm.tile[q] & 7u;
				fs ^= TF[pos * 8 + c];
				rs ^= TR[pos * 8 + c];
			}
			bool st = rs < fs;
			uint64_t b = st ? rs : fs;
			if (!fn(j * P.h2, b, st))
				return;
			for (uint32_t j2 = 1; j2 < P.h2; j2++)
				if (!fn(j * P.h2 + j2, multi_mix(b, P.mult[j2]), st))
					return;
		}
	}
}

// all h probes of a contiguous-ntHash BloomFilter query issued before any is consumed
template<int H, bool POW2>
BTL_HD bool bf_test_unrolled(const SeqParams& P, uint64_t b)
{
	const uint32_t* words = (const uint32_t*)P.filter;
	uint32_t acc = 1;
	uint32_t bit[H];
#pragma unroll
	for (int i = 0; i < H; i++) {
		uint64_t hv = i ? multi_mix(b, P.mult[i]) : b;
		uint64_t n = fastmod<POW2>(hv, P.fm);
		bit[i] = ld_ro(words + (n >> 5)) >> (uint32_t)(n & 31);
	}
#pragma unroll
	for (int i = 0; i < H; i++)
		acc &= bit[i];
	return acc & 1u;
}

template<int H, bool POW2>
BTL_HD uint32_t cbf_min_unrolled(const SeqParams& P, uint64_t b)
{
	const uint8_t* cnt = (const uint8_t*)P.filter;
	uint32_t v[H];
#pragma unroll
	for (int i = 0; i < H; i++) {
		uint64_t hv = i ? multi_mix(b, P.mult[i]) : b;
		v[i] = ld_ro(cnt + fastmod<POW2>(hv, P.fm));
	}
	uint32_t mn = 255;
#pragma unroll
	for (int i = 0; i < H; i++)
		mn = v[i] < mn ? v[i] : mn;
	return mn;
}

// saturating ++ of one 8-bit counter by CAS on its aligned 32-bit word (CountingBloomFilter.hpp:164-183)
BTL_HD void counter_sat_inc(uint8_t* cnt, uint64_t n)
{
	uint32_t* word = reinterpret_cast<uint32_t*>(cnt + (n & ~(uint64_t)3));
	uint32_t sh = (uint32_t)(n & 3) * 8;
	uint32_t old = ld_cg(word);
	for (;;) {
		if (((old >> sh) & 255u) == 255u)
			return;
		uint32_t seen = mem_atomic_cas(word, old, old + (1u << sh));
		if (seen == old)
			return;
		old = seen;
	}
}

// exact incrementMin of one k-mer that owns all its slots (CountingBloomFilter.hpp:134-162)
// returns the minimum BEFORE the update (CountingBloomFilter::insertAndCheck, :206-214, reports
// minCount >= threshold of that value)
template<bool SPACED, bool POW2>
BTL_HD uint32_t cbf_commit_one(const SeqParams& P, const TileSmem& sm, uint32_t w, uint64_t F, uint64_t RC)
{
	uint8_t* cnt = (uint8_t*)P.filter;
	uint32_t mn = 255;
	for_each_hash<SPACED>(P, sm, w, F, RC, [&](uint32_t, uint64_t hv, bool) {
		uint32_t v = ld_cg(cnt + fastmod<POW2>(hv, P.fm));
		mn = v < mn ? v : mn;
		return true;
	});
	if (mn == 255)
		return mn;
	for_each_hash<SPACED>(P, sm, w, F, RC, [&](uint32_t, uint64_t hv, bool) {
		uint64_t n = fastmod<POW2>(hv, P.fm);
		if (ld_cg(cnt + n) == mn)
			*(volatile uint8_t*)(cnt + n) = (uint8_t)(mn + 1);
		return true;
	});
	return mn;
}

// insertAndCheck of one k-mer that owns all its bits (BloomFilter.hpp:200-214): true when every bit was set
template<bool SPACED, bool POW2>
BTL_HD bool bfchk_commit_one(const SeqParams& P, const TileSmem& sm, uint32_t w, uint64_t F, uint64_t RC)
{
	uint32_t* words = (uint32_t*)P.filter;
	bool found = true;
	for_each_hash<SPACED>(P, sm, w, F, RC, [&](uint32_t, uint64_t hv, bool) {
		uint64_t n = fastmod<POW2>(hv, P.fm);
		uint32_t bit = 1u << (uint32_t)(n & 31);
		found &= (mem_atomic_or(words + (n >> 5), bit) & bit) != 0;
		return true;
	});
	return found;
}

// reservation sketch of the ordered updates: the two table positions of a slot, counting, reading
BTL_HD uint32_t resv_pos1(uint64_t slot, uint32_t log2)
{
	return (uint32_t)(slot & (((uint64_t)1 << log2) - 1));
}
BTL_HD uint32_t resv_pos2(uint64_t slot, uint32_t log2)
{
	return (uint32_t)((slot * 0x9e3779b97f4a7c15ULL + 0x7f4a7c159e3779b9ULL) >> (64 - log2));
}
BTL_HD void resv_count(const SeqParams& P, uint32_t e)
{
	uint32_t bit = 1u << (e & 31);
	if (atomic_or_keep(P.resv_touched + (e >> 5), bit) & bit)
		red_or_keep(P.resv_contended + (e >> 5), bit);
}
BTL_HD bool resv_twice(const SeqParams& P, uint32_t e)
{
	return (ld_keep(P.resv_contended + (e >> 5)) >> (e & 31)) & 1u;
}

// ---------------------------------------------------------------- the fused per-window operation
struct ThreadOut
{
	uint32_t validw, hitw; // bit s = window p0+s
};

template<int OP, bool SPACED, bool POW2>
BTL_HD void window_op(const SeqParams& P, const TileSmem& sm, uint64_t t0, uint32_t w, uint32_t s,
                      uint64_t F, uint64_t RC, ThreadOut& out)
{
	out.validw |= 1u << s;
	if (OP == OP_HASH) {
		uint64_t gw = t0 + w;
		for_each_hash<SPACED>(P, sm, w, F, RC, [&](uint32_t i, uint64_t hv, bool st) {
			if (P.hashes)
				P.hashes[gw * P.h + i] = hv;
			if (P.strands)
				P.strands[gw * P.h + i] = SPACED ? (uint8_t)st : 0;
			return true;
		});
	} else if (OP == OP_BF_INSERT) {
		uint32_t* words = (uint32_t*)P.filter;
		for_each_hash<SPACED>(P, sm, w, F, RC, [&](uint32_t, uint64_t hv, bool) {
			uint64_t n = fastmod<POW2>(hv, P.fm);
			mem_red_or(words + (n >> 5), 1u << (uint32_t)(n & 31));
			return true;
		});
	} else if (OP == OP_BF_CONTAINS) {
		bool hit;
		bool done = false;
		if (!SPACED && P.query_mode == 0) {
			uint64_t b = RC < F ? RC : F;
			done = true;
			switch (P.h) {
			case 1: hit = bf_test_unrolled<1, POW2>(P, b); break;
			case 2: hit = bf_test_unrolled<2, POW2>(P, b); break;
			case 3: hit = bf_test_unrolled<3, POW2>(P, b); break;
			case 4: hit = bf_test_unrolled<4, POW2>(P, b); break;
			case 5: hit = bf_test_unrolled<5, POW2>(P, b); break;
			case 6: hit = bf_test_unrolled<6, POW2>(P, b); break;
			case 7: hit = bf_test_unrolled<7, POW2>(P, b); break;
			case 8: hit = bf_test_unrolled<8, POW2>(P, b); break;
			default: done = false; hit = false; break;
			}
		}
		if (!done) {
			const uint32_t* words = (const uint32_t*)P.filter;
			hit = true;
			for_each_hash<SPACED>(P, sm, w, F, RC, [&](uint32_t, uint64_t hv, bool) {
				uint64_t n = fastmod<POW2>(hv, P.fm);
				hit = (ld_ro(words + (n >> 5)) >> (uint32_t)(n & 31)) & 1u;
				return hit; // early exit on the first miss, BloomFilter.hpp:257-259
			});
		}
		if (hit)
			out.hitw |= 1u << s;
	} else if (OP == OP_CBF_MINCOUNT) {
		uint32_t mn = 255;
		bool done = false;
		if (!SPACED) {
			uint64_t b = RC < F ? RC : F;
			done = true;
			switch (P.h) {
			case 1: mn = cbf_min_unrolled<1, POW2>(P, b); break;
			case 2: mn = cbf_min_unrolled<2, POW2>(P, b); break;
			case 3: mn = cbf_min_unrolled<3, POW2>(P, b); break;
			case 4: mn = cbf_min_unrolled<4, POW2>(P, b); break;
			case 5: mn = cbf_min_unrolled<5, POW2>(P, b); break;
			case 6: mn = cbf_min_unrolled<6, POW2>(P, b); break;
			case 7: mn = cbf_min_unrolled<7, POW2>(P, b); break;
			case 8: mn = cbf_min_unrolled<8, POW2>(P, b); break;
			default: done = false; break;
			}
		}
		if (!done) {
			const uint8_t* cnt = (const uint8_t*)P.filter;
			for_each_hash<SPACED>(P, sm, w, F, RC, [&](uint32_t, uint64_t hv, bool) {
				uint32_t v = ld_ro(cnt + fastmod<POW2>(hv, P.fm));
				mn = v < mn ? v : mn;
				return true;
			});
		}
		if (P.counts)
			P.counts[t0 + w] = (uint8_t)mn;
		if (mn >= P.threshold) // CountingBloomFilter.hpp:190-196
			out.hitw |= 1u << s;
	} else if (OP == OP_CBF_INCALL) {
		uint8_t* cnt = (uint8_t*)P.filter;
		for_each_hash<SPACED>(P, sm, w, F, RC, [&](uint32_t, uint64_t hv, bool) {
			counter_sat_inc(cnt, fastmod<POW2>(hv, P.fm));
			return true;
		});
	} else if (OP == OP_BF_BIN) {
		// partitioned build, pass 1: the bit index is split into (filter partition, offset inside it) and
		// the offset is appended to this CTA's private sub-bucket of that partition.  A sub-bucket that is
		// full (skewed input) falls back to the direct atomic -- OR is order-free, so any mix is exact.
		uint32_t* words = (uint32_t*)P.filter;
		for_each_hash<SPACED>(P, sm, w, F, RC, [&](uint32_t, uint64_t hv, bool) {
			uint64_t n = fastmod<POW2>(hv, P.fm);
			uint32_t part = (uint32_t)(n >> P.bin_shift);
			uint32_t pos = smem_atomic_inc(sm.cursors + part);
			if (pos < P.bin_cap)
				P.bin_items[((uint64_t)part * P.bin_writers + sm.writer) * P.bin_cap + pos] = (uint32_t)n & P.bin_mask;
			else
				mem_red_or(words + (n >> 5), 1u << (uint32_t)(n & 31));
			return true;
		});
	} else if (OP == OP_RESV_TOUCH) {
		// ordered updates, pass 1: count-min sketch with two-valued counters.  Every slot is counted at two
		// positions of the bit tables (touched = "at least once", contended = "at least twice").
		for_each_hash<SPACED>(P, sm, w, F, RC, [&](uint32_t, uint64_t hv, bool) {
			uint64_t slot = fastmod<POW2>(hv, P.fm);
			resv_count(P, resv_pos1(slot, P.resv_log2));
			resv_count(P, resv_pos2(slot, P.resv_log2));
			return true;
		});
	} else if (OP == OP_CBF_COMMIT || OP == OP_BFCHK_COMMIT) {
		// pass 2: a slot that some other k-mer of the batch also uses was counted twice at BOTH of its
		// positions.  A k-mer none of whose slots looks that way shares no slot with any other k-mer of the
		// batch, so its update commutes with all of them and is applied now; the others (true sharers plus the
		// few false alarms of the sketch) are deferred to the index-ordered residual rounds (list_round_*).
		bool contended = false;
		for_each_hash<SPACED>(P, sm, w, F, RC, [&](uint32_t, uint64_t hv, bool) {
			uint64_t slot = fastmod<POW2>(hv, P.fm);
			if (resv_twice(P, resv_pos1(slot, P.resv_log2)) && resv_twice(P, resv_pos2(slot, P.resv_log2)))
				contended = true;
			return true;
		});
		if (!contended) {
			bool found;
			if (OP == OP_CBF_COMMIT)
				found = cbf_commit_one<SPACED, POW2>(P, sm, w, F, RC) >= P.threshold;
			else
				found = bfchk_commit_one<SPACED, POW2>(P, sm, w, F, RC);
			if (found)
				out.hitw |= 1u << s;
		} else {
			uint32_t slot = mem_atomic_inc(P.pending_count);
			P.pending[slot] = (uint32_t)(t0 + w);
			if (P.pending_slots) {
				uint64_t* dst = P.pending_slots + (t0 + w) * P.h;
				for_each_hash<SPACED>(P, sm, w, F, RC, [&](uint32_t i, uint64_t hv, bool) {
					dst[i] = fastmod<POW2>(hv, P.fm);
					return true;
				});
			}
		}
	} else if (OP == OP_RESV_CLEAR) {
		// pass 3 (small batches; large ones clear the tables with a memset instead)
		for_each_hash<SPACED>(P, sm, w, F, RC, [&](uint32_t, uint64_t hv, bool) {
			uint64_t slot = fastmod<POW2>(hv, P.fm);
			uint32_t e1 = resv_pos1(slot, P.resv_log2), e2 = resv_pos2(slot, P.resv_log2);
			P.resv_touched[e1 >> 5] = 0;
			P.resv_contended[e1 >> 5] = 0;
			P.resv_touched[e2 >> 5] = 0;
			P.resv_contended[e2 >> 5] = 0;
			return true;
		});
	}
}

// ---------------------------------------------------------------- seeding
// State of the rolling hash after the k-1 bases p0 .. p0+k-2 of the thread's first window (2-bit path): what the
// step-by-step loop  F = R(F) ^ g_f[c_i],  RC = R(RC) ^ g_r[c_(k-2-i)],  ... RC = R(RC)  leaves, computed as
// F = XOR_i R^(k-2-i)(g_f[c_i]),  RC = XOR_i R^(i+1)(g_r[c_i])  (R is linear over XOR) from the per-position table
// tile_phase_a builds: one 16-byte shared-memory load and four XORs per base instead of two split rotations and four
// dependent loads.  g (hashable bases ending at the newest one, restarting at sequence starts) comes from the
// invalid / start planes with one count-leading-zeros.  p0 is a multiple of 32, k <= 33: the bases sit in one aligned
// 64-bit code word pair and one word of each plane.  NTMC64 base, vendor/nthash.hpp:667-692.
BTL_HD uint32_t clz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
	return (uint32_t)__clz((int)x);
#else
	return x ? (uint32_t)__builtin_clz(x) : 32u;
#endif
}

BTL_HD void seed_packed(const TileSmem& sm, uint32_t p0, uint32_t k, uint64_t codes, uint64_t& F, uint64_t& RC, uint32_t& g)
{
	uint32_t fl = 0, fh = 0, rl = 0, rh = 0;
	const uint4* T = sm.seedtab;
	for (uint32_t i = 0; i + 1 < k; i++) {
		const uint4 e = T[i * 4u + ((uint32_t)codes & 3u)];
		codes >>= 2;
		fl ^= e.x; fh ^= e.y; rl ^= e.z; rh ^= e.w;
	}
	F = ((uint64_t)fh << 32) | fl;
	RC = ((uint64_t)rh << 32) | rl;
	const uint32_t m = k - 1u >= 32u ? 0xffffffffu : (1u << (k - 1u)) - 1u;
	const uint32_t bad = sm.badw[p0 >> 5] & m, ev = bad | (sm.startw[p0 >> 5] & m);
	if (ev == 0) {
		g = k - 1u;
	} else {
		const uint32_t t = 31u - clz32(ev); // the last base that restarts the count
		g = ((bad >> t) & 1u) ? k - 2u - t : k - 1u - t;
	}
}

// ---------------------------------------------------------------- phase C: roll + operate
// Rolls the canonical ntHash over the thread's kWPT consecutive windows and calls fn(s, ok, F, RC) for
// every s in [0, kWPT) -- for every thread, in the same order, so that fn may contain warp-synchronous
// code.  ok: window p0+s is a k-mer the reference's iterator visits (all k bases hashable, inside one
// sequence, inside the launch's window range); F / RC are only meaningful when ok.
template<class Fn>
BTL_HD void roll_windows(const SeqParams& P, const TileSmem& sm, uint64_t t0, int tid, Fn&& fn)
{
	uint64_t nwin64 = P.n_windows > t0 ? P.n_windows - t0 : 0;
	uint32_t nwin = nwin64 > (uint64_t)kTile ? (uint32_t)kTile : (uint32_t)nwin64;
	uint32_t p0 = (uint32_t)tid * kWPT;
	const uint32_t k = P.k;
	const uint64_t* Gf = sm.gtab;
	const uint64_t* Gfk = sm.gtab + 16;
	const uint64_t* Gr = sm.gtab + 32;
	const uint64_t* Grk = sm.gtab + 48;
	uint64_t F = 0, RC = 0;
	uint32_t g = 0; // hashable bases of the current sequence ending at the newest base
	uint32_t q1 = p0 + k - 1; // first incoming base
	bool generic = P.force_generic || sm.scratch[1] != 0;

	if (generic) {
		// byte-class path: exact for every byte value (self-complementary raw bytes included)
		for (uint32_t i = 0; i + 1 < k; i++) {
			uint32_t qa = p0 + i, qb = p0 + k - 2 - i;
			uint32_t ca = sm.tile[qa], cb = sm.tile[qb];
			F = srol(F) ^ Gf[ca];
			RC = srol(RC) ^ Gr[cb];
			bool st = (sm.startw[qa >> 5] >> (qa & 31)) & 1u;
			g = (ca & kClsBad) ? 0u : (st ? 1u : g + 1u);
		}
		RC = srol(RC);
		for (uint32_t s = 0; s < (uint32_t)kWPT; s++) {
			uint32_t q = q1 + s;
			uint32_t cin = sm.tile[q];
			F = srol(F) ^ Gf[cin];
			RC ^= Grk[cin];
			if (s > 0) {
				uint32_t cout = sm.tile[p0 + s - 1];
				F ^= Gfk[cout];
				RC ^= Gr[cout];
			}
			RC = sror(RC);
			bool st = (sm.startw[q >> 5] >> (q & 31)) & 1u;
			g = (cin & kClsBad) ? 0u : (st ? 1u : g + 1u);
			fn(s, g >= k && p0 + s < nwin, F, RC);
		}
		return;
	}

	// 2-bit packed path
	uint64_t out_codes = ((uint64_t)sm.codes[(p0 >> 4) + 1] << 32) | sm.codes[p0 >> 4];
	if (seed_rows(k)) {
		seed_packed(sm, p0, k, out_codes, F, RC, g);
	} else {
		for (uint32_t i = 0; i + 1 < k; i++) {
			uint32_t qa = p0 + i, qb = p0 + k - 2 - i;
			uint32_t ca = (sm.codes[qa >> 4] >> (2 * (qa & 15))) & 3u;
			uint32_t cb = (sm.codes[qb >> 4] >> (2 * (qb & 15))) & 3u;
			F = srol(F) ^ Gf[ca];
			RC = srol(RC) ^ Gr[cb];
			bool bad = (sm.badw[qa >> 5] >> (qa & 31)) & 1u;
			bool st = (sm.startw[qa >> 5] >> (qa & 31)) & 1u;
			g = bad ? 0u : (st ? 1u : g + 1u);
		}
		RC = srol(RC);
	}
	// register streams: 32 incoming bases from q1 (unaligned), 32 outgoing bases from p0 (aligned)
	uint32_t a = q1 >> 4, sh2 = 2 * (q1 & 15);
	uint32_t in_lo = funnel_r(sm.codes[a], sm.codes[a + 1], sh2);
	uint32_t in_hi = funnel_r(sm.codes[a + 1], sm.codes[a + 2], sh2);
	uint32_t bw = q1 >> 5, sh1 = q1 & 31;
	uint32_t in_bad = funnel_r(sm.badw[bw], sm.badw[bw + 1], sh1);
	uint32_t in_start = funnel_r(sm.startw[bw], sm.startw[bw + 1], sh1);
	uint64_t in_codes = ((uint64_t)in_hi << 32) | in_lo;
#pragma unroll 4
	for (uint32_t s = 0; s < (uint32_t)kWPT; s++) {
		uint32_t cin = (uint32_t)in_codes & 3u;
		in_codes >>= 2;
		F = srol(F) ^ Gf[cin];
		RC ^= Grk[cin];
		if (s > 0) {
			uint32_t cout = (uint32_t)out_codes & 3u;
			out_codes >>= 2;
			F ^= Gfk[cout];
			RC ^= Gr[cout];
		}
		RC = sror(RC);
		g = ((in_bad >> s) & 1u) ? 0u : (((in_start >> s) & 1u) ? 1u : g + 1u);
		fn(s, g >= k && p0 + s < nwin, F, RC);
	}
}

// The rolling hash of roll_windows() as an explicit state machine: init() seeds the thread's first window
// (O(k)), one step() per window after that.  Used where the windows are consumed in groups (sort_bin.cuh,
// the grouped commit pass below).
struct Roller
{
	uint64_t F, RC, in_codes, out_codes;
	uint32_t in_bad, in_start, g, p0, q1, nwin;
	bool generic;

	BTL_HD void init(const SeqParams& P, const TileSmem& sm, uint64_t t0, int tid, uint32_t tile)
	{
		uint64_t nwin64 = P.n_windows > t0 ? P.n_windows - t0 : 0;
		nwin = nwin64 > (uint64_t)tile ? tile : (uint32_t)nwin64;
		p0 = (uint32_t)tid * kWPT;
		const uint32_t k = P.k;
		const uint64_t* Gf = sm.gtab;
		const uint64_t* Gr = sm.gtab + 32;
		F = 0;
		RC = 0;
		g = 0;
		q1 = p0 + k - 1;
		generic = P.force_generic || sm.scratch[1] != 0;
		in_codes = out_codes = 0;
		in_bad = in_start = 0;
		if (generic) {
			// byte-class path: exact for every byte value (self-complementary raw bytes included)
			for (uint32_t i = 0; i + 1 < k; i++) {
				uint32_t qa = p0 + i, qb = p0 + k - 2 - i;
				uint32_t ca = sm.tile[qa], cb = sm.tile[qb];
				F = srol(F) ^ Gf[ca];
				RC = srol(RC) ^ Gr[cb];
				bool st = (sm.startw[qa >> 5] >> (qa & 31)) & 1u;
				g = (ca & kClsBad) ? 0u : (st ? 1u : g + 1u);
			}
			RC = srol(RC);
			return;
		}
		out_codes = ((uint64_t)sm.codes[(p0 >> 4) + 1] << 32) | sm.codes[p0 >> 4];
		if (seed_rows(k)) {
			seed_packed(sm, p0, k, out_codes, F, RC, g);
		} else {
			for (uint32_t i = 0; i + 1 < k; i++) {
				uint32_t qa = p0 + i, qb = p0 + k - 2 - i;
				uint32_t ca = (sm.codes[qa >> 4] >> (2 * (qa & 15))) & 3u;
				uint32_t cb = (sm.codes[qb >> 4] >> (2 * (qb & 15))) & 3u;
				F = srol(F) ^ Gf[ca];
				RC = srol(RC) ^ Gr[cb];
				bool bad = (sm.badw[qa >> 5] >> (qa & 31)) & 1u;
				bool st = (sm.startw[qa >> 5] >> (qa & 31)) & 1u;
				g = bad ? 0u : (st ? 1u : g + 1u);
			}
			RC = srol(RC);
		}
		// register streams: 32 incoming bases from q1 (unaligned), 32 outgoing bases from p0 (aligned)
		uint32_t a = q1 >> 4, sh2 = 2 * (q1 & 15);
		uint32_t in_lo = funnel_r(sm.codes[a], sm.codes[a + 1], sh2);
		uint32_t in_hi = funnel_r(sm.codes[a + 1], sm.codes[a + 2], sh2);
		uint32_t bw = q1 >> 5, sh1 = q1 & 31;
		in_bad = funnel_r(sm.badw[bw], sm.badw[bw + 1], sh1);
		in_start = funnel_r(sm.startw[bw], sm.startw[bw + 1], sh1);
		in_codes = ((uint64_t)in_hi << 32) | in_lo;
	}

	// advances to window p0+s (s = 0, 1, 2, ... in order); true when it is a k-mer the reference's
	// iterator visits (ntHashIterator.hpp:59-86)
	BTL_HD bool step(const SeqParams& P, const TileSmem& sm, uint32_t s)
	{
		const uint64_t* Gf = sm.gtab;
		const uint64_t* Gfk = sm.gtab + 16;
		const uint64_t* Gr = sm.gtab + 32;
		const uint64_t* Grk = sm.gtab + 48;
		if (generic) {
			uint32_t q = q1 + s;
			uint32_t cin = sm.tile[q];
			F = srol(F) ^ Gf[cin];
			RC ^= Grk[cin];
			if (s > 0) {
				uint32_t cout = sm.tile[p0 + s - 1];
				F ^= Gfk[cout];
				RC ^= Gr[cout];
			}
			RC = sror(RC);
			bool st = (sm.startw[q >> 5] >> (q & 31)) & 1u;
			g = (cin & kClsBad) ? 0u : (st ? 1u : g + 1u);
		} else {
			uint32_t cin = (uint32_t)in_codes & 3u;
			in_codes >>= 2;
			F = srol(F) ^ Gf[cin];
			RC ^= Grk[cin];
			if (s > 0) {
				uint32_t cout = (uint32_t)out_codes & 3u;
				out_codes >>= 2;
				F ^= Gfk[cout];
				RC ^= Gr[cout];
			}
			RC = sror(RC);
			g = ((in_bad >> s) & 1u) ? 0u : (((in_start >> s) & 1u) ? 1u : g + 1u);
		}
		return g >= P.k && p0 + s < nwin;
	}
};

// ---------------------------------------------------------------- ordered updates, pass 2, grouped
// The per-window form of OP_CBF_COMMIT / OP_BFCHK_COMMIT waits for three dependent round trips per window
// (sketch -> counters -> update).  Here a thread handles G windows at a time, so that the G*H*2 sketch reads,
// then the G*H counter reads (or bit atomics), are in flight together.  Exact for the same reason as the
// per-window form: only k-mers that share no slot with any other k-mer of the batch are touched, so the
// order in which their reads and writes are issued does not matter.  Contiguous ntHash, h <= 8.
template<int OP, int H, bool POW2>
BTL_HD ThreadOut tile_phase_c_commit(const SeqParams& P, const TileSmem& sm, uint64_t t0, int tid)
{
	constexpr int G = H <= 4 ? 4 : 2;
	ThreadOut out;
	out.validw = 0;
	out.hitw = 0;
	const uint64_t nwin64 = P.n_windows > t0 ? P.n_windows - t0 : 0;
	const uint32_t p0 = (uint32_t)tid * kWPT;
	if (p0 >= nwin64)
		return out;
	Roller r;
	r.init(P, sm, t0, tid, (uint32_t)kTile);
	for (uint32_t s0 = 0; s0 < (uint32_t)kWPT; s0 += G) {
		uint64_t slot[G][H];
		uint32_t okmask = 0;
#pragma unroll
		for (int g = 0; g < G; g++) {
			const bool ok = r.step(P, sm, s0 + g);
			okmask |= (uint32_t)ok << g;
			const uint64_t b = r.RC < r.F ? r.RC : r.F;
#pragma unroll
			for (int i = 0; i < H; i++)
				slot[g][i] = fastmod<POW2>(i ? multi_mix(b, P.mult[i]) : b, P.fm);
		}
		out.validw |= okmask << s0;
		if (!okmask)
			continue;
		// sketch: "counted twice" words of both positions of every slot
		uint32_t w1[G][H], w2[G][H];
#pragma unroll
		for (int g = 0; g < G; g++)
#pragma unroll
			for (int i = 0; i < H; i++) {
				const bool on = (okmask >> g) & 1u;
				w1[g][i] = on ? ld_keep(P.resv_contended + (resv_pos1(slot[g][i], P.resv_log2) >> 5)) : 0u;
				w2[g][i] = on ? ld_keep(P.resv_contended + (resv_pos2(slot[g][i], P.resv_log2) >> 5)) : 0u;
			}
		uint32_t freemask = 0; // windows that share no slot with any other k-mer of the batch
#pragma unroll
		for (int g = 0; g < G; g++) {
			uint32_t c = 0;
#pragma unroll
			for (int i = 0; i < H; i++)
				c |= (w1[g][i] >> (resv_pos1(slot[g][i], P.resv_log2) & 31)) &
				     (w2[g][i] >> (resv_pos2(slot[g][i], P.resv_log2) & 31)) & 1u;
			freemask |= (((okmask >> g) & 1u) & (c ^ 1u)) << g;
		}
		if (OP == OP_CBF_COMMIT) {
			uint8_t* cnt = (uint8_t*)P.filter;
			uint32_t v[G][H];
#pragma unroll
			for (int g = 0; g < G; g++)
#pragma unroll
				for (int i = 0; i < H; i++)
					v[g][i] = ((freemask >> g) & 1u) ? ld_stream(cnt + slot[g][i]) : 255u;
#pragma unroll
			for (int g = 0; g < G; g++) {
				if (!((freemask >> g) & 1u))
					continue;
				uint32_t mn = 255;
#pragma unroll
				for (int i = 0; i < H; i++)
					mn = v[g][i] < mn ? v[g][i] : mn;
				if (mn != 255) { // CountingBloomFilter.hpp:134-162
#pragma unroll
					for (int i = 0; i < H; i++)
						if (v[g][i] == mn)
							st_stream(cnt + slot[g][i], mn + 1);
				}
				if (mn >= P.threshold) // insertAndCheck reports the minimum before the update, :206-214
					out.hitw |= 1u << (s0 + g);
			}
		} else {
			uint32_t* words = (uint32_t*)P.filter;
			uint32_t old[G][H];
#pragma unroll
			for (int g = 0; g < G; g++)
#pragma unroll
				for (int i = 0; i < H; i++)
					old[g][i] = ((freemask >> g) & 1u)
					                ? mem_atomic_or(words + (slot[g][i] >> 5), 1u << (uint32_t)(slot[g][i] & 31))
					                : 0u;
#pragma unroll
			for (int g = 0; g < G; g++) {
				uint32_t found = (freemask >> g) & 1u;
#pragma unroll
				for (int i = 0; i < H; i++)
					found &= old[g][i] >> (uint32_t)(slot[g][i] & 31);
				out.hitw |= (found & 1u) << (s0 + g);
			}
		}
		// the others wait for the index-ordered residual rounds
		const uint32_t defer = okmask & ~freemask;
#pragma unroll
		for (int g = 0; g < G; g++) {
			if (!((defer >> g) & 1u))
				continue;
			const uint32_t w = p0 + s0 + g;
			const uint32_t at = mem_atomic_inc(P.pending_count);
			P.pending[at] = (uint32_t)(t0 + w);
			if (P.pending_slots) {
#pragma unroll
				for (int i = 0; i < H; i++)
					P.pending_slots[(t0 + w) * H + i] = slot[g][i];
			}
		}
	}
	return out;
}

template<int OP, bool POW2>
BTL_HD bool tile_phase_c_commit_any(const SeqParams& P, const TileSmem& sm, uint64_t t0, int tid, ThreadOut& out)
{
	switch (P.h) {
	case 1: out = tile_phase_c_commit<OP, 1, POW2>(P, sm, t0, tid); return true;
	case 2: out = tile_phase_c_commit<OP, 2, POW2>(P, sm, t0, tid); return true;
	case 3: out = tile_phase_c_commit<OP, 3, POW2>(P, sm, t0, tid); return true;
	case 4: out = tile_phase_c_commit<OP, 4, POW2>(P, sm, t0, tid); return true;
	case 5: out = tile_phase_c_commit<OP, 5, POW2>(P, sm, t0, tid); return true;
	case 6: out = tile_phase_c_commit<OP, 6, POW2>(P, sm, t0, tid); return true;
	case 7: out = tile_phase_c_commit<OP, 7, POW2>(P, sm, t0, tid); return true;
	case 8: out = tile_phase_c_commit<OP, 8, POW2>(P, sm, t0, tid); return true;
	}
	return false;
}

template<int OP, bool SPACED, bool POW2>
BTL_HD ThreadOut tile_phase_c(const SeqParams& P, const TileSmem& sm, uint64_t t0, int tid)
{
	if constexpr ((OP == OP_CBF_COMMIT || OP == OP_BFCHK_COMMIT) && !SPACED) { // (instantiated for these two only)
		if (!P.ungrouped_commit) {
			ThreadOut grouped;
			if (tile_phase_c_commit_any<OP, POW2>(P, sm, t0, tid, grouped))
				return grouped;
		}
	}
	ThreadOut out;
	out.validw = 0;
	out.hitw = 0;
	uint64_t nwin64 = P.n_windows > t0 ? P.n_windows - t0 : 0;
	const uint32_t p0 = (uint32_t)tid * kWPT;
	if (p0 >= nwin64)
		return out;
	roll_windows(P, sm, t0, tid, [&](uint32_t s, bool ok, uint64_t F, uint64_t RC) {
		if (ok)
			window_op<OP, SPACED, POW2>(P, sm, t0, p0 + s, s, F, RC, out);
	});
	return out;
}

// ---------------------------------------------------------------- deferred-list phases (exact counting insert)
// One thread = one deferred window; hashes are re-derived directly from the bases (O(k)).
template<bool POW2, class Fn>
BTL_HD void list_for_each_hash(const SeqParams& P, uint32_t w, Fn&& fn)
{
	if (P.pending_slots) { // recorded when the window was deferred
		const uint64_t* src = P.pending_slots + (uint64_t)w * P.h;
		for (uint32_t i = 0; i < P.h; i++)
			if (!fn(ld_cg(src + i)))
				return;
		return;
	}
	const uint8_t* s = P.bases + w;
	const uint32_t k = P.k;
	// class of base i of the window (a deferred window is valid: no invalid base inside)
	auto cls = [&](uint32_t i) -> unsigned {
		if (P.packed) {
			const uint64_t q = (uint64_t)w + i;
			return (P.bases[q >> 2] >> (2 * (uint32_t)(q & 3))) & 3u;
		}
		return base_class(s[i]);
	};
	uint64_t F = 0, RC = 0;
	for (uint32_t i = 0; i < k; i++) {
		F = srol(F) ^ class_fseed(cls(i));
		RC = srol(RC) ^ class_rseed(cls(k - 1 - i));
	}
	if (P.n_seeds == 0) {
		uint64_t b = RC < F ? RC : F;
		if (!fn(fastmod<POW2>(b, P.fm)))
			return;
		for (uint32_t i = 1; i < P.h; i++)
			if (!fn(fastmod<POW2>(multi_mix(b, P.mult[i]), P.fm)))
				return;
	} else {
		const uint64_t* TF = P.st_tab;
		const uint64_t* TR = P.st_tab + (size_t)k * 8;
		for (uint32_t j = 0; j < P.n_seeds; j++) {
			uint64_t fs = F, rs = RC;
			for (uint32_t t = P.st_dc_off[j]; t < P.st_dc_off[j + 1]; t++) {
				uint32_t pos = P.st_dc[t];
				uint32_t c = cls(pos) & 7u;
				fs ^= TF[pos * 8 + c];
				rs ^= TR[pos * 8 + c];
			}
			uint64_t b = rs < fs ? rs : fs;
			if (!fn(fastmod<POW2>(b, P.fm)))
				return;
			for (uint32_t j2 = 1; j2 < P.h2; j2++)
				if (!fn(fastmod<POW2>(multi_mix(b, P.mult[j2]), P.fm)))
					return;
		}
	}
}

// Residual rounds of the ordered updates ("deterministic reservations").  Every pending k-mer writes
// tag = (~epoch, window index) into the reservation table at each of its (hashed) slots with
// atomicMin: the smallest tag wins, i.e. the newest round and, within it, the earliest k-mer in
// reference order.  A k-mer holding all its reservations precedes every pending k-mer it conflicts
// with, so applying its update now is exactly the sequential result; the others wait a round.
// KIND 0: CountingBloomFilter::incrementMin, KIND 1: BloomFilter::insertAndCheck.
BTL_HD uint64_t list_tag(uint32_t epoch, uint32_t w)
{
	return ((uint64_t)(0xffffffffu - epoch) << 32) | w;
}

template<bool POW2>
BTL_HD void list_round_reserve(const SeqParams& P, uint64_t* R, uint64_t emask, uint32_t epoch, uint32_t w)
{
	uint64_t tag = list_tag(epoch, w);
	list_for_each_hash<POW2>(P, w, [&](uint64_t slot) {
		mem_atomic_min64(R + (slot & emask), tag);
		return true;
	});
}

// returns true when the k-mer was committed, false when it must wait for the next round
template<bool POW2, int KIND>
BTL_HD bool list_round_commit(const SeqParams& P, const uint64_t* R, uint64_t emask, uint32_t epoch, uint32_t w)
{
	uint64_t tag = list_tag(epoch, w);
	bool own = true;
	list_for_each_hash<POW2>(P, w, [&](uint64_t slot) {
		own = ld_cg(R + (slot & emask)) == tag;
		return own;
	});
	if (!own)
		return false;
	if (KIND == 0) {
		uint8_t* cnt = (uint8_t*)P.filter;
		uint32_t mn = 255;
		list_for_each_hash<POW2>(P, w, [&](uint64_t slot) {
			uint32_t v = ld_cg(cnt + slot);
			mn = v < mn ? v : mn;
			return true;
		});
		if (mn != 255)
			list_for_each_hash<POW2>(P, w, [&](uint64_t slot) {
				if (ld_cg(cnt + slot) == mn)
					*(volatile uint8_t*)(cnt + slot) = (uint8_t)(mn + 1);
				return true;
			});
		if (mn >= P.threshold && P.hit_bits)
			mem_red_or(P.hit_bits + (w >> 5), 1u << (w & 31));
	} else {
		uint32_t* words = (uint32_t*)P.filter;
		bool found = true;
		list_for_each_hash<POW2>(P, w, [&](uint64_t n) {
			uint32_t bit = 1u << (uint32_t)(n & 31);
			found &= (mem_atomic_or(words + (n >> 5), bit) & bit) != 0;
			return true;
		});
		if (found && P.hit_bits)
			mem_red_or(P.hit_bits + (w >> 5), 1u << (w & 31));
	}
	return true;
}

} // namespace btl
