// sort_bin.cuh -- pass 1 of the partitioned BloomFilter build / query, fast flavour (device only).
//
// Serves h <= kMaxSortHashes hashes per k-mer and <= kMaxSortBins filter partitions (everything the
// BASELINE configs use); other shapes fall back to bin_kernel_warp / bin_kernel_cta in kernels.cu.
//
// Persistent CTAs of kSortThreads threads; every CTA is the only writer of its own sub-bucket of each
// filter partition.  A thread rolls kWPT consecutive windows (same staging and rolling as seq_kernel);
// after every W of them the CTA counting-sorts the kSortThreads*W*h items it holds in registers:
//   A   item -> (partition, offset); its rank inside the partition is the return value of a
//       shared-memory atomicAdd on the round's histogram
//   S   exclusive scan of the histogram; per partition: advance the sub-bucket cursor and remember
//       gdelta = (global item index of the run's first item) - (its position in the sorted buffer)
//   C   scatter the items from registers into the sorted shared-memory buffer
//   D   copy the buffer out: consecutive threads write consecutive items, so each partition's run
//       leaves as one contiguous, coalesced piece of its sub-bucket
// That is ~7 warp-instructions per k-mer against ~20 for the warp-private staging-line kernel.
// The cursors live in bin_counts between launches: successive launches APPEND to the same sub-buckets
// (the host zeroes bin_counts when it starts a new accumulation), which lets pass 2 stream the filter
// once per many batches.  An item that does not fit its sub-bucket (skewed input) is applied / probed
// directly -- OR and AND-of-probes are order-free, so any mix of the two paths is exact.
#pragma once
#include "tile_core.cuh"

namespace btl {

constexpr int kSortThreads = 512;
constexpr int kSortTile = kSortThreads * kWPT; // windows per CTA pass (16384)
constexpr uint32_t kMaxSortBins = 1024;        // the scan handles two partitions per thread
constexpr int kMaxSortHashes = 8;

BTL_HD constexpr int sort_round_windows(int h)
{
	return h <= 2 ? 8 : h <= 4 ? 4 : 2;
}

// kernel entry points by shape; instantiated in sort_bin_build.cu (QUERY = false) and sort_bin_query.cu
const void* bin_sort_kernel_build(int h, bool spaced, bool pow2);
const void* bin_sort_kernel_query(int h, bool spaced, bool pow2);

#if defined(__CUDACC__)
__device__ __forceinline__ void bin_direct_or(const SeqParams& P, uint32_t part, uint32_t off)
{
	uint64_t n = ((uint64_t)part << P.bin_shift) | off;
	atomicOr((uint32_t*)P.filter + (n >> 5), 1u << (uint32_t)(n & 31));
}

// query flavour: the bit is tested right away; a miss clears the window's hit bit
__device__ __forceinline__ void bin_direct_probe(const SeqParams& P, uint32_t part, uint32_t off, uint32_t wid)
{
	uint64_t n = ((uint64_t)part << P.bin_shift) | off;
	if (!((__ldg((const uint32_t*)P.filter + (n >> 5)) >> (uint32_t)(n & 31)) & 1u))
		atomicAnd(P.hit_bits + (wid >> 5), ~(1u << (wid & 31)));
}

// the rolling hash of roll_windows() (tile_core.cuh) as an explicit state machine: one step() per window
struct Roller
{
	uint64_t F, RC, in_codes, out_codes;
	uint32_t in_bad, in_start, g, p0, q1, nwin;
	bool generic;

	__device__ __forceinline__ void init(const SeqParams& P, const TileSmem& sm, uint64_t t0, int tid, uint32_t tile)
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
		// register streams: 32 incoming bases from q1 (unaligned), 32 outgoing bases from p0 (aligned)
		uint32_t a = q1 >> 4, sh2 = 2 * (q1 & 15);
		uint32_t in_lo = funnel_r(sm.codes[a], sm.codes[a + 1], sh2);
		uint32_t in_hi = funnel_r(sm.codes[a + 1], sm.codes[a + 2], sh2);
		uint32_t bw = q1 >> 5, sh1 = q1 & 31;
		in_bad = funnel_r(sm.badw[bw], sm.badw[bw + 1], sh1);
		in_start = funnel_r(sm.startw[bw], sm.startw[bw + 1], sh1);
		in_codes = ((uint64_t)in_hi << 32) | in_lo;
		out_codes = ((uint64_t)sm.codes[(p0 >> 4) + 1] << 32) | sm.codes[p0 >> 4];
	}

	// advances to window p0+s (s = 0, 1, 2, ... in order); true when it is a k-mer the reference's
	// iterator visits (ntHashIterator.hpp:59-86)
	__device__ __forceinline__ bool step(const SeqParams& P, const TileSmem& sm, uint32_t s)
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

// the H hashes of one window, statically indexed (SPACED: H == n_seeds, one hash per seed mask;
// nthash.hpp:684-690 and :820-878)
template<int H, bool SPACED>
__device__ __forceinline__ void expand_hashes(const SeqParams& P, const TileSmem& sm, uint32_t w, uint64_t F, uint64_t RC,
                                              uint64_t (&hv)[H])
{
	if (!SPACED) {
		uint64_t b = RC < F ? RC : F;
		hv[0] = b;
#pragma unroll
		for (int i = 1; i < H; i++)
			hv[i] = multi_mix(b, P.mult[i]);
	} else {
		const uint64_t* TF = sm.sttab;
		const uint64_t* TR = sm.sttab + (size_t)P.k * 8;
#pragma unroll
		for (int j = 0; j < H; j++) {
			uint64_t fs = F, rs = RC;
			for (uint32_t t = P.st_dc_off[j]; t < P.st_dc_off[j + 1]; t++) {
				uint32_t pos = P.st_dc[t];
				uint32_t c = sm.tile[w + pos] & 7u;
				fs ^= TF[pos * 8 + c];
				rs ^= TR[pos * 8 + c];
			}
			hv[j] = rs < fs ? rs : fs;
		}
	}
}

#endif // __CUDACC__

// dynamic shared memory of bin_kernel_sort: the sort arrays, then the tile staging area
BTL_HD size_t sort_arrays_bytes(uint32_t n_bins, int h)
{
	const size_t nbr = (n_bins + 1u) & ~1u;
	const size_t cap = (size_t)kSortThreads * sort_round_windows(h) * h;
	size_t s = nbr * 8;                 // gdelta
	s += cap * 8;                       // sorted_off, sorted_aux
	s += nbr * 4 * 3;                   // hist, base, cursor
	s += (kSortThreads / 32 + 2) * 4;   // warp sums, total, overflow flag
	return (s + 15) / 16 * 16;
}

inline size_t sort_smem_bytes(uint32_t k, bool spaced, uint32_t n_bins, int h)
{
	return sort_arrays_bytes(n_bins, h) + tile_smem_bytes(k, spaced, 0, kSortTile);
}

#if defined(__CUDACC__)
template<int H, bool SPACED, bool POW2, bool QUERY>
__global__ void __launch_bounds__(kSortThreads, 2) bin_kernel_sort(const __grid_constant__ SeqParams P)
{
	constexpr int W = sort_round_windows(H);
	constexpr int ITEMS = W * H;
	constexpr uint32_t CAPACITY = (uint32_t)kSortThreads * ITEMS;
	constexpr uint32_t NW = kSortThreads / 32;
	constexpr uint32_t kNone = 0xffffffffu;
	extern __shared__ __align__(16) uint8_t smem_raw[];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t nb = P.n_bins, nbr = (nb + 1u) & ~1u;
	const uint32_t writer = blockIdx.x;

	uint8_t* p = smem_raw;
	uint64_t* const gdelta = reinterpret_cast<uint64_t*>(p);     p += (size_t)nbr * 8;
	uint32_t* const sorted_off = reinterpret_cast<uint32_t*>(p); p += (size_t)CAPACITY * 4;
	uint32_t* const sorted_aux = reinterpret_cast<uint32_t*>(p); p += (size_t)CAPACITY * 4;
	uint32_t* const hist = reinterpret_cast<uint32_t*>(p);       p += (size_t)nbr * 4;
	uint32_t* const base = reinterpret_cast<uint32_t*>(p);       p += (size_t)nbr * 4;
	uint32_t* const cursor = reinterpret_cast<uint32_t*>(p);     p += (size_t)nbr * 4;
	uint32_t* const wsum = reinterpret_cast<uint32_t*>(p); // [NW] warp sums, [NW] total, [NW+1] overflow flag
	const TileSmem sm = carve_smem(smem_raw + sort_arrays_bytes(nb, H), P.k, SPACED, 0, kSortTile);

	for (uint32_t b = tid; b < nbr; b += kSortThreads) {
		hist[b] = 0;
		cursor[b] = b < nb ? P.bin_counts[(uint64_t)b * P.bin_writers + writer] : 0u;
	}

	// tile t belongs to writer (t + bin_rot) % gridDim.x: the host advances bin_rot from launch to launch so
	// that a stream of small batches still spreads evenly over the writers' sub-buckets
	const uint64_t tiles = (P.n_windows + kSortTile - 1) / kSortTile;
	const uint32_t first = (blockIdx.x + gridDim.x - P.bin_rot % gridDim.x) % gridDim.x;
	for (uint64_t t = first; t < tiles; t += gridDim.x) {
		const uint64_t t0 = t * kSortTile;
		__syncthreads(); // the previous tile is fully consumed before its staging area is overwritten
		tile_phase_a(P, sm, t0, tid, kSortThreads);
		__syncthreads();
		tile_phase_b(P, sm, t0, tid, kSortThreads);
		__syncthreads();
		Roller r;
		r.init(P, sm, t0, tid, kSortTile);
		uint32_t validw = 0;
		for (uint32_t round = 0; round < (uint32_t)(kWPT / W); round++) {
			// ---- A: hash, split into (partition, offset), rank by histogram atomics
			uint32_t it_off[ITEMS], it_pr[ITEMS];
#pragma unroll
			for (int ws = 0; ws < W; ws++) {
				const uint32_t s = round * W + ws;
				const bool ok = r.step(P, sm, s);
				validw |= (uint32_t)ok << s;
				uint64_t hv[H];
				expand_hashes<H, SPACED>(P, sm, r.p0 + s, r.F, r.RC, hv);
#pragma unroll
				for (int i = 0; i < H; i++) {
					const uint64_t n = fastmod<POW2>(hv[i], P.fm);
					const uint32_t part = (uint32_t)(n >> P.bin_shift);
					it_off[ws * H + i] = (uint32_t)n & P.bin_mask;
					it_pr[ws * H + i] = ok ? (part << 16) | atomicAdd(hist + part, 1u) : kNone;
				}
			}
			__syncthreads();
			// ---- S: exclusive scan over the partitions (two per thread), cursors, gdelta
			uint32_t v0 = 0, v1 = 0;
			if (2u * tid < nbr) {
				v0 = hist[2 * tid];
				v1 = hist[2 * tid + 1];
				hist[2 * tid] = 0;
				hist[2 * tid + 1] = 0;
			}
			uint32_t incl = v0 + v1;
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
				if (lane >= o)
					incl += y;
			}
			if (lane == 31)
				wsum[warp] = incl;
			if (tid == 0)
				wsum[NW + 1] = 0;
			__syncthreads();
			uint32_t before = 0, total = 0;
#pragma unroll
			for (uint32_t w = 0; w < NW; w++) {
				uint32_t x = wsum[w];
				before += w < (uint32_t)warp ? x : 0u;
				total += x;
			}
			if (tid == 0)
				wsum[NW] = total;
			if (2u * tid < nbr) {
				uint32_t e0 = before + incl - v0 - v1, e1 = e0 + v0;
#pragma unroll
				for (int j = 0; j < 2; j++) {
					const uint32_t b = 2 * tid + j, e = j ? e1 : e0, v = j ? v1 : v0;
					base[b] = e;
					if (v) {
						const uint32_t c = cursor[b];
						uint32_t cn = c + v;
						cn = cn < c ? kNone : cn;
						cursor[b] = cn;
						gdelta[b] = ((uint64_t)b * P.bin_writers + writer) * P.bin_cap + c - e;
						if (cn > P.bin_cap)
							wsum[NW + 1] = 1;
					}
				}
			}
			__syncthreads();
			// ---- C: scatter from registers into the sorted buffer
#pragma unroll
			for (int ws = 0; ws < W; ws++) {
#pragma unroll
				for (int i = 0; i < H; i++) {
					const uint32_t pr = it_pr[ws * H + i];
					if (pr != kNone) {
						const uint32_t part = pr >> 16;
						const uint32_t pos = base[part] + (pr & 0xffffu);
						sorted_off[pos] = it_off[ws * H + i];
						sorted_aux[pos] = QUERY ? part | ((r.p0 + round * W + ws) << 12) : part;
					}
				}
			}
			__syncthreads();
			// ---- D: copy out (the next round's phase A only touches hist, so no barrier is needed after this)
			total = wsum[NW];
			const bool overflow = wsum[NW + 1] != 0;
			for (uint32_t pos = tid; pos < total; pos += kSortThreads) {
				const uint32_t off = sorted_off[pos], aux = sorted_aux[pos];
				const uint32_t part = QUERY ? aux & 0xfffu : aux;
				const uint64_t idx = gdelta[part] + pos;
				const uint32_t wid = (uint32_t)t0 + (aux >> 12);
				if (overflow && idx - ((uint64_t)part * P.bin_writers + writer) * P.bin_cap >= P.bin_cap) {
					if (QUERY)
						bin_direct_probe(P, part, off, wid);
					else
						bin_direct_or(P, part, off);
					continue;
				}
				if (QUERY)
					reinterpret_cast<uint2*>(P.bin_items)[idx] = make_uint2(off, wid);
				else
					P.bin_items[idx] = off;
			}
		}
		const uint64_t widx = (t0 >> 5) + tid;
		if (P.valid_bits && widx < P.out_words)
			P.valid_bits[widx] = validw;
		if (P.stats) {
			uint32_t nv = __reduce_add_sync(0xffffffffu, __popc(validw));
			if (lane == 0 && nv)
				atomicAdd((unsigned long long*)&P.stats[0], (unsigned long long)nv);
		}
	}
	__syncthreads();
	for (uint32_t b = tid; b < nb; b += kSortThreads)
		P.bin_counts[(uint64_t)b * P.bin_writers + writer] = cursor[b];
}

template<bool QUERY, int H>
static const void* bin_sort_kernel_shape(bool spaced, bool pow2)
{
	if (spaced)
		return pow2 ? (const void*)bin_kernel_sort<H, true, true, QUERY> : (const void*)bin_kernel_sort<H, true, false, QUERY>;
	return pow2 ? (const void*)bin_kernel_sort<H, false, true, QUERY> : (const void*)bin_kernel_sort<H, false, false, QUERY>;
}

template<bool QUERY>
static const void* bin_sort_kernel_any(int h, bool spaced, bool pow2)
{
	switch (h) {
	case 1: return bin_sort_kernel_shape<QUERY, 1>(spaced, pow2);
	case 2: return bin_sort_kernel_shape<QUERY, 2>(spaced, pow2);
	case 3: return bin_sort_kernel_shape<QUERY, 3>(spaced, pow2);
	case 4: return bin_sort_kernel_shape<QUERY, 4>(spaced, pow2);
	case 5: return bin_sort_kernel_shape<QUERY, 5>(spaced, pow2);
	case 6: return bin_sort_kernel_shape<QUERY, 6>(spaced, pow2);
	case 7: return bin_sort_kernel_shape<QUERY, 7>(spaced, pow2);
	case 8: return bin_sort_kernel_shape<QUERY, 8>(spaced, pow2);
	}
	return nullptr;
}
#endif // __CUDACC__

} // namespace btl
