// sort_bin.cuh -- pass 1 of the partitioned BloomFilter build / query, fast flavour (device only).
//
// Serves h <= kMaxSortHashes hashes per k-mer and <= kMaxSortBins filter partitions (everything the
// BASELINE configs use); other shapes fall back to bin_kernel_warp / bin_kernel_cta in kernels.cu.
//
// Persistent CTAs of 256 or 512 threads (see sort_threads_for()); every CTA is the only writer of its own sub-bucket of each
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

// Two CTA shapes are built: 256 threads (four CTAs per SM: cheaper barriers, the faster one whenever a round still
// gives every partition a run of ~16 items) and 512 threads (two per SM: rounds twice as large, for many partitions
// or few hashes per k-mer).  sort_threads_for() picks.
BTL_HD constexpr int sort_ctas_per_sm(int threads) { return 1024 / threads; } // 64 registers per thread: 1024 threads fill the register file
BTL_HD constexpr uint32_t sort_tile(int threads) { return (uint32_t)threads * kWPT; } // windows per CTA pass
BTL_HD constexpr uint32_t sort_max_bins(int threads) { return 2u * (uint32_t)threads; } // the scan handles two partitions per thread
constexpr int kMaxSortHashes = 8;

BTL_HD constexpr int sort_round_windows(int h)
{
	return h <= 2 ? 8 : h <= 4 ? 4 : 2;
}

// kernel entry points by shape; instantiated in sort_bin_build.cu (QUERY = false) and sort_bin_query.cu
const void* bin_sort_kernel_build_256(int h, bool spaced, bool pow2);
const void* bin_sort_kernel_build_512(int h, bool spaced, bool pow2);
const void* bin_sort_kernel_query_256(int h, bool spaced, bool pow2);
const void* bin_sort_kernel_query_512(int h, bool spaced, bool pow2);

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
	const bool ok = P.bin_counting ? __ldg((const uint8_t*)P.filter + n) >= P.threshold // CountingBloomFilter.hpp:190-196
	                               : ((__ldg((const uint32_t*)P.filter + (n >> 5)) >> (uint32_t)(n & 31)) & 1u) != 0;
	if (!ok)
		atomicAnd(P.hit_bits + (wid >> 5), ~(1u << (wid & 31)));
}

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
		const bool packed = !P.force_generic && sm.scratch[1] == 0; // see for_each_hash (tile_core.cuh)
#pragma unroll
		for (int j = 0; j < H; j++) {
			uint64_t fs = F, rs = RC;
			for (uint32_t t = P.st_dc_off[j]; t < P.st_dc_off[j + 1]; t++) {
				uint32_t pos = P.st_dc[t];
				uint32_t q = w + pos;
				uint32_t c = packed ? (sm.codes[q >> 4] >> (2 * (q & 15))) & 3u : sm.tile[q] & 7u;
				fs ^= TF[pos * 8 + c];
				rs ^= TR[pos * 8 + c];
			}
			hv[j] = rs < fs ? rs : fs;
		}
	}
}

#endif // __CUDACC__

// dynamic shared memory of bin_kernel_sort: the sort arrays, then the tile staging area
BTL_HD size_t sort_arrays_bytes(uint32_t n_bins, int h, int threads)
{
	const size_t nbr = (n_bins + 1u) & ~1u;
	const size_t cap = (size_t)threads * sort_round_windows(h) * h;
	size_t s = nbr * 8;                 // gdelta
	s += cap * 8;                       // sorted: (offset, partition [| window]) pairs
	s += nbr * 4 * 3 + 32 * 4;          // hist (+ 32 per-lane dump slots for invalid windows), base, cursor
	s += (size_t)(threads / 32 + 2) * 4;   // warp sums, total, overflow flag
	return (s + 15) / 16 * 16;
}

inline size_t sort_smem_bytes(uint32_t k, bool spaced, uint32_t n_bins, int h, int threads)
{
	return sort_arrays_bytes(n_bins, h, threads) + tile_smem_bytes(k, spaced, 0, sort_tile(threads));
}

#if defined(__CUDACC__)
template<int THREADS, int H, bool SPACED, bool POW2, bool QUERY>
__global__ void __launch_bounds__(THREADS, sort_ctas_per_sm(THREADS)) bin_kernel_sort(const __grid_constant__ SeqParams P)
{
	constexpr int kSortThreads = THREADS;
	constexpr uint32_t kSortTile = sort_tile(THREADS);
	constexpr int W = sort_round_windows(H);
	constexpr int ITEMS = W * H;
	constexpr uint32_t CAPACITY = (uint32_t)kSortThreads * ITEMS;
	constexpr uint32_t NW = kSortThreads / 32;
	extern __shared__ __align__(16) uint8_t smem_raw[];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t nb = P.n_bins, nbr = (nb + 1u) & ~1u;
	const uint32_t writer = blockIdx.x;
	if (P.gate && *P.gate != P.gate_want) // adaptive query: the other path was chosen for this batch
		return;

	uint8_t* p = smem_raw;
	uint64_t* const gdelta = reinterpret_cast<uint64_t*>(p);     p += (size_t)nbr * 8;
	uint2* const sorted = reinterpret_cast<uint2*>(p);           p += (size_t)CAPACITY * 8; // gdelta keeps this 8-byte aligned
	uint32_t* const hist = reinterpret_cast<uint32_t*>(p);       p += (size_t)(nbr + 32) * 4;
	uint32_t* const base = reinterpret_cast<uint32_t*>(p);       p += (size_t)nbr * 4;
	uint32_t* const cursor = reinterpret_cast<uint32_t*>(p);     p += (size_t)nbr * 4;
	uint32_t* const wsum = reinterpret_cast<uint32_t*>(p); // [NW] warp sums, [NW] total, [NW+1] overflow flag
	const TileSmem sm = carve_smem(smem_raw + sort_arrays_bytes(nb, H, THREADS), P.k, SPACED, 0, kSortTile);

	for (uint32_t b = tid; b < nbr; b += kSortThreads) {
		hist[b] = 0;
		cursor[b] = b < nb ? P.bin_counts[(uint64_t)b * P.bin_writers + writer] : 0u;
	}
	if (tid < 32)
		hist[nbr + tid] = 0;
	// Items of windows that are not k-mers are counted in a per-lane dump slot behind the histogram instead of
	// being branched around: the four atomics of a window then issue back to back and their latencies overlap.
	const uint32_t dump = nbr + (uint32_t)lane;

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
					const uint32_t bin = ok ? part : dump;
					it_off[ws * H + i] = (uint32_t)n & P.bin_mask;
					it_pr[ws * H + i] = (bin << 16) | (atomicAdd(hist + bin, 1u) & 0xffffu);
				}
			}
			__syncthreads();
			// ---- S: exclusive scan over the partitions (two per thread), cursors, gdelta
			uint32_t v0 = 0, v1 = 0, incl = 0;
			if (64u * (uint32_t)warp < nbr) { // warps whose 64 partitions do not exist skip the scan (warp-uniform)
				if (2u * tid < nbr) {
					v0 = hist[2 * tid];
					v1 = hist[2 * tid + 1];
					hist[2 * tid] = 0;
					hist[2 * tid + 1] = 0;
				}
				incl = v0 + v1;
#pragma unroll
				for (int o = 1; o < 32; o <<= 1) {
					uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
					if (lane >= o)
						incl += y;
				}
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
						cn = cn < c ? 0xffffffffu : cn;
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
					const uint32_t part = pr >> 16;
					if (part < nbr) {
						const uint32_t pos = base[part] + (pr & 0xffffu);
						sorted[pos] = make_uint2(it_off[ws * H + i], QUERY ? part | ((r.p0 + round * W + ws) << 12) : part);
					}
				}
			}
			__syncthreads();
			// ---- D: copy out (the next round's phase A only touches hist, so no barrier is needed after this)
			total = wsum[NW];
			const bool overflow = wsum[NW + 1] != 0;
#pragma unroll 4
			for (uint32_t pos = tid; pos < total; pos += kSortThreads) {
				const uint2 item = sorted[pos];
				const uint32_t off = item.x, aux = item.y;
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

template<int THREADS, bool QUERY, int H>
static const void* bin_sort_kernel_shape(bool spaced, bool pow2)
{
	if (spaced)
		return pow2 ? (const void*)bin_kernel_sort<THREADS, H, true, true, QUERY>
		            : (const void*)bin_kernel_sort<THREADS, H, true, false, QUERY>;
	return pow2 ? (const void*)bin_kernel_sort<THREADS, H, false, true, QUERY>
	            : (const void*)bin_kernel_sort<THREADS, H, false, false, QUERY>;
}

template<int THREADS, bool QUERY>
static const void* bin_sort_kernel_any(int h, bool spaced, bool pow2)
{
	switch (h) {
	case 1: return bin_sort_kernel_shape<THREADS, QUERY, 1>(spaced, pow2);
	case 2: return bin_sort_kernel_shape<THREADS, QUERY, 2>(spaced, pow2);
	case 3: return bin_sort_kernel_shape<THREADS, QUERY, 3>(spaced, pow2);
	case 4: return bin_sort_kernel_shape<THREADS, QUERY, 4>(spaced, pow2);
	case 5: return bin_sort_kernel_shape<THREADS, QUERY, 5>(spaced, pow2);
	case 6: return bin_sort_kernel_shape<THREADS, QUERY, 6>(spaced, pow2);
	case 7: return bin_sort_kernel_shape<THREADS, QUERY, 7>(spaced, pow2);
	case 8: return bin_sort_kernel_shape<THREADS, QUERY, 8>(spaced, pow2);
	}
	return nullptr;
}
#endif // __CUDACC__

} // namespace btl
