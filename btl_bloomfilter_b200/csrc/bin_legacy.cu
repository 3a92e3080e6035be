// bin_legacy.cu -- pass 1 of the partitioned build / query for the shapes the sort-bin kernel (sort_bin.cuh) does
// not serve: more than kMaxSortHashes hashes per k-mer, spaced seeds with h2 > 1, more partitions than the sort
// handles.  bin_kernel_warp: warp-private cursors and 32-byte staging lines; bin_kernel_cta: CTA cursors.
#include "sort_bin.cuh"
#include "seq_kernel.cuh"

namespace btl {

// ---------------------------------------------------------------- partitioned build, pass 1
// Persistent CTAs.  Every WARP is the only writer of its own sub-bucket of each filter partition:
//   * the warp's append cursors and one 32-byte staging line per partition live in shared memory;
//   * lanes that hit the same partition in one step are serialised by an optimistic claim on the cursor
//     word (warp_bin_emit) -- no atomics;
//   * offsets are written to the staging line and leave for HBM as whole 32-byte sectors.
// Shared memory per warp: n_bins * 36 bytes, so this kernel serves n_bins <= kMaxWarpBins; filters with
// more partitions use bin_kernel_cta below (CTA-shared cursors, shared-memory atomics, 4-byte stores).
constexpr uint32_t kFullMask = 0xffffffffu;

struct WarpBins
{
	uint32_t* cursor;   // [n_bins] cursor words of this warp: items appended so far << kTagBits | claimant tag
	uint32_t* line;     // [n_bins][8] current partial line
	uint32_t cursor_sa; // the same two arrays as shared-window byte addresses (cheap addressing in the hot loop)
	uint32_t line_sa;
	uint32_t writer;    // global warp index
};

__device__ __forceinline__ uint32_t lds_u32(uint32_t sa)
{
	uint32_t v;
	asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sa) : "memory");
	return v;
}
__device__ __forceinline__ void sts_u32(uint32_t sa, uint32_t v)
{
	asm volatile("st.shared.u32 [%0], %1;" ::"r"(sa), "r"(v) : "memory");
}
__device__ __forceinline__ void sts_u64(uint32_t sa, uint32_t lo, uint32_t hi)
{
	asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sa), "r"(lo), "r"(hi) : "memory");
}

// Stores one full 32-byte staging line at sub-bucket item position `start`: 8 offsets (build), or 4
// (offset, window) pairs (QUERY).  A sub-bucket that is full (skewed input) handles the line's items
// directly instead -- OR / AND-of-probes are order-free, so any mix of the two paths is exact.
template<bool QUERY>
__device__ __forceinline__ void bin_flush_line(const SeqParams& P, const WarpBins& wb, uint32_t part, uint32_t start)
{
	const uint4* src = reinterpret_cast<const uint4*>(wb.line + part * 8);
	uint4 a = src[0], b = src[1];
	if (start < P.bin_cap) {
		uint64_t item0 = ((uint64_t)part * P.bin_writers + wb.writer) * P.bin_cap + start;
		uint4* dst = reinterpret_cast<uint4*>(P.bin_items + (QUERY ? item0 * 2 : item0));
		dst[0] = a;
		dst[1] = b;
	} else if (QUERY) {
		bin_direct_probe(P, part, a.x, a.y); bin_direct_probe(P, part, a.z, a.w);
		bin_direct_probe(P, part, b.x, b.y); bin_direct_probe(P, part, b.z, b.w);
	} else {
		bin_direct_or(P, part, a.x); bin_direct_or(P, part, a.y); bin_direct_or(P, part, a.z); bin_direct_or(P, part, a.w);
		bin_direct_or(P, part, b.x); bin_direct_or(P, part, b.y); bin_direct_or(P, part, b.z); bin_direct_or(P, part, b.w);
	}
}

// warp-synchronous: every lane of the warp calls this the same number of times.
// Appends one item per active lane to the warp's sub-buckets.  Lanes that target the same partition are
// serialised by an optimistic claim on the partition's cursor word (count << kTagBits | claiming lane):
// all pending lanes read the word, all write count+1 tagged with their lane, and the lane whose tag
// survived owns position `count`; the others retry against the updated word.  With 32 lanes over a few
// hundred partitions this takes two rounds on average and needs no atomics, ballots or match instructions.
constexpr uint32_t kTagBits = 5, kTagMask = (1u << kTagBits) - 1u;

template<bool QUERY>
__device__ __forceinline__ void warp_bin_emit(const SeqParams& P, const WarpBins& wb, uint64_t n, uint32_t wid, bool active)
{
	constexpr uint32_t L = QUERY ? 4u : 8u; // items per 32-byte line
	const uint32_t lane = threadIdx.x & 31;
	const uint32_t part = (uint32_t)(n >> P.bin_shift);
	const uint32_t off = (uint32_t)n & P.bin_mask;
	const uint32_t cur = wb.cursor_sa + part * 4u;
	bool pending = active;
	while (__any_sync(kFullMask, pending)) {
		uint32_t c = 0;
		if (pending)
			c = lds_u32(cur);
		__syncwarp(); // every read of this round precedes every write
		if (pending)
			sts_u32(cur, ((c & ~kTagMask) + (kTagMask + 1u)) | lane);
		__syncwarp();
		if (pending && (lds_u32(cur) & kTagMask) == lane) {
			const uint32_t pos = c >> kTagBits;
			if (QUERY)
				sts_u64(wb.line_sa + part * 32u + (pos & (L - 1)) * 8u, off, wid);
			else
				sts_u32(wb.line_sa + part * 32u + (pos & (L - 1)) * 4u, off);
			// the item that completes a line stores it: its other slots were filled in earlier rounds (at
			// most one item per partition is placed per round, and two warp barriers separate the rounds)
			if ((pos & (L - 1)) == L - 1)
				bin_flush_line<QUERY>(P, wb, part, pos - (L - 1));
			pending = false;
		}
	}
}

template<bool SPACED, bool POW2, bool QUERY>
__global__ void __launch_bounds__(kTPB) bin_kernel_warp(const __grid_constant__ SeqParams P)
{
	constexpr uint32_t L = QUERY ? 4u : 8u;
	extern __shared__ __align__(32) uint8_t smem_raw[];
	const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
	// [warp][n_bins] cursors, [warp][n_bins][8] lines, then the tile staging area
	uint32_t* cur_all = reinterpret_cast<uint32_t*>(smem_raw);
	const uint32_t nb = P.n_bins, nb8 = (nb + 7u) & ~7u;
	uint32_t* line_all = cur_all + (kTPB / 32) * nb8;
	uint8_t* tile_raw = reinterpret_cast<uint8_t*>(line_all + (size_t)(kTPB / 32) * nb8 * 8);
	const TileSmem sm = carve_smem(tile_raw, P.k, SPACED);
	WarpBins wb;
	wb.cursor = cur_all + warp * nb8;
	wb.line = line_all + (size_t)warp * nb8 * 8;
	wb.cursor_sa = (uint32_t)__cvta_generic_to_shared(wb.cursor);
	wb.line_sa = (uint32_t)__cvta_generic_to_shared(wb.line);
	wb.writer = blockIdx.x * (kTPB / 32) + warp;
	for (uint32_t i = lane; i < nb; i += 32)
		wb.cursor[i] = 0;
	__syncwarp();

	const uint64_t tiles = (P.n_windows + kTile - 1) / kTile;
	for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
		const uint64_t t0 = t * kTile;
		__syncthreads(); // the previous tile is fully consumed before its staging area is overwritten
		tile_phase_a(P, sm, t0, tid, kTPB);
		__syncthreads();
		tile_phase_b(P, sm, t0, tid, kTPB);
		__syncthreads();
		uint32_t validw = 0;
		const uint32_t p0 = (uint32_t)tid * kWPT;
		roll_windows(P, sm, t0, tid, [&](uint32_t s, bool ok, uint64_t F, uint64_t RC) {
			validw |= (uint32_t)ok << s;
			const uint32_t wid = (uint32_t)t0 + p0 + s;
			for_each_hash<SPACED>(P, sm, p0 + s, F, RC, [&](uint32_t, uint64_t hv, bool) {
				warp_bin_emit<QUERY>(P, wb, fastmod<POW2>(hv, P.fm), wid, ok);
				return true;
			});
		});
		uint64_t widx = (t0 >> 5) + tid;
		if (P.valid_bits && widx < P.out_words)
			P.valid_bits[widx] = validw;
		if (P.stats) {
			uint32_t nv = __reduce_add_sync(kFullMask, __popc(validw));
			if (lane == 0 && nv)
				atomicAdd((unsigned long long*)&P.stats[0], (unsigned long long)nv);
		}
	}
	// drain the partial lines and publish the cursors
	__syncwarp();
	for (uint32_t part = lane; part < nb; part += 32) {
		const uint32_t c = wb.cursor[part] >> kTagBits, start = c & ~(L - 1);
		const uint64_t item0 = ((uint64_t)part * P.bin_writers + wb.writer) * P.bin_cap + start;
		for (uint32_t i = 0; i < (c & (L - 1)); i++) {
			if (QUERY) {
				uint2 it = *reinterpret_cast<const uint2*>(wb.line + part * 8 + i * 2);
				if (start < P.bin_cap)
					*reinterpret_cast<uint2*>(P.bin_items + (item0 + i) * 2) = it;
				else
					bin_direct_probe(P, part, it.x, it.y);
			} else {
				uint32_t off = wb.line[part * 8 + i];
				if (start < P.bin_cap)
					P.bin_items[item0 + i] = off;
				else
					bin_direct_or(P, part, off);
			}
		}
		P.bin_counts[(uint64_t)part * P.bin_writers + wb.writer] = c;
	}
}

size_t bin_warp_smem_bytes(uint32_t k, bool spaced, uint32_t n_bins)
{
	uint32_t nb8 = (n_bins + 7u) & ~7u;
	return (size_t)(kTPB / 32) * nb8 * 36 + tile_smem_bytes(k, spaced) + 32;
}

// Fallback for filters with more than kMaxWarpBins partitions: CTA-shared cursors bumped with
// shared-memory atomics, offsets stored one by one (window_op<OP_BF_BIN>).
template<bool SPACED, bool POW2>
__global__ void __launch_bounds__(kTPB) bin_kernel_cta(const __grid_constant__ SeqParams P)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	TileSmem sm = carve_smem(smem_raw, P.k, SPACED, P.n_bins);
	sm.writer = blockIdx.x;
	for (uint32_t i = threadIdx.x; i < P.n_bins; i += kTPB)
		sm.cursors[i] = 0;
	const uint64_t tiles = (P.n_windows + kTile - 1) / kTile;
	for (uint64_t t = blockIdx.x; t < tiles; t += gridDim.x) {
		__syncthreads(); // previous tile fully consumed (and the cursors initialised) before restaging
		run_tile<OP_BF_BIN, SPACED, POW2>(P, sm, t * kTile, threadIdx.x);
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < P.n_bins; i += kTPB)
		P.bin_counts[(uint64_t)i * P.bin_writers + blockIdx.x] = sm.cursors[i];
}

template<bool QUERY>
static const void* warp_fn(bool spaced, bool pow2)
{
	if (spaced)
		return pow2 ? (const void*)bin_kernel_warp<true, true, QUERY> : (const void*)bin_kernel_warp<true, false, QUERY>;
	return pow2 ? (const void*)bin_kernel_warp<false, true, QUERY> : (const void*)bin_kernel_warp<false, false, QUERY>;
}

static const void* cta_fn(bool spaced, bool pow2)
{
	if (spaced)
		return pow2 ? (const void*)bin_kernel_cta<true, true> : (const void*)bin_kernel_cta<true, false>;
	return pow2 ? (const void*)bin_kernel_cta<false, true> : (const void*)bin_kernel_cta<false, false>;
}

const void* bin_warp_kernel(bool query, bool spaced, bool pow2)
{
	return query ? warp_fn<true>(spaced, pow2) : warp_fn<false>(spaced, pow2);
}

const void* bin_cta_kernel(bool spaced, bool pow2)
{
	return cta_fn(spaced, pow2);
}

} // namespace btl
