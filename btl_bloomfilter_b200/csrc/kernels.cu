// kernels.cu -- sm_100a kernels of the btl_bloomfilter k-mer hot path.
//
//   seq_kernel<OP,SPACED,POW2>   fused  stage -> classify/pack -> roll ntHash -> multi-hash -> filter op
//                                (K1 bf_insert, K2 bf_contains, K3/K4 counting ops, K5 spaced seeds, K6 hash)
//   cbf_list_kernel              residual rounds of the exact (reference-order) counting insert
//   popcount_kernel (K8), merge_kernel (K7 local step), synth_*_kernel, random_probe_kernel
//
// Memory behaviour: the input is read once with coalesced 16-byte loads and lives in
// shared memory as 2-bit codes; the filter traffic is one random 32-byte sector per hash (gather for
// queries, RED.OR / byte update for inserts), so the kernels are bound by HBM random-sector rate.
#include "sort_bin.cuh"
#include "seq_kernel.cuh"

#include <cooperative_groups.h>

namespace btl {

// ---------------------------------------------------------------- partitioned build, pass 2
// Blocks are ordered partition-major, so at any moment the whole GPU is ORing into one or two
// 2^bin_shift-bit regions of the filter, which stay resident in L2 (one HBM read and one write-back per
// line instead of one 128-byte fetch per random bit).  One warp drains one sub-bucket at a time.
#ifndef BTL_APPLY_THREADS
#define BTL_APPLY_THREADS 256
#endif
constexpr int kApplyThreads = BTL_APPLY_THREADS;
__global__ void __launch_bounds__(kApplyThreads) apply_bins_kernel(const __grid_constant__ SeqParams P, uint32_t blocks_per_part)
{
	const uint32_t part = P.bin_part0 + blockIdx.x / blocks_per_part, sub = blockIdx.x % blocks_per_part;
	uint32_t* region = (uint32_t*)P.filter + ((uint64_t)part << (P.bin_shift - 5));
	// Software pipeline: while partition `part` is being updated, pull the next partition's lines into L2
	// with full-line prefetches, so that its atomics hit in L2 instead of waiting on one HBM sector each.
	// (bin_prefetch == 2: this CTA's share of its OWN partition instead -- full 128-byte lines ahead of the demand
	// misses, without a second partition in L2; for partitions too large to keep two of them resident)
	const uint32_t pf_part = part + (P.bin_prefetch == 2 ? 0u : 1u);
	if (P.bin_prefetch && pf_part < P.n_bins) {
		const uint64_t next_bit0 = (uint64_t)pf_part << P.bin_shift;
		uint64_t bits = P.fm.m - next_bit0;
		if (bits > ((uint64_t)1 << P.bin_shift))
			bits = (uint64_t)1 << P.bin_shift;
		const uint64_t lines = (bits + 1023) >> 10; // 128-byte lines
		const char* base = (const char*)P.filter + (next_bit0 >> 3);
		for (uint64_t l = (uint64_t)sub * kApplyThreads + threadIdx.x; l < lines; l += (uint64_t)blocks_per_part * kApplyThreads)
			asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (l << 7)));
	}
	// work unit = one segment of one sub-bucket (bin_segs segments each, so that a partition keeps the whole
	// GPU busy even when there are only a few hundred writers)
	const uint32_t warps = kApplyThreads / 32, lane = threadIdx.x & 31;
	const uint32_t units = P.bin_writers * P.bin_segs;
	for (uint32_t u = sub * warps + (threadIdx.x >> 5); u < units; u += blocks_per_part * warps) {
		const uint32_t w = u / P.bin_segs, sg = u - w * P.bin_segs;
		uint32_t n = __ldg(P.bin_counts + (uint64_t)part * P.bin_writers + w);
		n = n < P.bin_cap ? n : P.bin_cap;
		const uint32_t seg = ((n + P.bin_segs - 1) / P.bin_segs + 3u) & ~3u; // whole 16-byte vectors
		const uint32_t lo = sg * seg;
		if (lo >= n)
			continue;
		n = n - lo < seg ? n - lo : seg;
		const uint32_t* items = P.bin_items + ((uint64_t)part * P.bin_writers + w) * P.bin_cap + lo;
		const uint4* v = reinterpret_cast<const uint4*>(items); // bin_cap and lo are multiples of 4
		const uint32_t nv = n / 4;
		for (uint32_t i = lane; i < nv; i += 32) {
			uint4 x = __ldcs(v + i);
			atomicOr(region + (x.x >> 5), 1u << (x.x & 31));
			atomicOr(region + (x.y >> 5), 1u << (x.y & 31));
			atomicOr(region + (x.z >> 5), 1u << (x.z & 31));
			atomicOr(region + (x.w >> 5), 1u << (x.w & 31));
		}
		for (uint32_t i = nv * 4 + lane; i < n; i += 32) {
			uint32_t o = __ldcs(items + i);
			atomicOr(region + (o >> 5), 1u << (o & 31));
		}
	}
}

// Partitioned query, pass 2: same schedule; every (offset, window) pair tests its bit in the L2-resident
// region and a miss clears the window's hit bit (the hit words start as all ones and are ANDed with the
// valid words by finalize_hits_kernel).  COUNTING: the region holds 8-bit counters and the test is
// counter >= threshold (CountingBloomFilter.hpp:190-196: contains == minCount >= threshold, and a minimum is
// >= t exactly when every counter is).
// ld_mode (experiment knob "probe_ld"): 0 = read-only path (ld.global.nc), 1 = ld.global.cg, 2 = L1::no_allocate
template<bool COUNTING>
__device__ __forceinline__ uint32_t probe_load(const void* region, uint32_t off, uint32_t ld_mode)
{
	if (COUNTING) {
		const uint8_t* a = reinterpret_cast<const uint8_t*>(region) + off;
		if (ld_mode == 1)
			return __ldcg(a);
		if (ld_mode == 2) {
			uint32_t v;
			asm volatile("ld.global.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(a));
			return v;
		}
		return __ldg(a);
	}
	const uint32_t* a = reinterpret_cast<const uint32_t*>(region) + (off >> 5);
	if (ld_mode == 1)
		return __ldcg(a);
	if (ld_mode == 2) {
		uint32_t v;
		asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(a));
		return v;
	}
	return __ldg(a);
}

template<bool COUNTING>
__device__ __forceinline__ void probe_test(uint32_t v, uint32_t threshold, uint32_t* hit, uint32_t off, uint32_t wid)
{
	const bool ok = COUNTING ? v >= threshold : ((v >> (off & 31)) & 1u) != 0;
	if (!ok)
		atomicAnd(hit + (wid >> 5), ~(1u << (wid & 31)));
}

// The kernel may share the SMs with pass 1 of the next sub-batch (binned_query runs the two passes of successive
// sub-batches on two streams), which leaves it a fraction of the thread slots: every thread therefore keeps
// 2 * UNROLL probes in flight.
template<bool COUNTING, int UNROLL>
__global__ void __launch_bounds__(kApplyThreads) probe_bins_kernel(const __grid_constant__ SeqParams P, uint32_t blocks_per_part)
{
	if (P.gate && *P.gate != P.gate_want)
		return;
	const uint32_t part = blockIdx.x / blocks_per_part, sub = blockIdx.x % blocks_per_part;
	const void* region = COUNTING ? (const void*)((const uint8_t*)P.filter + ((uint64_t)part << P.bin_shift))
	                              : (const void*)((const uint32_t*)P.filter + ((uint64_t)part << (P.bin_shift - 5)));
	const uint32_t pf_part = part + (P.bin_prefetch == 2 ? 0u : 1u); // 2: the CTA's share of its own partition (see apply_bins_kernel)
	if (P.bin_prefetch && pf_part < P.n_bins) {
		const uint64_t next0 = (uint64_t)pf_part << P.bin_shift; // bits, or counters
		uint64_t units = P.fm.m - next0;
		if (units > ((uint64_t)1 << P.bin_shift))
			units = (uint64_t)1 << P.bin_shift;
		const uint64_t lines = COUNTING ? (units + 127) >> 7 : (units + 1023) >> 10;
		const char* base = (const char*)P.filter + (COUNTING ? next0 : next0 >> 3);
		for (uint64_t l = (uint64_t)sub * kApplyThreads + threadIdx.x; l < lines; l += (uint64_t)blocks_per_part * kApplyThreads)
			asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (l << 7)));
	}
	const uint32_t warps = kApplyThreads / 32, lane = threadIdx.x & 31;
	const uint32_t units = P.bin_writers * P.bin_segs;
	for (uint32_t u = sub * warps + (threadIdx.x >> 5); u < units; u += blocks_per_part * warps) {
		const uint32_t w = u / P.bin_segs, sg = u - w * P.bin_segs;
		uint32_t n = __ldg(P.bin_counts + (uint64_t)part * P.bin_writers + w);
		n = n < P.bin_cap ? n : P.bin_cap;
		const uint32_t seg = ((n + P.bin_segs - 1) / P.bin_segs + 3u) & ~3u;
		const uint32_t lo = sg * seg;
		if (lo >= n)
			continue;
		n = n - lo < seg ? n - lo : seg;
		const uint32_t* items = P.bin_items + (((uint64_t)part * P.bin_writers + w) * P.bin_cap + lo) * 2;
		const uint4* v = reinterpret_cast<const uint4*>(items); // two items per 16 bytes
		const uint32_t nv = n / 2;
		for (uint32_t i0 = lane; i0 < nv; i0 += 32 * UNROLL) {
			uint4 x[UNROLL];
			uint32_t a[UNROLL], b[UNROLL];
#pragma unroll
			for (int j = 0; j < UNROLL; j++) {
				const uint32_t i = i0 + 32 * j;
				// (a slot past the end re-reads the last vector: harmless, its result is not used)
				x[j] = __ldcs(v + (i < nv ? i : nv - 1));
			}
#pragma unroll
			for (int j = 0; j < UNROLL; j++) {
				a[j] = probe_load<COUNTING>(region, x[j].x, P.probe_ld);
				b[j] = probe_load<COUNTING>(region, x[j].z, P.probe_ld);
			}
#pragma unroll
			for (int j = 0; j < UNROLL; j++) {
				if (i0 + 32 * j < nv) {
					probe_test<COUNTING>(a[j], P.threshold, P.hit_bits, x[j].x, x[j].y);
					probe_test<COUNTING>(b[j], P.threshold, P.hit_bits, x[j].z, x[j].w);
				}
			}
		}
		if ((n & 1u) && lane == 0) {
			uint2 x = __ldcs(reinterpret_cast<const uint2*>(items) + (n - 1));
			probe_test<COUNTING>(probe_load<COUNTING>(region, x.x, P.probe_ld), P.threshold, P.hit_bits, x.x, x.y);
		}
	}
}

// hit &= valid, and the number of hits is added to stats[1]
__global__ void __launch_bounds__(256) finalize_hits_kernel(uint32_t* hit, const uint32_t* valid, uint64_t n_words,
                                                            unsigned long long* hits_out, const uint32_t* gate,
                                                            uint32_t gate_want)
{
	if (gate && *gate != gate_want)
		return;
	unsigned long long acc = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t h = hit[i] & valid[i];
		hit[i] = h;
		acc += __popc(h);
	}
	for (int o = 16; o > 0; o >>= 1)
		acc += __shfl_xor_sync(0xffffffffu, acc, o);
	if ((threadIdx.x & 31) == 0 && acc && hits_out)
		atomicAdd(hits_out, acc);
}

// ---------------------------------------------------------------- pass-1 kernel selection
// BIN_SORT: bin_kernel_sort (sort_bin.cuh); BIN_WARP / BIN_CTA: the general-shape kernels above.
struct BinKernel
{
	const void* fn;
	size_t smem;
	int threads;
	int mode;
	uint32_t tile; // windows per CTA pass
};

// CTA size of the sort-bin kernel for this shape: 256 threads when a round (threads * W * h items) still gives
// every partition a run of ~14 items, else 512; 0 when the shape is not served at all.
static int sort_threads_for(const SeqParams& P, uint32_t n_bins)
{
	if (P.bin_legacy || P.h > (uint32_t)kMaxSortHashes || (P.n_seeds && P.h2 != 1))
		return 0;
	const bool spaced = P.n_seeds != 0;
	for (int threads : { 256, 512 }) {
		if (n_bins > sort_max_bins(threads))
			continue;
		if (sort_smem_bytes(P.k, spaced, n_bins, (int)P.h, threads) > (size_t)(220 / sort_ctas_per_sm(threads)) * 1024)
			continue;
		const uint64_t round_items = (uint64_t)threads * sort_round_windows((int)P.h) * P.h;
		if (threads == 256 && round_items < (uint64_t)14 * n_bins)
			continue;
		return threads;
	}
	return 0;
}

bool bin_sort_eligible(const SeqParams& P, uint32_t n_bins)
{
	return sort_threads_for(P, n_bins) != 0;
}

static cudaError_t bin_select(const SeqParams& P, uint32_t n_bins, bool query, BinKernel* K, int* occ)
{
	const bool spaced = P.n_seeds != 0, pow2 = P.fm.pow2 != 0;
	if (const int threads = sort_threads_for(P, n_bins)) {
		K->mode = BIN_SORT;
		K->threads = threads;
		K->tile = sort_tile(threads);
		K->smem = sort_smem_bytes(P.k, spaced, n_bins, (int)P.h, threads);
		if (threads == 256)
			K->fn = query ? bin_sort_kernel_query_256((int)P.h, spaced, pow2) : bin_sort_kernel_build_256((int)P.h, spaced, pow2);
		else
			K->fn = query ? bin_sort_kernel_query_512((int)P.h, spaced, pow2) : bin_sort_kernel_build_512((int)P.h, spaced, pow2);
	} else if (n_bins <= kMaxWarpBins) {
		K->mode = BIN_WARP;
		K->threads = kTPB;
		K->tile = kTile;
		K->smem = bin_warp_smem_bytes(P.k, spaced, n_bins);
		K->fn = bin_warp_kernel(query, spaced, pow2);
	} else {
		if (query)
			return cudaErrorNotSupported;
		K->mode = BIN_CTA;
		K->threads = kTPB;
		K->tile = kTile;
		K->smem = tile_smem_bytes(P.k, spaced, n_bins);
		K->fn = bin_cta_kernel(spaced, pow2);
	}
	if (K->smem > 48 * 1024) {
		cudaError_t e = cudaFuncSetAttribute(K->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)K->smem);
		if (e != cudaSuccess)
			return e;
	}
	// same L1/shared split as the pass-2 kernels, so that the two passes can share an SM (kernels that ask
	// for different carve-outs cannot be co-resident)
	cudaFuncSetAttribute(K->fn, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
	return cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, K->fn, K->threads, K->smem);
}

uint32_t bin_sort_tile(const SeqParams& P, uint32_t n_bins)
{
	const int threads = sort_threads_for(P, n_bins);
	return threads ? sort_tile(threads) : (uint32_t)kTile;
}

bool bin_query_supported(const SeqParams& P, uint32_t n_bins)
{
	return n_bins <= kMaxWarpBins || bin_sort_eligible(P, n_bins);
}

// number of sub-bucket writers (CTAs or warps) the bin kernel will run with, its grid size and flavour.
// BIN_SORT: the writer count does not depend on the batch (so that launches can append to the same
// sub-buckets); only the first `grid` writers run when the batch has fewer tiles than that.
cudaError_t bin_plan(const SeqParams& P, uint32_t n_bins, bool query, uint32_t* writers, uint32_t* grid, int* mode)
{
	BinKernel K;
	int occ = 0, dev = 0, sms = 0;
	cudaError_t e = bin_select(P, n_bins, query, &K, &occ);
	if (e != cudaSuccess)
		return e;
	if ((e = cudaGetDevice(&dev)) != cudaSuccess)
		return e;
	if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
		return e;
	if (occ < 1)
		return cudaErrorLaunchOutOfResources;
	if (P.bin_ctas_per_sm && (int)P.bin_ctas_per_sm < occ)
		occ = (int)P.bin_ctas_per_sm;
	uint64_t tiles = (P.n_windows + K.tile - 1) / K.tile;
	uint64_t full = (uint64_t)sms * (uint64_t)occ;
	uint64_t g = tiles < full ? (tiles ? tiles : 1) : full;
	*grid = (uint32_t)g;
	*mode = K.mode;
	*writers = (uint32_t)(K.mode == BIN_SORT ? full : K.mode == BIN_WARP ? g * (kTPB / 32) : g);
	return cudaSuccess;
}

cudaError_t launch_bin(const SeqParams& P, bool query, uint32_t grid, cudaStream_t stream)
{
	BinKernel K;
	int occ = 0;
	cudaError_t e = bin_select(P, P.n_bins, query, &K, &occ);
	if (e != cudaSuccess)
		return e;
	void* args[1] = { (void*)&P };
	return cudaLaunchKernel(K.fn, dim3(grid), dim3((unsigned)K.threads), args, K.smem, stream);
}

// number of SMs of the current device (queried once per device)
static uint32_t sm_count()
{
	static int cached[64] = { 0 };
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64)
		return 148;
	if (!cached[dev]) {
		int n = 0;
		if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n < 1)
			n = 148;
		cached[dev] = n;
	}
	return (uint32_t)cached[dev];
}

static uint32_t blocks_per_partition(const SeqParams& P)
{
	// enough blocks per partition to fill the GPU (4 per SM), never more warps than sub-buckets
	const uint32_t full = 4u * sm_count();
	uint32_t want = (P.bin_writers * P.bin_segs + (kApplyThreads / 32) - 1) / (kApplyThreads / 32);
	return want < full ? (want ? want : 1u) : full;
}

cudaError_t launch_probe_bins(const SeqParams& P, bool counting, int unroll, bool maxshared, cudaStream_t stream)
{
	uint32_t bpp = blocks_per_partition(P);
	uint64_t grid = (uint64_t)P.n_bins * bpp;
	if (grid == 0)
		return cudaSuccess;
	if (grid > 0x7fffffffULL)
		return cudaErrorInvalidValue;
	const void* fn;
	if (counting)
		fn = unroll >= 4 ? (const void*)probe_bins_kernel<true, 4> : unroll >= 2 ? (const void*)probe_bins_kernel<true, 2>
		                                                                        : (const void*)probe_bins_kernel<true, 1>;
	else
		fn = unroll >= 4 ? (const void*)probe_bins_kernel<false, 4> : unroll >= 2 ? (const void*)probe_bins_kernel<false, 2>
		                                                                         : (const void*)probe_bins_kernel<false, 1>;
	// The gathers need the large L1 (every outstanding miss holds a line there: with the small one the kernel runs
	// at half speed, measured).  maxshared != 0 (overlapped sub-batches): the pass-1 kernel's split instead, without
	// which the two kernels cannot be co-resident at all.
	cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout,
	                     maxshared ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault);
	void* args[2] = { (void*)&P, (void*)&bpp };
	return cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(kApplyThreads), args, 0, stream);
}

// adaptive query: picks the path from the hit fraction of a sample of the batch, on the device
__global__ void query_gate_kernel(const unsigned long long* sample_stats, uint32_t* flag, uint32_t pct)
{
	const unsigned long long valid = sample_stats[0], hits = sample_stats[1];
	*flag = (valid > 0 && hits * 100ull < valid * (unsigned long long)pct) ? 1u : 0u;
}

cudaError_t launch_query_gate(const unsigned long long* sample_stats, uint32_t* flag, uint32_t pct, cudaStream_t stream)
{
	query_gate_kernel<<<1, 1, 0, stream>>>(sample_stats, flag, pct);
	return cudaGetLastError();
}

cudaError_t launch_finalize_hits(uint32_t* hit, const uint32_t* valid, uint64_t n_words, unsigned long long* hits_out,
                                 const uint32_t* gate, uint32_t gate_want, cudaStream_t stream)
{
	if (n_words == 0)
		return cudaSuccess;
	uint64_t want = (n_words + 255) / 256;
	unsigned grid = (unsigned)(want > sm_count() * 8 ? sm_count() * 8 : want);
	finalize_hits_kernel<<<grid, 256, 0, stream>>>(hit, valid, n_words, hits_out, gate, gate_want);
	return cudaGetLastError();
}

cudaError_t launch_apply_bins(const SeqParams& P, cudaStream_t stream)
{
	uint32_t bpp = blocks_per_partition(P);
	uint64_t grid = (uint64_t)(P.bin_part_count ? P.bin_part_count : P.n_bins) * bpp;
	if (grid == 0)
		return cudaSuccess;
	if (grid > 0x7fffffffULL)
		return cudaErrorInvalidValue;
	cudaFuncSetAttribute(apply_bins_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
	apply_bins_kernel<<<(unsigned)grid, kApplyThreads, 0, stream>>>(P, bpp);
	return cudaGetLastError();
}

size_t seq_kernel_smem_bytes(uint32_t k, bool spaced)
{
	return tile_smem_bytes(k, spaced);
}

cudaError_t launch_seq(SeqOp op, const SeqParams& P, cudaStream_t stream)
{
	switch (op) {
	case OP_HASH:
	case OP_BF_INSERT:
	case OP_BF_CONTAINS: return launch_seq_bloom(op, P, stream); // seq_ops_bloom.cu
	case OP_CBF_MINCOUNT: return launch_op<OP_CBF_MINCOUNT>(P, stream);
	case OP_CBF_INCALL: return launch_op<OP_CBF_INCALL>(P, stream);
	case OP_RESV_TOUCH: return launch_op<OP_RESV_TOUCH>(P, stream);
	case OP_CBF_COMMIT: return launch_seq_cbf_commit(P, stream);   // seq_commit_cbf.cu
	case OP_RESV_CLEAR: return launch_op<OP_RESV_CLEAR>(P, stream);
	case OP_BFCHK_COMMIT: return launch_seq_bfchk_commit(P, stream); // seq_commit_bfchk.cu
	}
	return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------- ordered updates, residual rounds
template<bool POW2, int KIND>
__global__ void __launch_bounds__(128) list_round_kernel(int phase, const __grid_constant__ SeqParams P,
                                                         const __grid_constant__ ListParams L)
{
	uint32_t item = blockIdx.x * blockDim.x + threadIdx.x;
	if (item >= *L.count_in)
		return;
	uint32_t w = L.list_in[item];
	uint64_t emask = ((uint64_t)1 << L.resv_log2) - 1;
	if (phase == 0)
		list_round_reserve<POW2>(P, L.resv, emask, L.epoch, w);
	else if (!list_round_commit<POW2, KIND>(P, L.resv, emask, L.epoch, w))
		L.list_out[atomicAdd(L.count_out, 1u)] = w;
}

cudaError_t launch_list_round(int phase, const SeqParams& P, const ListParams& L, cudaStream_t stream)
{
	if (L.max_items == 0)
		return cudaSuccess;
	unsigned grid = (L.max_items + 127) / 128;
	if (P.fm.pow2) {
		if (L.kind == 0)
			list_round_kernel<true, 0><<<grid, 128, 0, stream>>>(phase, P, L);
		else
			list_round_kernel<true, 1><<<grid, 128, 0, stream>>>(phase, P, L);
	} else {
		if (L.kind == 0)
			list_round_kernel<false, 0><<<grid, 128, 0, stream>>>(phase, P, L);
		else
			list_round_kernel<false, 1><<<grid, 128, 0, stream>>>(phase, P, L);
	}
	return cudaGetLastError();
}

constexpr int kDrainThreads = 1024;

template<bool POW2, int KIND>
__global__ void __launch_bounds__(kDrainThreads) list_drain_kernel(const __grid_constant__ SeqParams P,
                                                                    const __grid_constant__ ListParams L)
{
	__shared__ uint32_t s_out;
	const uint32_t* in = L.list_in;
	uint32_t* out = L.list_out;
	uint32_t* other = const_cast<uint32_t*>(L.list_in);
	uint32_t n = *L.count_in;
	uint32_t epoch = L.epoch, rounds = 0;
	uint64_t emask = ((uint64_t)1 << L.resv_log2) - 1;
	while (n > 0) {
		if (threadIdx.x == 0)
			s_out = 0;
		for (uint32_t i = threadIdx.x; i < n; i += kDrainThreads)
			list_round_reserve<POW2>(P, L.resv, emask, epoch, in[i]);
		__syncthreads();
		for (uint32_t i = threadIdx.x; i < n; i += kDrainThreads) {
			uint32_t w = in[i];
			if (!list_round_commit<POW2, KIND>(P, L.resv, emask, epoch, w))
				out[atomicAdd(&s_out, 1u)] = w;
		}
		__syncthreads();
		n = s_out;
		const uint32_t* t = out;
		out = other;
		in = t;
		other = const_cast<uint32_t*>(t);
		epoch++;
		rounds++;
		__syncthreads();
	}
	if (threadIdx.x == 0) {
		*L.count_out = 0;
		if (L.rounds_out)
			*L.rounds_out = rounds;
	}
}

// Cooperative variant: every round is [reserve | grid barrier | commit or re-queue | grid barrier].
// L.counts[0] is the length of list_in on entry; the epoch counter is read from and written back to
// L.d_epoch, and the reservation table is re-armed here when the 32-bit epoch space runs low.
constexpr int kCoopThreads = 256;

template<bool POW2, int KIND>
__global__ void __launch_bounds__(kCoopThreads) list_drain_coop_kernel(const __grid_constant__ SeqParams P,
                                                                       const __grid_constant__ ListParams L)
{
	cooperative_groups::grid_group grid = cooperative_groups::this_grid();
	const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
	const uint64_t emask = ((uint64_t)1 << L.resv_log2) - 1;
	uint32_t epoch = *L.d_epoch;
	uint32_t n = L.counts[0];
	if (n == 0)
		return; // uniform across the grid
	if (gtid == 0)
		L.counts[4] += n; // statistics: k-mers that had to wait for the residual rounds
	if (epoch > 0xf0000000u) {
		for (uint64_t i = gtid; i <= emask; i += gthreads)
			L.resv[i] = ~0ull;
		epoch = 0;
		grid.sync();
	}
	const uint32_t* in = L.list_in;
	uint32_t* out = L.list_out;
	uint32_t cur = 0, rounds = 0;
	while (n > 0) {
		++epoch;
		for (uint32_t i = gtid; i < n; i += gthreads)
			list_round_reserve<POW2>(P, L.resv, emask, epoch, in[i]);
		grid.sync();
		for (uint32_t i = gtid; i < n; i += gthreads) {
			uint32_t w = in[i];
			if (!list_round_commit<POW2, KIND>(P, L.resv, emask, epoch, w))
				out[atomicAdd(L.counts + (cur ^ 1u), 1u)] = w;
		}
		grid.sync();
		n = ld_cg(L.counts + (cur ^ 1u));
		if (gtid == 0)
			L.counts[cur] = 0; // becomes the output counter of the next round (ordered by its first barrier)
		const uint32_t* t = out;
		out = const_cast<uint32_t*>(in);
		in = t;
		cur ^= 1u;
		rounds++;
	}
	if (gtid == 0) {
		*L.d_epoch = epoch;
		L.counts[0] = 0;
		L.counts[1] = 0;
		if (L.rounds_out)
			atomicAdd(L.rounds_out, rounds);
	}
}

cudaError_t launch_list_drain_coop(const SeqParams& P, const ListParams& L, cudaStream_t stream)
{
	const void* kern;
	if (P.fm.pow2)
		kern = L.kind == 0 ? (const void*)list_drain_coop_kernel<true, 0> : (const void*)list_drain_coop_kernel<true, 1>;
	else
		kern = L.kind == 0 ? (const void*)list_drain_coop_kernel<false, 0> : (const void*)list_drain_coop_kernel<false, 1>;
	int occ = 0, dev = 0, sms = 0;
	cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kCoopThreads, 0);
	if (e != cudaSuccess)
		return e;
	if ((e = cudaGetDevice(&dev)) != cudaSuccess)
		return e;
	if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess)
		return e;
	if (occ < 1)
		return cudaErrorLaunchOutOfResources;
	int per_sm = occ < 2 ? occ : 2; // a few hundred thousand items at most: two CTAs per SM are plenty
	void* args[2] = { (void*)&P, (void*)&L };
	return cudaLaunchCooperativeKernel(kern, dim3((unsigned)(sms * per_sm)), dim3(kCoopThreads), args, 0, stream);
}

cudaError_t launch_list_drain(const SeqParams& P, const ListParams& L, cudaStream_t stream)
{
	if (P.fm.pow2) {
		if (L.kind == 0)
			list_drain_kernel<true, 0><<<1, kDrainThreads, 0, stream>>>(P, L);
		else
			list_drain_kernel<true, 1><<<1, kDrainThreads, 0, stream>>>(P, L);
	} else {
		if (L.kind == 0)
			list_drain_kernel<false, 0><<<1, kDrainThreads, 0, stream>>>(P, L);
		else
			list_drain_kernel<false, 1><<<1, kDrainThreads, 0, stream>>>(P, L);
	}
	return cudaGetLastError();
}

// ---------------------------------------------------------------- precomputed-hash (legacy per-k-mer) interface
__global__ void __launch_bounds__(128) hashes_kernel(int op, void* filter, FastMod fm, uint32_t h, uint32_t threshold,
                                                     const uint64_t* __restrict__ hashes, uint64_t n, uint8_t* out)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n)
		return;
	const uint64_t* hv = hashes + i * h;
	if (op == 0) {
		uint32_t* words = (uint32_t*)filter;
		for (uint32_t j = 0; j < h; j++) {
			uint64_t b = fm.pow2 ? fastmod<true>(hv[j], fm) : fastmod<false>(hv[j], fm);
			atomicOr(words + (b >> 5), 1u << (uint32_t)(b & 31));
		}
	} else if (op == 1) {
		const uint32_t* words = (const uint32_t*)filter;
		uint32_t hit = 1;
		for (uint32_t j = 0; j < h && hit; j++) {
			uint64_t b = fm.pow2 ? fastmod<true>(hv[j], fm) : fastmod<false>(hv[j], fm);
			hit = (__ldcg(words + (b >> 5)) >> (uint32_t)(b & 31)) & 1u;
		}
		out[i] = (uint8_t)hit;
	} else if (op == 2) {
		const uint8_t* cnt = (const uint8_t*)filter;
		uint32_t mn = 255;
		for (uint32_t j = 0; j < h; j++) {
			uint32_t v = __ldcg(cnt + (fm.pow2 ? fastmod<true>(hv[j], fm) : fastmod<false>(hv[j], fm)));
			mn = v < mn ? v : mn;
		}
		out[i] = (uint8_t)mn;
	} else if (op == 4) {
		for (uint32_t j = 0; j < h; j++)
			counter_sat_inc((uint8_t*)filter, fm.pow2 ? fastmod<true>(hv[j], fm) : fastmod<false>(hv[j], fm));
	}
}

__global__ void hashes_serial_kernel(int op, void* filter, FastMod fm, uint32_t h, uint32_t threshold,
                                     const uint64_t* __restrict__ hashes, uint64_t n, uint8_t* out)
{
	if (blockIdx.x != 0 || threadIdx.x != 0)
		return;
	for (uint64_t i = 0; i < n; i++) {
		const uint64_t* hv = hashes + i * h;
		if (op == 3) {
			volatile uint8_t* cnt = (volatile uint8_t*)filter;
			uint32_t mn = 255;
			for (uint32_t j = 0; j < h; j++) {
				uint32_t v = cnt[fm.pow2 ? fastmod<true>(hv[j], fm) : fastmod<false>(hv[j], fm)];
				mn = v < mn ? v : mn;
			}
			if (out)
				out[i] = mn >= threshold;
			if (mn == 255)
				continue;
			for (uint32_t j = 0; j < h; j++) {
				uint64_t p = fm.pow2 ? fastmod<true>(hv[j], fm) : fastmod<false>(hv[j], fm);
				if (cnt[p] == mn)
					cnt[p] = (uint8_t)(mn + 1);
			}
		} else {
			volatile uint32_t* words = (volatile uint32_t*)filter;
			uint32_t found = 1;
			for (uint32_t j = 0; j < h; j++) {
				uint64_t b = fm.pow2 ? fastmod<true>(hv[j], fm) : fastmod<false>(hv[j], fm);
				uint32_t bit = 1u << (uint32_t)(b & 31);
				uint32_t old = words[b >> 5];
				found &= (old & bit) != 0;
				words[b >> 5] = old | bit;
			}
			if (out)
				out[i] = (uint8_t)found;
		}
	}
}

cudaError_t launch_hashes_op(int op, void* filter, FastMod fm, uint32_t h, uint32_t threshold,
                             const uint64_t* d_hashes, uint64_t n, uint8_t* d_out, cudaStream_t stream)
{
	if (n == 0)
		return cudaSuccess;
	if (op == 3 || op == 5)
		hashes_serial_kernel<<<1, 32, 0, stream>>>(op, filter, fm, h, threshold, d_hashes, n, d_out);
	else
		hashes_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(op, filter, fm, h, threshold, d_hashes, n, d_out);
	return cudaGetLastError();
}

// ---------------------------------------------------------------- K8: popcount / count_nonzero / count_ge
// mode 0: set bits, 1: non-zero bytes, 2: bytes >= threshold
__global__ void __launch_bounds__(256) popcount_kernel(const uint8_t* __restrict__ data, uint64_t nbytes, int mode,
                                                       unsigned threshold, unsigned long long* out)
{
	uint64_t nvec = nbytes / 16;
	const uint4* v = reinterpret_cast<const uint4*>(data);
	uint32_t thr4 = (threshold & 255u) * 0x01010101u;
	unsigned long long acc = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
		uint4 x = __ldg(v + i);
		if (mode == 0)
			acc += __popc(x.x) + __popc(x.y) + __popc(x.z) + __popc(x.w);
		else if (mode == 1)
			acc += (__popc(__vcmpne4(x.x, 0)) + __popc(__vcmpne4(x.y, 0)) + __popc(__vcmpne4(x.z, 0)) +
			        __popc(__vcmpne4(x.w, 0))) >> 3;
		else
			acc += (__popc(__vcmpgeu4(x.x, thr4)) + __popc(__vcmpgeu4(x.y, thr4)) +
			        __popc(__vcmpgeu4(x.z, thr4)) + __popc(__vcmpgeu4(x.w, thr4))) >> 3;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		for (uint64_t b = nvec * 16; b < nbytes; b++) {
			uint8_t c = data[b];
			acc += mode == 0 ? __popc((unsigned)c) : mode == 1 ? (c != 0) : (c >= threshold);
		}
	}
	for (int o = 16; o > 0; o >>= 1)
		acc += __shfl_xor_sync(0xffffffffu, acc, o);
	__shared__ unsigned long long red[8];
	if ((threadIdx.x & 31) == 0)
		red[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned long long s = 0;
		for (int w = 0; w < 8; w++)
			s += red[w];
		if (s)
			atomicAdd(out, s);
	}
}

cudaError_t launch_popcount(const void* data, uint64_t nbytes, int mode, unsigned threshold,
                            unsigned long long* d_out, cudaStream_t stream)
{
	uint64_t nvec = nbytes / 16;
	uint64_t want = (nvec + 255) / 256;
	unsigned grid = (unsigned)(want < 1 ? 1 : want > sm_count() * 16 ? sm_count() * 16 : want);
	popcount_kernel<<<grid, 256, 0, stream>>>((const uint8_t*)data, nbytes, mode, threshold, d_out);
	return cudaGetLastError();
}

// ---------------------------------------------------------------- K7 local step: dst |= src  /  dst = sat(dst + src)
__global__ void __launch_bounds__(256) merge_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src,
                                                    uint64_t nbytes, int sat_add)
{
	uint64_t nvec = nbytes / 16;
	uint4* d = reinterpret_cast<uint4*>(dst);
	const uint4* s = reinterpret_cast<const uint4*>(src);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (uint64_t)gridDim.x * blockDim.x) {
		uint4 a = d[i], b = s[i];
		if (sat_add) {
			a.x = __vaddus4(a.x, b.x); a.y = __vaddus4(a.y, b.y);
			a.z = __vaddus4(a.z, b.z); a.w = __vaddus4(a.w, b.w);
		} else {
			a.x |= b.x; a.y |= b.y; a.z |= b.z; a.w |= b.w;
		}
		d[i] = a;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) {
		for (uint64_t b = nvec * 16; b < nbytes; b++) {
			unsigned x = dst[b], y = src[b];
			dst[b] = sat_add ? (uint8_t)(x + y > 255 ? 255 : x + y) : (uint8_t)(x | y);
		}
	}
}

// ---------------------------------------------------------------- K7 fused: reduce-scatter + all-gather over peer memory
// One kernel per GPU replaces  all-to-all of 1/N slices -> local reduce -> all-gather:  this rank owns
// the byte range [lo, hi) of the filter; every 16-byte vector of it is loaded from all N partial filters
// (its own HBM and the peers' memory mapped over NVLink), reduced (OR / saturating add) in registers and
// stored into all N filters.  No staging buffer, one pass; the NVLink loads of one vector overlap the
// stores of the previous ones.  Ranges of different ranks are disjoint, so the kernels of all ranks run
// concurrently; the host brackets them with barriers.
template<int WORLD, int UNROLL>
__global__ void __launch_bounds__(256) peer_merge_kernel(const __grid_constant__ PeerMergeParams M)
{
	const uint64_t nvec = (M.hi - M.lo) / 16;
	const int world = WORLD ? WORLD : (int)M.world;
	// a CTA handles UNROLL consecutive groups of 256 vectors per iteration: UNROLL * WORLD 16-byte loads per
	// thread are in flight before the first one is consumed
	const uint64_t stride = (uint64_t)gridDim.x * blockDim.x * UNROLL;
	for (uint64_t i0 = (uint64_t)blockIdx.x * blockDim.x * UNROLL + threadIdx.x; i0 < nvec; i0 += stride) {
		uint4 v[UNROLL][WORLD ? WORLD : 1];
		if (WORLD) {
#pragma unroll
			for (int u = 0; u < UNROLL; u++) {
				const uint64_t i = i0 + (uint64_t)u * blockDim.x;
#pragma unroll
				for (int p = 0; p < WORLD; p++)
					v[u][p] = i < nvec && (p == 0 || M.mode != 2) ? *reinterpret_cast<const uint4*>(M.base[p] + M.lo + i * 16)
					                                             : make_uint4(0, 0, 0, 0);
			}
		}
#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
			const uint64_t i = i0 + (uint64_t)u * blockDim.x;
			if (i >= nvec)
				break;
			const uint64_t o = M.lo + i * 16;
			uint4 acc;
			if (WORLD) {
				acc = v[u][0];
#pragma unroll
				for (int p = 1; p < WORLD; p++) {
					if (M.sat_add) {
						acc.x = __vaddus4(acc.x, v[u][p].x); acc.y = __vaddus4(acc.y, v[u][p].y);
						acc.z = __vaddus4(acc.z, v[u][p].z); acc.w = __vaddus4(acc.w, v[u][p].w);
					} else {
						acc.x |= v[u][p].x; acc.y |= v[u][p].y; acc.z |= v[u][p].z; acc.w |= v[u][p].w;
					}
				}
			} else {
				acc = *reinterpret_cast<const uint4*>(M.base[0] + o);
				for (int p = 1; p < world; p++) {
					uint4 b = *reinterpret_cast<const uint4*>(M.base[p] + o);
					if (M.sat_add) {
						acc.x = __vaddus4(acc.x, b.x); acc.y = __vaddus4(acc.y, b.y);
						acc.z = __vaddus4(acc.z, b.z); acc.w = __vaddus4(acc.w, b.w);
					} else {
						acc.x |= b.x; acc.y |= b.y; acc.z |= b.z; acc.w |= b.w;
					}
				}
			}
			for (int p = 0; p < (M.mode == 1 ? 1 : world); p++)
				*reinterpret_cast<uint4*>(M.base[p] + o) = acc;
		}
	}
}

cudaError_t launch_peer_merge(const PeerMergeParams& M, cudaStream_t stream)
{
	if (M.world < 1 || M.world > (uint32_t)kMaxPeers || M.lo > M.hi || ((M.lo | M.hi) & 15u))
		return cudaErrorInvalidValue;
	uint64_t nvec = (M.hi - M.lo) / 16;
	if (nvec == 0)
		return cudaSuccess;
	const unsigned unroll = M.unroll == 2 || M.unroll == 4 ? M.unroll : 1;
	uint64_t want = (nvec + 256 * unroll - 1) / (256 * unroll);
	const uint64_t cap = M.grid ? M.grid : sm_count() * 8;
	unsigned grid = (unsigned)(want > cap ? cap : want);
#define BTL_PEER(W)                                                                  \
	do {                                                                             \
		if (unroll == 4) peer_merge_kernel<W, 4><<<grid, 256, 0, stream>>>(M);       \
		else if (unroll == 2) peer_merge_kernel<W, 2><<<grid, 256, 0, stream>>>(M);  \
		else peer_merge_kernel<W, 1><<<grid, 256, 0, stream>>>(M);                   \
	} while (0)
	switch (M.world) {
	case 2: BTL_PEER(2); break;
	case 4: BTL_PEER(4); break;
	case 8: BTL_PEER(8); break;
	default: peer_merge_kernel<0, 1><<<grid, 256, 0, stream>>>(M); break;
	}
#undef BTL_PEER
	return cudaGetLastError();
}

// ---------------------------------------------------------------- K7 in the switch: OR merge over an NVLS multicast mapping
// `mc` is a multicast address of the N partial filters (one replica per GPU, bound to one NVSwitch multicast
// object).  multimem.ld_reduce.or returns the OR of all N replicas of a word -- the reduction happens inside the
// switch, so this GPU receives 1/N of the filter instead of (N-1)/N -- and multimem.st writes the result to all N
// replicas (the switch fans it out).  Per NVLink direction a GPU moves ~|filter| bytes instead of the
// 2 (N-1)/N |filter| of the peer-memory kernel plus its own loads.  Bitwise reductions exist for .b32 / .b64
// only (vector forms are floating point), so a thread moves UNROLL independent 8-byte words per iteration.
// Saturating add has no multimem form: counting filters stay on peer_merge_kernel.
template<int UNROLL>
__device__ __forceinline__ void multimem_or_range(uint64_t* __restrict__ mc, uint64_t n_words, uint32_t cta, uint32_t n_ctas)
{
	const uint64_t stride = (uint64_t)n_ctas * blockDim.x * UNROLL;
	for (uint64_t i0 = (uint64_t)cta * blockDim.x * UNROLL + threadIdx.x; i0 < n_words; i0 += stride) {
		uint64_t v[UNROLL];
#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
			const uint64_t i = i0 + (uint64_t)u * blockDim.x;
			v[u] = 0;
			if (i < n_words)
				asm volatile("multimem.ld_reduce.relaxed.sys.global.or.b64 %0, [%1];" : "=l"(v[u]) : "l"(mc + i) : "memory");
		}
#pragma unroll
		for (int u = 0; u < UNROLL; u++) {
			const uint64_t i = i0 + (uint64_t)u * blockDim.x;
			if (i < n_words)
				asm volatile("multimem.st.relaxed.sys.global.b64 [%0], %1;" ::"l"(mc + i), "l"(v[u]) : "memory");
		}
	}
}

template<int UNROLL>
__global__ void __launch_bounds__(512) multimem_or_kernel(uint64_t* __restrict__ mc, uint64_t n_words)
{
	multimem_or_range<UNROLL>(mc, n_words, blockIdx.x, gridDim.x);
}

// Both mechanisms at once: the first mm_ctas CTAs reduce bytes [lo, mid) inside the switch (multimem), the others
// bytes [mid, hi) with peer loads and stores.  The two paths stress different parts of the fabric (the switch's
// reduction / multicast engines against plain link bandwidth), so running them side by side can finish sooner than
// either alone.  OR only.
template<int WORLD>
__global__ void __launch_bounds__(256) hybrid_merge_kernel(const __grid_constant__ PeerMergeParams M, uint64_t* __restrict__ mc,
                                                            uint64_t mid, uint32_t mm_ctas)
{
	if (blockIdx.x < mm_ctas) {
		multimem_or_range<4>(mc + M.lo / 8, (mid - M.lo) / 8, blockIdx.x, mm_ctas);
		return;
	}
	const uint32_t cta = blockIdx.x - mm_ctas, n_ctas = gridDim.x - mm_ctas;
	const uint64_t nvec = (M.hi - mid) / 16;
	const uint64_t stride = (uint64_t)n_ctas * blockDim.x;
	for (uint64_t i = (uint64_t)cta * blockDim.x + threadIdx.x; i < nvec; i += stride) {
		const uint64_t o = mid + i * 16;
		uint4 v[WORLD];
#pragma unroll
		for (int p = 0; p < WORLD; p++)
			v[p] = *reinterpret_cast<const uint4*>(M.base[p] + o);
		uint4 acc = v[0];
#pragma unroll
		for (int p = 1; p < WORLD; p++) {
			acc.x |= v[p].x; acc.y |= v[p].y; acc.z |= v[p].z; acc.w |= v[p].w;
		}
#pragma unroll
		for (int p = 0; p < WORLD; p++)
			*reinterpret_cast<uint4*>(M.base[p] + o) = acc;
	}
}

cudaError_t launch_hybrid_merge(const PeerMergeParams& M, void* mc_base, unsigned mm_pct, cudaStream_t stream)
{
	if (M.world < 2 || M.world > (uint32_t)kMaxPeers || M.lo > M.hi || ((M.lo | M.hi) & 15u) || !mc_base || mm_pct > 100)
		return cudaErrorInvalidValue;
	if (M.hi == M.lo)
		return cudaSuccess;
	uint64_t mid = M.lo + (M.hi - M.lo) / 100 * mm_pct / 16 * 16;
	if (mm_pct == 100)
		mid = M.hi;
	const unsigned total = M.grid ? M.grid : sm_count() * 8;
	unsigned mm_ctas = (unsigned)((uint64_t)total * mm_pct / 100);
	if (mid > M.lo && mm_ctas == 0)
		mm_ctas = 1;
	if (mid == M.lo)
		mm_ctas = 0;
	unsigned grid = mid < M.hi ? (total > mm_ctas ? total : mm_ctas + 1) : mm_ctas;
	uint64_t* mc = reinterpret_cast<uint64_t*>(mc_base);
	switch (M.world) {
	case 2: hybrid_merge_kernel<2><<<grid, 256, 0, stream>>>(M, mc, mid, mm_ctas); break;
	case 4: hybrid_merge_kernel<4><<<grid, 256, 0, stream>>>(M, mc, mid, mm_ctas); break;
	case 8: hybrid_merge_kernel<8><<<grid, 256, 0, stream>>>(M, mc, mid, mm_ctas); break;
	default: return cudaErrorInvalidValue;
	}
	return cudaGetLastError();
}

cudaError_t launch_multimem_or(void* mc_base, uint64_t lo, uint64_t hi, unsigned unroll, unsigned grid_ctas, cudaStream_t stream)
{
	if (!mc_base || lo > hi || ((lo | hi | (uint64_t)(uintptr_t)mc_base) & 15u))
		return cudaErrorInvalidValue;
	const uint64_t n_words = (hi - lo) / 8;
	if (n_words == 0)
		return cudaSuccess;
	uint64_t* p = reinterpret_cast<uint64_t*>(static_cast<uint8_t*>(mc_base) + lo);
	const unsigned u = unroll == 1 || unroll == 2 || unroll == 8 ? unroll : 4;
	uint64_t want = (n_words + 512ull * u - 1) / (512ull * u);
	const uint64_t cap = grid_ctas ? grid_ctas : sm_count() * 4;
	const unsigned grid = (unsigned)(want > cap ? cap : want);
	switch (u) {
	case 1: multimem_or_kernel<1><<<grid, 512, 0, stream>>>(p, n_words); break;
	case 2: multimem_or_kernel<2><<<grid, 512, 0, stream>>>(p, n_words); break;
	case 8: multimem_or_kernel<8><<<grid, 512, 0, stream>>>(p, n_words); break;
	default: multimem_or_kernel<4><<<grid, 512, 0, stream>>>(p, n_words); break;
	}
	return cudaGetLastError();
}

cudaError_t launch_merge(void* dst, const void* src, uint64_t nbytes, int saturating_add, cudaStream_t stream)
{
	uint64_t nvec = nbytes / 16;
	uint64_t want = (nvec + 255) / 256;
	unsigned grid = (unsigned)(want < 1 ? 1 : want > sm_count() * 8 ? sm_count() * 8 : want);
	merge_kernel<<<grid, 256, 0, stream>>>((uint8_t*)dst, (const uint8_t*)src, nbytes, saturating_add);
	return cudaGetLastError();
}

// ---------------------------------------------------------------- synthetic inputs
__global__ void __launch_bounds__(256) synth_genome_kernel(uint8_t* out, uint64_t start, uint64_t n, uint64_t seed)
{
	// one thread = 16 output bases (a 16-byte store when aligned)
	uint64_t nchunks = (n + 15) / 16;
	for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < nchunks; c += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t o = c * 16;
		uint8_t b[16];
#pragma unroll
		for (int j = 0; j < 16; j++)
			b[j] = "ACGT"[synth_code(start + o + j, seed)];
		if (o + 16 <= n && ((((uintptr_t)out) + o) & 15u) == 0) {
			uint4 v;
			v.x = b[0] | (b[1] << 8) | (b[2] << 16) | ((uint32_t)b[3] << 24);
			v.y = b[4] | (b[5] << 8) | (b[6] << 16) | ((uint32_t)b[7] << 24);
			v.z = b[8] | (b[9] << 8) | (b[10] << 16) | ((uint32_t)b[11] << 24);
			v.w = b[12] | (b[13] << 8) | (b[14] << 16) | ((uint32_t)b[15] << 24);
			*reinterpret_cast<uint4*>(out + o) = v;
		} else {
			for (int j = 0; j < 16 && o + j < n; j++)
				out[o + j] = b[j];
		}
	}
}

cudaError_t launch_synth_genome(uint8_t* out, uint64_t start, uint64_t n, uint64_t seed, cudaStream_t stream)
{
	if (n == 0)
		return cudaSuccess;
	uint64_t want = ((n + 15) / 16 + 255) / 256;
	unsigned grid = (unsigned)(want > sm_count() * 32 ? sm_count() * 32 : want);
	synth_genome_kernel<<<grid, 256, 0, stream>>>(out, start, n, seed);
	return cudaGetLastError();
}

__global__ void __launch_bounds__(256) synth_reads_kernel(uint8_t* out, uint64_t first_read, uint64_t n_reads,
                                                          unsigned read_len, uint64_t g_start, uint64_t g_len,
                                                          uint64_t gseed, uint64_t rseed)
{
	uint64_t total = n_reads * read_len;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t r = i / read_len;
		unsigned j = (unsigned)(i - r * read_len);
		uint64_t st = g_start + splitmix64(rseed + first_read + r) % (g_len - read_len);
		out[i] = "ACGT"[synth_code(st + j, gseed)];
	}
}

cudaError_t launch_synth_reads(uint8_t* out, uint64_t first_read, uint64_t n_reads, unsigned read_len,
                               uint64_t g_start, uint64_t g_len, uint64_t gseed, uint64_t rseed, cudaStream_t stream)
{
	uint64_t total = n_reads * read_len;
	if (total == 0)
		return cudaSuccess;
	uint64_t want = (total + 255) / 256;
	unsigned grid = (unsigned)(want > sm_count() * 64 ? sm_count() * 64 : want);
	synth_reads_kernel<<<grid, 256, 0, stream>>>(out, first_read, n_reads, read_len, g_start, g_len, gseed, rseed);
	return cudaGetLastError();
}

// ---------------------------------------------------------------- random-sector probe (roofline denominator)
__global__ void __launch_bounds__(256) random_probe_kernel(uint32_t* arr, uint64_t n_words, uint64_t n_access, int mode,
                                                           unsigned long long* sink)
{
	uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
	uint32_t acc = 0;
	FastMod fm = make_fastmod(n_words);
	for (uint64_t i = tid * 4; i < n_access; i += nthreads * 4) {
		uint64_t r0 = splitmix64(i), r1 = splitmix64(i + 1), r2 = splitmix64(i + 2), r3 = splitmix64(i + 3);
		uint64_t a0 = fm.pow2 ? fastmod<true>(r0, fm) : fastmod<false>(r0, fm);
		uint64_t a1 = fm.pow2 ? fastmod<true>(r1, fm) : fastmod<false>(r1, fm);
		uint64_t a2 = fm.pow2 ? fastmod<true>(r2, fm) : fastmod<false>(r2, fm);
		uint64_t a3 = fm.pow2 ? fastmod<true>(r3, fm) : fastmod<false>(r3, fm);
		if (mode == 0) {
			uint32_t v0 = __ldg(arr + a0), v1 = __ldg(arr + a1), v2 = __ldg(arr + a2), v3 = __ldg(arr + a3);
			acc += v0 + v1 + v2 + v3;
		} else {
			atomicOr(arr + a0, 1u << (r0 >> 59));
			atomicOr(arr + a1, 1u << (r1 >> 59));
			atomicOr(arr + a2, 1u << (r2 >> 59));
			atomicOr(arr + a3, 1u << (r3 >> 59));
		}
	}
	if (mode == 0 && acc == 0x9e3779b9u)
		atomicAdd(sink, 1ull);
}

cudaError_t launch_random_probe(uint32_t* arr, uint64_t n_words, uint64_t n_access, int mode,
                                unsigned long long* d_sink, cudaStream_t stream)
{
	random_probe_kernel<<<sm_count() * 8, 256, 0, stream>>>(arr, n_words, n_access, mode, d_sink);
	return cudaGetLastError();
}

} // namespace btl
