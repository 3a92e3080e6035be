// apply2.cu -- pass 2 of the partitioned BloomFilter build, two-level flavour.
//
// apply_bins_kernel (kernels.cu) ORs every binned offset into its L2-resident filter partition with one
// RED.OR per item, and tops out at the rate at which an SM can issue L2 atomics (~190 G/s on B200).  With
// the builds accumulating ~1.8 G items per pass, that is the larger half of the build.  Here the items of a
// partition are first split once more, by 64 KiB slice of the partition (refine_kernel: a counting sort in
// shared memory, the S / C / D phases of sort_bin.cuh on items that are loaded instead of hashed), and
// then every slice is ORed in SHARED memory (apply_slices_kernel: load the slice with coalesced 16-byte
// loads, shared-memory atomicOr per item, store it back) -- the filter is read and written exactly once,
// with plain vector accesses, and the per-item work moves from the L2 atomic units to the SMs' shared memory.
// An item that does not fit its level-2 bucket (skew) is ORed into the filter directly by refine_kernel, which
// finishes before apply_slices_kernel starts.  Bit-identical to the one-level pass: OR is order-free.
//
// STATUS: optional (context option bin_two_level=1), off by default.  Measured on B200 (cfg2, 627 M items):
// refine 2.14 ms (293 G items/s, 42 instructions per item, 37 % warps) + slices 2.92 ms (3.5 TB/s with three
// 64 KiB CTAs per SM) = 5.1 ms against ~3.5 ms for the one-level pass on the same items; kept with its parity
// test because the accounting (what a second split costs against what the L2 atomic rate costs) is the
// point of reference for any future attempt.
#include "kernels.cuh"

namespace btl {

constexpr int kRefineThreads = 256;
constexpr int kRefineItems = 16; // per thread and round
constexpr uint32_t kRefineRound = kRefineThreads * kRefineItems;
constexpr uint32_t kMaxSlices = 256; // per partition: one histogram bin per thread

__global__ void __launch_bounds__(kRefineThreads) refine_kernel(const __grid_constant__ Apply2Params A)
{
	__shared__ uint32_t sorted[kRefineRound];
	__shared__ uint16_t slice_of[kRefineRound];
	__shared__ uint32_t hist[kMaxSlices + 32], base[kMaxSlices], cursor[kMaxSlices];
	__shared__ uint64_t gdelta[kMaxSlices];
	__shared__ uint32_t wsum[kRefineThreads / 32 + 2];
	constexpr uint32_t NW = kRefineThreads / 32;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const uint32_t part = blockIdx.x / A.writers2, me = blockIdx.x % A.writers2;
	const uint32_t ns = A.n_sub;
	uint32_t* region = A.filter + ((uint64_t)part << (A.bin_shift - 5));

	hist[tid] = 0;
	cursor[tid] = 0;
	if (tid < 32)
		hist[kMaxSlices + tid] = 0;
	const uint32_t dump = kMaxSlices + (uint32_t)lane;
	__syncthreads();

	for (uint32_t w = me; w < A.writers; w += A.writers2) {
		uint32_t n = __ldg(A.counts + (uint64_t)part * A.writers + w);
		n = n < A.cap ? n : A.cap;
		const uint32_t* items = A.items + ((uint64_t)part * A.writers + w) * A.cap;
		const uint4* vec = reinterpret_cast<const uint4*>(items); // cap is a multiple of 8
		for (uint32_t i0 = 0; i0 < n; i0 += kRefineRound) {
			// ---- A: load 16 items (coalesced 16-byte vectors), slice = offset >> sub_shift, rank by histogram atomics
			uint32_t off[kRefineItems], pr[kRefineItems];
#pragma unroll
			for (int v = 0; v < kRefineItems / 4; v++) {
				const uint32_t first = i0 + ((uint32_t)v * kRefineThreads + (uint32_t)tid) * 4u;
				uint4 x = make_uint4(0, 0, 0, 0);
				if (first < n)
					x = __ldcs(vec + first / 4);
				off[v * 4 + 0] = x.x; off[v * 4 + 1] = x.y; off[v * 4 + 2] = x.z; off[v * 4 + 3] = x.w;
#pragma unroll
				for (int e = 0; e < 4; e++) {
					const bool ok = first + (uint32_t)e < n;
					const uint32_t bin = ok ? off[v * 4 + e] >> A.sub_shift : dump;
					pr[v * 4 + e] = (bin << 16) | (atomicAdd(hist + bin, 1u) & 0xffffu);
				}
			}
			__syncthreads();
			// ---- S: exclusive scan over the slices (one per thread), cursors, gdelta
			const uint32_t cnt = (uint32_t)tid < ns ? hist[tid] : 0u;
			hist[tid] = 0;
			uint32_t incl = cnt;
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
			for (uint32_t x = 0; x < NW; x++) {
				const uint32_t y = wsum[x];
				before += x < (uint32_t)warp ? y : 0u;
				total += y;
			}
			if (tid == 0)
				wsum[NW] = total;
			const uint32_t excl = before + incl - cnt;
			base[tid] = excl;
			if (cnt) {
				const uint32_t c = cursor[tid];
				cursor[tid] = c + cnt; // < 2^32: a level-2 bucket sees at most the items of one partition
				gdelta[tid] = (((uint64_t)part * ns + (uint32_t)tid) * A.writers2 + me) * A.cap2 + c - excl;
				if (c + cnt > A.cap2)
					wsum[NW + 1] = 1;
			}
			__syncthreads();
			// ---- C: scatter into the sorted buffer
#pragma unroll
			for (int e = 0; e < kRefineItems; e++) {
				const uint32_t bin = pr[e] >> 16;
				if (bin < kMaxSlices) {
					const uint32_t pos = base[bin] + (pr[e] & 0xffffu);
					sorted[pos] = off[e];
					slice_of[pos] = (uint16_t)bin;
				}
			}
			__syncthreads();
			// ---- D: copy out, consecutive threads write consecutive items
			const bool overflow = wsum[NW + 1] != 0;
			total = wsum[NW];
			for (uint32_t pos = tid; pos < total; pos += kRefineThreads) {
				const uint32_t o = sorted[pos], bin = slice_of[pos];
				const uint64_t idx = gdelta[bin] + pos;
				if (overflow && idx - (((uint64_t)part * ns + bin) * A.writers2 + me) * A.cap2 >= A.cap2) {
					atomicOr(region + (o >> 5), 1u << (o & 31)); // full bucket: straight into the filter
					continue;
				}
				A.items2[idx] = o;
			}
			// (the next round's phase A only touches hist and registers; its first barrier orders the rest)
		}
	}
	__syncthreads();
	if ((uint32_t)tid < ns)
		A.counts2[((uint64_t)part * ns + (uint32_t)tid) * A.writers2 + me] = cursor[tid];
}

// one CTA = one slice of one partition: OR its level-2 buckets into the slice in shared memory
__global__ void __launch_bounds__(256) apply_slices_kernel(const __grid_constant__ Apply2Params A)
{
	extern __shared__ __align__(16) uint32_t slice[];
	__shared__ uint32_t any;
	const int tid = threadIdx.x;
	const uint32_t part = blockIdx.x / A.n_sub, sl = blockIdx.x % A.n_sub;
	const uint64_t bit0 = ((uint64_t)part << A.bin_shift) + ((uint64_t)sl << A.sub_shift);
	if (bit0 >= A.m)
		return;
	const uint64_t word0 = bit0 >> 5;
	uint64_t nwords = (uint64_t)1 << (A.sub_shift - 5);
	if (word0 + nwords > A.alloc_words)
		nwords = A.alloc_words - word0; // the filter's allocation is a whole number of 16-byte vectors
	const uint32_t* cnt = A.counts2 + ((uint64_t)part * A.n_sub + sl) * A.writers2;
	if (tid == 0)
		any = 0;
	__syncthreads();
	uint32_t mine = 0;
	for (uint32_t w = tid; w < A.writers2; w += blockDim.x)
		mine |= __ldg(cnt + w);
	if (mine)
		any = 1;
	__syncthreads();
	if (!any)
		return; // nothing landed in this slice: leave it alone
	uint4* g = reinterpret_cast<uint4*>(A.filter + word0);
	uint4* s4 = reinterpret_cast<uint4*>(slice);
	const uint32_t nvec = (uint32_t)(nwords / 4);
	for (uint32_t i = tid; i < nvec; i += blockDim.x)
		s4[i] = g[i];
	__syncthreads();
	const uint32_t mask = (1u << A.sub_shift) - 1u;
	const uint32_t warps = blockDim.x / 32, lane = tid & 31;
	for (uint32_t w = tid >> 5; w < A.writers2; w += warps) {
		uint32_t n = __ldg(cnt + w);
		n = n < A.cap2 ? n : A.cap2;
		const uint32_t* items = A.items2 + (((uint64_t)part * A.n_sub + sl) * A.writers2 + w) * A.cap2;
		const uint4* v = reinterpret_cast<const uint4*>(items); // cap2 is a multiple of 4
		const uint32_t nv = n / 4;
		for (uint32_t i = lane; i < nv; i += 32) {
			const uint4 x = __ldcs(v + i);
			atomicOr(slice + ((x.x & mask) >> 5), 1u << (x.x & 31));
			atomicOr(slice + ((x.y & mask) >> 5), 1u << (x.y & 31));
			atomicOr(slice + ((x.z & mask) >> 5), 1u << (x.z & 31));
			atomicOr(slice + ((x.w & mask) >> 5), 1u << (x.w & 31));
		}
		for (uint32_t i = nv * 4 + lane; i < n; i += 32) {
			const uint32_t o = __ldcs(items + i);
			atomicOr(slice + ((o & mask) >> 5), 1u << (o & 31));
		}
	}
	__syncthreads();
	for (uint32_t i = tid; i < nvec; i += blockDim.x)
		g[i] = s4[i];
}

bool apply2_geometry(uint32_t bin_shift, uint32_t* sub_shift, uint32_t* n_sub)
{
	// 64 KiB slices (three CTAs per SM), at most kMaxSlices of them per partition
	uint32_t ss = 19;
	if (bin_shift < ss)
		ss = bin_shift < 7 ? 7 : bin_shift;
	while (bin_shift - ss > 8)
		ss++;
	if (ss > 20 || ss < 7 || ss > bin_shift) // a slice must fit shared memory (128 KiB at most) and hold whole vectors
		return false;
	*sub_shift = ss;
	*n_sub = 1u << (bin_shift - ss);
	return *n_sub <= kMaxSlices;
}

cudaError_t launch_apply2(const Apply2Params& A, cudaStream_t stream)
{
	if (A.n_bins == 0)
		return cudaSuccess;
	const uint64_t g1 = (uint64_t)A.n_bins * A.writers2, g2 = (uint64_t)A.n_bins * A.n_sub;
	if (g1 > 0x7fffffffULL || g2 > 0x7fffffffULL || A.n_sub > kMaxSlices)
		return cudaErrorInvalidValue;
	refine_kernel<<<(unsigned)g1, kRefineThreads, 0, stream>>>(A);
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		return e;
	const size_t smem = (size_t)1 << (A.sub_shift - 3);
	if (smem > 48 * 1024) {
		e = cudaFuncSetAttribute(apply_slices_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess)
			return e;
	}
	apply_slices_kernel<<<(unsigned)g2, 256, smem, stream>>>(A);
	return cudaGetLastError();
}

} // namespace btl
