// seq_kernel.cuh -- the fused sequence kernel (one CTA per tile) and its launcher, as templates over the
// operation.  kernels.cu instantiates the plain operations; the two ordered-update commit passes, whose grouped
// form is the bulk of the compile time, are instantiated in seq_commit_cbf.cu / seq_commit_bfchk.cu.
#pragma once
#include "tile_core.cuh"

namespace btl {

// one tile: stage -> classify/pack -> roll + fused operation -> per-window result words + statistics
template<int OP, bool SPACED, bool POW2>
__device__ __forceinline__ void run_tile(const SeqParams& P, const TileSmem& sm, uint64_t t0, int tid)
{
	tile_phase_a(P, sm, t0, tid, kTPB);
	__syncthreads();
	tile_phase_b(P, sm, t0, tid, kTPB);
	__syncthreads();
	ThreadOut out = tile_phase_c<OP, SPACED, POW2>(P, sm, t0, tid);

	// per-thread result words: 32 consecutive windows -> one coalesced 32-bit store per plane
	uint64_t widx = (t0 >> 5) + tid;
	if (widx < P.out_words) {
		if (P.valid_bits)
			P.valid_bits[widx] = out.validw;
		if (P.hit_bits)
			P.hit_bits[widx] = out.hitw;
	}
	if (P.stats) {
		uint32_t nv = __popc(out.validw), nh = __popc(out.hitw);
		nv = __reduce_add_sync(0xffffffffu, nv);
		nh = __reduce_add_sync(0xffffffffu, nh);
		uint32_t* red = reinterpret_cast<uint32_t*>(sm.scratch + 2);
		if ((tid & 31) == 0) {
			red[(tid >> 5) * 2] = nv;
			red[(tid >> 5) * 2 + 1] = nh;
		}
		__syncthreads();
		if (tid == 0) {
			uint32_t sv = 0, sh = 0;
			for (int w = 0; w < kTPB / 32; w++) {
				sv += red[w * 2];
				sh += red[w * 2 + 1];
			}
			if (sv)
				atomicAdd((unsigned long long*)&P.stats[0], (unsigned long long)sv);
			if (sh)
				atomicAdd((unsigned long long*)&P.stats[1], (unsigned long long)sh);
		}
	}
}

template<int OP, bool SPACED, bool POW2>
__global__ void __launch_bounds__(kTPB) seq_kernel(const __grid_constant__ SeqParams P)
{
	extern __shared__ __align__(16) uint8_t smem_raw[];
	if (P.gate && *P.gate != P.gate_want)
		return;
	const TileSmem sm = carve_smem(smem_raw, P.k, SPACED);
	if (P.tile_count) { // a strided sample of the tiles
		run_tile<OP, SPACED, POW2>(P, sm, ((uint64_t)P.tile_first + (uint64_t)blockIdx.x * P.tile_stride) * kTile, threadIdx.x);
		return;
	}
	const uint32_t per = P.tiles_per_cta ? P.tiles_per_cta : 1u;
	const uint64_t tiles = (P.n_windows + kTile - 1) / kTile;
	for (uint32_t j = 0; j < per; j++) {
		const uint64_t tile = (uint64_t)blockIdx.x * per + j;
		if (tile >= tiles)
			break;
		if (j)
			__syncthreads(); // the previous tile is fully consumed before its staging area is overwritten
		run_tile<OP, SPACED, POW2>(P, sm, tile * kTile, threadIdx.x);
	}
}

template<int OP, bool SPACED, bool POW2>
inline cudaError_t launch_one(const SeqParams& P, cudaStream_t stream)
{
	size_t smem = tile_smem_bytes(P.k, SPACED);
	auto kern = seq_kernel<OP, SPACED, POW2>;
	if (smem > 48 * 1024) {
		cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		if (e != cudaSuccess)
			return e;
	}
	uint64_t tiles = (P.n_windows + kTile - 1) / kTile;
	if (P.tile_count)
		tiles = P.tile_count;
	else if (P.tiles_per_cta > 1)
		tiles = (tiles + P.tiles_per_cta - 1) / P.tiles_per_cta;
	if (tiles == 0)
		return cudaSuccess;
	if (tiles > 0x7fffffffULL)
		return cudaErrorInvalidValue;
	kern<<<(unsigned)tiles, kTPB, smem, stream>>>(P);
	return cudaGetLastError();
}

template<int OP>
inline cudaError_t launch_op(const SeqParams& P, cudaStream_t stream)
{
	bool spaced = P.n_seeds != 0, pow2 = P.fm.pow2 != 0;
	if (spaced)
		return pow2 ? launch_one<OP, true, true>(P, stream) : launch_one<OP, true, false>(P, stream);
	return pow2 ? launch_one<OP, false, true>(P, stream) : launch_one<OP, false, false>(P, stream);
}

cudaError_t launch_seq_bloom(SeqOp op, const SeqParams& P, cudaStream_t stream); // seq_ops_bloom.cu
cudaError_t launch_seq_cbf_commit(const SeqParams& P, cudaStream_t stream);
cudaError_t launch_seq_bfchk_commit(const SeqParams& P, cudaStream_t stream);

// general-shape pass-1 kernels of the partitioned paths (bin_legacy.cu)
constexpr uint32_t kMaxWarpBins = 256;
const void* bin_warp_kernel(bool query, bool spaced, bool pow2);
const void* bin_cta_kernel(bool spaced, bool pow2);
size_t bin_warp_smem_bytes(uint32_t k, bool spaced, uint32_t n_bins);

} // namespace btl
