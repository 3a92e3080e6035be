// kernels.cuh -- launch interface between the C ABI (capi.cu) and the sm_100a kernels (kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "nthash_dev.cuh"

namespace btl {

constexpr int kMaxHash = 64;  // hashes per k-mer (hash_num, or n_seeds*h2)
constexpr int kMaxSeeds = 16; // spaced seeds
constexpr int kTPB = 128;     // threads per CTA of the sequence kernels
constexpr int kWPT = 32;      // consecutive windows rolled by one thread
constexpr int kTile = kTPB * kWPT; // windows per tile (4096)

enum SeqOp {
	OP_HASH = 0,         // emit raw hashes / strands / valid bits
	OP_BF_INSERT = 1,    // BloomFilter::insert
	OP_BF_CONTAINS = 2,  // BloomFilter::contains
	OP_CBF_MINCOUNT = 3, // CountingBloomFilter::minCount (+ contains via threshold)
	OP_CBF_INCALL = 4,   // CountingBloomFilter::incrementAll
	// order-dependent updates (incrementMin, insertAndCheck) reproduced exactly, three passes per batch:
	OP_RESV_TOUCH = 5,   // pass 1: mark reservation bits, detect contended slots
	OP_CBF_COMMIT = 6,   // pass 2 (incrementMin): commit uncontended k-mers, defer the rest
	OP_RESV_CLEAR = 7,   // pass 3: clear the reservation bits this batch set
	OP_BFCHK_COMMIT = 8, // pass 2 (insertAndCheck)
	OP_BF_BIN = 9        // partitioned BloomFilter build, pass 1: bin bit indices by filter partition
};

struct SeqParams
{
	// input batch (device pointers)
	const uint8_t* bases;    // chunk of the flat base stream
	uint64_t n_bases;        // bytes readable at `bases`
	uint64_t base0;          // flat position of bases[0] in the whole batch (for offsets lookup)
	uint64_t n_windows;      // windows [0, n_windows) of this chunk are processed by this launch
	const uint64_t* offsets; // n_seqs + 1 flat sequence offsets of the whole batch
	uint64_t n_seqs;
	// 2-bit packed input (btlbf_*_seqs_packed): `bases` then holds 4 bases per byte (base i = bits 2*(i&3) of byte
	// i>>2, A0 C1 G2 T3) and `invalid` one bit per base (bit i&7 of byte i>>3; may be null: every base valid);
	// n_bases, base0, n_windows keep counting bases
	const uint8_t* invalid;
	uint32_t packed;
	// hashing
	uint32_t k, h;           // k-mer size, hashes per k-mer
	uint32_t n_seeds, h2;    // spaced seeds: h == n_seeds*h2
	const uint64_t* st_tab;  // spaced: TF[k][8] then TR[k][8]
	const uint16_t* st_dc;   // spaced: concatenated don't-care positions
	uint32_t st_dc_off[kMaxSeeds + 1];
	uint64_t mult[kMaxHash]; // mult[i] = i ^ k*multiSeed (i >= 1)
	uint64_t g_f[16], g_fk[16], g_r[16], g_rk[16]; // seed tables by base class; *k = R^k applied
	// filter
	void* filter;
	FastMod fm;
	uint32_t threshold;
	// exact counting insert: reservation bit tables (touched / contended), 2^resv_log2 bits each
	uint32_t* resv_touched;
	uint32_t* resv_contended;
	uint32_t resv_log2;
	uint32_t* pending;       // deferred window list (chunk-local window indices)
	uint32_t* pending_count;
	uint64_t* pending_slots; // optional: the h slots of deferred window w at [w*h ..), written when it is deferred,
	                         // so that the residual rounds need not re-derive them from the bases
	// partitioned build: sub-bucket (partition p, writer w) holds up to bin_cap 32-bit offsets at
	// bin_items[(p*bin_writers + w)*bin_cap ...]; bin_counts[p*bin_writers + w] = appended (may exceed cap)
	uint32_t* bin_items;
	uint32_t* bin_counts;
	uint32_t bin_shift;   // log2(bits per partition)
	uint32_t bin_mask;    // (1 << bin_shift) - 1
	uint32_t bin_cap;
	uint32_t bin_writers;
	uint32_t n_bins;
	uint32_t bin_segs;    // pass 2 splits every sub-bucket into this many segments (work units)
	uint32_t bin_legacy;  // knob: never use the sort-bin kernel
	uint32_t bin_rot;     // sort-bin kernel: tile t goes to writer (t + bin_rot) % bin_writers
	uint32_t bin_counting;    // partitioned query of a counting filter: an item is a counter index, the test is >= threshold
	uint32_t probe_ld;        // pass 2 of the query, experiment knob: 0 ld.global.nc, 1 ld.global.cg, 2 L1::no_allocate
	uint32_t bin_prefetch;    // pass 2: pull the next partition into L2 while this one is processed
	uint32_t bin_part0, bin_part_count; // pass 2 of the build over partitions [bin_part0, bin_part0 + bin_part_count) only
	                                    // (count 0: all of them) -- btlbf_filter_flush_parts
	uint32_t bin_ctas_per_sm; // pass 1 (query): persistent CTAs per SM (0 = as many as fit); fewer leave room for a
	                          // concurrent pass 2
	// outputs (chunk-local indexing by window)
	uint32_t* hit_bits;
	uint32_t* valid_bits;
	uint64_t out_words;      // uint32 words available in hit_bits / valid_bits
	uint8_t* counts;
	uint64_t* hashes;
	uint8_t* strands;
	uint64_t* stats; // [0] += valid k-mers, [1] += hits
	// knobs
	uint32_t force_generic;
	uint32_t query_mode;
	// device-side path selection (adaptive query): a kernel whose gate is set runs only when *gate == gate_want
	const uint32_t* gate;
	uint32_t gate_want;
	// seq_kernel over a subset of the tiles: tile = tile_first + blockIdx.x * tile_stride, tile_count tiles
	// (tile_count == 0: every tile of the chunk)
	uint32_t tile_first, tile_stride, tile_count;
	uint32_t tiles_per_cta; // seq_kernel: consecutive tiles handled by one CTA (0 = 1)
	uint32_t ungrouped_commit; // ordered updates, pass 2: one window at a time (the general form) even where the grouped form applies
};

size_t seq_kernel_smem_bytes(uint32_t k, bool spaced);
// launches seq_kernel<op> for P on stream; returns cudaGetLastError()
cudaError_t launch_seq(SeqOp op, const SeqParams& P, cudaStream_t stream);

// Partitioned BloomFilter build.  bin_plan: number of persistent writer CTAs the bin kernel will use for
// P (so that the caller can size the sub-buckets); launch_bin: pass 1 (hash + bin, persistent CTAs);
// launch_apply_bins: pass 2 (per partition, OR the binned offsets into the L2-resident filter region).
// query == true: the partitioned QUERY (items are (offset, window) pairs; pass 2 = launch_probe_bins, then
// launch_finalize_hits ANDs the hit words with the valid words and counts the hits).
enum BinMode { BIN_SORT = 0, BIN_WARP = 1, BIN_CTA = 2 };
cudaError_t bin_plan(const SeqParams& P, uint32_t n_bins, bool query, uint32_t* writers, uint32_t* grid, int* mode);
cudaError_t launch_bin(const SeqParams& P, bool query, uint32_t grid, cudaStream_t stream);
cudaError_t launch_apply_bins(const SeqParams& P, cudaStream_t stream);
// counting: the filter holds 8-bit counters, the test is counter >= P.threshold; unroll: item vectors in flight per thread
cudaError_t launch_probe_bins(const SeqParams& P, bool counting, int unroll, bool maxshared, cudaStream_t stream);
cudaError_t launch_finalize_hits(uint32_t* hit, const uint32_t* valid, uint64_t n_words, unsigned long long* hits_out,
                                 const uint32_t* gate, uint32_t gate_want, cudaStream_t stream);
// *flag = 1 (direct early-exit kernel) when fewer than pct percent of the sampled k-mers {stats[0] valid,
// stats[1] hits} were hits, else 0 (partitioned query)
cudaError_t launch_query_gate(const unsigned long long* sample_stats, uint32_t* flag, uint32_t pct, cudaStream_t stream);
bool bin_query_supported(const SeqParams& P, uint32_t n_bins);
bool bin_sort_eligible(const SeqParams& P, uint32_t n_bins); // the sort-bin kernel (sort_bin.cuh) serves this shape
uint32_t bin_sort_tile(const SeqParams& P, uint32_t n_bins); // windows per CTA pass of the sort-bin kernel for this shape

// Two-level pass 2 of the partitioned build (apply2.cu): the items of every partition are split once more by
// slice (refine), then each slice is ORed in shared memory and written back once.
struct Apply2Params
{
	const uint32_t* items;  // level 1: sub-bucket (partition p, writer w) at items[(p*writers + w)*cap ...]
	const uint32_t* counts; // level 1: appended per sub-bucket (may exceed cap)
	uint32_t n_bins, writers, cap, bin_shift;
	uint32_t* items2;       // level 2: bucket (partition p, slice s, writer j) at items2[((p*n_sub + s)*writers2 + j)*cap2 ...]
	uint32_t* counts2;
	uint32_t sub_shift, n_sub, writers2, cap2; // log2(bits per slice), slices per partition, refine CTAs per partition
	uint32_t* filter;
	uint64_t m;             // filter bits
	uint64_t alloc_words;   // 32-bit words of the filter's allocation (a multiple of 4)
};
bool apply2_geometry(uint32_t bin_shift, uint32_t* sub_shift, uint32_t* n_sub);
cudaError_t launch_apply2(const Apply2Params& A, cudaStream_t stream);

// residual rounds of the ordered updates on the compacted list of deferred windows
struct ListParams
{
	const uint32_t* list_in;
	const uint32_t* count_in;
	uint32_t* list_out;
	uint32_t* count_out;  // must be zero before a commit round
	uint64_t* resv;       // 2^resv_log2 reservation words, initialised to all-ones
	uint32_t resv_log2;
	uint32_t epoch;       // strictly increasing across rounds
	uint32_t max_items;   // upper bound of *count_in (sizes the grid)
	uint32_t kind;        // 0: incrementMin, 1: insertAndCheck
	uint32_t* rounds_out; // drain kernels: number of rounds they ran (added)
	uint32_t* d_epoch;    // cooperative drain: the epoch counter lives on the device
	uint32_t* counts;     // cooperative drain: the two ping-pong list counters {count(list_in), count(list_out)}
};
// phase 0: reserve, phase 1: commit or re-queue (one grid-wide launch each)
cudaError_t launch_list_round(int phase, const SeqParams& P, const ListParams& L, cudaStream_t stream);
// one CTA loops reserve/commit rounds until the list is empty (short lists, long dependency chains)
cudaError_t launch_list_drain(const SeqParams& P, const ListParams& L, cudaStream_t stream);
// the whole GPU loops reserve/commit rounds (grid-wide barriers) until the list is empty: no host round
// trip at all, so batches of ordered updates can be queued back to back.  Cooperative launch.
cudaError_t launch_list_drain_coop(const SeqParams& P, const ListParams& L, cudaStream_t stream);

// Legacy per-k-mer interface of the reference classes (caller supplies the h hash values of each k-mer,
// BloomFilter.hpp:185-262, CountingBloomFilter.hpp:53-64,134-214): hashes = n x h values.
// op 0: BF insert, 1: BF contains -> out[i], 2: CBF minCount -> out[i], 3: CBF incrementMin (+ out[i] =
// minCount >= threshold beforehand, i.e. insertAndCheck), 4: CBF incrementAll, 5: BF insertAndCheck -> out[i].
// Ops 3 and 5 are order-dependent and run the n k-mers one after the other in a single thread.
cudaError_t launch_hashes_op(int op, void* filter, FastMod fm, uint32_t h, uint32_t threshold,
                             const uint64_t* d_hashes, uint64_t n, uint8_t* d_out, cudaStream_t stream);
cudaError_t launch_popcount(const void* data, uint64_t nbytes, int mode, unsigned threshold,
                            unsigned long long* d_out, cudaStream_t stream);
cudaError_t launch_merge(void* dst, const void* src, uint64_t nbytes, int saturating_add,
                         cudaStream_t stream);
// fused multi-GPU merge over peer-mapped memory: base[p] = rank p's partial filter (base[rank] is local);
// this rank reduces bytes [lo, hi) of all of them and stores the result into all of them
constexpr int kMaxPeers = 16;
struct PeerMergeParams
{
	uint8_t* base[kMaxPeers];
	uint32_t world;
	int sat_add;
	uint64_t lo, hi; // multiples of 16
	uint32_t unroll; // 16-byte vectors per thread and peer in flight (1, 2 or 4)
	uint32_t grid;   // CTAs (0: 8 per SM)
	uint32_t mode;   // 0: the merge; measurement only: 1 = peer loads without peer stores, 2 = peer stores without peer loads
};
cudaError_t launch_peer_merge(const PeerMergeParams& M, cudaStream_t stream);
// OR of all replicas of bytes [lo, hi) of an NVLS multicast mapping, written back to all replicas
cudaError_t launch_multimem_or(void* mc_base, uint64_t lo, uint64_t hi, unsigned unroll, unsigned grid_ctas, cudaStream_t stream);
// the first mm_pct per cent of [M.lo, M.hi) through the multicast mapping, the rest through the peer pointers, in one kernel
cudaError_t launch_hybrid_merge(const PeerMergeParams& M, void* mc_base, unsigned mm_pct, cudaStream_t stream);
cudaError_t launch_synth_genome(uint8_t* out, uint64_t start, uint64_t n, uint64_t seed,
                                cudaStream_t stream);
cudaError_t launch_synth_reads(uint8_t* out, uint64_t first_read, uint64_t n_reads,
                               unsigned read_len, uint64_t g_start, uint64_t g_len, uint64_t gseed,
                               uint64_t rseed, cudaStream_t stream);
cudaError_t launch_random_probe(uint32_t* arr, uint64_t n_words, uint64_t n_access, int mode,
                                unsigned long long* d_sink, cudaStream_t stream);

} // namespace btl
