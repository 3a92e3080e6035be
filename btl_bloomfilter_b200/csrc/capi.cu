// capi.cu -- the C ABI of libbtlbf_cuda.so (include/btlbf.h): handles, device memory, host<->device
// pipelines and kernel orchestration.  No CPU implementation of the hot path lives here: every batched
// operation is a launch of the sm_100a kernels in kernels.cu.
#include "../../include/btlbf.h"
#include "kernels.cuh"
#include "host_params.hpp"

#include <atomic>
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

using namespace btl;

// pack.cu: ASCII -> 2-bit codes + invalid plane on host threads
void btl_pack_bases(const char* bases, uint64_t n_bases, uint8_t* codes, uint8_t* invalid, int threads, uint64_t* n_invalid,
                    uint64_t* first_raw);

// ---------------------------------------------------------------- errors
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof g_err, fmt, ap);
	va_end(ap);
	return code;
}

#define CU(call)                                                                                   \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if (e_ != cudaSuccess)                                                                     \
			return fail(e_ == cudaErrorMemoryAllocation ? BTLBF_ERR_NOMEM : BTLBF_ERR_CUDA,       \
			            "%s failed: %s", #call, cudaGetErrorString(e_));                           \
	} while (0)
#define TRY(call)                                                                                   \
	do {                                                                                           \
		int rc_ = (call);                                                                          \
		if (rc_ != BTLBF_OK)                                                                       \
			return rc_;                                                                            \
	} while (0)

// ---------------------------------------------------------------- handles
namespace {

struct DevBuf
{
	void* p = nullptr;
	size_t cap = 0;
};

struct Slot // one stage of the host-buffer pipeline (H2D copy | kernel | D2H copy)
{
	DevBuf bases, invalid, hit, valid, counts, hashes, strands;
	cudaEvent_t ev_h2d = nullptr, ev_kernel = nullptr, ev_d2h = nullptr;
	bool used = false;
	// pinned staging of the host-side packing ("host_pack"): the chunk as 2-bit codes + invalid plane
	uint8_t* pk_codes = nullptr;
	uint8_t* pk_inv = nullptr;
	size_t pk_cap = 0; // bases the staging buffers hold
};

struct Ticket // per-call state of the host-buffer pipeline (several calls may be in flight)
{
	DevBuf offsets;
	unsigned long long* d_stats = nullptr; // {n_kmers, n_hits}
	cudaEvent_t done = nullptr;
	bool in_use = false;
};
constexpr int kTickets = 4;

struct HashCfg // everything the kernels need to turn a window into its hashes
{
	uint32_t k = 0, h = 0, n_seeds = 0, h2 = 0;
	uint64_t* d_st_tab = nullptr;
	uint16_t* d_st_dc = nullptr;
	SeqParams proto;
};

// Per (host thread, filter) staging of the legacy per-k-mer updates: a thread appends to its own queue under the queue's
// own flag (uncontended in the normal case) and takes the context lock only once per kTlCap k-mers, so that an OpenMP
// loop of bloom.insert(*itr) over one shared filter (Tests/AdHoc/ParallelFilter.cpp:104-122) does not serialise on the
// context.  Any thread that touches the filter drains all of its queues first (joined()).
struct TlQueue
{
	std::atomic<int> busy{ 0 };
	uint32_t n = 0;
	int op = -1;
	uint64_t* data = nullptr; // kTlCap x h hash values, in call order
	std::thread::id owner;    // the host thread that appends to it
};
constexpr uint32_t kTlCap = 512;

struct TlGuard
{
	TlQueue* q;
	explicit TlGuard(TlQueue* q_) : q(q_)
	{
		int expect = 0;
		while (!q->busy.compare_exchange_weak(expect, 1, std::memory_order_acquire))
			expect = 0;
	}
	~TlGuard() { q->busy.store(0, std::memory_order_release); }
};

} // namespace

struct btlbf_ctx
{
	// One lock per context: every entry point that touches the context (or a filter of it) holds it for the
	// duration of the call, so the handles may be shared between host threads (the reference's insert is
	// thread-safe by atomics, BloomFilter.hpp:171-194, and its OpenMP driver calls it concurrently,
	// Tests/AdHoc/ParallelFilter.cpp:104-122).  Recursive: entry points call each other.
	std::recursive_mutex mu;
	int device = 0;
	cudaEvent_t ev_switch = nullptr;  // orders a newly selected active stream after the old one
	btlbf_filter* hq_owner = nullptr; // filter whose per-k-mer queue (legacy interface) holds unapplied updates
	bool in_hq_flush = false;
	std::atomic<uint64_t> tl_pending{ 0 };  // per-thread queues (TlQueue) of this context's filters that hold updates
	std::vector<btlbf_filter*> tl_filters;  // filters that own per-thread queues
	cudaStream_t own = nullptr, active = nullptr, copy_in = nullptr, copy_out = nullptr;
	// Background stream: pass 2 of the partitioned build (memory-bound) runs here, concurrently with
	// whatever the active stream does next (typically the compute-bound pass 1 of the following batch).
	cudaStream_t aux = nullptr;
	cudaEvent_t ev_bin_done[2] = { nullptr, nullptr }, ev_apply_done[2] = { nullptr, nullptr };
	bool slot_used[2] = { false, false };
	int bin_slot = 0;
	cudaEvent_t ev_aux_last = nullptr; // last event recorded on aux (one of ev_apply_done)
	bool aux_pending = false;          // the active stream has not been ordered after ev_aux_last yet
	int64_t overlap = 0; // 1: run pass 2 of the partitioned build on the background stream
	uint64_t launches = 0;
	unsigned long long* d_scalars = nullptr; // 16 device words: [0..1] stats, [2] popcount, [4..7] list counters
	unsigned long long* h_scalars = nullptr; // pinned mirror
	int64_t force_generic = 0, query_mode = 0, ungrouped_commit = 0;
	int64_t chunk_bases = (int64_t)64 << 20; // windows per pipeline stage of the host-buffer calls
	int64_t cbf_batch = (int64_t)4 << 20;    // windows per batch of the ordered (exact) updates
	int64_t resv_log2 = 28, list_log2 = 24;  // reservation sketch bits per table (2 x 32 MiB: stays in L2) / residual-round table entries
	int64_t drain_threshold = 4096;
	int64_t ordered_coop = 1; // residual rounds of the ordered updates: 1 cooperative grid kernel, 0 host-driven rounds
	int64_t bin_mode = 0;       // partitioned BloomFilter build: 0 auto, 1 always, -1 never
	int64_t bin_part_log2 = 27; // bits per filter partition (2^27 bits = 16 MiB: two of them resident in L2)
	int64_t bin_slack_pct = 20;
	int64_t bin_query_mode = 0; // partitioned query: 0 auto, 1 always (when supported), -1 never
	Slot slot[2];
	Ticket ticket[kTickets];
	uint64_t slot_seq = 0, ticket_seq = 0;
	DevBuf ibin_items[2], ibin_counts[2], qbin_items[2], qbin_counts[2], q_hit, q_valid;
	// The partitioned query runs in sub-batches: pass 1 (hash + bin, issue-bound) of sub-batch i+1 on the active
	// stream shares the SMs with pass 2 (probe, L2-bound) of sub-batch i on the background stream.
	cudaEvent_t ev_qp1[2] = { nullptr, nullptr }, ev_qp2[2] = { nullptr, nullptr };
	bool qslot_used[2] = { false, false };
	int64_t query_sub = 0;          // sub-batches per partitioned query (0 auto, 1: no overlap)
	int64_t query_p1_ctas = 0;      // pass-1 CTAs per SM while overlapping (0 auto: one fewer than fit)
	int64_t query_probe_unroll = 0; // item vectors in flight per thread of pass 2 (0 auto)
	int64_t probe_ld = 0, probe_carveout = -1; // experiment knobs of the probe kernel (load flavour; L1 split: -1 auto, 0 large L1, 1 small)
	int64_t bin_prefetch = -1;      // pass 2 prefetch into L2: -1 auto (see bin_setup), 0 none, 1 the next partition, 2 the CTA's share of its own
	DevBuf ibin2_items, ibin2_counts; // level-2 buckets of the two-level pass 2 (apply2.cu)
	int64_t bin_two_level = 0;        // 1: two-level pass 2 (apply2.cu); measured slower than the L2-atomics pass on B200
	int64_t bin_two_level_min = (int64_t)1 << 26; // items below which the one-level pass is used
	uint64_t two_level_passes = 0;
	uint64_t binned_launches = 0;
	// Accumulation of the partitioned build: pass 1 of successive batches appends to the same sub-buckets and
	// pass 2 runs once they are full or the filter contents are needed (settle()).
	struct
	{
		btlbf_filter* f = nullptr; // filter whose k-mers are parked in the sub-buckets (nullptr: none)
		SeqParams P;               // geometry of the sub-buckets + the filter they belong to
		uint64_t windows = 0, capacity = 0, tiles = 0;
		int slot = 0;
	} acc;
	// btlbf_filter_flush_parts in progress: the filter, its partition geometry and the number of chunks (n_bins 0: the
	// parked work could not be applied in parts and chunk 0 applied all of it)
	struct
	{
		btlbf_filter* f = nullptr;
		uint32_t n_bins = 0, shift = 0, n_chunks = 0;
	} parts;
	int64_t bin_accum_bytes = (int64_t)8 << 30; // sub-bucket storage one accumulation may use
	int64_t bin_kernel = 0;                     // 0 auto (sort-bin kernel when the shape allows), 1 legacy kernels only
	int64_t bin_max_parts = 512;                // the sort-bin path widens the partitions until there are at most this many
	int settle_error = 0;
	int64_t peer_unroll = 1, peer_grid = 0, peer_mode = 0; // fused multi-GPU merge: vectors in flight per thread and peer, CTAs
	int64_t mm_unroll = 4, mm_grid = 0;                    // in-switch (multimem) merge: 8-byte words in flight per thread, CTAs
	int64_t wrap_accumulate = 0; // 1: filters over caller-owned memory may park k-mers too (the caller flushes before reading)
	// 1: the ASCII host-buffer calls (insert / contains) pack every chunk to 2 bits per base on host threads before the
	// H2D copy (a quarter of the PCIe bytes; the CPU packs chunk i+1 while the GPU works on chunk i)
	int64_t host_pack = 0, host_pack_threads = 8;
	int64_t query_adaptive = 1;      // partitioned query: sample the batch, fall back to the early-exit kernel when few k-mers hit
	int64_t query_adaptive_pct = 20; // ... fewer than this percentage of the sampled k-mers
	int64_t query_adaptive_min_tiles = 256; // batches below this many 4096-window tiles are not sampled
	int64_t query_chunk_factor = 4; // BloomFilter queries of the host-buffer calls run in chunks of this many chunk_bases
};

struct btlbf_filter
{
	btlbf_ctx* ctx = nullptr;
	int kind = BTLBF_BLOOM;
	uint64_t size = 0, bytes = 0, cap = 0;
	unsigned threshold = 1;
	uint8_t* d_data = nullptr;
	bool owned = false;
	bool bitvector = false; // BTLBF_BITVECTOR: a BLOOM filter of any size in 64-bit words, without a file format
	HashCfg hc;
	// ordered-update state (lazy)
	uint32_t* d_touched = nullptr;
	uint32_t* d_contended = nullptr;
	uint32_t resv_log2 = 0;
	uint32_t* d_pending[2] = { nullptr, nullptr };
	uint32_t pending_cap = 0;
	uint64_t* d_pending_slots = nullptr; // [pending_cap * h] slots of deferred windows (when that fits 1 GiB)
	uint32_t pending_slots_h = 0, pending_slots_cap = 0;
	uint64_t* d_list_resv = nullptr;
	uint32_t list_log2 = 0;
	uint32_t epoch = 0;
	uint32_t* d_ord = nullptr; // device words of the cooperative path: [0..1] list counters, [2] epoch, [3] rounds, [4] deferred
	uint64_t deferred_total = 0, rounds_total = 0;
	bool wrapped = false; // caller-owned memory: nothing is ever left parked when a call returns
	// Queue of the legacy per-k-mer interface: updates that return nothing (insert / incrementAll / incrementMin
	// without `found`) are collected in pinned host memory and applied by one kernel per kHashQueue k-mers, or as
	// soon as anything reads or writes the filter (joined()).  Order inside the queue is the call order.
	uint64_t* hq = nullptr; // pinned: hq_n x h hash values
	uint64_t hq_n = 0;
	int hq_op = -1;
	std::vector<TlQueue*> tlq; // the host threads' private queues in front of hq (registered under the context lock)
	uint64_t uid = 0;          // never reused: keys the threads' queue caches
};
constexpr uint64_t kHashQueue = 1u << 16;

// ---------------------------------------------------------------- small helpers
#define LOCKED(ctx) std::lock_guard<std::recursive_mutex> lock_((ctx)->mu)

static int use(btlbf_ctx* ctx)
{
	if (!ctx)
		return fail(BTLBF_ERR_ARG, "null context");
	CU(cudaSetDevice(ctx->device));
	return BTLBF_OK;
}

// The active stream, ordered after all background (aux-stream) work: every operation that reads or
// writes filter contents on the active stream goes through this.
static int settle(btlbf_ctx* ctx);
static int hq_flush(btlbf_ctx* ctx);

static int tl_drain_all(btlbf_ctx* ctx);
static void tl_discard(btlbf_filter* f, bool destroy);

static cudaStream_t joined(btlbf_ctx* ctx)
{
	if (!ctx->in_hq_flush && ctx->tl_pending.load(std::memory_order_acquire) != 0) {
		int rc = tl_drain_all(ctx);
		if (rc != BTLBF_OK && !ctx->settle_error)
			ctx->settle_error = rc;
	}
	if (ctx->hq_owner && !ctx->in_hq_flush) {
		int rc = hq_flush(ctx);
		if (rc != BTLBF_OK && !ctx->settle_error)
			ctx->settle_error = rc;
	}
	if (ctx->acc.f)
		settle(ctx); // a failure is kept in ctx->settle_error
	if (ctx->aux_pending) {
		cudaError_t e = cudaStreamWaitEvent(ctx->active, ctx->ev_aux_last, 0);
		if (e != cudaSuccess && !ctx->settle_error)
			ctx->settle_error = BTLBF_ERR_CUDA;
		ctx->aux_pending = false;
	}
	return ctx->active;
}

static int take_settle_error(btlbf_ctx* ctx);

// joined() for entry points that report: the deferred work is queued on *s, or the call fails with the
// reason the deferred pass could not be launched (the parked k-mers are lost in that case)
static int join(btlbf_ctx* ctx, cudaStream_t* s)
{
	*s = joined(ctx);
	return take_settle_error(ctx);
}

static int ensure(DevBuf& b, size_t bytes)
{
	if (bytes <= b.cap)
		return BTLBF_OK;
	if (b.p)
		CU(cudaFree(b.p));
	b.p = nullptr;
	b.cap = 0;
	size_t want = (bytes + 255) / 256 * 256;
	CU(cudaMalloc(&b.p, want));
	b.cap = want;
	return BTLBF_OK;
}

static void release(DevBuf& b)
{
	if (b.p)
		cudaFree(b.p);
	b.p = nullptr;
	b.cap = 0;
}

static int launch(btlbf_ctx* ctx, SeqOp op, const SeqParams& P, cudaStream_t s)
{
	if (P.n_windows == 0)
		return BTLBF_OK;
	cudaError_t e = launch_seq(op, P, s);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "kernel launch (op %d) failed: %s", (int)op, cudaGetErrorString(e));
	ctx->launches++;
	return BTLBF_OK;
}

static void hashcfg_free(HashCfg& hc)
{
	if (hc.d_st_tab)
		cudaFree(hc.d_st_tab);
	if (hc.d_st_dc)
		cudaFree(hc.d_st_dc);
	hc.d_st_tab = nullptr;
	hc.d_st_dc = nullptr;
}

// Fills hc.proto with the launch-invariant hashing constants (host_params.hpp) and uploads the
// spaced-seed tables.
static int hashcfg_init(HashCfg& hc, unsigned k, unsigned h, const char* const* seeds, unsigned n_seeds,
                        unsigned h2)
{
	hashcfg_free(hc);
	HostSeedTables t;
	std::string err = build_hash_proto(hc.proto, t, k, h, seeds, n_seeds, h2);
	if (!err.empty())
		return fail(BTLBF_ERR_ARG, "%s", err.c_str());
	hc.k = hc.proto.k;
	hc.h = hc.proto.h;
	hc.n_seeds = hc.proto.n_seeds;
	hc.h2 = hc.proto.h2;
	if (hc.n_seeds) {
		CU(cudaMalloc(&hc.d_st_tab, t.tab.size() * 8));
		CU(cudaMemcpy(hc.d_st_tab, t.tab.data(), t.tab.size() * 8, cudaMemcpyHostToDevice));
		CU(cudaMalloc(&hc.d_st_dc, (t.dc.size() + 1) * 2));
		if (!t.dc.empty())
			CU(cudaMemcpy(hc.d_st_dc, t.dc.data(), t.dc.size() * 2, cudaMemcpyHostToDevice));
		hc.proto.st_tab = hc.d_st_tab;
		hc.proto.st_dc = hc.d_st_dc;
	}
	return BTLBF_OK;
}

static SeqParams filter_params(btlbf_filter* f)
{
	SeqParams P = f->hc.proto;
	P.filter = f->d_data;
	P.fm = make_fastmod(f->size);
	P.threshold = f->threshold;
	P.force_generic = (uint32_t)f->ctx->force_generic;
	P.query_mode = (uint32_t)f->ctx->query_mode;
	P.ungrouped_commit = (uint32_t)f->ctx->ungrouped_commit;
	return P;
}

// the part of chunk C that starts at its window b0 (a multiple of 32): input pointers and positions of P
static void advance_input(SeqParams& P, const SeqParams& C, uint64_t b0)
{
	P.bases = C.bases + (C.packed ? b0 >> 2 : b0);
	P.invalid = C.invalid ? C.invalid + (b0 >> 3) : nullptr;
	P.n_bases = C.n_bases > b0 ? C.n_bases - b0 : 0;
	P.base0 = C.base0 + b0;
}

// ---------------------------------------------------------------- misc entry points
extern "C" int btlbf_set_error(int code, const char* msg) // used by ingest.cu
{
	return fail(code, "%s", msg ? msg : "");
}

extern "C" int btlbf_filter_ctx(btlbf_filter* f, btlbf_ctx** ctx)
{
	if (!f || !ctx)
		return fail(BTLBF_ERR_ARG, "null argument");
	*ctx = f->ctx;
	return BTLBF_OK;
}

extern "C" const char* btlbf_last_error(void)
{
	return g_err;
}

extern "C" int btlbf_version(void)
{
	return BTLBF_VERSION;
}

extern "C" int btlbf_device_count(int* count)
{
	if (!count)
		return fail(BTLBF_ERR_ARG, "null argument");
	*count = 0;
	CU(cudaGetDeviceCount(count));
	return BTLBF_OK;
}

// ---------------------------------------------------------------- context
extern "C" int btlbf_ctx_create(int device, btlbf_ctx** out)
{
	if (!out)
		return fail(BTLBF_ERR_ARG, "null argument");
	*out = nullptr;
	int n = 0;
	CU(cudaGetDeviceCount(&n));
	if (device < 0 || device >= n)
		return fail(BTLBF_ERR_CUDA, "device %d not available (%d CUDA devices); there is no CPU fallback", device, n);
	cudaDeviceProp prop;
	CU(cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10)
		return fail(BTLBF_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device,
		            prop.major, prop.minor);
	CU(cudaSetDevice(device));
	btlbf_ctx* ctx = new (std::nothrow) btlbf_ctx();
	if (!ctx)
		return fail(BTLBF_ERR_NOMEM, "out of host memory");
	ctx->device = device;
	cudaError_t e = cudaStreamCreateWithFlags(&ctx->own, cudaStreamNonBlocking);
	if (e == cudaSuccess)
		e = cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking);
	if (e == cudaSuccess)
		e = cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking);
	if (e == cudaSuccess)
		e = cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking);
	if (e == cudaSuccess)
		e = cudaEventCreateWithFlags(&ctx->ev_switch, cudaEventDisableTiming);
	for (int i = 0; i < 2 && e == cudaSuccess; i++) {
		e = cudaEventCreateWithFlags(&ctx->ev_bin_done[i], cudaEventDisableTiming);
		if (e == cudaSuccess)
			e = cudaEventCreateWithFlags(&ctx->ev_apply_done[i], cudaEventDisableTiming);
		if (e == cudaSuccess)
			e = cudaEventCreateWithFlags(&ctx->ev_qp1[i], cudaEventDisableTiming);
		if (e == cudaSuccess)
			e = cudaEventCreateWithFlags(&ctx->ev_qp2[i], cudaEventDisableTiming);
	}
	if (e == cudaSuccess)
		e = cudaMalloc(&ctx->d_scalars, 16 * sizeof(unsigned long long));
	if (e == cudaSuccess)
		e = cudaMemset(ctx->d_scalars, 0, 16 * sizeof(unsigned long long));
	if (e == cudaSuccess)
		e = cudaHostAlloc(&ctx->h_scalars, 16 * sizeof(unsigned long long), cudaHostAllocDefault);
	for (int i = 0; i < kTickets && e == cudaSuccess; i++) {
		e = cudaEventCreateWithFlags(&ctx->ticket[i].done, cudaEventDisableTiming);
		if (e == cudaSuccess)
			e = cudaMalloc(&ctx->ticket[i].d_stats, 16);
	}
	for (int i = 0; i < 2 && e == cudaSuccess; i++) {
		e = cudaEventCreateWithFlags(&ctx->slot[i].ev_h2d, cudaEventDisableTiming);
		if (e == cudaSuccess)
			e = cudaEventCreateWithFlags(&ctx->slot[i].ev_kernel, cudaEventDisableTiming);
		if (e == cudaSuccess)
			e = cudaEventCreateWithFlags(&ctx->slot[i].ev_d2h, cudaEventDisableTiming);
	}
	if (e != cudaSuccess) {
		btlbf_ctx_destroy(ctx);
		return fail(BTLBF_ERR_CUDA, "context creation failed: %s", cudaGetErrorString(e));
	}
	ctx->active = ctx->own;
	*out = ctx;
	return BTLBF_OK;
}

extern "C" int btlbf_ctx_destroy(btlbf_ctx* ctx)
{
	if (!ctx)
		return BTLBF_OK;
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	for (int i = 0; i < 2; i++) {
		Slot& s = ctx->slot[i];
		release(s.bases); release(s.invalid); release(s.hit); release(s.valid); release(s.counts); release(s.hashes); release(s.strands);
		if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
		if (s.ev_kernel) cudaEventDestroy(s.ev_kernel);
		if (s.ev_d2h) cudaEventDestroy(s.ev_d2h);
	}
	for (int i = 0; i < kTickets; i++) {
		release(ctx->ticket[i].offsets);
		if (ctx->ticket[i].d_stats) cudaFree(ctx->ticket[i].d_stats);
		if (ctx->ticket[i].done) cudaEventDestroy(ctx->ticket[i].done);
	}
	for (int i = 0; i < 2; i++) {
		release(ctx->ibin_items[i]);
		release(ctx->ibin_counts[i]);
		if (ctx->ev_bin_done[i]) cudaEventDestroy(ctx->ev_bin_done[i]);
		if (ctx->ev_apply_done[i]) cudaEventDestroy(ctx->ev_apply_done[i]);
		if (ctx->ev_qp1[i]) cudaEventDestroy(ctx->ev_qp1[i]);
		if (ctx->ev_qp2[i]) cudaEventDestroy(ctx->ev_qp2[i]);
		release(ctx->qbin_items[i]);
		release(ctx->qbin_counts[i]);
	}
	release(ctx->ibin2_items);
	release(ctx->ibin2_counts);
	if (ctx->aux) cudaStreamDestroy(ctx->aux);
	if (ctx->ev_switch) cudaEventDestroy(ctx->ev_switch);
	release(ctx->q_hit);
	release(ctx->q_valid);
	if (ctx->d_scalars) cudaFree(ctx->d_scalars);
	if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
	for (int i = 0; i < 2; i++) {
		if (ctx->slot[i].pk_codes) cudaFreeHost(ctx->slot[i].pk_codes);
		if (ctx->slot[i].pk_inv) cudaFreeHost(ctx->slot[i].pk_inv);
	}
	if (ctx->own) cudaStreamDestroy(ctx->own);
	if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
	if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
	delete ctx;
	return BTLBF_OK;
}

extern "C" int btlbf_ctx_set_stream(btlbf_ctx* ctx, void* cuda_stream)
{
	TRY(use(ctx));
	LOCKED(ctx);
	cudaStream_t next = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own;
	if (ctx->aux_pending) // both the old and the new stream are ordered after the background work
		CU(cudaStreamWaitEvent(next, ctx->ev_aux_last, 0));
	// Everything queued so far -- including the deferred pass 2 of a partitioned build, which joined() launches on
	// the OLD stream right here -- happens before anything the caller queues on the new stream.
	cudaStream_t prev;
	TRY(join(ctx, &prev));
	if (next != prev) {
		CU(cudaEventRecord(ctx->ev_switch, prev));
		CU(cudaStreamWaitEvent(next, ctx->ev_switch, 0));
	}
	ctx->active = next;
	return BTLBF_OK;
}

extern "C" int btlbf_ctx_sync(btlbf_ctx* ctx)
{
	TRY(use(ctx));
	LOCKED(ctx);
	CU(cudaStreamSynchronize(ctx->copy_in));
	CU(cudaStreamSynchronize(joined(ctx)));
	CU(cudaStreamSynchronize(ctx->copy_out));
	for (int i = 0; i < kTickets; i++)
		ctx->ticket[i].in_use = false;
	return take_settle_error(ctx);
}

// a failed pass-2 launch inside joined() (which cannot report it) is reported by the next sync / flush
static int take_settle_error(btlbf_ctx* ctx)
{
	int rc = ctx->settle_error;
	ctx->settle_error = 0;
	return rc == 0 ? BTLBF_OK : fail(rc, "a deferred pass of the partitioned build failed to launch");
}

extern "C" int btlbf_ctx_flush(btlbf_ctx* ctx)
{
	TRY(use(ctx));
	LOCKED(ctx);
	joined(ctx);
	return take_settle_error(ctx);
}

extern "C" int btlbf_ctx_aux_stream(btlbf_ctx* ctx, void** cuda_stream)
{
	if (!ctx || !cuda_stream)
		return fail(BTLBF_ERR_ARG, "null argument");
	*cuda_stream = (void*)ctx->aux;
	return BTLBF_OK;
}

extern "C" int btlbf_ctx_counter(btlbf_ctx* ctx, const char* name, uint64_t* value)
{
	if (!ctx || !name || !value)
		return fail(BTLBF_ERR_ARG, "null argument");
	std::string k(name);
	if (k == "launches") *value = ctx->launches;
	else if (k == "binned_launches") *value = ctx->binned_launches;
	else if (k == "two_level_passes") *value = ctx->two_level_passes;
	else
		return fail(BTLBF_ERR_ARG, "unknown counter '%s'", name);
	return BTLBF_OK;
}

extern "C" int btlbf_ctx_launch_count(btlbf_ctx* ctx, uint64_t* count)
{
	if (!ctx || !count)
		return fail(BTLBF_ERR_ARG, "null argument");
	*count = ctx->launches;
	return BTLBF_OK;
}

extern "C" int btlbf_ctx_set_option(btlbf_ctx* ctx, const char* key, int64_t value)
{
	if (!ctx || !key)
		return fail(BTLBF_ERR_ARG, "null argument");
	LOCKED(ctx);
	std::string k(key);
	if (k == "force_generic")
		ctx->force_generic = value != 0;
	else if (k == "query_mode")
		ctx->query_mode = value != 0;
	else if (k == "ungrouped_commit")
		ctx->ungrouped_commit = value != 0;
	else if (k == "chunk_bases") {
		if (value < kTile || value > ((int64_t)1 << 31))
			return fail(BTLBF_ERR_ARG, "chunk_bases out of range");
		ctx->chunk_bases = value / kTile * kTile;
	} else if (k == "cbf_batch") {
		if (value < kTile || value > ((int64_t)1 << 30))
			return fail(BTLBF_ERR_ARG, "cbf_batch out of range");
		ctx->cbf_batch = value / kTile * kTile;
	} else if (k == "resv_log2") {
		if (value < 10 || value > 32)
			return fail(BTLBF_ERR_ARG, "resv_log2 out of range");
		ctx->resv_log2 = value;
	} else if (k == "list_log2") {
		if (value < 6 || value > 28)
			return fail(BTLBF_ERR_ARG, "list_log2 out of range");
		ctx->list_log2 = value;
	} else if (k == "l2_fetch_granularity") {
		// device-wide hint: bytes fetched from HBM per L2 sector miss (32, 64 or 128).  Random 32-byte
		// sector probes waste bandwidth at the larger settings.
		if (value != 32 && value != 64 && value != 128)
			return fail(BTLBF_ERR_ARG, "l2_fetch_granularity must be 32, 64 or 128");
		TRY(use(ctx));
		CU(cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)value));
	} else if (k == "bin_mode") {
		if (value < -1 || value > 1)
			return fail(BTLBF_ERR_ARG, "bin_mode must be -1, 0 or 1");
		ctx->bin_mode = value;
	} else if (k == "overlap") {
		ctx->overlap = value != 0;
	} else if (k == "bin_query_mode") {
		if (value < -1 || value > 1)
			return fail(BTLBF_ERR_ARG, "bin_query_mode must be -1, 0 or 1");
		ctx->bin_query_mode = value;
	} else if (k == "bin_part_log2") {
		if (value < 8 || value > 31)
			return fail(BTLBF_ERR_ARG, "bin_part_log2 out of range [8,31]");
		ctx->bin_part_log2 = value;
	} else if (k == "bin_slack_pct") {
		if (value < 0 || value > 1000)
			return fail(BTLBF_ERR_ARG, "bin_slack_pct out of range");
		ctx->bin_slack_pct = value;
	} else if (k == "bin_accum_bytes") {
		if (value < 0)
			return fail(BTLBF_ERR_ARG, "bin_accum_bytes out of range");
		ctx->bin_accum_bytes = value;
	} else if (k == "peer_unroll") {
		ctx->peer_unroll = value;
	} else if (k == "peer_mode") {
		ctx->peer_mode = value;
	} else if (k == "peer_grid") {
		ctx->peer_grid = value < 0 ? 0 : value;
	} else if (k == "mm_unroll") {
		ctx->mm_unroll = value;
	} else if (k == "mm_grid") {
		ctx->mm_grid = value < 0 ? 0 : value;
	} else if (k == "wrap_accumulate") {
		ctx->wrap_accumulate = value != 0;
	} else if (k == "host_pack") {
		ctx->host_pack = value != 0;
	} else if (k == "host_pack_threads") {
		if (value < 1 || value > 64)
			return fail(BTLBF_ERR_ARG, "host_pack_threads out of range [1,64]");
		ctx->host_pack_threads = value;
	} else if (k == "query_adaptive") {
		ctx->query_adaptive = value != 0;
	} else if (k == "query_adaptive_min_tiles") {
		if (value < 1)
			return fail(BTLBF_ERR_ARG, "query_adaptive_min_tiles must be >= 1");
		ctx->query_adaptive_min_tiles = value;
	} else if (k == "query_adaptive_pct") {
		if (value < 0 || value > 101)
			return fail(BTLBF_ERR_ARG, "query_adaptive_pct out of range [0,101]");
		ctx->query_adaptive_pct = value;
	} else if (k == "query_chunk_factor") {
		if (value < 1 || value > 16)
			return fail(BTLBF_ERR_ARG, "query_chunk_factor out of range [1,16]");
		ctx->query_chunk_factor = value;
	} else if (k == "query_sub") {
		if (value < 0 || value > 64)
			return fail(BTLBF_ERR_ARG, "query_sub out of range [0,64]");
		ctx->query_sub = value;
	} else if (k == "query_p1_ctas") {
		if (value < 0 || value > 32)
			return fail(BTLBF_ERR_ARG, "query_p1_ctas out of range [0,32]");
		ctx->query_p1_ctas = value;
	} else if (k == "query_probe_unroll") {
		if (value != 0 && value != 1 && value != 2 && value != 4)
			return fail(BTLBF_ERR_ARG, "query_probe_unroll must be 0, 1, 2 or 4");
		ctx->query_probe_unroll = value;
	} else if (k == "probe_ld") {
		ctx->probe_ld = value < 0 || value > 2 ? 0 : value;
	} else if (k == "probe_carveout") {
		ctx->probe_carveout = value < -1 || value > 1 ? -1 : value;
	} else if (k == "bin_prefetch") {
		if (value < -1 || value > 2)
			return fail(BTLBF_ERR_ARG, "bin_prefetch must be -1, 0, 1 or 2");
		ctx->bin_prefetch = value;
	} else if (k == "bin_two_level") {
		ctx->bin_two_level = value != 0;
	} else if (k == "bin_two_level_min") {
		ctx->bin_two_level_min = value < 0 ? 0 : value;
	} else if (k == "bin_kernel") {
		ctx->bin_kernel = value != 0;
	} else if (k == "bin_max_parts") {
		if (value < 1 || value > 1024)
			return fail(BTLBF_ERR_ARG, "bin_max_parts out of range [1,1024]");
		ctx->bin_max_parts = value;
	} else if (k == "ordered_coop") {
		ctx->ordered_coop = value != 0;
	} else if (k == "drain_threshold") {
		if (value < 0)
			return fail(BTLBF_ERR_ARG, "drain_threshold out of range");
		ctx->drain_threshold = value;
	} else
		return fail(BTLBF_ERR_ARG, "unknown option '%s'", key);
	return BTLBF_OK;
}

// ---------------------------------------------------------------- filters
static int filter_make(btlbf_ctx* ctx, int kind, uint64_t size, unsigned h, unsigned k, unsigned threshold,
                       void* wrap_ptr, uint64_t wrap_cap, btlbf_filter** out)
{
	if (!out)
		return fail(BTLBF_ERR_ARG, "null argument");
	*out = nullptr;
	TRY(use(ctx));
	LOCKED(ctx);
	if (kind != BTLBF_BLOOM && kind != BTLBF_COUNTING8 && kind != BTLBF_BITVECTOR)
		return fail(BTLBF_ERR_ARG, "unknown filter kind %d", kind);
	if (size == 0)
		return fail(BTLBF_ERR_ARG, "filter size must be > 0");
	const bool bitvector = kind == BTLBF_BITVECTOR;
	if (bitvector)
		kind = BTLBF_BLOOM; // same bits, same kernels
	if (kind == BTLBF_BLOOM && !bitvector && size % 8 != 0) // BloomFilter.hpp:389-394
		return fail(BTLBF_ERR_ARG, "Filter Size \"%llu\" is not a multiple of 8", (unsigned long long)size);
	btlbf_filter* f = new (std::nothrow) btlbf_filter();
	if (f) {
		static std::atomic<uint64_t> next_uid{ 1 };
		f->uid = next_uid.fetch_add(1);
	}
	if (!f)
		return fail(BTLBF_ERR_NOMEM, "out of host memory");
	f->ctx = ctx;
	f->kind = kind;
	f->size = size;
	f->bitvector = bitvector;
	f->bytes = bitvector ? (size + 63) / 64 * 8 : kind == BTLBF_BLOOM ? size / 8 : size;
	f->threshold = threshold;
	int rc = hashcfg_init(f->hc, k, h, nullptr, 0, 0);
	if (rc != BTLBF_OK) {
		delete f;
		return rc;
	}
	uint64_t need = (f->bytes + 15) / 16 * 16;
	if (wrap_ptr) {
		if (wrap_cap < need || ((uintptr_t)wrap_ptr & 15u)) {
			delete f;
			return fail(BTLBF_ERR_ARG, "wrapped memory must be 16-byte aligned and hold %llu bytes",
			            (unsigned long long)need);
		}
		f->d_data = (uint8_t*)wrap_ptr;
		f->cap = wrap_cap;
		f->owned = false;
		f->wrapped = true;
	} else {
		cudaError_t e = cudaMalloc(&f->d_data, need);
		if (e == cudaSuccess)
			e = cudaMemsetAsync(f->d_data, 0, need, ctx->active);
		if (e != cudaSuccess) {
			delete f;
			return fail(e == cudaErrorMemoryAllocation ? BTLBF_ERR_NOMEM : BTLBF_ERR_CUDA,
			            "allocating a %llu-byte filter failed: %s", (unsigned long long)need, cudaGetErrorString(e));
		}
		f->cap = need;
		f->owned = true;
	}
	*out = f;
	return BTLBF_OK;
}

extern "C" int btlbf_filter_create(btlbf_ctx* ctx, int kind, uint64_t size, unsigned hash_num, unsigned kmer_size,
                                   unsigned threshold, btlbf_filter** filter)
{
	return filter_make(ctx, kind, size, hash_num, kmer_size, threshold, nullptr, 0, filter);
}

extern "C" int btlbf_filter_wrap(btlbf_ctx* ctx, int kind, uint64_t size, unsigned hash_num, unsigned kmer_size,
                                 unsigned threshold, void* device_ptr, uint64_t capacity_bytes, btlbf_filter** filter)
{
	if (!device_ptr)
		return fail(BTLBF_ERR_ARG, "null device pointer");
	return filter_make(ctx, kind, size, hash_num, kmer_size, threshold, device_ptr, capacity_bytes, filter);
}

extern "C" int btlbf_filter_destroy(btlbf_filter* f)
{
	if (!f)
		return BTLBF_OK;
	btlbf_ctx* ctx = f->ctx;
	LOCKED(ctx);
	cudaSetDevice(f->ctx->device);
	if (f->ctx->acc.f == f)
		f->ctx->acc.f = nullptr; // parked k-mers of a filter that is going away
	tl_discard(f, true);
	if (ctx->hq_owner == f)
		ctx->hq_owner = nullptr;
	if (f->hq)
		cudaFreeHost(f->hq);
	cudaStreamSynchronize(joined(f->ctx));
	if (f->owned && f->d_data)
		cudaFree(f->d_data);
	hashcfg_free(f->hc);
	if (f->d_touched) cudaFree(f->d_touched);
	if (f->d_contended) cudaFree(f->d_contended);
	if (f->d_pending[0]) cudaFree(f->d_pending[0]);
	if (f->d_pending[1]) cudaFree(f->d_pending[1]);
	if (f->d_pending_slots) cudaFree(f->d_pending_slots);
	if (f->d_list_resv) cudaFree(f->d_list_resv);
	if (f->d_ord) cudaFree(f->d_ord);
	delete f;
	return BTLBF_OK;
}

extern "C" int btlbf_filter_clear(btlbf_filter* f)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	TRY(use(f->ctx));
	LOCKED(f->ctx);
	if (f->ctx->acc.f == f)
		f->ctx->acc.f = nullptr; // k-mers still parked in the partition buckets vanish with the rest
	tl_discard(f, false); // ... and so do queued per-k-mer updates
	if (f->ctx->hq_owner == f) {
		f->ctx->hq_owner = nullptr;
		f->hq_n = 0;
	}
	cudaStream_t s;
	TRY(join(f->ctx, &s));
	CU(cudaMemsetAsync(f->d_data, 0, (f->bytes + 15) / 16 * 16, s));
	return BTLBF_OK;
}

extern "C" int btlbf_filter_info(btlbf_filter* f, int* kind, uint64_t* size, uint64_t* size_bytes, unsigned* hash_num,
                                 unsigned* kmer_size, unsigned* threshold)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	if (kind) *kind = f->bitvector ? BTLBF_BITVECTOR : f->kind;
	if (size) *size = f->size;
	if (size_bytes) *size_bytes = f->bytes;
	if (hash_num) *hash_num = f->hc.h;
	if (kmer_size) *kmer_size = f->hc.k;
	if (threshold) *threshold = f->threshold;
	return BTLBF_OK;
}

extern "C" int btlbf_filter_set_threshold(btlbf_filter* f, unsigned threshold)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	f->threshold = threshold;
	return BTLBF_OK;
}

extern "C" int btlbf_filter_upload(btlbf_filter* f, const void* host, uint64_t nbytes)
{
	if (!f || (!host && nbytes))
		return fail(BTLBF_ERR_ARG, "null argument");
	if (nbytes != f->bytes)
		return fail(BTLBF_ERR_ARG, "upload of %llu bytes into a %llu-byte filter", (unsigned long long)nbytes,
		            (unsigned long long)f->bytes);
	TRY(use(f->ctx));
	LOCKED(f->ctx);
	if (f->ctx->acc.f == f)
		f->ctx->acc.f = nullptr; // overwritten anyway
	tl_discard(f, false);
	if (f->ctx->hq_owner == f) {
		f->ctx->hq_owner = nullptr;
		f->hq_n = 0;
	}
	cudaStream_t s;
	TRY(join(f->ctx, &s));
	CU(cudaMemcpyAsync(f->d_data, host, nbytes, cudaMemcpyHostToDevice, s));
	uint64_t pad = (f->bytes + 15) / 16 * 16 - f->bytes;
	if (pad)
		CU(cudaMemsetAsync(f->d_data + f->bytes, 0, pad, f->ctx->active));
	CU(cudaStreamSynchronize(f->ctx->active));
	return BTLBF_OK;
}

extern "C" int btlbf_filter_download(btlbf_filter* f, void* host, uint64_t nbytes)
{
	if (!f || (!host && nbytes))
		return fail(BTLBF_ERR_ARG, "null argument");
	if (nbytes != f->bytes)
		return fail(BTLBF_ERR_ARG, "download of %llu bytes from a %llu-byte filter", (unsigned long long)nbytes,
		            (unsigned long long)f->bytes);
	TRY(use(f->ctx));
	LOCKED(f->ctx);
	cudaStream_t s;
	TRY(join(f->ctx, &s));
	CU(cudaMemcpyAsync(host, f->d_data, nbytes, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	return BTLBF_OK;
}

extern "C" int btlbf_filter_device_ptr(btlbf_filter* f, void** device_ptr, uint64_t* nbytes)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	LOCKED(f->ctx);
	if (f->ctx->acc.f == f || f->ctx->aux_pending || f->ctx->hq_owner == f || f->ctx->tl_pending.load() != 0) { // parked k-mers reach the filter now, in stream order
		TRY(use(f->ctx));
		cudaStream_t s;
		TRY(join(f->ctx, &s));
	}
	if (device_ptr) *device_ptr = f->d_data;
	if (nbytes) *nbytes = f->bytes;
	return BTLBF_OK;
}

static int reduce_count(btlbf_filter* f, int mode, unsigned threshold, uint64_t* count)
{
	if (!f || !count)
		return fail(BTLBF_ERR_ARG, "null argument");
	btlbf_ctx* ctx = f->ctx;
	TRY(use(ctx));
	LOCKED(ctx);
	cudaStream_t s0;
	TRY(join(ctx, &s0));
	CU(cudaMemsetAsync(ctx->d_scalars + 2, 0, 8, s0));
	cudaError_t e = launch_popcount(f->d_data, f->bytes, mode, threshold, ctx->d_scalars + 2, ctx->active);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "popcount launch failed: %s", cudaGetErrorString(e));
	ctx->launches++;
	CU(cudaMemcpyAsync(ctx->h_scalars + 2, ctx->d_scalars + 2, 8, cudaMemcpyDeviceToHost, ctx->active));
	CU(cudaStreamSynchronize(ctx->active));
	*count = ctx->h_scalars[2];
	return BTLBF_OK;
}

extern "C" int btlbf_filter_popcount(btlbf_filter* f, uint64_t* count)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	return reduce_count(f, f->kind == BTLBF_BLOOM ? 0 : 1, 0, count);
}

extern "C" int btlbf_filter_count_ge(btlbf_filter* f, unsigned threshold, uint64_t* count)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	if (f->kind != BTLBF_COUNTING8)
		return fail(BTLBF_ERR_STATE, "count_ge needs a counting filter");
	if (threshold > 255) {
		if (count) *count = 0;
		return count ? BTLBF_OK : fail(BTLBF_ERR_ARG, "null argument");
	}
	if (threshold == 0) {
		if (count) *count = f->size;
		return count ? BTLBF_OK : fail(BTLBF_ERR_ARG, "null argument");
	}
	return reduce_count(f, 2, threshold, count);
}

extern "C" int btlbf_filter_set_seeds(btlbf_filter* f, const char* const* seeds, unsigned n_seeds, unsigned h2)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	TRY(use(f->ctx));
	LOCKED(f->ctx);
	{
		cudaStream_t s; // parked / queued k-mers were hashed (or will be applied) under the old seeds
		TRY(join(f->ctx, &s));
		CU(cudaStreamSynchronize(s));
	}
	unsigned k = f->hc.k, h = f->hc.h;
	if (n_seeds == 0)
		return hashcfg_init(f->hc, k, h, nullptr, 0, 0);
	if ((uint64_t)n_seeds * h2 != h)
		return fail(BTLBF_ERR_ARG, "spaced seeds need hash_num == n_seeds*h2 (%u != %u*%u)", h, n_seeds, h2);
	HashCfg tmp;
	int rc = hashcfg_init(tmp, k, h, seeds, n_seeds, h2);
	if (rc != BTLBF_OK) {
		hashcfg_free(tmp);
		return rc;
	}
	hashcfg_free(f->hc);
	f->hc = tmp;
	return BTLBF_OK;
}

extern "C" int btlbf_filter_merge_from_device(btlbf_filter* f, const void* src_device, uint64_t nbytes)
{
	if (!f || !src_device)
		return fail(BTLBF_ERR_ARG, "null argument");
	if (nbytes != f->bytes)
		return fail(BTLBF_ERR_ARG, "merge of %llu bytes into a %llu-byte filter", (unsigned long long)nbytes,
		            (unsigned long long)f->bytes);
	if ((uintptr_t)src_device & 15u)
		return fail(BTLBF_ERR_ARG, "merge source must be 16-byte aligned");
	TRY(use(f->ctx));
	LOCKED(f->ctx);
	cudaStream_t s;
	TRY(join(f->ctx, &s));
	cudaError_t e = launch_merge(f->d_data, src_device, nbytes, f->kind == BTLBF_COUNTING8, s);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "merge launch failed: %s", cudaGetErrorString(e));
	f->ctx->launches++;
	return BTLBF_OK;
}

extern "C" int btlbf_merge_device_buffers(btlbf_ctx* ctx, int kind, void* dst_device, const void* src_device,
                                          uint64_t nbytes)
{
	TRY(use(ctx));
	if (kind != BTLBF_BLOOM && kind != BTLBF_COUNTING8)
		return fail(BTLBF_ERR_ARG, "unknown filter kind %d", kind);
	if (!dst_device || !src_device)
		return fail(BTLBF_ERR_ARG, "null argument");
	if (((uintptr_t)dst_device | (uintptr_t)src_device) & 15u)
		return fail(BTLBF_ERR_ARG, "merge buffers must be 16-byte aligned");
	LOCKED(ctx);
	cudaStream_t s;
	TRY(join(ctx, &s));
	cudaError_t e = launch_merge(dst_device, src_device, nbytes, kind == BTLBF_COUNTING8, s);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "merge launch failed: %s", cudaGetErrorString(e));
	ctx->launches++;
	return BTLBF_OK;
}

// ---------------------------------------------------------------- fused multi-GPU merge (peer memory)
extern "C" int btlbf_merge_slice(uint64_t nbytes, int world, int rank, uint64_t* lo, uint64_t* hi)
{
	if (world < 1 || rank < 0 || rank >= world || !lo || !hi)
		return fail(BTLBF_ERR_ARG, "bad world / rank");
	uint64_t nvec = (nbytes + 15) / 16;
	uint64_t per = (nvec + (uint64_t)world - 1) / (uint64_t)world;
	uint64_t a = (uint64_t)rank * per, b = a + per;
	a = a > nvec ? nvec : a;
	b = b > nvec ? nvec : b;
	*lo = a * 16;
	*hi = b * 16;
	return BTLBF_OK;
}

extern "C" int btlbf_ipc_export(btlbf_ctx* ctx, const void* device_ptr, void* handle64, uint64_t* offset)
{
	if (!device_ptr || !handle64 || !offset)
		return fail(BTLBF_ERR_ARG, "null argument");
	TRY(use(ctx));
	LOCKED(ctx);
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
	// the handle names a whole cudaMalloc allocation: find its base (the pointer may sit inside a block of
	// a caching allocator)
	cudaPointerAttributes attr;
	CU(cudaPointerGetAttributes(&attr, device_ptr));
	if (attr.type != cudaMemoryTypeDevice)
		return fail(BTLBF_ERR_ARG, "not a device pointer");
	// (driver entry point fetched through the runtime: the library does not link libcuda)
	typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
	void* fn = nullptr;
	cudaDriverEntryPointQueryResult qr;
	CU(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
	if (!fn || qr != cudaDriverEntryPointSuccess)
		return fail(BTLBF_ERR_CUDA, "cuMemGetAddressRange is not available");
	unsigned long long base = 0;
	size_t size = 0;
	int r = ((range_fn)fn)(&base, &size, (unsigned long long)(uintptr_t)device_ptr);
	if (r != 0)
		return fail(BTLBF_ERR_CUDA, "cuMemGetAddressRange failed (%d)", r);
	cudaIpcMemHandle_t h;
	CU(cudaIpcGetMemHandle(&h, (void*)(uintptr_t)base));
	memcpy(handle64, &h, 64);
	*offset = (uint64_t)((unsigned long long)(uintptr_t)device_ptr - base);
	return BTLBF_OK;
}

extern "C" int btlbf_ipc_open(btlbf_ctx* ctx, const void* handle64, void** mapped_base)
{
	if (!handle64 || !mapped_base)
		return fail(BTLBF_ERR_ARG, "null argument");
	TRY(use(ctx));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, 64);
	CU(cudaIpcOpenMemHandle(mapped_base, h, cudaIpcMemLazyEnablePeerAccess));
	return BTLBF_OK;
}

extern "C" int btlbf_ipc_close(btlbf_ctx* ctx, void* mapped_base)
{
	TRY(use(ctx));
	if (mapped_base)
		CU(cudaIpcCloseMemHandle(mapped_base));
	return BTLBF_OK;
}

extern "C" int btlbf_merge_peers(btlbf_ctx* ctx, int kind, void* const* bases, int world, int rank, uint64_t nbytes)
{
	if (!bases || world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
		return fail(BTLBF_ERR_ARG, "bad peer list (world %d, rank %d; at most %d peers)", world, rank, kMaxPeers);
	if (kind != BTLBF_BLOOM && kind != BTLBF_COUNTING8)
		return fail(BTLBF_ERR_ARG, "bad filter kind");
	TRY(use(ctx));
	LOCKED(ctx);
	PeerMergeParams M;
	memset(&M, 0, sizeof M);
	// start with the local copy, then the peers in ring order so that the ranks do not all hit one GPU at once
	for (int i = 0; i < world; i++) {
		void* b = bases[(rank + i) % world];
		if (!b || ((uintptr_t)b & 15u))
			return fail(BTLBF_ERR_ARG, "peer base %d is null or not 16-byte aligned", (rank + i) % world);
		M.base[i] = (uint8_t*)b;
	}
	M.world = (uint32_t)world;
	M.sat_add = kind == BTLBF_COUNTING8;
	M.unroll = (uint32_t)ctx->peer_unroll;
	M.grid = (uint32_t)ctx->peer_grid;
	M.mode = (uint32_t)ctx->peer_mode;
	TRY(btlbf_merge_slice(nbytes, world, rank, &M.lo, &M.hi));
	cudaStream_t s;
	TRY(join(ctx, &s));
	cudaError_t e = launch_peer_merge(M, s);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "peer merge launch failed: %s", cudaGetErrorString(e));
	ctx->launches++;
	return BTLBF_OK;
}

// The OR merge inside NVSwitch.  mc_base: a multicast (NVLS) address that maps the SAME byte range of every rank's
// partial filter (e.g. torch.distributed._symmetric_memory: rendezvous(...).multicast_ptr of the tensor the filter
// was wrapped around with btlbf_filter_wrap).  This rank reduces bytes btlbf_merge_slice(nbytes, world, rank) of all
// replicas with multimem.ld_reduce.or and broadcasts them with multimem.st.  The caller brackets the launches of
// all ranks with barriers, exactly as for btlbf_merge_peers.
extern "C" int btlbf_merge_multimem(btlbf_ctx* ctx, int kind, void* mc_base, int world, int rank, uint64_t nbytes)
{
	if (kind != BTLBF_BLOOM)
		return fail(BTLBF_ERR_ARG, "the in-switch merge is an OR: counting filters (saturating add) use btlbf_merge_peers");
	if (!mc_base || ((uintptr_t)mc_base & 15u) || world < 1 || rank < 0 || rank >= world)
		return fail(BTLBF_ERR_ARG, "bad multicast base / world / rank");
	TRY(use(ctx));
	LOCKED(ctx);
	uint64_t lo, hi;
	TRY(btlbf_merge_slice(nbytes, world, rank, &lo, &hi));
	cudaStream_t s;
	TRY(join(ctx, &s));
	cudaError_t e = launch_multimem_or(mc_base, lo, hi, (unsigned)ctx->mm_unroll, (unsigned)ctx->mm_grid, s);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "multimem merge launch failed: %s", cudaGetErrorString(e));
	ctx->launches++;
	return BTLBF_OK;
}

// Both at once (BLOOM, world 2 / 4 / 8): mm_pct per cent of this rank's byte range through the multicast mapping, the
// rest through the peer pointers, in one kernel.
extern "C" int btlbf_merge_hybrid(btlbf_ctx* ctx, int kind, void* mc_base, void* const* bases, int world, int rank,
                                  uint64_t nbytes, unsigned mm_pct)
{
	if (kind != BTLBF_BLOOM)
		return fail(BTLBF_ERR_ARG, "the in-switch merge is an OR: counting filters (saturating add) use btlbf_merge_peers");
	if (!mc_base || ((uintptr_t)mc_base & 15u) || !bases || (world != 2 && world != 4 && world != 8) || rank < 0 || rank >= world ||
	    mm_pct > 100)
		return fail(BTLBF_ERR_ARG, "bad multicast base / peer list / world (2, 4 or 8) / rank / percentage");
	TRY(use(ctx));
	LOCKED(ctx);
	PeerMergeParams M;
	memset(&M, 0, sizeof M);
	for (int i = 0; i < world; i++) {
		void* b = bases[(rank + i) % world];
		if (!b || ((uintptr_t)b & 15u))
			return fail(BTLBF_ERR_ARG, "peer base %d is null or not 16-byte aligned", (rank + i) % world);
		M.base[i] = (uint8_t*)b;
	}
	M.world = (uint32_t)world;
	M.grid = (uint32_t)ctx->peer_grid;
	TRY(btlbf_merge_slice(nbytes, world, rank, &M.lo, &M.hi));
	cudaStream_t s;
	TRY(join(ctx, &s));
	cudaError_t e = launch_hybrid_merge(M, mc_base, mm_pct, s);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "hybrid merge launch failed: %s", cudaGetErrorString(e));
	ctx->launches++;
	return BTLBF_OK;
}

// Pass 2 of the parked build in parts, so that a multi-GPU merge can follow the partitions as they are finished
// (pass 2 is partition-major).  Chunk `chunk` of `n_chunks` applies the parked k-mers of its share of the partitions on
// the active stream and reports the byte range [lo, hi) of the filter array those partitions cover (multiples of 16).
// Call with chunk = 0, 1, ..., n_chunks-1; after the last one nothing is parked.  When the parked work cannot be
// split (nothing parked for f, two-level or background pass 2, queued per-k-mer updates), chunk 0 applies everything
// and reports the whole array, the other chunks report empty ranges.  Any other call in between settles all of it
// (harmless: OR is idempotent); the remaining chunks then only report their ranges.
extern "C" int btlbf_filter_flush_parts(btlbf_filter* f, unsigned chunk, unsigned n_chunks, uint64_t* byte_lo, uint64_t* byte_hi)
{
	if (!f || !byte_lo || !byte_hi || n_chunks == 0 || chunk >= n_chunks)
		return fail(BTLBF_ERR_ARG, "bad filter / chunk");
	btlbf_ctx* ctx = f->ctx;
	TRY(use(ctx));
	LOCKED(ctx);
	*byte_lo = *byte_hi = 0;
	const uint64_t whole = (f->bytes + 15) / 16 * 16;
	if (chunk == 0) {
		const bool ok = ctx->acc.f == f && f->kind == BTLBF_BLOOM && !ctx->overlap && !ctx->bin_two_level && !ctx->hq_owner &&
		                ctx->tl_pending.load() == 0 && !ctx->aux_pending && ctx->acc.P.n_bins >= n_chunks && ctx->acc.P.bin_shift >= 7;
		ctx->parts.f = f;
		ctx->parts.n_chunks = n_chunks;
		ctx->parts.n_bins = ok ? ctx->acc.P.n_bins : 0;
		ctx->parts.shift = ok ? ctx->acc.P.bin_shift : 0;
		if (!ok) {
			cudaStream_t s;
			TRY(join(ctx, &s));
			*byte_hi = whole;
			return BTLBF_OK;
		}
	} else if (ctx->parts.f != f || ctx->parts.n_chunks != n_chunks) {
		return fail(BTLBF_ERR_STATE, "flush_parts: chunk %u without chunk 0 of the same sequence", chunk);
	}
	if (ctx->parts.n_bins == 0)
		return BTLBF_OK; // chunk 0 reported the whole array
	const uint32_t nb = ctx->parts.n_bins;
	const uint32_t p0 = (uint32_t)((uint64_t)nb * chunk / n_chunks), p1 = (uint32_t)((uint64_t)nb * (chunk + 1) / n_chunks);
	if (ctx->acc.f == f && p1 > p0) { // still parked: apply this chunk's partitions
		SeqParams P = ctx->acc.P;
		P.bin_part0 = p0;
		P.bin_part_count = p1 - p0;
		cudaError_t e = launch_apply_bins(P, ctx->active);
		if (e != cudaSuccess)
			return fail(BTLBF_ERR_CUDA, "partitioned build (pass 2, partitions %u..%u) launch failed: %s", p0, p1, cudaGetErrorString(e));
		ctx->launches++;
	}
	const uint64_t part_bytes = ((uint64_t)1 << ctx->parts.shift) >> 3;
	uint64_t lo = (uint64_t)p0 * part_bytes, hi = (uint64_t)p1 * part_bytes;
	if (chunk + 1 == n_chunks) {
		hi = whole;
		if (ctx->acc.f == f) { // the accumulation is closed, as settle() would have done
			CU(cudaEventRecord(ctx->ev_apply_done[ctx->acc.slot], ctx->active));
			ctx->slot_used[ctx->acc.slot] = true;
			ctx->acc.f = nullptr;
		}
		ctx->parts.f = nullptr;
	}
	*byte_lo = lo < whole ? lo : whole;
	*byte_hi = hi < whole ? hi : whole;
	return BTLBF_OK;
}

// btlbf_merge_peers for bytes [lo, hi) of the arrays only (multiples of 16), on the given stream, WITHOUT applying
// deferred work first: the caller has just flushed exactly that range (btlbf_filter_flush_parts) and the rest of the
// parked k-mers must stay parked.  This rank reduces its 1/world share of the range.
extern "C" int btlbf_merge_peers_range(btlbf_ctx* ctx, int kind, void* const* bases, int world, int rank, uint64_t lo,
                                       uint64_t hi, void* cuda_stream)
{
	if (!bases || world < 1 || world > kMaxPeers || rank < 0 || rank >= world || lo > hi || ((lo | hi) & 15u))
		return fail(BTLBF_ERR_ARG, "bad peer list / byte range");
	if (kind != BTLBF_BLOOM && kind != BTLBF_COUNTING8)
		return fail(BTLBF_ERR_ARG, "bad filter kind");
	TRY(use(ctx));
	LOCKED(ctx);
	PeerMergeParams M;
	memset(&M, 0, sizeof M);
	for (int i = 0; i < world; i++) {
		void* b = bases[(rank + i) % world];
		if (!b || ((uintptr_t)b & 15u))
			return fail(BTLBF_ERR_ARG, "peer base %d is null or not 16-byte aligned", (rank + i) % world);
		M.base[i] = (uint8_t*)b;
	}
	M.world = (uint32_t)world;
	M.sat_add = kind == BTLBF_COUNTING8;
	M.unroll = (uint32_t)ctx->peer_unroll;
	M.grid = (uint32_t)ctx->peer_grid;
	uint64_t a = 0, b = 0;
	TRY(btlbf_merge_slice(hi - lo, world, rank, &a, &b));
	M.lo = lo + a;
	M.hi = lo + (b < hi - lo ? b : hi - lo);
	if (M.lo > M.hi)
		M.lo = M.hi;
	cudaError_t e = launch_peer_merge(M, cuda_stream ? (cudaStream_t)cuda_stream : ctx->active);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "peer merge launch failed: %s", cudaGetErrorString(e));
	ctx->launches++;
	return BTLBF_OK;
}

// ---------------------------------------------------------------- ordered (exact) updates
static int ordered_state(btlbf_filter* f, uint32_t batch)
{
	btlbf_ctx* ctx = f->ctx;
	uint32_t want_log2 = (uint32_t)ctx->resv_log2;
	if (f->d_touched && f->resv_log2 != want_log2) {
		CU(cudaFree(f->d_touched));
		CU(cudaFree(f->d_contended));
		f->d_touched = f->d_contended = nullptr;
	}
	if (!f->d_touched) {
		size_t bytes = (((size_t)1 << want_log2) + 7) / 8;
		bytes = bytes < 4 ? 4 : bytes;
		CU(cudaMalloc(&f->d_touched, bytes));
		CU(cudaMalloc(&f->d_contended, bytes));
		CU(cudaMemsetAsync(f->d_touched, 0, bytes, ctx->active));
		CU(cudaMemsetAsync(f->d_contended, 0, bytes, ctx->active));
		f->resv_log2 = want_log2;
	}
	if (f->pending_cap < batch) {
		for (int i = 0; i < 2; i++) {
			if (f->d_pending[i])
				CU(cudaFree(f->d_pending[i]));
			f->d_pending[i] = nullptr;
			CU(cudaMalloc(&f->d_pending[i], (size_t)batch * 4));
		}
		f->pending_cap = batch;
	}
	// slot cache of the deferred windows (h may change with btlbf_filter_set_seeds)
	if (f->d_pending_slots && (f->pending_slots_h != f->hc.h || f->pending_slots_cap < batch)) {
		CU(cudaFree(f->d_pending_slots));
		f->d_pending_slots = nullptr;
	}
	if (!f->d_pending_slots && (f->pending_slots_h != f->hc.h || f->pending_slots_cap < batch) &&
	    (uint64_t)batch * f->hc.h * 8 <= ((uint64_t)1 << 30)) {
		if (cudaMalloc(&f->d_pending_slots, (size_t)batch * f->hc.h * 8) != cudaSuccess) {
			cudaGetLastError();
			f->d_pending_slots = nullptr; // optional: the rounds then re-derive the slots
		}
		f->pending_slots_h = f->hc.h;
		f->pending_slots_cap = f->d_pending_slots ? batch : 0;
	}
	uint32_t want_list = (uint32_t)ctx->list_log2;
	if (f->d_list_resv && f->list_log2 != want_list) {
		CU(cudaFree(f->d_list_resv));
		f->d_list_resv = nullptr;
	}
	if (!f->d_list_resv) {
		size_t bytes = ((size_t)1 << want_list) * 8;
		CU(cudaMalloc(&f->d_list_resv, bytes));
		CU(cudaMemsetAsync(f->d_list_resv, 0xff, bytes, ctx->active));
		f->list_log2 = want_list;
		f->epoch = 0;
		if (f->d_ord)
			CU(cudaMemsetAsync(f->d_ord, 0, 32, ctx->active));
	}
	if (!f->d_ord) {
		CU(cudaMalloc(&f->d_ord, 32));
		CU(cudaMemsetAsync(f->d_ord, 0, 32, ctx->active));
	}
	return BTLBF_OK;
}

// Applies an order-dependent update (kind 0: incrementMin, 1: insertAndCheck) to the windows of a
// device-resident chunk so that the result equals the reference's single-threaded, read-order,
// position-order loop.  P describes the whole chunk (outputs included).
static int ordered_apply(btlbf_filter* f, const SeqParams& chunk, int kind, cudaStream_t s)
{
	btlbf_ctx* ctx = f->ctx;
	uint64_t batch = (uint64_t)ctx->cbf_batch;
	TRY(ordered_state(f, (uint32_t)batch));
	uint32_t* d_cnt = reinterpret_cast<uint32_t*>(ctx->d_scalars + 4); // [0],[1]: list counters, [2]: rounds
	volatile uint32_t* h_cnt = reinterpret_cast<volatile uint32_t*>(ctx->h_scalars + 4);
	const bool coop = ctx->ordered_coop != 0;
	if (coop)
		d_cnt = f->d_ord;
	for (uint64_t b0 = 0; b0 < chunk.n_windows; b0 += batch) {
		SeqParams P = chunk;
		uint64_t bw = chunk.n_windows - b0 < batch ? chunk.n_windows - b0 : batch;
		advance_input(P, chunk, b0);
		P.n_windows = bw;
		if (chunk.hit_bits) P.hit_bits = chunk.hit_bits + (b0 >> 5);
		if (chunk.valid_bits) P.valid_bits = chunk.valid_bits + (b0 >> 5);
		P.out_words = chunk.out_words > (b0 >> 5) ? chunk.out_words - (b0 >> 5) : 0;
		P.resv_touched = f->d_touched;
		P.resv_contended = f->d_contended;
		P.resv_log2 = f->resv_log2;
		P.pending = f->d_pending[0];
		P.pending_count = d_cnt;
		P.pending_slots = f->d_pending_slots;
		if (!coop)
			CU(cudaMemsetAsync(d_cnt, 0, 16, s));
		// pass 1 carries no outputs; pass 2 writes valid/hit words and the k-mer statistics
		SeqParams T = P;
		T.hit_bits = T.valid_bits = nullptr;
		T.stats = nullptr;
		TRY(launch(ctx, OP_RESV_TOUCH, T, s));
		TRY(launch(ctx, kind == 0 ? OP_CBF_COMMIT : OP_BFCHK_COMMIT, P, s));
		// re-arm the sketch: a sweep over the tables when the batch touched a good part of them, else the
		// clear pass (which re-derives the positions)
		const size_t table_bytes = (((size_t)1 << f->resv_log2) + 7) / 8;
		if ((double)bw * P.h * 64.0 >= (double)table_bytes) {
			CU(cudaMemsetAsync(f->d_touched, 0, table_bytes < 4 ? 4 : table_bytes, s));
			CU(cudaMemsetAsync(f->d_contended, 0, table_bytes < 4 ? 4 : table_bytes, s));
		} else
			TRY(launch(ctx, OP_RESV_CLEAR, T, s));
		if (coop) {
			// the residual rounds run to completion on the device (grid-wide barriers): nothing to wait for
			ListParams L;
			memset(&L, 0, sizeof L);
			L.list_in = f->d_pending[0];
			L.list_out = f->d_pending[1];
			L.counts = d_cnt;
			L.count_in = d_cnt;
			L.count_out = d_cnt + 1;
			L.resv = f->d_list_resv;
			L.resv_log2 = f->list_log2;
			L.kind = (uint32_t)kind;
			L.d_epoch = d_cnt + 2;
			L.rounds_out = d_cnt + 3;
			cudaError_t e = launch_list_drain_coop(P, L, s);
			if (e != cudaSuccess)
				return fail(BTLBF_ERR_CUDA, "cooperative drain launch failed: %s", cudaGetErrorString(e));
			ctx->launches++;
			continue;
		}
		CU(cudaMemcpyAsync((void*)h_cnt, d_cnt, 4, cudaMemcpyDeviceToHost, s));
		CU(cudaStreamSynchronize(s));
		uint32_t n = h_cnt[0];
		f->deferred_total += n;
		int cur = 0;
		while (n > 0) {
			ListParams L;
			memset(&L, 0, sizeof L);
			L.list_in = f->d_pending[cur];
			L.count_in = d_cnt + cur;
			L.list_out = f->d_pending[1 - cur];
			L.count_out = d_cnt + (1 - cur);
			L.resv = f->d_list_resv;
			L.resv_log2 = f->list_log2;
			L.kind = (uint32_t)kind;
			L.max_items = n;
			if (f->epoch > 0xfff00000u) { // re-arm the reservation table long before the epoch wraps
				CU(cudaMemsetAsync(f->d_list_resv, 0xff, ((size_t)1 << f->list_log2) * 8, s));
				f->epoch = 0;
			}
			L.epoch = ++f->epoch;
			if ((int64_t)n > ctx->drain_threshold) {
				CU(cudaMemsetAsync(d_cnt + (1 - cur), 0, 4, s));
				cudaError_t e = launch_list_round(0, P, L, s);
				if (e == cudaSuccess)
					e = launch_list_round(1, P, L, s);
				if (e != cudaSuccess)
					return fail(BTLBF_ERR_CUDA, "list round launch failed: %s", cudaGetErrorString(e));
				ctx->launches += 2;
				f->rounds_total++;
				CU(cudaMemcpyAsync((void*)(h_cnt + (1 - cur)), d_cnt + (1 - cur), 4, cudaMemcpyDeviceToHost, s));
				CU(cudaStreamSynchronize(s));
				n = h_cnt[1 - cur];
				cur = 1 - cur;
			} else {
				L.rounds_out = d_cnt + 2;
				cudaError_t e = launch_list_drain(P, L, s);
				if (e != cudaSuccess)
					return fail(BTLBF_ERR_CUDA, "list drain launch failed: %s", cudaGetErrorString(e));
				ctx->launches++;
				f->epoch += n + 1; // the drain kernel runs at most n rounds, one epoch each
				f->rounds_total += 1;
				n = 0;
			}
		}
	}
	return BTLBF_OK;
}

// ---------------------------------------------------------------- partitioned BloomFilter build
// Random single-bit atomics into a filter much larger than L2 cost one 128-byte HBM fetch plus a
// write-back each and top out near 20 G updates/s on B200; the same atomics into an L2-resident region run
// ~9x faster.  So large batches are built in two passes: (1) hash every k-mer and append the bit offsets to
// per-partition buckets (sequential traffic, 4 bytes per hash), (2) partition by partition, OR the offsets
// into the filter while that 2^bin_part_log2-bit region sits in L2.  The result is bit-identical to the
// direct path because OR is commutative and idempotent.
static bool want_binned(const btlbf_filter* f, const SeqParams& P)
{
	const btlbf_ctx* ctx = f->ctx;
	if (ctx->bin_mode < 0 || f->kind != BTLBF_BLOOM)
		return false;
	if (ctx->bin_mode > 0)
		return true;
	// auto: the filter does not fit in L2 and the batch is large enough to amortise streaming it.  Pass 2 runs
	// once per accumulation (several batches), hence the lower bar than for the partitioned query.
	return f->bytes >= ((uint64_t)96 << 20) && P.n_windows * P.h >= f->bytes / 512;
}

// log2(bits per partition) for filter f hashed as P describes (P.bin_legacy set); false: too many partitions
static bool bin_geometry(const btlbf_filter* f, const SeqParams& P, uint32_t* shift_out)
{
	const btlbf_ctx* ctx = f->ctx;
	auto parts_at = [&](uint32_t sh) { return (f->size + (((uint64_t)1 << sh) - 1)) >> sh; };
	// bin_part_log2 is in bits; a counting filter's partition of the same byte size holds 2^(log2 - 3) counters
	uint32_t shift = (uint32_t)(f->kind == BTLBF_COUNTING8 ? (ctx->bin_part_log2 > 11 ? ctx->bin_part_log2 - 3 : 8) : ctx->bin_part_log2);
	while (parts_at(shift) > 4096 && shift < 31)
		shift++;
	if (parts_at(shift) > 4096)
		return false;
	// the sort-bin kernel wants few partitions (long runs per partition and round); widen them when that
	// makes the filter eligible
	if (!P.bin_legacy) {
		uint32_t sh2 = shift;
		while (parts_at(sh2) > (uint64_t)ctx->bin_max_parts && sh2 < 31)
			sh2++;
		if (sh2 != shift && parts_at(sh2) <= (uint64_t)ctx->bin_max_parts && bin_sort_eligible(P, (uint32_t)parts_at(sh2)))
			shift = sh2;
	}
	*shift_out = shift;
	return true;
}

// partition geometry + sub-bucket storage shared by the partitioned build and query.
// capacity_windows: windows the sub-buckets are sized for (>= P.n_windows).
static int bin_setup(btlbf_filter* f, SeqParams& P, bool query, uint64_t capacity_windows, uint32_t* grid, int* mode,
                     DevBuf& items, DevBuf& counts)
{
	btlbf_ctx* ctx = f->ctx;
	P.bin_legacy = (uint32_t)ctx->bin_kernel;
	P.bin_rot = 0;
	uint32_t shift = 0;
	if (!bin_geometry(f, P, &shift))
		return fail(BTLBF_ERR_ARG, "filter too large for the partitioned path");
	auto parts_at = [&](uint32_t sh) { return (f->size + (((uint64_t)1 << sh) - 1)) >> sh; };
	uint64_t n_bins = parts_at(shift);
	P.n_bins = (uint32_t)n_bins;
	P.bin_shift = shift;
	P.bin_mask = (uint32_t)(((uint64_t)1 << shift) - 1);
	P.bin_counting = f->kind == BTLBF_COUNTING8;
	{
		// Prefetch rules of pass 2, measured on B200 (profiles/r2_prefetch_sweep.txt).  Partitions up to 16 MiB: pull the
		// NEXT partition into L2 while this one is processed (cfg2 build 27.9 against 23.8 Gk-mer/s without).  Larger
		// partitions (the 16 GB filters: 512 x 32 MiB): two of them plus the item stream do not stay resident, so the
		// build prefetches the CTA's share of its OWN partition in full lines (19.1 against 17.1 with the next one and
		// 16.7 with none) and the query none at all (16 GiB BloomFilter 24.7 against 22.4 / 21.3; counting filter 24.5
		// against 23.4 / 20.5).  1024 partitions of 16 MiB instead: slower in every case (pass 1 runs get too short).
		const uint64_t part_bytes = f->kind == BTLBF_COUNTING8 ? (uint64_t)1 << shift : ((uint64_t)1 << shift) >> 3;
		const bool small = part_bytes <= ((uint64_t)16 << 20);
		P.bin_prefetch = ctx->bin_prefetch < 0 ? (small ? 1u : query ? 0u : 2u) : (uint32_t)ctx->bin_prefetch;
	}
	uint32_t writers = 0;
	cudaError_t e = bin_plan(P, P.n_bins, query, &writers, grid, mode);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "planning the partitioned pass failed: %s", cudaGetErrorString(e));
	P.bin_writers = writers;
	P.bin_segs = writers >= 2048 ? 1u : (4096u + writers - 1) / writers;
	double parts = (double)f->size / (double)((uint64_t)1 << shift); // fractional: the last one is partial
	double expect = (double)capacity_windows * P.h / parts / writers;
	uint64_t cap = (uint64_t)(expect * (1.0 + ctx->bin_slack_pct / 100.0)) + 96;
	cap = (cap + 7) / 8 * 8; // whole 32-byte lines
	if (cap > 0x7ffffff8ULL)
		cap = 0x7ffffff8ULL;
	P.bin_cap = (uint32_t)cap;
	if (n_bins * writers * cap * (query ? 8 : 4) > items.cap || n_bins * writers * 4 > counts.cap) {
		// growing a buffer frees the old one: nothing in flight may still use it
		CU(cudaStreamSynchronize(ctx->aux));
		CU(cudaStreamSynchronize(ctx->active));
	}
	TRY(ensure(items, n_bins * writers * cap * (query ? 8 : 4)));
	TRY(ensure(counts, n_bins * writers * 4));
	P.bin_items = (uint32_t*)items.p;
	P.bin_counts = (uint32_t*)counts.p;
	return BTLBF_OK;
}

// pass 2 of the partitioned build for the sub-buckets described by P: on the background stream when
// overlapping (it then runs under the next batches' pass 1), else on s
static int apply_pass(btlbf_ctx* ctx, const SeqParams& P, int slot, cudaStream_t s, uint64_t windows)
{
	cudaStream_t s2 = ctx->overlap ? ctx->aux : s;
	if (ctx->overlap) {
		CU(cudaEventRecord(ctx->ev_bin_done[slot], s));
		CU(cudaStreamWaitEvent(s2, ctx->ev_bin_done[slot], 0));
	}
	// Two-level pass 2 (apply2.cu) when there are enough items to pay for splitting them a second time.  Not in
	// overlap mode: its plain stores of whole slices must not race with the direct atomics of a concurrent pass 1.
	uint32_t sub_shift = 0, n_sub = 0;
	const uint64_t items = windows * P.h;
	bool two = ctx->bin_two_level && !ctx->overlap && P.bin_segs > 1 && items >= (uint64_t)ctx->bin_two_level_min &&
	           apply2_geometry(P.bin_shift, &sub_shift, &n_sub);
	cudaError_t e = cudaSuccess;
	if (two) {
		Apply2Params A;
		memset(&A, 0, sizeof A);
		A.items = P.bin_items; A.counts = P.bin_counts;
		A.n_bins = P.n_bins; A.writers = P.bin_writers; A.cap = P.bin_cap; A.bin_shift = P.bin_shift;
		A.sub_shift = sub_shift; A.n_sub = n_sub;
		A.writers2 = P.bin_writers < 32 ? P.bin_writers : 32;
		double parts = (double)P.fm.m / (double)((uint64_t)1 << P.bin_shift);
		double expect = (double)items / parts / n_sub / A.writers2;
		uint64_t cap2 = (uint64_t)(expect * 1.3) + 32;
		cap2 = (cap2 + 3) / 4 * 4;
		const uint64_t buckets = (uint64_t)P.n_bins * n_sub * A.writers2;
		if (cap2 > 0x7ffffff0ULL || buckets * cap2 * 4 > ((uint64_t)64 << 30))
			two = false;
		else {
			A.cap2 = (uint32_t)cap2;
			if (buckets * cap2 * 4 > ctx->ibin2_items.cap || buckets * 4 > ctx->ibin2_counts.cap)
				CU(cudaStreamSynchronize(s2)); // growing a buffer frees the old one
			int rc = ensure(ctx->ibin2_items, buckets * cap2 * 4);
			if (rc == BTLBF_OK)
				rc = ensure(ctx->ibin2_counts, buckets * 4);
			if (rc != BTLBF_OK) {
				cudaGetLastError();
				two = false; // not enough memory for the second level: the one-level pass needs none
			} else {
				A.items2 = (uint32_t*)ctx->ibin2_items.p;
				A.counts2 = (uint32_t*)ctx->ibin2_counts.p;
				A.filter = (uint32_t*)P.filter;
				A.m = P.fm.m;
				A.alloc_words = ((P.fm.m + 127) / 128) * 4;
				e = launch_apply2(A, s2);
				ctx->launches++; // two kernels
				ctx->two_level_passes++;
			}
		}
	}
	if (!two)
		e = launch_apply_bins(P, s2);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "partitioned build (pass 2) launch failed: %s", cudaGetErrorString(e));
	CU(cudaEventRecord(ctx->ev_apply_done[slot], s2));
	ctx->slot_used[slot] = true;
	if (ctx->overlap) {
		ctx->ev_aux_last = ctx->ev_apply_done[slot];
		ctx->aux_pending = true;
	}
	ctx->launches++;
	return BTLBF_OK;
}

// Runs pass 2 for the k-mers parked in the accumulation, if any.  Everything that reads or writes filter
// contents goes through joined(), which calls this first.
static int settle(btlbf_ctx* ctx)
{
	if (!ctx->acc.f)
		return BTLBF_OK;
	ctx->acc.f = nullptr;
	int rc = apply_pass(ctx, ctx->acc.P, ctx->acc.slot, ctx->active, ctx->acc.windows);
	if (rc != BTLBF_OK)
		ctx->settle_error = rc;
	return rc;
}

static int binned_insert(btlbf_filter* f, SeqParams P, cudaStream_t s)
{
	btlbf_ctx* ctx = f->ctx;
	uint32_t grid = 0;
	int mode = 0;
	// accumulation target in windows: the storage budget, but no more than makes pass 2's one sweep over
	// the filter a small part of the item traffic, and never less than this batch
	uint64_t budget = (uint64_t)ctx->bin_accum_bytes;
	if (budget > 8 * f->bytes)
		budget = 8 * f->bytes;
	uint64_t target = (uint64_t)((double)budget / (4.0 * P.h * (1.0 + ctx->bin_slack_pct / 100.0)));
	if (target < P.n_windows)
		target = P.n_windows;

	if (ctx->acc.f) {
		// append to the running accumulation when it is this filter's and there is room
		SeqParams Q = P;
		const SeqParams& A = ctx->acc.P;
		Q.bin_legacy = (uint32_t)ctx->bin_kernel;
		uint32_t w = 0, g = 0;
		int m = -1;
		bool same = ctx->acc.f == f && A.filter == P.filter && A.fm.m == P.fm.m && !Q.bin_legacy &&
		            ctx->acc.windows + P.n_windows <= ctx->acc.capacity &&
		            bin_plan(Q, A.n_bins, false, &w, &g, &m) == cudaSuccess && m == BIN_SORT && w == A.bin_writers;
		if (same) {
			P.n_bins = A.n_bins; P.bin_shift = A.bin_shift; P.bin_mask = A.bin_mask; P.bin_cap = A.bin_cap;
			P.bin_writers = A.bin_writers; P.bin_segs = A.bin_segs; P.bin_items = A.bin_items; P.bin_counts = A.bin_counts;
			P.bin_legacy = 0;
			P.bin_rot = (uint32_t)(ctx->acc.tiles % A.bin_writers);
			cudaError_t e = launch_bin(P, false, A.bin_writers, s);
			if (e != cudaSuccess)
				return fail(BTLBF_ERR_CUDA, "partitioned build (pass 1) launch failed: %s", cudaGetErrorString(e));
			ctx->acc.windows += P.n_windows;
			ctx->acc.tiles += (P.n_windows + bin_sort_tile(P, P.n_bins) - 1) / bin_sort_tile(P, P.n_bins);
			ctx->launches++;
			ctx->binned_launches++;
			return BTLBF_OK;
		}
		TRY(settle(ctx));
	}

	const int slot = ctx->bin_slot;
	if (ctx->overlap)
		ctx->bin_slot ^= 1;
	TRY(bin_setup(f, P, false, target, &grid, &mode, ctx->ibin_items[slot], ctx->ibin_counts[slot]));
	// pass 1 (hash + bin) on the active stream; it may not overwrite sub-buckets pass 2 is still reading
	if (ctx->slot_used[slot])
		CU(cudaStreamWaitEvent(s, ctx->ev_apply_done[slot], 0));
	if (mode == BIN_SORT) {
		CU(cudaMemsetAsync(P.bin_counts, 0, (size_t)P.n_bins * P.bin_writers * 4, s));
		grid = P.bin_writers; // every writer runs (idle ones just keep their cursors)
	}
	cudaError_t e = launch_bin(P, false, grid, s);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "partitioned build (pass 1) launch failed: %s", cudaGetErrorString(e));
	ctx->launches++;
	ctx->binned_launches++;
	if (mode == BIN_SORT) {
		ctx->acc.f = f;
		ctx->acc.P = P;
		ctx->acc.windows = P.n_windows;
		ctx->acc.tiles = (P.n_windows + bin_sort_tile(P, P.n_bins) - 1) / bin_sort_tile(P, P.n_bins);
		ctx->acc.capacity = target;
		ctx->acc.slot = slot;
		if (ctx->acc.windows >= ctx->acc.capacity)
			TRY(settle(ctx));
		return BTLBF_OK;
	}
	return apply_pass(ctx, P, slot, s, P.n_windows);
}

// Partitioned query: bin (offset, window) pairs by filter partition, test them while the partition is
// resident in L2, clear the hit bit of every window with a missing bit (BloomFilter) or a counter below the
// threshold (CountingBloomFilter::contains).  Same booleans as the direct path.
static bool want_binned_query(const btlbf_filter* f, const SeqParams& P)
{
	const btlbf_ctx* ctx = f->ctx;
	if (ctx->bin_query_mode < 0 || P.n_windows > 0xffffffffULL)
		return false;
	if (f->kind == BTLBF_COUNTING8 && (P.counts || P.threshold == 0 || P.threshold > 255))
		return false; // per-window minima need the gather kernel; threshold 0 / > 255 are constant answers there
	if (!P.hit_bits && !P.stats)
		return false;
	if (((uintptr_t)P.hit_bits | (uintptr_t)P.valid_bits) & 3u)
		return false;
	SeqParams Q = P;
	Q.bin_legacy = (uint32_t)ctx->bin_kernel;
	uint32_t shift = 0;
	if (!bin_geometry(f, Q, &shift) || !bin_query_supported(Q, (uint32_t)((f->size + (((uint64_t)1 << shift) - 1)) >> shift)))
		return false;
	if (ctx->bin_query_mode > 0)
		return true;
	// auto (widened partitions included: without the prefetch of the next partition a 16 GiB filter answers 24.2
	// Gk-mer/s partitioned against 22.3 direct, the 16e9-counter filter 24.7 against 11.3)
	return f->bytes >= ((uint64_t)96 << 20) && P.n_windows * P.h >= f->bytes / 64;
}

static int binned_query(btlbf_filter* f, SeqParams P, cudaStream_t s)
{
	btlbf_ctx* ctx = f->ctx;
	const bool counting = f->kind == BTLBF_COUNTING8;
	{
		cudaStream_t js; // parked k-mers first: the probes (and pass 1's overflow path) read the filter
		TRY(join(ctx, &js));
	}
	// ---- how many pass-1 CTAs fit on an SM for this shape (0: not the sort-bin kernel)
	int64_t fit = 0;
	{
		SeqParams T = P;
		T.bin_legacy = (uint32_t)ctx->bin_kernel;
		T.bin_ctas_per_sm = 0;
		T.n_windows = ~0ull >> 8;
		uint32_t w = 0, g = 0, shift = 0;
		int m = 0;
		if (bin_geometry(f, T, &shift)) {
			const uint32_t nb = (uint32_t)((f->size + (((uint64_t)1 << shift) - 1)) >> shift);
			if (bin_plan(T, nb, true, &w, &g, &m) == cudaSuccess && m == BIN_SORT) {
				int dev = 0, sms = 1;
				cudaGetDevice(&dev);
				cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
				fit = w / (uint32_t)sms;
			}
		}
	}
	// ---- sub-batches: windows [c0, c0 + sub) of the chunk, whole 16384-window tiles (both CTA shapes, whole
	// result words).  Overlapping needs the sort-bin kernel in the shape that leaves an SM room for the probe kernel.
	uint64_t n_sub = (uint64_t)ctx->query_sub;
	if (fit == 0 || (n_sub == 0 && fit < 3))
		n_sub = 1;
	if (n_sub == 0)
		n_sub = 1; // measured on B200 (cfg2): 15.0 ms in one pass, 16.6 / 18.0 / 19.4 ms in 2 / 4 / 8 overlapped sub-batches
		           // (same L1 configuration) -- the pass-1 CTAs' shared memory leaves pass 2 too small an L1
	const uint64_t kAlign = 16384;
	uint64_t sub = ((P.n_windows + n_sub - 1) / n_sub + kAlign - 1) / kAlign * kAlign;
	n_sub = (P.n_windows + sub - 1) / sub;
	const bool overlap = n_sub > 1;

	SeqParams G = P; // geometry for one sub-batch
	G.n_windows = sub < P.n_windows ? sub : P.n_windows;
	uint32_t grid = 0;
	int mode = 0;
	if (overlap) {
		// pass 1 leaves room on every SM for the probe kernel of the previous sub-batch
		int64_t ctas = ctx->query_p1_ctas;
		if (ctas == 0)
			ctas = fit > 1 ? fit - 1 : 1;
		G.bin_ctas_per_sm = (uint32_t)ctas;
	}
	TRY(bin_setup(f, G, true, G.n_windows, &grid, &mode, ctx->qbin_items[0], ctx->qbin_counts[0]));
	const size_t item_bytes = (size_t)G.n_bins * G.bin_writers * G.bin_cap * 8, count_bytes = (size_t)G.n_bins * G.bin_writers * 4;
	if (overlap) {
		if (item_bytes > ctx->qbin_items[1].cap || count_bytes > ctx->qbin_counts[1].cap) {
			CU(cudaStreamSynchronize(ctx->aux));
			CU(cudaStreamSynchronize(ctx->active));
		}
		TRY(ensure(ctx->qbin_items[1], item_bytes));
		TRY(ensure(ctx->qbin_counts[1], count_bytes));
	}
	const uint64_t words = P.out_words;
	if (!P.hit_bits) {
		TRY(ensure(ctx->q_hit, words * 4));
		P.hit_bits = (uint32_t*)ctx->q_hit.p;
	}
	if (!P.valid_bits) {
		TRY(ensure(ctx->q_valid, words * 4));
		P.valid_bits = (uint32_t*)ctx->q_valid.p;
	}
	uint64_t* stats = P.stats;
	cudaError_t e = cudaSuccess;
	// Adaptive path selection, entirely on the device (no host round trip, so calls still queue back to back):
	// the early-exit kernel runs over a strided sample of the tiles; a one-thread kernel turns the sample's hit
	// fraction into a flag; then BOTH paths are launched and the kernels of the one that lost return at once.
	// Absent k-mers cost the early-exit kernel ~1/(1-occupancy) probes but the partitioned path all h of them
	// (plus an atomic per missing bit), so read sets that mostly miss are faster on the direct kernel.
	const uint64_t tiles = (P.n_windows + kTile - 1) / kTile;
	const bool adaptive = ctx->query_adaptive && !counting && mode == BIN_SORT &&
	                      tiles >= (uint64_t)ctx->query_adaptive_min_tiles && P.n_seeds == 0;
	CU(cudaMemsetAsync(P.hit_bits, 0xff, words * 4, s)); // the partitioned path clears the bits of missing k-mers
	const uint32_t* gate = nullptr;
	if (adaptive) {
		unsigned long long* sample = ctx->d_scalars + 8; // [8] valid, [9] hits, [10] flag
		uint32_t* flag = reinterpret_cast<uint32_t*>(ctx->d_scalars + 10);
		CU(cudaMemsetAsync(sample, 0, 24, s));
		SeqParams S = P;
		S.hit_bits = S.valid_bits = nullptr;
		S.stats = reinterpret_cast<uint64_t*>(sample);
		S.tile_count = (uint32_t)(tiles < 64 ? tiles : 64);
		S.tile_stride = (uint32_t)(tiles / S.tile_count);
		S.tile_first = S.tile_stride / 2;
		S.query_mode = 0; // all h probes of a window in flight together: the sample's latency is what matters
		TRY(launch(ctx, OP_BF_CONTAINS, S, s));
		e = launch_query_gate(sample, flag, (uint32_t)ctx->query_adaptive_pct, s);
		if (e != cudaSuccess)
			return fail(BTLBF_ERR_CUDA, "query gate launch failed: %s", cudaGetErrorString(e));
		SeqParams D = P;
		D.gate = flag;
		D.gate_want = 1;
		D.query_mode = 1;
		D.tiles_per_cta = 8; // fewer CTAs to retire when the partitioned path is the one that runs
		TRY(launch(ctx, OP_BF_CONTAINS, D, s));
		ctx->launches++;
		gate = flag;
	}
	int unroll = (int)ctx->query_probe_unroll;
	if (unroll == 0)
		unroll = overlap ? 4 : 1;
	for (uint64_t i = 0, c0 = 0; c0 < P.n_windows; i++, c0 += sub) {
		const int b = overlap ? (int)(i & 1) : 0;
		SeqParams Q = G;
		advance_input(Q, P, c0);
		Q.n_windows = P.n_windows - c0 < sub ? P.n_windows - c0 : sub;
		if (Q.n_bases > Q.n_windows + P.k - 1)
			Q.n_bases = Q.n_windows + P.k - 1; // a sub-batch reads its windows + the k-1 halo, like a pipeline chunk
		Q.hit_bits = P.hit_bits + (c0 >> 5);
		Q.valid_bits = P.valid_bits + (c0 >> 5);
		Q.out_words = words - (c0 >> 5);
		Q.stats = stats;
		Q.gate = gate;
		Q.gate_want = 0;
		Q.probe_ld = (uint32_t)ctx->probe_ld;
		Q.bin_items = (uint32_t*)ctx->qbin_items[b].p;
		Q.bin_counts = (uint32_t*)ctx->qbin_counts[b].p;
		// pass 1 on the active stream; the buffers are free once the probe kernel that last read them is done
		if (overlap && ctx->qslot_used[b])
			CU(cudaStreamWaitEvent(s, ctx->ev_qp2[b], 0));
		if (mode == BIN_SORT)
			CU(cudaMemsetAsync(Q.bin_counts, 0, count_bytes, s));
		uint32_t g1 = grid;
		if (mode != BIN_SORT) { // the general-shape kernels size their grid by the batch
			uint32_t w = 0;
			int m = 0;
			if (bin_plan(Q, Q.n_bins, true, &w, &g1, &m) != cudaSuccess || w > Q.bin_writers)
				g1 = grid;
		}
		e = launch_bin(Q, true, g1, s);
		if (e != cudaSuccess)
			return fail(BTLBF_ERR_CUDA, "partitioned query (pass 1) launch failed: %s", cudaGetErrorString(e));
		// pass 2: on the background stream when overlapping
		cudaStream_t s2 = s;
		if (overlap) {
			s2 = ctx->aux;
			CU(cudaEventRecord(ctx->ev_qp1[b], s));
			CU(cudaStreamWaitEvent(s2, ctx->ev_qp1[b], 0));
		}
		// the item words of a sub-batch carry window indices relative to the sub-batch
		e = launch_probe_bins(Q, counting, unroll, ctx->probe_carveout < 0 ? overlap : ctx->probe_carveout != 0, s2);
		if (e != cudaSuccess)
			return fail(BTLBF_ERR_CUDA, "partitioned query (pass 2) launch failed: %s", cudaGetErrorString(e));
		if (overlap) {
			CU(cudaEventRecord(ctx->ev_qp2[b], s2));
			ctx->qslot_used[b] = true;
		}
		ctx->launches += 2;
	}
	if (overlap) { // the results (and everything queued after this call) follow the last probes
		CU(cudaStreamWaitEvent(s, ctx->ev_qp2[0], 0));
		if (n_sub > 1)
			CU(cudaStreamWaitEvent(s, ctx->ev_qp2[1], 0));
	}
	e = launch_finalize_hits(P.hit_bits, P.valid_bits, words, stats ? (unsigned long long*)(stats + 1) : nullptr, gate, 0, s);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "partitioned query launch failed: %s", cudaGetErrorString(e));
	ctx->launches++;
	ctx->binned_launches++;
	return BTLBF_OK;
}

// ---------------------------------------------------------------- operations on a device-resident chunk
enum PublicOp { PUB_INSERT, PUB_CONTAINS, PUB_INSERT_CHECK, PUB_MINCOUNT, PUB_INCALL, PUB_HASH };

struct ChunkIO
{
	const uint8_t* d_bases = nullptr;
	const uint8_t* d_invalid = nullptr; // packed input: the invalid plane (may be null)
	bool packed = false;    // d_bases holds 2-bit codes
	uint64_t n_bases = 0;   // bases readable at d_bases
	uint64_t base0 = 0;     // flat position of d_bases[0]
	uint64_t n_windows = 0; // windows of this chunk
	const uint64_t* d_offsets = nullptr;
	uint64_t n_seqs = 0;
	uint32_t* d_hit = nullptr;
	uint32_t* d_valid = nullptr;
	uint8_t* d_counts = nullptr;
	uint64_t* d_hashes = nullptr;
	uint8_t* d_strands = nullptr;
	uint64_t* d_stats = nullptr;
};

static void fill_io(SeqParams& P, const ChunkIO& io)
{
	P.bases = io.d_bases;
	P.invalid = io.packed ? io.d_invalid : nullptr;
	P.packed = io.packed;
	if (io.packed)
		P.force_generic = 0; // the byte-class path needs the ASCII tile
	P.n_bases = io.n_bases;
	P.base0 = io.base0;
	P.n_windows = io.n_windows;
	P.offsets = io.d_offsets;
	P.n_seqs = io.n_seqs;
	P.hit_bits = io.d_hit;
	P.valid_bits = io.d_valid;
	P.out_words = (io.n_windows + 31) / 32;
	P.counts = io.d_counts;
	P.hashes = io.d_hashes;
	P.strands = io.d_strands;
	P.stats = io.d_stats;
}

static int filter_op_dev(btlbf_filter* f, PublicOp op, const ChunkIO& io, cudaStream_t s)
{
	SeqParams P = filter_params(f);
	fill_io(P, io);
	btlbf_ctx* ctx = f->ctx;
	const bool binned = P.n_windows && ((op == PUB_INSERT && f->kind == BTLBF_BLOOM && want_binned(f, P)) ||
	                                    (op == PUB_CONTAINS && want_binned_query(f, P)));
	if (!binned) { // the direct kernels read / write the filter right away
		cudaStream_t js;
		TRY(join(ctx, &js));
	}
	switch (op) {
	case PUB_INSERT:
		if (f->kind == BTLBF_BLOOM) {
			if (P.n_windows && want_binned(f, P))
				return binned_insert(f, P, s);
			return launch(ctx, OP_BF_INSERT, P, s);
		}
		return ordered_apply(f, P, 0, s);
	case PUB_CONTAINS:
		if (P.n_windows && want_binned_query(f, P))
			return binned_query(f, P, s);
		return launch(ctx, f->kind == BTLBF_BLOOM ? OP_BF_CONTAINS : OP_CBF_MINCOUNT, P, s);
	case PUB_INSERT_CHECK:
		return ordered_apply(f, P, f->kind == BTLBF_BLOOM ? 1 : 0, s);
	case PUB_MINCOUNT:
		if (f->kind != BTLBF_COUNTING8)
			return fail(BTLBF_ERR_STATE, "mincount_seqs needs a counting filter");
		return launch(ctx, OP_CBF_MINCOUNT, P, s);
	case PUB_INCALL:
		if (f->kind != BTLBF_COUNTING8)
			return fail(BTLBF_ERR_STATE, "increment_all_seqs needs a counting filter");
		return launch(ctx, OP_CBF_INCALL, P, s);
	default:
		return fail(BTLBF_ERR_ARG, "bad operation");
	}
}

// Caller-owned filter memory (btlbf_filter_wrap): the caller may read it on the active stream without going
// through this library, so no k-mer stays parked in the partition buckets when a call returns.
static int settle_if_wrapped(btlbf_filter* f)
{
	if (f->wrapped && !f->ctx->wrap_accumulate && f->ctx->acc.f == f)
		return settle(f->ctx);
	return BTLBF_OK;
}

static int check_dev_args(btlbf_filter* f, const void* d_bases, uint64_t n_bases, const uint64_t* d_offsets)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	if (n_bases && (!d_bases || !d_offsets))
		return fail(BTLBF_ERR_ARG, "null device batch");
	if ((uintptr_t)d_bases & 15u)
		return fail(BTLBF_ERR_ARG, "device bases must be 16-byte aligned");
	return use(f->ctx);
}

extern "C" int btlbf_insert_seqs_dev(btlbf_filter* f, const void* d_bases, uint64_t n_bases, const uint64_t* d_offsets,
                                     uint64_t n_seqs, uint64_t* d_stats)
{
	TRY(check_dev_args(f, d_bases, n_bases, d_offsets));
	ChunkIO io;
	io.d_bases = (const uint8_t*)d_bases;
	io.n_bases = io.n_windows = n_bases;
	io.d_offsets = d_offsets;
	io.n_seqs = n_seqs;
	io.d_stats = d_stats;
	LOCKED(f->ctx);
	TRY(filter_op_dev(f, PUB_INSERT, io, f->ctx->active));
	return settle_if_wrapped(f);
}

extern "C" int btlbf_contains_seqs_dev(btlbf_filter* f, const void* d_bases, uint64_t n_bases,
                                       const uint64_t* d_offsets, uint64_t n_seqs, uint32_t* d_hit_bits,
                                       uint32_t* d_valid_bits, uint64_t* d_stats)
{
	TRY(check_dev_args(f, d_bases, n_bases, d_offsets));
	ChunkIO io;
	io.d_bases = (const uint8_t*)d_bases;
	io.n_bases = io.n_windows = n_bases;
	io.d_offsets = d_offsets;
	io.n_seqs = n_seqs;
	io.d_hit = d_hit_bits;
	io.d_valid = d_valid_bits;
	io.d_stats = d_stats;
	LOCKED(f->ctx);
	return filter_op_dev(f, PUB_CONTAINS, io, f->ctx->active);
}

extern "C" int btlbf_mincount_seqs_dev(btlbf_filter* f, const void* d_bases, uint64_t n_bases,
                                       const uint64_t* d_offsets, uint64_t n_seqs, uint8_t* d_counts,
                                       uint32_t* d_valid_bits, uint64_t* d_stats)
{
	TRY(check_dev_args(f, d_bases, n_bases, d_offsets));
	ChunkIO io;
	io.d_bases = (const uint8_t*)d_bases;
	io.n_bases = io.n_windows = n_bases;
	io.d_offsets = d_offsets;
	io.n_seqs = n_seqs;
	io.d_counts = d_counts;
	io.d_valid = d_valid_bits;
	io.d_stats = d_stats;
	LOCKED(f->ctx);
	if (d_counts)
		CU(cudaMemsetAsync(d_counts, 0, n_bases, f->ctx->active));
	return filter_op_dev(f, PUB_MINCOUNT, io, f->ctx->active);
}

// ---------------------------------------------------------------- host-buffer pipeline
struct HostIO
{
	const char* bases = nullptr;
	const uint8_t* invalid = nullptr; // packed input: one bit per base (may be null)
	bool packed = false;              // bases holds 2-bit codes, 4 per byte
	const uint64_t* offsets = nullptr;
	uint64_t n_seqs = 0;
	uint8_t* hit_bits = nullptr;   // ceil(n/32)*4 bytes each
	uint8_t* valid_bits = nullptr;
	uint8_t* counts = nullptr;     // n bytes
	uint64_t* hashes = nullptr;    // n*H
	uint8_t* strands = nullptr;    // n*H
	uint64_t* n_kmers = nullptr;
	uint64_t* n_hits = nullptr;
};

// grows a pipeline buffer; since that frees the old allocation, nothing may be in flight when it happens
static int ensure_idle(btlbf_ctx* ctx, DevBuf& b, size_t bytes)
{
	if (bytes <= b.cap)
		return BTLBF_OK;
	CU(cudaStreamSynchronize(ctx->copy_in));
	CU(cudaStreamSynchronize(ctx->active));
	CU(cudaStreamSynchronize(ctx->copy_out));
	return ensure(b, bytes);
}

// Streams the flat batch through the GPU in chunks of ctx->chunk_bases windows: chunk i+1 is copied
// in (copy_in stream) and chunk i-1's results are copied out (copy_out stream) while chunk i runs.
// async: return once everything is queued; counts_out (host, 2 words) receives {n_kmers, n_hits} when the
// call completes, and up to kTickets calls may be in flight (H2D of one call overlaps kernels of another).
static int host_pipeline(btlbf_ctx* ctx, btlbf_filter* f, const HashCfg* hash_only, PublicOp op, const HostIO& h,
                         bool async = false, uint64_t* counts_out = nullptr)
{
	TRY(use(ctx));
	LOCKED(ctx);
	if (h.n_seqs && !h.offsets)
		return fail(BTLBF_ERR_ARG, "null offsets");
	if (h.n_kmers) *h.n_kmers = 0;
	if (h.n_hits) *h.n_hits = 0;
	if (counts_out) counts_out[0] = counts_out[1] = 0;
	uint64_t n_bases = h.n_seqs ? h.offsets[h.n_seqs] : 0;
	if (h.n_seqs && h.offsets[0] != 0)
		return fail(BTLBF_ERR_ARG, "offsets[0] must be 0");
	if (n_bases && !h.bases)
		return fail(BTLBF_ERR_ARG, "null bases");
	if (n_bases == 0)
		return BTLBF_OK;
	for (uint64_t i = 0; i < h.n_seqs; i++)
		if (h.offsets[i] > h.offsets[i + 1])
			return fail(BTLBF_ERR_ARG, "offsets must be non-decreasing (sequence %llu)", (unsigned long long)i);
	const HashCfg& hc = hash_only ? *hash_only : f->hc;
	const uint32_t k = hc.k, H = hc.h;
	cudaStream_t s = ctx->active;

	uint64_t chunk = (uint64_t)ctx->chunk_bases;
	// the partitioned query streams the whole filter once per chunk: larger chunks for it
	if (op == PUB_CONTAINS && f && f->kind == BTLBF_BLOOM && f->bytes >= ((uint64_t)96 << 20))
		chunk *= (uint64_t)(async || ctx->query_chunk_factor < 2 ? ctx->query_chunk_factor : ctx->query_chunk_factor / 2);
	// (a blocking call has nothing else in flight: two chunks at least, so that its own copies and kernels overlap)
	// ... while pass 1 of the partitioned build appends to the same sub-buckets whatever the chunk size: small
	// chunks let the H2D copy of one overlap the kernel of the previous one inside a single blocking call
	if (op == PUB_INSERT && f && f->kind == BTLBF_BLOOM && f->bytes >= ((uint64_t)96 << 20) && chunk > ((uint64_t)16 << 20))
		chunk = (chunk / 4 + kTile - 1) / kTile * kTile;
	if (op == PUB_HASH) { // keep the per-chunk hash buffer around 256 MiB
		uint64_t lim = ((uint64_t)256 << 20) / ((uint64_t)H * 9);
		lim = lim / kTile * kTile;
		if (lim < (uint64_t)kTile) lim = kTile;
		if (chunk > lim) chunk = lim;
	}
	{
		// near-equal chunks (a batch slightly larger than chunk_bases is not split into one big and one tiny
		// chunk: every chunk of a large filter costs a pass over the whole filter)
		uint64_t n_chunks = (n_bases + chunk + chunk / 4 - 1) / (chunk + chunk / 4);
		if (n_chunks < 1) n_chunks = 1;
		uint64_t per = (n_bases + n_chunks - 1) / n_chunks;
		chunk = (per + kTile - 1) / kTile * kTile;
	}

	Ticket& tk = ctx->ticket[ctx->ticket_seq++ % kTickets];
	if (tk.in_use) {
		CU(cudaEventSynchronize(tk.done));
		tk.in_use = false;
	}
	const bool want_hit = h.hit_bits != nullptr, want_valid = h.valid_bits != nullptr;
	const bool host_pack = ctx->host_pack && !h.packed && f && (op == PUB_INSERT || op == PUB_CONTAINS);
	const bool need_hit_dev = want_hit || op == PUB_INSERT_CHECK; // list rounds OR into hit words
	auto run_chunks = [&]() -> int {
		TRY(ensure_idle(ctx, tk.offsets, (h.n_seqs + 1) * 8));
		CU(cudaMemcpyAsync(tk.offsets.p, h.offsets, (h.n_seqs + 1) * 8, cudaMemcpyHostToDevice, ctx->copy_in));
		CU(cudaMemsetAsync(tk.d_stats, 0, 16, s));
		for (uint64_t c0 = 0; c0 < n_bases; c0 += chunk) {
			Slot& sl = ctx->slot[ctx->slot_seq++ & 1];
			uint64_t cw = n_bases - c0 < chunk ? n_bases - c0 : chunk;
			uint64_t cb = n_bases - c0 < cw + k - 1 ? n_bases - c0 : cw + k - 1; // chunk + halo
			uint64_t words = (cw + 31) / 32;
			TRY(ensure_idle(ctx, sl.bases, (size_t)chunk + k + 64));
			if (need_hit_dev) TRY(ensure_idle(ctx, sl.hit, (size_t)(chunk / 32 + 1) * 4));
			if (want_valid) TRY(ensure_idle(ctx, sl.valid, (size_t)(chunk / 32 + 1) * 4));
			if (h.counts) TRY(ensure_idle(ctx, sl.counts, (size_t)chunk));
			if (h.hashes) TRY(ensure_idle(ctx, sl.hashes, (size_t)chunk * H * 8));
			if (h.strands) TRY(ensure_idle(ctx, sl.strands, (size_t)chunk * H));
			// the slot is free once its previous results have left the device
			if (sl.used)
				CU(cudaStreamWaitEvent(ctx->copy_in, sl.ev_d2h, 0));
			// the chunk as the device will see it: 2-bit planes (the caller's, or packed here on host threads) or ASCII
			bool c_packed = h.packed;
			const uint8_t* c_codes = h.packed ? (const uint8_t*)h.bases + (c0 >> 2) : nullptr; // c0 is a multiple of the tile
			const uint8_t* c_inv = h.packed && h.invalid ? h.invalid + (c0 >> 3) : nullptr;    // size: whole bytes of both planes
			if (host_pack) {
				if (sl.pk_cap < cb) {
					if (sl.used)
						CU(cudaEventSynchronize(sl.ev_h2d));
					if (sl.pk_codes) cudaFreeHost(sl.pk_codes);
					if (sl.pk_inv) cudaFreeHost(sl.pk_inv);
					sl.pk_codes = sl.pk_inv = nullptr;
					sl.pk_cap = 0;
					const size_t want = (size_t)chunk + k + 64;
					CU(cudaHostAlloc((void**)&sl.pk_codes, want / 4 + 64, cudaHostAllocDefault));
					CU(cudaHostAlloc((void**)&sl.pk_inv, want / 8 + 64, cudaHostAllocDefault));
					sl.pk_cap = want;
				} else if (sl.used) {
					CU(cudaEventSynchronize(sl.ev_h2d)); // the copy that last read this staging buffer has finished
				}
				uint64_t n_bad = 0, raw = ~0ull;
				btl_pack_bases(h.bases + c0, cb, sl.pk_codes, sl.pk_inv, (int)ctx->host_pack_threads, &n_bad, &raw);
				if (raw == ~0ull) { // (a raw byte 1 3 4 5 7 has no packed form: such a chunk travels as ASCII)
					c_packed = true;
					c_codes = sl.pk_codes;
					c_inv = n_bad ? sl.pk_inv : nullptr;
				}
			}
			if (c_packed) {
				CU(cudaMemcpyAsync(sl.bases.p, c_codes, (cb + 3) >> 2, cudaMemcpyHostToDevice, ctx->copy_in));
				if (c_inv) {
					TRY(ensure_idle(ctx, sl.invalid, (size_t)chunk / 8 + k + 64));
					CU(cudaMemcpyAsync(sl.invalid.p, c_inv, (cb + 7) >> 3, cudaMemcpyHostToDevice, ctx->copy_in));
				}
			} else
				CU(cudaMemcpyAsync(sl.bases.p, h.bases + c0, cb, cudaMemcpyHostToDevice, ctx->copy_in));
			CU(cudaEventRecord(sl.ev_h2d, ctx->copy_in));
			CU(cudaStreamWaitEvent(s, sl.ev_h2d, 0)); // also orders the offsets upload before the kernels
			if (sl.used)
				CU(cudaStreamWaitEvent(s, sl.ev_d2h, 0));
			ChunkIO io;
			io.d_bases = (const uint8_t*)sl.bases.p;
			io.packed = c_packed;
			io.d_invalid = c_packed && c_inv ? (const uint8_t*)sl.invalid.p : nullptr;
			io.n_bases = cb;
			io.base0 = c0;
			io.n_windows = cw;
			io.d_offsets = (const uint64_t*)tk.offsets.p;
			io.n_seqs = h.n_seqs;
			io.d_hit = need_hit_dev ? (uint32_t*)sl.hit.p : nullptr;
			io.d_valid = want_valid ? (uint32_t*)sl.valid.p : nullptr;
			io.d_counts = h.counts ? (uint8_t*)sl.counts.p : nullptr;
			io.d_hashes = h.hashes ? (uint64_t*)sl.hashes.p : nullptr;
			io.d_strands = h.strands ? (uint8_t*)sl.strands.p : nullptr;
			io.d_stats = (uint64_t*)tk.d_stats;
			if (io.d_counts) CU(cudaMemsetAsync(io.d_counts, 0, cw, s));
			if (io.d_hashes) CU(cudaMemsetAsync(io.d_hashes, 0, cw * H * 8, s));
			if (io.d_strands) CU(cudaMemsetAsync(io.d_strands, 0, cw * H, s));
			if (op == PUB_HASH) {
				SeqParams P = hc.proto;
				P.fm = make_fastmod(1);
				P.force_generic = (uint32_t)ctx->force_generic;
				fill_io(P, io);
				TRY(launch(ctx, OP_HASH, P, s));
			} else {
				TRY(filter_op_dev(f, op, io, s));
			}
			CU(cudaEventRecord(sl.ev_kernel, s));
			CU(cudaStreamWaitEvent(ctx->copy_out, sl.ev_kernel, 0));
			if (want_hit)
				CU(cudaMemcpyAsync(h.hit_bits + (c0 >> 3), sl.hit.p, words * 4, cudaMemcpyDeviceToHost, ctx->copy_out));
			if (want_valid)
				CU(cudaMemcpyAsync(h.valid_bits + (c0 >> 3), sl.valid.p, words * 4, cudaMemcpyDeviceToHost, ctx->copy_out));
			if (h.counts)
				CU(cudaMemcpyAsync(h.counts + c0, sl.counts.p, cw, cudaMemcpyDeviceToHost, ctx->copy_out));
			if (h.hashes)
				CU(cudaMemcpyAsync(h.hashes + c0 * H, sl.hashes.p, cw * H * 8, cudaMemcpyDeviceToHost, ctx->copy_out));
			if (h.strands)
				CU(cudaMemcpyAsync(h.strands + c0 * H, sl.strands.p, cw * H, cudaMemcpyDeviceToHost, ctx->copy_out));
			CU(cudaEventRecord(sl.ev_d2h, ctx->copy_out));
			sl.used = true;
		}
		if (op == PUB_INSERT && f)
			TRY(settle_if_wrapped(f));
		// copy_out is ordered after the last kernel: the call's statistics follow its results
		CU(cudaMemcpyAsync(async ? (void*)counts_out : (void*)ctx->h_scalars, tk.d_stats, 16, cudaMemcpyDeviceToHost,
		                   ctx->copy_out));
		CU(cudaEventRecord(tk.done, ctx->copy_out));
		tk.in_use = true;
		return BTLBF_OK;
	};
	int rc = run_chunks();
	if (async && rc == BTLBF_OK)
		return BTLBF_OK;
	// drain: everything queued on the three streams must finish before the caller reads its buffers
	cudaError_t e1 = cudaStreamSynchronize(ctx->copy_in);
	cudaError_t e2 = cudaStreamSynchronize(s);
	cudaError_t e3 = cudaStreamSynchronize(ctx->copy_out);
	for (int i = 0; i < kTickets; i++)
		ctx->ticket[i].in_use = false;
	if (rc != BTLBF_OK)
		return rc;
	CU(e1);
	CU(e2);
	CU(e3);
	if (h.n_kmers) *h.n_kmers = ctx->h_scalars[0];
	if (h.n_hits) *h.n_hits = ctx->h_scalars[1];
	return BTLBF_OK;
}

extern "C" int btlbf_insert_seqs(btlbf_filter* f, const char* bases, const uint64_t* offsets, uint64_t n_seqs,
                                 uint64_t* n_kmers)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	HostIO h;
	h.bases = bases; h.offsets = offsets; h.n_seqs = n_seqs; h.n_kmers = n_kmers;
	return host_pipeline(f->ctx, f, nullptr, PUB_INSERT, h);
}

extern "C" int btlbf_contains_seqs(btlbf_filter* f, const char* bases, const uint64_t* offsets, uint64_t n_seqs,
                                   uint8_t* hit_bits, uint8_t* valid_bits, uint64_t* n_kmers, uint64_t* n_hits)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	HostIO h;
	h.bases = bases; h.offsets = offsets; h.n_seqs = n_seqs; h.hit_bits = hit_bits; h.valid_bits = valid_bits;
	h.n_kmers = n_kmers; h.n_hits = n_hits;
	return host_pipeline(f->ctx, f, nullptr, PUB_CONTAINS, h);
}

extern "C" int btlbf_insert_seqs_async(btlbf_filter* f, const char* bases, const uint64_t* offsets, uint64_t n_seqs,
                                       uint64_t* counts_out)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	HostIO h;
	h.bases = bases; h.offsets = offsets; h.n_seqs = n_seqs;
	uint64_t dummy[2];
	return host_pipeline(f->ctx, f, nullptr, PUB_INSERT, h, counts_out != nullptr, counts_out ? counts_out : dummy);
}

extern "C" int btlbf_contains_seqs_async(btlbf_filter* f, const char* bases, const uint64_t* offsets, uint64_t n_seqs,
                                         uint8_t* hit_bits, uint8_t* valid_bits, uint64_t* counts_out)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	if (!counts_out)
		return fail(BTLBF_ERR_ARG, "counts_out is required (2 host words: n_kmers, n_hits)");
	HostIO h;
	h.bases = bases; h.offsets = offsets; h.n_seqs = n_seqs; h.hit_bits = hit_bits; h.valid_bits = valid_bits;
	return host_pipeline(f->ctx, f, nullptr, PUB_CONTAINS, h, true, counts_out);
}

// ---------------------------------------------------------------- 2-bit packed input
// The same calls for callers that hold their reads as 2 bits per base (codes A0 C1 G2 T3 + an optional invalid
// plane; btlbf_pack_seqs in pack.cu produces both from ASCII): a quarter to three eighths of the host->device
// bytes, and the kernels skip the classification of the input (tile_phase_a copies the planes as they are).
static int packed_host_call(btlbf_filter* f, PublicOp op, const uint8_t* codes, const uint8_t* invalid, const uint64_t* offsets,
                            uint64_t n_seqs, uint8_t* hit_bits, uint8_t* valid_bits, uint64_t* n_kmers, uint64_t* n_hits,
                            bool async, uint64_t* counts_out)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	HostIO h;
	h.bases = (const char*)codes; h.invalid = invalid; h.packed = true; h.offsets = offsets; h.n_seqs = n_seqs;
	h.hit_bits = hit_bits; h.valid_bits = valid_bits; h.n_kmers = n_kmers; h.n_hits = n_hits;
	return host_pipeline(f->ctx, f, nullptr, op, h, async, counts_out);
}

extern "C" int btlbf_insert_seqs_packed(btlbf_filter* f, const uint8_t* codes, const uint8_t* invalid, const uint64_t* offsets,
                                        uint64_t n_seqs, uint64_t* n_kmers)
{
	return packed_host_call(f, PUB_INSERT, codes, invalid, offsets, n_seqs, nullptr, nullptr, n_kmers, nullptr, false, nullptr);
}

extern "C" int btlbf_contains_seqs_packed(btlbf_filter* f, const uint8_t* codes, const uint8_t* invalid, const uint64_t* offsets,
                                          uint64_t n_seqs, uint8_t* hit_bits, uint8_t* valid_bits, uint64_t* n_kmers,
                                          uint64_t* n_hits)
{
	return packed_host_call(f, PUB_CONTAINS, codes, invalid, offsets, n_seqs, hit_bits, valid_bits, n_kmers, n_hits, false, nullptr);
}

extern "C" int btlbf_insert_seqs_packed_async(btlbf_filter* f, const uint8_t* codes, const uint8_t* invalid,
                                              const uint64_t* offsets, uint64_t n_seqs, uint64_t* counts_out)
{
	uint64_t dummy[2];
	return packed_host_call(f, PUB_INSERT, codes, invalid, offsets, n_seqs, nullptr, nullptr, nullptr, nullptr,
	                        counts_out != nullptr, counts_out ? counts_out : dummy);
}

extern "C" int btlbf_contains_seqs_packed_async(btlbf_filter* f, const uint8_t* codes, const uint8_t* invalid,
                                                const uint64_t* offsets, uint64_t n_seqs, uint8_t* hit_bits,
                                                uint8_t* valid_bits, uint64_t* counts_out)
{
	if (!counts_out)
		return fail(BTLBF_ERR_ARG, "counts_out is required (2 host words: n_kmers, n_hits)");
	return packed_host_call(f, PUB_CONTAINS, codes, invalid, offsets, n_seqs, hit_bits, valid_bits, nullptr, nullptr, true,
	                        counts_out);
}

static int packed_dev_call(btlbf_filter* f, PublicOp op, const void* d_codes, const void* d_invalid, uint64_t n_bases,
                           const uint64_t* d_offsets, uint64_t n_seqs, uint32_t* d_hit_bits, uint32_t* d_valid_bits,
                           uint64_t* d_stats)
{
	TRY(check_dev_args(f, d_codes, n_bases, d_offsets));
	if ((uintptr_t)d_invalid & 15u)
		return fail(BTLBF_ERR_ARG, "the device invalid plane must be 16-byte aligned");
	ChunkIO io;
	io.d_bases = (const uint8_t*)d_codes;
	io.d_invalid = (const uint8_t*)d_invalid;
	io.packed = true;
	io.n_bases = io.n_windows = n_bases;
	io.d_offsets = d_offsets;
	io.n_seqs = n_seqs;
	io.d_hit = d_hit_bits;
	io.d_valid = d_valid_bits;
	io.d_stats = d_stats;
	LOCKED(f->ctx);
	TRY(filter_op_dev(f, op, io, f->ctx->active));
	return op == PUB_INSERT ? settle_if_wrapped(f) : BTLBF_OK;
}

extern "C" int btlbf_insert_seqs_packed_dev(btlbf_filter* f, const void* d_codes, const void* d_invalid, uint64_t n_bases,
                                            const uint64_t* d_offsets, uint64_t n_seqs, uint64_t* d_stats)
{
	return packed_dev_call(f, PUB_INSERT, d_codes, d_invalid, n_bases, d_offsets, n_seqs, nullptr, nullptr, d_stats);
}

extern "C" int btlbf_contains_seqs_packed_dev(btlbf_filter* f, const void* d_codes, const void* d_invalid, uint64_t n_bases,
                                              const uint64_t* d_offsets, uint64_t n_seqs, uint32_t* d_hit_bits,
                                              uint32_t* d_valid_bits, uint64_t* d_stats)
{
	return packed_dev_call(f, PUB_CONTAINS, d_codes, d_invalid, n_bases, d_offsets, n_seqs, d_hit_bits, d_valid_bits, d_stats);
}

extern "C" int btlbf_insert_and_check_seqs(btlbf_filter* f, const char* bases, const uint64_t* offsets,
                                           uint64_t n_seqs, uint8_t* found_bits, uint8_t* valid_bits,
                                           uint64_t* n_kmers)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	HostIO h;
	h.bases = bases; h.offsets = offsets; h.n_seqs = n_seqs; h.hit_bits = found_bits; h.valid_bits = valid_bits;
	h.n_kmers = n_kmers;
	return host_pipeline(f->ctx, f, nullptr, PUB_INSERT_CHECK, h);
}

extern "C" int btlbf_mincount_seqs(btlbf_filter* f, const char* bases, const uint64_t* offsets, uint64_t n_seqs,
                                   uint8_t* counts, uint8_t* valid_bits, uint64_t* n_kmers)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	if (f->kind != BTLBF_COUNTING8)
		return fail(BTLBF_ERR_STATE, "mincount_seqs needs a counting filter");
	HostIO h;
	h.bases = bases; h.offsets = offsets; h.n_seqs = n_seqs; h.counts = counts; h.valid_bits = valid_bits;
	h.n_kmers = n_kmers;
	return host_pipeline(f->ctx, f, nullptr, PUB_MINCOUNT, h);
}

extern "C" int btlbf_increment_all_seqs(btlbf_filter* f, const char* bases, const uint64_t* offsets, uint64_t n_seqs,
                                        uint64_t* n_kmers)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	if (f->kind != BTLBF_COUNTING8)
		return fail(BTLBF_ERR_STATE, "increment_all_seqs needs a counting filter");
	HostIO h;
	h.bases = bases; h.offsets = offsets; h.n_seqs = n_seqs; h.n_kmers = n_kmers;
	return host_pipeline(f->ctx, f, nullptr, PUB_INCALL, h);
}

extern "C" int btlbf_hash_seqs(btlbf_ctx* ctx, unsigned hash_num, unsigned kmer_size, const char* const* seeds,
                               unsigned n_seeds, unsigned h2, const char* bases, const uint64_t* offsets,
                               uint64_t n_seqs, uint64_t* hashes, uint8_t* strands, uint8_t* valid_bits,
                               uint64_t* n_kmers)
{
	TRY(use(ctx));
	LOCKED(ctx);
	HashCfg hc;
	int rc = hashcfg_init(hc, kmer_size, hash_num, n_seeds ? seeds : nullptr, n_seeds, h2);
	if (rc == BTLBF_OK) {
		HostIO h;
		h.bases = bases; h.offsets = offsets; h.n_seqs = n_seqs; h.hashes = hashes; h.strands = strands;
		h.valid_bits = valid_bits; h.n_kmers = n_kmers;
		rc = host_pipeline(ctx, nullptr, &hc, PUB_HASH, h);
	}
	hashcfg_free(hc);
	return rc;
}

// ---------------------------------------------------------------- ordered-update statistics (tests / tuning)
extern "C" int btlbf_filter_ordered_stats(btlbf_filter* f, uint64_t* deferred, uint64_t* rounds)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	uint64_t d = f->deferred_total, r = f->rounds_total;
	if (f->d_ord) { // the cooperative path keeps its counters on the device
		TRY(use(f->ctx));
		LOCKED(f->ctx);
		uint32_t w[8];
		CU(cudaStreamSynchronize(joined(f->ctx)));
		CU(cudaMemcpy(w, f->d_ord, 32, cudaMemcpyDeviceToHost));
		r += w[3];
		d += w[4];
	}
	if (deferred) *deferred = d;
	if (rounds) *rounds = r;
	return BTLBF_OK;
}

// ---------------------------------------------------------------- synthetic inputs, probe
extern "C" int btlbf_synth_genome_dev(btlbf_ctx* ctx, void* d_out, uint64_t start, uint64_t n, uint64_t seed)
{
	TRY(use(ctx));
	LOCKED(ctx);
	if (n && !d_out)
		return fail(BTLBF_ERR_ARG, "null output");
	cudaError_t e = launch_synth_genome((uint8_t*)d_out, start, n, seed, ctx->active);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "synth_genome launch failed: %s", cudaGetErrorString(e));
	ctx->launches++;
	return BTLBF_OK;
}

extern "C" int btlbf_synth_reads_dev(btlbf_ctx* ctx, void* d_out, uint64_t first_read, uint64_t n_reads,
                                     unsigned read_len, uint64_t g_start, uint64_t g_len, uint64_t genome_seed,
                                     uint64_t read_seed)
{
	TRY(use(ctx));
	LOCKED(ctx);
	if (n_reads && !d_out)
		return fail(BTLBF_ERR_ARG, "null output");
	if (read_len == 0 || g_len <= read_len)
		return fail(BTLBF_ERR_ARG, "need 0 < read_len < g_len");
	cudaError_t e = launch_synth_reads((uint8_t*)d_out, first_read, n_reads, read_len, g_start, g_len, genome_seed, read_seed,
	                                   ctx->active);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "synth_reads launch failed: %s", cudaGetErrorString(e));
	ctx->launches++;
	return BTLBF_OK;
}

extern "C" int btlbf_random_access_probe(btlbf_ctx* ctx, void* d_array, uint64_t bytes, uint64_t n_access, int mode,
                                         float* elapsed_ms)
{
	TRY(use(ctx));
	LOCKED(ctx);
	if (!d_array || bytes < 4 || !elapsed_ms || (mode != 0 && mode != 1))
		return fail(BTLBF_ERR_ARG, "bad probe arguments");
	cudaEvent_t a, b;
	CU(cudaEventCreate(&a));
	CU(cudaEventCreate(&b));
	CU(cudaEventRecord(a, ctx->active));
	cudaError_t e = launch_random_probe((uint32_t*)d_array, bytes / 4, n_access, mode, ctx->d_scalars + 3, ctx->active);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "probe launch failed: %s", cudaGetErrorString(e));
	ctx->launches++;
	CU(cudaEventRecord(b, ctx->active));
	CU(cudaEventSynchronize(b));
	CU(cudaEventElapsedTime(elapsed_ms, a, b));
	cudaEventDestroy(a);
	cudaEventDestroy(b);
	return BTLBF_OK;
}

// ---------------------------------------------------------------- legacy per-k-mer interface (precomputed hashes)
// runs `op` over n k-mers whose hashes are in host memory; blocking
static int hashes_run(btlbf_filter* f, int op, const uint64_t* hashes, uint64_t n, uint8_t* out)
{
	btlbf_ctx* ctx = f->ctx;
	cudaStream_t s;
	TRY(join(ctx, &s));
	Slot& sl = ctx->slot[0];
	uint32_t h = f->hc.h;
	// slot 0 also serves the host-buffer pipeline: nothing of it may be in flight while its buffers are reused
	CU(cudaStreamSynchronize(ctx->copy_in));
	CU(cudaStreamSynchronize(s));
	CU(cudaStreamSynchronize(ctx->copy_out));
	TRY(ensure(sl.hashes, n * h * 8));
	TRY(ensure(sl.counts, n));
	CU(cudaMemcpyAsync(sl.hashes.p, hashes, n * h * 8, cudaMemcpyHostToDevice, s));
	cudaError_t e = launch_hashes_op(op, f->d_data, make_fastmod(f->size), h, f->threshold,
	                                 (const uint64_t*)sl.hashes.p, n, (uint8_t*)sl.counts.p, s);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "hashes kernel launch failed: %s", cudaGetErrorString(e));
	ctx->launches++;
	if (out)
		CU(cudaMemcpyAsync(out, sl.counts.p, n, cudaMemcpyDeviceToHost, s));
	CU(cudaStreamSynchronize(s));
	return BTLBF_OK;
}

// applies the queued per-k-mer updates of ctx->hq_owner (joined() calls this before anything touches a filter)
static int hq_flush(btlbf_ctx* ctx)
{
	btlbf_filter* f = ctx->hq_owner;
	if (!f)
		return BTLBF_OK;
	ctx->hq_owner = nullptr;
	const uint64_t n = f->hq_n;
	f->hq_n = 0;
	if (n == 0)
		return BTLBF_OK;
	ctx->in_hq_flush = true;
	int rc = hashes_run(f, f->hq_op, f->hq, n, nullptr);
	ctx->in_hq_flush = false;
	return rc;
}

// appends n k-mers (n x h hash values) to the filter's queue; context lock held
static int hq_append(btlbf_filter* f, int op, const uint64_t* hashes, uint64_t n)
{
	btlbf_ctx* ctx = f->ctx;
	const uint32_t h = f->hc.h;
	while (n) {
		if (ctx->hq_owner && (ctx->hq_owner != f || f->hq_op != op || f->hq_n == kHashQueue))
			TRY(hq_flush(ctx));
		if (!f->hq)
			CU(cudaHostAlloc(&f->hq, (size_t)kHashQueue * h * 8, cudaHostAllocDefault));
		const uint64_t take = kHashQueue - f->hq_n < n ? kHashQueue - f->hq_n : n;
		memcpy(f->hq + f->hq_n * h, hashes, take * h * 8);
		f->hq_n += take;
		f->hq_op = op;
		ctx->hq_owner = f;
		hashes += take * h;
		n -= take;
		if (f->hq_n == kHashQueue)
			TRY(hq_flush(ctx));
	}
	return BTLBF_OK;
}

// moves one thread's queue into the filter's queue; context lock held
static int tl_drain_one(btlbf_filter* f, TlQueue* q)
{
	TlGuard g(q);
	if (q->n == 0)
		return BTLBF_OK;
	const uint32_t n = q->n;
	q->n = 0;
	f->ctx->tl_pending.fetch_sub(1, std::memory_order_acq_rel);
	return hq_append(f, q->op, q->data, n);
}

static int tl_drain_all(btlbf_ctx* ctx)
{
	int rc = BTLBF_OK;
	for (btlbf_filter* f : ctx->tl_filters)
		for (TlQueue* q : f->tlq) {
			int r = tl_drain_one(f, q);
			if (r != BTLBF_OK && rc == BTLBF_OK)
				rc = r;
		}
	return rc;
}

// drops what the threads' queues of f hold (clear / upload / destroy); context lock held
static void tl_discard(btlbf_filter* f, bool destroy)
{
	btlbf_ctx* ctx = f->ctx;
	for (TlQueue* q : f->tlq) {
		{
			TlGuard g(q);
			if (q->n) {
				q->n = 0;
				ctx->tl_pending.fetch_sub(1, std::memory_order_acq_rel);
			}
		}
		if (destroy) {
			free(q->data);
			delete q;
		}
	}
	if (destroy) {
		f->tlq.clear();
		for (size_t i = 0; i < ctx->tl_filters.size(); i++)
			if (ctx->tl_filters[i] == f) {
				ctx->tl_filters.erase(ctx->tl_filters.begin() + (long)i);
				break;
			}
	}
}

// this thread's queue in front of filter f (created and registered on first use)
static TlQueue* tl_queue(btlbf_filter* f)
{
	struct Entry { uint64_t uid; TlQueue* q; };
	static thread_local Entry cache[4] = { { 0, nullptr }, { 0, nullptr }, { 0, nullptr }, { 0, nullptr } };
	static thread_local unsigned next = 0;
	for (const Entry& e : cache)
		if (e.uid == f->uid)
			return e.q;
	const std::thread::id me = std::this_thread::get_id();
	TlQueue* q = nullptr;
	{
		LOCKED(f->ctx);
		for (TlQueue* t : f->tlq) // a thread that works on more filters than its cache holds finds its queue again
			if (t->owner == me) {
				q = t;
				break;
			}
		if (!q) {
			q = new (std::nothrow) TlQueue;
			if (!q)
				return nullptr;
			q->data = (uint64_t*)malloc((size_t)kTlCap * f->hc.h * 8);
			if (!q->data) {
				delete q;
				return nullptr;
			}
			q->owner = me;
			if (f->tlq.empty())
				f->ctx->tl_filters.push_back(f);
			f->tlq.push_back(q);
		}
	}
	cache[next++ & 3u] = { f->uid, q };
	return q;
}

static int hashes_op(btlbf_filter* f, int op, const uint64_t* hashes, uint64_t n, uint8_t* out)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	if (n == 0)
		return BTLBF_OK;
	if (!hashes)
		return fail(BTLBF_ERR_ARG, "null hashes");
	btlbf_ctx* ctx = f->ctx;
	const uint32_t h = f->hc.h;
	// Updates that report nothing (a `while (itr != itr.end()) { bloom.insert(*itr); ++itr; }` loop, README.md:30-43,
	// possibly from many OpenMP threads, ParallelFilter.cpp:104-122) are queued: first in the calling thread's own
	// queue, kTlCap k-mers at a time in the filter's, one kernel per kHashQueue k-mers.  A thread's updates keep their
	// call order, which is what the order-dependent incrementMin (op 3) needs.
	if (!out && (op == 0 || op == 3 || op == 4) && n <= kHashQueue / 4) {
		TlQueue* q = n <= 64 ? tl_queue(f) : nullptr;
		if (q) {
			for (;;) {
				bool full = false, done = false;
				{
					TlGuard g(q);
					if (q->n == 0 || (q->op == op && q->n + n <= kTlCap)) {
						if (q->n == 0)
							ctx->tl_pending.fetch_add(1, std::memory_order_acq_rel);
						memcpy(q->data + (size_t)q->n * h, hashes, n * h * 8);
						q->n += (uint32_t)n;
						q->op = op;
						full = q->n == kTlCap;
						done = true;
					}
				}
				if (done && !full)
					return BTLBF_OK;
				TRY(use(ctx));
				LOCKED(ctx);
				TRY(tl_drain_one(f, q));
				if (done)
					return BTLBF_OK;
			}
		}
		TRY(use(ctx));
		LOCKED(ctx);
		for (TlQueue* t : f->tlq) // larger batches go straight to the filter's queue, behind what the threads queued
			TRY(tl_drain_one(f, t));
		return hq_append(f, op, hashes, n);
	}
	TRY(use(ctx));
	LOCKED(ctx);
	return hashes_run(f, op, hashes, n, out);
}

extern "C" int btlbf_insert_hashes(btlbf_filter* f, const uint64_t* hashes, uint64_t n_kmers, uint8_t* found)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	if (f->kind == BTLBF_BLOOM)
		return hashes_op(f, found ? 5 : 0, hashes, n_kmers, found);
	return hashes_op(f, 3, hashes, n_kmers, found);
}

extern "C" int btlbf_contains_hashes(btlbf_filter* f, const uint64_t* hashes, uint64_t n_kmers, uint8_t* hit)
{
	if (!f || (!hit && n_kmers))
		return fail(BTLBF_ERR_ARG, "null argument");
	if (f->kind == BTLBF_BLOOM)
		return hashes_op(f, 1, hashes, n_kmers, hit);
	TRY(hashes_op(f, 2, hashes, n_kmers, hit));
	for (uint64_t i = 0; i < n_kmers; i++)
		hit[i] = hit[i] >= f->threshold; // CountingBloomFilter.hpp:190-196
	return BTLBF_OK;
}

extern "C" int btlbf_mincount_hashes(btlbf_filter* f, const uint64_t* hashes, uint64_t n_kmers, uint8_t* counts)
{
	if (!f || (!counts && n_kmers))
		return fail(BTLBF_ERR_ARG, "null argument");
	if (f->kind != BTLBF_COUNTING8)
		return fail(BTLBF_ERR_STATE, "mincount_hashes needs a counting filter");
	return hashes_op(f, 2, hashes, n_kmers, counts);
}

extern "C" int btlbf_increment_all_hashes(btlbf_filter* f, const uint64_t* hashes, uint64_t n_kmers)
{
	if (!f)
		return fail(BTLBF_ERR_ARG, "null filter");
	if (f->kind != BTLBF_COUNTING8)
		return fail(BTLBF_ERR_STATE, "increment_all_hashes needs a counting filter");
	return hashes_op(f, 4, hashes, n_kmers, nullptr);
}

// ---------------------------------------------------------------- file layout (BTLBloomFilter_v1 / BTLCountingBloomFilter_v1)
// The reference writes the header through cpptoml (BloomFilter.hpp:264-288, CountingBloomFilter.hpp:
// 344-367): one table, tab-indented "key = value" lines in the iteration order of the
// std::unordered_map that cpptoml keeps the keys in (libstdc++), doubles with showpoint and 17
// significant digits (cpptoml.h:3477-3494), then "[HeaderEnd]\n" and the raw array.
static std::string toml_double(double v)
{
	char tmp[64];
	snprintf(tmp, sizeof tmp, "%#.17g", v);
	std::string s(tmp);
	size_t p;
	if ((p = s.find("e-0")) != std::string::npos) s.erase(p + 2, 1);
	else if ((p = s.find("e+0")) != std::string::npos) s.erase(p + 2, 1);
	else if ((p = s.find("e0")) != std::string::npos) s.erase(p + 1, 1);
	return s;
}

static std::string format_header(int kind, uint64_t size, uint64_t size_bytes, unsigned h, unsigned k, double dFPR,
                                 uint64_t nEntry, uint64_t tEntry)
{
	char buf[640];
	if (kind == BTLBF_BLOOM)
		snprintf(buf, sizeof buf,
		         "[BTLBloomFilter_v1]\n\tnEntry = %llu\n\tdFPR = %s\n\tEntry = %llu\n\tBloomFilterSizeInBytes = %llu\n"
		         "\tBloomFilterSize = %llu\n\tHashNum = %u\n\tKmerSize = %u\n[HeaderEnd]\n",
		         (unsigned long long)nEntry, toml_double(dFPR).c_str(), (unsigned long long)tEntry,
		         (unsigned long long)size_bytes, (unsigned long long)size, h, k);
	else
		snprintf(buf, sizeof buf,
		         "[BTLCountingBloomFilter_v1]\n\tBloomFilterSize = %llu\n\tHashNum = %u\n\tKmerSize = %u\n"
		         "\tBloomFilterSizeInBytes = %llu\n\tBitsPerCounter = 8\n[HeaderEnd]\n",
		         (unsigned long long)size, h, k, (unsigned long long)size_bytes);
	return buf;
}

extern "C" int btlbf_format_header(int kind, uint64_t size, uint64_t size_bytes, unsigned hash_num, unsigned kmer_size,
                                   double dFPR, uint64_t nEntry, uint64_t tEntry, char* buf, size_t cap, size_t* len)
{
	if (kind != BTLBF_BLOOM && kind != BTLBF_COUNTING8)
		return fail(BTLBF_ERR_ARG, "unknown filter kind %d", kind);
	std::string hdr = format_header(kind, size, size_bytes, hash_num, kmer_size, dFPR, nEntry, tEntry);
	if (len) *len = hdr.size();
	if (buf) {
		if (cap < hdr.size() + 1)
			return fail(BTLBF_ERR_ARG, "header buffer too small");
		memcpy(buf, hdr.c_str(), hdr.size() + 1);
	}
	return BTLBF_OK;
}

struct ParsedHeader
{
	uint64_t size = 0, size_bytes = 0, nEntry = 0, tEntry = 0;
	unsigned h = 0, k = 0, bits_per_counter = 8;
	double dFPR = 0;
	unsigned seen = 0; // bit per key
	size_t body_offset = 0;
};

// Mirrors loadHeader (BloomFilter.hpp:116-166, CountingBloomFilter.hpp:283-329): first line must be
// "[magic]", lines are collected up to "[HeaderEnd]", and every key the reference dereferences must exist.
// next(): the next byte of the header text, or EOF
template<class Next>
static int parse_header_from(Next&& next, int kind, ParsedHeader& H)
{
	const char* magic = kind == BTLBF_BLOOM ? "[BTLBloomFilter_v1]" : "[BTLCountingBloomFilter_v1]";
	std::string line;
	size_t consumed = 0;
	auto getline = [&](std::string& out) -> bool {
		out.clear();
		int c;
		bool any = false;
		while ((c = next()) != EOF) {
			any = true;
			consumed++;
			if (c == '\n')
				return true;
			out.push_back((char)c);
		}
		return any;
	};
	if (!getline(line) || line != magic)
		return fail(BTLBF_ERR_ARG, "magic string does not match (likely version mismatch): '%.80s' vs '%s'",
		            line.c_str(), magic);
	bool end = false;
	while (getline(line)) {
		if (line == "[HeaderEnd]") {
			end = true;
			break;
		}
		size_t eq = line.find('=');
		if (eq == std::string::npos)
			continue;
		auto trim = [](std::string s) {
			size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
			return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
		};
		std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
		size_t hash = val.find('#');
		if (hash != std::string::npos)
			val = trim(val.substr(0, hash));
		std::string digits;
		for (char c : val)
			if (c != '_')
				digits.push_back(c);
		if (key == "BloomFilterSize") { H.size = strtoull(digits.c_str(), nullptr, 10); H.seen |= 1; }
		else if (key == "HashNum") { H.h = (unsigned)strtoull(digits.c_str(), nullptr, 10); H.seen |= 2; }
		else if (key == "KmerSize") { H.k = (unsigned)strtoull(digits.c_str(), nullptr, 10); H.seen |= 4; }
		else if (key == "BloomFilterSizeInBytes") { H.size_bytes = strtoull(digits.c_str(), nullptr, 10); H.seen |= 8; }
		else if (key == "dFPR") { H.dFPR = strtod(digits.c_str(), nullptr); H.seen |= 16; }
		else if (key == "nEntry") { H.nEntry = strtoull(digits.c_str(), nullptr, 10); H.seen |= 32; }
		else if (key == "Entry") { H.tEntry = strtoull(digits.c_str(), nullptr, 10); H.seen |= 64; }
		else if (key == "BitsPerCounter") { H.bits_per_counter = (unsigned)strtoull(digits.c_str(), nullptr, 10); H.seen |= 128; }
	}
	if (!end)
		return fail(BTLBF_ERR_ARG, "pre-built bloom filter does not have the correct header end.");
	unsigned need = kind == BTLBF_BLOOM ? (1 | 2 | 4 | 8 | 16 | 32 | 64) : (1 | 2 | 4 | 8 | 128);
	if ((H.seen & need) != need)
		return fail(BTLBF_ERR_ARG, "filter header lacks a required key (mask %#x of %#x present)", H.seen & need, need);
	H.body_offset = consumed;
	return BTLBF_OK;
}

static int parse_header(FILE* fp, int kind, ParsedHeader& H)
{
	return parse_header_from([&]() { return fgetc(fp); }, kind, H);
}

extern "C" int btlbf_parse_header(int kind, const char* text, size_t len, uint64_t* size, uint64_t* size_bytes,
                                  unsigned* hash_num, unsigned* kmer_size, double* dFPR, uint64_t* nEntry,
                                  uint64_t* tEntry, size_t* header_len)
{
	if (kind != BTLBF_BLOOM && kind != BTLBF_COUNTING8)
		return fail(BTLBF_ERR_ARG, "unknown filter kind %d", kind);
	if (!text && len)
		return fail(BTLBF_ERR_ARG, "null argument");
	ParsedHeader H;
	size_t at = 0;
	TRY(parse_header_from([&]() { return at < len ? (int)(unsigned char)text[at++] : EOF; }, kind, H));
	if (kind == BTLBF_COUNTING8 && H.bits_per_counter != 8)
		return fail(BTLBF_ERR_ARG, "only 8-bit counting filters are supported (BitsPerCounter %u)", H.bits_per_counter);
	if (size) *size = H.size;
	if (size_bytes) *size_bytes = H.size_bytes;
	if (hash_num) *hash_num = H.h;
	if (kmer_size) *kmer_size = H.k;
	if (dFPR) *dFPR = H.dFPR;
	if (nEntry) *nEntry = H.nEntry;
	if (tEntry) *tEntry = H.tEntry;
	if (header_len) *header_len = H.body_offset;
	return BTLBF_OK;
}

extern "C" int btlbf_filter_store(btlbf_filter* f, const char* path, double dFPR, uint64_t nEntry, uint64_t tEntry)
{
	if (!f || !path)
		return fail(BTLBF_ERR_ARG, "null argument");
	if (f->bitvector)
		return fail(BTLBF_ERR_STATE, "a BTLBF_BITVECTOR has no file format of its own (download the words instead)");
	btlbf_ctx* ctx = f->ctx;
	TRY(use(ctx));
	LOCKED(ctx);
	{
		cudaStream_t s; // deferred updates first: a failure there must not produce a file that lacks k-mers
		TRY(join(ctx, &s));
	}
	FILE* fp = fopen(path, "wb");
	if (!fp)
		return fail(BTLBF_ERR_ARG, "error: `%s': %s", path, strerror(errno));
	std::string hdr = format_header(f->kind, f->size, f->bytes, f->hc.h, f->hc.k, dFPR, nEntry, tEntry);
	bool ok = fwrite(hdr.data(), 1, hdr.size(), fp) == hdr.size();
	// stream the array through a pinned bounce buffer
	const size_t CH = (size_t)64 << 20;
	void* bounce = nullptr;
	size_t bsz = f->bytes < CH ? (size_t)f->bytes : CH;
	cudaError_t e = cudaHostAlloc(&bounce, bsz ? bsz : 1, cudaHostAllocDefault);
	for (uint64_t off = 0; ok && e == cudaSuccess && off < f->bytes; off += CH) {
		size_t n = f->bytes - off < CH ? (size_t)(f->bytes - off) : CH;
		e = cudaMemcpyAsync(bounce, f->d_data + off, n, cudaMemcpyDeviceToHost, joined(ctx));
		if (e == cudaSuccess)
			e = cudaStreamSynchronize(ctx->active);
		if (e == cudaSuccess)
			ok = fwrite(bounce, 1, n, fp) == n;
	}
	if (bounce)
		cudaFreeHost(bounce);
	ok = ok && fflush(fp) == 0;
	int saved = errno;
	fclose(fp);
	if (e != cudaSuccess)
		return fail(BTLBF_ERR_CUDA, "reading the filter back failed: %s", cudaGetErrorString(e));
	if (!ok)
		return fail(BTLBF_ERR_ARG, "error: `%s': %s", path, strerror(saved));
	return BTLBF_OK;
}

extern "C" int btlbf_filter_load(btlbf_ctx* ctx, const char* path, int kind, unsigned threshold, btlbf_filter** out,
                                 double* dFPR, uint64_t* nEntry, uint64_t* tEntry)
{
	if (!ctx || !path || !out)
		return fail(BTLBF_ERR_ARG, "null argument");
	*out = nullptr;
	TRY(use(ctx));
	LOCKED(ctx);
	if (kind != BTLBF_BLOOM && kind != BTLBF_COUNTING8)
		return fail(BTLBF_ERR_ARG, "unknown filter kind %d", kind);
	FILE* fp = fopen(path, "rb");
	if (!fp)
		return fail(BTLBF_ERR_ARG, "error: `%s': %s", path, strerror(errno));
	ParsedHeader H;
	int rc = parse_header(fp, kind, H);
	if (rc != BTLBF_OK) {
		fclose(fp);
		return rc;
	}
	// BloomFilter: the array is BloomFilterSize/8 bytes (initSize, :389-399) and BloomFilterSizeInBytes
	// bytes are read into it; CountingBloomFilter<uint8_t>: BloomFilterSizeInBytes counters are read
	// and indexed modulo BloomFilterSize (:268-281).
	uint64_t size = H.size, body = H.size_bytes;
	if (kind == BTLBF_BLOOM && (size % 8 != 0 || body != size / 8)) {
		fclose(fp);
		return fail(BTLBF_ERR_ARG, "inconsistent header: BloomFilterSize %llu, BloomFilterSizeInBytes %llu",
		            (unsigned long long)size, (unsigned long long)body);
	}
	if (kind == BTLBF_COUNTING8 && (H.bits_per_counter != 8 || body != size)) {
		fclose(fp);
		return fail(BTLBF_ERR_ARG, "only 8-bit counting filters are supported (BitsPerCounter %u, size %llu, bytes %llu)",
		            H.bits_per_counter, (unsigned long long)size, (unsigned long long)body);
	}
	btlbf_filter* f = nullptr;
	rc = filter_make(ctx, kind, size, H.h, H.k, threshold, nullptr, 0, &f);
	if (rc != BTLBF_OK) {
		fclose(fp);
		return rc;
	}
	const size_t CH = (size_t)64 << 20;
	void* bounce = nullptr;
	cudaError_t e = cudaHostAlloc(&bounce, body < CH ? (size_t)(body ? body : 1) : CH, cudaHostAllocDefault);
	bool ok = true;
	for (uint64_t off = 0; ok && e == cudaSuccess && off < body; off += CH) {
		size_t n = body - off < CH ? (size_t)(body - off) : CH;
		ok = fread(bounce, 1, n, fp) == n;
		if (ok)
			e = cudaMemcpyAsync(f->d_data + off, bounce, n, cudaMemcpyHostToDevice, joined(ctx));
		if (ok && e == cudaSuccess)
			e = cudaStreamSynchronize(ctx->active);
	}
	if (bounce)
		cudaFreeHost(bounce);
	fclose(fp);
	if (e != cudaSuccess || !ok) {
		btlbf_filter_destroy(f);
		if (!ok)
			return fail(BTLBF_ERR_ARG, "error: `%s': file shorter than its header says", path);
		return fail(BTLBF_ERR_CUDA, "uploading the filter failed: %s", cudaGetErrorString(e));
	}
	if (dFPR) *dFPR = H.dFPR;
	if (nEntry) *nEntry = H.nEntry;
	if (tEntry) *tEntry = H.tEntry;
	*out = f;
	return BTLBF_OK;
}
