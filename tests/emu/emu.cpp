// emu.cpp -- TEST INFRASTRUCTURE ONLY: runs the per-thread phase functions of the CUDA kernels
// (btl_bloomfilter_b200/csrc/tile_core.cuh, the very source nvcc compiles for sm_100a) thread by
// thread on the CPU, with the same chunk / batch / residual-round orchestration as capi.cu, so the
// kernel logic can be checked against the oracle in a container without a GPU.  It is not a
// fallback: nothing in the product loads this file, and it is not built by __graft_entry__.build().
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "tile_core.cuh"
#include "host_params.hpp"

namespace btl {
size_t seq_kernel_smem_bytes(uint32_t k, bool spaced) { return tile_smem_bytes(k, spaced); }
}
using namespace btl;

namespace {

template<int OP, bool SPACED, bool POW2>
void run_grid(const SeqParams& P)
{
	uint64_t tiles = (P.n_windows + kTile - 1) / kTile;
	std::vector<uint8_t> raw(tile_smem_bytes(P.k, SPACED) + 64);
	uint8_t* base = raw.data() + ((16 - ((uintptr_t)raw.data() & 15)) & 15);
	for (uint64_t b = 0; b < tiles; b++) {
		TileSmem sm = carve_smem(base, P.k, SPACED);
		uint64_t t0 = b * kTile;
		for (int tid = 0; tid < kTPB; tid++) tile_phase_a(P, sm, t0, tid, kTPB);
		for (int tid = 0; tid < kTPB; tid++) tile_phase_b(P, sm, t0, tid, kTPB);
		for (int tid = 0; tid < kTPB; tid++) {
			ThreadOut out = tile_phase_c<OP, SPACED, POW2>(P, sm, t0, tid);
			uint64_t widx = (t0 >> 5) + tid;
			if (widx < P.out_words) {
				if (P.valid_bits) P.valid_bits[widx] = out.validw;
				if (P.hit_bits) P.hit_bits[widx] = out.hitw;
			}
			if (P.stats) {
				P.stats[0] += __builtin_popcount(out.validw);
				P.stats[1] += __builtin_popcount(out.hitw);
			}
		}
	}
}

template<int OP>
void run_op(const SeqParams& P)
{
	bool sp = P.n_seeds != 0, p2 = P.fm.pow2 != 0;
	if (sp) { if (p2) run_grid<OP, true, true>(P); else run_grid<OP, true, false>(P); }
	else    { if (p2) run_grid<OP, false, true>(P); else run_grid<OP, false, false>(P); }
}

void run(SeqOp op, const SeqParams& P)
{
	switch (op) {
	case OP_HASH: run_op<OP_HASH>(P); break;
	case OP_BF_INSERT: run_op<OP_BF_INSERT>(P); break;
	case OP_BF_CONTAINS: run_op<OP_BF_CONTAINS>(P); break;
	case OP_CBF_MINCOUNT: run_op<OP_CBF_MINCOUNT>(P); break;
	case OP_CBF_INCALL: run_op<OP_CBF_INCALL>(P); break;
	case OP_RESV_TOUCH: run_op<OP_RESV_TOUCH>(P); break;
	case OP_CBF_COMMIT: run_op<OP_CBF_COMMIT>(P); break;
	case OP_RESV_CLEAR: run_op<OP_RESV_CLEAR>(P); break;
	case OP_BFCHK_COMMIT: run_op<OP_BFCHK_COMMIT>(P); break;
	}
}

// mirrors bin_kernel / apply_bins_kernel of kernels.cu (persistent writers, private sub-buckets)
template<bool SPACED, bool POW2>
void run_bin_grid(const SeqParams& P)
{
	uint64_t tiles = (P.n_windows + kTile - 1) / kTile;
	std::vector<uint8_t> raw(tile_smem_bytes(P.k, SPACED, P.n_bins) + 64);
	uint8_t* base = raw.data() + ((16 - ((uintptr_t)raw.data() & 15)) & 15);
	for (uint32_t w = 0; w < P.bin_writers; w++) {
		TileSmem sm = carve_smem(base, P.k, SPACED, P.n_bins);
		sm.writer = w;
		for (uint32_t i = 0; i < P.n_bins; i++) sm.cursors[i] = 0;
		for (uint64_t t = w; t < tiles; t += P.bin_writers) {
			uint64_t t0 = t * kTile;
			for (int tid = 0; tid < kTPB; tid++) tile_phase_a(P, sm, t0, tid, kTPB);
			for (int tid = 0; tid < kTPB; tid++) tile_phase_b(P, sm, t0, tid, kTPB);
			for (int tid = 0; tid < kTPB; tid++) {
				ThreadOut out = tile_phase_c<OP_BF_BIN, SPACED, POW2>(P, sm, t0, tid);
				uint64_t widx = (t0 >> 5) + tid;
				if (widx < P.out_words && P.valid_bits) P.valid_bits[widx] = out.validw;
				if (P.stats) P.stats[0] += __builtin_popcount(out.validw);
			}
		}
		for (uint32_t i = 0; i < P.n_bins; i++) P.bin_counts[(uint64_t)i * P.bin_writers + w] = sm.cursors[i];
	}
}

uint64_t g_bin_overflow = 0;

void binned_insert(SeqParams P, uint64_t size_bits, uint32_t shift, uint32_t writers, uint32_t slack_pct)
{
	uint64_t n_bins = (size_bits + (((uint64_t)1 << shift) - 1)) >> shift;
	P.n_bins = (uint32_t)n_bins;
	P.bin_shift = shift;
	P.bin_mask = (uint32_t)(((uint64_t)1 << shift) - 1);
	uint64_t tiles = (P.n_windows + kTile - 1) / kTile;
	P.bin_writers = (uint32_t)(tiles < writers ? (tiles ? tiles : 1) : writers);
	double parts = (double)size_bits / (double)((uint64_t)1 << shift);
	double expect = (double)P.n_windows * P.h / parts / P.bin_writers;
	uint64_t cap = (uint64_t)(expect * (1.0 + slack_pct / 100.0)) + 4;
	P.bin_cap = (uint32_t)((cap + 3) / 4 * 4);
	std::vector<uint32_t> items(n_bins * P.bin_writers * P.bin_cap), counts(n_bins * P.bin_writers, 0xdeadbeefu);
	P.bin_items = items.data();
	P.bin_counts = counts.data();
	bool sp = P.n_seeds != 0, p2 = P.fm.pow2 != 0;
	if (sp) { if (p2) run_bin_grid<true, true>(P); else run_bin_grid<true, false>(P); }
	else    { if (p2) run_bin_grid<false, true>(P); else run_bin_grid<false, false>(P); }
	uint32_t* words = (uint32_t*)P.filter;
	for (uint64_t part = 0; part < n_bins; part++)
		for (uint32_t w = 0; w < P.bin_writers; w++) {
			uint32_t n = counts[part * P.bin_writers + w];
			if (n > P.bin_cap) { g_bin_overflow += n - P.bin_cap; n = P.bin_cap; }
			const uint32_t* it = items.data() + (part * P.bin_writers + w) * P.bin_cap;
			uint32_t* region = words + (part << (shift - 5));
			for (uint32_t i = 0; i < n; i++) region[it[i] >> 5] |= 1u << (it[i] & 31);
		}
}

struct Ordered
{
	std::vector<uint32_t> touched, contended, pend[2];
	std::vector<uint64_t> resv;
	uint32_t resv_log2, list_log2, epoch = 0;
	uint64_t deferred = 0, rounds = 0;
};

// mirrors ordered_apply() of capi.cu; threads of a "launch" run in a scrambled order so that the
// result cannot depend on the order in which a real grid would execute them
void ordered_apply(Ordered& st, const SeqParams& chunk, int kind, uint64_t batch)
{
	for (uint64_t b0 = 0; b0 < chunk.n_windows; b0 += batch) {
		SeqParams P = chunk;
		uint64_t bw = chunk.n_windows - b0 < batch ? chunk.n_windows - b0 : batch;
		P.bases = chunk.bases + (chunk.packed ? b0 >> 2 : b0); // advance_input() of capi.cu
		P.invalid = chunk.invalid ? chunk.invalid + (b0 >> 3) : nullptr;
		P.n_bases = chunk.n_bases > b0 ? chunk.n_bases - b0 : 0;
		P.base0 = chunk.base0 + b0;
		P.n_windows = bw;
		if (chunk.hit_bits) P.hit_bits = chunk.hit_bits + (b0 >> 5);
		if (chunk.valid_bits) P.valid_bits = chunk.valid_bits + (b0 >> 5);
		P.out_words = chunk.out_words > (b0 >> 5) ? chunk.out_words - (b0 >> 5) : 0;
		P.resv_touched = st.touched.data();
		P.resv_contended = st.contended.data();
		P.resv_log2 = st.resv_log2;
		uint32_t cnt[2] = { 0, 0 };
		st.pend[0].assign(bw + 1, 0);
		st.pend[1].assign(bw + 1, 0);
		P.pending = st.pend[0].data();
		P.pending_count = &cnt[0];
		SeqParams T = P;
		T.hit_bits = T.valid_bits = nullptr;
		T.stats = nullptr;
		run(OP_RESV_TOUCH, T);
		run(kind == 0 ? OP_CBF_COMMIT : OP_BFCHK_COMMIT, P);
		run(OP_RESV_CLEAR, T);
		for (uint32_t v : st.touched) if (v) abort();   // the clear pass must leave the tables empty
		for (uint32_t v : st.contended) if (v) abort();
		st.deferred += cnt[0];
		int cur = 0;
		uint32_t n = cnt[0];
		uint64_t emask = ((uint64_t)1 << st.list_log2) - 1;
		while (n > 0) {
			uint32_t epoch = ++st.epoch;
			st.rounds++;
			// scrambled execution order
			std::vector<uint32_t> order(n);
			for (uint32_t i = 0; i < n; i++) order[i] = i;
			for (uint32_t i = n; i > 1; i--) std::swap(order[i - 1], order[(uint32_t)(splitmix64(epoch * 7919ull + i) % i)]);
			bool p2 = P.fm.pow2 != 0;
			for (uint32_t i : order) {
				uint32_t w = st.pend[cur][i];
				if (p2) list_round_reserve<true>(P, st.resv.data(), emask, epoch, w);
				else list_round_reserve<false>(P, st.resv.data(), emask, epoch, w);
			}
			uint32_t out = 0;
			for (uint32_t i : order) {
				uint32_t w = st.pend[cur][i];
				bool done;
				if (kind == 0)
					done = p2 ? list_round_commit<true, 0>(P, st.resv.data(), emask, epoch, w)
					          : list_round_commit<false, 0>(P, st.resv.data(), emask, epoch, w);
				else
					done = p2 ? list_round_commit<true, 1>(P, st.resv.data(), emask, epoch, w)
					          : list_round_commit<false, 1>(P, st.resv.data(), emask, epoch, w);
				if (!done)
					st.pend[1 - cur][out++] = w;
			}
			n = out;
			cur = 1 - cur;
		}
	}
}

} // namespace

extern "C" {

// pub_op: 0 insert, 1 contains, 2 insert_and_check, 3 mincount, 4 increment_all, 5 hash
// kind: 0 BloomFilter (size = bits), 1 CountingBloomFilter<uint8_t> (size = counters)
// filter: host array of round_up(bytes, 16) bytes.  Outputs indexed by flat window (see btlbf.h).
// info[0] = deferred k-mers, info[1] = residual rounds.  Returns 0, or -1 and msg on bad arguments.
int emu_seq_op(int pub_op, int kind, uint64_t size, unsigned h, unsigned k, unsigned threshold,
               const char* const* seeds, unsigned n_seeds, unsigned h2, uint8_t* filter, const uint8_t* bases,
               const uint64_t* offsets, uint64_t n_seqs, uint32_t* hit, uint32_t* valid, uint8_t* counts,
               uint64_t* hashes, uint8_t* strands, uint64_t* stats, int force_generic, int query_mode,
               uint64_t chunk, uint64_t batch, unsigned resv_log2, unsigned list_log2, uint64_t* info,
               char* msg, size_t msg_cap, unsigned bin_shift, unsigned bin_writers, unsigned bin_slack_pct,
               const uint8_t* invalid, int packed)
{
	// packed != 0: `bases` holds 2-bit codes (4 per byte) and `invalid` one bit per base or null (btlbf.h,
	// "2-bit packed input"); offsets keep counting bases
	SeqParams proto;
	HostSeedTables t;
	std::string err = build_hash_proto(proto, t, k, h, n_seeds ? seeds : nullptr, n_seeds, h2);
	if (!err.empty()) {
		snprintf(msg, msg_cap, "%s", err.c_str());
		return -1;
	}
	const uint32_t H = proto.h;
	proto.st_tab = t.tab.data();
	proto.st_dc = t.dc.data();
	proto.filter = filter;
	proto.fm = make_fastmod(pub_op == 5 ? 1 : size);
	proto.threshold = threshold;
	proto.force_generic = force_generic;
	proto.query_mode = query_mode;
	uint64_t n_bases = n_seqs ? offsets[n_seqs] : 0;
	if (n_bases == 0)
		return 0;
	chunk = chunk / kTile * kTile;
	batch = batch / kTile * kTile;
	if (chunk < (uint64_t)kTile) chunk = kTile;
	if (batch < (uint64_t)kTile) batch = kTile;
	Ordered st;
	st.resv_log2 = resv_log2;
	st.list_log2 = list_log2;
	st.touched.assign((((size_t)1 << resv_log2) + 31) / 32, 0);
	st.contended.assign((((size_t)1 << resv_log2) + 31) / 32, 0);
	st.resv.assign((size_t)1 << list_log2, ~0ull);
	for (uint64_t c0 = 0; c0 < n_bases; c0 += chunk) {
		uint64_t cw = n_bases - c0 < chunk ? n_bases - c0 : chunk;
		uint64_t cb = n_bases - c0 < cw + k - 1 ? n_bases - c0 : cw + k - 1;
		// the device chunk buffer is a private copy: reads past cb must not see the caller's bytes
		std::vector<uint8_t> dev(cb + 32), inv(cb / 8 + 32);
		if (packed) { // whole 16-byte words of both planes, the bytes past the copies hold junk
			memset(dev.data(), 0xb7, dev.size());
			memset(inv.data(), 0x00, inv.size());
			memcpy(dev.data(), bases + (c0 >> 2), (cb + 3) >> 2);
			if (invalid) memcpy(inv.data(), invalid + (c0 >> 3), (cb + 7) >> 3);
		} else {
			memcpy(dev.data() + 0, bases + c0, cb);
			memset(dev.data() + cb, 'A', 16);
		}
		SeqParams P = proto;
		P.packed = packed ? 1 : 0;
		P.invalid = packed && invalid ? inv.data() : nullptr;
		if (packed) P.force_generic = 0;
		P.bases = dev.data();
		P.n_bases = cb;
		P.base0 = c0;
		P.n_windows = cw;
		P.offsets = offsets;
		P.n_seqs = n_seqs;
		P.hit_bits = hit ? hit + (c0 >> 5) : nullptr;
		P.valid_bits = valid ? valid + (c0 >> 5) : nullptr;
		P.out_words = (cw + 31) / 32;
		P.counts = counts ? counts + c0 : nullptr;
		P.hashes = hashes ? hashes + c0 * H : nullptr;
		P.strands = strands ? strands + c0 * H : nullptr;
		P.stats = stats;
		switch (pub_op) {
		case 0:
			if (kind == 0 && bin_shift && size % 32 == 0) binned_insert(P, size, bin_shift, bin_writers, bin_slack_pct);
			else if (kind == 0) run(OP_BF_INSERT, P);
			else ordered_apply(st, P, 0, batch);
			break;
		case 1: run(kind == 0 ? OP_BF_CONTAINS : OP_CBF_MINCOUNT, P); break;
		case 2: ordered_apply(st, P, 1, batch); break;
		case 3: run(OP_CBF_MINCOUNT, P); break;
		case 4: run(OP_CBF_INCALL, P); break;
		case 5: run(OP_HASH, P); break;
		default: snprintf(msg, msg_cap, "bad op"); return -1;
		}
	}
	if (info) {
		info[0] = st.deferred;
		info[1] = st.rounds;
		info[2] = g_bin_overflow;
	}
	return 0;
}

} // extern "C"

// ---- arithmetic helpers of nthash_dev.cuh exposed for tests/test_host_arith.py
extern "C" {
uint64_t emu_fastmod(uint64_t x, uint64_t m)
{
	FastMod f = make_fastmod(m);
	return f.pow2 ? fastmod<true>(x, f) : fastmod<false>(x, f);
}
uint64_t emu_srol(uint64_t v) { return srol(v); }
uint64_t emu_sror(uint64_t v) { return sror(v); }
uint64_t emu_srol_n(uint64_t v, unsigned n) { return srol_n(v, n); }
uint64_t emu_multi_mix(uint64_t b, unsigned i, unsigned k) { return multi_mix(b, multi_mult(i, k)); }
uint64_t emu_class_seeds(unsigned c, int rev) { uint8_t cls = base_class(c); return rev ? class_rseed(cls) : class_fseed(cls); }
}
