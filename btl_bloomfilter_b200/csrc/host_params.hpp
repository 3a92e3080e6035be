// host_params.hpp -- host-side construction of the launch-invariant hashing constants of SeqParams
// (no CUDA calls; used by capi.cu and by the CPU emulation harness under tests/emu/).
#pragma once
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.cuh"

namespace btl {

struct HostSeedTables
{
	std::vector<uint64_t> tab; // TF[k][8] then TR[k][8]
	std::vector<uint16_t> dc;  // concatenated don't-care positions
};

// seeds == nullptr / n_seeds == 0: contiguous ntHash (ntHashIterator) with h hashes per k-mer;
// else stHashIterator: n_seeds masks of exactly k characters ('1' = care, stHashIterator.hpp:23-33)
// times h2 hashes per mask.  Returns "" on success, else the reason.
inline std::string build_hash_proto(SeqParams& P, HostSeedTables& t, unsigned k, unsigned h,
                                    const char* const* seeds, unsigned n_seeds, unsigned h2)
{
	char msg[160];
	if (k == 0)
		return "kmer_size must be >= 1";
	if (k > 65535)
		return "kmer_size too large (max 65535)";
	if (n_seeds) {
		if (!seeds || h2 == 0 || n_seeds > (unsigned)kMaxSeeds || (uint64_t)n_seeds * h2 > (uint64_t)kMaxHash) {
			snprintf(msg, sizeof msg, "bad spaced-seed set (n_seeds %u, h2 %u)", n_seeds, h2);
			return msg;
		}
		h = n_seeds * h2;
	}
	if (h == 0 || h > (unsigned)kMaxHash) {
		snprintf(msg, sizeof msg, "hash_num %u out of range [1,%d]", h, kMaxHash);
		return msg;
	}
	memset(&P, 0, sizeof P);
	P.k = k;
	P.h = h;
	P.n_seeds = n_seeds;
	P.h2 = n_seeds ? h2 : 0;
	for (unsigned i = 0; i < (unsigned)kMaxHash; i++)
		P.mult[i] = multi_mult(i, k); // nthash.hpp:686
	for (unsigned c = 0; c < 16; c++) {
		P.g_f[c] = class_fseed(c);
		P.g_fk[c] = srol_n(class_fseed(c), k);
		P.g_r[c] = class_rseed(c);
		P.g_rk[c] = srol_n(class_rseed(c), k);
	}
	t.tab.clear();
	t.dc.clear();
	if (n_seeds) {
		for (unsigned j = 0; j < n_seeds; j++) {
			if (!seeds[j] || strlen(seeds[j]) != k) {
				snprintf(msg, sizeof msg, "spaced seed %u must have exactly kmer_size (%u) characters", j, k);
				return msg;
			}
			P.st_dc_off[j] = (uint32_t)t.dc.size();
			for (unsigned p = 0; p < k; p++)
				if (seeds[j][p] != '1')
					t.dc.push_back((uint16_t)p);
		}
		P.st_dc_off[n_seeds] = (uint32_t)t.dc.size();
		t.tab.resize((size_t)k * 16);
		for (unsigned p = 0; p < k; p++)
			for (unsigned c = 0; c < 8; c++) {
				t.tab[(size_t)p * 8 + c] = srol_n(class_fseed(c), k - 1 - p);         // nthash.hpp:843
				t.tab[(size_t)k * 8 + (size_t)p * 8 + c] = srol_n(class_rseed(c), p); // nthash.hpp:844
			}
		if (seq_kernel_smem_bytes(k, true) > 200 * 1024)
			return "kmer_size too large for spaced seeds";
	}
	if (seq_kernel_smem_bytes(k, false) > 200 * 1024)
		return "kmer_size too large";
	return "";
}

} // namespace btl
