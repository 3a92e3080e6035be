// sort_bin_query_256.cu -- instantiates the sort-bin kernel (sort_bin.cuh), 256-thread CTAs, for the partitioned QUERY ((offset, window) pairs).
#include "sort_bin.cuh"

namespace btl {

const void* bin_sort_kernel_query_256(int h, bool spaced, bool pow2)
{
	return bin_sort_kernel_any<256, true>(h, spaced, pow2);
}

} // namespace btl
