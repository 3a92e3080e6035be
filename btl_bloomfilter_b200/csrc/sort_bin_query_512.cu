// sort_bin_query_512.cu -- instantiates the sort-bin kernel (sort_bin.cuh), 512-thread CTAs, for the partitioned QUERY ((offset, window) pairs).
#include "sort_bin.cuh"

namespace btl {

const void* bin_sort_kernel_query_512(int h, bool spaced, bool pow2)
{
	return bin_sort_kernel_any<512, true>(h, spaced, pow2);
}

} // namespace btl
