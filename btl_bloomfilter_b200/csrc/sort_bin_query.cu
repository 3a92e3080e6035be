// sort_bin_query.cu -- instantiates the sort-bin kernel (sort_bin.cuh) for the partitioned QUERY ((offset, window) pairs).
#include "sort_bin.cuh"

namespace btl {

const void* bin_sort_kernel_query(int h, bool spaced, bool pow2)
{
	return bin_sort_kernel_any<true>(h, spaced, pow2);
}

} // namespace btl
