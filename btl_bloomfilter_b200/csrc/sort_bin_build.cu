// sort_bin_build.cu -- instantiates the sort-bin kernel (sort_bin.cuh) for the partitioned BUILD (offsets only).
#include "sort_bin.cuh"

namespace btl {

const void* bin_sort_kernel_build(int h, bool spaced, bool pow2)
{
	return bin_sort_kernel_any<false>(h, spaced, pow2);
}

} // namespace btl
