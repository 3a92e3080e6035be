// sort_bin_build_512.cu -- instantiates the sort-bin kernel (sort_bin.cuh), 512-thread CTAs, for the partitioned BUILD (offsets only).
#include "sort_bin.cuh"

namespace btl {

const void* bin_sort_kernel_build_512(int h, bool spaced, bool pow2)
{
	return bin_sort_kernel_any<512, false>(h, spaced, pow2);
}

} // namespace btl
