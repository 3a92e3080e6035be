// seq_commit_bfchk.cu -- instantiates seq_kernel for pass 2 of the exact insertAndCheck (OP_BFCHK_COMMIT).
#include "seq_kernel.cuh"

namespace btl {

cudaError_t launch_seq_bfchk_commit(const SeqParams& P, cudaStream_t stream)
{
	return launch_op<OP_BFCHK_COMMIT>(P, stream);
}

} // namespace btl
