// seq_commit_cbf.cu -- instantiates seq_kernel for pass 2 of the exact counting insert (OP_CBF_COMMIT).
#include "seq_kernel.cuh"

namespace btl {

cudaError_t launch_seq_cbf_commit(const SeqParams& P, cudaStream_t stream)
{
	return launch_op<OP_CBF_COMMIT>(P, stream);
}

} // namespace btl
