// seq_ops_bloom.cu -- instantiates seq_kernel for the raw hashes and the direct BloomFilter insert / contains.
#include "seq_kernel.cuh"

namespace btl {

cudaError_t launch_seq_bloom(SeqOp op, const SeqParams& P, cudaStream_t stream)
{
	switch (op) {
	case OP_HASH: return launch_op<OP_HASH>(P, stream);
	case OP_BF_INSERT: return launch_op<OP_BF_INSERT>(P, stream);
	case OP_BF_CONTAINS: return launch_op<OP_BF_CONTAINS>(P, stream);
	default: return cudaErrorInvalidValue;
	}
}

} // namespace btl
