// placeholder until the tcgen05 search kernel lands (next commit)
#include "vqb_common.cuh"
namespace vqb {
int forward_tensor_workspace(const vqb_fwd_args*, size_t* bytes) { *bytes = 0; return VQB_OK; }
int launch_forward_tensor(const vqb_fwd_args*, cudaStream_t) { return invalid("tensor-core search is not built into this library"); }
}
