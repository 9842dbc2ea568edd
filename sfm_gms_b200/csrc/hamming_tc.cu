// hamming_tc.cu — placeholder until the tcgen05 kernel lands: reports "not available" so the context
// selects the popc kernel.  (No CPU path: both kernels are sm_100a CUDA.)
#include "hamming_tc.cuh"

namespace sfmgms {
bool tc_available() { return false; }
const char* tc_last_error() { return "tensor-core Hamming kernel not built"; }
void tc_invalidate(TcState& s) { s.set_valid = false; }
void tc_release(TcState& s) {
    if (s.d_ops) cudaFree(s.d_ops);
    if (s.d_work) cudaFree(s.d_work);
    if (s.h_work) cudaFreeHost(s.h_work);
    s = TcState();
}
int launch_hamming_tc(TcState&, const PairDesc*, const PairDesc*, int, int, cudaStream_t) { return -1; }
}  // namespace sfmgms
