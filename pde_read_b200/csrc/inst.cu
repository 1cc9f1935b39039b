// One translation unit per (activation, NX, NT): compiled several times with
//   -DPDR_ACT=<0|1|2> -DPDR_NX=<n> -DPDR_NT=<m> -DPDR_BWD=<0|1>
// so that the template instantiations build in parallel (see pde_read_b200/build.py).
#include "mlp_jet_kernels.cuh"
#include "mlp_plain_kernels.cuh"
#include "tc_jet_kernels.cuh"

namespace pdr {
template int launch_fwd<PDR_ACT, PDR_NX, PDR_NT>(const FwdArgs&, cudaStream_t);
#if PDR_NX == 0 && PDR_NT == 0
template int launch_plain_fwd<PDR_ACT>(const FwdArgs&, cudaStream_t);
template int plain_bwd_grid<PDR_ACT>(int, int, int, long long, int*);
template int launch_plain_bwd<PDR_ACT>(const BwdArgs&, int, cudaStream_t);
#endif
template int launch_tc_fwd<PDR_ACT, PDR_NX, PDR_NT>(const TcFwdArgs&, cudaStream_t);
#if PDR_BWD
template int bwd_grid<PDR_ACT, PDR_NX, PDR_NT>(int, int, int, long long, int*);
template int launch_bwd_impl<PDR_ACT, PDR_NX, PDR_NT>(const BwdArgs&, int, cudaStream_t);
#endif
}  // namespace pdr
