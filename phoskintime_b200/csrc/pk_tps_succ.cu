// Thread-per-system kernels, successive model (models/succmod.py): instantiations for 1..8 sites.
#include "pk_tps_launch.cuh"
namespace pkh {
cudaError_t launch_tps_succ(pk_handle_s* h, const pk::LocalArgs& a) { return dispatch_tps<pk::SuccModel>(h, a); }
}  // namespace pkh
