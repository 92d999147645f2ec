// Thread-per-system kernels, distributive model (models/distmod.py): instantiations for 1..8 sites.
#include "pk_tps_launch.cuh"
namespace pkh {
cudaError_t launch_tps_dist(pk_handle_s* h, const pk::LocalArgs& a) { return dispatch_tps<pk::DistModel>(h, a); }
}  // namespace pkh
