// dev-only: one instantiation of the thread-per-system kernel for quick ptxas / SASS checks
#include "../../phoskintime_b200/csrc/local_tps.cuh"
#ifndef DEV_MODEL
#define DEV_MODEL SuccModel<5>
#endif
#ifndef DEV_MINB
#define DEV_MINB 3
#endif
#ifndef DEV_CSMEM
#define DEV_CSMEM false
#endif
template __global__ void pk::local_tps_kernel<pk::DEV_MODEL, DEV_MINB, true, DEV_CSMEM>(const pk::LocalArgs);
template __global__ void pk::local_tps_kernel<pk::DEV_MODEL, DEV_MINB, false, DEV_CSMEM>(const pk::LocalArgs);
