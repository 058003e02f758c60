// tc_dispatch.cuh -- entry points of the tcgen05 (sm_100a tensor-core) implementations.
// Each function returns false when it does not handle the request, in which case the caller
// runs the CUDA-core kernels instead.
#pragma once
#include "tower.cuh"

namespace pgpd { namespace tc {

inline bool available() { return false; }
inline bool tower_forward_tc(const TowerArgs&, TowerWs&, float*) { return false; }
inline bool tower_backward_tc(const TowerArgs&, TowerWs&, const pgpd_tower_grad&, const float*, float*) { return false; }

}}  // namespace pgpd::tc
