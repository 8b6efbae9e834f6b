// kernels_funnel.cu — sm_100a kernels of the NUTS engine for FunnelModel (see launch_impl.cuh)
#include "launch_impl.cuh"

namespace nb200 {
NB200_INSTANTIATE_MODEL(FunnelModel)
}
