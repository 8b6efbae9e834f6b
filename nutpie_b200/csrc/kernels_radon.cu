// kernels_radon.cu — sm_100a kernels of the NUTS engine for RadonModel (see launch_impl.cuh)
#include "launch_impl.cuh"

namespace nb200 {
NB200_INSTANTIATE_MODEL(RadonModel)
}
