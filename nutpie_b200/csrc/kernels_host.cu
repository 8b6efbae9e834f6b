// kernels_host.cu — sm_100a kernels of the NUTS engine for HostModel, the reference's host
// plug-in ABI served through a mailbox in mapped pinned memory (models.cuh, nb200_api.cu)
#include "launch_impl.cuh"

namespace nb200 {
NB200_INSTANTIATE_MODEL(HostModel)
}
