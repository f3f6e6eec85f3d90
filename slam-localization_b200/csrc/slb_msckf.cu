// placeholder until the MSCKF kernels land
#include "slb_internal.h"
namespace slb {
int launch_msckf_predict(int, const FilterArgs &, cudaStream_t) { return set_error(SLB_ERR_INVALID, "msckf kernels not built"); }
int launch_msckf_update(int, const FilterArgs &, cudaStream_t) { return set_error(SLB_ERR_INVALID, "msckf kernels not built"); }
}
