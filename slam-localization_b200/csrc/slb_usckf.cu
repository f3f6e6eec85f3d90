// placeholder until the USCKF kernels land
#include "slb_internal.h"
namespace slb {
int launch_usckf(int, int, bool, bool, const FilterArgs &, cudaStream_t) { return set_error(SLB_ERR_INVALID, "usckf kernels not built"); }
int launch_usckf_clone(int, const FilterArgs &, cudaStream_t) { return set_error(SLB_ERR_INVALID, "usckf kernels not built"); }
int launch_usckf_set_measurement(int, const FilterArgs &, cudaStream_t) { return set_error(SLB_ERR_INVALID, "usckf kernels not built"); }
}
