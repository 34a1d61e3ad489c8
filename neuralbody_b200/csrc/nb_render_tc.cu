// tcgen05 render kernel (NB_PRECISION_TC_FP16) -- placeholder until the kernel lands.
#include "nb_internal.h"
namespace nb {
bool tc_available() { return false; }
int launch_render_tc(const RenderParams&, int, cudaStream_t) {
    set_error("NB_PRECISION_TC_FP16 is not built yet");
    return NB_ERR_UNSUPPORTED;
}
}  // namespace nb
