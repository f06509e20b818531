// Tensor-core (tcgen05 / TMEM) emission scoring -- placeholder until the kernel lands.
#include "common.cuh"
namespace loe {
int emission_tc_launch(const float*, int64_t, const float*, const float*, const float*, int, float*, int, cudaStream_t) {
    set_error("tensor-core emission path not built yet");
    return LOE_ERR_UNSUPPORTED;
}
}  // namespace loe
