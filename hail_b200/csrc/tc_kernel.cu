// Kernel 2 (tcgen05 int8-sliced exact-integer sweep) -- placeholder until the tensor-core path lands.
#include "common.cuh"

namespace lrr {
bool tc_supported(const Ctx*) { return false; }
void tc_release(Ctx*) {}
int launch_tc_sweep(Ctx* c, const uint8_t*, int64_t, int64_t, cudaStream_t) {
  return fail(c, LRR_EINVAL, "tensor-core kernel not built");
}
}  // namespace lrr
