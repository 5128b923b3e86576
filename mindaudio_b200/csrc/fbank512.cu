// fbank512.cu -- placeholder until the specialised kernel lands (next commit).
#include "common.cuh"
namespace mafe {
bool fast_plan_supported(const mafe_frontend_desc*) { return false; }
int fast_plan_init(mafe_ctx*, mafe_plan*, const mafe_frontend_desc*) { return MAFE_E_UNSUPPORTED; }
void fast_plan_free(mafe_plan*) {}
int fast_tile_frames() { return 32; }
int fast_run(mafe_ctx*, const mafe_plan*, mafe_batch*, const void*, int, float, float*) { return MAFE_E_UNSUPPORTED; }
}  // namespace mafe
