// placeholder replaced below in the same round
#include "common.cuh"
namespace pcs {
bool umma_supported(int, int) { return false; }
size_t umma_weight_image(const float*, int, const int*, int, int, int, int, std::vector<uint16_t>&) { return 0; }
int launch_conv_umma(pcs_ctx* ctx, const UmmaConvArgs&) { return set_err(ctx, PCS_ERR_STATE, "umma engine not built"); }
}
