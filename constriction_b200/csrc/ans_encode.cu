// ans_encode.cu -- instantiations of ans_encode_kernel (K1) and their dispatch.
#include "launch.cuh"

namespace ctr {

template <int BLOCK>
static cudaError_t go(const LaunchCfg &cfg, const AnsParams &p) {
    if (cfg.f64) CTR_LAYOUT_DISPATCH(ans_encode_kernel, BLOCK, CTR_COMMA true);
    CTR_LAYOUT_DISPATCH(ans_encode_kernel, BLOCK, CTR_COMMA false);
}

cudaError_t launch_ans_encode(const LaunchCfg &cfg, const AnsParams &p) {
    if (cfg.block == (unsigned)kSmallBlock) return go<kSmallBlock>(cfg, p);
    if (cfg.block == (unsigned)kAnsBlock) return go<kAnsBlock>(cfg, p);
    return cudaErrorInvalidConfiguration;
}

}  // namespace ctr
