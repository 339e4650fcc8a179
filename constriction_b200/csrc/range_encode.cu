// range_encode.cu -- instantiations of range_encode_kernel (K3) and their dispatch.
#include "launch.cuh"
#include "range_kernels.cuh"

namespace ctr {

template <int BLOCK>
static cudaError_t go(const LaunchCfg &cfg, const AnsParams &p) {
    CTR_LAYOUT_DISPATCH(range_encode_kernel, BLOCK);
}

cudaError_t launch_range_encode(const LaunchCfg &cfg, const AnsParams &p) {
    if (cfg.block == (unsigned)kSmallBlock) return go<kSmallBlock>(cfg, p);
    if (cfg.block == (unsigned)kAnsBlock) return go<kAnsBlock>(cfg, p);
    return cudaErrorInvalidConfiguration;
}

}  // namespace ctr
