// chain_decode.cu -- instantiations of decode_chain_kernel (few long streams, contiguous layout) and their dispatch.
#include "chain_kernels.cuh"
#include "launch.cuh"

namespace ctr {

template <bool RANGE>
static cudaError_t go(const LaunchCfg &cfg, const AnsParams &p) {
    if (cfg.pool) return launch_kernel(decode_chain_kernel<RANGE, kTablePool, false>, cfg, p);
    return p.model.alphabet <= 256 ? launch_kernel(decode_chain_kernel<RANGE, kTableLut, true>, cfg, p)
                                   : launch_kernel(decode_chain_kernel<RANGE, kTableLut, false>, cfg, p);
}

cudaError_t launch_decode_chain(const LaunchCfg &cfg, const AnsParams &p, bool range) {
    return range ? go<true>(cfg, p) : go<false>(cfg, p);
}

}  // namespace ctr
