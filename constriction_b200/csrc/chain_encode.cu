// chain_encode.cu -- instantiations of encode_chain_kernel (few long streams, contiguous layout) and their dispatch.
#include "chain_kernels.cuh"
#include "launch.cuh"

namespace ctr {

cudaError_t launch_encode_chain(const LaunchCfg &cfg, const AnsParams &p, bool ans) {
    if (ans) {
        if (cfg.f64) return cfg.shared ? launch_kernel(encode_chain_kernel<true, true, true>, cfg, p) : launch_kernel(encode_chain_kernel<true, false, true>, cfg, p);
        return cfg.shared ? launch_kernel(encode_chain_kernel<true, true, false>, cfg, p) : launch_kernel(encode_chain_kernel<true, false, false>, cfg, p);
    }
    return cfg.shared ? launch_kernel(encode_chain_kernel<false, true, false>, cfg, p) : launch_kernel(encode_chain_kernel<false, false, false>, cfg, p);
}

}  // namespace ctr
