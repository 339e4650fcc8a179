// range_decode.cu -- instantiations of range_decode_kernel (K4) and their dispatch.
#include "launch.cuh"
#include "range_kernels.cuh"

namespace ctr {

// SMALL (alphabet <= 256) only matters for the shared-memory quantile index
template <int BLOCK>
static cudaError_t go(const LaunchCfg &cfg, const AnsParams &p) {
    if (cfg.shared) {
        if (p.model.alphabet <= 256)
            return cfg.contig ? launch_kernel(range_decode_kernel<BLOCK, kTableLut, true, false, true>, cfg, p)
                              : launch_kernel(range_decode_kernel<BLOCK, kTableLut, false, false, true>, cfg, p);
        return cfg.contig ? launch_kernel(range_decode_kernel<BLOCK, kTableLut, true, false, false>, cfg, p)
                          : launch_kernel(range_decode_kernel<BLOCK, kTableLut, false, false, false>, cfg, p);
    }
    if (cfg.persym)
        return cfg.contig ? launch_kernel(range_decode_kernel<BLOCK, kTableGlobal, true, true, false>, cfg, p)
                          : launch_kernel(range_decode_kernel<BLOCK, kTableGlobal, false, true, false>, cfg, p);
    return cfg.contig ? launch_kernel(range_decode_kernel<BLOCK, kTableGlobal, true, false, false>, cfg, p)
                      : launch_kernel(range_decode_kernel<BLOCK, kTableGlobal, false, false, false>, cfg, p);
}

cudaError_t launch_range_decode(const LaunchCfg &cfg, const AnsParams &p) {
    if (cfg.gauss) {  // per-symbol Gaussian parameters (gauss_kernels.cuh); always a model index per symbol
        if (cfg.block == (unsigned)kSmallBlock)
            return cfg.contig ? launch_kernel(range_decode_kernel<kSmallBlock, kTableGauss, true, true, false>, cfg, p)
                              : launch_kernel(range_decode_kernel<kSmallBlock, kTableGauss, false, true, false>, cfg, p);
        if (cfg.block == (unsigned)kAnsBlock)
            return cfg.contig ? launch_kernel(range_decode_kernel<kAnsBlock, kTableGauss, true, true, false>, cfg, p)
                              : launch_kernel(range_decode_kernel<kAnsBlock, kTableGauss, false, true, false>, cfg, p);
        return cudaErrorInvalidConfiguration;
    }
    if (cfg.pool) {  // model set in shared memory, CTA size decided by the caller
        if (cfg.persym)
            return cfg.contig ? launch_kernel(range_decode_kernel<0, kTablePool, true, true, false>, cfg, p)
                              : launch_kernel(range_decode_kernel<0, kTablePool, false, true, false>, cfg, p);
        return cfg.contig ? launch_kernel(range_decode_kernel<0, kTablePool, true, false, false>, cfg, p)
                          : launch_kernel(range_decode_kernel<0, kTablePool, false, false, false>, cfg, p);
    }
    if (cfg.block == (unsigned)kDecBlockShared) {  // shared model, interleaved deal, large batch
        if (!cfg.shared || cfg.contig) return cudaErrorInvalidConfiguration;
        return p.model.alphabet <= 256 ? launch_kernel(range_decode_kernel<kDecBlockShared, kTableLut, false, false, true>, cfg, p)
                                       : launch_kernel(range_decode_kernel<kDecBlockShared, kTableLut, false, false, false>, cfg, p);
    }
    if (cfg.block == (unsigned)kSmallBlock) return go<kSmallBlock>(cfg, p);
    if (cfg.block == (unsigned)kAnsBlock) return go<kAnsBlock>(cfg, p);
    return cudaErrorInvalidConfiguration;
}

}  // namespace ctr
