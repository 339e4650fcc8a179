// launch.cuh -- host-side launch interface between capi.cu and the kernel translation units.
//
// Every coder kernel family is instantiated in its own .cu file (ans_encode.cu, ans_decode.cu, range_encode.cu,
// range_decode.cu) so that the library builds in parallel; capi.cu decides the launch geometry and calls these.
#pragma once
#include <cuda_runtime.h>

#include "ans_kernels.cuh"

namespace ctr {

constexpr int kSmallBlock = 64;  // CTA size for batches that cannot fill the GPU with kAnsBlock-thread CTAs

struct LaunchCfg {
    bool shared;   // one model for the whole batch, tables staged in shared memory
    bool contig;   // contiguous layout (else interleaved deal)
    bool persym;   // a model index per symbol
    bool f64;      // ANS encoder: FP64 quotient estimate
    bool pool;     // decoders: the whole model set is staged in shared memory (p.model.pool_*_bytes)
    bool gauss;    // decoders: per-symbol Gaussian parameters instead of tables (p.gauss_*; implies persym)
    unsigned grid, block;
    size_t smem;   // dynamic shared memory per CTA
    cudaStream_t stream;
};

// Each returns the CUDA error of the attribute call / launch (cudaSuccess if the kernel was enqueued).
cudaError_t launch_ans_encode(const LaunchCfg &cfg, const AnsParams &p);
cudaError_t launch_ans_decode(const LaunchCfg &cfg, const AnsParams &p);
cudaError_t launch_range_encode(const LaunchCfg &cfg, const AnsParams &p);
cudaError_t launch_range_decode(const LaunchCfg &cfg, const AnsParams &p);
// few long streams (contiguous layout): one coder warp + producer warps per CTA (chain_kernels.cuh); cfg.grid = ceil(K / 32)
cudaError_t launch_encode_chain(const LaunchCfg &cfg, const AnsParams &p, bool ans);
// decoders: cfg.shared (one model, quantile index in shared memory) or cfg.pool (model set in shared memory); 64 threads
cudaError_t launch_decode_chain(const LaunchCfg &cfg, const AnsParams &p, bool range);
constexpr int kChainCtaThreads = 128;
constexpr int kChainRingSlots = 4;

// shared by the translation units: opt in to > 48 KB of dynamic shared memory, launch, report
template <typename Kernel>
cudaError_t launch_kernel(Kernel kernel, const LaunchCfg &cfg, const AnsParams &p) {
    if (cfg.smem > 48 * 1024) {
        const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cfg.smem);
        if (e != cudaSuccess) return e;
    }
    kernel<<<cfg.grid, cfg.block, cfg.smem, cfg.stream>>>(p);
    return cudaGetLastError();
}

// SHARED implies one model for the whole batch, hence no per-symbol index: six (SHARED, CONTIG, PERSYM) cases.
#define CTR_LAYOUT_DISPATCH(KERNEL, BLOCK, ...)                                                                  \
    do {                                                                                                         \
        if (cfg.shared)                                                                                          \
            return cfg.contig ? launch_kernel(KERNEL<BLOCK, true, true, false __VA_ARGS__>, cfg, p)              \
                              : launch_kernel(KERNEL<BLOCK, true, false, false __VA_ARGS__>, cfg, p);            \
        if (cfg.persym)                                                                                          \
            return cfg.contig ? launch_kernel(KERNEL<BLOCK, false, true, true __VA_ARGS__>, cfg, p)              \
                              : launch_kernel(KERNEL<BLOCK, false, false, true __VA_ARGS__>, cfg, p);            \
        return cfg.contig ? launch_kernel(KERNEL<BLOCK, false, true, false __VA_ARGS__>, cfg, p)                 \
                          : launch_kernel(KERNEL<BLOCK, false, false, false __VA_ARGS__>, cfg, p);               \
    } while (0)
#define CTR_COMMA ,

}  // namespace ctr
