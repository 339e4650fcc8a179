// host_common.h -- error reporting and launch accounting shared by the host translation units of the library
// (capi.cu owns the thread-local error text behind ctr_last_cuda_error()).
#pragma once
#include <cuda_runtime.h>

#include <string>

namespace ctr {

int host_cuda_fail(cudaError_t e, const char *what);  // records the text, returns CTR_ERR_CUDA
int host_fail(const std::string &what);               // same for driver-API / library failures
void host_count_launch();                             // ctr_kernel_launch_count()
void host_keep_pool_memory();                         // the stream-ordered pool keeps freed memory (no re-mapping per call)

}  // namespace ctr

#define CTR_HOST_TRY(expr)                                             \
    do {                                                               \
        cudaError_t e__ = (expr);                                      \
        if (e__ != cudaSuccess) return ctr::host_cuda_fail(e__, #expr); \
    } while (0)
