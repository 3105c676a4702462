// Shared by the translation units of libmetis_b200.so (not part of the C ABI).
#ifndef METIS_INTERNAL_H
#define METIS_INTERNAL_H
#include <cuda_runtime.h>

namespace metis {
// record the message returned by metis_last_error() and return METIS_E_CUDA / METIS_E_ARG
int fail_cuda(cudaError_t e, const char *what);
int fail_arg(const char *what);
}  // namespace metis
#endif
