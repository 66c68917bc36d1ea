// stream_bn.cuh -- internal interface of the bulk-TMA streaming BatchNorm kernels (stream_bn.cu); the C-ABI entry
// points in dense_ops.cu route [M, 64] bf16 operands here and everything else to their generic kernels.
#pragma once
#include "p2r_common.cuh"   // cuda_runtime.h, or the host emulator's twin of it in the CPU tests

bool p2r_stream_bn_ok(int dtype, long long M, int C, const void* p0, const void* p1 = nullptr, const void* p2 = nullptr,
                      const void* p3 = nullptr, const void* p4 = nullptr);
int p2r_stream_col_stats(const void* x, long long M, double* s1, double* s2, cudaStream_t st);
int p2r_stream_colsum_period(const void* x, long long rows, int period, double* out, cudaStream_t st);
int p2r_stream_col_bwd_stats(const void* dy, const void* x, const void* y, long long M, const float* mean,
                             const float* rstd, int relu, double* s1, double* s2, const float* scale,
                             const float* shift, cudaStream_t st);
int p2r_stream_affine_act(const void* x, long long M, const float* scale, const float* shift, const void* residual,
                          int relu, void* y, unsigned char* relu_mask, cudaStream_t st);
int p2r_stream_bn_bwd_apply(const void* dy, const void* x, const void* y, long long M, const float* mean,
                            const float* rstd, const float* scale, const double* s1, const double* s2, int relu,
                            void* dx, void* dres, const float* shift, double* colsum, int period, cudaStream_t st,
                            const double* sums64 = nullptr, float* sums32 = nullptr);
