// TEST HARNESS (CPU): the SIMT-only kernels of the product (loss_ops.cu, gmm_ops.cu, vote_ops.cu, geometry_ops.cu, dataloader_ops.cu, pointnet2_ops.cu, stream_bn.cu, graph_conv.cu) -- kernels AND their C-ABI launchers,
// unmodified -- compiled for the host on top of the execution-model emulator in cuda_emu.h.  The resulting library
// exports the same p2r_* symbols as libp2r_b200.so for these entry points; tests/test_kernels_emulated.py calls them with
// host arrays.  Built by the test (g++ -DP2R_HOST_EMULATION -ffp-contract=off -pthread); never shipped.
#ifndef P2R_HOST_EMULATION
#error "compile with -DP2R_HOST_EMULATION"
#endif
#include <stdio.h>
#include <string>

static std::string g_last_error;
extern "C" void p2r_set_last_error(const char* where, int code) { g_last_error = std::string(where) + " (" + std::to_string(code) + ")"; }
extern "C" const char* emu_last_error() { return g_last_error.c_str(); }

#include "../../pose2room_b200/csrc/loss_ops.cu"
#include "../../pose2room_b200/csrc/gmm_ops.cu"
#include "../../pose2room_b200/csrc/vote_ops.cu"
#include "../../pose2room_b200/csrc/geometry_ops.cu"
#include "../../pose2room_b200/csrc/dataloader_ops.cu"
#include "../../pose2room_b200/csrc/pointnet2_ops.cu"
#include "../../pose2room_b200/csrc/stream_bn.cu"
#include "../../pose2room_b200/csrc/graph_conv.cu"

// the streaming BatchNorm kernels have an internal C++ interface (stream_bn.cuh): plain-C doors for the test
extern "C" int emu_stream_col_stats(const void* x, long long M, double* s1, double* s2) {
  return p2r_stream_col_stats(x, M, s1, s2, nullptr);
}
extern "C" int emu_stream_col_bwd_stats(const void* dy, const void* x, const void* y, long long M, const float* mean,
                                        const float* rstd, int relu, double* s1, double* s2, const float* scale,
                                        const float* shift) {
  return p2r_stream_col_bwd_stats(dy, x, y, M, mean, rstd, relu, s1, s2, scale, shift, nullptr);
}
extern "C" int emu_stream_affine_act(const void* x, long long M, const float* scale, const float* shift,
                                     const void* residual, int relu, void* y, unsigned char* relu_mask) {
  return p2r_stream_affine_act(x, M, scale, shift, residual, relu, y, relu_mask, nullptr);
}
extern "C" int emu_stream_bn_bwd_apply(const void* dy, const void* x, const void* y, long long M, const float* mean,
                                       const float* rstd, const float* scale, const double* s1, const double* s2, int relu,
                                       void* dx, void* dres, const float* shift, double* colsum, int period) {
  return p2r_stream_bn_bwd_apply(dy, x, y, M, mean, rstd, scale, s1, s2, relu, dx, dres, shift, colsum, period, nullptr);
}

extern "C" int emu_stream_colsum_period(const void* x, long long rows, int period, double* out) {
  return p2r_stream_colsum_period(x, rows, period, out, nullptr);
}

extern "C" unsigned long long emu_launch_count() { return emu_launches; }
