// TEST HARNESS (CPU): the SIMT-only kernels of the product (loss_ops.cu, gmm_ops.cu, vote_ops.cu, geometry_ops.cu, dataloader_ops.cu, pointnet2_ops.cu) -- kernels AND their C-ABI launchers,
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

extern "C" unsigned long long emu_launch_count() { return emu_launches; }
