// tcgen05 / TMA tensor-core kernels (throughput mode).  STUB: filled in by the next milestone.
#pragma once
#include "common.cuh"
#include "../../include/fluoro_unet.h"

#define FU_TC_BUILD "0"

namespace fu {

struct TcConv {
  bool enabled = false;
};

inline void tc_carve(TcConv&, int, int, int, bool, bool, Bump&) {}
inline int tc_pack(TcConv&, const float*, cudaStream_t, fu_counters*) { return 0; }
inline const char* tc_last_error() { return "tensor-core path not built"; }
inline bool tc_conv_eligible(const TcConv&, const void*, int, const void*, int, const void*, int) { return false; }
inline int tc_conv_forward(TcConv&, const void*, int, void*, int, int, int, int, const float*, int, double*,
                           const void*, int, const float*, const float*, int, cudaStream_t, fu_counters*) { return -1; }
inline bool tc_down_eligible(const TcConv&, const void*, int, const void*, int) { return false; }
inline int tc_down_forward(TcConv&, const void*, int, void*, int, int, int, int, const float*, cudaStream_t, fu_counters*) { return -1; }
inline bool tc_up_eligible(const TcConv&, const void*, int, const void*, int) { return false; }
inline int tc_up_forward(TcConv&, const void*, int, void*, int, int, int, int, const float*, cudaStream_t, fu_counters*) { return -1; }
inline bool tc_dgrad_eligible(const TcConv&, const void*, int, const void*, int) { return false; }
inline int tc_conv_dgrad(TcConv&, const void*, int, void*, int, int, int, int, int, cudaStream_t, fu_counters*) { return -1; }
inline bool tc_wgrad_eligible(const TcConv&, const void*, int, const void*, int) { return false; }
inline int tc_conv_wgrad(TcConv&, const void*, int, const void*, int, int, int, int, float*, cudaStream_t, fu_counters*) { return -1; }

}  // namespace fu
namespace fu {
inline int tc_test_conv(int, int, int, int, int, int, int, int, int, int, const void*, const float*, const float*,
                        void*, const void*, float*, double*, cudaStream_t, fu_counters*) { return -1; }
}
