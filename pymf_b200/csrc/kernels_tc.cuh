// kernels_tc.cuh - tcgen05 / TMA kernels (3xTF32) of the streaming passes.  [stub: filled in next]
#pragma once
#include <string>
#include "common.cuh"
namespace pymfb {
struct TcPlan { bool ready = false; };
inline bool tc_supported(int64_t, int64_t, int, int64_t, const float*, std::string* why) { *why = "tcgen05 kernels not built yet"; return false; }
inline int tc_plan(TcPlan&, int, int, int64_t, int64_t, int, int, const float*, int64_t, int64_t) { return 1; }
inline void tc_release(TcPlan&) {}
inline int tc_after_gram(TcPlan&, const DevState*, const float*, const float*, cudaStream_t, int64_t*) { return 1; }
inline int tc_h_update(TcPlan&, const DevState*, const float*, int64_t, const float*, float*, cudaStream_t, int64_t*) { return 1; }
inline int tc_xht(TcPlan&, const DevState*, const float*, int64_t, const float*, float*, cudaStream_t, int64_t*) { return 1; }
}  // namespace pymfb
