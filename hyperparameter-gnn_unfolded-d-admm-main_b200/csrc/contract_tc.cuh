// contract_tc.cuh -- tcgen05 (5th-gen tensor core) contraction, fp32-accurate via 3xTF32.
// Placeholder until the kernel lands: reports every shape as unsupported so AUTO uses SIMT.
#pragma once
#include "common.cuh"
namespace dadmm { namespace tc {
inline bool dims_supported(int, int, int, int) { return false; }
inline bool shape_supported(int, int, int, int, const void*, int64_t, int64_t, int64_t, const void*, int64_t, int64_t,
                            int64_t, const void*, int64_t, int64_t, int64_t) { return false; }
inline size_t workspace_bytes(int, int, int, int) { return 0; }
inline int launch(int, int, int, int, const float*, const float*, float*, int64_t, int64_t, int, void*, size_t, cudaStream_t) {
    DADMM_FAIL(-4, "tcgen05 contraction not built");
}
}}
