// Shared device helpers for the NUNet-TLS sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nunet {

constexpr float LN_EPS = 1e-8f;   // LayerNormalization(epsilon=1e-8), models/proposed.py:201
constexpr int LSTM_UNITS = 21;    // models/proposed.py:21
constexpr int LSTM_GATES = 84;
constexpr int CTFA_WINDOW = 32;   // models/proposed.py:126 time_seq

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// 16-byte async copy global -> shared; src_bytes == 0 zero-fills the destination (LDGSTS).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gsrc),
                 "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float f4_get(const float4& v, int i) {
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

}  // namespace nunet
