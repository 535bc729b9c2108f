// xquant.cuh — the block-scaled int8x3 activation quantiser (see kernels_internal.h "xq"), shared by the
// stand-alone quantise/rmsnorm kernels (global layout) and the GEMV's fused prologue (swizzled shared layout).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace nt { namespace b200 {

// Quantises the 32 values held one per lane. Returns the three int8 terms of this lane; *scale_out is
// absmax/127/16384 (valid in all lanes), *sum16_out the exact F32 sum of this lane's 16-element half.
__device__ __forceinline__ void quantize_lane32(float v, int& q1, int& q2, int& q3, float& scale_out, float& sum16_out) {
    float amax = fabsf(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xFFFFFFFFu, amax, o));
    const float s = __fdiv_rn(amax, 127.0f);
    const float t = (amax > 0.f) ? __fdiv_rn(v, s) : 0.f;
    const float f1 = rintf(t);
    const float r1 = __fmul_rn(__fsub_rn(t, f1), 128.0f);
    const float f2 = rintf(r1);
    const float r2 = __fmul_rn(__fsub_rn(r1, f2), 128.0f);
    const float f3 = rintf(r2);
    q1 = (int)f1; q2 = (int)f2; q3 = (int)f3;
    float sh = v;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) sh = __fadd_rn(sh, __shfl_xor_sync(0xFFFFFFFFu, sh, o));
    scale_out = __fmul_rn(s, 1.0f / 16384.0f);
    sum16_out = sh;
}

// XOR swizzle of a byte offset inside an x plane in shared memory (16-byte columns ^ half-block index).
__device__ __forceinline__ uint32_t xq_swizzle(uint32_t e) { return e ^ (((e >> 7) & 7u) << 4); }

// Quantise the 32-element block held one element per lane into the global xq layout (kernels_internal.h "xq").
__device__ __forceinline__ void quantize_block32(float v, int blk, int lane, int8_t* xq, int K) {
    int q1, q2, q3;
    float sc, s16;
    quantize_lane32(v, q1, q2, q3, sc, s16);
    // planes are stored in the GEMV's shared-memory order (16-byte columns XOR-swizzled inside each 128-byte line)
    // so the consumer stages them with one TMA bulk copy
    const int e = (int)xq_swizzle((uint32_t)(blk * 32 + lane));
    xq[e] = (int8_t)q1;
    xq[K + e] = (int8_t)q2;
    xq[2 * K + e] = (int8_t)q3;
    float* scale = reinterpret_cast<float*>(xq + 3 * (size_t)K);
    float* sum16 = scale + K / 32;
    if (lane == 0) scale[blk] = sc;
    if ((lane & 15) == 0) sum16[blk * 2 + (lane >> 4)] = s16;
}

}}  // namespace nt::b200
