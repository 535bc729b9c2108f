// gemv_generic.cu — GEMV for every dtype and any (2-byte) alignment with F32 activations.
//
// Covers reference kernels K1 (Q4_0), K2 (Q8_0), K6 (F16), K7 (F32) — src/cuda/gemm.cu:32-152, 546-671 —
// and is the fallback for K-quant shapes the TMA kernel (gemv_kquant.cu) cannot take (rows whose pitch is
// not a multiple of 16 bytes).  One warp per pair of rows, lanes stride over quantisation blocks, x is
// read through L1 (ld.global.nc); 18/34/210-byte blocks are only 2-byte aligned so codes are fetched as
// 16-bit words.  F32 accumulate, warp-shuffle reduction.
#include "kernels_internal.h"
#include "ring.cuh"
#include <cuda_fp16.h>
#include <algorithm>

namespace nt { namespace b200 {

namespace {

constexpr int GW = 8;        // warps per CTA
constexpr int RPW = 2;       // rows per warp

__device__ __forceinline__ float h2f16(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ uint16_t ld16(const uint8_t* p) { return __ldg(reinterpret_cast<const uint16_t*>(p)); }

// 6-bit (scale, min) pair j of a Q4_K/Q5_K super-block
__device__ __forceinline__ void k4_scale_min(const uint8_t* s, int j, float& sc, float& mn) {
    if (j < 4) {
        sc = (float)(s[j] & 63);
        mn = (float)(s[j + 4] & 63);
    } else {
        sc = (float)((s[j + 4] & 0x0F) | ((s[j - 4] >> 6) << 4));
        mn = (float)((s[j + 4] >> 4) | ((s[j] >> 6) << 4));
    }
}

template <int DT>
__device__ float row_dot(const uint8_t* __restrict__ row, const float* __restrict__ x, int in, int lane) {
    float sum = 0.f;
    if (DT == (int)DType::F32) {
        const float* w = reinterpret_cast<const float*>(row);
        for (int i = lane; i < in; i += 32) sum = fmaf(__ldg(w + i), __ldg(x + i), sum);
    } else if (DT == (int)DType::F16) {
        const bool vec = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
        int i0 = 0;
        if (vec) {
            const int n8 = in / 8;
            for (int v = lane; v < n8; v += 32) {
                uint4 wv = __ldg(reinterpret_cast<const uint4*>(row) + v);
                float4 xa = __ldg(reinterpret_cast<const float4*>(x) + 2 * v);
                float4 xb = __ldg(reinterpret_cast<const float4*>(x) + 2 * v + 1);
                float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&wv.x));
                float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&wv.y));
                float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&wv.z));
                float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&wv.w));
                sum = fmaf(f0.x, xa.x, sum); sum = fmaf(f0.y, xa.y, sum);
                sum = fmaf(f1.x, xa.z, sum); sum = fmaf(f1.y, xa.w, sum);
                sum = fmaf(f2.x, xb.x, sum); sum = fmaf(f2.y, xb.y, sum);
                sum = fmaf(f3.x, xb.z, sum); sum = fmaf(f3.y, xb.w, sum);
            }
            i0 = n8 * 8;
        }
        for (int i = i0 + lane; i < in; i += 32) sum = fmaf(h2f16(ld16(row + 2 * (size_t)i)), __ldg(x + i), sum);
    } else if (DT == (int)DType::Q8_0) {
        const int nb = in / 32;
        for (int b = lane; b < nb; b += 32) {
            const uint8_t* blk = row + (size_t)b * 34;
            const float* xb = x + b * 32;
            float d = h2f16(ld16(blk));
            float bs = 0.f;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                uint16_t w = ld16(blk + 2 + 2 * j);
                bs = fmaf((float)(int8_t)(w & 0xFF), __ldg(xb + 2 * j), bs);
                bs = fmaf((float)(int8_t)(w >> 8), __ldg(xb + 2 * j + 1), bs);
            }
            sum = fmaf(d, bs, sum);
        }
    } else if (DT == (int)DType::Q4_0) {
        const int nb = in / 32;
        for (int b = lane; b < nb; b += 32) {
            const uint8_t* blk = row + (size_t)b * 18;
            const float* xb = x + b * 32;
            float d = h2f16(ld16(blk));
            float bs = 0.f;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                uint16_t w = ld16(blk + 2 + 2 * j);
                int b0 = w & 0xFF, b1 = w >> 8;
                bs = fmaf((float)((b0 & 15) - 8), __ldg(xb + 2 * j), bs);
                bs = fmaf((float)((b0 >> 4) - 8), __ldg(xb + 2 * j + 16), bs);
                bs = fmaf((float)((b1 & 15) - 8), __ldg(xb + 2 * j + 1), bs);
                bs = fmaf((float)((b1 >> 4) - 8), __ldg(xb + 2 * j + 17), bs);
            }
            sum = fmaf(d, bs, sum);
        }
    } else if (DT == (int)DType::Q4_K_M || DT == (int)DType::Q5_K) {
        constexpr bool Q5 = (DT == (int)DType::Q5_K);
        constexpr int BLK = Q5 ? 176 : 144;
        // lanes stride over (super-block, 64-weight chunk) units
        const int nunits = (in / 256) * 4;
        for (int u = lane; u < nunits; u += 32) {
            const int b = u >> 2, c = u & 3;
            const uint8_t* blk = row + (size_t)b * BLK;
            const float* xb = x + b * 256 + c * 64;
            float d = h2f16(ld16(blk)), dmin = h2f16(ld16(blk + 2));
            uint8_t sc[12];
#pragma unroll
            for (int j = 0; j < 6; j++) { uint16_t w = ld16(blk + 4 + 2 * j); sc[2 * j] = w & 0xFF; sc[2 * j + 1] = w >> 8; }
            float s_lo, m_lo, s_hi, m_hi;
            k4_scale_min(sc, 2 * c, s_lo, m_lo);
            k4_scale_min(sc, 2 * c + 1, s_hi, m_hi);
            const uint8_t* ql = blk + (Q5 ? 48 : 16) + c * 32;
            const uint8_t* qh = blk + 16;
            float a_lo = 0.f, a_hi = 0.f, x_lo = 0.f, x_hi = 0.f;
#pragma unroll
            for (int j = 0; j < 16; j++) {
                uint16_t w = ld16(ql + 2 * j);
                int b0 = w & 0xFF, b1 = w >> 8;
                int l0 = b0 & 15, h0 = b0 >> 4, l1 = b1 & 15, h1 = b1 >> 4;
                if (Q5) {
                    uint16_t hw = ld16(qh + 2 * j);
                    int hb0 = (hw & 0xFF) >> (2 * c), hb1 = (hw >> 8) >> (2 * c);
                    l0 += (hb0 & 1) << 4; h0 += (hb0 & 2) << 3;
                    l1 += (hb1 & 1) << 4; h1 += (hb1 & 2) << 3;
                }
                float xa = __ldg(xb + 2 * j), xb1 = __ldg(xb + 2 * j + 1);
                float xc = __ldg(xb + 32 + 2 * j), xd = __ldg(xb + 33 + 2 * j);
                a_lo = fmaf((float)l0, xa, a_lo); a_lo = fmaf((float)l1, xb1, a_lo);
                a_hi = fmaf((float)h0, xc, a_hi); a_hi = fmaf((float)h1, xd, a_hi);
                x_lo += xa + xb1; x_hi += xc + xd;
            }
            sum += d * (s_lo * a_lo + s_hi * a_hi) - dmin * (m_lo * x_lo + m_hi * x_hi);
        }
    } else if (DT == (int)DType::Q6_K) {
        // lanes stride over (super-block, 128-weight half) units
        const int nunits = (in / 256) * 2;
        for (int u = lane; u < nunits; u += 32) {
            const int b = u >> 1, hf = u & 1;
            const uint8_t* blk = row + (size_t)b * 210;
            const float* xb = x + b * 256 + hf * 128;
            const uint8_t* ql = blk + 64 * hf;
            const uint8_t* qh = blk + 128 + 32 * hf;
            float d = h2f16(ld16(blk + 208));
            float scf[8];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint16_t w = ld16(blk + 192 + 8 * hf + 2 * j);
                scf[2 * j] = (float)(int8_t)(w & 0xFF);
                scf[2 * j + 1] = (float)(int8_t)(w >> 8);
            }
            float a = 0.f;
#pragma unroll
            for (int j = 0; j < 16; j++) {        // l = 2j, 2j+1
                uint16_t wa = ld16(ql + 2 * j), wb = ld16(ql + 32 + 2 * j), wh = ld16(qh + 2 * j);
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    int l = 2 * j + t;
                    int qa = (wa >> (8 * t)) & 0xFF, qb = (wb >> (8 * t)) & 0xFF, hh = (wh >> (8 * t)) & 0xFF;
                    int q1 = ((qa & 15) | ((hh & 3) << 4)) - 32;
                    int q2 = ((qb & 15) | (((hh >> 2) & 3) << 4)) - 32;
                    int q3 = ((qa >> 4) | (((hh >> 4) & 3) << 4)) - 32;
                    int q4 = ((qb >> 4) | (((hh >> 6) & 3) << 4)) - 32;
                    int is = l >> 4;
                    a = fmaf(scf[is] * (float)q1, __ldg(xb + l), a);
                    a = fmaf(scf[is + 2] * (float)q2, __ldg(xb + l + 32), a);
                    a = fmaf(scf[is + 4] * (float)q3, __ldg(xb + l + 64), a);
                    a = fmaf(scf[is + 6] * (float)q4, __ldg(xb + l + 96), a);
                }
            }
            sum = fmaf(d, a, sum);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
    return sum;
}

template <int DT>
__global__ void __launch_bounds__(GW * 32) gemv_generic_kernel(float* __restrict__ y, const uint8_t* __restrict__ W,
                                                               const float* __restrict__ x, int out, int in,
                                                               size_t pitch, int ep) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * GW + warp) * RPW;
    pdl_launch_dependents();
    pdl_wait();
#pragma unroll
    for (int r = 0; r < RPW; r++) {
        int row = row0 + r;
        if (row >= out) return;
        float v = row_dot<DT>(W + (size_t)row * pitch, x, in, lane);
        if (lane == 0) { if (ep == GEMV_ADD) y[row] += v; else y[row] = v; }
    }
}

// ---- F16 weights, 16-byte aligned rows (reference K6, gemm.cu:546-610): x staged once per CTA in shared memory, a warp
// streams two rows at a time with four 16-byte loads per row in flight per lane (ld.global.nc, no L1 allocation). ----
constexpr int F16_WARPS = 8;
__device__ __forceinline__ uint4 ldg_stream16(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float dot8(const uint4& w, const float4& xa, const float4& xb, float acc) {
    const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&w.x));
    const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&w.y));
    const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&w.z));
    const float2 f3 = __half22float2(*reinterpret_cast<const __half2*>(&w.w));
    acc = fmaf(f0.x, xa.x, acc); acc = fmaf(f0.y, xa.y, acc); acc = fmaf(f1.x, xa.z, acc); acc = fmaf(f1.y, xa.w, acc);
    acc = fmaf(f2.x, xb.x, acc); acc = fmaf(f2.y, xb.y, acc); acc = fmaf(f3.x, xb.z, acc); acc = fmaf(f3.y, xb.w, acc);
    return acc;
}
__global__ void __launch_bounds__(F16_WARPS * 32) gemv_f16_kernel(float* __restrict__ y, const uint8_t* __restrict__ W,
                                                                  const float* __restrict__ x, int out, int in, size_t pitch, int ep) {
    extern __shared__ __align__(16) float xs[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    pdl_launch_dependents();
    pdl_wait();
    for (int i = threadIdx.x * 4; i < in; i += F16_WARPS * 32 * 4) *reinterpret_cast<float4*>(xs + i) = *reinterpret_cast<const float4*>(x + i);
    __syncthreads();
    const int n8 = in / 8;
    const float4* x4 = reinterpret_cast<const float4*>(xs);
    for (int pair = blockIdx.x * F16_WARPS + warp; 2 * pair < out; pair += gridDim.x * F16_WARPS) {
        const int r0 = 2 * pair, r1 = min(r0 + 1, out - 1);
        const uint4* w0 = reinterpret_cast<const uint4*>(W + (size_t)r0 * pitch);
        const uint4* w1 = reinterpret_cast<const uint4*>(W + (size_t)r1 * pitch);
        float a0 = 0.f, a1 = 0.f;
        int v = lane;
        for (; v + 96 < n8; v += 128) {
            uint4 p[4], q[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { p[u] = ldg_stream16(w0 + v + 32 * u); q[u] = ldg_stream16(w1 + v + 32 * u); }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const float4 xa = x4[2 * (v + 32 * u)], xb = x4[2 * (v + 32 * u) + 1];
                a0 = dot8(p[u], xa, xb, a0);
                a1 = dot8(q[u], xa, xb, a1);
            }
        }
        for (; v < n8; v += 32) {
            const uint4 p = ldg_stream16(w0 + v), q = ldg_stream16(w1 + v);
            const float4 xa = x4[2 * v], xb = x4[2 * v + 1];
            a0 = dot8(p, xa, xb, a0);
            a1 = dot8(q, xa, xb, a1);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { a0 += __shfl_xor_sync(0xFFFFFFFFu, a0, o); a1 += __shfl_xor_sync(0xFFFFFFFFu, a1, o); }
        if (lane == 0) {
            if (ep == GEMV_ADD) { y[r0] += a0; if (r0 + 1 < out) y[r0 + 1] += a1; }
            else { y[r0] = a0; if (r0 + 1 < out) y[r0 + 1] = a1; }
        }
    }
}
constexpr int F16_MAX_SMEM = 200 * 1024;

}  // namespace

void gemv_generic(float* y, const void* W, const float* x, int out, int in, DType dt, size_t row_pitch,
                  GemvEpilogue ep, cudaStream_t s) {
    if (out <= 0) return;
    size_t pitch = row_pitch ? row_pitch : dtype_row_size(dt, (size_t)in);
    const uint8_t* w = static_cast<const uint8_t*>(W);
    if (dt == DType::F16 && in % 8 == 0 && pitch % 16 == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (size_t)in * 4 <= (size_t)F16_MAX_SMEM && ep != GEMV_SWIGLU) {
        static unsigned long long configured = 0;      // bit per device id
        opt_in_dynamic_smem(gemv_f16_kernel, (int)(F16_MAX_SMEM), configured);
        const int pairs = (out + 1) / 2;
        int grid16 = (pairs + F16_WARPS - 1) / F16_WARPS;
        const int resident = (int)std::min<size_t>(8, (size_t)(220 * 1024) / ((size_t)in * 4 + 1024)) * 148;   // CTAs the chip holds
        if (grid16 > resident && resident > 0) grid16 = resident;
        launch_k(gemv_f16_kernel, dim3(grid16), dim3(F16_WARPS * 32), (size_t)in * 4, s, y, w, x, out, in, pitch, (int)ep);
        count_launch();
        return;
    }
    int grid = (out + GW * RPW - 1) / (GW * RPW);
#define NT_LAUNCH(DTV) launch_k(gemv_generic_kernel<(int)DTV>, dim3(grid), dim3(GW * 32), 0, s, y, w, x, out, in, pitch, (int)ep)
    switch (dt) {
        case DType::F32: NT_LAUNCH(DType::F32); break;
        case DType::F16: NT_LAUNCH(DType::F16); break;
        case DType::Q8_0: NT_LAUNCH(DType::Q8_0); break;
        case DType::Q4_0: NT_LAUNCH(DType::Q4_0); break;
        case DType::Q4_K_M: NT_LAUNCH(DType::Q4_K_M); break;
        case DType::Q5_K: NT_LAUNCH(DType::Q5_K); break;
        case DType::Q6_K: NT_LAUNCH(DType::Q6_K); break;
        default:
            // same observable behaviour as the reference launcher (gemm.cu:801-803): report and do nothing
            fprintf(stderr, "Unsupported dtype for GEMV: %s\n", dtype_name(dt));
            return;
    }
#undef NT_LAUNCH
    count_launch();
}

}}  // namespace nt::b200
