// prefill_dequant.cu — GGUF block formats -> F16 hi/lo pair for the tensor-core prefill GEMM (csrc/prefill_gemm.cu).
//
// The reference only ever dequantises inside its per-token GEMV kernels (src/cuda/gemm.cu:32-470) and, for embeddings, on
// the CPU (src/model/transformer.cpp:394-599).  For a prompt of T tokens the batched path instead expands one weight matrix
// at a time into a scratch pair W = W_hi + W_lo (both F16, ~22 mantissa bits together, so the F32 dequantised value
// d*sc*q - dmin*m survives) and runs two tcgen05 GEMMs over it; the expansion costs ~12 B of traffic per weight once per
// prompt chunk, against 2*T MMAs per weight.  Element formulas follow the reference's: Q8_0 gemm.cu:96-152, Q4_0 :32-90,
// Q4_K :158-262, Q5_K :300-350, Q6_K :356-470 (tests compare every format with the CPU checker bit for bit on the host).
#include "kernels_internal.h"
#include <cuda_fp16.h>

namespace nt { namespace b200 {

namespace {

__device__ __forceinline__ float ldh(const uint8_t* p) { return __half2float(__ushort_as_half(*reinterpret_cast<const uint16_t*>(p))); }

// 6-bit (scale, min) pair `is` of a Q4_K / Q5_K super-block (gemm.cu:199-217)
__device__ __forceinline__ void k4_scale_min(const uint8_t* s, int is, int& sc, int& m) {
    if (is < 4) { sc = s[is] & 63; m = s[is + 4] & 63; }
    else { sc = (s[is + 4] & 15) | ((s[is - 4] >> 6) << 4); m = (s[is + 4] >> 4) | ((s[is] >> 6) << 4); }
}

// value of element i of a row stored in GGUF layout `dt`
__device__ float weight_at(const uint8_t* row, int dt, int i) {
    switch (dt) {
        case (int)DType::F32: return reinterpret_cast<const float*>(row)[i];
        case (int)DType::F16: return ldh(row + 2 * (size_t)i);
        case (int)DType::Q8_0: {
            const uint8_t* b = row + (size_t)(i >> 5) * 34;
            return ldh(b) * (float)(int8_t)b[2 + (i & 31)];
        }
        case (int)DType::Q4_0: {
            const uint8_t* b = row + (size_t)(i >> 5) * 18;
            const int j = i & 31;
            const uint8_t byte = b[2 + (j & 15)];
            const int qv = (j < 16) ? (byte & 15) : (byte >> 4);
            return ldh(b) * (float)(qv - 8);
        }
        case (int)DType::Q4_K_M: {
            const uint8_t* b = row + (size_t)(i >> 8) * 144;
            const int n = i & 255, chunk = n >> 6, l = n & 31, hi = (n >> 5) & 1;
            int sc, m;
            k4_scale_min(b + 4, 2 * chunk + hi, sc, m);
            const uint8_t byte = b[16 + chunk * 32 + l];
            const int qv = hi ? (byte >> 4) : (byte & 15);
            return (ldh(b) * sc) * qv - ldh(b + 2) * m;
        }
        case (int)DType::Q5_K: {
            const uint8_t* b = row + (size_t)(i >> 8) * 176;          // d, dmin, scales[12], qh[32], ql[128]
            const int n = i & 255, chunk = n >> 6, l = n & 31, hi = (n >> 5) & 1;
            int sc, m;
            k4_scale_min(b + 4, 2 * chunk + hi, sc, m);
            const uint8_t byte = b[48 + chunk * 32 + l];
            const int bit = (b[16 + l] >> (2 * chunk + hi)) & 1;
            const int qv = (hi ? (byte >> 4) : (byte & 15)) + (bit ? 16 : 0);
            return (ldh(b) * sc) * qv - ldh(b + 2) * m;
        }
        case (int)DType::Q6_K: {
            const uint8_t* b = row + (size_t)(i >> 8) * 210;          // ql[128], qh[64], scales[16], d
            const int n = i & 255, hf = n >> 7, r = n & 127, run = r >> 5, l = r & 31;
            const uint8_t* ql = b + 64 * hf;
            const uint8_t* qh = b + 128 + 32 * hf;
            const int8_t* sc = reinterpret_cast<const int8_t*>(b + 192 + 8 * hf);
            const uint8_t qb = ql[l + ((run & 1) ? 32 : 0)];
            const int lo = (run >= 2) ? (qb >> 4) : (qb & 15);
            const int q = (lo | (((qh[l] >> (2 * run)) & 3) << 4)) - 32;
            return ldh(b + 208) * (float)sc[(l >> 4) + 2 * run] * q;
        }
        default: return 0.f;
    }
}

// one CTA per row; thread t expands element pairs (2t, 2t+1), (2t + 512, ...)
__global__ void __launch_bounds__(256) dequant_split_kernel(__half2* __restrict__ hi, __half2* __restrict__ lo,
                                                            const uint8_t* __restrict__ W, int dt, size_t row_pitch, int K) {
    const uint8_t* row = W + (size_t)blockIdx.x * row_pitch;
    __half2* h = hi + (size_t)blockIdx.x * (K / 2);
    __half2* l = lo + (size_t)blockIdx.x * (K / 2);
    for (int p = threadIdx.x; p < K / 2; p += 256) {
        const float a = weight_at(row, dt, 2 * p), b = weight_at(row, dt, 2 * p + 1);
        const __half2 hv = __floats2half2_rn(a, b);
        const float2 f = __half22float2(hv);
        h[p] = hv;
        l[p] = __floats2half2_rn(a - f.x, b - f.y);
    }
}

}  // namespace

bool dequant_split_supported(DType dt) {
    switch (dt) {
        case DType::F32: case DType::F16: case DType::Q8_0: case DType::Q4_0: case DType::Q4_K_M: case DType::Q5_K: case DType::Q6_K: return true;
        default: return false;
    }
}

// W (rows x cols, GGUF layout dt, row_pitch bytes between rows) -> w_hi, w_lo: dense F16 [rows][cols] with W = hi + lo
void dequant_split(void* w_hi, void* w_lo, const void* W, DType dt, size_t row_pitch, int rows, int cols, cudaStream_t s) {
    if (rows <= 0) return;
    NT_CHECK(cols % 2 == 0 && dequant_split_supported(dt), "dequant_split: unsupported dtype or odd row length");
    const size_t pitch = row_pitch ? row_pitch : dtype_row_size(dt, (size_t)cols);
    dequant_split_kernel<<<rows, 256, 0, s>>>(static_cast<__half2*>(w_hi), static_cast<__half2*>(w_lo), static_cast<const uint8_t*>(W),
                                              (int)dt, pitch, cols);
    count_launch();
}

}}  // namespace nt::b200
