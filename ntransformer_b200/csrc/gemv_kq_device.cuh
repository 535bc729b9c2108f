// gemv_kq_device.cuh — per-lane data path of the K-quant / Q8_0 GEMV (see gemv_kquant.cu for the design notes): stage layout
// constants, the lane's activation registers and the dp4a inner loops for one ring stage.  Shared by the stand-alone GEMV
// kernel (gemv_kquant.cu) and the persistent decode-step kernel (engine/decode_megakernel.cu).
#pragma once
#include "kernels_internal.h"
#include "ring.cuh"
#include <cuda_fp16.h>

namespace nt { namespace b200 {
namespace {

constexpr int RG = 4;            // rows per row-group (one stage = RG rows of one chunk)
constexpr int BS = 16;           // super-blocks per chunk (two lanes per super-block)

template <int FMT> struct Fmt;
template <> struct Fmt<0> { static constexpr int BLK = 144; };   // Q4_K
template <> struct Fmt<1> { static constexpr int BLK = 176; };   // Q5_K
template <> struct Fmt<2> { static constexpr int BLK = 210; };   // Q6_K
template <> struct Fmt<3> { static constexpr int BLK = 272; };   // Q8_0: 8 blocks of 34 B = 256 weights
template <> struct Fmt<4> { static constexpr int BLK = 144; };   // Q4_0: 8 blocks of 18 B = 256 weights

// The lane's 128 activation elements: three int8 planes (32 words each, natural order), the four 32-block scales and
// the eight 16-element sums.
struct XRegs {
    int x[3][32];
    float sx[4];
    float s16[8];
};

__device__ __forceinline__ float h2f(uint32_t h16) { return __half2float(__ushort_as_half((unsigned short)h16)); }
__device__ __forceinline__ int combine3(int s0, int s1, int s2) { return (s0 * 128 + s1) * 128 + s2; }

// float(byte IDX of w) without I2F: PRMT the byte under the exponent of 2^23, then subtract 2^23 (exact for 0..255).
__device__ __forceinline__ float byte_to_float(uint32_t w, int idx) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440u + idx)) - 8388608.0f;
}

// One stage (RG rows x up to BS super-blocks) of format FMT: this lane's half super-block (blk, h) against X.
// NR (default RG): rows of the stage that hold data — the persistent kernel cuts the last, partly filled round of a phase
// into 1- or 2-row stages so that its tail costs a fraction of a round (engine/decode_megakernel.cu).
template <int FMT, int NR = RG>
__device__ __forceinline__ void process_stage(const uint8_t* __restrict__ slot_base, int blk, int h, const XRegs& X,
                                              float (&acc)[RG]) {
    constexpr int BLK = Fmt<FMT>::BLK;
    constexpr int ROWP = BS * BLK;            // row pitch inside a stage slot
    const uint8_t* base = slot_base + blk * BLK;
    if (FMT <= 1) {
        // ---------------- Q4_K / Q5_K ----------------
        constexpr int QS = (FMT == 0) ? 16 : 48;
#pragma unroll
        for (int r = 0; r < NR; r++) {
            const int4 hd = *reinterpret_cast<const int4*>(base + r * ROWP);
            const uint32_t w0 = hd.y, w1 = hd.z, w2 = hd.w;
            // packed 6-bit scales / mins of the 4 sub-blocks of this half
            const uint32_t sa = w0 & 0x3F3F3F3Fu, ma = w1 & 0x3F3F3F3Fu;
            const uint32_t sb = (w2 & 0x0F0F0F0Fu) | ((w0 >> 2) & 0x30303030u);
            const uint32_t mb = ((w2 >> 4) & 0x0F0F0F0Fu) | ((w1 >> 2) & 0x30303030u);
            const uint32_t s4 = h ? sb : sa, m4 = h ? mb : ma;
            int4 qv[4];
#pragma unroll
            for (int i = 0; i < 4; i++) qv[i] = *reinterpret_cast<const int4*>(base + r * ROWP + QS + h * 64 + i * 16);
            int4 ha = {0, 0, 0, 0}, hv = {0, 0, 0, 0};
            if (FMT == 1) {
                ha = *reinterpret_cast<const int4*>(base + r * ROWP + 16);
                hv = *reinterpret_cast<const int4*>(base + r * ROWP + 32);
            }
            const uint32_t qh[8] = {(uint32_t)ha.x, (uint32_t)ha.y, (uint32_t)ha.z, (uint32_t)ha.w,
                                    (uint32_t)hv.x, (uint32_t)hv.y, (uint32_t)hv.z, (uint32_t)hv.w};
            float A = 0.f, B = 0.f;
#pragma unroll
            for (int c2 = 0; c2 < 2; c2++) {                      // 64-weight chunk: low nibbles then high nibbles
                const uint32_t q[8] = {(uint32_t)qv[2 * c2].x, (uint32_t)qv[2 * c2].y, (uint32_t)qv[2 * c2].z, (uint32_t)qv[2 * c2].w,
                                       (uint32_t)qv[2 * c2 + 1].x, (uint32_t)qv[2 * c2 + 1].y, (uint32_t)qv[2 * c2 + 1].z,
                                       (uint32_t)qv[2 * c2 + 1].w};
                int l0 = 0, l1 = 0, l2 = 0, h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
                for (int w = 0; w < 8; w++) {
                    uint32_t lo = q[w] & 0x0F0F0F0Fu;
                    // Q4_K: keep the high nibbles in place (codes x16, u8 <= 240) and divide the exact integer sums by 16
                    // once per sub-block instead of shifting every word; Q5_K needs the shift (5-bit codes x16 overflow u8)
                    uint32_t hi = (FMT == 0) ? (q[w] & 0xF0F0F0F0u) : ((q[w] >> 4) & 0x0F0F0F0Fu);
                    if (FMT == 1) {
                        const uint32_t t = qh[w] >> (2 * (2 * h + c2));   // bit0 -> low sub-block, bit1 -> high sub-block
                        lo |= (t << 4) & 0x10101010u;
                        hi |= (t << 3) & 0x10101010u;
                    }
                    const int xl = c2 * 16 + w, xh = c2 * 16 + 8 + w;
                    l0 = dp4a_us(lo, X.x[0][xl], l0); l1 = dp4a_us(lo, X.x[1][xl], l1); l2 = dp4a_us(lo, X.x[2][xl], l2);
                    h0 = dp4a_us(hi, X.x[0][xh], h0); h1 = dp4a_us(hi, X.x[1][xh], h1); h2 = dp4a_us(hi, X.x[2][xh], h2);
                }
                const float flo = (float)combine3(l0, l1, l2) * X.sx[2 * c2];
                const int ihi = (FMT == 0) ? (((h0 * 128 + h1) >> 4) * 128 + (h2 >> 4)) : combine3(h0, h1, h2);
                const float fhi = (float)ihi * X.sx[2 * c2 + 1];
                // 6-bit scale/min byte -> float on the ALU + FMA pipes (PRMT into a 2^23 mantissa, subtract 2^23): keeps
                // the conversion unit, which the dp4a stream already saturates, out of the scale path
                A = fmaf(byte_to_float(s4, 2 * c2), flo, fmaf(byte_to_float(s4, 2 * c2 + 1), fhi, A));
                B = fmaf(byte_to_float(m4, 2 * c2), X.s16[4 * c2] + X.s16[4 * c2 + 1],
                         fmaf(byte_to_float(m4, 2 * c2 + 1), X.s16[4 * c2 + 2] + X.s16[4 * c2 + 3], B));
            }
            acc[r] += h2f((uint32_t)hd.x & 0xFFFFu) * A - h2f((uint32_t)hd.x >> 16) * B;
        }
    } else if (FMT == 3) {
        // ---------------- Q8_0: this lane's 128 weights = 4 blocks of [fp16 d][32 x int8] = 136 bytes (8-byte aligned).
        // Blocks 0 and 2 start 2 bytes into a word: realign with PRMT; blocks 1 and 3 are word aligned. ----------------
        const uint8_t* lb = base + h * 136;
#pragma unroll
        for (int r = 0; r < NR; r++) {
            uint32_t w[34];
#pragma unroll
            for (int i = 0; i < 17; i++) {
                const uint2 v = *reinterpret_cast<const uint2*>(lb + r * ROWP + 8 * i);
                w[2 * i] = v.x; w[2 * i + 1] = v.y;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int w0 = (j == 0) ? 0 : (j == 1) ? 9 : (j == 2) ? 17 : 26;      // first word holding codes
                const bool mis = (j & 1) == 0;
                const float d = h2f((j == 0) ? (w[0] & 0xFFFFu) : (j == 1) ? (w[8] >> 16) : (j == 2) ? (w[17] & 0xFFFFu) : (w[25] >> 16));
                int s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int q = (int)(mis ? __byte_perm(w[w0 + i], w[w0 + i + 1], 0x5432u) : w[w0 + i]);
                    s0 = dp4a_ss(q, X.x[0][8 * j + i], s0);
                    s1 = dp4a_ss(q, X.x[1][8 * j + i], s1);
                    s2 = dp4a_ss(q, X.x[2][8 * j + i], s2);
                }
                // |s0| <= 32*127*127: s0*128+s1 fits s32, the last x128 step is done in F32
                const float f = fmaf((float)(s0 * 128 + s1), 128.0f, (float)s2);
                acc[r] = fmaf(d * X.sx[j], f, acc[r]);
            }
        }
    } else if (FMT == 4) {
        // ---------------- Q4_0 (reference K1, gemm.cu:32-90): this lane's 128 weights = 4 blocks of [fp16 d][16 bytes: low
        // nibbles = weights 0..15, high nibbles = weights 16..31], value = d * (nibble - 8).  72 bytes per lane, 8-byte aligned;
        // blocks 0 and 2 hold their codes 2 bytes into a word (PRMT realign), blocks 1 and 3 are word aligned.  The -8 offset
        // goes through the exact 16-element sums of x, like dmin in Q4_K. ----------------
        const uint8_t* lb = base + h * 72;
#pragma unroll
        for (int r = 0; r < NR; r++) {
            uint32_t w[18];
#pragma unroll
            for (int i = 0; i < 9; i++) {
                const uint2 v = *reinterpret_cast<const uint2*>(lb + r * ROWP + 8 * i);
                w[2 * i] = v.x; w[2 * i + 1] = v.y;
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // block j starts at byte 18 j: word (18 j) / 4 = {0, 4, 9, 13}, byte offset inside it {0, 2, 0, 2}
                const int w0 = (j == 0) ? 0 : (j == 1) ? 4 : (j == 2) ? 9 : 13;
                const bool odd = (j & 1) != 0;               // d in the high half of word w0, codes word aligned from w0 + 1
                const float d = h2f(odd ? (w[w0] >> 16) : (w[w0] & 0xFFFFu));
                int l0 = 0, l1 = 0, l2 = 0, h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t q = odd ? w[w0 + 1 + i] : __byte_perm(w[w0 + i], w[w0 + i + 1], 0x5432u);
                    const uint32_t lo = q & 0x0F0F0F0Fu, hi = (q >> 4) & 0x0F0F0F0Fu;
                    l0 = dp4a_us(lo, X.x[0][8 * j + i], l0); l1 = dp4a_us(lo, X.x[1][8 * j + i], l1); l2 = dp4a_us(lo, X.x[2][8 * j + i], l2);
                    h0 = dp4a_us(hi, X.x[0][8 * j + 4 + i], h0); h1 = dp4a_us(hi, X.x[1][8 * j + 4 + i], h1); h2 = dp4a_us(hi, X.x[2][8 * j + 4 + i], h2);
                }
                const float f = (float)combine3(l0 + h0, l1 + h1, l2 + h2) * X.sx[j];
                acc[r] = fmaf(d, fmaf(-8.0f, X.s16[2 * j] + X.s16[2 * j + 1], f), acc[r]);
            }
        }
    } else {
        // ---------------- Q6_K (210-byte blocks: 2-byte aligned, realigned with PRMT) ----------------
        const uint32_t mis = (uint32_t)(blk & 1) * 2u;           // (blk * 210) & 2
        const uint32_t sel = mis ? 0x5432u : 0x3210u;
        const uint8_t* ab = base - mis;                            // 4-byte aligned view of the block
#pragma unroll
        for (int r = 0; r < NR; r++) {
            const uint32_t* sp = reinterpret_cast<const uint32_t*>(ab + r * ROWP + 192 + 8 * h);
            const uint32_t a0 = sp[0], a1 = sp[1], a2 = sp[2];
            const uint32_t sc0 = __byte_perm(a0, a1, sel), sc1 = __byte_perm(a1, a2, sel);
            const uint32_t dw = *reinterpret_cast<const uint32_t*>(ab + r * ROWP + 208);
            const float d = h2f(mis ? (dw >> 16) : (dw & 0xFFFFu));
            const uint32_t* pq = reinterpret_cast<const uint32_t*>(ab + r * ROWP + 64 * h);          // ql half: 16 words (+1)
            const uint32_t* ph = reinterpret_cast<const uint32_t*>(ab + r * ROWP + 128 + 32 * h);    // qh half: 8 words (+1)
            float A = 0.f;
#pragma unroll
            for (int kk = 0; kk < 2; kk++) {                     // l in [16kk, 16kk + 16)
                uint32_t ra[5], rb[5], rh[5];
#pragma unroll
                for (int i = 0; i < 5; i++) { ra[i] = pq[4 * kk + i]; rb[i] = pq[8 + 4 * kk + i]; rh[i] = ph[4 * kk + i]; }
                int s[4][3] = {};
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t qa = __byte_perm(ra[i], ra[i + 1], sel), qb = __byte_perm(rb[i], rb[i + 1], sel);
                    const uint32_t hh = __byte_perm(rh[i], rh[i + 1], sel);
                    const uint32_t q1 = (qa & 0x0F0F0F0Fu) | ((hh << 4) & 0x30303030u);
                    const uint32_t q2 = (qb & 0x0F0F0F0Fu) | ((hh << 2) & 0x30303030u);
                    const uint32_t q3 = ((qa >> 4) & 0x0F0F0F0Fu) | (hh & 0x30303030u);
                    const uint32_t q4 = ((qb >> 4) & 0x0F0F0F0Fu) | ((hh >> 2) & 0x30303030u);
                    const int xi = 4 * kk + i;                   // word inside a 32-element run
#pragma unroll
                    for (int pl = 0; pl < 3; pl++) {
                        s[0][pl] = dp4a_us(q1, X.x[pl][0 + xi], s[0][pl]);
                        s[1][pl] = dp4a_us(q2, X.x[pl][8 + xi], s[1][pl]);
                        s[2][pl] = dp4a_us(q3, X.x[pl][16 + xi], s[2][pl]);
                        s[3][pl] = dp4a_us(q4, X.x[pl][24 + xi], s[3][pl]);
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int b = 2 * j + kk;                      // scale index inside the half
                    const int sc = (int)(signed char)(((b < 4 ? sc0 : sc1) >> (8 * (b & 3))) & 0xFFu);
                    // sum over 16 weights of sc * (q - 32) * x = sc * (S * sx - 32 * sum16)
                    A = fmaf((float)sc, fmaf((float)combine3(s[j][0], s[j][1], s[j][2]), X.sx[j], -32.0f * X.s16[2 * j + kk]), A);
                }
            }
            acc[r] = fmaf(d, A, acc[r]);
        }
    }
}

__host__ __device__ constexpr int max_blk(int mask) { return (mask & 8) ? 272 : (mask & 4) ? 210 : (mask & 2) ? 176 : 144; }   // bit 16 (Q4_0): 144

// 4-row transpose-reduce over the warp: on return lanes with (lane & 7) == 0 hold row (lane>>4)*2 + ((lane>>3)&1).
__device__ __forceinline__ float reduce4(const float (&acc)[RG], int lane) {
    const bool b4 = lane & 16, b3 = lane & 8;
    const float s0 = b4 ? acc[0] : acc[2], s1 = b4 ? acc[1] : acc[3];
    float k0 = b4 ? acc[2] : acc[0], k1 = b4 ? acc[3] : acc[1];
    k0 += __shfl_xor_sync(0xFFFFFFFFu, s0, 16);
    k1 += __shfl_xor_sync(0xFFFFFFFFu, s1, 16);
    const float sv = b3 ? k0 : k1;
    float kv = b3 ? k1 : k0;
    kv += __shfl_xor_sync(0xFFFFFFFFu, sv, 8);
    kv += __shfl_xor_sync(0xFFFFFFFFu, kv, 4);
    kv += __shfl_xor_sync(0xFFFFFFFFu, kv, 2);
    kv += __shfl_xor_sync(0xFFFFFFFFu, kv, 1);
    return kv;
}

}  // namespace
}}  // namespace nt::b200
