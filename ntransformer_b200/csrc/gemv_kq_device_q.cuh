// gemv_kq_device_q.cuh — quarter-block data path of the K-quant / Q8_0 / Q4_0 GEMV (gemv_kquant_q.cu).
//
// Same arithmetic as gemv_kq_device.cuh (exact IDP.4A partial sums against the three int8 activation planes, one F32 scale per
// 16 / 32 weights) with a finer lane mapping: lane <-> QUARTER super-block (64 weights) instead of half, so a lane keeps 54
// activation registers instead of 108, a warp covers 8 super-blocks (2048 weights) of a row, and the kernel affords 16 warps per
// SM instead of 12 (and 14 on the 28672-wide down projection, where the half-block mapping is stuck at 7 = the number of
// 4096-weight chunks).  ncu of the half-block kernel: issue slots 58 % busy with 3 warps per scheduler, ~4 stall cycles per
// instruction per warp — a fourth warp per scheduler is what the issue rate lacks (profiles/r02_gemv_launches_ncu_summary.txt).
// Which 64 weights of the 256 a lane owns follows each format's own block structure:
//   Q4_K / Q5_K: quarter q = 2 * h + c2  -> sub-blocks {2q, 2q + 1} (the low and the high nibbles of qs[32 q .. 32 q + 32))
//   Q6_K:        h = q >> 1, kk = q & 1  -> l in [16 kk, 16 kk + 16) of half h: 16 weights of each of its four 32-weight runs
//   Q8_0 / Q4_0: 32-weight blocks {2q, 2q + 1} of the 8 that make up 256 weights
#pragma once
#include "gemv_kq_device.cuh"

namespace nt { namespace b200 {
namespace {

constexpr int BSQ = 8;           // super-blocks per chunk in the quarter-block mapping (four lanes per super-block)

// The lane's 64 activation elements: three int8 planes (16 words each), their 32-block scales and four 16-element sums.
//   contiguous formats (everything but Q6_K): elements [64 u, 64 u + 64) of the row, u = quarter index; sx[0..1], s16[0..3] in order;
//   Q6_K: elements 128 hh + 32 j + 16 kk + [0, 16) for j = 0..3 (hh = u >> 1, kk = u & 1): x[pl][4 j + i], sx[j] = scale of the
//   32-block the group lies in (shared with the lane of the other kk), s16[j] = the group's own sum.
struct XRegsQ {
    int x[3][16];
    float sx[4];
    float s16[4];
};
// byte offset (inside a plane, unswizzled) of the 16-byte piece `piece` (0..3) of quarter `u` of the row, and the indices of the
// scale / 16-element sum that go with it
template <bool Q6> __device__ __forceinline__ uint32_t xq_piece_offset(uint32_t u, int piece) {
    return Q6 ? ((u >> 1) * 128u + 32u * piece + 16u * (u & 1u)) : (u * 64u + 16u * piece);
}

// One stage (RG rows x up to BSQ super-blocks) of format FMT: this lane's quarter (blk, q) against X.
template <int FMT, int NR = RG>
__device__ __forceinline__ void process_stage_q(const uint8_t* __restrict__ slot_base, int blk, int q, const XRegsQ& X,
                                                float (&acc)[RG]) {
    constexpr int BLK = Fmt<FMT>::BLK;
    constexpr int ROWP = BSQ * BLK;           // row pitch inside a stage slot
    const uint8_t* base = slot_base + blk * BLK;
    if (FMT <= 1) {
        // ---------------- Q4_K / Q5_K: sub-blocks 2q (low nibbles) and 2q + 1 (high nibbles) of qs[32 q, 32 q + 32) ----------------
        constexpr int QS = (FMT == 0) ? 16 : 48;
        const int h = q >> 1, c2 = q & 1;
#pragma unroll
        for (int r = 0; r < NR; r++) {
            const int4 hd = *reinterpret_cast<const int4*>(base + r * ROWP);
            const uint32_t w0 = hd.y, w1 = hd.z, w2 = hd.w;
            // 6-bit scale / min of sub-blocks j = 2q, 2q + 1: bytes j of (w0, w1) for j < 4, packed across w2 / the top bits for j >= 4
            const uint32_t sa = w0 & 0x3F3F3F3Fu, ma = w1 & 0x3F3F3F3Fu;
            const uint32_t sb = (w2 & 0x0F0F0F0Fu) | ((w0 >> 2) & 0x30303030u);
            const uint32_t mb = ((w2 >> 4) & 0x0F0F0F0Fu) | ((w1 >> 2) & 0x30303030u);
            const uint32_t s4 = (h ? sb : sa) >> (16 * c2), m4 = (h ? mb : ma) >> (16 * c2);     // bytes 0, 1 = this quarter's pair
            const int4 qa = *reinterpret_cast<const int4*>(base + r * ROWP + QS + q * 32);
            const int4 qb = *reinterpret_cast<const int4*>(base + r * ROWP + QS + q * 32 + 16);
            uint32_t qh[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            if (FMT == 1) {
                const int4 ha = *reinterpret_cast<const int4*>(base + r * ROWP + 16);
                const int4 hv = *reinterpret_cast<const int4*>(base + r * ROWP + 32);
                qh[0] = ha.x; qh[1] = ha.y; qh[2] = ha.z; qh[3] = ha.w; qh[4] = hv.x; qh[5] = hv.y; qh[6] = hv.z; qh[7] = hv.w;
            }
            const uint32_t qw[8] = {(uint32_t)qa.x, (uint32_t)qa.y, (uint32_t)qa.z, (uint32_t)qa.w,
                                    (uint32_t)qb.x, (uint32_t)qb.y, (uint32_t)qb.z, (uint32_t)qb.w};
            int l0 = 0, l1 = 0, l2 = 0, h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                uint32_t lo = qw[w] & 0x0F0F0F0Fu;
                uint32_t hi = (FMT == 0) ? (qw[w] & 0xF0F0F0F0u) : ((qw[w] >> 4) & 0x0F0F0F0Fu);
                if (FMT == 1) {
                    const uint32_t t = qh[w] >> (2 * q);              // bit 0 -> sub-block 2q, bit 1 -> sub-block 2q + 1
                    lo |= (t << 4) & 0x10101010u;
                    hi |= (t << 3) & 0x10101010u;
                }
                l0 = dp4a_us(lo, X.x[0][w], l0); l1 = dp4a_us(lo, X.x[1][w], l1); l2 = dp4a_us(lo, X.x[2][w], l2);
                h0 = dp4a_us(hi, X.x[0][8 + w], h0); h1 = dp4a_us(hi, X.x[1][8 + w], h1); h2 = dp4a_us(hi, X.x[2][8 + w], h2);
            }
            const float flo = (float)combine3(l0, l1, l2) * X.sx[0];
            const int ihi = (FMT == 0) ? (((h0 * 128 + h1) >> 4) * 128 + (h2 >> 4)) : combine3(h0, h1, h2);
            const float fhi = (float)ihi * X.sx[1];
            const float A = fmaf(byte_to_float(s4, 0), flo, byte_to_float(s4, 1) * fhi);
            const float B = fmaf(byte_to_float(m4, 0), X.s16[0] + X.s16[1], byte_to_float(m4, 1) * (X.s16[2] + X.s16[3]));
            acc[r] += h2f((uint32_t)hd.x & 0xFFFFu) * A - h2f((uint32_t)hd.x >> 16) * B;
        }
    } else if (FMT == 3) {
        // ---------------- Q8_0: blocks 2q, 2q + 1 of [fp16 d][32 x int8] = 68 bytes at offset 68 q (4-byte aligned).
        // Block 2q holds its codes 2 bytes into a word (PRMT realign), block 2q + 1 is word aligned. ----------------
        const uint8_t* lb = base + q * 68;
#pragma unroll
        for (int r = 0; r < NR; r++) {
            uint32_t w[17];
#pragma unroll
            for (int i = 0; i < 17; i++) w[i] = *reinterpret_cast<const uint32_t*>(lb + r * ROWP + 4 * i);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int w0 = (j == 0) ? 0 : 9;                 // first word holding codes
                const float d = h2f((j == 0) ? (w[0] & 0xFFFFu) : (w[8] >> 16));
                int s0 = 0, s1 = 0, s2 = 0;
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    const int c = (int)((j == 0) ? __byte_perm(w[w0 + i], w[w0 + i + 1], 0x5432u) : w[w0 + i]);
                    s0 = dp4a_ss(c, X.x[0][8 * j + i], s0);
                    s1 = dp4a_ss(c, X.x[1][8 * j + i], s1);
                    s2 = dp4a_ss(c, X.x[2][8 * j + i], s2);
                }
                const float f = fmaf((float)(s0 * 128 + s1), 128.0f, (float)s2);
                acc[r] = fmaf(d * X.sx[j], f, acc[r]);
            }
        }
    } else if (FMT == 4) {
        // ---------------- Q4_0: blocks 2q, 2q + 1 of [fp16 d][16 bytes] = 36 bytes at offset 36 q (4-byte aligned); value =
        // d * (nibble - 8), the -8 through the exact 16-element sums. ----------------
        const uint8_t* lb = base + q * 36;
#pragma unroll
        for (int r = 0; r < NR; r++) {
            uint32_t w[9];
#pragma unroll
            for (int i = 0; i < 9; i++) w[i] = *reinterpret_cast<const uint32_t*>(lb + r * ROWP + 4 * i);
#pragma unroll
            for (int j = 0; j < 2; j++) {
                const int w0 = (j == 0) ? 0 : 4;
                const bool odd = j != 0;                          // d in the high half of word w0, codes word aligned from w0 + 1
                const float d = h2f(odd ? (w[w0] >> 16) : (w[w0] & 0xFFFFu));
                int l0 = 0, l1 = 0, l2 = 0, h0 = 0, h1 = 0, h2 = 0;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const uint32_t c = odd ? w[w0 + 1 + i] : __byte_perm(w[w0 + i], w[w0 + i + 1], 0x5432u);
                    const uint32_t lo = c & 0x0F0F0F0Fu, hi = (c >> 4) & 0x0F0F0F0Fu;
                    l0 = dp4a_us(lo, X.x[0][8 * j + i], l0); l1 = dp4a_us(lo, X.x[1][8 * j + i], l1); l2 = dp4a_us(lo, X.x[2][8 * j + i], l2);
                    h0 = dp4a_us(hi, X.x[0][8 * j + 4 + i], h0); h1 = dp4a_us(hi, X.x[1][8 * j + 4 + i], h1); h2 = dp4a_us(hi, X.x[2][8 * j + 4 + i], h2);
                }
                const float f = (float)combine3(l0 + h0, l1 + h1, l2 + h2) * X.sx[j];
                acc[r] = fmaf(d, fmaf(-8.0f, X.s16[2 * j] + X.s16[2 * j + 1], f), acc[r]);
            }
        }
    } else {
        // ---------------- Q6_K (210-byte blocks, 2-byte aligned): half h = q >> 1, l in [16 kk, 16 kk + 16), kk = q & 1.  The lane's
        // 64 weights are 16 of each of the half's four 32-weight runs j = 0..3; X.x[pl][4 j + i] pairs with code word i of run j. ---
        const int h = q >> 1, kk = q & 1;
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
            const uint32_t* pq = reinterpret_cast<const uint32_t*>(ab + r * ROWP + 64 * h + 16 * kk);          // ql: 4 words (+1) at l
            const uint32_t* ph = reinterpret_cast<const uint32_t*>(ab + r * ROWP + 128 + 32 * h + 16 * kk);    // qh: 4 words (+1)
            uint32_t ra[5], rb[5], rh[5];
#pragma unroll
            for (int i = 0; i < 5; i++) { ra[i] = pq[i]; rb[i] = pq[8 + i]; rh[i] = ph[i]; }
            int s[4][3] = {};
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t qa = __byte_perm(ra[i], ra[i + 1], sel), qb = __byte_perm(rb[i], rb[i + 1], sel);
                const uint32_t hh = __byte_perm(rh[i], rh[i + 1], sel);
                const uint32_t q1 = (qa & 0x0F0F0F0Fu) | ((hh << 4) & 0x30303030u);
                const uint32_t q2 = (qb & 0x0F0F0F0Fu) | ((hh << 2) & 0x30303030u);
                const uint32_t q3 = ((qa >> 4) & 0x0F0F0F0Fu) | (hh & 0x30303030u);
                const uint32_t q4 = ((qb >> 4) & 0x0F0F0F0Fu) | ((hh >> 2) & 0x30303030u);
#pragma unroll
                for (int pl = 0; pl < 3; pl++) {
                    s[0][pl] = dp4a_us(q1, X.x[pl][0 + i], s[0][pl]);
                    s[1][pl] = dp4a_us(q2, X.x[pl][4 + i], s[1][pl]);
                    s[2][pl] = dp4a_us(q3, X.x[pl][8 + i], s[2][pl]);
                    s[3][pl] = dp4a_us(q4, X.x[pl][12 + i], s[3][pl]);
                }
            }
            float A = 0.f;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                // scale of the 16-weight group (run j, l half kk): index 2 j + kk inside the half's 8 scales
                const uint32_t word = (j < 2) ? sc0 : sc1;
                const int sc = (int)(signed char)((word >> (8 * (2 * (j & 1) + kk))) & 0xFFu);
                // sum over 16 weights of sc * (q - 32) * x = sc * (S * sx - 32 * sum16)
                A = fmaf((float)sc, fmaf((float)combine3(s[j][0], s[j][1], s[j][2]), X.sx[j], -32.0f * X.s16[j]), A);
            }
            acc[r] = fmaf(d, A, acc[r]);
        }
    }
}

}  // namespace
}}  // namespace nt::b200
