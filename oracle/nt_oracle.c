/*
 * nt_oracle.c — CPU restatement of the reference's resident decode path.
 * TEST INFRASTRUCTURE ONLY (see nt_oracle.h).  Plain C + OpenMP.
 *
 * Citations are file:line under /root/reference.
 */
#include "nt_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ---------------- block layouts: src/core/types.h:96-138 ---------------- */
#pragma pack(push, 1)
typedef struct { uint16_t d; uint8_t qs[16]; } blk_q4_0;                         /* 18  */
typedef struct { uint16_t d; int8_t qs[32]; } blk_q8_0;                          /* 34  */
typedef struct { uint16_t d, dmin; uint8_t scales[12]; uint8_t qs[128]; } blk_q4_k;              /* 144 */
typedef struct { uint16_t d, dmin; uint8_t scales[12]; uint8_t qh[32]; uint8_t ql[128]; } blk_q5_k; /* 176 */
typedef struct { uint8_t ql[128]; uint8_t qh[64]; int8_t scales[16]; uint16_t d; } blk_q6_k;     /* 210 */
#pragma pack(pop)
_Static_assert(sizeof(blk_q4_0) == 18, "q4_0");
_Static_assert(sizeof(blk_q8_0) == 34, "q8_0");
_Static_assert(sizeof(blk_q4_k) == 144, "q4_k");
_Static_assert(sizeof(blk_q5_k) == 176, "q5_k");
_Static_assert(sizeof(blk_q6_k) == 210, "q6_k");

/* src/core/types.h:38-53 */
size_t nto_dtype_size(int dt) {
    switch (dt) {
        case NTO_F32: return 4;
        case NTO_F16: return 2;
        case NTO_Q8_0: return 34;
        case NTO_Q4_0: return 18;
        case NTO_Q4_K: return 144;
        case NTO_Q5_K: return 176;
        case NTO_Q6_K: return 210;
        default: return 0;
    }
}
/* src/core/types.h:55-68 */
size_t nto_dtype_block_size(int dt) {
    switch (dt) {
        case NTO_Q8_0: case NTO_Q4_0: return 32;
        case NTO_Q4_K: case NTO_Q5_K: case NTO_Q6_K: return 256;
        default: return 1;
    }
}
/* src/core/types.h:84-88 */
size_t nto_row_bytes(int dt, int64_t n) {
    size_t bs = nto_dtype_block_size(dt);
    return (size_t)(n / (int64_t)bs) * nto_dtype_size(dt);
}

/* src/model/transformer.cpp:394-417 */
float nto_fp16_to_fp32(uint16_t h) {
    uint32_t sign = (h >> 15) & 1;
    int32_t exp = (h >> 10) & 0x1F;
    uint32_t mant = h & 0x3FF;
    uint32_t f;
    if (exp == 0) {
        if (mant == 0) {
            f = sign << 31;
        } else {
            exp = 1;
            while (!(mant & 0x400)) { mant <<= 1; exp--; }
            mant &= 0x3FF;
            f = (sign << 31) | ((uint32_t)(exp + 127 - 15) << 23) | (mant << 13);
        }
    } else if (exp == 31) {
        f = (sign << 31) | 0x7F800000u | (mant << 13);
    } else {
        f = (sign << 31) | ((uint32_t)(exp + 127 - 15) << 23) | (mant << 13);
    }
    float r;
    memcpy(&r, &f, 4);
    return r;
}

/* Round-to-nearest-even F32 -> F16, the conversion __float2half performs
 * (src/cuda/attention.cu:338-339). */
uint16_t nto_fp32_to_fp16(float fl) {
    uint32_t x;
    memcpy(&x, &fl, 4);
    uint32_t sign = (x >> 16) & 0x8000u;
    uint32_t absx = x & 0x7FFFFFFFu;
    if (absx >= 0x7F800000u) {                         /* inf / nan */
        return (uint16_t)(sign | 0x7C00u | ((absx > 0x7F800000u) ? 0x200u : 0));
    }
    if (absx >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u); /* rounds to inf (>= 65520) */
    if (absx < 0x33000001u) return (uint16_t)sign;               /* < 2^-25 -> 0 (ties-to-even at 2^-25) */
    int32_t e = (int32_t)(absx >> 23) - 127;
    uint32_t m = (absx & 0x7FFFFFu) | 0x800000u;
    uint32_t half;
    if (e < -14) {                                      /* subnormal half */
        int shift = -14 - e + 13;                       /* bits to drop */
        uint32_t q = m >> shift;
        uint32_t rem = m & ((1u << shift) - 1);
        uint32_t halfway = 1u << (shift - 1);
        if (rem > halfway || (rem == halfway && (q & 1))) q++;
        half = q;
    } else {
        uint32_t q = m >> 13;
        uint32_t rem = m & 0x1FFFu;
        if (rem > 0x1000u || (rem == 0x1000u && (q & 1))) q++;
        half = ((uint32_t)(e + 15) << 10) + (q - 0x400u);   /* carry propagates into exponent */
    }
    return (uint16_t)(sign | half);
}

/* 6-bit scale/min unpack shared by Q4_K and Q5_K — src/cuda/gemm.cu:199-217 */
static void k4_scale_min(const uint8_t* s, int is, uint8_t* sc, uint8_t* m) {
    if (is < 4) {
        *sc = s[is] & 0x3F;
        *m = s[is + 4] & 0x3F;
    } else {
        *sc = (uint8_t)((s[is + 4] & 0x0F) | ((s[is - 4] >> 6) << 4));
        *m = (uint8_t)((s[is + 4] >> 4) | ((s[is] >> 6) << 4));
    }
}

void nto_dequant_row(int dt, const void* row, int64_t n, float* out) {
    const uint8_t* raw = (const uint8_t*)row;
    if (dt == NTO_F32) {                                 /* transformer.cpp:428-437 */
        memcpy(out, raw, (size_t)n * 4);
    } else if (dt == NTO_F16) {                          /* transformer.cpp:438-448 */
        const uint16_t* h = (const uint16_t*)raw;
        for (int64_t i = 0; i < n; i++) out[i] = nto_fp16_to_fp32(h[i]);
    } else if (dt == NTO_Q8_0) {                         /* transformer.cpp:449-470 */
        int64_t nb = n / 32;
        for (int64_t b = 0; b < nb; b++) {
            const blk_q8_0* blk = (const blk_q8_0*)(raw + b * sizeof(blk_q8_0));
            float d = nto_fp16_to_fp32(blk->d);
            for (int j = 0; j < 32; j++) out[b * 32 + j] = d * blk->qs[j];
        }
    } else if (dt == NTO_Q4_0) {                         /* transformer.cpp:471-497 */
        int64_t nb = n / 32;
        for (int64_t b = 0; b < nb; b++) {
            const blk_q4_0* blk = (const blk_q4_0*)(raw + b * sizeof(blk_q4_0));
            float d = nto_fp16_to_fp32(blk->d);
            for (int j = 0; j < 16; j++) {
                uint8_t byte = blk->qs[j];
                int8_t lo = (int8_t)((byte & 0x0F) - 8);
                int8_t hi = (int8_t)((byte >> 4) - 8);
                out[b * 32 + j] = d * lo;
                out[b * 32 + j + 16] = d * hi;
            }
        }
    } else if (dt == NTO_Q6_K) {                         /* transformer.cpp:498-538 */
        int64_t nb = n / 256;
        for (int64_t b = 0; b < nb; b++) {
            const blk_q6_k* blk = (const blk_q6_k*)(raw + b * sizeof(blk_q6_k));
            float d = nto_fp16_to_fp32(blk->d);
            float* y = out + b * 256;
            const uint8_t* ql = blk->ql;
            const uint8_t* qh = blk->qh;
            const int8_t* sc = blk->scales;
            for (int half = 0; half < 2; half++) {
                for (int l = 0; l < 32; l++) {
                    int is = l / 16;
                    int q1 = (int)((ql[l] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
                    int q2 = (int)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                    int q3 = (int)((ql[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
                    int q4 = (int)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
                    y[l] = d * (float)sc[is + 0] * q1;
                    y[l + 32] = d * (float)sc[is + 2] * q2;
                    y[l + 64] = d * (float)sc[is + 4] * q3;
                    y[l + 96] = d * (float)sc[is + 6] * q4;
                }
                y += 128; ql += 64; qh += 32; sc += 8;
            }
        }
    } else if (dt == NTO_Q4_K) {                         /* transformer.cpp:539-594 */
        int64_t nb = n / 256;
        for (int64_t b = 0; b < nb; b++) {
            const blk_q4_k* blk = (const blk_q4_k*)(raw + b * sizeof(blk_q4_k));
            float d = nto_fp16_to_fp32(blk->d);
            float dmin = nto_fp16_to_fp32(blk->dmin);
            float* y = out + b * 256;
            const uint8_t* q = blk->qs;
            for (int chunk = 0; chunk < 4; chunk++) {
                uint8_t sc_lo, m_lo, sc_hi, m_hi;
                k4_scale_min(blk->scales, chunk * 2, &sc_lo, &m_lo);
                k4_scale_min(blk->scales, chunk * 2 + 1, &sc_hi, &m_hi);
                float d1 = d * sc_lo, m1 = dmin * m_lo;
                float d2 = d * sc_hi, m2 = dmin * m_hi;
                for (int l = 0; l < 32; l++) {
                    y[chunk * 64 + l] = d1 * (q[l] & 0xF) - m1;
                    y[chunk * 64 + l + 32] = d2 * (q[l] >> 4) - m2;
                }
                q += 32;
            }
        }
    } else if (dt == NTO_Q5_K) {       /* gemm.cu:300-350 ; tools/decompose_gguf.py:318-367 */
        int64_t nb = n / 256;
        for (int64_t b = 0; b < nb; b++) {
            const blk_q5_k* blk = (const blk_q5_k*)(raw + b * sizeof(blk_q5_k));
            float d = nto_fp16_to_fp32(blk->d);
            float dmin = nto_fp16_to_fp32(blk->dmin);
            float* y = out + b * 256;
            uint8_t u1 = 1, u2 = 2;
            for (int chunk = 0; chunk < 4; chunk++) {
                uint8_t sc_lo, m_lo, sc_hi, m_hi;
                k4_scale_min(blk->scales, chunk * 2, &sc_lo, &m_lo);
                k4_scale_min(blk->scales, chunk * 2 + 1, &sc_hi, &m_hi);
                float d1 = d * sc_lo, m1 = dmin * m_lo;
                float d2 = d * sc_hi, m2 = dmin * m_hi;
                const uint8_t* ql = blk->ql + chunk * 32;
                const uint8_t* qh = blk->qh;
                for (int l = 0; l < 32; l++) {
                    int lo = (ql[l] & 0x0F) + ((qh[l] & u1) ? 16 : 0);
                    int hi = (ql[l] >> 4) + ((qh[l] & u2) ? 16 : 0);
                    y[chunk * 64 + l] = d1 * lo - m1;
                    y[chunk * 64 + l + 32] = d2 * hi - m2;
                }
                u1 <<= 2; u2 <<= 2;
            }
        }
    } else {
        fprintf(stderr, "nto_dequant_row: unsupported dtype %d\n", dt);
        memset(out, 0, (size_t)n * 4);
    }
}

/* One row of the GEMV with the reference's per-block formulas. */
static double gemv_row(int dt, const uint8_t* row, const float* x, int in) {
    double sum = 0.0;
    if (dt == NTO_Q4_0) {                                /* gemm.cu:60-75 */
        int nb = in / 32;
        for (int b = 0; b < nb; b++) {
            const blk_q4_0* blk = (const blk_q4_0*)(row + (size_t)b * 18);
            float d = nto_fp16_to_fp32(blk->d);
            double bs = 0.0;
            for (int j = 0; j < 16; j++) {
                uint8_t byte = blk->qs[j];
                int lo = (byte & 0x0F) - 8, hi = (byte >> 4) - 8;
                bs += (double)lo * x[b * 32 + j] + (double)hi * x[b * 32 + j + 16];
            }
            sum += (double)d * bs;
        }
    } else if (dt == NTO_Q8_0) {                         /* gemm.cu:128-140 */
        int nb = in / 32;
        for (int b = 0; b < nb; b++) {
            const blk_q8_0* blk = (const blk_q8_0*)(row + (size_t)b * 34);
            float d = nto_fp16_to_fp32(blk->d);
            double bs = 0.0;
            for (int j = 0; j < 32; j++) bs += (double)blk->qs[j] * x[b * 32 + j];
            sum += (double)d * bs;
        }
    } else if (dt == NTO_Q4_K) {                         /* gemm.cu:186-243 */
        int nb = in / 256;
        for (int b = 0; b < nb; b++) {
            const blk_q4_k* blk = (const blk_q4_k*)(row + (size_t)b * 144);
            float d = nto_fp16_to_fp32(blk->d), dmin = nto_fp16_to_fp32(blk->dmin);
            const float* xb = x + b * 256;
            double bsum = 0.0;
            for (int chunk = 0; chunk < 4; chunk++) {
                uint8_t sc_lo, m_lo, sc_hi, m_hi;
                k4_scale_min(blk->scales, chunk * 2, &sc_lo, &m_lo);
                k4_scale_min(blk->scales, chunk * 2 + 1, &sc_hi, &m_hi);
                float d1 = d * sc_lo, m1 = dmin * m_lo, d2 = d * sc_hi, m2 = dmin * m_hi;
                const uint8_t* q = blk->qs + chunk * 32;
                double s_lo = 0, s_hi = 0, sx_lo = 0, sx_hi = 0;
                for (int l = 0; l < 32; l++) {
                    s_lo += (double)(q[l] & 0x0F) * xb[chunk * 64 + l];
                    s_hi += (double)(q[l] >> 4) * xb[chunk * 64 + l + 32];
                    sx_lo += xb[chunk * 64 + l];
                    sx_hi += xb[chunk * 64 + l + 32];
                }
                bsum += (double)d1 * s_lo - (double)m1 * sx_lo + (double)d2 * s_hi - (double)m2 * sx_hi;
            }
            sum += bsum;
        }
    } else if (dt == NTO_Q5_K) {                         /* gemm.cu:293-352 */
        int nb = in / 256;
        for (int b = 0; b < nb; b++) {
            const blk_q5_k* blk = (const blk_q5_k*)(row + (size_t)b * 176);
            float d = nto_fp16_to_fp32(blk->d), dmin = nto_fp16_to_fp32(blk->dmin);
            const float* xb = x + b * 256;
            double bsum = 0.0;
            uint8_t u1 = 1, u2 = 2;
            for (int chunk = 0; chunk < 4; chunk++) {
                uint8_t sc_lo, m_lo, sc_hi, m_hi;
                k4_scale_min(blk->scales, chunk * 2, &sc_lo, &m_lo);
                k4_scale_min(blk->scales, chunk * 2 + 1, &sc_hi, &m_hi);
                float d1 = d * sc_lo, m1 = dmin * m_lo, d2 = d * sc_hi, m2 = dmin * m_hi;
                const uint8_t* ql = blk->ql + chunk * 32;
                const uint8_t* qh = blk->qh;
                double s_lo = 0, s_hi = 0, sx_lo = 0, sx_hi = 0;
                for (int l = 0; l < 32; l++) {
                    int lo = (ql[l] & 0x0F) + ((qh[l] & u1) ? 16 : 0);
                    int hi = (ql[l] >> 4) + ((qh[l] & u2) ? 16 : 0);
                    s_lo += (double)lo * xb[chunk * 64 + l];
                    s_hi += (double)hi * xb[chunk * 64 + l + 32];
                    sx_lo += xb[chunk * 64 + l];
                    sx_hi += xb[chunk * 64 + l + 32];
                }
                bsum += (double)d1 * s_lo - (double)m1 * sx_lo + (double)d2 * s_hi - (double)m2 * sx_hi;
                u1 <<= 2; u2 <<= 2;
            }
            sum += bsum;
        }
    } else if (dt == NTO_Q6_K) {                         /* gemm.cu:421-457 */
        int nb = in / 256;
        for (int b = 0; b < nb; b++) {
            const blk_q6_k* blk = (const blk_q6_k*)(row + (size_t)b * 210);
            float d = nto_fp16_to_fp32(blk->d);
            const uint8_t* ql = blk->ql;
            const uint8_t* qh = blk->qh;
            const int8_t* sc = blk->scales;
            double bs = 0.0;
            for (int half = 0; half < 2; half++) {
                const float* xh = x + b * 256 + half * 128;
                for (int l = 0; l < 32; l++) {
                    int is = l / 16;
                    int q1 = (int)((ql[l] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
                    int q2 = (int)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                    int q3 = (int)((ql[l] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
                    int q4 = (int)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
                    bs += (double)sc[is + 0] * q1 * xh[l];
                    bs += (double)sc[is + 2] * q2 * xh[l + 32];
                    bs += (double)sc[is + 4] * q3 * xh[l + 64];
                    bs += (double)sc[is + 6] * q4 * xh[l + 96];
                }
                ql += 64; qh += 32; sc += 8;
            }
            sum += (double)d * bs;
        }
    } else if (dt == NTO_F16) {                          /* gemm.cu:546-612 */
        const uint16_t* w = (const uint16_t*)row;
        for (int i = 0; i < in; i++) sum += (double)nto_fp16_to_fp32(w[i]) * x[i];
    } else if (dt == NTO_F32) {                          /* gemm.cu:617-671 */
        const float* w = (const float*)row;
        for (int i = 0; i < in; i++) sum += (double)w[i] * x[i];
    }
    return sum;
}

/* gemm.cu:748-805: unsupported dtype prints and leaves y untouched. */
void nto_gemv(float* y, const void* W, const float* x, int out, int in, int dt) {
    size_t rb = nto_row_bytes(dt, in);
    if (rb == 0) {
        fprintf(stderr, "Unsupported dtype for GEMV: %d\n", dt);
        return;
    }
    const uint8_t* w = (const uint8_t*)W;
#pragma omp parallel for schedule(static)
    for (int r = 0; r < out; r++) y[r] = (float)gemv_row(dt, w + (size_t)r * rb, x, in);
}

/* src/cuda/rmsnorm.cu:17-70 */
void nto_rmsnorm(float* y, const float* x, const float* w, int rows, int hidden, float eps) {
    for (int r = 0; r < rows; r++) {
        const float* xr = x + (size_t)r * hidden;
        float* yr = y + (size_t)r * hidden;
        double ss = 0.0;
        for (int i = 0; i < hidden; i++) ss += (double)xr[i] * xr[i];
        float mean_sq = (float)(ss / hidden);
        float rms_inv = (float)(1.0 / sqrt((double)(mean_sq + eps)));
        for (int i = 0; i < hidden; i++) yr[i] = xr[i] * rms_inv * w[i];
    }
}

/* src/cuda/rotary.cu:16-62 / 65-107 */
void nto_rope(float* q, float* k, const int* positions, int seq_len, int n_heads,
              int n_kv_heads, int head_dim, float theta_base, float freq_scale, int interleaved) {
    int half_dim = head_dim / 2;
    for (int pass = 0; pass < 2; pass++) {
        float* data = pass ? k : q;
        int n_h = pass ? n_kv_heads : n_heads;
        for (int s = 0; s < seq_len; s++) {
            int pos = positions[s];
            for (int h = 0; h < n_h; h++) {
                float* v = data + ((size_t)s * n_h + h) * head_dim;
                for (int i = 0; i < half_dim; i++) {
                    float freq = (float)(1.0 / pow((double)theta_base, (double)((2.0f * i) / head_dim)));
                    float angle = pos * freq * freq_scale;
                    float c = (float)cos((double)angle), sn = (float)sin((double)angle);
                    int i0 = interleaved ? 2 * i : i;
                    int i1 = interleaved ? 2 * i + 1 : i + half_dim;
                    float x0 = v[i0], x1 = v[i1];
                    v[i0] = x0 * c - x1 * sn;
                    v[i1] = x1 * c + x0 * sn;
                }
            }
        }
    }
}

/* src/cuda/attention.cu:316-342 */
void nto_copy_to_kv_cache(uint16_t* kc, uint16_t* vc, const float* k, const float* v,
                          int seq_len, int n_kv, int hd, int start_pos, int max_seq) {
    for (int s = 0; s < seq_len; s++) {
        int cp = start_pos + s;
        if (cp >= max_seq) continue;                      /* attention.cu:336 */
        for (int i = 0; i < n_kv * hd; i++) {
            kc[(size_t)cp * n_kv * hd + i] = nto_fp32_to_fp16(k[(size_t)s * n_kv * hd + i]);
            vc[(size_t)cp * n_kv * hd + i] = nto_fp32_to_fp16(v[(size_t)s * n_kv * hd + i]);
        }
    }
}

/* One query row against cache rows [0, n_keys) — attention.cu:131-201 / 245-310 */
static void attend_one(float* out, const float* qv, const uint16_t* kc, const uint16_t* vc,
                       int n_keys, int kv_head, int n_kv, int hd, float scale, double* sc) {
    double mx = -1e300;
    for (int p = 0; p < n_keys; p++) {
        const uint16_t* kp = kc + ((size_t)p * n_kv + kv_head) * hd;
        double s = 0.0;
        for (int d = 0; d < hd; d++) s += (double)qv[d] * nto_fp16_to_fp32(kp[d]);
        sc[p] = (double)((float)s * scale);
        if (sc[p] > mx) mx = sc[p];
    }
    double sum = 0.0;
    for (int p = 0; p < n_keys; p++) { sc[p] = exp(sc[p] - mx); sum += sc[p]; }
    double inv = sum > 0.0 ? 1.0 / sum : 0.0;
    for (int d = 0; d < hd; d++) {
        double acc = 0.0;
        for (int p = 0; p < n_keys; p++)
            acc += sc[p] * inv * nto_fp16_to_fp32(vc[((size_t)p * n_kv + kv_head) * hd + d]);
        out[d] = (float)acc;
    }
}

/* src/cuda/attention.cu:108-202 */
void nto_attention_decode(float* out, const float* q, const uint16_t* kc, const uint16_t* vc,
                          int seq_len, int n_heads, int n_kv, int hd, int max_seq, float scale) {
    (void)max_seq;
#pragma omp parallel
    {
        double* sc = (double*)malloc(sizeof(double) * (size_t)(seq_len > 0 ? seq_len : 1));
#pragma omp for schedule(static)
        for (int h = 0; h < n_heads; h++) {
            int kvh = h / (n_heads / n_kv);
            attend_one(out + (size_t)h * hd, q + (size_t)h * hd, kc, vc, seq_len, kvh, n_kv, hd, scale, sc);
        }
        free(sc);
    }
}

/* src/cuda/attention.cu:216-311 */
void nto_attention_prefill(float* out, const float* Q, const uint16_t* kc, const uint16_t* vc,
                           int seq_len, int start_pos, int n_heads, int n_kv, int hd,
                           int max_seq, float scale) {
    (void)max_seq;
#pragma omp parallel
    {
        double* sc = (double*)malloc(sizeof(double) * (size_t)(start_pos + seq_len + 1));
#pragma omp for schedule(dynamic) collapse(2)
        for (int qi = 0; qi < seq_len; qi++) {
            for (int h = 0; h < n_heads; h++) {
                int kvh = h / (n_heads / n_kv);
                int n_keys = start_pos + qi + 1;
                attend_one(out + ((size_t)qi * n_heads + h) * hd, Q + ((size_t)qi * n_heads + h) * hd,
                           kc, vc, n_keys, kvh, n_kv, hd, scale, sc);
            }
        }
        free(sc);
    }
}

/* src/cuda/gemm.cu:713-725 */
void nto_silu_mul(float* out, const float* gate, const float* up, int n) {
    for (int i = 0; i < n; i++) {
        float g = gate[i];
        float silu = (float)((double)g / (1.0 + exp(-(double)g)));
        out[i] = silu * up[i];
    }
}
/* src/cuda/elementwise.cu:23-32 */
void nto_add_inplace(float* a, const float* b, int n) {
    for (int i = 0; i < n; i++) a[i] += b[i];
}

/* ---------------- whole model: src/model/transformer.cpp:604-669 ---------------- */
struct nto_model {
    nto_config cfg;
    const void* token_embd; int dt_embd;
    const void* output_w;   int dt_output;
    const float* output_norm;
    nto_layer* layers;
    uint16_t *k_cache, *v_cache;       /* [L, max_seq, n_kv, hd]  transformer.cpp:340-346 */
    float *hidden, *residual, *ws;
};

nto_model* nto_model_create(const nto_config* cfg) {
    nto_model* m = (nto_model*)calloc(1, sizeof(nto_model));
    m->cfg = *cfg;
    m->layers = (nto_layer*)calloc((size_t)cfg->n_layers, sizeof(nto_layer));
    size_t kv = (size_t)cfg->n_layers * cfg->max_seq_len * cfg->n_kv_heads * cfg->head_dim;
    m->k_cache = (uint16_t*)calloc(kv, 2);
    m->v_cache = (uint16_t*)calloc(kv, 2);
    return m;
}
void nto_model_destroy(nto_model* m) {
    if (!m) return;
    free(m->layers); free(m->k_cache); free(m->v_cache);
    free(m->hidden); free(m->residual); free(m->ws);
    free(m);
}
void nto_model_set_globals(nto_model* m, const void* te, int dte, const void* ow, int dto, const float* on) {
    m->token_embd = te; m->dt_embd = dte; m->output_w = ow; m->dt_output = dto; m->output_norm = on;
}
void nto_model_set_layer(nto_model* m, int i, const nto_layer* l) { m->layers[i] = *l; }

int nto_model_forward(nto_model* m, const int* tokens, int seq_len, int start_pos,
                      float* logits, int n_layers_run) {
    const nto_config* c = &m->cfg;
    int hidden = c->hidden_size, inter = c->intermediate_size;
    int nh = c->n_heads, nkv = c->n_kv_heads, hd = c->head_dim, max_seq = c->max_seq_len;
    int q_dim = nh * hd, kv_dim = nkv * hd;
    if (n_layers_run < 0 || n_layers_run > c->n_layers) n_layers_run = c->n_layers;

    size_t attn_ws = (size_t)seq_len * (2 * q_dim + 2 * kv_dim);     /* attention.cpp:106-118 */
    size_t ffn_ws = (size_t)2 * seq_len * inter;                       /* ffn.cpp:85-90 */
    size_t ws_n = attn_ws > ffn_ws ? attn_ws : ffn_ws;
    m->hidden = (float*)realloc(m->hidden, sizeof(float) * (size_t)seq_len * hidden);
    m->residual = (float*)realloc(m->residual, sizeof(float) * (size_t)seq_len * hidden);
    m->ws = (float*)realloc(m->ws, sizeof(float) * ws_n);
    float *h = m->hidden, *res = m->residual, *ws = m->ws;

    /* 1. embedding lookup — transformer.cpp:419-599 (Q5_K unsupported -> zeros, :595-598) */
    size_t erb = nto_row_bytes(m->dt_embd, hidden);
    for (int t = 0; t < seq_len; t++) {
        if (m->dt_embd == NTO_Q5_K) {
            memset(h + (size_t)t * hidden, 0, sizeof(float) * hidden);
        } else {
            nto_dequant_row(m->dt_embd, (const uint8_t*)m->token_embd + (size_t)tokens[t] * erb,
                            hidden, h + (size_t)t * hidden);
        }
    }
    /* 2. positions — transformer.cpp:619-623 */
    int* pos = (int*)malloc(sizeof(int) * (size_t)seq_len);
    for (int i = 0; i < seq_len; i++) pos[i] = start_pos + i;
    float scale = 1.0f / sqrtf((float)hd);                              /* attention.cpp:20 */
    size_t kv_stride = (size_t)max_seq * nkv * hd;                      /* transformer.cpp:629 */

    for (int i = 0; i < n_layers_run; i++) {
        const nto_layer* L = &m->layers[i];
        int n = seq_len * hidden;
        nto_rmsnorm(res, h, L->attn_norm, seq_len, hidden, c->norm_eps);          /* :635 */
        float* q_buf = ws;                                                        /* attention.cpp:133-137 */
        float* k_buf = q_buf + (size_t)seq_len * q_dim;
        float* v_buf = k_buf + (size_t)seq_len * kv_dim;
        float* attn_out = v_buf + (size_t)seq_len * kv_dim;
        for (int t = 0; t < seq_len; t++) {                                       /* attention.cpp:144-162 */
            const float* inp = res + (size_t)t * hidden;
            nto_gemv(q_buf + (size_t)t * q_dim, L->wq, inp, q_dim, hidden, L->dt_q);
            nto_gemv(k_buf + (size_t)t * kv_dim, L->wk, inp, kv_dim, hidden, L->dt_k);
            nto_gemv(v_buf + (size_t)t * kv_dim, L->wv, inp, kv_dim, hidden, L->dt_v);
        }
        nto_rope(q_buf, k_buf, pos, seq_len, nh, nkv, hd, c->rope_theta, 1.0f, 0); /* :165-170 */
        uint16_t* kc = m->k_cache + (size_t)i * kv_stride;
        uint16_t* vc = m->v_cache + (size_t)i * kv_stride;
        nto_copy_to_kv_cache(kc, vc, k_buf, v_buf, seq_len, nkv, hd, start_pos, max_seq);
        if (seq_len == 1)                                                         /* :182-197 */
            nto_attention_decode(attn_out, q_buf, kc, vc, start_pos + 1, nh, nkv, hd, max_seq, scale);
        else
            nto_attention_prefill(attn_out, q_buf, kc, vc, seq_len, start_pos, nh, nkv, hd, max_seq, scale);
        for (int t = 0; t < seq_len; t++)                                         /* :200-210 */
            nto_gemv(res + (size_t)t * hidden, L->wo, attn_out + (size_t)t * q_dim, hidden, q_dim, L->dt_o);
        nto_add_inplace(h, res, n);                                               /* transformer.cpp:647 */

        nto_rmsnorm(res, h, L->ffn_norm, seq_len, hidden, c->norm_eps);           /* :650 */
        float* gate = ws;                                                         /* ffn.cpp:96-133 */
        float* up = gate + (size_t)seq_len * inter;
        for (int t = 0; t < seq_len; t++) {
            const float* inp = res + (size_t)t * hidden;
            float* g = gate + (size_t)t * inter;
            float* u = up + (size_t)t * inter;
            nto_gemv(g, L->w_gate, inp, inter, hidden, L->dt_gate);
            nto_gemv(u, L->w_up, inp, inter, hidden, L->dt_up);
            nto_silu_mul(g, g, u, inter);
            nto_gemv(res + (size_t)t * hidden, L->w_down, g, hidden, inter, L->dt_down);
        }
        nto_add_inplace(h, res, n);                                               /* :654 */
    }
    /* final norm on last token in place, LM head — transformer.cpp:657-665 */
    float* last = h + (size_t)(seq_len - 1) * hidden;
    nto_rmsnorm(last, last, m->output_norm, 1, hidden, c->norm_eps);
    nto_gemv(logits, m->output_w, last, c->vocab_size, hidden, m->dt_output);
    free(pos);
    return 0;
}

int nto_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
