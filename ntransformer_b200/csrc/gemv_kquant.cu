// gemv_kquant.cu — dequant-fused GEMV over GGUF Q4_K / Q5_K / Q6_K super-blocks for sm_100a.
//
// Replaces reference kernels K3/K4/K5 (src/cuda/gemm.cu:158-470) behind launch_gemv (gemm.cu:748-805).
// The reference walks one warp per row with per-lane byte loads, an F32 convert + 2 FMAs per weight and
// x re-read from shared memory for every row.  On B200 (7.3 TB/s reads, ~48 Q4_K weights/clk/SM) that
// instruction stream is ~2x over the issue budget, so this kernel is organised differently:
//
//   * weights: each warp owns a private ring of TMA bulk copies (cp.async.bulk + mbarrier): one stage =
//     4 rows x 16 super-blocks (>= 2304 B per copy), filled by the warp's lane 0 and consumed by the
//     same warp, so there is no CTA-wide barrier in the streaming loop and HBM requests stay in flight
//     across row-group boundaries;
//   * activations: block-scaled 3-term int8 (kernels_internal.h "xq"), staged once per CTA in shared
//     memory (XOR-swizzled so the lane -> half-block mapping is bank-conflict free) and re-used for the
//     4 rows of a stage from registers;
//   * math: IDP.4A on the 4/5/6-bit codes (exact integer partial sums), one F32 scale per 16/32 weights,
//     6-bit scale unpack amortised over a 128-weight half super-block per lane;
//   * reduction: a 4-row transpose-reduce (6 shuffles) once per row-group.
// No tensor cores: the path is HBM-bound (BASELINE.json north_star).
#include "kernels_internal.h"
#include "ring.cuh"
#include <cuda_fp16.h>

namespace nt { namespace b200 {

namespace {

constexpr int RG = 4;            // rows per warp stage
constexpr int BS = 16;           // super-blocks per stage chunk (two lanes per super-block)
constexpr int MAX_MATS = 3;

struct KqMat {
    const uint8_t* W;
    float* y;
    int out;
    int groups;        // ceil(out / RG)
    int blk_bytes;     // 144 / 176 / 210
    int fmt;           // 0 = Q4_K, 1 = Q5_K, 2 = Q6_K
    long long row_pitch;
};
struct KqParams {
    KqMat mat[MAX_MATS];
    int n_mat;
    int K, NB, NC;             // elements, super-blocks per row, chunks per row
    const int8_t* xq;
    int total_groups;          // SWIGLU: groups of mat[0]
    int n_seg;                 // segments (matrices) per task: 2 for SWIGLU else 1
    int epilogue;
    int stages;                // ring depth per warp
    int slot_bytes;            // RG * BS * max blk_bytes
};

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ int4 lds128(uint32_t addr) {
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float h2f(uint32_t h16) { return __half2float(__ushort_as_half((unsigned short)h16)); }

// Byte offset of element e inside an x plane: 16-byte columns are XORed with the half-block index so
// that the 8 lanes of a quarter-warp (8 consecutive half-blocks, 128 B apart) hit 8 different columns.
__device__ __forceinline__ uint32_t xswz(uint32_t e) { return e ^ (((e >> 7) & 7u) << 4); }

__device__ __forceinline__ int combine3(const int s[3]) { return (s[0] * 128 + s[1]) * 128 + s[2]; }

// ---- one 64-weight pass of a Q4_K / Q5_K half super-block for RG rows --------------------------
// xr[p][0..7] = x terms for the low sub-block (32 elements), xr[p][8..15] for the high sub-block.
template <int FMT>
__device__ __forceinline__ void pass_q45(const int (&xr)[3][16], float sx_lo, float sx_hi, float sum_lo, float sum_hi,
                                         uint32_t q_addr, uint32_t qh_addr, int row_pitch, int cg /* chunk 0..3 */,
                                         const uint32_t (&sc4)[RG], const uint32_t (&m4)[RG], int cc,
                                         float (&A)[RG], float (&B)[RG]) {
#pragma unroll
    for (int r = 0; r < RG; r++) {
        int4 qa = lds128(q_addr + r * row_pitch);
        int4 qb = lds128(q_addr + r * row_pitch + 16);
        uint32_t q[8] = {(uint32_t)qa.x, (uint32_t)qa.y, (uint32_t)qa.z, (uint32_t)qa.w,
                         (uint32_t)qb.x, (uint32_t)qb.y, (uint32_t)qb.z, (uint32_t)qb.w};
        uint32_t qh[8];
        if (FMT == 1) {
            int4 ha = lds128(qh_addr + r * row_pitch);
            int4 hb = lds128(qh_addr + r * row_pitch + 16);
            qh[0] = ha.x; qh[1] = ha.y; qh[2] = ha.z; qh[3] = ha.w;
            qh[4] = hb.x; qh[5] = hb.y; qh[6] = hb.z; qh[7] = hb.w;
        }
        int slo[3] = {0, 0, 0}, shi[3] = {0, 0, 0};
#pragma unroll
        for (int w = 0; w < 8; w++) {
            uint32_t lo = q[w] & 0x0F0F0F0Fu;
            uint32_t hi = (q[w] >> 4) & 0x0F0F0F0Fu;
            if (FMT == 1) {
                uint32_t t = qh[w] >> (2 * cg);             // bit0 -> low sub-block, bit1 -> high sub-block
                lo |= (t << 4) & 0x10101010u;
                hi |= (t << 3) & 0x10101010u;
            }
#pragma unroll
            for (int p = 0; p < 3; p++) {
                slo[p] = dp4a_us(lo, xr[p][w], slo[p]);
                shi[p] = dp4a_us(hi, xr[p][8 + w], shi[p]);
            }
        }
        float flo = (float)combine3(slo) * sx_lo;
        float fhi = (float)combine3(shi) * sx_hi;
        float sc_lo = (float)((sc4[r] >> (16 * cc)) & 0xFFu), sc_hi = (float)((sc4[r] >> (16 * cc + 8)) & 0xFFu);
        float m_lo = (float)((m4[r] >> (16 * cc)) & 0xFFu), m_hi = (float)((m4[r] >> (16 * cc + 8)) & 0xFFu);
        A[r] = fmaf(sc_lo, flo, fmaf(sc_hi, fhi, A[r]));
        B[r] = fmaf(m_lo, sum_lo, fmaf(m_hi, sum_hi, B[r]));
    }
}

// Load n+1 aligned words starting at (addr & ~3) and funnel them into n words starting at addr
// (addr is 2-byte aligned: Q6_K super-blocks are 210 B).
template <int N>
__device__ __forceinline__ void lds_funnel(uint32_t addr, uint32_t sel, uint32_t (&out)[N]) {
    uint32_t a = addr & ~3u;
    uint32_t prev = lds32(a);
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint32_t next = lds32(a + 4 * (i + 1));
        out[i] = __byte_perm(prev, next, sel);
        prev = next;
    }
}

// ---- one pass (kk = 0/1: l in [16kk, 16kk+16)) of a Q6_K half super-block for RG rows -----------
// xr[p][4*j + i]: x terms for run j (elements j*32 + 16kk + 4i ..), j = 0..3.
__device__ __forceinline__ void pass_q6(const int (&xr)[3][16], const float (&sx)[4], const float (&c32)[4],
                                        uint32_t half_addr /* ql half base of row 0 */, uint32_t qh_addr, int row_pitch,
                                        uint32_t sel, int kk, const uint32_t (&scw)[RG][2], float (&A)[RG]) {
#pragma unroll
    for (int r = 0; r < RG; r++) {
        uint32_t qa[4], qb[4], qh[4];
        lds_funnel<4>(half_addr + r * row_pitch + 16 * kk, sel, qa);        // ql[l]
        lds_funnel<4>(half_addr + r * row_pitch + 32 + 16 * kk, sel, qb);   // ql[l + 32]
        lds_funnel<4>(qh_addr + r * row_pitch + 16 * kk, sel, qh);
        int s[4][3] = {};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            uint32_t h = qh[i];
            uint32_t q1 = (qa[i] & 0x0F0F0F0Fu) | ((h << 4) & 0x30303030u);
            uint32_t q2 = (qb[i] & 0x0F0F0F0Fu) | ((h << 2) & 0x30303030u);
            uint32_t q3 = ((qa[i] >> 4) & 0x0F0F0F0Fu) | (h & 0x30303030u);
            uint32_t q4 = ((qb[i] >> 4) & 0x0F0F0F0Fu) | ((h >> 2) & 0x30303030u);
#pragma unroll
            for (int p = 0; p < 3; p++) {
                s[0][p] = dp4a_us(q1, xr[p][0 + i], s[0][p]);
                s[1][p] = dp4a_us(q2, xr[p][4 + i], s[1][p]);
                s[2][p] = dp4a_us(q3, xr[p][8 + i], s[2][p]);
                s[3][p] = dp4a_us(q4, xr[p][12 + i], s[3][p]);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int b = 2 * j + kk;                                   // scale index inside the half
            int sc = (int)(signed char)((scw[r][b >> 2] >> (8 * (b & 3))) & 0xFFu);
            // sum over the 16 weights of sc * (q - 32) * x = sc * (S * sx - 32 * sum16)
            A[r] = fmaf((float)sc, fmaf((float)combine3(s[j]), sx[j], -c32[j]), A[r]);
        }
    }
}

template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) gemv_kq_kernel(const __grid_constant__ KqParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int K = p.K;
    // smem carve-up: [x planes 3K][scale K/32 f32][sum16 K/16 f32][pad to 128][rings][mbarriers]
    uint8_t* xs = smem;
    float* xscale = reinterpret_cast<float*>(smem + 3 * (size_t)K);
    float* xsum16 = xscale + K / 32;
    size_t ring_off = (3 * (size_t)K + (size_t)(K / 32) * 4 + (size_t)(K / 16) * 4 + 127) & ~(size_t)127;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = smem + ring_off + (size_t)warp * p.stages * p.slot_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ring_off + (size_t)WARPS * p.stages * p.slot_bytes) + warp * p.stages;

    const int gw = blockIdx.x * WARPS + warp, nw = gridDim.x * WARPS;
    const int n_tasks = (gw < p.total_groups) ? (p.total_groups - gw + nw - 1) / nw : 0;
    const int per_task = p.n_seg * p.NC;
    const int n_stages_total = n_tasks * per_task;

    // ---- producer: issue the copies of flattened stage s into ring slot ----
    auto issue = [&](int s, int slot) {
        int task = s / per_task, rem = s - task * per_task;
        int seg = rem / p.NC, chunk = rem - seg * p.NC;
        int g = gw + task * nw;
        int mi = seg;
        if (p.n_seg == 1) {
            mi = 0;
            while (mi + 1 < p.n_mat && g >= p.mat[mi].groups) { g -= p.mat[mi].groups; mi++; }
        }
        const KqMat& m = p.mat[mi];
        int nbc = min(BS, p.NB - chunk * BS);
        // copy size rounded up to 16 B (a 210-byte Q6_K tail may spill into the row's padding; checked on the host)
        uint32_t bytes = ((uint32_t)(nbc * m.blk_bytes) + 15u) & ~15u;
        uint64_t* bar = bars + slot;
        mbar_expect_tx(bar, bytes * RG);
        uint8_t* dst = ring + (size_t)slot * p.slot_bytes;
#pragma unroll
        for (int r = 0; r < RG; r++) {
            int row = min(g * RG + r, m.out - 1);
            bulk_g2s(dst + r * (BS * m.blk_bytes), m.W + (long long)row * m.row_pitch + (long long)chunk * BS * m.blk_bytes,
                     bytes, bar);
        }
    };

    if (lane == 0) {
        for (int s = 0; s < p.stages; s++) mbar_init(bars + s, 1);
        mbar_fence_init();
    }
    __syncwarp();
    // Weights do not depend on the previous kernel: start streaming before touching x.
    if (lane == 0) {
        for (int s = 0; s < p.stages && s < n_stages_total; s++) issue(s, s);
    }
    pdl_wait();   // no-op unless launched with programmatic stream serialization

    // ---- stage xq into shared memory (swizzled planes) ----
    {
        const int n16 = 3 * K / 16;
        const int4* src = reinterpret_cast<const int4*>(p.xq);
        for (int i = threadIdx.x; i < n16; i += WARPS * 32) {
            uint32_t byte = (uint32_t)i * 16u;
            uint32_t plane = byte / (uint32_t)K, e = byte - plane * (uint32_t)K;
            *reinterpret_cast<int4*>(xs + plane * (uint32_t)K + xswz(e)) = __ldg(src + i);
        }
        const float* fsrc = reinterpret_cast<const float*>(p.xq + 3 * (size_t)K);
        const int nf = K / 32 + K / 16;
        for (int i = threadIdx.x; i < nf; i += WARPS * 32) xscale[i] = __ldg(fsrc + i);
    }
    __syncthreads();
    pdl_launch_dependents();

    const uint32_t xs_a = smem_u32(xs);
    const int blk = lane >> 1, h = lane & 1;
    float acc[RG] = {0.f, 0.f, 0.f, 0.f};
    float gate_keep = 0.f;

    for (int s = 0; s < n_stages_total; s++) {
        const int slot = s % p.stages;
        const uint32_t parity = (uint32_t)((s / p.stages) & 1);
        int task = s / per_task, rem = s - task * per_task;
        int seg = rem / p.NC, chunk = rem - seg * p.NC;
        int g = gw + task * nw;
        int mi = seg;
        if (p.n_seg == 1) {
            mi = 0;
            while (mi + 1 < p.n_mat && g >= p.mat[mi].groups) { g -= p.mat[mi].groups; mi++; }
        }
        const KqMat& m = p.mat[mi];
        const int nbc = min(BS, p.NB - chunk * BS);
        const int row_pitch = BS * m.blk_bytes;

        mbar_wait(bars + slot, parity);

        if (blk < nbc) {
            const uint32_t base = smem_u32(ring + (size_t)slot * p.slot_bytes) + blk * m.blk_bytes;
            const uint32_t hb = (uint32_t)((chunk * BS + blk) * 2 + h);      // global half-block index
            if (m.fmt <= 1) {
                // ---------------- Q4_K / Q5_K ----------------
                const uint32_t qs_off = (m.fmt == 0) ? 16u : 48u;
                uint32_t sc4[RG], m4[RG];
                float d[RG], dmin[RG];
#pragma unroll
                for (int r = 0; r < RG; r++) {
                    int4 hd = lds128(base + r * row_pitch);
                    uint32_t w0 = hd.y, w1 = hd.z, w2 = hd.w;
                    d[r] = h2f((uint32_t)hd.x & 0xFFFFu);
                    dmin[r] = h2f((uint32_t)hd.x >> 16);
                    if (h == 0) {
                        sc4[r] = w0 & 0x3F3F3F3Fu;
                        m4[r] = w1 & 0x3F3F3F3Fu;
                    } else {
                        sc4[r] = (w2 & 0x0F0F0F0Fu) | ((w0 >> 2) & 0x30303030u);
                        m4[r] = ((w2 >> 4) & 0x0F0F0F0Fu) | ((w1 >> 2) & 0x30303030u);
                    }
                }
                float A[RG] = {0.f, 0.f, 0.f, 0.f}, B[RG] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int cc = 0; cc < 2; cc++) {
                    int xr[3][16];
#pragma unroll
                    for (int pl = 0; pl < 3; pl++) {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            uint32_t col = (uint32_t)(cc * 4 + i) ^ (hb & 7u);
                            int4 v = lds128(xs_a + pl * K + hb * 128u + col * 16u);
                            xr[pl][4 * i + 0] = v.x; xr[pl][4 * i + 1] = v.y; xr[pl][4 * i + 2] = v.z; xr[pl][4 * i + 3] = v.w;
                        }
                    }
                    const int b32 = hb * 4 + cc * 2, b16 = hb * 8 + cc * 4;
                    float sx_lo = xscale[b32], sx_hi = xscale[b32 + 1];
                    float sum_lo = xsum16[b16] + xsum16[b16 + 1], sum_hi = xsum16[b16 + 2] + xsum16[b16 + 3];
                    if (m.fmt == 0)
                        pass_q45<0>(xr, sx_lo, sx_hi, sum_lo, sum_hi, base + qs_off + h * 64 + cc * 32, 0, row_pitch,
                                    2 * h + cc, sc4, m4, cc, A, B);
                    else
                        pass_q45<1>(xr, sx_lo, sx_hi, sum_lo, sum_hi, base + qs_off + h * 64 + cc * 32, base + 16, row_pitch,
                                    2 * h + cc, sc4, m4, cc, A, B);
                }
#pragma unroll
                for (int r = 0; r < RG; r++) acc[r] += d[r] * A[r] - dmin[r] * B[r];
            } else {
                // ---------------- Q6_K ----------------
                const uint32_t sel = (base & 2u) ? 0x5432u : 0x3210u;
                uint32_t scw[RG][2];
                float d[RG];
#pragma unroll
                for (int r = 0; r < RG; r++) {
                    lds_funnel<2>(base + r * row_pitch + 192 + 8 * h, sel, scw[r]);
                    uint32_t da = base + r * row_pitch + 208;
                    uint32_t w = lds32(da & ~3u);
                    d[r] = h2f((da & 2u) ? (w >> 16) : (w & 0xFFFFu));
                }
                float A[RG] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int kk = 0; kk < 2; kk++) {
                    int xr[3][16];
                    float sx[4], c32[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        uint32_t col = (uint32_t)(2 * j + kk) ^ (hb & 7u);
#pragma unroll
                        for (int pl = 0; pl < 3; pl++) {
                            int4 v = lds128(xs_a + pl * K + hb * 128u + col * 16u);
                            xr[pl][4 * j + 0] = v.x; xr[pl][4 * j + 1] = v.y; xr[pl][4 * j + 2] = v.z; xr[pl][4 * j + 3] = v.w;
                        }
                        sx[j] = xscale[hb * 4 + j];
                        c32[j] = 32.0f * xsum16[hb * 8 + 2 * j + kk];
                    }
                    pass_q6(xr, sx, c32, base + 64 * h, base + 128 + 32 * h, row_pitch, sel, kk, scw, A);
                }
#pragma unroll
                for (int r = 0; r < RG; r++) acc[r] = fmaf(d[r], A[r], acc[r]);
            }
        }
        __syncwarp();
        if (lane == 0 && s + p.stages < n_stages_total) issue(s + p.stages, slot);

        if (chunk == p.NC - 1) {
            // ---- 4-row transpose-reduce: lanes with (lane & 7) == 0 end up holding one row each ----
            const bool b4 = lane & 16, b3 = lane & 8;
            float s0 = b4 ? acc[0] : acc[2], s1 = b4 ? acc[1] : acc[3];
            float k0 = b4 ? acc[2] : acc[0], k1 = b4 ? acc[3] : acc[1];
            k0 += __shfl_xor_sync(0xFFFFFFFFu, s0, 16);
            k1 += __shfl_xor_sync(0xFFFFFFFFu, s1, 16);
            float sv = b3 ? k0 : k1, kv = b3 ? k1 : k0;
            kv += __shfl_xor_sync(0xFFFFFFFFu, sv, 8);
            kv += __shfl_xor_sync(0xFFFFFFFFu, kv, 4);
            kv += __shfl_xor_sync(0xFFFFFFFFu, kv, 2);
            kv += __shfl_xor_sync(0xFFFFFFFFu, kv, 1);
            const int row = g * RG + (b4 ? 2 : 0) + (b3 ? 1 : 0);
            if ((lane & 7) == 0) {
                if (p.epilogue == GEMV_SWIGLU) {
                    if (seg == 0) {
                        gate_keep = kv;
                    } else if (row < m.out) {
                        // silu(g) * u with the reference's fast-math expression (gemm.cu:713-725)
                        float gv = gate_keep;
                        p.mat[0].y[row] = __fdividef(gv, 1.0f + __expf(-gv)) * kv;
                    }
                } else if (row < m.out) {
                    if (p.epilogue == GEMV_ADD) m.y[row] += kv; else m.y[row] = kv;
                }
            }
            acc[0] = acc[1] = acc[2] = acc[3] = 0.f;
        }
    }
}

int fmt_of(DType dt) { return dt == DType::Q4_K_M ? 0 : dt == DType::Q5_K ? 1 : dt == DType::Q6_K ? 2 : -1; }

int g_num_sms = 0;
int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    return g_num_sms;
}

template <int WARPS>
void launch_kq(const KqParams& p, size_t smem, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        NT_CUDA_CHECK(cudaFuncSetAttribute(gemv_kq_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = true;
    }
    int grid = num_sms();
    int need = (p.total_groups + WARPS - 1) / WARPS;
    if (need < grid) grid = need;
    gemv_kq_kernel<WARPS><<<grid, WARPS * 32, smem, s>>>(p);
    count_launch();
}

}  // namespace

bool gemv_kq_supported(const GemvMat* mats, int n_mat, int K) {
    if (n_mat < 1 || n_mat > MAX_MATS || K % 256 != 0 || K <= 0) return false;
    if (3.375 * K > 120 * 1024) return false;   // xq must leave room for the rings
    for (int i = 0; i < n_mat; i++) {
        int f = fmt_of(mats[i].dtype);
        if (f < 0 || mats[i].out <= 0) return false;
        size_t pitch = mats[i].row_pitch ? mats[i].row_pitch : dtype_row_size(mats[i].dtype, K);
        if ((reinterpret_cast<uintptr_t>(mats[i].W) & 15) || (pitch & 15)) return false;
        // the last chunk of a row is copied in 16-byte units and must stay inside the row pitch
        const size_t blk = dtype_size(mats[i].dtype);
        const int NB = K / 256, NC = (NB + BS - 1) / BS, last = NB - (NC - 1) * BS;
        const size_t tail_end = (size_t)(NC - 1) * BS * blk + ((last * blk + 15) & ~(size_t)15);
        if (tail_end > pitch) return false;
    }
    return true;
}

void gemv_kq(const GemvMat* mats, int n_mat, int K, const void* xq, GemvEpilogue ep, cudaStream_t s) {
    NT_CHECK(gemv_kq_supported(mats, n_mat, K), "gemv_kq: unsupported shape/dtype/alignment");
    KqParams p{};
    p.n_mat = n_mat;
    p.K = K; p.NB = K / 256; p.NC = (p.NB + BS - 1) / BS;
    p.xq = static_cast<const int8_t*>(xq);
    p.epilogue = (int)ep;
    int max_blk = 0, total = 0;
    for (int i = 0; i < n_mat; i++) {
        KqMat& m = p.mat[i];
        m.W = static_cast<const uint8_t*>(mats[i].W);
        m.y = mats[i].y;
        m.out = mats[i].out;
        m.groups = (mats[i].out + RG - 1) / RG;
        m.blk_bytes = (int)dtype_size(mats[i].dtype);
        m.fmt = fmt_of(mats[i].dtype);
        m.row_pitch = (long long)(mats[i].row_pitch ? mats[i].row_pitch : dtype_row_size(mats[i].dtype, K));
        if (m.blk_bytes > max_blk) max_blk = m.blk_bytes;
        total += m.groups;
    }
    if (ep == GEMV_SWIGLU) {
        NT_CHECK(n_mat == 2 && mats[0].out == mats[1].out, "gemv_kq: SWIGLU needs gate and up of equal rows");
        p.n_seg = 2;
        p.total_groups = p.mat[0].groups;
    } else {
        p.n_seg = 1;
        p.total_groups = total;
    }
    p.slot_bytes = RG * BS * max_blk;
    const size_t xq_sz = ((size_t)3 * K + (size_t)(K / 32) * 4 + (size_t)(K / 16) * 4 + 127) & ~(size_t)127;
    const size_t budget = 227 * 1024 - xq_sz - 256;
    // pick the widest CTA that still affords a 2-deep ring, then the deepest ring (<= 4)
    int warps = 8;
    while (warps > 4 && (size_t)warps * 2 * (p.slot_bytes + 8) > budget) warps -= 2;
    int stages = (int)(budget / ((size_t)warps * (p.slot_bytes + 8)));
    if (stages > 4) stages = 4;
    NT_CHECK(stages >= 1, "gemv_kq: shared memory budget exceeded");
    p.stages = stages;
    size_t smem = xq_sz + (size_t)warps * stages * (p.slot_bytes + 8);
    if (warps == 8) launch_kq<8>(p, smem, s);
    else if (warps == 6) launch_kq<6>(p, smem, s);
    else launch_kq<4>(p, smem, s);
}

}}  // namespace nt::b200
