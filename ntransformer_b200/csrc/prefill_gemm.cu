// prefill_gemm.cu — tensor-core GEMM for prefill: C[M,N] (F32) = A[M,K] (F32) . W[N,K]^T (F16 weights), sm_100a.
//
// The reference has no batched matmul on its forward path: prefill is a host loop of per-token GEMVs that re-reads
// every weight matrix once per prompt token (src/model/attention.cpp:144-162, src/model/ffn.cpp:96-133; its
// launch_gemm_f32, src/cuda/gemm.cu:677-694, is dead code).  This is the dense contraction BASELINE.json config 5
// names, written for Blackwell's 5th-generation tensor cores:
//   * operands staged by TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) into a 4-stage shared-memory ring,
//   * tcgen05.mma.cta_group::1.kind::f16 issued by one elected thread, 128x128 F32 accumulator in TMEM,
//   * tcgen05.commit -> mbarrier hand-offs (smem slot free / accumulator ready), epilogue via tcgen05.ld.
// F32 activations keep (almost) their precision on F16 tensor cores by splitting a = hi + lo (two F16 values,
// ~22 mantissa bits) and accumulating both products into the same TMEM tile; weights are F16 exactly.
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..5 = epilogue.
#include "kernels_internal.h"
#include "ring.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <mutex>

namespace nt { namespace b200 {

namespace {

constexpr int BM = 128, BK = 64;                     // CTA tile rows; BK halfs = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int TILE_A = BM * BK * 2;                  // 16 KB
// Tile width BN (= UMMA N = TMEM columns) is 256 when N allows it, else 128.  The kernel is bound by operand fetch, not by
// the tensor pipe: per 16-deep k-step a CTA pulls 64*BM + 32*BN bytes (A twice: hi and lo) for BM*BN/128 MMA cycles, i.e.
// 96 B/clk/SM at 128x128 against ~43 B/clk/SM of L2 bandwidth; 128x256 needs 64 B/clk (profiles/r01_prefill_*).
template <int BN> struct Cfg {
    static constexpr int TILE_B = BN * BK * 2;
    static constexpr int STAGE_BYTES = 2 * TILE_A + TILE_B;     // A_hi, A_lo, B
    static constexpr int STAGES = BN == 256 ? 3 : 4;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 1024;
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A = B = F16 (0), both K-major, N >> 3 at
    // [17,23), M >> 4 at [24,29)
    static constexpr uint32_t IDESC = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

// a = hi + lo with hi = f16(a), lo = f16(a - hi): 4 floats per thread
__global__ void split_f32_kernel(__half2* __restrict__ hi, __half2* __restrict__ lo, const float4* __restrict__ a, size_t n4) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n4) {
        const float4 v = a[i];
        const __half2 h0 = __floats2half2_rn(v.x, v.y), h1 = __floats2half2_rn(v.z, v.w);
        const float2 f0 = __half22float2(h0), f1 = __half22float2(h1);
        hi[2 * i] = h0; hi[2 * i + 1] = h1;
        lo[2 * i] = __floats2half2_rn(v.x - f0.x, v.y - f0.y);
        lo[2 * i + 1] = __floats2half2_rn(v.z - f1.x, v.w - f1.y);
    }
}

// bounded mbarrier wait: a mis-programmed pipeline traps instead of hanging the GPU
__device__ __forceinline__ void wait_or_trap(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 24); spin++) {
        asm volatile(
            "{\n .reg .pred p;\n"
            " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address >> 4 in
// [0,14), leading byte offset [16,30) (unused for swizzled K-major, canonical value 1), stride byte offset [32,46) =
// 8 rows x 128 B = 1024 B, version 1 at [46,48), layout type SWIZZLE_128B = 2 at [61,64).
__device__ __forceinline__ uint64_t make_desc(const void* smem_ptr) {
    const uint64_t addr = (uint64_t)(smem_u32(smem_ptr) & 0x3FFFFu) >> 4;
    return addr | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
        " tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 consecutive F32 accumulator columns of this warp's 32 TMEM lanes (one row per thread)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

enum : int { MODE_STORE = 0, MODE_ADD = 1, MODE_SWIGLU = 2 };

// Persistent, warp-specialised GEMM: CTA b handles output tiles b, b + grid, ...; the shared-memory ring and the two TMEM
// accumulators run across tile boundaries, so the epilogue of tile i overlaps the MMAs of tile i + 1.
//   MODE_STORE  C = A.W^T              MODE_ADD  C += A.W^T (residual)
//   K-chunked accumulation (STORE / ADD): the tensor cores truncate the F32 accumulator once per 16-deep MMA step, an error that
//   grows with the number of steps summed into one accumulator (~K/16 * 2^-25 relative, twice that with the hi/lo activation
//   split).  A tile's K range is therefore cut into chunks of kc_blocks k-blocks; each chunk gets a fresh TMEM accumulator and the
//   epilogue adds the chunk's partial sum into C with an ordinary (round-to-nearest) F32 add.  The same CTA walks a tile's chunks
//   back to back, so the read-modify-write of C is race-free and overlaps the next chunk's MMAs through the second accumulator.
//   MODE_SWIGLU (BN = 256 as 128 gate + 128 up columns, W = gate via map_b, up via map_b2):
//               split_out = F16 hi/lo split of silu(A.Wg^T) * (A.Wu^T), i.e. the down projection's GEMM input, so the
//               F32 gate/up activations never touch HBM (ffn.cpp:96-133 computes them as three launches).
template <int BN, int MODE>
__global__ void __launch_bounds__(192, 1) gemm_f16_tc_kernel(const __grid_constant__ CUtensorMap map_a,
                                                             const __grid_constant__ CUtensorMap map_b,
                                                             const __grid_constant__ CUtensorMap map_b2,
                                                             float* __restrict__ C, __half* __restrict__ split_out,
                                                             int M, int Mp, int N, int K, int tiles_m, int tiles_n, int n_tiles, int n_outer,
                                                             int kc_blocks) {
    constexpr int STAGES = Cfg<BN>::STAGES, STAGE_BYTES = Cfg<BN>::STAGE_BYTES, TMEM_COLS = 2 * BN;
    constexpr int TN = MODE == MODE_SWIGLU ? 128 : BN;          // output columns per tile
    constexpr uint32_t IDESC = MODE == MODE_SWIGLU ? Cfg<128>::IDESC : Cfg<BN>::IDESC;
    static_assert(MODE != MODE_SWIGLU || BN == 256, "SwiGLU tiles pair 128 gate with 128 up columns");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb = K / BK;
    const int n_kc = (MODE == MODE_SWIGLU) ? 1 : (num_kb + kc_blocks - 1) / kc_blocks;      // accumulation chunks per tile

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; a++) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 4); }   // 4 epilogue warps
        mbar_fence_init();
    }
    if (warp == 1) {                                   // whole warp: allocate both accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {                               // ---- TMA producer ----
            uint32_t it = 0;                           // k-block counter across all of this CTA's tiles
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (n_outer ? tile % tiles_m : tile / tiles_n) * BM, n0 = (n_outer ? tile / tiles_m : tile % tiles_n) * TN;
                for (int kb = 0; kb < num_kb; kb++, it++) {         // chunk boundaries do not matter to the producer
                    const uint32_t s = it % STAGES;
                    wait_or_trap(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
                    uint8_t* st = smem + (size_t)s * STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    tma_load_2d(st, &map_a, kb * BK, m0, &full_bar[s]);                 // A_hi rows [m0, m0+128)
                    tma_load_2d(st + TILE_A, &map_a, kb * BK, Mp + m0, &full_bar[s]);   // A_lo lives below A_hi
                    tma_load_2d(st + 2 * TILE_A, &map_b, kb * BK, n0, &full_bar[s]);
                    if (MODE == MODE_SWIGLU) tma_load_2d(st + 2 * TILE_A + 128 * BK * 2, &map_b2, kb * BK, n0, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {                               // ---- MMA issuer ----
            uint32_t it = 0, lt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int kc = 0; kc < n_kc; kc++, lt++) {
                const uint32_t acc = lt & 1, acc_phase = (lt >> 1) & 1;
                wait_or_trap(&tmem_empty_bar[acc], acc_phase ^ 1);         // epilogue drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + acc * BN;
                const int kb0 = (MODE == MODE_SWIGLU) ? 0 : kc * kc_blocks, kb1 = (MODE == MODE_SWIGLU) ? num_kb : min(num_kb, kb0 + kc_blocks);
                for (int kb = kb0; kb < kb1; kb++, it++) {
                    const uint32_t s = it % STAGES;
                    wait_or_trap(&full_bar[s], (it / STAGES) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint8_t* st = smem + (size_t)s * STAGE_BYTES;
                    const uint64_t da_hi = make_desc(st), da_lo = make_desc(st + TILE_A), db = make_desc(st + 2 * TILE_A);
                    const uint64_t db2 = make_desc(st + 2 * TILE_A + 128 * BK * 2);
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; k++) {  // advance 32 bytes (>> 4 = 2) inside the 128-byte swizzle row
                        umma_f16(d, da_hi + 2 * k, db + 2 * k, IDESC, ((kb - kb0) | k) ? 1u : 0u);
                        if (MODE == MODE_SWIGLU) umma_f16(d + 128, da_hi + 2 * k, db2 + 2 * k, IDESC, (kb | k) ? 1u : 0u);
                    }
#pragma unroll
                    for (int k = 0; k < BK / UMMA_K; k++) {
                        umma_f16(d, da_lo + 2 * k, db + 2 * k, IDESC, 1u);
                        if (MODE == MODE_SWIGLU) umma_f16(d + 128, da_lo + 2 * k, db2 + 2 * k, IDESC, 1u);
                    }
                    umma_commit(&empty_bar[s]);        // smem slot reusable once these MMAs retire
                }
                umma_commit(&tmem_full_bar[acc]);      // accumulator complete
            }
        }
    } else {
        // ---- epilogue: TMEM -> registers -> global (warp w may only touch TMEM lanes [32 * (w % 4), +32)) ----
        const int quarter = warp & 3;
        uint32_t lt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int kc = 0; kc < n_kc; kc++, lt++) {
            const uint32_t acc = lt & 1, acc_phase = (lt >> 1) & 1;
            const bool add_c = MODE == MODE_ADD || kc > 0;           // later chunks add their partial sums to what is in C
            const int m0 = (n_outer ? tile % tiles_m : tile / tiles_n) * BM, n0 = (n_outer ? tile / tiles_m : tile % tiles_n) * TN;
            const int row = m0 + quarter * 32 + lane;
            wait_or_trap(&tmem_full_bar[acc], acc_phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t tacc = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
            for (int c = 0; c < TN / 32; c++) {
                uint32_t v[32];
                tmem_ld32(tacc + (uint32_t)(c * 32), v);
                if (MODE == MODE_SWIGLU) {
                    uint32_t u[32];
                    tmem_ld32(tacc + 128u + (uint32_t)(c * 32), u);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (row < M) {
                        __half* hi = split_out + (size_t)row * N + n0 + c * 32;
                        __half* lo = hi + (size_t)Mp * N;
#pragma unroll
                        for (int j = 0; j < 32; j += 8) {
                            uint32_t ph[4], pl[4];
#pragma unroll
                            for (int e = 0; e < 4; e++) {
                                const float g0 = __uint_as_float(v[j + 2 * e]), g1 = __uint_as_float(v[j + 2 * e + 1]);
                                const float a0 = g0 / (1.0f + expf(-g0)) * __uint_as_float(u[j + 2 * e]);      // elementwise.cu silu_mul
                                const float a1 = g1 / (1.0f + expf(-g1)) * __uint_as_float(u[j + 2 * e + 1]);
                                const __half2 h = __floats2half2_rn(a0, a1);
                                const float2 f = __half22float2(h);
                                const __half2 l = __floats2half2_rn(a0 - f.x, a1 - f.y);
                                ph[e] = *reinterpret_cast<const uint32_t*>(&h);
                                pl[e] = *reinterpret_cast<const uint32_t*>(&l);
                            }
                            *reinterpret_cast<uint4*>(hi + j) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                            *reinterpret_cast<uint4*>(lo + j) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                        }
                    }
                } else {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (row < M) {
                        float* crow = C + (size_t)row * N + n0 + c * 32;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            float4* dst = reinterpret_cast<float4*>(crow + j);
                            float4 o = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                            if (add_c) { const float4 r = *dst; o.x += r.x; o.y += r.y; o.z += r.z; o.w += r.w; }   // residual / earlier chunks
                            *dst = o;
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);             // this warp's quarter of the accumulator is free
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    __syncwarp();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS));
}

// RMSNorm (rmsnorm.cu:17-70 math) writing the F16 hi/lo split the GEMM consumes: one CTA per row
__global__ void __launch_bounds__(512) rmsnorm_split_kernel(__half* __restrict__ hi, __half* __restrict__ lo, const float* __restrict__ xin,
                                                            const float* __restrict__ w, int hidden, float eps) {
    __shared__ float red[16];
    const float* x = xin + (size_t)blockIdx.x * hidden;
    float ss = 0.f;
    for (int i = threadIdx.x; i < hidden; i += blockDim.x) { const float v = x[i]; ss += v * v; }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < (int)(blockDim.x >> 5); i++) tot += red[i];
    const float rms_inv = rsqrtf(tot / hidden + eps);
    __half2* h2 = reinterpret_cast<__half2*>(hi + (size_t)blockIdx.x * hidden);
    __half2* l2 = reinterpret_cast<__half2*>(lo + (size_t)blockIdx.x * hidden);
    for (int i = threadIdx.x; i < hidden / 2; i += blockDim.x) {
        const float2 xv = reinterpret_cast<const float2*>(x)[i], wv = reinterpret_cast<const float2*>(w)[i];
        const float a = xv.x * rms_inv * wv.x, b = xv.y * rms_inv * wv.y;
        const __half2 h = __floats2half2_rn(a, b);
        const float2 f = __half22float2(h);
        h2[i] = h;
        l2[i] = __floats2half2_rn(a - f.x, b - f.y);
    }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}
// 2-D F16 tensor [rows][cols] (cols contiguous), box = 64 cols x box_rows rows (<= 256), 128-byte swizzle
bool make_map(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 2};
    const cuuint32_t box[2] = {BK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace

size_t gemm_f16_tc_workspace_bytes(int M, int K) {
    const size_t Mp = ((size_t)M + BM - 1) / BM * BM;
    return 2 * Mp * (size_t)K * sizeof(__half);
}

bool gemm_f16_tc_supported(const void* W_f16, int N, int K, size_t row_pitch) {
    return N > 0 && K > 0 && N % 128 == 0 && K % BK == 0 && (row_pitch == 0 || row_pitch == (size_t)K * 2) &&
           (reinterpret_cast<uintptr_t>(W_f16) & 15) == 0 && encode_fn() != nullptr;
}

// workspace <- F16 hi/lo split of A[M,K] (rows padded with zeros to a multiple of 128)
void split_activations(void* workspace, const float* A, int M, int K, cudaStream_t s) {
    const size_t Mp = ((size_t)M + BM - 1) / BM * BM;
    __half* hi = static_cast<__half*>(workspace);
    __half* lo = hi + Mp * (size_t)K;
    if (Mp != (size_t)M) NT_CUDA_CHECK(cudaMemsetAsync(workspace, 0, gemm_f16_tc_workspace_bytes(M, K), s));
    const size_t n4 = (size_t)M * K / 4;
    split_f32_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(reinterpret_cast<__half2*>(hi), reinterpret_cast<__half2*>(lo),
                                                                  reinterpret_cast<const float4*>(A), n4);
    count_launch();
}

// rmsnorm(x) * w for `rows` rows, written straight into the split workspace
void rmsnorm_split(void* workspace, const float* x, const float* w, int rows, int hidden, float eps, cudaStream_t s) {
    const size_t Mp = ((size_t)rows + BM - 1) / BM * BM;
    __half* hi = static_cast<__half*>(workspace);
    rmsnorm_split_kernel<<<rows, 512, 0, s>>>(hi, hi + Mp * (size_t)hidden, x, w, hidden, eps);
    count_launch();
}

namespace {

int sm_count() {
    static int n = 0;
    if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
    return n;
}

template <int BN, int MODE>
bool launch_gemm(float* C, __half* split_out, const void* workspace, const void* W, const void* W2, int M, size_t Mp, int N, int K,
                 cudaStream_t s) {
    constexpr int TN = MODE == MODE_SWIGLU ? 128 : BN;
    CUtensorMap map_a, map_b, map_b2;
    if (!make_map(&map_a, workspace, 2 * Mp, (uint64_t)K, BM) || !make_map(&map_b, W, (uint64_t)N, (uint64_t)K, TN) ||
        !make_map(&map_b2, W2 ? W2 : W, (uint64_t)N, (uint64_t)K, TN))
        return false;
    static unsigned long long configured = 0;      // bit per device id
    opt_in_dynamic_smem(gemm_f16_tc_kernel<BN, MODE>, (int)(Cfg<BN>::SMEM), configured);
    const int tiles_m = (int)(Mp / BM), tiles_n = N / TN, n_tiles = tiles_n * tiles_m;
    const int grid = n_tiles < sm_count() ? n_tiles : sm_count();
    // Tile order: concurrently running CTAs share the operand that is walked in the inner loop, the other one should stay
    // L2-resident (126 MB) across the whole launch.  Keep the smaller operand resident: n-outer when the split activations
    // (2 * Mp * K halfs) are smaller than the weights, m-outer otherwise (ncu: 1.5 GB of DRAM reads for 184 MB of operands
    // with the wrong order, profiles/r01_gemm_tc_ncu_summary.txt).
    const size_t a_bytes = 4 * Mp * (size_t)K, w_bytes = (MODE == MODE_SWIGLU ? 4 : 2) * (size_t)N * K;
    static const bool force_m_outer = getenv("NT_B200_GEMM_M_OUTER") != nullptr;
    const int n_outer = (a_bytes < w_bytes && !force_m_outer) ? 1 : 0;
    // accumulation chunk: 4096 elements of K (64 k-blocks) per fresh accumulator — the hidden-sized GEMMs keep one accumulator, the
    // 14336-wide down projection gets four (1024-element chunks measured 29.2k vs 32.0k tok/s on the 8B F16 4096-token prompt and
    // 13.5k vs 17.6k on Q4_K_M: the read-modify-write of C per chunk is not free); NT_B200_GEMM_KC=<k-blocks> overrides (0: whole K)
    static const int kc_env = [] { const char* e = getenv("NT_B200_GEMM_KC"); return e ? atoi(e) : 64; }();
    const int kc_blocks = kc_env > 0 ? kc_env : K / BK;
    gemm_f16_tc_kernel<BN, MODE><<<grid, 192, Cfg<BN>::SMEM, s>>>(map_a, map_b, map_b2, C, split_out, M, (int)Mp, N, K, tiles_m, tiles_n,
                                                                 n_tiles, n_outer, kc_blocks);
    count_launch();
    return true;
}

}  // namespace

// C[M,N] (+)= split(A)[M,K] . W[N,K]^T with A already split into `workspace`
bool gemm_f16_tc_ws(float* C, const void* workspace, const void* W_f16, int M, int N, int K, bool add, cudaStream_t s) {
    if (M <= 0 || !gemm_f16_tc_supported(W_f16, N, K, 0)) return false;
    const size_t Mp = ((size_t)M + BM - 1) / BM * BM;
    static const bool force128 = getenv("NT_B200_GEMM_BN128") != nullptr;
    if (N % 256 == 0 && !force128)
        return add ? launch_gemm<256, MODE_ADD>(C, nullptr, workspace, W_f16, nullptr, M, Mp, N, K, s)
                   : launch_gemm<256, MODE_STORE>(C, nullptr, workspace, W_f16, nullptr, M, Mp, N, K, s);
    return add ? launch_gemm<128, MODE_ADD>(C, nullptr, workspace, W_f16, nullptr, M, Mp, N, K, s)
               : launch_gemm<128, MODE_STORE>(C, nullptr, workspace, W_f16, nullptr, M, Mp, N, K, s);
}

// workspace_out <- split(silu(A.Wgate^T) * (A.Wup^T)) for A already split into workspace_in; N = rows of Wgate/Wup
bool gemm_f16_tc_swiglu_ws(void* workspace_out, const void* workspace_in, const void* Wgate_f16, const void* Wup_f16, int M, int N, int K,
                           cudaStream_t s) {
    if (M <= 0 || !gemm_f16_tc_supported(Wgate_f16, N, K, 0) || !gemm_f16_tc_supported(Wup_f16, N, K, 0)) return false;
    const size_t Mp = ((size_t)M + BM - 1) / BM * BM;
    return launch_gemm<256, MODE_SWIGLU>(nullptr, static_cast<__half*>(workspace_out), workspace_in, Wgate_f16, Wup_f16, M, Mp, N, K, s);
}

// C[M,N] = A[M,K] . W[N,K]^T ; N % 128 == 0, K % 64 == 0; workspace >= gemm_f16_tc_workspace_bytes(M, K)
bool gemm_f16_tc(float* C, const float* A, const void* W_f16, int M, int N, int K, void* workspace, cudaStream_t s) {
    if (M <= 0 || !gemm_f16_tc_supported(W_f16, N, K, 0)) return false;
    split_activations(workspace, A, M, K, s);
    return gemm_f16_tc_ws(C, workspace, W_f16, M, N, K, false, s);
}

}}  // namespace nt::b200
