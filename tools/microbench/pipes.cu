// Pipe-throughput and HBM-read microbenchmark for sm_100a (B200).
// Dev tool: decides the GEMV inner-loop design (which pipe the dequant math
// should sit on). Not part of the product path.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
    printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;
constexpr int NCH = 8;   // independent dependency chains per thread

enum Op { FFMA, FFMA2, LOP3, PRMT, SHF, DP4A, I2F, HFMA2, CVT_H2F, IMAD, FADD2,
          MIX_DP4A_LOP3, MIX_FFMA_LOP3, MIX_FFMA2_PRMT, MIX_DP4A_FFMA, MIX_DP4A_DP4A_LOP,
          MIX_PRMT_FFMA2_2TO1, NOPS };
static const char* names[] = {"FFMA", "FFMA2(f32x2)", "LOP3", "PRMT", "SHF", "DP4A", "I2F.s32",
    "HFMA2", "CVT f16->f32", "IMAD", "FADD2(f32x2)",
    "mix DP4A+LOP3 (1:1)", "mix FFMA+LOP3 (1:1)", "mix FFMA2+PRMT (1:1)", "mix DP4A+FFMA (1:1)",
    "mix 3xDP4A+1xLOP3", "mix 2xPRMT+1xFFMA2"};

template <int OP>
__global__ void __launch_bounds__(1024, 1) pipe_kernel(unsigned* out, unsigned long long* cycles, unsigned seed) {
    unsigned r[NCH];
    float f[NCH];
    unsigned long long d[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        r[i] = seed + threadIdx.x * 17 + i;
        f[i] = 1.0f + (float)i * 1e-3f;
        d[i] = ((unsigned long long)__float_as_uint(1.0f + i) << 32) | __float_as_uint(1.0f);
    }
    unsigned a = seed | 1, b = seed ^ 0x0F0F0F0F;
    float fa = 1.0000001f, fb = 1e-9f;
    unsigned long long da = ((unsigned long long)__float_as_uint(1.0000001f) << 32) | __float_as_uint(0.9999999f);
    __syncthreads();
    unsigned long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
            if (OP == FFMA2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(da));
            if (OP == FADD2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(d[i]) : "l"(da));
            if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (OP == DP4A) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (OP == I2F) { asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f[i]) : "r"(r[i])); asm volatile("mov.b32 %0, %1;" : "=r"(r[i]) : "f"(f[i])); }
            if (OP == HFMA2) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (OP == CVT_H2F) { unsigned short h = (unsigned short)r[i]; asm volatile("cvt.f32.f16 %0, %1;" : "=f"(f[i]) : "h"(h)); asm volatile("mov.b32 %0, %1;" : "=r"(r[i]) : "f"(f[i])); }
            if (OP == IMAD) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
            if (OP == MIX_DP4A_LOP3) {
                if (i & 1) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a), "r"(b));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b));
            }
            if (OP == MIX_FFMA_LOP3) {
                if (i & 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b));
            }
            if (OP == MIX_FFMA2_PRMT) {
                if (i & 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(da));
                else asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
            }
            if (OP == MIX_DP4A_FFMA) {
                if (i & 1) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a), "r"(b));
                else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(fa), "f"(fb));
            }
            if (OP == MIX_DP4A_DP4A_LOP) {
                if ((i & 3) != 3) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(r[i]) : "r"(a), "r"(b));
                else asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b));
            }
            if (OP == MIX_PRMT_FFMA2_2TO1) {
                if ((i % 3) != 2) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));
                else asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(d[i]) : "l"(da));
            }
        }
    }
    unsigned long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) acc ^= r[i] ^ __float_as_uint(f[i]) ^ (unsigned)d[i] ^ (unsigned)(d[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run_pipe(unsigned* out, unsigned long long* cyc, int nsm) {
    pipe_kernel<OP><<<nsm, 1024>>>(out, cyc, 12345u);
    CK(cudaDeviceSynchronize());
    pipe_kernel<OP><<<nsm, 1024>>>(out, cyc, 12345u);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(nsm);
    CK(cudaMemcpy(h.data(), cyc, nsm * 8, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto v : h) avg += (double)v; avg /= nsm;
    double ops = (double)ITERS * NCH * 1024;   // thread-instructions per CTA (=per SM)
    printf("%-28s %8.1f thread-instr/clk/SM  (%.2f warp-instr/clk/SM)\n", names[OP], ops / avg, ops / avg / 32);
}

// ---------------- HBM read bandwidth: LDG.128 streaming ----------------
__global__ void __launch_bounds__(512) read_ldg(const uint4* __restrict__ p, size_t n16, unsigned* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    unsigned acc = 0;
    for (; i + 3 * stride < n16; i += 4 * stride) {
        uint4 v0 = __ldcs(p + i), v1 = __ldcs(p + i + stride), v2 = __ldcs(p + i + 2 * stride), v3 = __ldcs(p + i + 3 * stride);
        acc ^= v0.x ^ v0.y ^ v0.z ^ v0.w ^ v1.x ^ v1.y ^ v1.z ^ v1.w ^ v2.x ^ v2.y ^ v2.z ^ v2.w ^ v3.x ^ v3.y ^ v3.z ^ v3.w;
    }
    for (; i < n16; i += stride) { uint4 v = __ldcs(p + i); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) out[0] = acc;
}

// ---------------- HBM read bandwidth: cp.async.bulk ring per warp ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n .reg .pred p;\n WAIT_LOOP:\n"
        " mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        " @p bra DONE;\n bra WAIT_LOOP;\n DONE:\n}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// Each warp owns STAGES stages of CHUNK bytes; lane 0 is the producer; all lanes consume (LDS.128 + xor).
template <int CHUNK, int STAGES, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) read_bulk(const uint8_t* __restrict__ p, size_t nbytes, unsigned* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint64_t bars[WARPS * STAGES];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* my = smem + (size_t)warp * STAGES * CHUNK;
    uint64_t* mybar = bars + warp * STAGES;
    if (lane == 0) for (int s = 0; s < STAGES; s++) mbar_init(mybar + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const size_t nchunks = nbytes / CHUNK;
    const size_t gw = (size_t)blockIdx.x * WARPS + warp, nw = (size_t)gridDim.x * WARPS;
    // prologue
    size_t next = gw;
    if (lane == 0) {
        for (int s = 0; s < STAGES; s++) {
            if (next < nchunks) { mbar_expect_tx(mybar + s, CHUNK); bulk_g2s(my + s * CHUNK, p + next * CHUNK, CHUNK, mybar + s); }
            next += nw;
        }
    }
    unsigned acc = 0;
    int s = 0; uint32_t phase = 0;
    for (size_t c = gw; c < nchunks; c += nw) {
        mbar_wait(mybar + s, phase);
        const uint4* q = reinterpret_cast<const uint4*>(my + s * CHUNK);
#pragma unroll
        for (int i = lane; i < CHUNK / 16; i += 32) { uint4 v = q[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
        __syncwarp();
        if (lane == 0) {
            size_t nc = c + (size_t)STAGES * nw;
            if (nc < nchunks) { mbar_expect_tx(mybar + s, CHUNK); bulk_g2s(my + s * CHUNK, p + nc * CHUNK, CHUNK, mybar + s); }
        }
        if (++s == STAGES) { s = 0; phase ^= 1; }
    }
    if (acc == 0x12345678u) out[0] = acc;
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(a)); f(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    return best;
}

template <int CHUNK, int STAGES, int WARPS>
void run_bulk(const uint8_t* buf, size_t nbytes, unsigned* out, int nsm, int ctas_per_sm) {
    size_t smem = (size_t)CHUNK * STAGES * WARPS;
    CK(cudaFuncSetAttribute(read_bulk<CHUNK, STAGES, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    float ms = time_ms([&] { read_bulk<CHUNK, STAGES, WARPS><<<nsm * ctas_per_sm, WARPS * 32, smem>>>(buf, nbytes, out); }, 5);
    CK(cudaGetLastError());
    printf("bulk ring chunk=%5d stages=%d warps=%2d ctas/sm=%d smem/cta=%6zu : %8.1f GB/s\n",
           CHUNK, STAGES, WARPS, ctas_per_sm, smem, nbytes / ms / 1e6);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int nsm = prop.multiProcessorCount;
    printf("device %s  SMs %d  smem/SM %zu  L2 %d MB  clock %d kHz\n", prop.name, nsm,
           prop.sharedMemPerMultiprocessor, prop.l2CacheSize >> 20, prop.clockRate);
    unsigned* out; unsigned long long* cyc;
    CK(cudaMalloc(&out, (size_t)nsm * 1024 * 4 * 4)); CK(cudaMalloc(&cyc, nsm * 8));
    run_pipe<FFMA>(out, cyc, nsm); run_pipe<FFMA2>(out, cyc, nsm); run_pipe<FADD2>(out, cyc, nsm);
    run_pipe<LOP3>(out, cyc, nsm); run_pipe<PRMT>(out, cyc, nsm); run_pipe<SHF>(out, cyc, nsm);
    run_pipe<DP4A>(out, cyc, nsm); run_pipe<I2F>(out, cyc, nsm); run_pipe<HFMA2>(out, cyc, nsm);
    run_pipe<CVT_H2F>(out, cyc, nsm); run_pipe<IMAD>(out, cyc, nsm);
    run_pipe<MIX_DP4A_LOP3>(out, cyc, nsm); run_pipe<MIX_FFMA_LOP3>(out, cyc, nsm);
    run_pipe<MIX_FFMA2_PRMT>(out, cyc, nsm); run_pipe<MIX_DP4A_FFMA>(out, cyc, nsm);
    run_pipe<MIX_DP4A_DP4A_LOP>(out, cyc, nsm); run_pipe<MIX_PRMT_FFMA2_2TO1>(out, cyc, nsm);

    // HBM read tests: 8 GiB buffer (>> 126 MB L2)
    size_t nbytes = (size_t)8 << 30;
    uint8_t* buf; CK(cudaMalloc(&buf, nbytes)); CK(cudaMemset(buf, 1, nbytes));
    for (int mult : {4, 8, 16}) {
        float ms = time_ms([&] { read_ldg<<<nsm * mult, 512>>>((const uint4*)buf, nbytes / 16, out); }, 5);
        printf("LDG.128 stream grid=%dxSM x512 : %8.1f GB/s\n", mult, nbytes / ms / 1e6);
    }
    run_bulk<2304, 4, 16>(buf, nbytes, out, nsm, 1);
    run_bulk<2304, 6, 8>(buf, nbytes, out, nsm, 1);
    run_bulk<4608, 3, 16>(buf, nbytes, out, nsm, 1);
    run_bulk<4608, 4, 8>(buf, nbytes, out, nsm, 1);
    run_bulk<4608, 4, 8>(buf, nbytes, out, nsm, 1);
    run_bulk<9216, 2, 8>(buf, nbytes, out, nsm, 1);
    run_bulk<9216, 3, 8>(buf, nbytes, out, nsm, 1);
    run_bulk<1152, 8, 16>(buf, nbytes, out, nsm, 1);
    run_bulk<576, 8, 16>(buf, nbytes, out, nsm, 1);
    run_bulk<2304, 4, 8>(buf, nbytes, out, nsm, 2);
    run_bulk<4608, 2, 8>(buf, nbytes, out, nsm, 2);
    // memcpy D2D reference (read+write)
    uint8_t* buf2; CK(cudaMalloc(&buf2, nbytes / 2));
    float ms = time_ms([&] { CK(cudaMemcpyAsync(buf2, buf, nbytes / 2, cudaMemcpyDeviceToDevice)); }, 5);
    printf("cudaMemcpy D2D (r+w bytes)      : %8.1f GB/s\n", (double)nbytes / ms / 1e6);
    ms = time_ms([&] { CK(cudaMemsetAsync(buf, 0, nbytes)); }, 3);
    printf("cudaMemset (write)              : %8.1f GB/s\n", (double)nbytes / ms / 1e6);
    return 0;
}
