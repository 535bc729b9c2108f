// nt_types.h — wire format shared by both sides of the drop-in boundary.
//
// ABI contract (must stay byte-compatible with the reference):
//   * nt::DType numeric values            — reference src/core/types.h:24-35
//   * GGUF block layouts and their sizes  — reference src/core/types.h:96-138
//   * dtype_size / block_size / row_size  — reference src/core/types.h:38-88
// Written from the GGUF/GGML block specification; no reference code is included.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace nt {

enum class DType : uint8_t {
    F32 = 0, F16 = 1, Q8_0 = 2, Q4_0 = 3, Q4_K_M = 4, Q6_K = 5, Q5_K = 6, Q2_K = 7, I32 = 8, COUNT
};

// Bytes per element (plain types) or per quantisation block (quantised types).
constexpr size_t dtype_size(DType dt) {
    switch (dt) {
        case DType::F32: case DType::I32: return 4;
        case DType::F16:    return 2;
        case DType::Q8_0:   return 34;
        case DType::Q4_0:   return 18;
        case DType::Q4_K_M: return 144;
        case DType::Q5_K:   return 176;
        case DType::Q6_K:   return 210;
        case DType::Q2_K:   return 84;
        default:            return 0;
    }
}
// Weights per quantisation block.
constexpr size_t dtype_block_size(DType dt) {
    switch (dt) {
        case DType::Q8_0: case DType::Q4_0: return 32;
        case DType::Q4_K_M: case DType::Q5_K: case DType::Q6_K: case DType::Q2_K: return 256;
        default: return 1;
    }
}
constexpr size_t dtype_row_size(DType dt, size_t n) { return n / dtype_block_size(dt) * dtype_size(dt); }

inline const char* dtype_name(DType dt) {
    switch (dt) {
        case DType::F32: return "F32";     case DType::F16: return "F16";
        case DType::Q8_0: return "Q8_0";   case DType::Q4_0: return "Q4_0";
        case DType::Q4_K_M: return "Q4_K_M"; case DType::Q5_K: return "Q5_K";
        case DType::Q6_K: return "Q6_K";   case DType::Q2_K: return "Q2_K";
        case DType::I32: return "I32";     default: return "UNKNOWN";
    }
}

#pragma pack(push, 1)
struct BlockQ4_0 { uint16_t d; uint8_t qs[16]; };
struct BlockQ8_0 { uint16_t d; int8_t qs[32]; };
struct BlockQ4_K { uint16_t d, dmin; uint8_t scales[12]; uint8_t qs[128]; };
struct BlockQ5_K { uint16_t d, dmin; uint8_t scales[12]; uint8_t qh[32]; uint8_t ql[128]; };
struct BlockQ6_K { uint8_t ql[128]; uint8_t qh[64]; int8_t scales[16]; uint16_t d; };
#pragma pack(pop)
static_assert(sizeof(BlockQ4_0) == 18 && sizeof(BlockQ8_0) == 34, "GGUF block size");
static_assert(sizeof(BlockQ4_K) == 144 && sizeof(BlockQ5_K) == 176 && sizeof(BlockQ6_K) == 210, "GGUF block size");

// GGML tensor-type ids that map onto a DType (reference src/core/types.h:168-217).
enum class GGMLType : uint32_t { F32 = 0, F16 = 1, Q4_0 = 2, Q8_0 = 8, Q2_K = 10, Q4_K = 12, Q5_K = 13, Q6_K = 14, I32 = 26 };
inline DType ggml_to_dtype(uint32_t t) {
    switch (t) {
        case 0: return DType::F32;   case 1: return DType::F16;  case 8: return DType::Q8_0;
        case 2: return DType::Q4_0;  case 12: return DType::Q4_K_M; case 13: return DType::Q5_K;
        case 14: return DType::Q6_K; case 10: return DType::Q2_K; case 26: return DType::I32;
        default: return DType::F32;   // reference falls back to F32 for unknown ids
    }
}

}  // namespace nt

// Host-side invariant failures abort, like the reference's NT_CHECK (src/core/types.h:220-228).
#define NT_CHECK(cond, msg) \
    do { if (!(cond)) { fprintf(stderr, "NT ERROR: %s at %s:%d\n", (msg), __FILE__, __LINE__); abort(); } } while (0)
#define NT_CUDA_CHECK(expr) \
    do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { \
        fprintf(stderr, "CUDA error: %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); abort(); } } while (0)
