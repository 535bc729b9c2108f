// decode_mega_sched.h — the GEMV phase schedule of the persistent decode kernel: which rows of which matrix the stage
// (round, segment) of warp slot (cta, gsub) holds.  One definition, used by the device producer and consumer
// (decode_megakernel.cu) and by the host-side plan builder / schedule replay (decode_mega_plan.cu).
#pragma once
#include "decode_mega.h"
#include "../gemv_kq_device.cuh"

namespace nt { namespace b200 {
namespace {

constexpr size_t MEGA_STATIC_SMEM = 4096;                                   // upper bound of the kernel's static __shared__
constexpr size_t MEGA_DYN_SMEM = 227 * 1024 - MEGA_STATIC_SMEM;             // TMA rings; aliased by the attention scratch
constexpr int MEGA_NTHREADS = MEGA_WARPS * 32;

// Identical schedule to gemv_kq_kernel: in round r CTA b handles row-groups (r * grid + b) * gpc + [0, gpc); warp w holds
// chunk (w % NC) of row-group slot (w / NC).  Flattened stage index = round * n_seg + seg.
struct Producer {
    int issued;        // stages issued so far in the phase (uniform across the warp)
    int round, seg;    // round and segment of the next stage to fetch
};

__host__ __device__ __forceinline__ int blk_bytes(int fmt) { return fmt == 1 ? 176 : fmt == 2 ? 210 : fmt == 3 ? 272 : 144; }   // 0 Q4_K, 4 Q4_0: 144

__host__ __device__ __forceinline__ void locate(const MegaPhase& d, int g, int seg, int& mi, int& gl) {
    if (d.n_seg == 2) { mi = seg; gl = g; return; }
    mi = 0;
    while (mi + 1 < d.n_mat && g >= d.mat[mi].groups) { g -= d.mat[mi].groups; mi++; }
    gl = g;
}

// What the stage (round, seg) of warp slot (cta, gsub) holds: nrows rows, starting at row0, of chunk `chunk` of matrix mi.
// The one function both the device producer, the device consumer and the host-side schedule check derive the schedule from.
//   full rounds:  row-group (round * grid + cta) * gpc + gsub, all RG rows;
//   tail round (MEGA_SPLIT_TAIL, tail_nr < RG): the remaining tail_groups row-groups are dealt out tail_nr rows at a time over
//   the slots in slot order, so a partly filled last round costs tail_nr / RG of a round instead of a whole one.
struct StageRef {
    bool empty;            // nothing to do: the stage only completes its mbarrier phase
    int mi, gl, row0, nrows;
    int blkb;              // bytes per 256 weights of the matrix's format
    uint32_t bytes;        // bytes copied per row (multiple of 16)
    long long src_off;     // offset of row row0's part inside the matrix
};
__host__ __device__ __forceinline__ StageRef stage_ref(const MegaPhase& d, int grid, int cta, int gsub, int round, int seg, int chunk,
                                                       int nbc) {
    StageRef r;
    r.empty = true; r.mi = 0; r.gl = 0; r.row0 = 0; r.nrows = RG; r.blkb = 0; r.bytes = 0; r.src_off = 0;
    int g, row_sub = 0, nr = RG;
    if (round < d.full_rounds || d.tail_nr >= RG) {
        g = (round * grid + cta) * d.gpc + gsub;
        if (g >= d.total_groups) return r;
    } else {
        nr = d.tail_nr;
        const int per = RG / nr, u = cta * d.gpc + gsub;
        if (u >= d.tail_groups * per) return r;
        g = d.full_rounds * grid * d.gpc + u / per;
        row_sub = u % per;
    }
    locate(d, g, seg, r.mi, r.gl);
    const MegaMat& m = d.mat[r.mi];
    r.row0 = r.gl * RG + row_sub * nr;
    if (r.row0 >= m.out) return r;                           // a slice of a ragged last group that holds no row
    r.empty = false;
    r.nrows = nr;
    r.blkb = blk_bytes(m.fmt);
    r.bytes = ((uint32_t)(nbc * r.blkb) + 15u) & ~15u;       // a 210-byte tail may spill into row padding (host-checked)
    r.src_off = (long long)chunk * (BS * r.blkb) + (long long)r.row0 * m.pitch;
    return r;
}
__host__ __device__ __forceinline__ void producer_advance(const MegaPhase& d, Producer& pr) {
    if (++pr.seg == d.n_seg) { pr.seg = 0; pr.round++; }
}

}  // namespace
}}  // namespace nt::b200
