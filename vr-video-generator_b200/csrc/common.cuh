// Shared device helpers for the SBS warp kernels (sm_100a only).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/vrsbs.h"

namespace vrsbs {

// ---------------------------------------------------------------------------------------------
// Per-frame record shared by the table builder (writer) and the warp / blur kernels (readers).
// ---------------------------------------------------------------------------------------------
struct FrameTab {
    int32_t  layers;        // L
    int32_t  fill_off;      // off[int(L*3/5)] mod W, in [0,W)
    int32_t  strip;         // columns [0,strip) are restored from the input
    uint32_t status;        // VRSBS_FRAME_* bits
    float    guess_scale;   // layer guess = floor(d*scale + bias); exactness never depends on it
    float    guess_bias;
    int32_t  limit_step;
    int32_t  fill_layer;
    float    depth_max;
    float    pad;
    double   range[2];
    unsigned long long holes;
    uint32_t fast;          // 1: the colour-key fast path tables (blob) are valid for this frame
    uint32_t lut_shift;     // cell = fp16 bits >> lut_shift
    uint32_t lut_cells;     // number of non-negative cells; index lut_cells = "negative" cell
    uint32_t pad2;
};

// ---------------------------------------------------------------------------------------------
// Per-frame "blob" consumed by k_warp_fused (one TMA bulk copy per image row):
//   [BlobHdr 16 B][LayerEnt x (ent_cap+1)][cell LUT, lut_cap bytes]
// LayerEnt e (e = 0..L) describes the two layers a depth value of a cell with LUT value e can
// belong to: layer e-1 (painted iff d < hi) and layer e (painted iff !(d < lo_next)).
// ---------------------------------------------------------------------------------------------
struct BlobHdr {
    int32_t  fill_off;      // fill layer's offset mod W
    uint32_t shift;
    uint32_t ncells;
    uint32_t flags;         // bit 0: fast; bits 8..: L
};
struct LayerEnt {
    uint32_t hi_lo;         // half2: .x (low 16) = hi of layer e-1 (-inf for e = 0), .y = lo of layer e (+inf for e = L)
    uint32_t off4;          // low 16: 4*(signed offset of layer e-1 + key_pad); high 16: same for layer e
};
__host__ __device__ inline uint32_t blob_ent_bytes(int ent_cap, bool f32 = false) {
    return (uint32_t)(((ent_cap + 1) * (f32 ? 16 : 8) + 15) / 16 * 16);
}
__host__ __device__ inline uint32_t blob_bytes(int ent_cap, int lut_cap, bool f32 = false) {
    return 16u + blob_ent_bytes(ent_cap, f32) + (uint32_t)((lut_cap + 15) / 16 * 16);
}
// fp32 depth: 16-byte entries - the two bounds as floats (hi of layer e-1: -inf for e = 0; lo of layer e: +inf for e = L) and
// the same offset pair; BlobHdr.shift = shift | base << 8 with cell = max(fp32 bits >> shift, base) - base
struct LayerEnt32 {
    float hi_prev, lo_cur;
    uint32_t off4, pad;
};
constexpr uint32_t kF32LutFloorBits = 0x3C800000u;   // 2^-6: every smaller non-negative depth shares cell 0

// Range-EMA state that survives between batches (SbsProcessor.last_offset_range).
struct RangeState {
    double range[2];
    int    has_last;
    int    pad;
};

// ---------------------------------------------------------------------------------------------
// fp16 arithmetic exactly as torch applies it: fp32 opmath, one rounding to fp16 per op.
// Intrinsics with explicit rounding are never contracted into FMAs by nvcc.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float h2f(__half h) { return __half2float(h); }
__device__ __forceinline__ __half f2h(float f) { return __float2half_rn(f); }

__device__ __forceinline__ __half smooth3(__half cur, __half p1, __half p2, float w0, float w1, float w2) {
    __half d = f2h(__fmul_rn(h2f(cur), w0));                // depth *= 0.58
    __half t = f2h(__fmul_rn(h2f(p1), w1));                 // list[1] * 0.3
    d = f2h(__fadd_rn(h2f(d), h2f(t)));                     // depth += ...
    t = f2h(__fmul_rn(h2f(p2), w2));                        // list[0] * 0.12
    return f2h(__fadd_rn(h2f(d), h2f(t)));
}

// order-preserving float -> uint32 (so integer atomicMax / redux.max implement a float max)
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t u) {
    uint32_t v = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
#ifdef __CUDA_ARCH__
    return __uint_as_float(v);
#else
    float f; memcpy(&f, &v, 4); return f;
#endif
}

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk-copy (TMA engine, non-tensor form) PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// explicit shared-window loads (keep address arithmetic in 32 bits, no generic->shared conversion per access)
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr)); return v; }
__device__ __forceinline__ uint2 lds_u64(uint32_t addr) { uint2 v; asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr)); return v; }
__device__ __forceinline__ uint4 lds_u128(uint32_t addr) { uint4 v; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr)); return v; }
__device__ __forceinline__ void sts_u32(uint32_t addr, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_zero128(uint32_t addr) { asm volatile("st.shared.v4.u32 [%0], {%1, %1, %1, %1};" :: "r"(addr), "r"(0u) : "memory"); }
__device__ __forceinline__ void reds_or(uint32_t addr, uint32_t v) { asm volatile("red.shared.or.b32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
// mbarrier / bulk-copy wrappers that take 32-bit shared-window addresses
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s_a(uint32_t dst_smem, const void *src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void bulk_s2g_a(void *dst_gmem, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global, tracked by the issuing thread's bulk async-group
__device__ __forceinline__ void bulk_s2g(void *dst_gmem, const void *src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem),
                 "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// make generic-proxy smem writes visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// become resident while its predecessor in the stream is still draining; pdl_wait() blocks until the predecessor grid
// has completed and its writes are visible (a no-op for a normal launch), pdl_launch_dependents() lets the successor's
// CTAs be scheduled as soon as every CTA of this grid has got that far.  Every kernel of the default route executes
// pdl_wait() before its first global read, so completion is transitive along the chain.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__host__ __device__ __forceinline__ int wrap_mod(int v, int n) {
    v %= n;
    return v < 0 ? v + n : v;
}
__host__ __device__ __forceinline__ int reflect_idx(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}
__host__ __device__ __forceinline__ size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace vrsbs
