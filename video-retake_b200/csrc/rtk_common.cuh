// Shared device helpers for the sm_100a kernels (bf16 rounding primitives, mbarrier / bulk-copy PTX,
// radix-order key transform).  Everything here is header-only and internal to the library.
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdint.h>

#include "../../include/rtk_b200.h"

namespace rtk {

extern long long g_launches;  // host-side launch counter (rtk_launch_count)

#define RTK_CHECK_LAUNCH()                                   \
    do {                                                     \
        ++::rtk::g_launches;                                 \
        cudaError_t e__ = cudaGetLastError();                \
        if (e__ != cudaSuccess) return (int)e__;             \
    } while (0)

constexpr int kWarp = 32;

// NVTX range covering one C-ABI call on the host thread (ranges at the kernel entry points).
// nvtx3 is header-only and binds to the profiler's injection library lazily: without a profiler push / pop are no-ops.
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};
#define RTK_NVTX(name) ::rtk::NvtxRange nvtx_range__(name)

// optional outputs of the streaming cosine kernel (dpselect.cu) used by the MA-LLM compressors (mallm.cu):
// sim[i * si + p * sp] = bf16 cosine(frame i, frame i + 1), nrm[i * si + p * sp] = clamped bf16 norm of frame i
struct DisAux {
    float* sim;
    float* nrm;
    long long si, sp;
};
int dpselect_sim_nrm(const void* x, int64_t T, int64_t N, int64_t C, DisAux aux, cudaStream_t st);

// Layers of one chunk scored by one chain of launches (rtk_pivot_update_batch -> pivot_score.cu).  Per-layer tables travel
// as kernel parameters (CUDA >= 12.1 accepts up to 32 KB of them), so a batch is capped at kMaxBatchLayers layers.
constexpr int kMaxBatchLayers = 32;
struct ScoreBatch {
    int n;                                           // layers
    int64_t H, KVH, L, D;                            // shared by all layers
    const void* q[kMaxBatchLayers];                  // bf16 [H, L, D] views
    const void* k[kMaxBatchLayers];                  // bf16 [KVH, L, D] views
    void* head_scores[kMaxBatchLayers];              // bf16 [KVH, L] each
    int64_t q_stride_h[kMaxBatchLayers], q_stride_l[kMaxBatchLayers], k_stride_h[kMaxBatchLayers], k_stride_l[kMaxBatchLayers];
    // Key elision (optional, all three set or none): key-patch tokens get score 1.0 whatever their column sum is
    // (longvideo_cache.py:272-274), so pass 2 only visits the other keys.  kc[l]: bf16 [n_keys[l], KVH, D] token-major copy of
    // the rows of k[l] that are NOT key patches, in order; slot[l][j]: row of key j in kc[l], -1 for a key patch;
    // n_keys: device array [n].  head_scores of key patches are written as 1.0.
    const void* kc[kMaxBatchLayers];
    const int32_t* slot[kMaxBatchLayers];
    const int32_t* n_keys;
};
int pivot_score_batch(const ScoreBatch& b, void* workspace, size_t workspace_bytes, cudaStream_t st);

// ------------------------------------------------------------------------ programmatic dependent launch (PDL)
// The operators are chains of small dependent kernels on one stream.  Launched with the programmatic-stream-
// serialization attribute a kernel may become resident while its predecessor is still running; every kernel
// therefore starts with pdl_enter(): it lets ITS successor be scheduled (launch_dependents) and then blocks until
// the predecessor grid has completed and its writes are visible (wait).  For a normal launch both are no-ops.
// RTK_NO_PDL=1 in the environment turns the attribute off (A/B and debugging).
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef RTK_PDL_EARLY_SMALL
#define RTK_PDL_EARLY_SMALL 1   // small kernels of the PivotKV / MA-LLM chains let their successor in right away
#endif
#ifndef RTK_PDL_EARLY_SCORE
#define RTK_PDL_EARLY_SCORE 1
#endif
__device__ __forceinline__ void pdl_enter() {
    if (RTK_PDL_EARLY_SMALL) pdl_trigger();
    pdl_wait();
}

bool pdl_enabled();  // pivot_misc.cu

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
#define RTK_LAUNCH_PDL(kern, grid, block, smem, st, ...)                                                  \
    do {                                                                                                  \
        ++::rtk::g_launches;                                                                              \
        cudaError_t le__ = ::rtk::launch_pdl(kern, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__); \
        if (le__ != cudaSuccess) return (int)le__;                                                        \
    } while (0)

// ---------------------------------------------------------------------------------------------- bf16 bits
// A bf16 value widened to fp32 is its 16 bits in the upper half of the word.
__device__ __forceinline__ float bf16lo_to_f32(uint32_t packed) { return __uint_as_float(packed << 16); }
__device__ __forceinline__ float bf16hi_to_f32(uint32_t packed) { return __uint_as_float(packed & 0xffff0000u); }

// cvt.rn.bf16x2.f32 d, a, b : d.hi = bf16(a), d.lo = bf16(b)   (round to nearest even, one instruction)
__device__ __forceinline__ uint32_t pack_bf16x2_rn(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ float round_bf16(float x) {  // fp32 -> bf16 (RN) -> fp32
    return __bfloat162float(__float2bfloat16_rn(x));
}
// exact bf16 product pair: RN(a*b) per half (the fp32 product of two bf16 is exact, so this equals ATen's
// float multiply followed by the bf16 cast)
__device__ __forceinline__ uint32_t mul_bf16x2_rn(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

// Mixed-precision fp32 accumulate straight from the halves of a packed bf16x2 register (sm_100: SASS FHFMA.BF16 /
// FHADD.BF16) - no widening instructions.  acc_lo += lo(p)*lo(p), acc_hi += hi(p)*hi(p)
__device__ __forceinline__ void fma_sq_bf16x2(float& acc_lo, float& acc_hi, uint32_t p) {
    asm("{\n\t.reg .b16 lo, hi;\n\t"
        "mov.b32 {lo, hi}, %2;\n\t"
        "fma.rn.f32.bf16 %0, lo, lo, %0;\n\t"
        "fma.rn.f32.bf16 %1, hi, hi, %1;\n\t}"
        : "+f"(acc_lo), "+f"(acc_hi)
        : "r"(p));
}
// acc_lo += lo(p), acc_hi += hi(p)
__device__ __forceinline__ void add_bf16x2(float& acc_lo, float& acc_hi, uint32_t p) {
    asm("{\n\t.reg .b16 lo, hi;\n\t"
        "mov.b32 {lo, hi}, %2;\n\t"
        "add.rn.f32.bf16 %0, lo, %0;\n\t"
        "add.rn.f32.bf16 %1, hi, %1;\n\t}"
        : "+f"(acc_lo), "+f"(acc_hi)
        : "r"(p));
}
// (lo, hi) of a packed bf16x2 times a scalar, rounded back to bf16x2
__device__ __forceinline__ uint32_t scale_bf16x2_rn(uint32_t p, float s) {
    return pack_bf16x2_rn(bf16lo_to_f32(p) * s, bf16hi_to_f32(p) * s);
}

// Radix order used by ATen's CUDA top-k (SortingRadixSelect.cuh TopKTypeConfig<float>): NaN is the largest
// key, -0 sorts below +0.
__device__ __forceinline__ uint32_t f32_to_ordered(float v) {
    uint32_t x = __float_as_uint(v);
    uint32_t mask = (x & 0x80000000u) ? 0xffffffffu : 0x80000000u;
    return (v == v) ? (x ^ mask) : 0xffffffffu;
}

// ------------------------------------------------------------------------------- mbarrier / bulk async copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// 1-D bulk copy global -> shared (TMA engine, SASS UBLKCP); bytes % 16 == 0, both sides 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// 1-D bulk copy shared -> global.
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src_smem), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

}  // namespace rtk
