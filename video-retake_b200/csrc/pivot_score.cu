// PivotKV scoring (sm_100a, tcgen05 + TMEM + TMA).  Replaces retake/longvideo_cache.py:260-269:
//   S = bf16(Q K^T);  S = bf16(S * f32(1/sqrt D));  P = bf16(softmax_f32(S));  a = bf16(sum_q P);
//   head_scores = bf16(mean over the G heads of each KV group)
// without ever materialising an L x L tensor.
//
// Column sums of a row-normalised softmax need the row statistics first, so the contraction runs twice:
//   pass 1 (rows = queries)  D = Q_tile . K_tile^T : per query row the running max / sum  -> c_q = m*log2e + log2(l)
//   pass 2 (rows = keys)     D = K_tile . Q_tile^T : p = bf16(exp2(s*log2e - c_q)), accumulated per key row
// In both passes a CTA keeps a PAIR of 128-row "stationary" tiles in shared memory (two slots of two tiles: the next
// unit's pair is prefetched) and streams the other operand through a 2-stage TMA ring; every streamed tile feeds two
// MMAs, one per stationary tile, which halves the L2 -> shared-memory traffic per flop - the kernel runs at the board's
// power cap and that traffic is a quarter of the budget.  Accumulators are 128x128 fp32 tiles in
// TMEM (4 buffers = all 512 columns: two streamed tiles x two stationary tiles).
// Because the softmax side owns whole TMEM lanes, pass 1 reduces along columns inside a thread (no shuffles)
// and pass 2 accumulates the column sums inside a thread as well - hence the transposed second pass.
//
// Warp roles (576 threads): warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer (one elected lane), warps 2-17
// four softmax groups of four warps (one warp per TMEM lane quarter): groups 0, 1 serve the accumulators of the first
// stationary tile, groups 2, 3 those of the second, each taking 64 of the 128 columns with one tcgen05.ld.32x32b.x64.
// Per logit the softmax side issues two in-place fp32->bf16 roundings (F2FP with a zero low half), packed f32x2
// scale / FMA / add, one MUFU.EX2 and (pass 1) half an FMNMX3.
// This file holds the shipped configuration only; the A/B variants of round 1 (ring depths, 16/32-column TMEM loads,
// more softmax groups, lazy rescale ...) live in tests/probes/lab/pivot_score_r1_variants.cu with their results in
// profiles/r1_score_ab_experiments.md, and round 2's FMA-pipe polynomial exp2 (no gain under sustained load) in
// tests/probes/lab/pivot_score_r2_poly.cu / profiles/r2_score_experiments.md.
#include <cuda.h>

#include "rtk_common.cuh"

namespace rtk {

constexpr int kTile = 128;            // rows of both operand tiles, = UMMA M = UMMA N
constexpr int kHeadDim = 128;         // largest D: two 64-element swizzle atoms (D = 64 uses one)
constexpr int kPair = 2;              // stationary tiles per unit: every streamed tile feeds two MMAs
constexpr int kStages = 2;            // streamed-operand ring depth (a stage lasts two MMAs)
constexpr int kASlots = 2;            // stationary-pair slots: the next unit's pair is fetched while this unit computes
constexpr int kAccBufs = 4;           // 128-column TMEM buffers
// pass 2: per-tile c_q rows.  A slot is rewritten for tile c after MMA(c - kStages) was issued, i.e. after tile
// c - kStages - kAccBufs was released (= fully processed) by the softmax side.
constexpr int kStatSlots = kStages + kAccBufs;
constexpr uint32_t kTileBytes = kTile * kHeadDim * 2;        // 32 KiB
constexpr uint32_t kHalfBytes = kTile * 64 * 2;              // one [128][64] swizzle-128B box

struct ScoreSmem {
    // offsets inside dynamic shared memory (1024-byte aligned base)
    static constexpr uint32_t a_tile = 0;                                             // [kASlots][kPair] stationary tiles
    static constexpr uint32_t b_ring = kASlots * kPair * kTileBytes;
    static constexpr uint32_t stats = b_ring + kStages * kTileBytes;                 // [kStatSlots][128] f32
    static constexpr uint32_t merge = stats + kStatSlots * kTile * 4;                // [4][128][2] f32
    static constexpr uint32_t bars = merge + 4 * kTile * 2 * 4;
    // barriers: a_full[2], a_empty[2], b_full[kStages], b_empty[kStages], t_full[4], t_empty[4], st_full[kStatSlots]
    static constexpr uint32_t n_bars = 4 + 2 * kStages + 2 * kAccBufs + kStatSlots;
    static constexpr uint32_t tmem_ptr = bars + n_bars * 8;
    static constexpr uint32_t total = tmem_ptr + 16;
};

// ------------------------------------------------------------------------------------------ tcgen05 wrappers
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]),
          "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
          "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]),
          "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]),
          "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr)
        : "memory");
}
// acc(fp32) += one half of a packed bf16x2 (SASS FHADD.BF16, no unpack needed)
__device__ __forceinline__ void add_bf16_pair(float& acc_lo, float& acc_hi, uint32_t packed) {
    asm("{\n\t.reg .b16 lo, hi;\n\t"
        "mov.b32 {lo, hi}, %2;\n\t"
        "add.rn.f32.bf16 %0, lo, %0;\n\t"
        "add.rn.f32.bf16 %1, hi, %1;\n\t}"
        : "+f"(acc_lo), "+f"(acc_hi)
        : "r"(packed));
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float lg2f(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// K-major, 128-byte-swizzled operand tile: 8-row groups 1024 B apart (SBO), version 1, layout SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3ffffu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// D fp32, A/B bf16, both K-major, M = N = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kTile >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);

// Several layers of one chunk can be scored by the same pair of launches (rtk_pivot_update_batch): the unit space is
// (layer, head, stationary tile pair) and every layer brings its own pair of tensor maps.
template <int NL>
struct ScoreMaps {
    CUtensorMap q[NL], k[NL];
    CUtensorMap kc[NL];               // key elision: the compact copy of the keys that are not key patches (pass 2 only)
};

struct ScoreParams {
    int H, G, L, nt;                  // heads of ALL layers in the launch, heads per KV group, chunk length, tiles = ceil(L / 128)
    int Hl;                           // heads per layer (H = layers * Hl)
    int n_atoms;                      // D / 64
    int q_dim1_is_l, k_dim1_is_l;     // tensor-map coordinate order (dims are sorted by stride on the host)
    float inv_sqrt_d;                 // fp32(1 / fp32(sqrt D))
    const int32_t* n_keys;            // key elision: rows of the compact key copy per layer (device); nullptr = every key
    int kc_dim1_is_l;
    float2* ml_part;                  // [3][H][nt*128]  pass 1 partials (row max in scaled-logit domain, row sum): a unit is
                                      //                 shared by at most three CTAs - first / middle / last part
    float* stats;                     // [H][nt*128]     c_q = m*log2e + log2(l)   (merge kernel output, pass 2 input)
    float* colsum_part;               // [3][H][nt*128]  pass 2 partial fp32 sums over queries of bf16(P)
};

constexpr float kLog2e = 1.4426950408889634f;

// ------------------------------------------------------------------------------- packed fp32x2 helpers (sm_100)
__device__ __forceinline__ uint64_t pk2(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
__device__ __forceinline__ void upk2(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// wait used by the producer and MMA-issuer warps
__device__ __forceinline__ void mbar_wait_feeder(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// fp32 -> nearest bf16, returned as fp32: one F2FP.BF16.F32.PACK_AB whose low half is RZ - the packed result with a
// zero low half IS the rounded fp32 value, so no widening instruction follows
__device__ __forceinline__ float round_bf16_inplace(float a) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(a), "f"(0.f));
    return __uint_as_float(d);
}

// the reference's logit rounding chain on a pair of raw accumulators: bf16(acc) then bf16(x * inv_sqrt_d); fp32 out
__device__ __forceinline__ uint64_t logit_chain2(uint32_t r0, uint32_t r1, uint64_t inv2) {
    float t0, t1;
    upk2(mul2(pk2(round_bf16_inplace(__uint_as_float(r0)), round_bf16_inplace(__uint_as_float(r1))), inv2), t0, t1);
    return pk2(round_bf16_inplace(t0), round_bf16_inplace(t1));
}
__device__ __forceinline__ float logit_chain1(float acc, float inv) { return round_bf16(round_bf16(acc) * inv); }

// exp2 of a packed pair: two MUFU.EX2
__device__ __forceinline__ uint64_t ex2_mufu2(uint64_t x2) {
    float x0, x1;
    upk2(x2, x0, x1);
    return pk2(ex2f(x0), ex2f(x1));
}

struct SoftmaxState {
    float m = -INFINITY;                                     // pass 1: running max (scaled-logit domain)
    uint64_t acc = 0;                                        // pass 1: packed fp32x2 row sum
    float c0 = 0.f, c1 = 0.f;                                // pass 2: column sums
};

// 64 consecutive accumulator columns of one TMEM lane (one row of the stationary tile)
template <int PASS>
__device__ __forceinline__ void softmax_cols(uint32_t (&r)[64], int valid, SoftmaxState& st, const float* cq, float inv,
                                             uint64_t inv2, uint64_t l2e2) {
    constexpr int NC = 64;
    if (PASS == 1) {
        if (valid < NC) {
#pragma unroll
            for (int i = 0; i < NC; ++i)
                if (i >= valid) r[i] = 0xff800000u;          // -inf: padded key column
        }
        // the rounding chain is monotone, so the row max of the rounded logits is the chain of the raw max
        float mx = fmaxf(fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1])), __uint_as_float(r[2]));
#pragma unroll
        for (int i = 3; i < NC - 1; i += 2) mx = fmaxf(fmaxf(mx, __uint_as_float(r[i])), __uint_as_float(r[i + 1]));
        mx = fmaxf(mx, __uint_as_float(r[NC - 1]));
        const float mn = fmaxf(st.m, logit_chain1(mx, inv));
        if (!(mn > -INFINITY)) return;
        const float mm = mn * kLog2e;
        const float sc = ex2f(fmaf(st.m, kLog2e, -mm));      // m == -inf -> 0
        st.acc = mul2(st.acc, pk2(sc, sc));
        st.m = mn;
        const uint64_t nmm = pk2(-mm, -mm);
#pragma unroll
        for (int i = 0; i < NC; i += 4) {
            st.acc = add2(st.acc, ex2_mufu2(fma2(logit_chain2(r[i], r[i + 1], inv2), l2e2, nmm)));
            st.acc = add2(st.acc, ex2_mufu2(fma2(logit_chain2(r[i + 2], r[i + 3], inv2), l2e2, nmm)));
        }
    } else {
#pragma unroll
        for (int i = 0; i < NC; i += 4) {
            const float4 cc = *reinterpret_cast<const float4*>(cq + i);
            float e0, e1, e2, e3;
            upk2(ex2_mufu2(fma2(logit_chain2(r[i], r[i + 1], inv2), l2e2, pk2(-cc.x, -cc.y))), e0, e1);
            upk2(ex2_mufu2(fma2(logit_chain2(r[i + 2], r[i + 3], inv2), l2e2, pk2(-cc.z, -cc.w))), e2, e3);
            add_bf16_pair(st.c0, st.c1, pack_bf16x2_rn(e0, e1));
            add_bf16_pair(st.c0, st.c1, pack_bf16x2_rn(e2, e3));
        }
    }
}

// One CTA's share of the flattened (unit, streamed tile) space of a layer: a contiguous range, so every SM gets the same
// number of tile-steps (+-1).  The host sizes the grid so that a range is at least half a unit long (floor(Gl / grid) >=
// ceil(nt / 2)): a unit is then shared by at most THREE CTAs - the one with its first tiles, possibly a middle one, the one
// with its last tiles - each writing its own partial plane.  (Launches with fewer units than SMs - a rank of the KV-head
// split holds 7 heads = 112 units - still fill the machine that way.)  With several layers in the launch the CTA takes ITS
// range of every layer in turn - the cut points inside a layer are those of a single-layer launch, so batched and
// per-layer scoring fold their fp32 partials in the same order and agree bit for bit.
struct TileRange {
    // 32-bit arithmetic: Gl = heads per layer x units per head x streamed tiles <= 256 * 64 * 128 = 2^21 for L <= 16384, and
    // Gl * (grid + 2) stays below 2^31 (checked on the host)
    int g, g1, Gl;
    int nt;                           // streamed tiles of a unit
    int nt_a, nta;                    // stationary tiles / units per head of the current layer
    int layer, n_layers, Hl, geff;
    const int32_t* n_keys;            // pass 2 with key elision: stationary rows per layer
    __device__ __forceinline__ void set_layer_range() {
        nt_a = n_keys ? (n_keys[layer] + kTile - 1) / kTile : nt;
        nta = (nt_a + kPair - 1) / kPair;
        Gl = Hl * nta * nt;
        // CTAs that share the layer: all of them unless that would make a range shorter than half a unit (the host sizes the
        // grid for the full key count; with most keys elided fewer CTAs take part)
        int ge = Gl / ((nt + 1) / 2);
        ge = ge < 1 ? 1 : (ge > (int)gridDim.x ? (int)gridDim.x : ge);
        geff = ge;
        if ((int)blockIdx.x < geff) {
            g = (int)((unsigned)Gl * blockIdx.x / (unsigned)geff);
            g1 = (int)((unsigned)Gl * (blockIdx.x + 1) / (unsigned)geff);
        } else {
            g = g1 = 0;
        }
    }
    // does the NEXT CTA's range of the current layer reach the end of unit `ul` (then the unit has no middle part)?
    __device__ __forceinline__ bool next_cta_reaches_end_of(int ul) const {
        return (int)((unsigned)Gl * (blockIdx.x + 2) / (unsigned)geff) >= (ul + 1) * nt;
    }
    // H: heads of all layers, Hl: heads per layer; a head has ceil(nt_a / kPair) units of kPair stationary tiles
    __device__ __forceinline__ TileRange(int H, int Hl_, int nt_, const int32_t* n_keys_) : nt(nt_), layer(0), Hl(Hl_), n_keys(n_keys_) {
        n_layers = H / Hl_;
        set_layer_range();
    }
    // ul: unit index inside the current layer (= head-of-layer * nta + stationary tile pair); the layer is `layer`
    __device__ __forceinline__ bool next(int& ul, int& tb0, int& tb1) {
        while (g >= g1) {
            if (++layer >= n_layers) return false;
            set_layer_range();
        }
        ul = g / nt;
        tb0 = g - ul * nt;
        const int left = g1 - g;
        tb1 = (left < nt - tb0) ? tb0 + left : nt;
        g += tb1 - tb0;
        return true;
    }
};

constexpr int kGroups = 4;                          // softmax groups of four warps (one warp per TMEM lane quarter)
constexpr int kSoftmaxWarps = 4 * kGroups;
constexpr int kScoreThreads2 = (2 + kSoftmaxWarps) * 32;
constexpr int kTileArrivals = 8;                    // softmax-warp arrivals that free one accumulator buffer (two groups)

template <int PASS, int NL>
__global__ void __launch_bounds__(kScoreThreads2, 1)
pivot_score_kernel(const __grid_constant__ ScoreMaps<NL> maps, ScoreParams prm) {
    if (RTK_PDL_EARLY_SCORE) pdl_trigger();
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // dynamic smem base is only guaranteed 16-byte aligned: round up to 1024 for the swizzle atoms
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t bar0 = base + ScoreSmem::bars;
    auto a_full = [&](int i) { return bar0 + 8 * i; };
    auto a_empty = [&](int i) { return bar0 + 16 + 8 * i; };
    auto b_full = [&](int s) { return bar0 + 32 + 8 * s; };
    auto b_empty = [&](int s) { return bar0 + 32 + 8 * (kStages + s); };
    auto t_full = [&](int b) { return bar0 + 32 + 8 * (2 * kStages + b); };
    auto t_empty = [&](int b) { return bar0 + 32 + 8 * (2 * kStages + kAccBufs + b); };
    auto st_full = [&](int i) { return bar0 + 32 + 8 * (2 * kStages + 2 * kAccBufs + i); };

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(a_full(i), 1); mbar_init(a_empty(i), 1); }
        for (int s = 0; s < kStages; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        for (int b = 0; b < kAccBufs; ++b) { mbar_init(t_full(b), 1); mbar_init(t_empty(b), kTileArrivals); }
        for (int i = 0; i < kStatSlots; ++i) mbar_init(st_full(i), 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(base + ScoreSmem::tmem_ptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                                              // everything above overlapped the previous kernel's tail
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(smem + ScoreSmem::tmem_ptr);

    const int nt = prm.nt;
    const size_t hl = (size_t)prm.H * nt * kTile;            // elements of one [H][Lpad] plane
    // the stationary operand is Q in pass 1 and K in pass 2
    // (pass 2 with key elision: the stationary rows are the compact copy of the keys that are not key patches)
    const bool elide = (PASS == 2) && prm.n_keys != nullptr;
    const CUtensorMap* a_maps = (PASS == 1) ? maps.q : (elide ? maps.kc : maps.k);
    const CUtensorMap* b_maps = (PASS == 1) ? maps.k : maps.q;
    const int a_l1 = (PASS == 1) ? prm.q_dim1_is_l : (elide ? prm.kc_dim1_is_l : prm.k_dim1_is_l);
    const int b_l1 = (PASS == 1) ? prm.k_dim1_is_l : prm.q_dim1_is_l;
    TileRange range(prm.H, prm.Hl, nt, elide ? prm.n_keys : nullptr);
    int ul, tb0, tb1;

    if (warp == 0) {
        // ===================================================================== TMA producer
        if (lane == 0) {
            uint32_t cnt = 0, ucnt = 0;
            while (range.next(ul, tb0, tb1)) {
                const int layer = range.layer;
                const int h = ul / range.nta, ta = (ul - h * range.nta) * kPair;      // head of the layer; first stationary tile
                const int hh = layer * prm.Hl + h;                                      // head index over all layers of the launch
                const int nt_a = range.nt_a;
                const CUtensorMap* a_map = a_maps + layer;
                const CUtensorMap* b_map = b_maps + layer;
                const int a_head = (PASS == 1) ? h : h / prm.G;
                const int b_head = (PASS == 1) ? h / prm.G : h;
                // stationary pair of this unit into slot ucnt % kASlots: it is on its way while the MMAs of the
                // previous unit are still running (no pipeline bubble at unit boundaries)
                const int as = ucnt % kASlots;
                mbar_wait_feeder(a_empty(as), ((ucnt / kASlots) & 1u) ^ 1u);
                mbar_arrive_expect_tx(a_full(as), kPair * prm.n_atoms * kHalfBytes);
                for (int pa = 0; pa < kPair; ++pa)
                    for (int kk = 0; kk < prm.n_atoms; ++kk) {
                        // (an odd tile count leaves the last unit one tile short: its second slot repeats the first, the
                        //  softmax side drops that result)
                        const int row = ((ta + pa < nt_a) ? ta + pa : ta) * kTile;
                        tma_load_3d(base + ScoreSmem::a_tile + (as * kPair + pa) * kTileBytes + kk * kHalfBytes, a_map, kk * 64,
                                    a_l1 ? row : a_head, a_l1 ? a_head : row, a_full(as));
                    }
                ++ucnt;
                for (int tb = tb0; tb < tb1; ++tb, ++cnt) {
                    const int s = cnt % kStages;
                    mbar_wait_feeder(b_empty(s), ((cnt / kStages) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(b_full(s), prm.n_atoms * kHalfBytes);
                    const uint32_t dst = base + ScoreSmem::b_ring + s * kTileBytes;
                    for (int kk = 0; kk < prm.n_atoms; ++kk) {
                        const int row = tb * kTile;
                        tma_load_3d(dst + kk * kHalfBytes, b_map, kk * 64, b_l1 ? row : b_head, b_l1 ? b_head : row,
                                    b_full(s));
                    }
                    if (PASS == 2) {
                        // c_q of the 128 streamed queries; slot cnt % kStatSlots was last read for tile cnt - kStatSlots,
                        // which the softmax side finished before MMA(cnt - kStages) could start, i.e. before b_empty(s) fired
                        const int sl = cnt % kStatSlots;
                        mbar_arrive_expect_tx(st_full(sl), kTile * 4u);
                        bulk_g2s(base + ScoreSmem::stats + sl * kTile * 4u,
                                 prm.stats + (size_t)hh * nt * kTile + (size_t)tb * kTile, kTile * 4u, st_full(sl));
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ======================================================================= MMA issuer
        uint32_t cnt = 0, ucnt = 0;
        while (range.next(ul, tb0, tb1)) {
            const int as = ucnt % kASlots;
            mbar_wait_feeder(a_full(as), (ucnt / kASlots) & 1u);
            ++ucnt;
            // streamed tile cnt feeds both stationary tiles: accumulators 2*(cnt&1) (tile 0) and 2*(cnt&1)+1 (tile 1)
            for (int tb = tb0; tb < tb1; ++tb, ++cnt) {
                const int s = cnt % kStages, bp = 2 * (int)(cnt & 1u);
                mbar_wait_feeder(b_full(s), (cnt / kStages) & 1u);
                mbar_wait_feeder(t_empty(bp), ((cnt >> 1) & 1u) ^ 1u);
                mbar_wait_feeder(t_empty(bp + 1), ((cnt >> 1) & 1u) ^ 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t b0 = base + ScoreSmem::b_ring + s * kTileBytes;
                    const int nks = prm.n_atoms * 4;
#pragma unroll
                    for (int pa = 0; pa < 2; ++pa) {
                        const uint32_t a0 = base + ScoreSmem::a_tile + (as * 2 + pa) * kTileBytes;
#pragma unroll 8
                        for (int ks = 0; ks < nks; ++ks) {
                            const uint32_t off = (ks >> 2) * kHalfBytes + (ks & 3) * 32;
                            umma_bf16(tmem_base + (bp + pa) * kTile, umma_desc_sw128(a0 + off), umma_desc_sw128(b0 + off), kIdesc,
                                      ks > 0 ? 1u : 0u);
                        }
                        tc_commit(t_full(bp + pa));
                    }
                    tc_commit(b_empty(s));
                    if (tb == tb1 - 1) tc_commit(a_empty(as));
                }
                __syncwarp();
            }
        }
    } else {
        // ====== softmax: 4 groups x 4 warps; groups 0, 1 serve the unit's first stationary tile, groups 2, 3 its second
        //        one, each taking one 64-column half of every accumulator
        const int sw = warp - 2;                 // 0..15
        const int grp = sw >> 2;                 // 0..3
        const int quarter = warp & 3;            // TMEM lane quarter this warp may touch
        const int row = quarter * 32 + lane;     // row of the stationary tile
        const int half = grp & 1;                // which 64 columns of the tile
        const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + half * 64;
        float* merge = reinterpret_cast<float*>(smem + ScoreSmem::merge);        // [4][128][2]
        const float inv = prm.inv_sqrt_d;
        const uint64_t inv2 = pk2(inv, inv), l2e2 = pk2(kLog2e, kLog2e);
        uint32_t cnt = 0;
        while (range.next(ul, tb0, tb1)) {
            const int hl_ = ul / range.nta, ta = (ul - hl_ * range.nta) * 2 + (grp >> 1);
            const int h = range.layer * prm.Hl + hl_;          // head index over all layers of the launch
            const bool has_tile = ta < range.nt_a;   // odd tile count: the last unit's second tile is a repeat, dropped below
            SoftmaxState st;
            for (int tb = tb0; tb < tb1; ++tb, ++cnt) {
                const int b = 2 * (int)(cnt & 1u) + (grp >> 1);
                mbar_wait(t_full(b), (cnt >> 1) & 1u);
                tc_fence_after();
                if (PASS == 2) mbar_wait(st_full(cnt % kStatSlots), (cnt / kStatSlots) & 1u);
                const int valid = prm.L - tb * kTile - half * 64;      // streamed rows of this half that exist
                const float* cq = reinterpret_cast<const float*>(smem + ScoreSmem::stats) + (cnt % kStatSlots) * kTile + half * 64;
                if (has_tile) {              // (the spare slot of an odd tile count holds a repeat: nothing to read there)
                    uint32_t r[64];
                    tmem_ld64(lane_addr + b * kTile, r);
                    tmem_ld_wait();
                    softmax_cols<PASS>(r, valid, st, cq, inv, inv2, l2e2);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(t_empty(b));
            }
            const float m = st.m;
            // ---- fold the two column halves and write this CTA's share of the unit
            float a0, a1;
            upk2(st.acc, a0, a1);
            if (PASS == 2) { a0 = st.c0; a1 = st.c1; }
            // this CTA's part of the unit: 0 = has its first tiles, 2 = has its last tiles (and not the first), 1 = in between.
            // The CTA with the first tiles also writes the neutral element into the planes nobody else will write.
            const bool first = (tb0 == 0), last = (tb1 == nt);
            const int part = first ? 0 : (last ? 2 : 1);
            const bool fill1 = first && (last || range.next_cta_reaches_end_of(ul)), fill2 = first && last;
            const size_t o = (size_t)h * nt * kTile + (size_t)ta * kTile + row;
            const int g_lo = grp & ~1;               // the two groups of this stationary tile
            const bool folder = (grp == g_lo) && has_tile;
            if (PASS == 1) {
                merge[(grp * kTile + row) * 2] = m;
                merge[(grp * kTile + row) * 2 + 1] = a0 + a1;
                asm volatile("bar.sync 1, %0;" ::"n"(kSoftmaxWarps * 32) : "memory");
                if (folder) {
                    float mn = -INFINITY;
                    for (int g = g_lo; g < g_lo + 2; ++g) mn = fmaxf(mn, merge[(g * kTile + row) * 2]);
                    float lt = 0.f;
                    for (int g = g_lo; g < g_lo + 2; ++g) {
                        const float mg = merge[(g * kTile + row) * 2];
                        if (mg > -INFINITY) lt += merge[(g * kTile + row) * 2 + 1] * ex2f((mg - mn) * kLog2e);
                    }
                    prm.ml_part[(size_t)part * hl + o] = make_float2(mn, lt);
                    if (fill1) prm.ml_part[hl + o] = make_float2(-INFINITY, 0.f);
                    if (fill2) prm.ml_part[2 * hl + o] = make_float2(-INFINITY, 0.f);
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kSoftmaxWarps * 32) : "memory");
            } else {
                merge[grp * kTile + row] = a0 + a1;
                asm volatile("bar.sync 1, %0;" ::"n"(kSoftmaxWarps * 32) : "memory");
                if (folder) {
                    const float cs = merge[g_lo * kTile + row] + merge[(g_lo + 1) * kTile + row];
                    prm.colsum_part[(size_t)part * hl + o] = cs;
                    if (fill1) prm.colsum_part[hl + o] = 0.f;
                    if (fill2) prm.colsum_part[2 * hl + o] = 0.f;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(kSoftmaxWarps * 32) : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// pass 1 epilogue: fold the (at most three) partial row statistics of every query into c_q = m*log2e + log2(l)
__global__ void pivot_stats_merge_kernel(const float2* __restrict__ ml_part, int H, int L, int Lpad, float* __restrict__ stats) {
    pdl_enter();
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    const int h = blockIdx.y;
    if (q >= Lpad) return;
    const size_t o = (size_t)h * Lpad + q, hl = (size_t)H * Lpad;
    const float2 p0 = ml_part[o], p1 = ml_part[hl + o], p2 = ml_part[2 * hl + o];
    const float mn = fmaxf(fmaxf(p0.x, p1.x), p2.x);
    float lt = 0.f;
    if (p0.x > -INFINITY) lt += p0.y * ex2f((p0.x - mn) * kLog2e);
    if (p1.x > -INFINITY) lt += p1.y * ex2f((p1.x - mn) * kLog2e);
    if (p2.x > -INFINITY) lt += p2.y * ex2f((p2.x - mn) * kLog2e);
    stats[o] = (q < L) ? fmaf(mn, kLog2e, lg2f(lt)) : INFINITY;
}

// a = bf16(colsum);  head_scores[g] = bf16((sum over the G heads of group g, ATen 4-accumulator order) * f32(1/G)).
// blockIdx.y runs over the KV heads of all layers of the launch; every layer has its own [KVH, L] output.
struct HeadScoreOut {
    __nv_bfloat16* p[kMaxBatchLayers];
    const int32_t* slot[kMaxBatchLayers];   // key elision: column of key k in the partial planes, -1 = key patch (score 1.0); or nullptr
    int KVH;                          // KV heads per layer
};

__global__ void pivot_head_reduce_kernel(const float* __restrict__ colsum_part, int H, int G, int L, int Lpad,
                                         const __grid_constant__ HeadScoreOut out) {
    pdl_enter();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y;
    if (k >= L) return;
    const size_t hl = (size_t)H * Lpad;
    const int layer = g / out.KVH;
    int col = k;
    if (out.slot[layer]) {
        col = out.slot[layer][k];
        if (col < 0) {                // a key patch: its score is 1.0 by definition (longvideo_cache.py:272-274), pass 2 skipped it
            out.p[layer][(size_t)(g - layer * out.KVH) * L + k] = __float2bfloat16_rn(1.0f);
            return;
        }
    }
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    for (int j = 0; j < G; ++j) {
        const size_t o = (size_t)(g * G + j) * Lpad + col;
        v[j & 3] += round_bf16((colsum_part[o] + colsum_part[hl + o]) + colsum_part[2 * hl + o]);
    }
    const float s = ((v[0] + v[1]) + v[2]) + v[3];
    out.p[layer][(size_t)(g - layer * out.KVH) * L + k] = __float2bfloat16_rn(s * (1.0f / (float)G));
}

// ---------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// [heads, L, 128] bf16 view with element strides (stride_h, stride_l, 1) -> 3-D map, box = 64 x 128 rows x 1 head.
// Dimensions 1 and 2 are ordered by ascending stride; *dim1_is_l tells the kernel which coordinate is the row.
static int make_map(CUtensorMap* map, const void* ptr, int64_t heads, int64_t L, int64_t D, int64_t stride_h,
                    int64_t stride_l, int* dim1_is_l) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) return RTK_E_DRIVER;
    const bool l_first = (heads == 1) || (stride_l <= stride_h);
    *dim1_is_l = l_first ? 1 : 0;
    cuuint64_t dims[3] = {(cuuint64_t)D, (cuuint64_t)(l_first ? L : heads), (cuuint64_t)(l_first ? heads : L)};
    cuuint64_t strides[2] = {(cuuint64_t)((l_first ? stride_l : stride_h) * 2), (cuuint64_t)((l_first ? stride_h : stride_l) * 2)};
    if (heads == 1) strides[1] = (cuuint64_t)(stride_l * 2) * (cuuint64_t)L;      // never stepped; keep it well-formed
    cuuint32_t box[3] = {64u, (cuuint32_t)(l_first ? kTile : 1), (cuuint32_t)(l_first ? 1 : kTile)};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : RTK_E_DRIVER;
}

}  // namespace rtk

using namespace rtk;

extern "C" size_t rtk_pivot_score_workspace_bytes(int64_t H, int64_t L) {
    if (H < 1 || L < 1) return 0;
    const size_t lpad = (size_t)((L + kTile - 1) / kTile) * kTile;
    return 10 * (size_t)H * lpad * sizeof(float);      // ml_part (3 x float2) + stats + colsum_part (3)
}

namespace rtk {

template <int NL>
static int score_launch(const ScoreBatch& b, ScoreParams prm, cudaStream_t st) {
    ScoreMaps<NL> maps;
    for (int l = 0; l < b.n; ++l) {
        int ql = 0, kl = 0;
        int rc = make_map(&maps.q[l], b.q[l], b.H, b.L, b.D, b.q_stride_h[l], b.q_stride_l[l], &ql);
        if (rc) return rc;
        rc = make_map(&maps.k[l], b.k[l], b.KVH, b.L, b.D, b.k_stride_h[l], b.k_stride_l[l], &kl);
        if (rc) return rc;
        if (l == 0) { prm.q_dim1_is_l = ql; prm.k_dim1_is_l = kl; }
        else if (ql != prm.q_dim1_is_l || kl != prm.k_dim1_is_l) return RTK_E_UNSUPPORTED;   // layers must share a layout
        if (prm.n_keys) {
            int cl = 0;                                             // compact key copy: token-major [rows, KVH, D]
            rc = make_map(&maps.kc[l], b.kc[l], b.KVH, b.L, b.D, b.D, b.KVH * b.D, &cl);
            if (rc) return rc;
            prm.kc_dim1_is_l = cl;
        } else {
            maps.kc[l] = maps.k[l];
        }
    }
    for (int l = b.n; l < NL; ++l) { maps.q[l] = maps.q[0]; maps.k[l] = maps.k[0]; maps.kc[l] = maps.kc[0]; }
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int units = prm.Hl * ((prm.nt + kPair - 1) / kPair);   // per layer: the grid (and so every cut point) is that of one layer
    // a CTA's range of a layer must be at least half a unit long (at most three CTAs per unit, see TileRange)
    const long long gmax = (long long)units * prm.nt / ((prm.nt + 1) / 2);
    const int grid = (int)(gmax < sms ? (gmax < 1 ? 1 : gmax) : sms);
    const size_t smem = ScoreSmem::total + 1024;
    cudaError_t e = cudaFuncSetAttribute(pivot_score_kernel<1, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    e = cudaFuncSetAttribute(pivot_score_kernel<2, NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    HeadScoreOut out;
    out.KVH = (int)b.KVH;
    for (int l = 0; l < kMaxBatchLayers; ++l) {
        out.p[l] = reinterpret_cast<__nv_bfloat16*>(b.head_scores[l < b.n ? l : 0]);
        out.slot[l] = prm.n_keys ? b.slot[l < b.n ? l : 0] : nullptr;
    }
    RTK_LAUNCH_PDL((pivot_score_kernel<1, NL>), grid, kScoreThreads2, smem, st, maps, prm);
    dim3 g1((unsigned)((prm.nt * kTile + 255) / 256), (unsigned)prm.H);
    RTK_LAUNCH_PDL(pivot_stats_merge_kernel, g1, 256, 0, st, prm.ml_part, prm.H, prm.L, prm.nt * kTile, prm.stats);
    RTK_LAUNCH_PDL((pivot_score_kernel<2, NL>), grid, kScoreThreads2, smem, st, maps, prm);
    dim3 g2((unsigned)((prm.L + 255) / 256), (unsigned)(b.KVH * b.n));
    RTK_LAUNCH_PDL(pivot_head_reduce_kernel, g2, 256, 0, st, prm.colsum_part, prm.H, prm.G, prm.L, prm.nt * kTile, out);
    return 0;
}

// scoring of b.n layers of one chunk (same H, KVH, L, D) by one chain of four launches
int pivot_score_batch(const ScoreBatch& b, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (b.n < 1 || b.n > kMaxBatchLayers || !workspace || b.H < 1 || b.KVH < 1 || b.L < 1) return RTK_E_BADARG;
    if ((b.D != 64 && b.D != 128) || b.H % b.KVH != 0 || b.L > 16384 || b.H > 256) return RTK_E_UNSUPPORTED;   // (TileRange: 32-bit)
    if (((uintptr_t)workspace & 15u) != 0) return RTK_E_ALIGN;
    for (int l = 0; l < b.n; ++l) {
        if (!b.q[l] || !b.k[l] || !b.head_scores[l]) return RTK_E_BADARG;
        if ((((uintptr_t)b.q[l] | (uintptr_t)b.k[l]) & 15u) != 0) return RTK_E_ALIGN;
        if ((b.q_stride_h[l] | b.q_stride_l[l] | b.k_stride_h[l] | b.k_stride_l[l]) % 8 != 0) return RTK_E_ALIGN;
    }
    if (workspace_bytes < rtk_pivot_score_workspace_bytes(b.H * b.n, b.L)) return RTK_E_WORKSPACE;

    ScoreParams prm;
    prm.H = (int)(b.H * b.n);
    prm.Hl = (int)b.H;
    prm.G = (int)(b.H / b.KVH);
    prm.L = (int)b.L;
    prm.nt = (int)((b.L + kTile - 1) / kTile);
    prm.n_atoms = (int)(b.D / 64);
    prm.q_dim1_is_l = prm.k_dim1_is_l = prm.kc_dim1_is_l = 0;
    prm.n_keys = b.n_keys;
    if (b.n_keys)
        for (int l = 0; l < b.n; ++l)
            if (!b.kc[l] || !b.slot[l] || ((uintptr_t)b.kc[l] & 15u) != 0) return RTK_E_BADARG;
    const float sq = (float)sqrt((double)b.D);
    prm.inv_sqrt_d = 1.0f / sq;
    const size_t plane = (size_t)prm.H * prm.nt * kTile;
    prm.ml_part = reinterpret_cast<float2*>(workspace);
    prm.stats = reinterpret_cast<float*>(workspace) + 6 * plane;
    prm.colsum_part = prm.stats + plane;
    return b.n == 1 ? score_launch<1>(b, prm, st) : score_launch<kMaxBatchLayers>(b, prm, st);
}

}  // namespace rtk

extern "C" int rtk_pivot_score(const void* q, int64_t H, int64_t q_stride_h, int64_t q_stride_l, const void* k, int64_t KVH,
                               int64_t k_stride_h, int64_t k_stride_l, int64_t L, int64_t D, void* head_scores,
                               void* workspace, size_t workspace_bytes, void* stream) {
    RTK_NVTX("rtk_pivot_score");
    if (!q || !k || !head_scores || !workspace || H < 1 || KVH < 1 || L < 1) return RTK_E_BADARG;
    ScoreBatch b = {};
    b.n = 1; b.H = H; b.KVH = KVH; b.L = L; b.D = D;
    b.q[0] = q; b.k[0] = k; b.head_scores[0] = head_scores;
    b.q_stride_h[0] = q_stride_h; b.q_stride_l[0] = q_stride_l; b.k_stride_h[0] = k_stride_h; b.k_stride_l[0] = k_stride_l;
    return pivot_score_batch(b, workspace, workspace_bytes, (cudaStream_t)stream);
}
