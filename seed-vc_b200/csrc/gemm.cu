// svc_gemm: every dense contraction of the hot path (Linear / concat-free Linear / Conv1d taps /
// polyphase ConvTranspose1d) as one segmented GEMM with a fused epilogue.
//
//  * bf16 operands  -> persistent tcgen05 kernel (grid <= 148, n-fastest tile order): UMMA 128 x N x 16
//                      with fp32 accumulators in TMEM (two buffers, so the epilogue of tile i overlaps
//                      the MMAs of tile i+1); operands staged by TMA (cp.async.bulk.tensor, 128B
//                      swizzle) through a 4-8 stage mbarrier ring that runs ahead across tiles;
//                      warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-9 = two
//                      epilogue groups.  Epilogues (template EPI): register / LSU (any pattern),
//                      direct row-layout TMA store (bf16 or fp32 tile, fp32 reduce-add for in-place
//                      residuals), two-output variant; see gemm_tc_kernel.
//  * fp32 operands  -> FFMA shared-memory tiled mainloop ("fp32 mode", small-M conditioning
//                      GEMMs, and a debug cross-check of the tensor-core path).
//
// Reference sites replaced: nn.Linear / nn.Conv1d / nn.ConvTranspose1d calls listed in
// include/seedvc_b200.h next to svc_gemm.
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <type_traits>

#include "epilogue.cuh"

namespace svc {

// ------------------------------------------------------------------------------------------
// tcgen05 path
// ------------------------------------------------------------------------------------------
constexpr int kMaxMaps = 4;
constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row

struct TcSeg {
    int a_map, w_map, shift, w_row0, nkb;
};

struct alignas(64) TcParams {
    CUtensorMap amap[kMaxMaps];
    CUtensorMap wmap[kMaxMaps];
    TcSeg seg[SVC_MAX_SEG];
    int n_seg, total_kb;
    int B, T, tiles_per_batch, n_tiles;
    int ab_f16;                // operands are IEEE half (else bf16)
    // CTA pairs (cluster of 2 along M): both CTAs work on the same N tile of two adjacent M tiles, each TMA-loads
    // HALF of the weight tile and multicasts it to the pair - the weight tile crosses L2 -> SM once per pair.
    int mc, pair_tiles;
    CUtensorMap wmap_h[kMaxMaps];   // weight maps with a box of BN / 2 rows
#ifdef SVC_TRACE
    int dbg;   // trace builds only (SVC_DBG env): 1 = skip epilogue body, 2 = skip MMA issue (timing experiments)
#endif
    // TMA-store epilogue: 0 = off (register/LSU epilogue), 1 = bf16 tile -> out_op,
    // 2 = fp32 tile reduce-added into out_f32 (in-place residual), 3 = fp32 tile -> out_f32
    int store_mode;
    int direct;                // 1: row-layout epilogue writes the swizzled TMA tile directly (no transpose)
    int res_rows;              // direct: residual read in the row layout (one 128 B line per thread)
    int dual;                  // direct: fp32 tile -> out_f32 (omap) AND bf16 tile -> out_op (omap2)
    int res_tma;               // dual: the residual item is TMA-loaded into the staging tile (rmap), not read by LSU
    int epi5;                  // run the two-staging-tile epilogue (EPI 5): dual, or one fp32 output + TMA residual
    int wide;                  // 16-bit-only output in items of 64 columns (EPI 7): omap box is {64, 32}, SWIZZLE_128B
    CUtensorMap rmap;          // (N_out, T, B) view of the residual, same box / swizzle as omap
    CUtensorMap omap2;
    CUtensorMap omap;          // (N_out, T, B) view of the output, box {32, 32, 1}; swizzled when direct
    EpiParams epi;
};

// Work-skipping / path-forcing experiment switches exist only in -DSVC_TRACE builds
// (libseedvc_b200_trace.so, never shipped): the product library does not read the environment.
#ifdef SVC_TRACE
#define SVC_DBG_BITS(p) ((p).dbg)
static bool svc_env_flag(const char* name) { return getenv(name) != nullptr; }
#else
#define SVC_DBG_BITS(p) 0
static constexpr bool svc_env_flag(const char*) { return false; }
#endif

#ifdef SVC_TRACE
__device__ long long g_gemm_trace[2][128][8];   // [0]: epilogue warp 2 lane 0 per item, [1]: MMA thread per tile
#define GTRACE(role, idx, ev)                                                                       \
    do {                                                                                            \
        if (blockIdx.x == 0 && (idx) < 128) g_gemm_trace[role][idx][ev] = clock64();                \
    } while (0)
#else
#define GTRACE(role, idx, ev) do {} while (0)
#endif

constexpr int kEpiWarps = 8;             // two epilogue groups of 4 warps (one per TMEM buffer)
constexpr int kTcThreads = 64 + kEpiWarps * 32;    // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue
constexpr int kStageRowF = 32;           // fp32 row of the per-warp transpose buffer (XOR-swizzled)
// two-output epilogue (EPI 5): per warp two fp32 item tiles (the residual of item k+1 is TMA-loaded into one while
// item k is finished in the other) + one 16-bit tile
constexpr int kDualWarpBytes = 2 * 4096 + 2048;
constexpr int kDualOpOff = 2 * 4096;
// deepest operand ring that fits next to the two-output epilogue's staging tiles
constexpr int epi5_stages(int bn, int stages) {
    const int stage_bytes = (BM + bn) * BK * 2;
    const int fit = (232448 - 8 * kDualWarpBytes - 512 - 1024) / stage_bytes;
    return fit < stages - 1 ? fit : stages - 1;
}

template <int BN, int STAGES, int EPW = 4096>
struct TcSmem {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int EPI_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int EPI_WARP_BYTES = EPW;       // per epilogue warp: fp32 tile (+ bf16 tile when dual)
    static constexpr int EPI_BYTES = kEpiWarps * EPW;
    static constexpr int BAR_OFFSET = EPI_OFFSET + EPI_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 512 + 1024;  // + barriers + alignment slack
    static_assert(TOTAL <= 232448, "shared memory budget");
};

// Epilogue of one 32-row x 32-column accumulator chunk, executed by one warp.
//   prefetch (lane = 4 columns x 1 row per step): residual / accumulate inputs, issued before
//            the accumulator is even read so their latency hides behind phase 1
//   phase 1 (thread = row): bias, per-batch bias, activation (pairs / RoPE need adjacent columns)
//   transpose through an XOR-swizzled smem tile (float4 both ways, conflict free)
//   phase 2 (lane = 4 columns): gate, residual, alpha, accumulate and the stores - every global
//            access is a contiguous row segment (coalesced 16 B per lane).
// bf16-output tensor-core path: approximate transcendentals (1 MUFU each) are far below the
// 2^-9 output rounding; the fp32 SIMT path keeps the precise versions (epilogue.cuh).
__device__ __forceinline__ float fast_rcp(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_tanh(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
#ifdef SVC_SIGMOID_EX2
__device__ __forceinline__ float fast_sigmoid(float x) { return fast_rcp(1.0f + __expf(-x)); }
__device__ __forceinline__ float fast_silu(float x) { return x * fast_sigmoid(x); }
#else
// one MUFU instead of two (ex2 + rcp): sigmoid(x) = 0.5 + 0.5 tanh(x / 2).  The SwiGLU / gate epilogues are bound by
// the MUFU pipe (8 cycles per warp instruction, two epilogue warps per sub-partition); |error| <= 2^-12
__device__ __forceinline__ float fast_sigmoid(float x) { return fmaf(0.5f, fast_tanh(0.5f * x), 0.5f); }
__device__ __forceinline__ float fast_silu(float x) {
    const float h = 0.5f * x;
    return fmaf(h, fast_tanh(h), h);
}
#endif

struct EpiChunk {
    int co, c0;        // outputs per row in this chunk (32, or 16 after a pair activation), first column
    int lanes_per_row; // co / 4
    int rows_per_it;   // 32 / lanes_per_row
    int n_it;          // 32 / rows_per_it
    bool vec;          // vector path usable for this chunk
};

__device__ __forceinline__ EpiChunk epi_chunk_geom(const EpiParams& e, int n0) {
    EpiChunk g;
    const bool pair = e.act == SVC_ACT_SWIGLU_PAIR || e.act == SVC_ACT_TANH_SIG_PAIR;
    g.co = pair ? 16 : 32;
    g.c0 = pair ? (n0 >> 1) : n0;
    g.lanes_per_row = g.co >> 2;
    g.rows_per_it = 32 / g.lanes_per_row;
    g.n_it = 32 / g.rows_per_it;
    // partial last chunk is fine as long as whole float4 column groups are valid (N_out % 4 == 0,
    // implied by vec_ok's N_out % 8 == 0): lanes past N_out simply idle
    g.vec = e.vec_ok && !(e.act == SVC_ACT_ROPE && e.res != nullptr);
    return g;
}

// rr[it] <- raw residual (or, without a residual, the old output when accumulating) for the 4
// columns this lane owns in step `it` (vector path only).  The values are NOT touched here: any
// arithmetic on them would stall the warp until the loads land and defeat the prefetch.
__device__ __forceinline__ void epi_prefetch(const EpiParams& e, const EpiChunk& g, int lane, int b,
                                             int t_base, int T, float4 (&rr)[8]) {
    const int c = g.c0 + (lane % g.lanes_per_row) * 4;
    const int rsub = lane / g.lanes_per_row;
    if (c >= e.N_out) return;
    // one base pointer + a constant row stride per step; rows past T are clamped (their values
    // are never stored) so the loads need no per-row branches
    const float* base;
    long long rstride;
    if (e.act == SVC_ACT_ROPE) {
        if (c >= e.rope_cols) return;
        base = e.rope_tab + static_cast<long long>(e.rope_pos0) * 64 + (c & 63);   // (cos, sin) x 2 pairs
        rstride = 64;
    } else if (e.res != nullptr) {
        base = e.res + static_cast<long long>(b) * e.res_bstride + c;
        rstride = e.res_rstride;
    } else if (e.accumulate) {
        base = e.out_f32 + static_cast<long long>(b) * e.of_bstride + c;
        rstride = e.of_rstride;
    } else {
        return;
    }
    const int tmax = T - 1;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        if (it < g.n_it) {
            const int t = min(t_base + it * g.rows_per_it + rsub, tmax);
            rr[it] = __ldg(reinterpret_cast<const float4*>(base + static_cast<long long>(t) * rstride));
        }
    }
}

__device__ __forceinline__ void epilogue_chunk_coalesced(const EpiParams& e, const EpiChunk& g,
                                                         float* stage, int lane, int b, int t_base,
                                                         int T, int n0, float (&v)[32],
                                                         const float4 (&rr)[8], int dbg) {
    const int t_row = t_base + lane;
    // linear epilogues (no activation, or RoPE) add the biases after the transpose, 4 columns
    // per lane; activations need them first
    const bool bias_late = g.vec && (e.act == SVC_ACT_NONE || e.act == SVC_ACT_ROPE);
    if (dbg & 4) goto transpose;
    if (e.bias != nullptr && !bias_late) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (n0 + j < e.N) v[j] += __ldg(e.bias + n0 + j);
    }
    if (e.rowbias != nullptr && !bias_late) {
        const float* rb = e.rowbias + static_cast<long long>(b) * e.rowbias_bstride;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (n0 + j < e.N) v[j] += __ldg(rb + n0 + j);
    }
    if (e.act == SVC_ACT_SILU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fast_silu(v[j]);
    } else if (e.act == SVC_ACT_SWIGLU_PAIR) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fast_silu(v[2 * j]) * v[2 * j + 1];
    } else if (e.act == SVC_ACT_TANH_SIG_PAIR) {
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = fast_tanh(v[2 * j]) * fast_sigmoid(v[2 * j + 1]);
    } else if (e.act == SVC_ACT_ROPE && !g.vec) {
        if (n0 < e.rope_cols && t_row < T) {
            const float* tab = e.rope_tab + static_cast<long long>(e.rope_pos0 + t_row) * 64;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int i = ((n0 + 2 * j) & 63) >> 1;
                const float2 cs = __ldg(reinterpret_cast<const float2*>(tab) + i);
                const float x0 = v[2 * j], x1 = v[2 * j + 1];
                v[2 * j] = x0 * cs.x - x1 * cs.y;
                v[2 * j + 1] = x1 * cs.x + x0 * cs.y;
            }
        }
        if (n0 < e.q_cols) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= e.q_scale;
        }
    }
transpose:
    // ---- transpose: row-per-thread -> 4-columns-per-lane (slot q of row r lives at q ^ (r & 7))
    if (dbg & 16) return;
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q)
        if (q < g.lanes_per_row)
            *reinterpret_cast<float4*>(stage + lane * kStageRowF + ((q ^ (lane & 7)) << 2)) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    __syncwarp();
    const int q = lane % g.lanes_per_row;
    const int rsub = lane / g.lanes_per_row;
    const int c = g.c0 + q * 4;
    if (dbg & 8) return;
    if (g.vec) {
        if (c >= e.N_out) return;
        float4 gt = make_float4(e.alpha, e.alpha, e.alpha, e.alpha);
        const bool is_rope = e.act == SVC_ACT_ROPE;
        const bool rope = is_rope && c < e.rope_cols;
        const float qs = (is_rope && c < e.q_cols) ? e.q_scale : 1.0f;
        const bool has_rr = !is_rope && (e.res != nullptr || e.accumulate);
        const float rr_scale = e.res != nullptr ? e.alpha : 1.0f;
        const bool rmw = e.res != nullptr && e.accumulate;
        float4 bl = make_float4(0.f, 0.f, 0.f, 0.f);     // late biases for this lane's 4 columns
        if (bias_late) {
            if (e.bias != nullptr) bl = __ldg(reinterpret_cast<const float4*>(e.bias + c));
            if (e.rowbias != nullptr) {
                const float4 rb4 = __ldg(reinterpret_cast<const float4*>(
                    e.rowbias + static_cast<long long>(b) * e.rowbias_bstride + c));
                bl.x += rb4.x, bl.y += rb4.y, bl.z += rb4.z, bl.w += rb4.w;
            }
        }
        if (e.gate != nullptr) {
            const float4 gq = __ldg(reinterpret_cast<const float4*>(
                e.gate + static_cast<long long>(b) * e.gate_bstride + c));
            gt.x *= gq.x, gt.y *= gq.y, gt.z *= gq.z, gt.w *= gq.w;
        }
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = it * g.rows_per_it + rsub;
            const int t = t_base + r;
            if (it < g.n_it && t < T) {
                float4 a = *reinterpret_cast<const float4*>(stage + r * kStageRowF + ((q ^ (r & 7)) << 2));
                a.x += bl.x, a.y += bl.y, a.z += bl.z, a.w += bl.w;
                if (rope) {
                    const float4 cs = rr[it];
                    const float y0 = a.x * cs.x - a.y * cs.y, y1 = a.y * cs.x + a.x * cs.y;
                    const float y2 = a.z * cs.z - a.w * cs.w, y3 = a.w * cs.z + a.z * cs.w;
                    a = make_float4(y0 * qs, y1 * qs, y2 * qs, y3 * qs);
                }
                float4 x = make_float4(a.x * gt.x, a.y * gt.y, a.z * gt.z, a.w * gt.w);
                if (has_rr) {
                    x.x = fmaf(rr[it].x, rr_scale, x.x), x.y = fmaf(rr[it].y, rr_scale, x.y);
                    x.z = fmaf(rr[it].z, rr_scale, x.z), x.w = fmaf(rr[it].w, rr_scale, x.w);
                }
                float* of = e.out_f32 + static_cast<long long>(b) * e.of_bstride +
                            static_cast<long long>(t) * e.of_rstride + c;
                if (rmw) {                       // residual AND accumulate: old value read here
                    const float4 o = *reinterpret_cast<const float4*>(of);
                    x.x += o.x, x.y += o.y, x.z += o.z, x.w += o.w;
                }
                if (e.out_f32 != nullptr) *reinterpret_cast<float4*>(of) = x;
                if (e.out_op != nullptr) {
                    const long long off = static_cast<long long>(b) * e.oo_bstride +
                                          static_cast<long long>(t) * e.oo_rstride + c;
                    if (e.op_is_f32) {
                        *reinterpret_cast<float4*>(static_cast<float*>(e.out_op) + off) = x;
                    } else {
                        uint2 pk;
                        pk.x = pack_op16_rt(x.x, x.y, e.op_is_f16);
                        pk.y = pack_op16_rt(x.z, x.w, e.op_is_f16);
                        *reinterpret_cast<uint2*>(static_cast<uint16_t*>(e.out_op) + off) = pk;
                    }
                }
            }
        }
    } else {
        // scalar fallback (unaligned views or a partial last chunk): still row-contiguous
        for (int it = 0; it < g.n_it; ++it) {
            const int r = it * g.rows_per_it + rsub;
            const int t = t_base + r;
            if (t >= T) break;
            for (int k = 0; k < 4; ++k) {
                const int cc = c + k;
                if (cc >= e.N_out) break;
                float x = stage[r * kStageRowF + ((q ^ (r & 7)) << 2) + k];
                if (e.gate != nullptr) x *= __ldg(e.gate + static_cast<long long>(b) * e.gate_bstride + cc);
                if (e.res != nullptr)
                    x += __ldg(e.res + static_cast<long long>(b) * e.res_bstride +
                               static_cast<long long>(t) * e.res_rstride + cc);
                x *= e.alpha;
                if (e.out_f32 != nullptr) {
                    float* o = e.out_f32 + static_cast<long long>(b) * e.of_bstride +
                               static_cast<long long>(t) * e.of_rstride + cc;
                    if (e.accumulate) x += *o;
                    *o = x;
                }
                if (e.out_op != nullptr) {
                    const long long off = static_cast<long long>(b) * e.oo_bstride +
                                          static_cast<long long>(t) * e.oo_rstride + cc;
                    if (e.op_is_f32) static_cast<float*>(e.out_op)[off] = x;
                    else static_cast<uint16_t*>(e.out_op)[off] = cvt_op16_rt(x, e.op_is_f16);
                }
            }
        }
    }
}

__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tmap)),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const void* tmap, const void* src, int c0, int c1, int c2) {
    asm volatile(
        "cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
            reinterpret_cast<uint64_t>(tmap)),
        "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// TMA-store epilogue of one item = 32 rows x 32 OUTPUT columns (32 accumulator columns, or 64 for
// the pair activations), executed by one warp:
//   phase 1 (thread = row): early biases + activation, fp32 tile into XOR-swizzled smem
//   phase 2 (lane = 4 columns, registers): late biases, RoPE with coalesced prefetched table
//            values, gate, alpha
//   the finished tile goes back to smem in plain row-major output dtype and ONE elected lane
//   hands it to the TMA unit (store, or fp32 reduce-add for the in-place residual update):
//   no per-element address arithmetic, bounds handling by the tensor map, stores fully async.
template <bool PAIR>
__device__ __forceinline__ void epilogue_item_tma(const TcParams& p, float* stage, int lane, int b,
                                                  int t_base, int n0_acc, float (&v)[PAIR ? 64 : 32],
                                                  const float4 (&rr)[8]) {
    const EpiParams& e = p.epi;
    constexpr int NA = PAIR ? 64 : 32;
    const int c0 = PAIR ? (n0_acc >> 1) : n0_acc;          // first output column
    const bool bias_late = e.act == SVC_ACT_NONE || e.act == SVC_ACT_ROPE;
    if (!bias_late) {
        if (e.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < NA; ++j)
                if (n0_acc + j < e.N) v[j] += __ldg(e.bias + n0_acc + j);
        }
        if (e.rowbias != nullptr) {
            const float* rb = e.rowbias + static_cast<long long>(b) * e.rowbias_bstride;
#pragma unroll
            for (int j = 0; j < NA; ++j)
                if (n0_acc + j < e.N) v[j] += __ldg(rb + n0_acc + j);
        }
    }
    if (e.act == SVC_ACT_SILU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fast_silu(v[j]);
    } else if (PAIR && e.act == SVC_ACT_SWIGLU_PAIR) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fast_silu(v[2 * j]) * v[2 * j + 1];
    } else if (PAIR && e.act == SVC_ACT_TANH_SIG_PAIR) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fast_tanh(v[2 * j]) * fast_sigmoid(v[2 * j + 1]);
    }
    // the previous TMA store of this warp must have finished reading the staging tile
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
#pragma unroll
    for (int q = 0; q < 8; ++q)
        *reinterpret_cast<float4*>(stage + lane * kStageRowF + ((q ^ (lane & 7)) << 2)) =
            make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    __syncwarp();
    const int q = lane & 7;
    const int rsub = lane >> 3;
    const int c = c0 + q * 4;
    const bool is_rope = e.act == SVC_ACT_ROPE;
    const bool rope = is_rope && c < e.rope_cols;
    const float qs = (is_rope && c < e.q_cols) ? e.q_scale : 1.0f;
    float4 gt = make_float4(e.alpha, e.alpha, e.alpha, e.alpha);
    float4 bl = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < e.N_out) {
        if (e.gate != nullptr) {
            const float4 gq = __ldg(reinterpret_cast<const float4*>(
                e.gate + static_cast<long long>(b) * e.gate_bstride + c));
            gt.x *= gq.x, gt.y *= gq.y, gt.z *= gq.z, gt.w *= gq.w;
        }
        if (bias_late) {
            if (e.bias != nullptr) bl = __ldg(reinterpret_cast<const float4*>(e.bias + c));
            if (e.rowbias != nullptr) {
                const float4 rb4 = __ldg(reinterpret_cast<const float4*>(
                    e.rowbias + static_cast<long long>(b) * e.rowbias_bstride + c));
                bl.x += rb4.x, bl.y += rb4.y, bl.z += rb4.z, bl.w += rb4.w;
            }
        }
    }
    float4 a[8];
#pragma unroll
    for (int it = 0; it < 8; ++it) {
        const int r = it * 4 + rsub;
        a[it] = *reinterpret_cast<const float4*>(stage + r * kStageRowF + ((q ^ (r & 7)) << 2));
        a[it].x += bl.x, a[it].y += bl.y, a[it].z += bl.z, a[it].w += bl.w;
        if (rope) {
            const float4 cs = rr[it];
            const float y0 = a[it].x * cs.x - a[it].y * cs.y, y1 = a[it].y * cs.x + a[it].x * cs.y;
            const float y2 = a[it].z * cs.z - a[it].w * cs.w, y3 = a[it].w * cs.z + a[it].z * cs.w;
            a[it] = make_float4(y0 * qs, y1 * qs, y2 * qs, y3 * qs);
        }
        a[it].x *= gt.x, a[it].y *= gt.y, a[it].z *= gt.z, a[it].w *= gt.w;
    }
    __syncwarp();          // everyone has read the fp32 tile; overwrite it with the final tile
    if (p.store_mode == 1) {
        uint8_t* sb = reinterpret_cast<uint8_t*>(stage);
        const int f16 = e.op_is_f16;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + rsub;
            uint2 pk;
            pk.x = pack_op16_rt(a[it].x, a[it].y, f16);
            pk.y = pack_op16_rt(a[it].z, a[it].w, f16);
            *reinterpret_cast<uint2*>(sb + r * 64 + q * 8) = pk;
        }
    } else {
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            const int r = it * 4 + rsub;
            *reinterpret_cast<float4*>(stage + r * 32 + q * 4) = a[it];
        }
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        if (p.store_mode == 2) tma_reduce_add_3d(&p.omap, stage, c0, t_base, b);
        else tma_store_3d(&p.omap, stage, c0, t_base, b);
        bulk_commit();
    }
}

// (cos, sin) of the 16 pairs of a 32-column item for this lane's row, from the pair-major table:
// consecutive lanes = consecutive positions = consecutive addresses
__device__ __forceinline__ void rope_prefetch_rows(const EpiParams& e, int c0, int lane, int t_base,
                                                   float4 (&rr)[8]) {
    const int row = t_base + lane;
    const int pos = min(e.rope_pos0 + (e.rope_mod > 0 ? row % e.rope_mod : row), e.rope_ld - 1);
    const float2* tab = reinterpret_cast<const float2*>(e.rope_tab_t) +
                        static_cast<long long>((c0 & 63) >> 1) * e.rope_ld + pos;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float2 a = __ldg(tab + static_cast<long long>(2 * i) * e.rope_ld);
        const float2 b = __ldg(tab + static_cast<long long>(2 * i + 1) * e.rope_ld);
        rr[i] = make_float4(a.x, a.y, b.x, b.y);
    }
}

// residual values of this lane's row for a 32-column item (its own 128 B line; the eight requests of
// a warp touch the same 32 lines, so seven of them hit L1)
__device__ __forceinline__ void res_prefetch_rows(const EpiParams& e, int c0, int lane, int b, int t_base,
                                                  int T, float4 (&rr)[8]) {
    const int t = t_base + lane;
    if (t < T) {
        const float* rp = e.res + static_cast<long long>(b) * e.res_bstride +
                          static_cast<long long>(t) * e.res_rstride + c0;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (c0 + 4 * q < e.N_out) rr[q] = *reinterpret_cast<const float4*>(rp + 4 * q);
    }
}

// Direct TMA-store epilogue: the whole item is finished in the TMEM row layout (thread = output row,
// 32 consecutive output columns in registers; per-column operands are warp-uniform loads) and written
// once into a swizzled shared-memory tile that the TMA store reads (bf16: 64B rows, SWIZZLE_64B; fp32:
// 128B rows, SWIZZLE_128B) - no transpose round trip through shared memory.  Interleaved-pair RoPE
// works here because a pair sits in adjacent registers; its table must be pair-major so that the 32
// rows of a warp read consecutive addresses (rr = 16 (cos, sin) pairs prefetched one item ahead).
//
// EK (epilogue kind) resolves the launch-constant questions at compile time.  An epilogue warp runs alone on its
// scheduler slot, so every warp-uniform branch on a kernel parameter costs its full latency; clock stamps showed
// ~500 of the ~1 500 cycles of a 32-column item going to the dozen "is there a bias / gate / residual / alpha"
// branches.  The kernel picks EK once (epi_kind) and runs the item loop specialised for it:
//   0 generic (every check at run time)      1 no vector operand; act = none or (pair) SwiGLU
//   2 interleaved-pair RoPE, nothing else     3 bias only (+ row-layout residual in the two-output epilogue)
//   4 per-batch bias + tanh * sigmoid pair (WaveNet in-layers)   5 residual only, two outputs (w2 / wo with an operand copy)
//   6 / 7 folded RMS norm: row scale (rs) + bias, then RoPE / SwiGLU pair
constexpr int EK_GENERIC = 0, EK_PLAIN = 1, EK_ROPE = 2, EK_BIAS = 3, EK_ROWBIAS_TS = 4, EK_RES = 5, EK_RS_ROPE = 6,
              EK_RS_SWIGLU = 7;

// rbar / ridx: two-output epilogue with a TMA-loaded residual - the item's residual tile arrives in fp32 staging
// buffer (ridx & 1) under mbarrier rbar[ridx & 1]; the finished fp32 tile is written over it and stored from there.
template <bool PAIR, int EK, bool DUAL>
__device__ __forceinline__ void epilogue_item_direct(const TcParams& p, float* stage, int lane, int b,
                                                     int t_base, int n0_acc, float (&v)[PAIR ? 64 : 32],
                                                     const float4 (&rr)[8], uint64_t* rbar = nullptr,
                                                     uint32_t ridx = 0, float rs = 1.0f) {
    const EpiParams& e = p.epi;
    constexpr int NA = PAIR ? 64 : 32;
    constexpr bool G = EK == EK_GENERIC;
    const int c0 = PAIR ? (n0_acc >> 1) : n0_acc;          // first output column
    constexpr bool RS = EK == EK_RS_ROPE || EK == EK_RS_SWIGLU;
    const bool has_bias = G ? e.bias != nullptr : (EK == EK_BIAS || RS);
    const bool has_rowbias = G ? e.rowbias != nullptr : EK == EK_ROWBIAS_TS;
    const bool has_gate = G ? e.gate != nullptr : false;
    const bool has_res = G ? p.res_rows != 0 : EK == EK_RES ? true : (EK == EK_BIAS && DUAL && p.res_rows != 0);
    const bool res_tma = DUAL && p.res_tma;                 // warp-uniform
    uint8_t* const fbuf = reinterpret_cast<uint8_t*>(stage) + ((DUAL && res_tma) ? (ridx & 1) * 4096 : 0);
    const bool has_alpha = G ? e.alpha != 1.0f : false;
    const int act = G ? e.act
                      : (EK == EK_ROPE || EK == EK_RS_ROPE) ? SVC_ACT_ROPE
                      : EK == EK_ROWBIAS_TS ? SVC_ACT_TANH_SIG_PAIR
                      : ((EK == EK_PLAIN && PAIR) || EK == EK_RS_SWIGLU) ? SVC_ACT_SWIGLU_PAIR : SVC_ACT_NONE;
    if constexpr (RS) {     // folded RMS norm: this row's 1 / rms before the (folded) bias
#pragma unroll
        for (int j = 0; j < NA; ++j) v[j] *= rs;
    }
    if (has_bias) {
#pragma unroll
        for (int j = 0; j < NA; j += 4)
            if (n0_acc + j < e.N) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(e.bias + n0_acc + j));
                v[j] += q.x, v[j + 1] += q.y, v[j + 2] += q.z, v[j + 3] += q.w;
            }
    }
    if (has_rowbias) {
        const float* rb = e.rowbias + static_cast<long long>(b) * e.rowbias_bstride + n0_acc;
#pragma unroll
        for (int j = 0; j < NA; j += 4)
            if (n0_acc + j < e.N) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(rb + j));
                v[j] += q.x, v[j + 1] += q.y, v[j + 2] += q.z, v[j + 3] += q.w;
            }
    }
    if (act == SVC_ACT_SILU) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fast_silu(v[j]);
    } else if (PAIR && act == SVC_ACT_SWIGLU_PAIR) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fast_silu(v[2 * j]) * v[2 * j + 1];
    } else if (PAIR && act == SVC_ACT_TANH_SIG_PAIR) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fast_tanh(v[2 * j]) * fast_sigmoid(v[2 * j + 1]);
    } else if (!PAIR && act == SVC_ACT_ROPE && c0 < e.rope_cols) {
        const float qs = c0 < e.q_cols ? e.q_scale : 1.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 cs = rr[i];           // (cos, sin) of pairs 2i, 2i+1
            const float x0 = v[4 * i], x1 = v[4 * i + 1], x2 = v[4 * i + 2], x3 = v[4 * i + 3];
            v[4 * i] = (x0 * cs.x - x1 * cs.y) * qs;
            v[4 * i + 1] = (x1 * cs.x + x0 * cs.y) * qs;
            v[4 * i + 2] = (x2 * cs.z - x3 * cs.w) * qs;
            v[4 * i + 3] = (x3 * cs.z + x2 * cs.w) * qs;
        }
    }
    if (has_gate) {
        const float* gp = e.gate + static_cast<long long>(b) * e.gate_bstride + c0;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
            if (c0 + j < e.N_out) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(gp + j));
                v[j] *= q.x, v[j + 1] *= q.y, v[j + 2] *= q.z, v[j + 3] *= q.w;
            }
    }
    if (!PAIR && has_res) {
        if (res_tma) {
            mbar_wait_warp(&rbar[ridx & 1], (ridx >> 1) & 1);
            const uint8_t* row = fbuf + lane * 128;
            const int sw = lane & 7;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 r4 = *reinterpret_cast<const float4*>(row + ((q ^ sw) << 4));
                v[4 * q] += r4.x, v[4 * q + 1] += r4.y, v[4 * q + 2] += r4.z, v[4 * q + 3] += r4.w;
            }
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q)
                v[4 * q] += rr[q].x, v[4 * q + 1] += rr[q].y, v[4 * q + 2] += rr[q].z, v[4 * q + 3] += rr[q].w;
        }
    }
    if (has_alpha) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= e.alpha;
    }
    // the previous TMA store of this warp must have finished reading the staging tile
    if (lane == 0) bulk_wait_read0();
    __syncwarp();
    uint8_t* sb = reinterpret_cast<uint8_t*>(stage);
    if (p.store_mode == 1 || (DUAL && p.dual)) {
        uint8_t* row = sb + (DUAL ? kDualOpOff : 0) + lane * 64;
        const int sw = (lane >> 1) & 3;
        if (e.op_is_f16) {         // warp-uniform: one conversion flavour per launch
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(row + ((q ^ sw) << 4)) =
                    make_uint4(pack_f16(v[8 * q], v[8 * q + 1]), pack_f16(v[8 * q + 2], v[8 * q + 3]),
                               pack_f16(v[8 * q + 4], v[8 * q + 5]), pack_f16(v[8 * q + 6], v[8 * q + 7]));
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(row + ((q ^ sw) << 4)) =
                    make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                               pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
        }
    }
    if (p.store_mode != 1) {
        uint8_t* row = fbuf + lane * 128;
        const int sw = lane & 7;
#pragma unroll
        for (int q = 0; q < 8; ++q)
            *reinterpret_cast<float4*>(row + ((q ^ sw) << 4)) =
                make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        if (p.store_mode == 2) tma_reduce_add_3d(&p.omap, fbuf, c0, t_base, b);
        else tma_store_3d(&p.omap, fbuf, c0, t_base, b);
        if (DUAL && p.dual) tma_store_3d(&p.omap2, sb + kDualOpOff, c0, t_base, b);
        bulk_commit();
    }
}

// 16-bit-only output, 64 columns per item (EPI 7): the same row-layout epilogue with HALF as many items per tile.
// The per-item fixed costs (wait for the previous store's smem read, proxy fence, TMA issue, item bookkeeping) are
// what an epilogue warp alone on its scheduler slot spends its time on: the SwiGLU-pair epilogue, which always had
// 64 accumulator columns per item, reached 1 285 TFLOP/s on w13 where the 32-column RoPE / bias / plain items of wqkv
// stayed at 970 - 1 090 with half the FLOPs per tile.  Staging tile: 32 rows x 128 B, SWIZZLE_128B (box {64, 32}).
// RoPE: one item = one head, so the row's 32 (cos, sin) pairs are the same for every item of a tile (L1 hits after
// the first); they are read in two halves right where they are used instead of being carried in registers.
template <int EK>
__device__ __forceinline__ void epilogue_item_wide(const TcParams& p, float* stage, int lane, int b, int t_base,
                                                   int c0, float (&v)[64], int rope_pos) {
    const EpiParams& e = p.epi;
    constexpr bool G = EK == EK_GENERIC;
    const bool has_bias = G ? e.bias != nullptr : EK == EK_BIAS;
    const bool has_rowbias = G ? e.rowbias != nullptr : false;
    const bool has_gate = G ? e.gate != nullptr : false;
    const bool has_alpha = G ? e.alpha != 1.0f : false;
    const int act = G ? e.act : EK == EK_ROPE ? SVC_ACT_ROPE : SVC_ACT_NONE;
    if (has_bias) {
#pragma unroll
        for (int j = 0; j < 64; j += 4)
            if (c0 + j < e.N) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(e.bias + c0 + j));
                v[j] += q.x, v[j + 1] += q.y, v[j + 2] += q.z, v[j + 3] += q.w;
            }
    }
    if (has_rowbias) {
        const float* rb = e.rowbias + static_cast<long long>(b) * e.rowbias_bstride + c0;
#pragma unroll
        for (int j = 0; j < 64; j += 4)
            if (c0 + j < e.N) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(rb + j));
                v[j] += q.x, v[j + 1] += q.y, v[j + 2] += q.z, v[j + 3] += q.w;
            }
    }
    if (act == SVC_ACT_SILU) {
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = fast_silu(v[j]);
    } else if (act == SVC_ACT_ROPE && c0 < e.rope_cols) {
        const float qs = c0 < e.q_cols ? e.q_scale : 1.0f;
        const float2* tab = reinterpret_cast<const float2*>(e.rope_tab_t) + rope_pos;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float2 cs[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) cs[i] = __ldg(tab + static_cast<long long>(16 * h + i) * e.rope_ld);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const float x0 = v[32 * h + 2 * i], x1 = v[32 * h + 2 * i + 1];
                v[32 * h + 2 * i] = (x0 * cs[i].x - x1 * cs[i].y) * qs;
                v[32 * h + 2 * i + 1] = (x1 * cs[i].x + x0 * cs[i].y) * qs;
            }
        }
    }
    if (has_gate) {
        const float* gp = e.gate + static_cast<long long>(b) * e.gate_bstride + c0;
#pragma unroll
        for (int j = 0; j < 64; j += 4)
            if (c0 + j < e.N_out) {
                const float4 q = __ldg(reinterpret_cast<const float4*>(gp + j));
                v[j] *= q.x, v[j + 1] *= q.y, v[j + 2] *= q.z, v[j + 3] *= q.w;
            }
    }
    if (has_alpha) {
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] *= e.alpha;
    }
    if (lane == 0) bulk_wait_read0();       // the previous TMA store of this warp has read the staging tile
    __syncwarp();
    uint8_t* row = reinterpret_cast<uint8_t*>(stage) + lane * 128;
    const int sw = lane & 7;
    if (e.op_is_f16) {
#pragma unroll
        for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(row + ((q ^ sw) << 4)) =
                make_uint4(pack_f16(v[8 * q], v[8 * q + 1]), pack_f16(v[8 * q + 2], v[8 * q + 3]),
                           pack_f16(v[8 * q + 4], v[8 * q + 5]), pack_f16(v[8 * q + 6], v[8 * q + 7]));
    } else {
#pragma unroll
        for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(row + ((q ^ sw) << 4)) =
                make_uint4(pack_bf16(v[8 * q], v[8 * q + 1]), pack_bf16(v[8 * q + 2], v[8 * q + 3]),
                           pack_bf16(v[8 * q + 4], v[8 * q + 5]), pack_bf16(v[8 * q + 6], v[8 * q + 7]));
    }
    fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
        tma_store_3d(&p.omap, stage, c0, t_base, b);
        bulk_commit();
    }
}

// Which specialised item loop a launch can use (warp-uniform, evaluated once per kernel)
__device__ __forceinline__ int epi_kind(const TcParams& p, bool pair, bool dual) {
    const EpiParams& e = p.epi;
    if (e.gate != nullptr || e.alpha != 1.0f) return EK_GENERIC;
    const bool b = e.bias != nullptr, rb = e.rowbias != nullptr, rs = p.res_rows != 0;
    if (!b && !rb && !rs) {
        if (!pair && e.act == SVC_ACT_NONE) return EK_PLAIN;
        if (pair && e.act == SVC_ACT_SWIGLU_PAIR) return EK_PLAIN;
        if (!pair && e.act == SVC_ACT_ROPE) return EK_ROPE;
    }
    if (b && !rb && !pair && e.act == SVC_ACT_NONE && (dual || !rs)) return EK_BIAS;
    if (!b && !rb && rs && dual && !pair && e.act == SVC_ACT_NONE) return EK_RES;
    if (e.row_ss_in != nullptr && b && !rb && !rs) return pair ? EK_RS_SWIGLU : EK_RS_ROPE;   // host-validated
    if (rb && !b && !rs && pair && e.act == SVC_ACT_TANH_SIG_PAIR) return EK_ROWBIAS_TS;
    return EK_GENERIC;
}

// Persistent, warp-specialised tcgen05 GEMM.  One CTA per SM walks output tiles
// (n fastest, so CTAs running together share the A tile in L2); the accumulator is
// double-buffered in TMEM so the epilogue of tile i overlaps the mainloop of tile i+1.
// EPI: 0 = register/LSU epilogue (any pattern), 1 = TMA-store epilogue, 2 = TMA-store epilogue with a
// pair activation (items of 64 accumulator columns).  Separate instantiations keep each path's
// registers and code small.
template <int BN, int STAGES, int EPI>
__global__ void __launch_bounds__(kTcThreads, 1) gemm_tc_kernel(const __grid_constant__ TcParams p) {
    using S = TcSmem<BN, STAGES, EPI == 5 ? kDualWarpBytes : 4096>;
    constexpr int ACC_COLS = BN < 32 ? 32 : BN;
    // accumulator buffers in TMEM: four when they fit (tiles up to 128 columns), else two.  With two, a buffer is busy
    // from the first MMA of a tile to the last tcgen05.ld of its epilogue, so MMA and epilogue of the SAME group take
    // turns; with four the MMA warp runs up to two tiles ahead and both epilogue groups always find a finished tile.
#ifdef SVC_TMEM_NB2
    constexpr int NB = 2;
#else
    constexpr int NB = 4 * ACC_COLS <= 512 ? 4 : 2;
#endif
    constexpr int kTmemCols = NB * ACC_COLS <= 64 ? 64 : NB * ACC_COLS <= 128 ? 128 : NB * ACC_COLS <= 256 ? 256 : 512;   // power of two
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;     // [4] (NB used)
    uint64_t* tmem_empty_bar = tmem_full_bar + 4;     // [4]
    uint64_t* res_bar = tmem_empty_bar + 4;           // [kEpiWarps][2]: TMA-loaded residual items (EPI 5)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 2 * kEpiWarps);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int total_tiles = p.B * p.tiles_per_batch * p.n_tiles;
    const bool mc = p.mc != 0;
    const int crank = mc ? static_cast<int>(blockIdx.x & 1) : 0;        // rank in the CTA pair (1-D grid, cluster 2)
    // tile `it` of this CTA -> (m_tile, n_tile); pairs walk (m pair, n) with n fastest, rank = which M tile
    auto tile_of = [&](int it, int& m_tile, int& n_tile) {
        if (mc) {
            const int pt = static_cast<int>(blockIdx.x >> 1) + it * static_cast<int>(gridDim.x >> 1);
            if (pt >= p.pair_tiles) return false;
            n_tile = pt % p.n_tiles;
            m_tile = 2 * (pt / p.n_tiles) + crank;       // may lie one past the last M tile: computed on zeros, not stored
            return true;
        }
        const int tile = blockIdx.x + it * gridDim.x;
        if (tile >= total_tiles) return false;
        n_tile = tile % p.n_tiles;
        m_tile = tile / p.n_tiles;
        return true;
    };

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < kMaxMaps; ++i) {
            tma_prefetch_desc(&p.amap[i]);
            tma_prefetch_desc(&p.wmap[i]);
        }
        if (p.store_mode != 0) tma_prefetch_desc(&p.omap);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], mc ? 2 : 1);     // pair: the MMA warps of both CTAs release a stage
        }
        for (int a = 0; a < 4; ++a) {
            mbar_init(&tmem_full_bar[a], 1);
            mbar_init(&tmem_empty_bar[a], kEpiWarps / 2);
        }
        for (int a = 0; a < 2 * kEpiWarps; ++a) mbar_init(&res_bar[a], 1);
        if (EPI == 5 && p.res_tma) tma_prefetch_desc(&p.rmap);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    if (mc) cluster_sync_all();        // the peer's barriers are initialised before anything is multicast to them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int m_tile, n_tile;
            for (int it = 0; tile_of(it, m_tile, n_tile); ++it) {
                const int b = m_tile / p.tiles_per_batch;
                const int t0 = (m_tile % p.tiles_per_batch) * BM;
                const int n0 = n_tile * BN;
                for (int s = 0; s < p.n_seg; ++s) {
                    const TcSeg sg = p.seg[s];
                    for (int kb = 0; kb < sg.nkb; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
                        uint8_t* sa = smem + stage * S::STAGE_BYTES;
                        tma_load_3d(sa, &p.amap[sg.a_map], &full_bar[stage], kb * BK, t0 + sg.shift, b);
                        // weight maps are rank 3 too (K, rows, 1): instruction rank == map rank
                        if (mc)     // this CTA's half of the weight tile, to both CTAs of the pair
                            tma_load_3d_mc(sa + S::A_BYTES + crank * (S::B_BYTES / 2), &p.wmap_h[sg.w_map], &full_bar[stage],
                                           kb * BK, sg.w_row0 + n0 + crank * (BN / 2), 0, 3);
                        else
                            tma_load_3d(sa + S::A_BYTES, &p.wmap[sg.w_map], &full_bar[stage], kb * BK,
                                        sg.w_row0 + n0, 0);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            const uint32_t a_lo0 = desc_lo(smem_u32(smem));
            int m_tile, n_tile;
            for (; tile_of(it, m_tile, n_tile); ++it) {
                const int n0 = n_tile * BN;
                int n_umma = p.epi.N - n0;
                n_umma = n_umma > BN ? BN : ((n_umma + 15) & ~15);
                const uint32_t idesc = umma_idesc_bf16(BM, n_umma, 0, 0, p.ab_f16 ? 0u : 1u);
                const int acc = it % NB;
                GTRACE(1, it, 0);
                mbar_wait(&tmem_empty_bar[acc], ((it / NB) & 1) ^ 1);   // epilogue drained this buffer
                GTRACE(1, it, 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                for (int kb = 0; kb < p.total_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t alo = a_lo0 + stage * (S::STAGE_BYTES >> 4);
                    const uint32_t blo = alo + (S::A_BYTES >> 4);
                    if (!(SVC_DBG_BITS(p) & 2)) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k)
                            tc_mma_f16_lh(d_tmem, alo + k * 2, kDescHiSw128, blo + k * 2, kDescHiSw128,
                                          idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if (mc) tc_commit_mc(&empty_bar[stage], 3);     // ... in both CTAs: the peer multicasts into this slot
                    else tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                tc_commit(&tmem_full_bar[acc]);
                GTRACE(1, it, 2);
            }
        }
    } else {
        // ===================== epilogue: group g (4 warps) owns TMEM buffer g =====================
        // Tiles alternate between the two groups, so each group has two mainloop durations to
        // drain its accumulator.  Work items are (tile, 32-column chunk); the residual /
        // accumulate inputs of item k+1 are requested before item k is processed, so global
        // loads stay in flight the whole time.
        const int ew = warp - 2;
        const int group = ew >> 2;
        const int lg = warp & 3;             // TMEM lane group this warp may access
        float* stage_buf = reinterpret_cast<float*>(smem + S::EPI_OFFSET + ew * S::EPI_WARP_BYTES);
        const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(lg * 32) << 16);   // + buffer (tile index % NB) * ACC_COLS
        constexpr bool tma_mode = EPI != 0;
        constexpr bool pair = EPI == 2 || EPI == 4;
        constexpr bool wide = EPI == 7;            // 16-bit-only output, 64 columns per item
        constexpr bool direct = EPI >= 3;          // 3 direct, 4 direct pair, 5 direct with two outputs, 7 direct wide
        constexpr int acc_per_item = (pair || wide) ? 64 : 32;   // accumulator columns per work item
        struct Item {
            int it, ch, b, t_base, n0c, ncols, step, n0;
            bool valid, last;
        };
        // Tiles alternate between the two groups.  (Measured alternative: both groups sharing every tile, group g taking
        // the column chunks g, g + 2, ... so the accumulator buffer is released after half the epilogue time: qkv 283 ->
        // 275 us isolated, but 42.8 vs 42.2 ms per Euler step in situ at the power-capped clock, and ~40 spilled
        // registers in the epilogue warps - not kept.)
        // RoPE items of a full tile are visited even chunks first, then odd ones: chunks 64 columns apart
        // use the same 16 (cos, sin) pairs, so the table is read twice per tile instead of once per chunk
        const bool rope_order = direct && !pair && !wide && p.epi.act == SVC_ACT_ROPE;
        auto chunk_of = [&](int step, int ncols) {
            const int n = ncols / acc_per_item;
            if (!rope_order || ncols != BN || (n & 1)) return step;
            return step < n / 2 ? 2 * step : 2 * (step - n / 2) + 1;
        };
        auto make_item = [&](int it, int ch) {          // full (re)computation: once per tile
            Item x;
            x.it = it;
            x.ch = ch;
            int m_tile = 0, n_tile = 0;
            x.valid = tile_of(it, m_tile, n_tile);
            x.b = 0, x.t_base = 0, x.n0c = 0, x.ncols = 0, x.last = true, x.step = 0, x.n0 = 0;
            if (x.valid) {
                x.b = m_tile / p.tiles_per_batch;
                x.t_base = (m_tile - x.b * p.tiles_per_batch) * BM + lg * 32;
                if (x.b >= p.B) x.b = 0, x.t_base = p.T;      // pair: the M tile past the end is never stored
                const int n0 = n_tile * BN;
                x.n0 = n0;
                x.ncols = min(BN, p.epi.N - n0);
                x.step = 0;
                x.ch = chunk_of(0, x.ncols);
                x.n0c = n0 + x.ch * acc_per_item;
                x.last = acc_per_item >= x.ncols;
            }
            return x;
        };
        auto next_item = [&](const Item& c) {           // within a tile: no divisions
            if (c.last) return make_item(c.it + 2, 0);
            Item x = c;
            x.step = c.step + 1;
            x.ch = chunk_of(x.step, c.ncols);
            x.n0c = c.n0 + x.ch * acc_per_item;
            x.last = (x.step + 1) * acc_per_item >= c.ncols;
            return x;
        };
        // geometry used by the prefetch (4 columns per lane, 8 steps) in TMA mode
        EpiChunk g_tma;
        g_tma.co = 32, g_tma.lanes_per_row = 8, g_tma.rows_per_it = 4, g_tma.n_it = 8, g_tma.vec = true;
        // the item loop, specialised on the epilogue kind (see epilogue_item_direct); EK != 0 only for direct EPIs
        auto run_items = [&](auto ek_tag) {
        constexpr int EK = decltype(ek_tag)::value;
        Item cur = make_item(group, 0);
        EpiChunk g_cur = g_tma;
        if constexpr (tma_mode) g_cur.c0 = pair ? (cur.n0c >> 1) : cur.n0c;
        else g_cur = epi_chunk_geom(p.epi, cur.n0c);
        float4 rr_cur[8], rr_nxt[8];
        const bool want_prefetch = EK == EK_GENERIC && !direct && (!tma_mode || p.epi.act == SVC_ACT_ROPE);
        const bool rope_direct = !wide && (EK == EK_GENERIC ? (direct && p.epi.act == SVC_ACT_ROPE) : (EK == EK_ROPE || EK == EK_RS_ROPE));
        int rope_pos = 0;                                           // wide RoPE: this row's table position, once per tile
        // folded RMS norm: this thread's row scale (consumer) / running sum of squares of its row (producer)
        constexpr bool rs_in = EK == EK_RS_ROPE || EK == EK_RS_SWIGLU;
        const bool ss_out = EPI == 5 && p.epi.row_ss_out != nullptr;
        float rs_row = 1.0f, ss_acc = 0.f;
        if (cur.valid && want_prefetch && g_cur.vec && cur.t_base < p.T)
            epi_prefetch(p.epi, g_cur, lane, cur.b, cur.t_base, p.T, rr_cur);
        if (rope_direct && cur.valid && cur.n0c < p.epi.rope_cols)
            rope_prefetch_rows(p.epi, cur.n0c, lane, cur.t_base, rr_cur);
        const bool res_any = EK == EK_GENERIC ? (direct && !pair && p.res_rows)
                                              : ((EK == EK_BIAS || EK == EK_RES) && EPI == 5 && p.res_rows);
        const bool res_tma = EPI == 5 && res_any && p.res_tma;      // residual item through TMA into the staging tile
        const bool res_direct = res_any && !res_tma;                // ... or one 128-byte line per thread by LSU
        uint64_t* const rbar = res_bar + 2 * ew;
        uint32_t r_issued = 0, r_used = 0;                          // processed items only (t_base < T)
        auto res_issue = [&](const Item& x) {                       // lane 0
            mbar_expect_tx(&rbar[r_issued & 1], 4096);
            tma_load_3d(reinterpret_cast<uint8_t*>(stage_buf) + (r_issued & 1) * 4096, &p.rmap, &rbar[r_issued & 1],
                        x.n0c, x.t_base, x.b);
        };
        if (res_tma && cur.valid && cur.t_base < p.T) {
            if (lane == 0) res_issue(cur);
            ++r_issued;
        }
        if (res_direct && cur.valid) res_prefetch_rows(p.epi, cur.n0c, lane, cur.b, cur.t_base, p.T, rr_cur);
        int tr_i = 0;
        const bool tr_on = (warp == 2 && lane == 0);
        while (cur.valid) {
            if (tr_on) GTRACE(0, tr_i, 0);
            if (cur.step == 0) {
                if constexpr (rs_in) {
                    const int t = cur.t_base + lane;
                    if (t < p.T) {
                        const float4* qp = reinterpret_cast<const float4*>(
                            p.epi.row_ss_in + (static_cast<long long>(cur.b) * p.T + t) * SVC_SS_SLOTS);
                        const float4 q = __ldg(qp), q2 = __ldg(qp + 1);
                        rs_row = rsqrtf((((q.x + q.y) + (q.z + q.w)) + ((q2.x + q2.y) + (q2.z + q2.w))) * p.epi.rs_inv_dim +
                                        p.epi.rs_eps);
                    }
                }
                if constexpr (wide) {
                    const int row = cur.t_base + lane;
                    rope_pos = min(p.epi.rope_pos0 + (p.epi.rope_mod > 0 ? row % p.epi.rope_mod : row), p.epi.rope_ld - 1);
                }
                mbar_wait(&tmem_full_bar[cur.it % NB], (cur.it / NB) & 1);
                tc_fence_after();
            }
            uint32_t r[32], r2[(pair || wide) ? 32 : 1];
            const uint32_t taddr = taddr0 + (cur.it % NB) * ACC_COLS;
            tmem_ld_32x32(taddr + cur.ch * acc_per_item, r);
            if constexpr (pair || wide) tmem_ld_32x32(taddr + cur.ch * 64 + 32, r2);
            const Item nxt = next_item(cur);
            const bool nxt_on = nxt.valid;
            EpiChunk g_nxt = g_tma;
            if constexpr (tma_mode) g_nxt.c0 = pair ? (nxt.n0c >> 1) : nxt.n0c;
            else g_nxt = epi_chunk_geom(p.epi, nxt.n0c);
            if (nxt_on && want_prefetch && g_nxt.vec && nxt.t_base < p.T && !(SVC_DBG_BITS(p) & 32))
                epi_prefetch(p.epi, g_nxt, lane, nxt.b, nxt.t_base, p.T, rr_nxt);
            if (rope_direct && nxt_on && nxt.n0c < p.epi.rope_cols) {
                const bool same = nxt.it == cur.it && cur.n0c < p.epi.rope_cols && ((nxt.n0c ^ cur.n0c) & 63) == 0;
                if (same) {                 // same rows, same pair set: keep the table values
#pragma unroll
                    for (int i = 0; i < 8; ++i) rr_nxt[i] = rr_cur[i];
                } else {
                    rope_prefetch_rows(p.epi, nxt.n0c, lane, nxt.t_base, rr_nxt);
                }
            }
            if (res_direct && nxt_on) res_prefetch_rows(p.epi, nxt.n0c, lane, nxt.b, nxt.t_base, p.T, rr_nxt);
            if (res_tma && nxt_on && nxt.t_base < p.T) {
                if (lane == 0) {
                    bulk_wait_read0();          // the stores of the item before this one have read the target buffer
                    res_issue(nxt);
                }
                ++r_issued;
            }
            if (tr_on) GTRACE(0, tr_i, 1);
            tc_wait_ld();
            if (tr_on) GTRACE(0, tr_i, 2);
            if (cur.last) {                     // this warp has read its whole slice of the buffer
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[cur.it % NB]);
            }
            if (cur.t_base < p.T && !(SVC_DBG_BITS(p) & 1)) {
                if constexpr (EPI == 0) {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    epilogue_chunk_coalesced(p.epi, g_cur, stage_buf, lane, cur.b, cur.t_base, p.T,
                                             cur.n0c, v, rr_cur, SVC_DBG_BITS(p));
                } else if constexpr (wide) {
                    float v[64];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]), v[32 + j] = __uint_as_float(r2[j]);
                    epilogue_item_wide<EK>(p, stage_buf, lane, cur.b, cur.t_base, cur.n0c, v, rope_pos);
                } else if constexpr (pair) {
                    float v[64];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]), v[32 + j] = __uint_as_float(r2[j]);
                    if constexpr (direct) epilogue_item_direct<true, EK, EPI == 5>(p, stage_buf, lane, cur.b, cur.t_base, cur.n0c, v, rr_cur, nullptr, 0, rs_row);
                    else epilogue_item_tma<true>(p, stage_buf, lane, cur.b, cur.t_base, cur.n0c, v, rr_cur);
                } else {
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    if constexpr (direct) epilogue_item_direct<false, EK, EPI == 5>(p, stage_buf, lane, cur.b, cur.t_base, cur.n0c, v, rr_cur, rbar, r_used, rs_row);
                    else epilogue_item_tma<false>(p, stage_buf, lane, cur.b, cur.t_base, cur.n0c, v, rr_cur);
                    if constexpr (EPI == 5) {
                        if (ss_out) {           // v holds the final fp32 values of this row's 32 columns
#pragma unroll
                            for (int j = 0; j < 32; ++j) ss_acc = fmaf(v[j], v[j], ss_acc);
                            if (cur.last) {
                                const int t = cur.t_base + lane;
                                if (t < p.T)
                                    p.epi.row_ss_out[(static_cast<long long>(cur.b) * p.T + t) * SVC_SS_SLOTS + cur.n0 / BN] = ss_acc;
                                ss_acc = 0.f;
                            }
                        }
                    }
                }
                ++r_used;
            }
            if (tr_on) GTRACE(0, tr_i, 3);
            ++tr_i;
            cur = nxt;
            g_cur = g_nxt;
#pragma unroll
            for (int i = 0; i < 8; ++i) rr_cur[i] = rr_nxt[i];
        }
        };   // run_items
        using ek0 = std::integral_constant<int, EK_GENERIC>;
        int ek = EK_GENERIC;
        if constexpr (direct) ek = epi_kind(p, pair, EPI == 5);
        if (EPI != 5 && ek == EK_PLAIN) {
            if constexpr (direct && EPI != 5) run_items(std::integral_constant<int, EK_PLAIN>{});
        } else if ((EPI == 3 || EPI == 7) && ek == EK_ROPE) {
            if constexpr (EPI == 3 || EPI == 7) run_items(std::integral_constant<int, EK_ROPE>{});
        } else if ((EPI == 3 || EPI == 5 || EPI == 7) && ek == EK_BIAS) {
            if constexpr (EPI == 3 || EPI == 5 || EPI == 7) run_items(std::integral_constant<int, EK_BIAS>{});
        } else if (EPI == 4 && ek == EK_ROWBIAS_TS) {
            if constexpr (EPI == 4) run_items(std::integral_constant<int, EK_ROWBIAS_TS>{});
        } else if (EPI == 5 && ek == EK_RES) {
            if constexpr (EPI == 5) run_items(std::integral_constant<int, EK_RES>{});
        } else if (EPI == 3 && ek == EK_RS_ROPE) {
            if constexpr (EPI == 3) run_items(std::integral_constant<int, EK_RS_ROPE>{});
        } else if (EPI == 4 && ek == EK_RS_SWIGLU) {
            if constexpr (EPI == 4) run_items(std::integral_constant<int, EK_RS_SWIGLU>{});
        } else {
            run_items(ek0{});
        }
        if (tma_mode && lane == 0) bulk_wait_read0();   // smem must outlive the last bulk store
    }
    tc_fence_before();
    __syncthreads();
    if (mc) cluster_sync_all();        // no CTA leaves while its peer may still multicast into it
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------
// SIMT path (fp32 FFMA): operands of type T (float or bf16)
// ------------------------------------------------------------------------------------------
struct SimtSeg {
    const void* a;
    long long a_bstride, a_rstride;
    int a_rows, shift;
    const void* w;
    long long w_rstride;
    int K;
};
struct SimtParams {
    SimtSeg seg[SVC_MAX_SEG];
    int n_seg, B, T, tiles_per_batch, n_tiles;
    EpiParams epi;
};

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const __grid_constant__ SimtParams p) {
    constexpr int TM = 64, TN = 64, TK = 16;
    __shared__ float As[TK][TM + 4];
    __shared__ float Ws[TK][TN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int n_tile = blockIdx.x % p.n_tiles;
    const int m_tile = blockIdx.x / p.n_tiles;
    const int b = m_tile / p.tiles_per_batch;
    const int t0 = (m_tile % p.tiles_per_batch) * TM;
    const int n0 = n_tile * TN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int lr = tid >> 2;        // 0..63 : row (A) / col (W) loaded by this thread
    const int lk = (tid & 3) * 4;   // 0,4,8,12
    for (int s = 0; s < p.n_seg; ++s) {
        const SimtSeg sg = p.seg[s];
        const T* A = static_cast<const T*>(sg.a) + static_cast<long long>(b) * sg.a_bstride;
        const T* W = static_cast<const T*>(sg.w);
        const int ar = t0 + lr + sg.shift;
        const bool a_ok = ar >= 0 && ar < sg.a_rows;
        const bool w_ok = (n0 + lr) < p.epi.N;
        for (int k0 = 0; k0 < sg.K; k0 += TK) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + lk + j;
                float av = 0.f, wv = 0.f;
                if (a_ok && k < sg.K) av = to_f32<T>(A[static_cast<long long>(ar) * sg.a_rstride + k]);
                if (w_ok && k < sg.K) wv = to_f32<T>(W[static_cast<long long>(n0 + lr) * sg.w_rstride + k]);
                As[lk + j][lr] = av;
                Ws[lk + j][lr] = wv;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 w4 = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
                const float a[4] = {a4.x, a4.y, a4.z, a4.w};
                const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + ty * 4 + i;
        if (t < p.T && n0 + tx * 4 < p.epi.N) epilogue_chunk<4>(p.epi, b, t, n0 + tx * 4, acc[i]);
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, []() {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
                cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

// (inner = K elements, rows, batches) bf16 view, 128B swizzle, box {64, box_rows, 1}
bool encode_bf16_map(CUtensorMap* map, const void* ptr, int K, long long rows, long long rstride,
                     long long batches, long long bstride, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (fn == nullptr) return false;
    if (reinterpret_cast<uintptr_t>(ptr) % 16 != 0 || (rstride * 2) % 16 != 0) return false;
    if (batches > 1 && (bstride * 2) % 16 != 0) return false;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows),
                          static_cast<cuuint64_t>(batches)};
    if (batches <= 1) bstride = rows * rstride;
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(rstride * 2),
                             static_cast<cuuint64_t>(bstride * 2)};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides,
                    box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// (N_out, T, B) view of an output tensor for the TMA-store epilogue: box {32, 32, 1}, no swizzle
static bool encode_out_map(CUtensorMap* map, const void* ptr, bool f32, int n_out, long long rows,
                           long long rstride, long long batches, long long bstride, bool swizzled, int box_cols = 32) {
    EncodeTiledFn fn = get_encode_fn();
    const int es = f32 ? 4 : 2;
    if (fn == nullptr || ptr == nullptr) return false;
    if (reinterpret_cast<uintptr_t>(ptr) % 16 != 0 || (rstride * es) % 16 != 0) return false;
    if (batches > 1 && (bstride * es) % 16 != 0) return false;
    if (batches <= 1) bstride = rows * rstride;
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(n_out), static_cast<cuuint64_t>(rows),
                          static_cast<cuuint64_t>(batches)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(rstride * es), static_cast<cuuint64_t>(bstride * es)};
    cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), 32, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                    const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    !swizzled ? CU_TENSOR_MAP_SWIZZLE_NONE
                              : (f32 || box_cols == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

template <int BN, int STAGES, int EPI>
static int launch_tc_epi(const TcParams& p, int m_tiles, cudaStream_t stream) {
    using S = TcSmem<BN, STAGES, EPI == 5 ? kDualWarpBytes : 4096>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, EPI>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        attr_set = true;
    }
    const int tiles = m_tiles * p.n_tiles;
    if (p.mc) {         // CTA pairs: one cluster of 2 per pair of M tiles
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(kNumSMs & ~1), cfg.blockDim = dim3(kTcThreads), cfg.dynamicSmemBytes = S::TOTAL, cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2, at[0].val.clusterDim.y = 1, at[0].val.clusterDim.z = 1;
        cfg.attrs = at, cfg.numAttrs = 1;
        // a persistent grid must be co-resident: pairs the GPU can actually host at once (GPCs with an odd number of
        // usable SMs leave one SM without a partner)
        static int max_pairs = 0;
        if (max_pairs == 0) {
            int n = 0;
            if (cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<BN, STAGES, EPI>, &cfg) != cudaSuccess || n < 1) n = 1;
            max_pairs = n;
        }
        const int grid = 2 * (p.pair_tiles < max_pairs ? p.pair_tiles : max_pairs);
        cfg.gridDim = dim3(grid);
#ifdef SVC_TRACE
        static bool said = false;
        if (!said) fprintf(stderr, "seedvc_b200 trace: CTA pairs: max active clusters %d, grid %d\n", max_pairs, grid), said = true;
#endif
        cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, STAGES, EPI>, p);
        if (e != cudaSuccess) {
            svc_set_error(cudaGetErrorString(e));
            return SVC_ERR_CUDA;
        }
        return SVC_OK;
    }
    const int grid = tiles < kNumSMs ? tiles : kNumSMs;
    gemm_tc_kernel<BN, STAGES, EPI><<<grid, kTcThreads, S::TOTAL, stream>>>(p);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

template <int BN, int STAGES>
static int launch_tc(const TcParams& p, int m_tiles, cudaStream_t stream) {
    const bool pair = p.epi.act == SVC_ACT_SWIGLU_PAIR || p.epi.act == SVC_ACT_TANH_SIG_PAIR;
    if (p.store_mode == 0) return launch_tc_epi<BN, STAGES, 0>(p, m_tiles, stream);
    if constexpr (BN >= 64) {
        if (pair) return p.direct ? launch_tc_epi<BN, STAGES, 4>(p, m_tiles, stream)
                                  : launch_tc_epi<BN, STAGES, 2>(p, m_tiles, stream);
    }
    if (p.epi5) return launch_tc_epi<BN, epi5_stages(BN, STAGES), 5>(p, m_tiles, stream);
    if constexpr (BN >= 64) {
        if (p.wide && p.store_mode == 1) return launch_tc_epi<BN, STAGES, 7>(p, m_tiles, stream);
    }
    return p.direct ? launch_tc_epi<BN, STAGES, 3>(p, m_tiles, stream)
                    : launch_tc_epi<BN, STAGES, 1>(p, m_tiles, stream);
}

// Rows of different batch entries can share M tiles when every operand / output is contiguous across the
// batch and nothing in the epilogue is per batch entry: (B, T) -> (1, B*T).  T = 2581 wastes 4 % of every
// 128-row tile grid per batch entry (21 tiles for 20.16); flattened, config 2 runs 1291 tiles instead of 1344.
static bool flatten_batch(svc_gemm_desc& d, int* rope_mod) {
    *rope_mod = 0;
    static const int off = svc_env_flag("SVC_NO_FLATTEN") ? 1 : 0;
    static const bool no_direct_path = svc_env_flag("SVC_NO_DIRECT") || svc_env_flag("SVC_NO_TMA_STORE");
    if (off || d.B <= 1) return false;
    // per-batch-entry vectors are fine when every entry uses the same one (stride 0: v2 AdaLN gates)
    if ((d.rowbias != nullptr && d.rowbias_bstride != 0) || (d.gate != nullptr && d.gate_bstride != 0)) return false;
    const long long T = d.T;
    for (int s = 0; s < d.n_seg; ++s)
        if (d.a_shift[s] != 0 || d.a_rows[s] != d.T || d.a_bstride[s] != T * d.a_rstride[s]) return false;
    if (d.out_f32 != nullptr && d.of_bstride != T * d.of_rstride) return false;
    if (d.out_op != nullptr && d.oo_bstride != T * d.oo_rstride) return false;
    if (d.res != nullptr && d.res_bstride != T * d.res_rstride) return false;
    if (d.act == SVC_ACT_ROPE) {
        // positions restart per batch entry: only the row-layout epilogue knows how (rope_mod); make sure
        // that epilogue will be the one selected
        const bool direct_rope = d.rope_tab_t != nullptr && d.rope_ld > 0 && d.N % 4 == 0 &&
                                 reinterpret_cast<uintptr_t>(d.rope_tab_t) % 8 == 0 && d.out_op != nullptr &&
                                 d.out_f32 == nullptr && d.res == nullptr && !d.accumulate &&
                                 reinterpret_cast<uintptr_t>(d.out_op) % 16 == 0 && (d.oo_rstride * 2) % 16 == 0 &&
                                 !no_direct_path;
        if (!direct_rope) return false;
        *rope_mod = d.T;
    }
    if (T * d.B > 0x7fffffffLL) return false;
    d.T = static_cast<int>(T * d.B);
    d.B = 1;
    for (int s = 0; s < d.n_seg; ++s) d.a_rows[s] = d.T;
    return true;
}

static int gemm_tc(const svc_gemm_desc& d_in, cudaStream_t stream) {
    svc_gemm_desc d = d_in;
    int rope_mod = 0;
    const bool flattened = flatten_batch(d, &rope_mod);
    TcParams p;
    memset(&p, 0, sizeof(p));
    static const bool bn96 = !svc_env_flag("SVC_NO_BN96"), bn192 = !svc_env_flag("SVC_NO_BN192");
    int BN = 128;
    if (d.N <= 32) BN = 32;
    else if (d.N <= 64) BN = 64;
    else if (d.N == 96 && bn96) BN = 96;      // W box of exactly one tap's rows
    else if (d.N <= 128) BN = 128;
    else if ((d.N == 192 || d.N == 384) && bn192) BN = 192;   // exact tiles: no half-empty second tile / W box
    else BN = 256;
#ifdef SVC_PREFER_BN128
    if (BN == 256) {
        int ktot = 0;
        for (int s = 0; s < d.n_seg; ++s) ktot += d.K[s];
        if (ktot <= SVC_PREFER_BN128) BN = 128;
    }
#endif
    // ---- A maps: one per distinct view ---------------------------------------------------
    struct AKey { const void* ptr; long long bs, rs; int rows, K; };
    AKey akeys[kMaxMaps];
    int n_a = 0;
    // ---- W maps: segments whose weights sit at a whole number of rows from a common base
    struct WKey { const char* base; long long rs; int K; long long rows; };
    WKey wkeys[kMaxMaps];
    int n_w = 0;
    int total_kb = 0;
    for (int s = 0; s < d.n_seg; ++s) {
        int ai = -1;
        for (int i = 0; i < n_a; ++i)
            if (akeys[i].ptr == d.a_ptr[s] && akeys[i].bs == d.a_bstride[s] &&
                akeys[i].rs == d.a_rstride[s] && akeys[i].rows == d.a_rows[s] &&
                akeys[i].K == d.K[s])
                ai = i;
        if (ai < 0) {
            if (n_a == kMaxMaps) { svc_set_error("svc_gemm: too many distinct A views"); return SVC_ERR_ARG; }
            akeys[n_a] = {d.a_ptr[s], d.a_bstride[s], d.a_rstride[s], d.a_rows[s], d.K[s]};
            ai = n_a++;
        }
        int wi = -1;
        long long row0 = 0;
        const char* wp = static_cast<const char*>(d.w_ptr[s]);
        for (int i = 0; i < n_w; ++i) {
            const long long rb = wkeys[i].rs * 2;
            // merge only slices that start inside or directly behind the rows this map already covers (taps of
            // one (k, N, K) weight): two separately allocated weights never share a map by address coincidence
            if (wkeys[i].rs == d.w_rstride[s] && wkeys[i].K == d.K[s] && wp >= wkeys[i].base &&
                (wp - wkeys[i].base) % rb == 0 && (wp - wkeys[i].base) / rb <= wkeys[i].rows) {
                wi = i;
                row0 = (wp - wkeys[i].base) / rb;
            }
        }
        if (wi < 0) {
            if (n_w == kMaxMaps) { svc_set_error("svc_gemm: too many distinct W views"); return SVC_ERR_ARG; }
            wkeys[n_w] = {wp, d.w_rstride[s], d.K[s], 0};
            wi = n_w++;
            row0 = 0;
        }
        if (row0 + d.N > wkeys[wi].rows) wkeys[wi].rows = row0 + d.N;
        p.seg[s] = {ai, wi, d.a_shift[s], static_cast<int>(row0), (d.K[s] + BK - 1) / BK};
        total_kb += p.seg[s].nkb;
    }
    for (int i = 0; i < kMaxMaps; ++i) {
        const int ia = i < n_a ? i : 0, iw = i < n_w ? i : 0;
        if (!encode_bf16_map(&p.amap[i], akeys[ia].ptr, akeys[ia].K, akeys[ia].rows, akeys[ia].rs,
                             d.B, akeys[ia].bs, BM) ||
            !encode_bf16_map(&p.wmap[i], wkeys[iw].base, wkeys[iw].K, wkeys[iw].rows, wkeys[iw].rs,
                             1, 0, BN)) {
            svc_set_error("svc_gemm: cuTensorMapEncodeTiled failed (pointer/stride must be 16-byte aligned)");
            return SVC_ERR_ARG;
        }
    }
#ifdef SVC_TRACE
    static const int dbg = getenv("SVC_DBG") ? atoi(getenv("SVC_DBG")) : 0;
    p.dbg = dbg;
#endif
    p.n_seg = d.n_seg;
    p.ab_f16 = d.dtype == SVC_F16;
    p.total_kb = total_kb;
    p.B = d.B;
    p.T = d.T;
    p.tiles_per_batch = (d.T + BM - 1) / BM;
    p.n_tiles = (d.N + BN - 1) / BN;
    p.epi = make_epi_params(d);
    p.epi.rope_mod = rope_mod;
    // ---- TMA-store epilogue when the output pattern allows it ------------------------------
    static const int no_tma_store = svc_env_flag("SVC_NO_TMA_STORE") ? 1 : 0;
    static const int no_direct = svc_env_flag("SVC_NO_DIRECT") ? 1 : 0;
    static const int no_res_tma = svc_env_flag("SVC_NO_RES_TMA") ? 1 : 0;
    p.store_mode = 0;
    // row-layout epilogue (no transpose); RoPE needs the pair-major table
    p.direct = !no_direct && d.N % 4 == 0 &&
               (d.act != SVC_ACT_ROPE || (d.rope_tab_t != nullptr && d.rope_ld > 0 &&
                                          reinterpret_cast<uintptr_t>(d.rope_tab_t) % 8 == 0));
    if (p.epi.vec_ok && !no_tma_store) {
        const bool res_inplace = d.res != nullptr && d.res == d.out_f32 && d.res_bstride == d.of_bstride &&
                                 d.res_rstride == d.of_rstride;
        if (d.out_op != nullptr && d.out_f32 == nullptr && d.res == nullptr && !d.accumulate) {
            const bool pair_a = d.act == SVC_ACT_SWIGLU_PAIR || d.act == SVC_ACT_TANH_SIG_PAIR;
#ifdef SVC_NO_WIDE_ITEMS
            p.wide = 0;
#else
            p.wide = p.direct && !pair_a && BN >= 64 && d.row_ss_in == nullptr && d.N % 8 == 0;
#endif
            if (encode_out_map(&p.omap, d.out_op, false, p.epi.N_out, d.T, d.oo_rstride, d.B, d.oo_bstride, p.direct,
                               p.wide ? 64 : 32))
                p.store_mode = 1;
            else
                p.wide = 0;
        } else if (d.out_f32 != nullptr && d.out_op == nullptr && d.act != SVC_ACT_ROPE) {
            const bool add = (res_inplace && !d.accumulate && d.alpha == 1.0f) ||
                             (d.res == nullptr && d.accumulate);
            const bool plain = d.res == nullptr && !d.accumulate;
            if ((add || plain) &&
                encode_out_map(&p.omap, d.out_f32, true, p.epi.N_out, d.T, d.of_rstride, d.B, d.of_bstride, p.direct))
                p.store_mode = add ? 2 : 3;
        }
    }
    // two outputs (fp32 stream + bf16 operand copy), optionally with a residual read in the row layout
    const bool pair_act = d.act == SVC_ACT_SWIGLU_PAIR || d.act == SVC_ACT_TANH_SIG_PAIR;
    // ... or one fp32 output with a residual that is NOT the output itself (BigVGAN conv2 of an AMP pair, the hoisted
    // merge constants): the residual items come in by TMA, so this beats the transposing LSU epilogue (EPI 0) it used
    // to fall back to
#ifdef SVC_NO_RES_F32_DIRECT
    const bool res_f32_direct = false;
#else
    const bool res_f32_direct = d.res != nullptr && !no_res_tma;
#endif
    if (p.store_mode == 0 && p.direct && p.epi.vec_ok && !no_tma_store && !pair_act &&
        d.act != SVC_ACT_ROPE && d.out_f32 != nullptr && !(d.accumulate && d.out_op != nullptr) &&
        (d.out_op != nullptr || res_f32_direct)) {
        bool ok = encode_out_map(&p.omap, d.out_f32, true, p.epi.N_out, d.T, d.of_rstride, d.B, d.of_bstride, true);
        if (ok && d.out_op != nullptr)
            ok = encode_out_map(&p.omap2, d.out_op, false, p.epi.N_out, d.T, d.oo_rstride, d.B, d.oo_bstride, true);
        if (ok) {
            p.store_mode = d.accumulate ? 2 : 3;
            p.res_rows = d.res != nullptr;
            p.dual = d.out_op != nullptr;
            // residual items by TMA (same box and swizzle as the fp32 output tile they are finished in)
            p.res_tma = d.res != nullptr && !no_res_tma &&
                        encode_out_map(&p.rmap, d.res, true, p.epi.N_out, d.T, d.res_rstride, d.B, d.res_bstride, true);
            p.epi5 = 1;
            if (!p.dual && !p.res_tma) p.store_mode = 0, p.res_rows = 0, p.epi5 = 0;   // nothing gained: transposing epilogue
        }
    }
    if (BN < 64 && pair_act) p.store_mode = 0;
    if (d.row_ss_in != nullptr) {
        const bool ok = d.bias != nullptr && d.rowbias == nullptr && d.gate == nullptr && d.res == nullptr &&
                        !d.accumulate && d.alpha == 1.0f && p.direct && p.store_mode == 1 &&
                        (d.act == SVC_ACT_ROPE || d.act == SVC_ACT_SWIGLU_PAIR) &&
                        reinterpret_cast<uintptr_t>(d.row_ss_in) % 16 == 0;
        if (!ok) {
            svc_set_error("svc_gemm: row_ss_in needs bias + RoPE / SwiGLU-pair -> out_op on the row-layout tensor-core epilogue");
            return SVC_ERR_UNSUPPORTED;
        }
    }
    if (d.row_ss_out != nullptr && !(p.dual && d.N % 32 == 0 && p.n_tiles <= SVC_SS_SLOTS)) {
        svc_set_error("svc_gemm: row_ss_out needs the two-output tensor-core epilogue, N % 32 == 0 and <= 4 N tiles");
        return SVC_ERR_UNSUPPORTED;
    }
    if (flattened && rope_mod > 0 && !(p.direct && p.store_mode == 1)) {
        svc_set_error("svc_gemm: internal - flattened RoPE GEMM did not get the row-layout epilogue");
        return SVC_ERR_ARG;
    }
    const int m_tiles = d.B * p.tiles_per_batch;
#ifdef SVC_TRACE
    if (getenv("SVC_LOG_PATHS") != nullptr) {    // which epilogue every distinct call shape takes (trace builds only)
        static unsigned long long seen[256];
        static int n_seen = 0;
        const unsigned long long key = (static_cast<unsigned long long>(d.N) << 40) ^ (static_cast<unsigned long long>(total_kb) << 24) ^
                                       (p.store_mode << 20) ^ (p.direct << 19) ^ (p.dual << 18) ^ (p.res_rows << 17) ^ (p.res_tma << 16) ^
                                       (d.act << 8) ^ (d.bias != nullptr) ^ ((d.rowbias != nullptr) << 1) ^ ((d.gate != nullptr) << 2) ^
                                       ((d.accumulate != 0) << 3) ^ ((d.alpha != 1.0f) << 4);
        bool found = false;
        for (int i = 0; i < n_seen; ++i) found |= seen[i] == key;
        if (!found && n_seen < 256) {
            seen[n_seen++] = key;
            fprintf(stderr, "svc_gemm path: N %d kblocks %d BN %d rows %lld | store_mode %d direct %d dual %d res_rows %d res_tma %d epi5 %d | act %d bias %d rowbias %d gate %d acc %d alpha %g out_f32 %d out_op %d\n",
                    d.N, total_kb, BN, static_cast<long long>(d.B) * d.T, p.store_mode, p.direct, p.dual, p.res_rows, p.res_tma, p.epi5, d.act,
                    d.bias != nullptr, d.rowbias != nullptr, d.gate != nullptr, d.accumulate, d.alpha, d.out_f32 != nullptr, d.out_op != nullptr);
        }
    }
#endif
    // CTA pairs with the weight tile multicast (wide tiles on the row-layout epilogues, enough M tiles to pair up).
    // Built, parity-green (tests/test_gpu_kernels.py::test_gemm_cta_pairs ran on it) and measured NEUTRAL: qkv 285 vs
    // 278 us, plain 235 vs 240 us with 74 co-resident pairs - the K = 512 mainloop already runs at the measured tensor
    // peak (5 100-5 700 cycles per 128 x 256 x 512 tile = 1.6 PFLOP/s), what is lost is the MMA / epilogue overlap, not
    // L2 -> SM operand bandwidth.  Compiled in, switched on only by -DSVC_MC_PAIRS experiment builds.
#ifdef SVC_MC_PAIRS
    static const int no_mc = svc_env_flag("SVC_NO_MC") ? 1 : 0;
#else
    static const int no_mc = 1;
#endif
    if (!no_mc && BN == 256 && p.direct && p.store_mode != 0 && m_tiles >= 16) {
        bool ok = true;
        for (int i = 0; i < kMaxMaps && ok; ++i) {
            const int iw = i < n_w ? i : 0;
            ok = encode_bf16_map(&p.wmap_h[i], wkeys[iw].base, wkeys[iw].K, wkeys[iw].rows, wkeys[iw].rs, 1, 0, BN / 2);
        }
        if (ok) {
            p.mc = 1;
            p.pair_tiles = ((m_tiles + 1) / 2) * p.n_tiles;
        }
    }
    switch (BN) {
        case 32: return launch_tc<32, 8>(p, m_tiles, stream);
        case 64: return launch_tc<64, 8>(p, m_tiles, stream);
        case 96: return launch_tc<96, 6>(p, m_tiles, stream);
        case 128: return launch_tc<128, 6>(p, m_tiles, stream);
        case 192: return launch_tc<192, 4>(p, m_tiles, stream);
        default: return launch_tc<256, 4>(p, m_tiles, stream);
    }
}

static int gemm_simt(const svc_gemm_desc& d, cudaStream_t stream) {
    if (d.row_ss_in != nullptr || d.row_ss_out != nullptr) {
        svc_set_error("svc_gemm: row_ss_in / row_ss_out exist on the tensor-core path only");
        return SVC_ERR_UNSUPPORTED;
    }
    SimtParams p;
    memset(&p, 0, sizeof(p));
    for (int s = 0; s < d.n_seg; ++s)
        p.seg[s] = {d.a_ptr[s], d.a_bstride[s], d.a_rstride[s], d.a_rows[s], d.a_shift[s],
                    d.w_ptr[s], d.w_rstride[s], d.K[s]};
    p.n_seg = d.n_seg;
    p.B = d.B;
    p.T = d.T;
    p.tiles_per_batch = (d.T + 63) / 64;
    p.n_tiles = (d.N + 63) / 64;
    p.epi = make_epi_params(d);
    const long long blocks = static_cast<long long>(d.B) * p.tiles_per_batch * p.n_tiles;
    if (d.dtype == SVC_F32)
        gemm_simt_kernel<float><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
    else if (d.dtype == SVC_F16)
        gemm_simt_kernel<__half><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
    else
        gemm_simt_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(p);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

}  // namespace svc

extern "C" int svc_gemm(const svc_gemm_desc* d, int backend, void* stream) {
    if (d == nullptr || d->n_seg < 1 || d->n_seg > SVC_MAX_SEG || d->B < 1 || d->T < 1 || d->N < 1) {
        svc_set_error("svc_gemm: bad descriptor");
        return SVC_ERR_ARG;
    }
    if (d->dtype != SVC_BF16 && d->dtype != SVC_F32 && d->dtype != SVC_F16) {
        svc_set_error("svc_gemm: dtype must be SVC_BF16, SVC_F16 or SVC_F32");
        return SVC_ERR_ARG;
    }
    if ((d->act == SVC_ACT_SWIGLU_PAIR || d->act == SVC_ACT_TANH_SIG_PAIR) && (d->N % 2)) {
        svc_set_error("svc_gemm: pair activation needs even N");
        return SVC_ERR_ARG;
    }
    if (d->out_op_dtype_p1 != 0 &&
        (d->dtype == SVC_F32 || (d->out_op_dtype_p1 - 1 != SVC_BF16 && d->out_op_dtype_p1 - 1 != SVC_F16))) {
        svc_set_error("svc_gemm: out_op_dtype_p1 must name a 16-bit type and needs 16-bit operands");
        return SVC_ERR_ARG;
    }
    if (d->out_f32 == nullptr && d->out_op == nullptr) {
        svc_set_error("svc_gemm: no output");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d->dtype != SVC_F32 && backend == SVC_BACKEND_AUTO) return svc::gemm_tc(*d, st);
    return svc::gemm_simt(*d, st);
}

#ifdef SVC_TRACE
extern "C" int svc_debug_gemm_trace(long long* host, int n) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(host, svc::g_gemm_trace, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}
#endif
