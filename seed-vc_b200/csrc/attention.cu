// svc_attention: non-causal softmax attention, head_dim 64, keys masked by kv_len[b].
// Replaces F.scaled_dot_product_attention with the (B,1,T,T) bool key mask
// (reference: modules/diffusion_transformer.py:255,518-520).  q and k arrive already rotated
// (RoPE) and q pre-scaled by 1/sqrt(64) from the wqkv GEMM epilogue.
//
// bf16: flash-style tcgen05 kernel, one CTA = 384 queries (three Q tiles) of one (batch, head), keys in
//   blocks of 64: TMA producer warp, one MMA-issuing warp per Q tile, three softmax warpgroups (one
//   query row per thread), S / O / P all in TMEM (P is the A operand of the PV MMA straight from TMEM).
//   Details and the measured bounds are in the comment above attention_tc_kernel.
// fp32: FFMA kernel, one query per thread ("fp32 mode").
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace svc {

#ifdef SVC_TRACE
__device__ long long g_attn_trace[4][64][8];   // [role][block][event] clock64 stamps of CTA (0,0,0)
#define TRACE(role, j, ev)                                                                   \
    do {                                                                                     \
        if (trace_on && (j) < 64) g_attn_trace[role][j][ev] = clock64();                     \
    } while (0)
#else
#define TRACE(role, j, ev) do {} while (0)
#endif

constexpr int AT_BM = 128;   // queries per Q tile (one TMEM lane each)
constexpr int AT_QT = 3;     // Q tiles per CTA = softmax groups taking turns on the MUFU pipe
constexpr int AT_BN = 64;    // keys per block
constexpr int AT_HD = 64;
#ifndef AT_KST_V
#define AT_KST_V 4
#endif
#ifndef AT_POLY_EVERY
#define AT_POLY_EVERY 3   // every 3rd key pair takes the FMA-pipe exp2 (0 = all on MUFU)
#endif
#ifndef AT_SM_REGS
#define AT_SM_REGS 152
#endif
constexpr int AT_KST = AT_KST_V;    // K / V ring depth
constexpr int AT_THREADS = 128 + AT_QT * 128;   // warpgroup 0: TMA + one MMA issuer per Q tile; then one softmax warpgroup per tile

struct alignas(64) AttnTcParams {
    CUtensorMap qmap, kmap, vmap;  // (H*64, T, B) bf16 views, box {64, 128, 1}
    uint16_t* out;             // bf16 or fp16 (template parameter F16 of the kernel)
    long long o_bstride, o_rstride;
    const int* kv_len;
    int T, H;
    int B, q_tiles, total_items;   // persistent grid: item = (b * H + h) * q_tiles + q_tile
};

struct AttnSmem {
    static constexpr int QTILE = AT_BM * AT_HD * 2;         // 16 KB
    static constexpr int KTILE = AT_BN * AT_HD * 2;         // 8 KB
    static constexpr int Q_OFF = 0;
    static constexpr int K_OFF = Q_OFF + AT_QT * QTILE;
    static constexpr int V_OFF = K_OFF + AT_KST * KTILE;
    static constexpr int BAR_OFF = V_OFF + AT_KST * KTILE;
    static constexpr int TOTAL = BAR_OFF + 512 + 1024;
};

// descriptor for a MN-major (rows = K index, 64 contiguous MN elements per row) SW128 tile
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t addr) {
    return umma_desc_sw128(addr, 1024, 1024);
}

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// 2^x for a pair of values on the FMA pipe (no MUFU): Cody-Waite split x = n + f, f in [-0.5, 0.5],
// degree-3 minimax polynomial (relative error 7.5e-5, far below the bf16 rounding of P), exponent
// inserted with an integer add.  Inputs <= -126 flush to 2^-126.
__device__ __forceinline__ float2 poly_exp2_x2(float2 x) {
    const float kMagic = 12582912.f;   // 1.5 * 2^23: x + kMagic rounds x to an integer in the low mantissa bits
    x.x = fmaxf(x.x, -126.f);
    x.y = fmaxf(x.y, -126.f);
    const float2 t = __fadd2_rn(x, make_float2(kMagic, kMagic));
    const float2 n = __fadd2_rn(t, make_float2(-kMagic, -kMagic));
    const float2 f = __fadd2_rn(x, make_float2(-n.x, -n.y));
    float2 p = __ffma2_rn(f, make_float2(0.0551716685f, 0.0551716685f), make_float2(0.2426111251f, 0.2426111251f));
    p = __ffma2_rn(p, f, make_float2(0.6932609677f, 0.6932609677f));
    p = __ffma2_rn(p, f, make_float2(0.9999280572f, 0.9999280572f));
    float2 r;
    r.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
    r.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
    return r;
}

// 3-input max (FMNMX3, sm_100+): halves the instruction count of a row-max pass
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float r;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

template <int N>
__device__ __forceinline__ void reg_dealloc() {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_alloc() {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}


// 32 lanes x 32 columns store (thread i writes row lane_base + i)
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16,"
        " %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
        "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
        "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]),
        "r"(r[30]), "r"(r[31])
        : "memory");
}

// Persistent kernel: one CTA per SM walks work items = AT_QT x 128 = 384 queries (three Q tiles) of one
// (batch, head); keys in blocks of 64.  Ring stages and barrier parities continue across items (running
// block counts), so the next item's loads and first S MMAs overlap the previous item's O read-out.
//   warp 0    : TMA producer (Q tiles once, K_j / V_j through mbarrier rings shared by the tiles)
//   warps 1-3 : one MMA-issuing thread per Q tile g: S_g = Q_g K_j^T (TMEM, 64 columns) and
//               O_g += P_g V_j (TMEM, 64 columns, accumulated in place; A operand = P_g read
//               straight from TMEM, B = V_j MN-major from shared memory).  Each tile runs its own
//               S / PV chain, so no tile waits behind another tile's barriers.
//   warps 4-15: softmax group g = (warp - 4) / 4, one query row per thread: S row to registers in
//               one tcgen05.ld pass, exp2 (packed f32x2 FFMA / FADD around MUFU.EX2) against a
//               reference max that is only moved when the row max grows by more than 2^8 (then O
//               in TMEM and l are rescaled, rare), P -> packed bf16 pairs -> tcgen05.st into the
//               tile's P columns (no shared-memory round trip, no proxy fence).
// TMEM columns: S 3 x 64 | O 3 x 64 | P 3 x 32 = 480 of 512.  The softmax is bound by the MUFU
// pipe (8 cycles per warp instruction per SM sub-partition, measured: scripts/ubench/mufu.cu)
// and by the fixed latencies of a block (barrier round trips, tcgen05.ld/st); three free-running
// groups per SM keep that pipe ~70% busy, where two groups taking turns reached 60%.
// Measured dead ends (round 2, B = 64, T = 2580, H = 8, isolated): skipping the per-block row max after block 0
// ("lazy max", guarded by the block sum) ran at 666 TFLOP/s against 872 with the FMNMX3 row max kept - the extra
// control dependency costs more than the 72 instructions it removes; polynomial exp2 on every 2nd / 4th pair
// 585 / 681 (every 3rd stays best); the packed ex2.approx.bf16x2 / f16x2 forms compile to two MUFU.EX2
// instructions per pair (scripts/ubench/mufu2.cu), so they cannot lift the XU bound either.
// Scalar forms of the packed f32x2 arithmetic (an FFMA2 costs 3 FMA-pipe cycles, two FFMAs with an immediate operand
// 2: scripts/ubench/pipes.cu) lose as well - 829 (s * log2e - m), 793 (+ row-sum adds), 833 (polynomial only), 755 (all)
// against 868: the softmax warps are short of ISSUE slots, not of FMA-pipe cycles.
// F16: operands (Q, K, V, the P tile and the output) are IEEE half instead of bf16 - same instruction, the
// format bits of the instruction descriptor and the conversions differ.
template <bool F16>
__global__ void __launch_bounds__(AT_THREADS, 1) attention_tc_kernel(const __grid_constant__ AttnTcParams p) {
    using S = AttnSmem;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint64_t* q_full = bars;                // 1
    uint64_t* k_full = bars + 1;            // AT_KST
    uint64_t* k_empty = k_full + AT_KST;
    uint64_t* v_full = k_empty + AT_KST;
    uint64_t* v_empty = v_full + AT_KST;
    uint64_t* s_full = v_empty + AT_KST;    // per group
    uint64_t* s_empty = s_full + AT_QT;
    uint64_t* p_full = s_empty + AT_QT;
    uint64_t* p_empty = p_full + AT_QT;
    uint64_t* q_empty = p_empty + AT_QT;    // 1: every tile's last S of the work item has retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(q_empty + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // Persistent CTA: work items (q tile triple, head, batch entry) blockIdx.x, blockIdx.x + gridDim.x, ...
    // Every role walks the same item sequence and keeps running block counts, so ring stages and barrier
    // parities simply continue across items: the next item's Q / K / V loads and its first S MMAs overlap the
    // previous item's O read-out, and TMEM / barriers are set up once per SM.
    auto decode = [&](int item, int& q0, int& h, int& b, int& kv_len, int& n_blocks) {
        const int qt = item % p.q_tiles;
        const int bh = item / p.q_tiles;
        h = bh % p.H;
        b = bh / p.H;
        q0 = qt * (AT_BM * AT_QT);
        kv_len = p.kv_len != nullptr ? p.kv_len[b] : p.T;
        kv_len = max(1, min(kv_len, p.T));
        n_blocks = (kv_len + AT_BN - 1) / AT_BN;
    };
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.qmap);
        tma_prefetch_desc(&p.kmap);
        tma_prefetch_desc(&p.vmap);
        mbar_init(q_full, 1);
        mbar_init(q_empty, AT_QT);
        for (int i = 0; i < AT_KST; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], AT_QT);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], AT_QT);
        }
        for (int i = 0; i < AT_QT; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 128);
            mbar_init(&p_full[i], 128);
            mbar_init(&p_empty[i], 1);
        }
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base;                    // AT_QT x 128 columns
    const uint32_t tmem_O = tmem_base + AT_QT * AT_BN;    // AT_QT x 64 columns
    const uint32_t tmem_P = tmem_O + AT_QT * AT_HD;       // AT_QT x AT_BN/2 columns: P as packed bf16 pairs

    // warpgroup 0 (TMA, MMA issuers) hands its registers to the softmax warpgroups
    if (warp == 0) {
        reg_dealloc<40>();
        if (lane == 0) {
            int st = 0;
            uint32_t ph = 0;
            int ic = 0;
            for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++ic) {
                int q0, h, b, kv_len, n_blocks;
                decode(item, q0, h, b, kv_len, n_blocks);
                mbar_wait(q_empty, (ic & 1) ^ 1);              // previous item's S MMAs no longer read Q
                mbar_expect_tx(q_full, AT_QT * S::QTILE);
                for (int g = 0; g < AT_QT; ++g)
                    tma_load_3d(smem + S::Q_OFF + g * S::QTILE, &p.qmap, q_full, h * AT_HD, q0 + g * AT_BM, b);
                for (int j = 0; j < n_blocks; ++j) {
                    mbar_wait(&k_empty[st], ph ^ 1);
                    mbar_expect_tx(&k_full[st], S::KTILE);
                    tma_load_3d(smem + S::K_OFF + st * S::KTILE, &p.kmap, &k_full[st], h * AT_HD, j * AT_BN, b);
                    mbar_wait(&v_empty[st], ph ^ 1);
                    mbar_expect_tx(&v_full[st], S::KTILE);
                    tma_load_3d(smem + S::V_OFF + st * S::KTILE, &p.vmap, &v_full[st], h * AT_HD, j * AT_BN, b);
                    if (++st == AT_KST) st = 0, ph ^= 1;
                }
            }
        }
    } else if (warp >= 1 && warp <= AT_QT) {
        reg_dealloc<40>();
        // MMA issuer of Q tile g = warp - 1.  The whole warp runs this loop (uniform control flow, all
        // operands warp-uniform) and only the tcgen05 instructions sit under elect_one(): ptxas then
        // keeps descriptors in uniform registers.  Issuing from a divergent `if (lane == 0)` region
        // cost an elect / 5 x R2UR / branch sequence per MMA (~110 cycles each, measured), which made
        // this thread the pacing resource of the tile's S -> softmax -> PV chain.
        const int g = warp - 1;
        const uint32_t idesc_s = umma_idesc_bf16(AT_BM, AT_BN, 0, 0, F16 ? 0u : 1u);
        const uint32_t idesc_o = umma_idesc_bf16(AT_BM, AT_HD, 0, 1, F16 ? 0u : 1u);   // B (=V) MN-major
        // descriptor lo words of the smem regions (address field only varies)
        const uint32_t qlo = desc_lo(smem_u32(smem + S::Q_OFF)) + g * (S::QTILE >> 4);
        const uint32_t k_lo0 = desc_lo(smem_u32(smem + S::K_OFF));
        const uint32_t v_lo0 = desc_lo(smem_u32(smem + S::V_OFF), 1024);   // MN-major: LBO = 1024
        const uint32_t tS = tmem_S + g * AT_BN, tO = tmem_O + g * AT_HD, tP = tmem_P + g * (AT_BN / 2);
        auto issue_s = [&](int st, bool last_of_item) {
            tc_fence_after();
            if (elect_one()) {
                const uint32_t klo = k_lo0 + st * (S::KTILE >> 4);
#pragma unroll
                for (int k = 0; k < AT_HD / 16; ++k)
                    tc_mma_f16_lh(tS, qlo + k * 2, kDescHiSw128, klo + k * 2, kDescHiSw128, idesc_s, k != 0);
                tc_commit(&s_full[g]);
                tc_commit(&k_empty[st]);
                if (last_of_item) tc_commit(q_empty);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int st, bool first) {
            tc_fence_after();
            if (elect_one()) {
                const uint32_t vlo = v_lo0 + st * (S::KTILE >> 4);
#pragma unroll
                for (int k = 0; k < AT_BN / 16; ++k)      // A = P from TMEM: 8 columns per 16 keys
                    tc_mma_f16_ts(tO, tP + k * 8, vlo + k * (2048 >> 4), kDescHiSw128, idesc_o,
                                  (k != 0 || !first) ? 1u : 0u);
                tc_commit(&p_empty[g]);
                tc_commit(&v_empty[st]);
            }
            __syncwarp();
        };
        int st_s = 0, ph_s = 0;          // ring stage / phase of the next S to issue
        int st_v = 0, ph_v = 0;          // ring stage / phase of the next PV to issue
        uint32_t ns = 0;                 // S MMAs issued so far by this tile (global over items)
        uint32_t npv = 0;                // PV MMAs issued so far
        int ic = 0;
        auto next_s = [&](bool last_of_item) {
            mbar_wait_warp(&k_full[st_s], ph_s);
            if (ns > 0) mbar_wait_warp(&s_empty[g], (ns - 1) & 1);     // softmax g holds S(ns-1) in registers
            issue_s(st_s, last_of_item);
            if (++st_s == AT_KST) st_s = 0, ph_s ^= 1;
            ++ns;
        };
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x, ++ic) {
            int q0, h, b, kv_len, n_blocks;
            decode(item, q0, h, b, kv_len, n_blocks);
            mbar_wait_warp(q_full, ic & 1);
            next_s(n_blocks == 1);
            for (int j = 0; j < n_blocks; ++j) {
                if (j + 1 < n_blocks) next_s(j + 2 == n_blocks);
                mbar_wait_warp(&v_full[st_v], ph_v);
                mbar_wait_warp(&p_full[g], npv & 1);
                issue_pv(st_v, j == 0);
                if (++st_v == AT_KST) st_v = 0, ph_v ^= 1;
                ++npv;
            }
        }
    } else if (warp < 4) {
        reg_dealloc<40>();
    } else {
        // ===================== softmax groups: one query row per thread =====================
        reg_alloc<AT_QT == 2 ? 224 : AT_SM_REGS>();
        const int g = (warp - 4) >> 2;
        const int lg = warp & 3;
        const int row = lg * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(lg * 32) << 16;
        const uint32_t tS = tmem_S + g * AT_BN + lane_addr;
        const uint32_t tO = tmem_O + g * AT_HD + lane_addr;
        constexpr float kLog2e = 1.4426950408889634f;
        constexpr float kThresh = 8.0f;       // in log2 units
        const uint32_t tP = tmem_P + g * (AT_BN / 2) + lane_addr;
        float m_ref = 0.f, l_run = 0.f;       // m_ref in log2 units (already * log2e)
        uint32_t nb = 0;                      // blocks this tile has processed so far (all items): barrier parities
        int kv_len = 0, n_blocks = 0;

        // one KV block of this row; TAIL = the block holds keys >= kv_len (masked), only the last one
        auto block = [&](int j, auto tail_tag) {
            constexpr bool TAIL = decltype(tail_tag)::value;
            mbar_wait(&s_full[g], nb & 1);
            tc_fence_after();
            constexpr int NC = AT_BN / 32;
            float s[NC][32];
            {
                uint32_t r[NC][32];
#pragma unroll
                for (int c = 0; c < NC; ++c) tmem_ld_32x32(tS + c * 32, r[c]);
                tc_wait_ld();
                tc_fence_before();
                mbar_arrive(&s_empty[g]);
                const int kbase = j * AT_BN;
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        s[c][i] = __uint_as_float(r[c][i]);
                        if (TAIL && kbase + c * 32 + i >= kv_len) s[c][i] = -INFINITY;
                    }
            }
            // row max: 8 independent chains of 3-input max (a single fmax chain is AT_BN x 4 cycles of latency)
            auto rowmax = [&]() {
                float mxa[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) mxa[i] = -INFINITY;
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int i = 0; i < 32; i += 2) mxa[(i >> 1) & 7] = fmax3(mxa[(i >> 1) & 7], s[c][i], s[c][i + 1]);
                return kLog2e * fmax3(fmax3(mxa[0], mxa[1], mxa[2]), fmax3(mxa[3], mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7]));
            };
            // p = 2^(s * log2e - m) for the block, packed to bf16 pairs; returns the row sum.
            // Packed f32x2 arithmetic: one FFMA2 / FADD2 per pair of keys.
            uint32_t pk[AT_BN / 2];
            auto exps = [&](float m) {
                float2 la2[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
                const float2 sc2 = make_float2(kLog2e, kLog2e), mn2 = make_float2(-m, -m);
#pragma unroll
                for (int c = 0; c < NC; ++c)
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float2 x = __ffma2_rn(make_float2(s[c][i], s[c][i + 1]), sc2, mn2);
#if AT_POLY_EVERY > 0
                        float2 pp;
                        if (!TAIL && ((i >> 1) % AT_POLY_EVERY) == AT_POLY_EVERY - 1) pp = poly_exp2_x2(x);
                        else pp = make_float2(fast_exp2(x.x), fast_exp2(x.y));
#else
                        const float2 pp = make_float2(fast_exp2(x.x), fast_exp2(x.y));
#endif
                        la2[(i >> 1) & 3] = __fadd2_rn(la2[(i >> 1) & 3], pp);
                        pk[c * 16 + (i >> 1)] = pack_op16<F16>(pp.x, pp.y);
                    }
                const float2 l2 = __fadd2_rn(__fadd2_rn(la2[0], la2[1]), __fadd2_rn(la2[2], la2[3]));
                return l2.x + l2.y;
            };
            bool need = false;
            float m_new = m_ref, l_blk;
            const float mx = rowmax();
            if (j == 0) {
                m_new = mx;
            } else if (mx > m_ref + kThresh) {
                need = true;
                m_new = mx;
            }
            l_blk = exps(m_new);
            // PV of the previous block must have retired before P / O are touched
            mbar_wait(&p_empty[g], (nb & 1) ^ 1);
            if (__any_sync(0xffffffffu, need)) {
                tc_fence_after();
                const float alpha = need ? fast_exp2(m_ref - m_new) : 1.0f;
                l_run *= alpha;
#pragma unroll
                for (int c = 0; c < AT_HD; c += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(tO + c, r);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * alpha);
                    tmem_st_32x32(tO + c, r);
                }
                tc_wait_st();
            }
            m_ref = m_new;
            l_run += l_blk;
            // P row: AT_BN keys as packed bf16 pairs -> this row's lane of the P columns in TMEM
#pragma unroll
            for (int c = 0; c < AT_BN / 64; ++c)
                tmem_st_32x32(tP + c * 32, *reinterpret_cast<uint32_t (*)[32]>(&pk[c * 32]));
            tc_wait_st();
            tc_fence_before();
            mbar_arrive(&p_full[g]);
            ++nb;
        };
        for (int item = blockIdx.x; item < p.total_items; item += gridDim.x) {
            int q0, h, b;
            decode(item, q0, h, b, kv_len, n_blocks);
            m_ref = 0.f, l_run = 0.f;
            for (int j = 0; j < n_blocks - 1; ++j) block(j, std::false_type{});
            if (n_blocks * AT_BN > kv_len) block(n_blocks - 1, std::true_type{});
            else block(n_blocks - 1, std::false_type{});
            // last PV of the item retired -> O complete
            mbar_wait(&p_empty[g], (nb - 1) & 1);
            tc_fence_after();
            const int t = q0 + g * AT_BM + row;
            const float inv_l = 1.0f / l_run;
            uint16_t* o = p.out + static_cast<long long>(b) * p.o_bstride +
                               static_cast<long long>(t) * p.o_rstride + h * AT_HD;
#pragma unroll
            for (int c = 0; c < AT_HD; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tO + c, r);
                tc_wait_ld();
                if (t < p.T) {
#pragma unroll
                    for (int d = 0; d < 32; d += 8) {
                        uint4 q;
                        q.x = pack_op16<F16>(__uint_as_float(r[d]) * inv_l, __uint_as_float(r[d + 1]) * inv_l);
                        q.y = pack_op16<F16>(__uint_as_float(r[d + 2]) * inv_l, __uint_as_float(r[d + 3]) * inv_l);
                        q.z = pack_op16<F16>(__uint_as_float(r[d + 4]) * inv_l, __uint_as_float(r[d + 5]) * inv_l);
                        q.w = pack_op16<F16>(__uint_as_float(r[d + 6]) * inv_l, __uint_as_float(r[d + 7]) * inv_l);
                        *reinterpret_cast<uint4*>(o + c + d) = q;
                    }
                }
            }
            tc_fence_before();        // the next item's first PV overwrites O: order it after these reads
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------
// fp32 / debug path: one query per thread, K/V tiles of 32 keys staged in shared memory
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) attention_simt_kernel(
    const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, long long bstride,
    long long rstride, T* __restrict__ out, long long o_bstride, long long o_rstride, int Tn,
    const int* __restrict__ kv_len_p) {
    constexpr int KT = 32;
    __shared__ float ks[KT][AT_HD];
    __shared__ float vs[KT][AT_HD];
    const int h = blockIdx.y, b = blockIdx.z;
    const int t = blockIdx.x * 128 + threadIdx.x;
    int kv_len = kv_len_p != nullptr ? kv_len_p[b] : Tn;
    kv_len = max(1, min(kv_len, Tn));
    const long long base = static_cast<long long>(b) * bstride + h * AT_HD;
    float qr[AT_HD], o[AT_HD];
    const int tq = min(t, Tn - 1);
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) {
        qr[d] = to_f32<T>(q[base + static_cast<long long>(tq) * rstride + d]);
        o[d] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < kv_len; k0 += KT) {
        __syncthreads();
        for (int i = threadIdx.x; i < KT * AT_HD; i += 128) {
            const int r = i / AT_HD, d = i % AT_HD;
            const int kk = k0 + r;
            float kv = 0.f, vv = 0.f;
            if (kk < kv_len) {
                kv = to_f32<T>(k[base + static_cast<long long>(kk) * rstride + d]);
                vv = to_f32<T>(v[base + static_cast<long long>(kk) * rstride + d]);
            }
            ks[r][d] = kv;
            vs[r][d] = vv;
        }
        __syncthreads();
        const int nk = min(KT, kv_len - k0);
        for (int r = 0; r < nk; ++r) {
            float s = 0.f;
#pragma unroll
            for (int d = 0; d < AT_HD; ++d) s = fmaf(qr[d], ks[r][d], s);
            const float m_new = fmaxf(m, s);
            const float a = __expf(m - m_new);
            const float pz = __expf(s - m_new);
            l = l * a + pz;
#pragma unroll
            for (int d = 0; d < AT_HD; ++d) o[d] = fmaf(o[d], a, pz * vs[r][d]);
            m = m_new;
        }
    }
    if (t < Tn) {
        const float inv = 1.0f / l;
        T* op = out + static_cast<long long>(b) * o_bstride + static_cast<long long>(t) * o_rstride +
                h * AT_HD;
#pragma unroll
        for (int d = 0; d < AT_HD; ++d) op[d] = from_f32<T>(o[d] * inv);
    }
}

bool encode_bf16_map(CUtensorMap* map, const void* ptr, int K, long long rows, long long rstride,
                     long long batches, long long bstride, int box_rows);

}  // namespace svc

extern "C" int svc_attention(const void* q, const void* k, const void* v, long long qkv_bstride,
                             long long qkv_rstride, void* out, long long out_bstride,
                             long long out_rstride, int B, int T, int H, const int* kv_len,
                             int dtype, int backend, void* stream) {
    using namespace svc;
    if (B < 1 || T < 1 || H < 1 || q == nullptr || k == nullptr || v == nullptr || out == nullptr) {
        svc_set_error("svc_attention: bad arguments");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if ((dtype == SVC_BF16 || dtype == SVC_F16) && backend == SVC_BACKEND_AUTO) {
        AttnTcParams p;
        if (!encode_bf16_map(&p.qmap, q, H * AT_HD, T, qkv_rstride, B, qkv_bstride, AT_BM) ||
            !encode_bf16_map(&p.kmap, k, H * AT_HD, T, qkv_rstride, B, qkv_bstride, AT_BN) ||
            !encode_bf16_map(&p.vmap, v, H * AT_HD, T, qkv_rstride, B, qkv_bstride, AT_BN)) {
            svc_set_error("svc_attention: tensor map encode failed (alignment)");
            return SVC_ERR_ARG;
        }
        if (reinterpret_cast<uintptr_t>(out) % 16 || (out_bstride * 2) % 16 || (out_rstride * 2) % 16) {
            svc_set_error("svc_attention: out must be 16-byte aligned");
            return SVC_ERR_ARG;
        }
        p.out = static_cast<uint16_t*>(out);
        p.o_bstride = out_bstride;
        p.o_rstride = out_rstride;
        p.kv_len = kv_len;
        p.T = T;
        p.H = H;
        static bool attr_set = false;
        if (!attr_set) {
            cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 AttnSmem::TOTAL);
            cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 AttnSmem::TOTAL);
            attr_set = true;
        }
        p.B = B;
        p.q_tiles = (T + AT_BM * AT_QT - 1) / (AT_BM * AT_QT);
        const long long total = static_cast<long long>(p.q_tiles) * H * B;
        if (total > 0x7fffffffLL) {
            svc_set_error("svc_attention: too many work items");
            return SVC_ERR_ARG;
        }
        p.total_items = static_cast<int>(total);
        const int grid = p.total_items < 148 ? p.total_items : 148;     // one persistent CTA per SM
        if (dtype == SVC_F16) attention_tc_kernel<true><<<grid, AT_THREADS, AttnSmem::TOTAL, st>>>(p);
        else attention_tc_kernel<false><<<grid, AT_THREADS, AttnSmem::TOTAL, st>>>(p);
        SVC_CHECK_LAUNCH();
        return SVC_OK;
    }
    dim3 grid((T + 127) / 128, H, B);
    if (dtype == SVC_F32)
        attention_simt_kernel<float><<<grid, 128, 0, st>>>(
            static_cast<const float*>(q), static_cast<const float*>(k), static_cast<const float*>(v),
            qkv_bstride, qkv_rstride, static_cast<float*>(out), out_bstride, out_rstride, T, kv_len);
    else if (dtype == SVC_F16)
        attention_simt_kernel<__half><<<grid, 128, 0, st>>>(
            static_cast<const __half*>(q), static_cast<const __half*>(k), static_cast<const __half*>(v),
            qkv_bstride, qkv_rstride, static_cast<__half*>(out), out_bstride, out_rstride, T, kv_len);
    else
        attention_simt_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(
            static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k),
            static_cast<const __nv_bfloat16*>(v), qkv_bstride, qkv_rstride,
            static_cast<__nv_bfloat16*>(out), out_bstride, out_rstride, T, kv_len);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

#ifdef SVC_TRACE
extern "C" int svc_debug_attn_trace(long long* host, int n) {
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(host, svc::g_attn_trace, sizeof(long long) * n) == cudaSuccess ? 0 : -2;
}
#endif
