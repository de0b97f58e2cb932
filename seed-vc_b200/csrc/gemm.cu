// svc_gemm: every dense contraction of the hot path (Linear / concat-free Linear / Conv1d taps /
// polyphase ConvTranspose1d) as one segmented GEMM with a fused epilogue.
//
//  * bf16 operands  -> tcgen05.mma (UMMA 128 x N x 16, fp32 accumulators in TMEM), operands
//                      staged by TMA (cp.async.bulk.tensor, 128B swizzle) through an mbarrier
//                      ring; warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner,
//                      warps 2-5 = epilogue (tcgen05.ld -> registers -> fused epilogue -> HBM).
//  * fp32 operands  -> FFMA shared-memory tiled mainloop ("fp32 mode", small-M conditioning
//                      GEMMs, and a debug cross-check of the tensor-core path).
//
// Reference sites replaced: nn.Linear / nn.Conv1d / nn.ConvTranspose1d calls listed in
// include/seedvc_b200.h next to svc_gemm.
#include <string.h>

#include <mutex>

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
    EpiParams epi;
};

template <int BN, int STAGES>
struct TcSmem {
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int BAR_OFFSET = STAGES * STAGE_BYTES;
    static constexpr int TOTAL = BAR_OFFSET + 256 + 1024;  // + barriers + alignment slack
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(192) gemm_tc_kernel(const __grid_constant__ TcParams p) {
    using S = TcSmem<BN, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::BAR_OFFSET);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n_tile = blockIdx.x % p.n_tiles;
    const int m_tile = blockIdx.x / p.n_tiles;
    const int b = m_tile / p.tiles_per_batch;
    const int t0 = (m_tile % p.tiles_per_batch) * BM;
    const int n0 = n_tile * BN;

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < kMaxMaps; ++i) {
            tma_prefetch_desc(&p.amap[i]);
            tma_prefetch_desc(&p.wmap[i]);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, BN);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int s = 0; s < p.n_seg; ++s) {
                const TcSeg sg = p.seg[s];
                for (int kb = 0; kb < sg.nkb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], S::STAGE_BYTES);
                    uint8_t* sa = smem + stage * S::STAGE_BYTES;
                    tma_load_3d(sa, &p.amap[sg.a_map], &full_bar[stage], kb * BK, t0 + sg.shift, b);
                    // weight maps are rank 3 too (K, rows, 1): the TMA instruction rank must
                    // match the tensor-map rank
                    tma_load_3d(sa + S::A_BYTES, &p.wmap[sg.w_map], &full_bar[stage], kb * BK,
                                sg.w_row0 + n0, 0);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            int n_umma = p.epi.N - n0;
            n_umma = n_umma > BN ? BN : ((n_umma + 15) & ~15);
            const uint32_t idesc = umma_idesc_bf16(BM, n_umma, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < p.total_kb; ++it) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * S::STAGE_BYTES);
                const uint32_t sb = sa + S::A_BYTES;
#pragma unroll
                for (int k = 0; k < BK / 16; ++k) {
                    const uint64_t da = umma_desc_sw128(sa + k * 32, 0, 1024);
                    const uint64_t db = umma_desc_sw128(sb + k * 32, 0, 1024);
                    tc_mma_f16(tmem_base, da, db, idesc, (it | k) != 0 ? 1u : 0u);
                }
                tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                if (++stage == STAGES) {
                    stage = 0;
                    phase ^= 1;
                }
            }
            tc_commit(tmem_full_bar);
        }
    } else {
        // ===================== epilogue =====================
        const int lg = warp & 3;  // TMEM lane group this warp may access
        const int row = lg * 32 + lane;
        const int t = t0 + row;
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        const int ncols = min(BN, p.epi.N - n0);
        for (int c = 0; c < ncols; c += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(lg * 32) << 16) + c, r);
            tc_wait_ld();
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            if (t < p.T) epilogue_chunk<32>(p.epi, b, t, n0 + c, v);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, BN);
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

template <int BN, int STAGES>
static int launch_tc(const TcParams& p, int m_tiles, cudaStream_t stream) {
    using S = TcSmem<BN, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             S::TOTAL);
        attr_set = true;
    }
    gemm_tc_kernel<BN, STAGES><<<m_tiles * p.n_tiles, 192, S::TOTAL, stream>>>(p);
    SVC_CHECK_LAUNCH();
    return SVC_OK;
}

static int gemm_tc(const svc_gemm_desc& d, cudaStream_t stream) {
    TcParams p;
    memset(&p, 0, sizeof(p));
    int BN = 128;
    if (d.N <= 32) BN = 32;
    else if (d.N <= 64) BN = 64;
    else if (d.N <= 128) BN = 128;
    else if (d.N % 256 == 0 || d.N > 1024) BN = 256;
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
            if (wkeys[i].rs == d.w_rstride[s] && wkeys[i].K == d.K[s] && wp >= wkeys[i].base &&
                (wp - wkeys[i].base) % rb == 0 && (wp - wkeys[i].base) / rb < (1 << 24)) {
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
    p.n_seg = d.n_seg;
    p.total_kb = total_kb;
    p.B = d.B;
    p.T = d.T;
    p.tiles_per_batch = (d.T + BM - 1) / BM;
    p.n_tiles = (d.N + BN - 1) / BN;
    p.epi = make_epi_params(d);
    const int m_tiles = d.B * p.tiles_per_batch;
    switch (BN) {
        case 32: return launch_tc<32, 4>(p, m_tiles, stream);
        case 64: return launch_tc<64, 4>(p, m_tiles, stream);
        case 128: return launch_tc<128, 3>(p, m_tiles, stream);
        default: return launch_tc<256, 4>(p, m_tiles, stream);
    }
}

static int gemm_simt(const svc_gemm_desc& d, cudaStream_t stream) {
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
    if (d->dtype != SVC_BF16 && d->dtype != SVC_F32) {
        svc_set_error("svc_gemm: dtype must be SVC_BF16 or SVC_F32");
        return SVC_ERR_ARG;
    }
    if ((d->act == SVC_ACT_SWIGLU_PAIR || d->act == SVC_ACT_TANH_SIG_PAIR) && (d->N % 2)) {
        svc_set_error("svc_gemm: pair activation needs even N");
        return SVC_ERR_ARG;
    }
    if (d->out_f32 == nullptr && d->out_op == nullptr) {
        svc_set_error("svc_gemm: no output");
        return SVC_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d->dtype == SVC_BF16 && backend == SVC_BACKEND_AUTO) return svc::gemm_tc(*d, st);
    return svc::gemm_simt(*d, st);
}
