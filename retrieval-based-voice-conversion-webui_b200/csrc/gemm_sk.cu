// Cluster split-K variant of the implicit-GEMM engine for SMALL-M, LONG-K launches (RMVPE's deep 3x3 layers: M = 204 rows,
// K = 4608; flow / FFN layers with M <= 1600): with <= 40 output tiles the persistent streaming kernel leaves most SMs idle
// and each busy SM is bound by its own L2->SMEM ingest (~46 B/clk).  Here one output tile is owned by a thread-block CLUSTER
// of SK CTAs; CTA r accumulates k-blocks [K r/SK, K (r+1)/SK) in TMEM, parks its fp32 partial tile in shared memory, and after
// one cluster barrier every CTA reduces a 128/SK-row slab of the tile over DISTRIBUTED SHARED MEMORY in a fixed order
// (deterministic, no atomics) and applies the fused epilogue to it.  Same contract as gemm_tc.cu (gemm.cuh); 1-D taps and the
// 3-D (C, W, H) conv2d map; not for gate / up2 / per-row-bias / batched launches.
#include <cooperative_groups.h>

#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace rvcb {

namespace {

constexpr int SK_BK = 64;

template <int BN>
struct SKCfg {
    static constexpr int A_STAGE = BM * SK_BK * 2;
    static constexpr int B_STAGE = (BN * SK_BK * 2 + 1023) / 1024 * 1024;
    static constexpr int STAGE = A_STAGE + B_STAGE;
    static constexpr int P_STRIDE = BN + 4;                          // floats per row of the partial tile (16 B aligned)
    static constexpr int P_BYTES = BM * P_STRIDE * 4;
    static constexpr int STAGES = 6;
    static constexpr int TMEM_COLS = BN <= 32 ? 32 : 64;
    static constexpr int SMEM = STAGES * STAGE + 1024 + 256 + P_BYTES;
    static constexpr uint32_t TX_BYTES = A_STAGE + BN * SK_BK * 2;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_sk_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ KParams p) {
    using C = SKCfg<BN>;
    cg::cluster_group cluster = cg::this_cluster();
    const int SK = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + C::STAGES * C::A_STAGE;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + C::STAGES;
    uint64_t* tfull_bar = bars + 2 * C::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 1);
    float* part = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE + 256);      // this CTA's partial tile [128][BN + 4]

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        mbar_init(tfull_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int tile = blockIdx.x / SK;
    const int mt = tile / p.num_n_tiles;
    const int nt = tile - mt * p.num_n_tiles;
    const int kb0 = (int)((long)p.total_kb * rank / SK), kb1 = (int)((long)p.total_kb * (rank + 1) / SK);

    if (warp == 0) {
        // ---- TMA producer: this CTA's share of the k-blocks ----
        int stage = 0;
        uint32_t phase = 0;
        int kb = 0;
        for (int s = 0; s < p.nseg; ++s) {
            const SegPacked sg = p.seg[s];
            for (int kc = 0; kc < sg.nk; ++kc, ++kb) {
                if (kb < kb0 || kb >= kb1) continue;
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&full_bar[stage], C::TX_BYTES);
                    const int ac0 = sg.col + kc * SK_BK;
                    if (p.conv2d_W == 0) tma_load_3d(smem_a + stage * C::A_STAGE, &tmap_a, &full_bar[stage], ac0, mt * BM + sg.row, 0);
                    else tma_load_3d(smem_a + stage * C::A_STAGE, &tmap_a, &full_bar[stage], ac0, (int)sg.dw, mt * p.BH + sg.row);
                    tma_load_2d(smem_b + stage * C::B_STAGE, &tmap_b, &full_bar[stage], p.b_col0 + kb * SK_BK, nt * BN);
                }
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer ----
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        constexpr uint32_t desc_hi = (uint32_t)(((uint64_t)((8 * SK_BK * 2) >> 4) << 32 | (1ull << 46) | (2ull << 61)) >> 32);
        const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFF) >> 4) | (1u << 16);
        const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFF) >> 4) | (1u << 16);
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < kb1 - kb0; ++i) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a_lo = a_lo0 + (uint32_t)stage * (uint32_t)(C::A_STAGE / 16);
                const uint32_t b_lo = b_lo0 + (uint32_t)stage * (uint32_t)(C::B_STAGE / 16);
#pragma unroll
                for (int k = 0; k < SK_BK / 16; ++k) {
                    const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2 * k);
                    const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2 * k);
                    umma_f16(tmem_base, da, db, idesc, (i | k) != 0 ? 1u : 0u);
                }
                umma_commit(&empty_bar[stage]);
            }
            __syncwarp();
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
        if (elect_one()) umma_commit(tfull_bar);
        __syncwarp();
    } else {
        // ---- partial tile: TMEM -> registers -> shared memory (thread = row) ----
        constexpr int W = BN / 2;                            // columns per warp: 4 lane quadrants x 2 column halves
        const int quarter = warp & 3, chalf = (warp - 2) >> 2;
        const int row = quarter * 32 + lane;
        if (kb1 > kb0) {
            mbar_wait(tfull_bar, 0);
            tc_fence_after();
        }
        float* dst = part + row * C::P_STRIDE + chalf * W;
#pragma unroll
        for (int c = 0; c < W / 16; ++c) {
            uint32_t raw[16];
            if (kb1 > kb0) {
                tmem_ld16(tmem_base + chalf * W + c * 16 + ((uint32_t)(quarter * 32) << 16), raw);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) raw[i] = 0u;
            }
#pragma unroll
            for (int i = 0; i < 16; i += 4)
                *reinterpret_cast<float4*>(dst + c * 16 + i) = make_float4(__uint_as_float(raw[i]), __uint_as_float(raw[i + 1]),
                                                                           __uint_as_float(raw[i + 2]), __uint_as_float(raw[i + 3]));
        }
        tc_fence_before();
    }

    cluster.sync();                                          // every CTA's partial tile is in its shared memory

    if (warp >= 2) {
        // ---- reduce this CTA's row slab over the cluster (fixed order) + fused epilogue ----
        const int rows_per = BM / SK;
        constexpr int NV = BN / 4;
        const int tid = threadIdx.x - 64;
        const float* parts[8];
        for (int q = 0; q < SK; ++q) parts[q] = cluster.map_shared_rank(part, q);
        for (int it = tid; it < rows_per * NV; it += kEpiWarps * 32) {
            const int rr = rank * rows_per + it / NV;
            const int c4 = it - (it / NV) * NV;
            const int m = mt * BM + rr;
            const int col = nt * BN + 4 * c4;
            if (m >= p.M || col >= p.N) continue;
            float4 t = *reinterpret_cast<const float4*>(parts[0] + rr * C::P_STRIDE + 4 * c4);
            for (int q = 1; q < SK; ++q) {
                const float4 u = *reinterpret_cast<const float4*>(parts[q] + rr * C::P_STRIDE + 4 * c4);
                t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
            }
            if (p.bias) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col));
                t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
            }
            if (p.res1) {
                const float4 r = *reinterpret_cast<const float4*>(p.res1 + (long)m * p.ldres1 + col);
                t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
            }
            if (p.act1 != ACT_NONE) {
                t.x = apply_act(t.x, p.act1, p.act1_p); t.y = apply_act(t.y, p.act1, p.act1_p);
                t.z = apply_act(t.z, p.act1, p.act1_p); t.w = apply_act(t.w, p.act1, p.act1_p);
            }
            t.x *= p.alpha; t.y *= p.alpha; t.z *= p.alpha; t.w *= p.alpha;
            if (p.res2) {
                const float4 r = *reinterpret_cast<const float4*>(p.res2 + (long)m * p.ldres2 + col);
                t.x += r.x; t.y += r.y; t.z += r.z; t.w += r.w;
            }
            if (p.out32) *reinterpret_cast<float4*>(p.out32 + (long)m * p.ld32 + col) = t;
            if (p.out16) {
                float4 u = t;
                if (p.act2 != ACT_NONE) {
                    u.x = apply_act(u.x, p.act2, p.act2_p); u.y = apply_act(u.y, p.act2, p.act2_p);
                    u.z = apply_act(u.z, p.act2, p.act2_p); u.w = apply_act(u.w, p.act2, p.act2_p);
                }
                const __half2 h0 = __floats2half2_rn(u.x, u.y), h1 = __floats2half2_rn(u.z, u.w);
                *reinterpret_cast<uint2*>(p.out16 + (long)m * p.ld16 + col) =
                    make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
            }
        }
    }

    cluster.sync();                                          // nobody leaves while a peer may still read its partial tile
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

template <int BN>
void sk_launch(const CUtensorMap& ta, const CUtensorMap& tb, const KParams& p, int SK, cudaStream_t stream) {
    using C = SKCfg<BN>;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(gemm_sk_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.num_tiles * SK);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = C::SMEM;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = SK;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (gemm_prof_on()) gemm_prof_record_begin(stream);
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, gemm_sk_kernel<BN>, ta, tb, p));
    KERNEL_CHECK();
    if (gemm_prof_on()) gemm_prof_record_end(stream, {p.M, p.N, p.total_kb, SK_BK, BN, 1, p.nseg, p.num_tiles * SK});
    count_launch();
}

}  // namespace

// Returns false when the launch is not a split-K candidate (the caller falls through to the streaming kernel).
bool gemm_sk_try(const GemmArgs& g, cudaStream_t stream) {
    static const bool on = [] { const char* e = getenv("RVCB_SPLITK"); return !(e && e[0] == '0'); }();
    if (!on) return false;
    if (g.block_k != 64 || g.batch != 1 || g.gate || g.up2_C || g.bias_per_row || (g.N % 4) != 0) return false;
    auto al = [](const void* ptr, int a) { return ptr == nullptr || (reinterpret_cast<uintptr_t>(ptr) % a) == 0; };
    if (!(al(g.out32, 16) && al(g.out16, 8) && al(g.res1, 16) && al(g.res2, 16) && al(g.bias, 16))) return false;
    if ((g.out32 && g.ld32 % 4) || (g.out16 && g.ld16 % 4) || (g.res1 && g.ldres1 % 4) || (g.res2 && g.ldres2 % 4)) return false;
    int total_kb = 0;
    for (int s = 0; s < g.nseg; ++s) total_kb += g.seg[s].nk;
    const int m_tiles = ceil_div(g.M, BM);
    // narrowest tile that still leaves the cluster grid within one wave; split 4-way when K is long enough, else 2-way.
    // Measured (profiles/r1x): 204x512, K = 4608: 25.8 -> 18.5 us; 1598x192, K = 2304: 20.8 -> 16.9 us; with fewer than 16
    // k-blocks per CTA the two cluster barriers cost more than the shorter K loop saves (3264x128, K = 1152: 15.9 -> 16.8 us).
    // tuning knobs (measurement only): RVCB_SK_MAX = largest cluster tried (8 / 4 / 2), RVCB_SK_MINKB = fewest k-blocks per CTA,
    // RVCB_SK_BN64 = 1 tries the 64-wide tile first
    static const int sk_max = [] { const char* e = getenv("RVCB_SK_MAX"); return e ? atoi(e) : 4; }();
    static const int min_kb = [] { const char* e = getenv("RVCB_SK_MINKB"); return e ? atoi(e) : 16; }();
    static const bool bn64_first = [] { const char* e = getenv("RVCB_SK_BN64"); return e && e[0] == '1'; }();
    int BN = 0, SK = 0;
    const int bn_order[2] = {bn64_first ? 64 : 32, bn64_first ? 32 : 64};
    for (int bi = 0; bi < 2 && !BN; ++bi) {
        const int bn = bn_order[bi];
        if (bn > round_up(g.N, 32)) continue;
        const int tiles = m_tiles * ceil_div(g.N, bn);
        for (int sk : {8, 4, 2}) {
            if (sk > sk_max) continue;
            if (tiles * sk <= 148 && total_kb >= min_kb * sk) { BN = bn; SK = sk; break; }
        }
    }
    if (!BN) return false;
    if (g.conv2d_W) {
        if (!(g.conv2d_W >= 1 && g.conv2d_W <= 128 && (BM % g.conv2d_W) == 0)) return false;
    }
    KParams p{};
    p.M = g.M; p.N = g.N; p.nseg = g.nseg; p.batch = 1;
    p.num_m_tiles = m_tiles;
    p.num_n_tiles = ceil_div(g.N, BN);
    p.num_tiles = p.num_m_tiles * p.num_n_tiles;
    p.total_kb = total_kb;
    for (int s = 0; s < g.nseg; ++s) {
        p.seg[s].row = (short)g.seg[s].row_off;
        p.seg[s].col = (short)g.seg[s].col_off;
        p.seg[s].nk = (short)g.seg[s].nk;
        p.seg[s].dw = (signed char)g.seg[s].dw;
    }
    p.conv2d_W = g.conv2d_W;
    p.BH = g.conv2d_W ? BM / g.conv2d_W : 0;
    p.b_col0 = g.b_col0;
    p.bias = g.bias; p.res1 = g.res1; p.ldres1 = g.ldres1; p.res2 = g.res2; p.ldres2 = g.ldres2;
    p.alpha = g.alpha; p.act1 = g.act1; p.act1_p = g.act1_p; p.act2 = g.act2; p.act2_p = g.act2_p;
    p.out32 = g.out32; p.ld32 = g.ld32; p.out16 = g.out16; p.ld16 = g.ld16;
    p.vec_ok = 1;
    CUtensorMap ta, tb;
    if (g.conv2d_W == 0) {
        cuuint64_t dims[3] = {(cuuint64_t)g.a_cols, (cuuint64_t)g.a_rows, 1};
        cuuint64_t str[2] = {(cuuint64_t)g.lda * 2, (cuuint64_t)g.lda * 2 * (cuuint64_t)g.a_rows};
        cuuint32_t box[3] = {(cuuint32_t)SK_BK, (cuuint32_t)BM, 1};
        encode_map(&ta, g.A, 3, dims, str, box, SK_BK);
    } else {
        const int W = g.conv2d_W;
        cuuint64_t dims[3] = {(cuuint64_t)g.a_cols, (cuuint64_t)W, (cuuint64_t)g.a_rows};
        cuuint64_t str[2] = {(cuuint64_t)g.lda * 2, (cuuint64_t)g.lda * 2 * (cuuint64_t)W};
        cuuint32_t box[3] = {(cuuint32_t)SK_BK, (cuuint32_t)W, (cuuint32_t)(BM / W)};
        encode_map(&ta, g.A, 3, dims, str, box, SK_BK);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)g.b_cols, (cuuint64_t)g.b_rows};
        cuuint64_t str[1] = {(cuuint64_t)g.ldb * 2};
        cuuint32_t box[2] = {(cuuint32_t)SK_BK, (cuuint32_t)BN};
        encode_map(&tb, g.B, 2, dims, str, box, SK_BK);
    }
    if (BN == 32) sk_launch<32>(ta, tb, p, SK, stream);
    else sk_launch<64>(ta, tb, p, SK, stream);
    return true;
}

}  // namespace rvcb
