// tcgen05 + TMA implicit-GEMM kernel for sm_100a (see gemm.cuh for the contract).
//
// Persistent, warp-specialised CTA of 320 threads, one CTA per SM:
//   warp 0      TMA producer   (cp.async.bulk.tensor -> 128B/64B/32B-swizzled smem ring, mbarrier tx)
//   warp 1      MMA issuer     (warp-uniform loop, one elected lane issues tcgen05.mma kind::f16, fp32 accumulators in TMEM,
//                               tcgen05.commit releases smem slots / publishes the accumulator)
//   warps 2..9  epilogue       (tcgen05.ld TMEM -> registers -> warp-private smem transpose -> fused
//                               bias/act/residual with COALESCED residual loads and output stores)
// Two TMEM accumulator stages let the epilogue of tile i overlap the MMAs of tile i+1.
// Conv taps are K-segments whose A tile is the same matrix at a shifted TMA coordinate;
// zero padding comes from TMA out-of-bounds fill, so no im2col buffer ever exists.
#include "tc_common.cuh"

namespace rvcb {

template <int BN, int BK, bool RS = false>
struct Cfg {
    static constexpr int A_STAGE = BM * BK * 2;
    static constexpr int B_STAGE_RAW = BN * BK * 2;
    static constexpr int B_STAGE = (B_STAGE_RAW + 1023) / 1024 * 1024;
    static constexpr int STAGE = A_STAGE + B_STAGE;
    static constexpr int CW = BN >= 32 ? 32 : 16;                 // epilogue chunk width (columns)
    static constexpr int EPI_STRIDE = CW + 4;                     // floats per staged row (16B aligned, conflict-free)
    static constexpr int EPI_BYTES = kEpiWarps * 32 * EPI_STRIDE * 4;
    // RS ("residual staged"): the fp32 residual tile of res1 is TMA-prefetched into 128B-swizzled shared memory while the
    // tile's MMAs run, so the epilogue of large residual launches issues no global loads (panels of 32 columns x 128 rows)
    static constexpr int RES_BYTES = RS ? BM * BN * 4 : 0;
    static constexpr int PIPE_BUDGET = 226 * 1024 - EPI_BYTES - 2048 - 256 - RES_BYTES;
    // ring depth: what fits, up to 24 slots -- narrow tiles (BK = 16 / 32: 5-10 KB per slot) need many slots in flight to cover
    // the ~1 us L2->SMEM latency (8 slots of 5 KB bounded the C = 16 RMVPE layers at 18 B/clk per SM)
    static constexpr int STAGES = (PIPE_BUDGET / STAGE) > 24 ? 24 : (PIPE_BUDGET / STAGE);
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int SMEM = STAGES * STAGE + RES_BYTES + 1024 /*align slack*/ + 512 /*barriers*/ + EPI_BYTES;
    static constexpr uint32_t TX_BYTES = A_STAGE + B_STAGE_RAW;
};

template <int BN, int BK, bool RS>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_r, const __grid_constant__ KParams p) {
    using C = Cfg<BN, BK, RS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + C::STAGES * C::A_STAGE;
    uint8_t* smem_r = smem + C::STAGES * C::STAGE;                  // residual tile (RS only), 1024-byte aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE + C::RES_BYTES);
    uint64_t* full_bar = bars;                      // [STAGES]
    uint64_t* empty_bar = bars + C::STAGES;         // [STAGES]
    uint64_t* tfull_bar = bars + 2 * C::STAGES;     // [2]
    uint64_t* tempty_bar = bars + 2 * C::STAGES + 2;  // [2]
    uint64_t* r_full = bars + 2 * C::STAGES + 4;      // [1] residual tile landed
    uint64_t* r_empty = bars + 2 * C::STAGES + 5;     // [1] residual tile consumed by all epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::STAGES + 6);
    float* epi_smem = reinterpret_cast<float*>(smem + C::STAGES * C::STAGE + C::RES_BYTES + 512);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        for (int i = 0; i < C::STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], (BN / C::CW == 1) ? kEpiWarps / 2 : kEpiWarps);
        }
        if (RS) {
            prefetch_tmap(&tmap_r);
            mbar_init(r_full, 1);
            mbar_init(r_empty, kEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();      // dependents may start their prologue now ...
    pdl_wait();         // ... and nothing of ours touches global memory before the previous grids have completed

    const int tiles_per_z = p.num_m_tiles * p.num_n_tiles;

    if (warp == 0) {
        // ======================= TMA producer (warp-uniform loop; one elected lane issues) =======================
        int stage = 0;
        uint32_t phase = 0;
        uint32_t res_it = 0;                            // RS: tiles whose residual has been requested
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int z = tile / tiles_per_z;
            const int rem = tile - z * tiles_per_z;
            const int mt = rem / p.num_n_tiles;
            const int nt = rem - mt * p.num_n_tiles;
            int kb = 0;
            // RS: the single residual buffer frees up when the epilogue of the previous tile is done, i.e. some way into this
            // tile's MMAs: poll for it between operand loads so the operand pipeline never stalls behind it
            bool res_issued = !RS;
            auto issue_res = [&]() {
                if (elect_one()) {
                    mbar_expect_tx(r_full, (uint32_t)C::RES_BYTES);
#pragma unroll
                    for (int pn = 0; pn < BN / 32; ++pn)
                        tma_load_2d(smem_r + pn * (BM * 128), &tmap_r, r_full, nt * BN + pn * 32, mt * BM);
                }
                __syncwarp();
                res_issued = true;
            };
            for (int s = 0; s < p.nseg; ++s) {
                const SegPacked sg = p.seg[s];
                for (int kc = 0; kc < sg.nk; ++kc, ++kb) {
                    if (RS && !res_issued && mbar_test(r_empty, (res_it & 1) ^ 1)) issue_res();
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    if (elect_one()) {
                        mbar_expect_tx(&full_bar[stage], C::TX_BYTES);
                        const int ac0 = z * p.a_col_z + sg.col + kc * BK;
                        if (p.conv2d_W == 0) {
                            tma_load_3d(smem_a + stage * C::A_STAGE, &tmap_a, &full_bar[stage], ac0,
                                        mt * BM + sg.row + z * p.a_row_z, 0);
                        } else {
                            tma_load_3d(smem_a + stage * C::A_STAGE, &tmap_a, &full_bar[stage], ac0, (int)sg.dw,
                                        mt * p.BH + sg.row);
                        }
                        tma_load_2d(smem_b + stage * C::B_STAGE, &tmap_b, &full_bar[stage],
                                    p.b_col0 + z * p.b_col_z + kb * BK, nt * BN + z * p.b_row_z);
                    }
                    __syncwarp();
                    if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
                }
            }
            if (RS) {
                if (!res_issued) {
                    mbar_wait(r_empty, (res_it & 1) ^ 1);
                    issue_res();
                }
                ++res_it;
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer (warp-uniform loop; one elected lane issues) =======================
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        // smem descriptor = {lo: start address >> 4 | LBO(1) << 16, hi: SBO | version | layout}; a K step of 16 fp16 is +2 in lo
        constexpr uint32_t desc_hi = (uint32_t)(((uint64_t)((8 * BK * 2) >> 4) << 32 | (1ull << 46) |
                                                 ((BK == 64 ? 2ull : BK == 32 ? 4ull : 6ull) << 61)) >> 32);
        const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFF) >> 4) | (1u << 16);
        const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFF) >> 4) | (1u << 16);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
            tc_fence_after();
            const uint32_t tmem_c = tmem_base + acc * BN;
            for (int kb = 0; kb < p.total_kb; ++kb) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t a_lo = a_lo0 + (uint32_t)stage * (uint32_t)(C::A_STAGE / 16);
                    const uint32_t b_lo = b_lo0 + (uint32_t)stage * (uint32_t)(C::B_STAGE / 16);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2 * k);
                        const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2 * k);
                        umma_f16(tmem_c, da, db, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);     // frees the smem slot once these MMAs retire
                }
                __syncwarp();
                if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) umma_commit(&tfull_bar[acc]);   // accumulator complete
            __syncwarp();
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // ======================= epilogue (8 warps) =======================
        // TMEM lane quarter = warp % 4 (hardware rule); the two warps sharing a quarter split the column chunks.
        constexpr int CW = C::CW, ST = C::EPI_STRIDE, NV = CW / 4;     // NV float4 per staged row
        constexpr int RPI = 32 / NV;                                   // rows covered by one warp-wide float4 access
        constexpr int NIT = 32 / RPI;                                  // iterations to cover 32 rows
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int cgroup = ew >> 2;                                    // 0 / 1
        float* stg = epi_smem + ew * 32 * ST;
        const int c4 = lane % NV, rsub = lane / NV;
        // single-chunk tiles (BN == CW): the two warp groups take alternate tiles (group g <-> accumulator stage g);
        // otherwise both groups work on every tile and split its column chunks.
        constexpr bool kAlt = (BN / CW == 1);
        const int c_start = kAlt ? 0 : cgroup;
        constexpr int c_step = kAlt ? 1 : 2;
        int iter = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++iter) {
            if (kAlt && (iter & 1) != cgroup) continue;
            const int acc = iter & 1;
            const uint32_t acc_phase = (iter >> 1) & 1;
            const int z = tile / tiles_per_z;
            const int rem = tile - z * tiles_per_z;
            const int mt = rem / p.num_n_tiles;
            const int nt = rem - mt * p.num_n_tiles;
            const int m_warp0 = mt * BM + quarter * 32;                // first row of this warp's 32-row slab
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            if (RS) mbar_wait(r_full, (uint32_t)iter & 1);             // this tile's residual is in shared memory
            const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
            const float* biasz = p.bias ? p.bias + z * p.bias_z : nullptr;
            const float bias_row = (p.bias && p.bias_per_row && (m_warp0 + lane) < p.M) ? p.bias[m_warp0 + lane] : 0.f;
            const long zoff = z * p.c_z;

#pragma unroll 1
            for (int c = c_start; c < BN / CW; c += c_step) {
                const int n0 = nt * BN + c * CW;
                if (n0 >= p.N) break;                                   // warp-uniform
                // ---- TMEM -> registers (thread = row) ----
                float v[CW];
                {
                    uint32_t raw[16];
                    tmem_ld16(taddr + c * CW, raw);
                    if constexpr (CW == 32) {
                        uint32_t raw2[16];
                        tmem_ld16(taddr + c * CW + 16, raw2);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[16 + i] = __uint_as_float(raw2[i]);
                    } else {
                        tmem_ld_wait();
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
                }
                if (p.bias) {
                    if (p.bias_per_row) {
#pragma unroll
                        for (int i = 0; i < CW; ++i) v[i] += bias_row;
                    } else if (n0 + CW <= p.N) {
#pragma unroll
                        for (int i = 0; i < CW; i += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(biasz + n0 + i));
                            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < CW; ++i)
                            if (n0 + i < p.N) v[i] += __ldg(biasz + n0 + i);
                    }
                }
                // ---- transpose through warp-private smem: afterwards lane = (row rsub + RPI*it, float4 column c4) ----
                __syncwarp();
#pragma unroll
                for (int i = 0; i < CW; i += 4)
                    *reinterpret_cast<float4*>(stg + lane * ST + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                __syncwarp();
                const int col = n0 + 4 * c4;
                const bool col_full = (col + 3 < p.N);
                const bool vec = col_full && p.vec_ok;
                float4 t[NIT];
#pragma unroll
                for (int it = 0; it < NIT; ++it) t[it] = *reinterpret_cast<const float4*>(stg + (rsub + RPI * it) * ST + 4 * c4);
                // residual 1 (before the activation): from the TMA-staged tile (RS), else coalesced from global memory
                if (RS) {
#pragma unroll
                    for (int it = 0; it < NIT; ++it) {
                        const int row = quarter * 32 + rsub + RPI * it;               // tile row; panel = chunk c (CW == 32)
                        const float4 q = *reinterpret_cast<const float4*>(smem_r + c * (BM * 128) + row * 128 + ((c4 ^ (row & 7)) << 4));
                        t[it].x += q.x; t[it].y += q.y; t[it].z += q.z; t[it].w += q.w;
                    }
                } else if (p.res1) {
#pragma unroll
                    for (int it = 0; it < NIT; ++it) {
                        const int m = m_warp0 + rsub + RPI * it;
                        if (m < p.M && col < p.N) {
                            const float* r = p.res1 + zoff + (long)m * p.ldres1 + col;
                            if (vec) {
                                const float4 q = *reinterpret_cast<const float4*>(r);
                                t[it].x += q.x; t[it].y += q.y; t[it].z += q.z; t[it].w += q.w;
                            } else {
                                t[it].x += r[0];
                                if (col + 1 < p.N) t[it].y += r[1];
                                if (col + 2 < p.N) t[it].z += r[2];
                                if (col + 3 < p.N) t[it].w += r[3];
                            }
                        }
                    }
                }
                if (p.gate) {
                    // (tanh, sigmoid) column pairs -> N/2 outputs
#pragma unroll
                    for (int it = 0; it < NIT; ++it) {
                        const int m = m_warp0 + rsub + RPI * it;
                        if (m < p.M && col + 1 < p.N) {
                            const float g0 = tanhf(t[it].x) * (1.f / (1.f + expf(-t[it].y)));
                            const long o = zoff + (long)m * (p.out16 ? p.ld16 : p.ld32) + (col >> 1);
                            const bool two = (col + 3 < p.N);
                            const float g1 = two ? tanhf(t[it].z) * (1.f / (1.f + expf(-t[it].w))) : 0.f;
                            if (p.out16) {
                                if (two) *reinterpret_cast<__half2*>(p.out16 + o) = __floats2half2_rn(g0, g1);
                                else p.out16[o] = __float2half_rn(g0);
                            }
                            if (p.out32) {
                                const long o32 = zoff + (long)m * p.ld32 + (col >> 1);
                                p.out32[o32] = g0;
                                if (two) p.out32[o32 + 1] = g1;
                            }
                        }
                    }
                    continue;
                }
                // act1 (warp-uniform switch hoisted out of the element loops)
                switch (p.act1) {
                    case ACT_NONE: break;
#define RVCB_ACT_CASE(A)                                                                        \
    case A:                                                                                     \
        _Pragma("unroll") for (int it = 0; it < NIT; ++it) {                                    \
            t[it].x = apply_act(t[it].x, A, p.act1_p); t[it].y = apply_act(t[it].y, A, p.act1_p); \
            t[it].z = apply_act(t[it].z, A, p.act1_p); t[it].w = apply_act(t[it].w, A, p.act1_p); \
        }                                                                                       \
        break;
                    RVCB_ACT_CASE(ACT_RELU) RVCB_ACT_CASE(ACT_GELU) RVCB_ACT_CASE(ACT_LRELU) RVCB_ACT_CASE(ACT_TANH) RVCB_ACT_CASE(ACT_SIGMOID)
#undef RVCB_ACT_CASE
                    default: break;
                }
                if (p.alpha != 1.f) {
#pragma unroll
                    for (int it = 0; it < NIT; ++it) { t[it].x *= p.alpha; t[it].y *= p.alpha; t[it].z *= p.alpha; t[it].w *= p.alpha; }
                }
                if (p.res2) {
#pragma unroll
                    for (int it = 0; it < NIT; ++it) {
                        const int m = m_warp0 + rsub + RPI * it;
                        if (m < p.M && col < p.N) {
                            const float* r = p.res2 + zoff + (long)m * p.ldres2 + col;
                            if (vec) {
                                const float4 q = *reinterpret_cast<const float4*>(r);
                                t[it].x += q.x; t[it].y += q.y; t[it].z += q.z; t[it].w += q.w;
                            } else {
                                t[it].x += r[0];
                                if (col + 1 < p.N) t[it].y += r[1];
                                if (col + 2 < p.N) t[it].z += r[2];
                                if (col + 3 < p.N) t[it].w += r[3];
                            }
                        }
                    }
                }
                // ---- stores (coalesced: RPI rows x CW*4 bytes per warp instruction) ----
#pragma unroll
                for (int it = 0; it < NIT; ++it) {
                    const int m = m_warp0 + rsub + RPI * it;
                    if (m >= p.M || col >= p.N) continue;
                    long orow = m, ocol = col;
                    if (p.up2_C) {
                        const int W = p.conv2d_W, Cc = p.up2_C;
                        const int ab = col / Cc, co = col - ab * Cc;
                        const int ii = m / W, jj = m - ii * W;
                        orow = (long)(2 * ii + (ab >> 1)) * (2 * W) + 2 * jj + (ab & 1);
                        ocol = co;
                    }
                    if (p.out32) {
                        float* o = p.out32 + zoff + orow * p.ld32 + ocol;
                        if (vec) {
                            *reinterpret_cast<float4*>(o) = t[it];
                        } else {
                            o[0] = t[it].x;
                            if (col + 1 < p.N) o[1] = t[it].y;
                            if (col + 2 < p.N) o[2] = t[it].z;
                            if (col + 3 < p.N) o[3] = t[it].w;
                        }
                    }
                    if (p.out16) {
                        float4 u = t[it];
                        switch (p.act2) {
                            case ACT_NONE: break;
                            case ACT_LRELU:
                                u.x = u.x > 0.f ? u.x : u.x * p.act2_p; u.y = u.y > 0.f ? u.y : u.y * p.act2_p;
                                u.z = u.z > 0.f ? u.z : u.z * p.act2_p; u.w = u.w > 0.f ? u.w : u.w * p.act2_p;
                                break;
                            default:
                                u.x = apply_act(u.x, p.act2, p.act2_p); u.y = apply_act(u.y, p.act2, p.act2_p);
                                u.z = apply_act(u.z, p.act2, p.act2_p); u.w = apply_act(u.w, p.act2, p.act2_p);
                                break;
                        }
                        __half* o = p.out16 + zoff + orow * p.ld16 + ocol;
                        if (vec) {
                            const __half2 h0 = __floats2half2_rn(u.x, u.y), h1 = __floats2half2_rn(u.z, u.w);
                            *reinterpret_cast<uint2*>(o) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
                        } else {
                            o[0] = __float2half_rn(u.x);
                            if (col + 1 < p.N) o[1] = __float2half_rn(u.y);
                            if (col + 2 < p.N) o[2] = __float2half_rn(u.z);
                            if (col + 3 < p.N) o[3] = __float2half_rn(u.w);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&tempty_bar[acc]);
                if (RS) mbar_arrive(r_empty);              // all of this warp's reads of the residual tile are done
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ------------------------------------------------------------------------------------------------
// ---- optional per-launch timing (bench.py roofline): CUDA events on the launching stream ----
static bool g_prof_on = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_events;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_prof_pool;
static std::vector<ProfInfo> g_prof_info;
void gemm_prof_begin() {
    for (auto& e : g_prof_events) g_prof_pool.push_back(e);
    g_prof_events.clear();
    g_prof_info.clear();
    g_prof_on = true;
}
void gemm_prof_end(double* ms_total, unsigned long long* launches) {
    g_prof_on = false;
    CUDA_CHECK(cudaDeviceSynchronize());
    double tot = 0;
    for (auto& e : g_prof_events) {
        float ms = 0;
        CUDA_CHECK(cudaEventElapsedTime(&ms, e.first, e.second));
        tot += ms;
    }
    if (const char* path = getenv("RVCB_PROF_CSV")) {
        if (FILE* f = fopen(path, "w")) {
            fprintf(f, "idx,M,N,kblocks,BK,BN,batch,nseg,tiles,ms,tflops\n");
            for (size_t i = 0; i < g_prof_events.size() && i < g_prof_info.size(); ++i) {
                float ms = 0;
                cudaEventElapsedTime(&ms, g_prof_events[i].first, g_prof_events[i].second);
                const ProfInfo& q = g_prof_info[i];
                const double fl = 2.0 * q.M * q.N * (double)q.kb * q.BK * q.batch;
                fprintf(f, "%zu,%d,%d,%d,%d,%d,%d,%d,%d,%.4f,%.2f\n", i, q.M, q.N, q.kb, q.BK, q.BN, q.batch, q.nseg, q.tiles, ms, fl / (ms * 1e-3) / 1e12);
            }
            fclose(f);
        }
    }
    if (ms_total) *ms_total = tot;
    if (launches) *launches = g_prof_events.size();
}
void gemm_prof_classes(double ms[3], double launches[3], double flops[3], double bytes[3]) {
    for (int c = 0; c < 3; ++c) ms[c] = launches[c] = flops[c] = bytes[c] = 0;
    for (size_t i = 0; i < g_prof_events.size() && i < g_prof_info.size(); ++i) {
        float t = 0;
        cudaEventElapsedTime(&t, g_prof_events[i].first, g_prof_events[i].second);
        const ProfInfo& q = g_prof_info[i];
        const int c = q.BN <= -1000 ? 2 : (q.BN < 0 ? 1 : 0);
        ms[c] += t;
        launches[c] += 1;
        flops[c] += 2.0 * q.M * q.N * (double)q.kb * q.BK * q.batch;
        bytes[c] += q.bytes;
    }
}
bool gemm_prof_on() { return g_prof_on; }
static std::pair<cudaEvent_t, cudaEvent_t> prof_get();
static std::pair<cudaEvent_t, cudaEvent_t> g_prof_cur;
void gemm_prof_record_begin(cudaStream_t stream) {
    g_prof_cur = prof_get();
    CUDA_CHECK(cudaEventRecord(g_prof_cur.first, stream));
}
void gemm_prof_record_end(cudaStream_t stream, const ProfInfo& info) {
    CUDA_CHECK(cudaEventRecord(g_prof_cur.second, stream));
    g_prof_events.push_back(g_prof_cur);
    g_prof_info.push_back(info);
}
static std::pair<cudaEvent_t, cudaEvent_t> prof_get() {
    if (!g_prof_pool.empty()) {
        auto e = g_prof_pool.back();
        g_prof_pool.pop_back();
        return e;
    }
    cudaEvent_t a, b;
    CUDA_CHECK(cudaEventCreate(&a));
    CUDA_CHECK(cudaEventCreate(&b));
    return {a, b};
}

template <int BN, int BK, bool RS = false>
static void launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr, const KParams& p, cudaStream_t stream) {
    using C = Cfg<BN, BK, RS>;
    static_assert(C::STAGES >= 3, "operand ring too shallow");
    static bool configured = false;
    static int num_sms = 0;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<BN, BK, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
        int dev = 0;
        CUDA_CHECK(cudaGetDevice(&dev));
        CUDA_CHECK(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
        configured = true;
    }
    const int grid = p.num_tiles < sm_budget(num_sms) ? p.num_tiles : sm_budget(num_sms);
    if (g_prof_on) gemm_prof_record_begin(stream);
    launch_pdl(gemm_tc_kernel<BN, BK, RS>, grid, kThreads, C::SMEM, stream, ta, tb, tr, p);
    KERNEL_CHECK();
    if (g_prof_on) gemm_prof_record_end(stream, {p.M, p.N, p.total_kb, BK, BN, p.batch, p.nseg, p.num_tiles});
    count_launch();
}

static int pick_bn(int N) {
    if (N <= 16) return 16;
    if (N <= 32) return 32;
    if (N <= 64) return 64;
    return 128;
}

void gemm_tc(const GemmArgs& g, cudaStream_t stream) {
    RVCB_CHECK(g.A && g.B && g.M > 0 && g.N > 0 && g.nseg > 0 && g.nseg <= GEMM_MAX_SEG, "gemm: bad arguments");
    RVCB_CHECK(g.block_k == 64 || g.block_k == 32 || g.block_k == 16, "gemm: block_k must be 16/32/64");
    RVCB_CHECK(g.out32 || g.out16, "gemm: no output");
    if (gemm_ws_try(g, stream)) return;          // long stride-1 convolutions: weight-stationary halo kernel
    if (gemm_sk_try(g, stream)) return;          // small-M, long-K launches: cluster split-K
    const int BK = g.block_k;
    int BN = pick_bn(g.N);
    if (BK == 32 && BN > 64) BN = 64;
    if (BK == 16 && BN > 32) BN = 32;
    // small grids: prefer narrower N tiles so more SMs share the operand streaming (each SM's L2->smem
    // bandwidth is the bound for few-CTA launches)
    while (BN > 32 && (long)ceil_div(g.M, BM) * ceil_div(g.N, BN) * g.batch < 96 && g.N > BN / 2) BN /= 2;

    KParams p{};
    p.M = g.M; p.N = g.N; p.nseg = g.nseg; p.batch = g.batch;
    p.num_m_tiles = ceil_div(g.M, BM);
    p.num_n_tiles = ceil_div(g.N, BN);
    p.num_tiles = p.num_m_tiles * p.num_n_tiles * g.batch;
    p.total_kb = 0;
    for (int s = 0; s < g.nseg; ++s) {
        p.seg[s].row = (short)g.seg[s].row_off;
        p.seg[s].col = (short)g.seg[s].col_off;
        p.seg[s].nk = (short)g.seg[s].nk;
        p.seg[s].dw = (signed char)g.seg[s].dw;
        p.total_kb += g.seg[s].nk;
    }
    RVCB_CHECK(p.total_kb > 0, "gemm: empty K");
    p.conv2d_W = g.conv2d_W;
    p.BH = g.conv2d_W ? BM / g.conv2d_W : 0;
    p.a_row_z = (int)g.a_row_z; p.a_col_z = (int)g.a_col_z; p.b_row_z = (int)g.b_row_z; p.b_col_z = (int)g.b_col_z;
    p.b_col0 = g.b_col0;
    p.c_z = g.c_z; p.bias_z = g.bias_z;
    p.bias = g.bias; p.bias_per_row = g.bias_per_row;
    p.res1 = g.res1; p.ldres1 = g.ldres1; p.res2 = g.res2; p.ldres2 = g.ldres2;
    p.alpha = g.alpha; p.act1 = g.act1; p.act1_p = g.act1_p; p.act2 = g.act2; p.act2_p = g.act2_p;
    p.gate = g.gate;
    p.out32 = g.out32; p.ld32 = g.ld32; p.out16 = g.out16; p.ld16 = g.ld16;
    p.up2_C = g.up2_C;
    auto al = [](const void* ptr, int a) { return ptr == nullptr || (reinterpret_cast<uintptr_t>(ptr) % a) == 0; };
    bool v = al(g.out32, 16) && al(g.out16, 8) && al(g.res1, 16) && al(g.res2, 16) && al(g.bias, 16);
    if (g.out32) v = v && (g.ld32 % 4 == 0);
    if (g.out16) v = v && (g.ld16 % 4 == 0);
    if (g.res1) v = v && (g.ldres1 % 4 == 0);
    if (g.res2) v = v && (g.ldres2 % 4 == 0);
    v = v && (g.c_z % 4 == 0) && (g.bias_z % 4 == 0);
    if (g.up2_C) v = v && (g.up2_C % 4 == 0);
    p.vec_ok = v ? 1 : 0;
    if (g.bias && !g.bias_per_row) RVCB_CHECK(al(g.bias, 16) && g.bias_z % 4 == 0, "gemm: bias must be 16B aligned");

    // ---- tensor maps ----
    CUtensorMap ta, tb;
    if (g.conv2d_W == 0) {
        cuuint64_t dims[3] = {(cuuint64_t)g.a_cols, (cuuint64_t)g.a_rows, 1};
        cuuint64_t str[2] = {(cuuint64_t)g.lda * 2, (cuuint64_t)g.lda * 2 * (cuuint64_t)g.a_rows};
        cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, 1};
        encode_map(&ta, g.A, 3, dims, str, box, BK);
    } else {
        const int W = g.conv2d_W;
        RVCB_CHECK(W >= 1 && W <= 128 && (BM % W) == 0, "gemm: conv2d W must divide 128");
        cuuint64_t dims[3] = {(cuuint64_t)g.a_cols, (cuuint64_t)W, (cuuint64_t)g.a_rows};
        cuuint64_t str[2] = {(cuuint64_t)g.lda * 2, (cuuint64_t)g.lda * 2 * (cuuint64_t)W};
        cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)W, (cuuint32_t)(BM / W)};
        encode_map(&ta, g.A, 3, dims, str, box, BK);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)g.b_cols, (cuuint64_t)g.b_rows};
        cuuint64_t str[1] = {(cuuint64_t)g.ldb * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
        encode_map(&tb, g.B, 2, dims, str, box, BK);
    }

    // large launches with an fp32 residual can prefetch the residual tile by TMA (see Cfg::RES_BYTES).  Measured (profiles/README.md,
    // r1v): neutral on the k = 11 stage-1 convolutions, slower on stage 0 (the 64 KB tile costs two operand-ring stages), so
    // it is opt-in (RVCB_RS=1) and exercised by tests/test_gemm_gpu.py only.
    static const bool rs_on = [] { const char* e = getenv("RVCB_RS"); return e && e[0] == '1'; }();
    const bool rs = rs_on && g.res1 && !g.gate && !g.up2_C && g.batch == 1 && BK == 64 && (BN == 128 || BN == 64) && (g.N % BN) == 0 &&
                    g.M >= 8192 && (g.ldres1 % 4) == 0 && al(g.res1, 16);
    CUtensorMap tr = ta;
    if (rs) {
        cuuint64_t dims[2] = {(cuuint64_t)g.N, (cuuint64_t)g.M};
        cuuint64_t str[1] = {(cuuint64_t)g.ldres1 * 4};
        cuuint32_t box[2] = {32u, (cuuint32_t)BM};
        encode_map_ex(&tr, g.res1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (BN == 128) { launch<128, 64, true>(ta, tb, tr, p, stream); return; }
        launch<64, 64, true>(ta, tb, tr, p, stream);
        return;
    }
#define RVCB_LAUNCH(bn, bk)                          \
    if (BN == bn && BK == bk) {                      \
        launch<bn, bk>(ta, tb, tr, p, stream);       \
        return;                                      \
    }
    RVCB_LAUNCH(16, 64) RVCB_LAUNCH(32, 64) RVCB_LAUNCH(64, 64) RVCB_LAUNCH(128, 64)
    RVCB_LAUNCH(16, 32) RVCB_LAUNCH(32, 32) RVCB_LAUNCH(64, 32)
    RVCB_LAUNCH(16, 16) RVCB_LAUNCH(32, 16)
#undef RVCB_LAUNCH
    RVCB_CHECK(false, "gemm: no kernel instance");
}

}  // namespace rvcb
