// Weight-stationary, halo-streaming convolution kernel (tcgen05 + TMA) for the long stride-1 Conv1d layers of
// the NSF-HiFi-GAN vocoder (rvc/layers/residuals.py:68-85, nsf.py:145-191): M = time (up to 767 040 rows),
// N = C_out in {32, 64, 128}, K = C_in * k.
//
// The streaming kernel (gemm_tc.cu) re-fetches the A tile once per tap and the whole weight matrix once per M-tile;
// those L2->SMEM bytes, not the tensor pipe, bound it on these layers.  Here instead:
//   * each persistent CTA owns one N-slice of the output and keeps that slice of the packed weights RESIDENT in
//     shared memory for the whole launch (loaded once by TMA);
//   * per M-tile it streams ONE halo tile of activations per channel chunk: rows [m0 + r_min, m0 + r_max + 128),
//     double-buffered; every tap's A operand is the same smem tile at a row-shifted UMMA descriptor
//     (zero padding = TMA out-of-bounds fill) -- activations are read once per N-slice, not once per tap;
//   * the epilogue warps prefetch the residual operands of their tile BEFORE waiting for the accumulator, so the
//     HBM latency of the fp32 residual stream hides under the MMAs.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2..9 epilogue (two groups).
// Two kernels share that main loop: gemm_ws_kernel (v1: register prefetch of the residual, generic epilogue; used where the
// TMA staging of v2 does not fit, e.g. C = 128 with k = 7) and gemm_ws2_kernel (v2, further down: every byte of epilogue
// traffic moves by TMA; the vocoder's three epilogue patterns).
#include "tc_common.cuh"

namespace rvcb {

struct WSParams {
    KParams k;              // epilogue / shape fields reused (seg[] holds the taps)
    int rmin, HRp, nkc, n_slices, m_tiles, bo_mode;
    int has_r2, do16;       // ws2 only: second residual tile / fp16 output present
};

template <int BN, int BK>
struct WSCfg {
    static constexpr int CW = BN >= 32 ? 32 : 16;
    static constexpr int EPI_STRIDE = CW + 4;
    static constexpr int EPI_BYTES = kEpiWarps * 32 * EPI_STRIDE * 4;
    static constexpr int B_KB_BYTES = BN * BK * 2;                 // one K block of the resident weight slice
    static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : 128;
};

template <int BK>
__device__ __forceinline__ uint64_t make_desc_shifted(uint32_t smem_addr, int bo_mode) {
    constexpr uint64_t layout = (BK == 64) ? 2ull : (BK == 32) ? 4ull : 6ull;
    constexpr uint64_t sbo = (8 * BK * 2) >> 4;
    uint64_t d = (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
    // Measured on B200 (tests/test_gemm_gpu.py::test_conv1d_weight_stationary_halo_kernel): the UMMA swizzle is a function of
    // the ABSOLUTE shared-memory address, so a row-shifted start needs base_offset = 0; setting (addr >> 7) & 7 gives wrong results.
    if (bo_mode) d |= (uint64_t)((smem_addr >> 7) & 7) << 49;
    return d;
}

// EPI: compile-time epilogue specialisation (the generic code is instruction/latency bound with only 8 epilogue warps)
//   0  out16 = lrelu(acc + bias)                                   (resblock c1)
//   1  y = acc + bias + res1 ; out32 = y ; out16 = lrelu(y)         (resblock c2)
//   2  y = alpha*(acc + bias + res1) (+ res2) ; out32 = y ; (out16 = lrelu(y))   (last c2 of a resblock: 3-branch mean)
//   3  generic contract
template <int BN, int BK, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_ws_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ WSParams wp) {
    using C = WSCfg<BN, BK>;
    const KParams& p = wp.k;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int total_kb = p.nseg * wp.nkc;
    const int b_bytes = total_kb * C::B_KB_BYTES;
    const int a_chunk = wp.HRp * BK * 2;                         // one channel chunk of the halo tile
    const int a_buf = ((wp.nkc * a_chunk + 1023) / 1024) * 1024;
    uint8_t* smem_b = smem;
    uint8_t* smem_a = smem + ((b_bytes + 1023) / 1024) * 1024;
    uint8_t* tail = smem_a + 2 * a_buf;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    uint64_t* b_full = bars;            // [1]
    uint64_t* a_full = bars + 1;        // [2]
    uint64_t* a_empty = bars + 3;       // [2]
    uint64_t* tfull_bar = bars + 5;     // [2]
    uint64_t* tempty_bar = bars + 7;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);
    float* epi_smem = reinterpret_cast<float*>(tail + 128);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        mbar_init(b_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], (BN / C::CW == 1) ? kEpiWarps / 2 : kEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();

    const int slice = blockIdx.x % wp.n_slices;
    const int mt0 = blockIdx.x / wp.n_slices;
    const int mt_step = gridDim.x / wp.n_slices;

    if (warp == 0) {
        // ======================= TMA producer (warp-uniform loop, one elected lane issues) =======================
        if (elect_one()) {
            mbar_expect_tx(b_full, (uint32_t)b_bytes);
            for (int kb = 0; kb < total_kb; ++kb)
                tma_load_2d(smem_b + kb * C::B_KB_BYTES, &tmap_b, b_full, kb * BK, slice * BN);
        }
        __syncwarp();
        int it = 0;
        for (int mt = mt0; mt < wp.m_tiles; mt += mt_step, ++it) {
            const int buf = it & 1;
            mbar_wait(&a_empty[buf], ((it >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&a_full[buf], (uint32_t)(wp.nkc * a_chunk));
                for (int kc = 0; kc < wp.nkc; ++kc)
                    tma_load_3d(smem_a + buf * a_buf + kc * a_chunk, &tmap_a, &a_full[buf], kc * BK, mt * BM + wp.rmin, 0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ======================= MMA issuer (warp-uniform loop, one elected lane issues) =======================
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        mbar_wait(b_full, 0);
        int it = 0;
        for (int mt = mt0; mt < wp.m_tiles; mt += mt_step, ++it) {
            const int buf = it & 1, acc = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            mbar_wait(&tempty_bar[acc], ph ^ 1);
            mbar_wait(&a_full[buf], ph);
            tc_fence_after();
            const uint32_t tmem_c = tmem_base + acc * BN;
            const uint32_t a_base = smem_u32(smem_a + buf * a_buf);
            const uint32_t b_base = smem_u32(smem_b);
            uint32_t first = 0;
            for (int kc = 0; kc < wp.nkc; ++kc) {
                for (int j = 0; j < p.nseg; ++j) {
                    const uint32_t a_addr = a_base + kc * a_chunk + (uint32_t)(p.seg[j].row - wp.rmin) * (BK * 2);
                    const uint32_t b_addr = b_base + (uint32_t)(j * wp.nkc + kc) * C::B_KB_BYTES;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t da = make_desc_shifted<BK>(a_addr + k * 32, wp.bo_mode);
                            const uint64_t db = make_kmajor_desc<BK>(b_addr + k * 32);
                            umma_f16(tmem_c, da, db, idesc, first | (uint32_t)k);
                        }
                    }
                    first = 1;
                }
            }
            if (elect_one()) {
                umma_commit(&a_empty[buf]);       // halo buffer free once these MMAs retire
                umma_commit(&tfull_bar[acc]);     // accumulator complete
            }
            __syncwarp();
        }
    } else {
        // ======================= epilogue (8 warps, two groups) =======================
        constexpr int CW = C::CW, ST = C::EPI_STRIDE, NV = CW / 4;
        constexpr int RPI = 32 / NV, NIT = 32 / RPI;
        constexpr bool kAlt = (BN / CW == 1);
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int cgroup = ew >> 2;
        float* stg = epi_smem + ew * 32 * ST;
        const int c4 = lane % NV, rsub = lane / NV;
        const int c_start = kAlt ? 0 : cgroup;
        constexpr int c_step = kAlt ? 1 : 2;
        int it = 0;
        for (int mt = mt0; mt < wp.m_tiles; mt += mt_step, ++it) {
            if (kAlt && (it & 1) != cgroup) continue;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int m_warp0 = mt * BM + quarter * 32;
            const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(quarter * 32) << 16);
            bool waited = false;
#pragma unroll 1
            for (int c = c_start; c < BN / CW; c += c_step) {
                const int n0 = slice * BN + c * CW;
                if (n0 >= p.N) break;
                const int col = n0 + 4 * c4;
                if (EPI != 3 && m_warp0 + 32 <= p.M && n0 + CW <= p.N) {
                    // ================= specialised fast path: full 32-row slab, aligned, vectorised =================
                    float4 r1[NIT], r2[NIT];
                    const long rbase = (long)(m_warp0 + rsub) * p.ldres1 + col;
                    if (EPI >= 1) {
#pragma unroll
                        for (int i = 0; i < NIT; ++i) r1[i] = *reinterpret_cast<const float4*>(p.res1 + rbase + (long)(RPI * i) * p.ldres1);
                    }
                    const bool has_r2 = (EPI == 2) && p.res2 != nullptr;
                    if (has_r2) {
                        const long r2base = (long)(m_warp0 + rsub) * p.ldres2 + col;
#pragma unroll
                        for (int i = 0; i < NIT; ++i) r2[i] = *reinterpret_cast<const float4*>(p.res2 + r2base + (long)(RPI * i) * p.ldres2);
                    }
                    if (!waited) {
                        mbar_wait(&tfull_bar[acc], acc_phase);
                        tc_fence_after();
                        waited = true;
                    }
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
#pragma unroll
                    for (int i = 0; i < CW; i += 4) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
                        v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < CW; i += 4)
                        *reinterpret_cast<float4*>(stg + lane * ST + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    __syncwarp();
                    const float slope = p.act2_p;
                    const bool do16 = (EPI < 2) || p.out16 != nullptr;
#pragma unroll
                    for (int i = 0; i < NIT; ++i) {
                        const int m = m_warp0 + rsub + RPI * i;
                        float4 t = *reinterpret_cast<const float4*>(stg + (rsub + RPI * i) * ST + 4 * c4);
                        if (EPI >= 1) { t.x += r1[i].x; t.y += r1[i].y; t.z += r1[i].z; t.w += r1[i].w; }
                        if (EPI == 2) {
                            t.x *= p.alpha; t.y *= p.alpha; t.z *= p.alpha; t.w *= p.alpha;
                            if (has_r2) { t.x += r2[i].x; t.y += r2[i].y; t.z += r2[i].z; t.w += r2[i].w; }
                        }
                        if (EPI >= 1) *reinterpret_cast<float4*>(p.out32 + (long)m * p.ld32 + col) = t;
                        if (do16) {
                            t.x = t.x > 0.f ? t.x : t.x * slope; t.y = t.y > 0.f ? t.y : t.y * slope;
                            t.z = t.z > 0.f ? t.z : t.z * slope; t.w = t.w > 0.f ? t.w : t.w * slope;
                            const __half2 h0 = __floats2half2_rn(t.x, t.y), h1 = __floats2half2_rn(t.z, t.w);
                            *reinterpret_cast<uint2*>(p.out16 + (long)m * p.ld16 + col) =
                                make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
                        }
                    }
                    continue;
                }
                const bool vec = (col + 3 < p.N) && p.vec_ok;
                // ---- prefetch the residual operands (independent of the accumulator) ----
                float4 r1[NIT], r2[NIT];
#pragma unroll
                for (int i = 0; i < NIT; ++i) {
                    r1[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    r2[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    const int m = m_warp0 + rsub + RPI * i;
                    if (m < p.M && col < p.N) {
                        if (p.res1) {
                            const float* r = p.res1 + (long)m * p.ldres1 + col;
                            if (vec) r1[i] = *reinterpret_cast<const float4*>(r);
                            else {
                                r1[i].x = r[0];
                                if (col + 1 < p.N) r1[i].y = r[1];
                                if (col + 2 < p.N) r1[i].z = r[2];
                                if (col + 3 < p.N) r1[i].w = r[3];
                            }
                        }
                        if (p.res2) {
                            const float* r = p.res2 + (long)m * p.ldres2 + col;
                            if (vec) r2[i] = *reinterpret_cast<const float4*>(r);
                            else {
                                r2[i].x = r[0];
                                if (col + 1 < p.N) r2[i].y = r[1];
                                if (col + 2 < p.N) r2[i].z = r[2];
                                if (col + 3 < p.N) r2[i].w = r[3];
                            }
                        }
                    }
                }
                if (!waited) {
                    mbar_wait(&tfull_bar[acc], acc_phase);
                    tc_fence_after();
                    waited = true;
                }
                // ---- TMEM -> registers (thread = row) + bias ----
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
                    if (n0 + CW <= p.N) {
#pragma unroll
                        for (int i = 0; i < CW; i += 4) {
                            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + i));
                            v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < CW; ++i)
                            if (n0 + i < p.N) v[i] += __ldg(p.bias + n0 + i);
                    }
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < CW; i += 4)
                    *reinterpret_cast<float4*>(stg + lane * ST + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < NIT; ++i) {
                    const int m = m_warp0 + rsub + RPI * i;
                    float4 t = *reinterpret_cast<const float4*>(stg + (rsub + RPI * i) * ST + 4 * c4);
                    if (m >= p.M || col >= p.N) continue;
                    t.x += r1[i].x; t.y += r1[i].y; t.z += r1[i].z; t.w += r1[i].w;
                    if (p.act1 != ACT_NONE) {
                        t.x = apply_act(t.x, p.act1, p.act1_p); t.y = apply_act(t.y, p.act1, p.act1_p);
                        t.z = apply_act(t.z, p.act1, p.act1_p); t.w = apply_act(t.w, p.act1, p.act1_p);
                    }
                    t.x = fmaf(t.x, p.alpha, r2[i].x); t.y = fmaf(t.y, p.alpha, r2[i].y);
                    t.z = fmaf(t.z, p.alpha, r2[i].z); t.w = fmaf(t.w, p.alpha, r2[i].w);
                    if (p.out32) {
                        float* o = p.out32 + (long)m * p.ld32 + col;
                        if (vec) *reinterpret_cast<float4*>(o) = t;
                        else {
                            o[0] = t.x;
                            if (col + 1 < p.N) o[1] = t.y;
                            if (col + 2 < p.N) o[2] = t.z;
                            if (col + 3 < p.N) o[3] = t.w;
                        }
                    }
                    if (p.out16) {
                        float4 u = t;
                        if (p.act2 == ACT_LRELU) {
                            u.x = u.x > 0.f ? u.x : u.x * p.act2_p; u.y = u.y > 0.f ? u.y : u.y * p.act2_p;
                            u.z = u.z > 0.f ? u.z : u.z * p.act2_p; u.w = u.w > 0.f ? u.w : u.w * p.act2_p;
                        } else if (p.act2 != ACT_NONE) {
                            u.x = apply_act(u.x, p.act2, p.act2_p); u.y = apply_act(u.y, p.act2, p.act2_p);
                            u.z = apply_act(u.z, p.act2, p.act2_p); u.w = apply_act(u.w, p.act2, p.act2_p);
                        }
                        __half* o = p.out16 + (long)m * p.ld16 + col;
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
            if (!waited) mbar_wait(&tfull_bar[acc], acc_phase);      // keep the barrier protocol in lock-step even with no columns
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
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
// v2: same weight-stationary / halo-streaming main loop, but every byte of the epilogue's global traffic moves by TMA.
//   * the producer prefetches the fp32 residual tile(s) of tile i+1 into swizzled shared memory while tile i computes;
//   * the 8 epilogue warps go TMEM -> registers -> (+bias, +residual from smem) -> write y (fp32, in place over the residual
//     tile) and lrelu(y) (fp16) back to shared memory -- no global loads or stores, no exposed HBM latency;
//   * the producer warp (lane 0) stores each finished tile (cp.async.bulk.tensor, clipped at M), drains the stage and reuses
//     it for the next residual prefetch; the epilogue warps never wait on a store and need no block-wide barrier.
// Staging tiles use the TMA 128 B / 64 B swizzle, so "thread = row" 16-byte accesses are bank-conflict free.
// EPI as above (0: c1, 1: c2, 2: last c2 of a resblock); the generic contract stays on the v1 kernel.
// ------------------------------------------------------------------------------------------------
template <int BN>
struct WS2Cfg {
    static constexpr int W = BN / 2;                       // columns per epilogue warp (4 lane quadrants x 2 column halves)
    static constexpr int NPAN = BN / 32;                   // fp32 panels of 32 columns (128-byte rows, SW128)
    static constexpr int R_BYTES = BM * BN * 4;
    static constexpr int H_BYTES = BM * BN * 2;            // fp16 tile: BN = 64 -> 128-byte rows (SW128), BN = 32 -> 64-byte rows (SW64)
    static constexpr int TMEM_COLS = (2 * BN <= 64) ? 64 : 128;
    static constexpr int NS = 2;                           // staging stages
};

template <int BN, int BK, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_ws2_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_r1, const __grid_constant__ CUtensorMap tmap_r2,
                const __grid_constant__ CUtensorMap tmap_o32, const __grid_constant__ CUtensorMap tmap_o16,
                const __grid_constant__ WSParams wp) {
    using C = WS2Cfg<BN>;
    constexpr int B_KB_BYTES = BN * BK * 2;
    const KParams& p = wp.k;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int total_kb = p.nseg * wp.nkc;
    const int b_bytes = total_kb * B_KB_BYTES;
    const int a_chunk = wp.HRp * BK * 2;
    const int a_buf = ((wp.nkc * a_chunk + 1023) / 1024) * 1024;
    const bool has_r2 = (EPI == 2) && wp.has_r2;
    const bool do16 = (EPI < 2) || wp.do16;
    uint8_t* smem_b = smem;
    uint8_t* smem_a = smem + ((b_bytes + 1023) / 1024) * 1024;
    uint8_t* smem_r1 = smem_a + 2 * a_buf;
    uint8_t* smem_r2 = smem_r1 + (EPI >= 1 ? C::NS * C::R_BYTES : 0);
    uint8_t* smem_h = smem_r2 + (has_r2 ? C::NS * C::R_BYTES : 0);
    uint8_t* tail = smem_h + (do16 ? C::NS * C::H_BYTES : 0);
    uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
    uint64_t* b_full = bars;            // [1]
    uint64_t* a_full = bars + 1;        // [2]
    uint64_t* a_empty = bars + 3;       // [2]
    uint64_t* tfull_bar = bars + 5;     // [2]
    uint64_t* tempty_bar = bars + 7;    // [2]
    uint64_t* r_full = bars + 9;        // [2] residual tile(s) landed
    uint64_t* r_empty = bars + 11;      // [2] staging stage drained by the TMA stores (EPI 0 only: nothing else tells the epilogue)
    uint64_t* s_full = bars + 13;       // [2] epilogue warps finished writing the staging stage
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        if (EPI >= 1) { prefetch_tmap(&tmap_r1); prefetch_tmap(&tmap_o32); }
        if (has_r2) prefetch_tmap(&tmap_r2);
        if (do16) prefetch_tmap(&tmap_o16);
        mbar_init(b_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_full[i], 1);
            mbar_init(&a_empty[i], 1);
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], kEpiWarps);
            mbar_init(&r_full[i], 1);
            mbar_init(&r_empty[i], 1);
            mbar_init(&s_full[i], kEpiWarps);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();

    const int slice = blockIdx.x % wp.n_slices;
    const int mt0 = blockIdx.x / wp.n_slices;
    const int mt_step = gridDim.x / wp.n_slices;

    if (warp == 0) {
        // ======================= TMA producer + store issuer =======================
        // Warp-uniform loop; lane 0 issues (bulk async groups are per thread, so the same lane must issue and drain the stores).
        // Iteration `it`: prefetch the halo tile of tile it+1; load the residual of tile `it` (its stage was drained at the end
        // of the previous iteration); then, once the epilogue warps have written tile it-1, store it and drain the stage.
        const int n_it = (wp.m_tiles - mt0 + mt_step - 1) / mt_step;
        auto load_a = [&](int it) {
            const int buf = it & 1;
            mbar_wait(&a_empty[buf], ((it >> 1) & 1) ^ 1);
            if (lane == 0) {
                mbar_expect_tx(&a_full[buf], (uint32_t)(wp.nkc * a_chunk));
                for (int kc = 0; kc < wp.nkc; ++kc)
                    tma_load_3d(smem_a + buf * a_buf + kc * a_chunk, &tmap_a, &a_full[buf], kc * BK, (mt0 + it * mt_step) * BM + wp.rmin, 0);
            }
            __syncwarp();
        };
        auto store_tile = [&](int it) {
            const int s = it & 1;
            const int m0 = (mt0 + it * mt_step) * BM;
            mbar_wait(&s_full[s], (it >> 1) & 1);
            if (lane == 0) {
                if (EPI >= 1) {
#pragma unroll
                    for (int pn = 0; pn < C::NPAN; ++pn)
                        tma_store_2d(&tmap_o32, smem_r1 + s * C::R_BYTES + pn * (BM * 128), slice * BN + pn * 32, m0);
                }
                if (do16) tma_store_2d(&tmap_o16, smem_h + s * C::H_BYTES, slice * BN, m0);
                bulk_commit();
                bulk_wait_read0();                          // the stores have read the stage
                if (EPI == 0) mbar_arrive(&r_empty[s]);     // EPI >= 1: the next residual load into this stage is the signal
            }
            __syncwarp();
        };
        if (lane == 0) {
            mbar_expect_tx(b_full, (uint32_t)b_bytes);
            for (int kb = 0; kb < total_kb; ++kb)
                tma_load_2d(smem_b + kb * B_KB_BYTES, &tmap_b, b_full, kb * BK, slice * BN);
        }
        __syncwarp();
        if (n_it > 0) load_a(0);
        for (int it = 0; it < n_it; ++it) {
            if (it + 1 < n_it) load_a(it + 1);
            if (EPI >= 1) {
                const int s = it & 1;
                if (lane == 0) {
                    const int m0 = (mt0 + it * mt_step) * BM;
                    mbar_expect_tx(&r_full[s], (uint32_t)(C::R_BYTES * (has_r2 ? 2 : 1)));
#pragma unroll
                    for (int pn = 0; pn < C::NPAN; ++pn) {
                        tma_load_2d(smem_r1 + s * C::R_BYTES + pn * (BM * 128), &tmap_r1, &r_full[s], slice * BN + pn * 32, m0);
                        if (has_r2) tma_load_2d(smem_r2 + s * C::R_BYTES + pn * (BM * 128), &tmap_r2, &r_full[s], slice * BN + pn * 32, m0);
                    }
                }
                __syncwarp();
            }
            if (it >= 1) store_tile(it - 1);
        }
        if (n_it > 0) store_tile(n_it - 1);
        if (lane == 0) bulk_wait0();
        __syncwarp();
    } else if (warp == 1) {
        // ======================= MMA issuer (warp-uniform loop, one elected lane issues) =======================
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
        // descriptor = {lo: start address >> 4 | LBO(1) << 16, hi: SBO | version | layout}; K step = +32 B = +2 in lo
        constexpr uint32_t desc_hi = (uint32_t)(((uint64_t)((8 * BK * 2) >> 4) << 32 | (1ull << 46) |
                                                 ((BK == 64 ? 2ull : BK == 32 ? 4ull : 6ull) << 61)) >> 32);
        mbar_wait(b_full, 0);
        const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFF) >> 4) | (1u << 16);
        const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFF) >> 4) | (1u << 16);
        int it = 0;
        for (int mt = mt0; mt < wp.m_tiles; mt += mt_step, ++it) {
            const int buf = it & 1, acc = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            mbar_wait(&tempty_bar[acc], ph ^ 1);
            mbar_wait(&a_full[buf], ph);
            tc_fence_after();
            const uint32_t tmem_c = tmem_base + acc * BN;
            uint32_t first = 0;
            for (int kc = 0; kc < wp.nkc; ++kc) {
                const uint32_t a_kc = a_lo0 + (uint32_t)((buf * a_buf + kc * a_chunk) >> 4);
                for (int j = 0; j < p.nseg; ++j) {
                    const uint32_t a_lo = a_kc + (uint32_t)(p.seg[j].row - wp.rmin) * (uint32_t)(BK * 2 / 16);
                    const uint32_t b_lo = b_lo0 + (uint32_t)(j * wp.nkc + kc) * (uint32_t)(B_KB_BYTES / 16);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < BK / 16; ++k) {
                            const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2 * k);
                            const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2 * k);
                            umma_f16(tmem_c, da, db, idesc, first | (uint32_t)k);
                        }
                    }
                    first = 1;
                }
            }
            if (elect_one()) {
                umma_commit(&a_empty[buf]);
                umma_commit(&tfull_bar[acc]);
            }
            __syncwarp();
        }
    } else {
        // ======================= epilogue: 8 warps = 4 TMEM lane quadrants x 2 column halves =======================
        constexpr int W = C::W;
        const int ew = warp - 2;
        const int quarter = warp & 3;
        const int chalf = ew >> 2;
        const int r = quarter * 32 + lane;                 // tile row owned by this thread
        const int c0 = chalf * W;                          // first tile column owned by this thread
        // swizzled 16-byte chunk addressing (see header comment)
        const uint32_t r_off = (uint32_t)((c0 >> 5) * (BM * 128) + r * 128);
        const int r_cb = (c0 & 31) >> 2;
        const uint32_t h_off = (uint32_t)(r * (BN * 2));
        const int h_cb = c0 >> 3;
        const int h_x = (BN == 64) ? (r & 7) : ((r >> 1) & 3);
        float bias_r[W];
#pragma unroll
        for (int i = 0; i < W; i += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + slice * BN + c0 + i));
            bias_r[i] = b.x; bias_r[i + 1] = b.y; bias_r[i + 2] = b.z; bias_r[i + 3] = b.w;
        }
        const float slope = p.act2_p, alpha = p.alpha;
        int it = 0;
        for (int mt = mt0; mt < wp.m_tiles; mt += mt_step, ++it) {
            const int s = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            if (EPI >= 1) mbar_wait(&r_full[s], ph);       // residual landed (the producer waited for the stage to drain)
            else mbar_wait(&r_empty[s], ph ^ 1);           // staging stage drained
            mbar_wait(&tfull_bar[s], ph);
            tc_fence_after();
            const uint32_t taddr = tmem_base + s * BN + c0 + ((uint32_t)(quarter * 32) << 16);
            float v[W];
            {
                uint32_t raw[16];
                tmem_ld16(taddr, raw);
                if constexpr (W == 32) {
                    uint32_t raw2[16];
                    tmem_ld16(taddr + 16, raw2);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[16 + i] = __uint_as_float(raw2[i]);
                } else {
                    tmem_ld_wait();
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[s]);    // accumulator stage free as soon as it is in registers
            uint8_t* R1 = smem_r1 + s * C::R_BYTES + r_off;
            uint8_t* R2 = smem_r2 + s * C::R_BYTES + r_off;
            uint8_t* Hs = smem_h + s * C::H_BYTES + h_off;
#pragma unroll
            for (int i = 0; i < W / 4; ++i) {
                float4 t = make_float4(v[4 * i] + bias_r[4 * i], v[4 * i + 1] + bias_r[4 * i + 1], v[4 * i + 2] + bias_r[4 * i + 2],
                                       v[4 * i + 3] + bias_r[4 * i + 3]);
                if (EPI >= 1) {
                    float4* q = reinterpret_cast<float4*>(R1 + (((r_cb + i) ^ (r & 7)) << 4));
                    const float4 a = *q;
                    t.x += a.x; t.y += a.y; t.z += a.z; t.w += a.w;
                    if (EPI == 2) {
                        t.x *= alpha; t.y *= alpha; t.z *= alpha; t.w *= alpha;
                        if (has_r2) {
                            const float4 b = *reinterpret_cast<const float4*>(R2 + (((r_cb + i) ^ (r & 7)) << 4));
                            t.x += b.x; t.y += b.y; t.z += b.z; t.w += b.w;
                        }
                    }
                    *q = t;                                 // y, in place: this tile is stored to out32
                }
                v[4 * i] = t.x > 0.f ? t.x : t.x * slope; v[4 * i + 1] = t.y > 0.f ? t.y : t.y * slope;
                v[4 * i + 2] = t.z > 0.f ? t.z : t.z * slope; v[4 * i + 3] = t.w > 0.f ? t.w : t.w * slope;
            }
            if (do16) {
#pragma unroll
                for (int i = 0; i < W / 8; ++i) {
                    const __half2 h0 = __floats2half2_rn(v[8 * i], v[8 * i + 1]), h1 = __floats2half2_rn(v[8 * i + 2], v[8 * i + 3]);
                    const __half2 h2 = __floats2half2_rn(v[8 * i + 4], v[8 * i + 5]), h3 = __floats2half2_rn(v[8 * i + 6], v[8 * i + 7]);
                    *reinterpret_cast<uint4*>(Hs + (((h_cb + i) ^ h_x) << 4)) =
                        make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                   *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
                }
            }
            fence_proxy_async();                           // generic-proxy smem writes -> visible to the TMA store
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_full[s]);        // warp 0 stores the tile once all 8 warps have arrived
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<C::TMEM_COLS>(tmem_base);
    }
}

static size_t ws2_smem_bytes(int BN, int BK, int nseg, int nkc, int HRp, int epi, bool has_r2, bool do16) {
    const size_t b = ((size_t)nseg * nkc * BN * BK * 2 + 1023) / 1024 * 1024;
    const size_t a = ((size_t)nkc * HRp * BK * 2 + 1023) / 1024 * 1024;
    size_t st = 0;
    if (epi >= 1) st += 2 * (size_t)BM * BN * 4;
    if (epi == 2 && has_r2) st += 2 * (size_t)BM * BN * 4;
    if (epi < 2 || do16) st += 2 * (size_t)BM * BN * 2;
    return b + 2 * a + st + 256 + 1024;
}

template <int BN, int BK, int EPI>
static void ws2_launch_e(const CUtensorMap* tm, const WSParams& wp, int grid, size_t smem, cudaStream_t stream) {
    static size_t configured = 0;
    if (smem > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(gemm_ws2_kernel<BN, BK, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    launch_pdl(gemm_ws2_kernel<BN, BK, EPI>, grid, kThreads, smem, stream, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], wp);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
template <int BN, int BK, int EPI>
static void ws_launch_e(const CUtensorMap& ta, const CUtensorMap& tb, const WSParams& wp, int grid, size_t smem, cudaStream_t stream) {
    static size_t configured = 0;
    if (smem > configured) {
        CUDA_CHECK(cudaFuncSetAttribute(gemm_ws_kernel<BN, BK, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    launch_pdl(gemm_ws_kernel<BN, BK, EPI>, grid, kThreads, smem, stream, ta, tb, wp);
}

template <int BN, int BK>
static void ws_launch(const CUtensorMap& ta, const CUtensorMap& tb, const WSParams& wp, int grid, size_t smem, cudaStream_t stream, int epi) {
    if (gemm_prof_on()) gemm_prof_record_begin(stream);
    switch (epi) {
        case 0: ws_launch_e<BN, BK, 0>(ta, tb, wp, grid, smem, stream); break;
        case 1: ws_launch_e<BN, BK, 1>(ta, tb, wp, grid, smem, stream); break;
        case 2: ws_launch_e<BN, BK, 2>(ta, tb, wp, grid, smem, stream); break;
        default: ws_launch_e<BN, BK, 3>(ta, tb, wp, grid, smem, stream); break;
    }
    KERNEL_CHECK();
    if (gemm_prof_on()) {
        // algorithmic HBM bytes of this launch: activations in once, weights once, residual(s) in, outputs out
        const KParams& q = wp.k;
        const double mn = (double)q.M * q.N;
        const double bytes = (double)q.M * wp.nkc * BK * 2 + (double)q.N * q.nseg * wp.nkc * BK * 2 + (q.res1 ? mn * 4 : 0) +
                             (q.res2 ? mn * 4 : 0) + (q.out32 ? mn * 4 : 0) + (q.out16 ? mn * 2 : 0);
        ProfInfo info{q.M, q.N, q.nseg * wp.nkc, BK, -BN /*negative = WS kernel*/, 1, q.nseg, wp.m_tiles * wp.n_slices};
        info.bytes = bytes;
        gemm_prof_record_end(stream, info);
    }
    count_launch();
}

static size_t ws_smem_bytes(int BN, int BK, int nseg, int nkc, int HRp) {
    const int CW = BN >= 32 ? 32 : 16;
    const size_t epi = (size_t)kEpiWarps * 32 * (CW + 4) * 4;
    const size_t b = ((size_t)nseg * nkc * BN * BK * 2 + 1023) / 1024 * 1024;
    const size_t a = ((size_t)nkc * HRp * BK * 2 + 1023) / 1024 * 1024;
    return b + 2 * a + 128 + epi + 1024;
}

// v2 dispatch: the three vocoder epilogue patterns with TMA-legal operands.  Returns false to fall through to v1.
static bool ws2_try(const GemmArgs& g, cudaStream_t stream, int BK, int nkc, int rmin, int HRp, int m_tiles, int sms, int bo_mode) {
    static int mode = -1, n32 = 0, n32_off = 0;
    if (mode < 0) {
        const char* e = getenv("RVCB_WS2");
        mode = (e && e[0] == '0') ? 0 : 1;
        const char* f = getenv("RVCB_WS2_N32");      // 1: narrow slices wherever they fit; 0: never; unset: residual launches only
        n32 = (f && f[0] == '1') ? 1 : 0;
        n32_off = (f && f[0] == '0') ? 1 : 0;
    }
    if (!mode) return false;
    auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    if (!g.bias || !al16(g.bias) || g.act1 != ACT_NONE) return false;
    int epi = -1;
    if (!g.res1 && !g.res2 && !g.out32 && g.out16 && g.act2 == ACT_LRELU && g.alpha == 1.f) epi = 0;
    else if (g.res1 && !g.res2 && g.out32 && g.out16 && g.act2 == ACT_LRELU && g.alpha == 1.f) epi = 1;
    else if (g.res1 && g.out32 && (!g.out16 || g.act2 == ACT_LRELU)) epi = 2;
    if (epi < 0) return false;
    if (g.out16 && (!al16(g.out16) || g.ld16 % 8)) return false;
    if (g.out32 && (!al16(g.out32) || g.ld32 % 4)) return false;
    if (g.res1 && (!al16(g.res1) || g.ldres1 % 4)) return false;
    if (g.res2 && (!al16(g.res2) || g.ldres2 % 4)) return false;
    const bool has_r2 = g.res2 != nullptr, do16 = g.out16 != nullptr;
    int BN = 0;
    for (int cand : {64, 32}) {
        if (BK == 32 && cand > 32) continue;
        if (g.N % cand) continue;
        if (ws2_smem_bytes(cand, BK, g.nseg, nkc, HRp, epi, has_r2, do16) <= 226 * 1024) { BN = cand; break; }
    }
    if (BN == 0) return false;
    // Slices narrower than 64 columns over a multi-chunk C_in re-read the halo tile once per slice and run N = 32 MMAs: measured
    // slower than the alternatives (k = 7, C = 128: 84 vs 53 us) EXCEPT for residual launches that fit no 64-column variant and
    // would fall to the streaming kernel, whose epilogue is the bottleneck there (k = 11, C = 128, c2: 119 vs 178 us).
    if (BN < 64 && BN < g.N && nkc > 1 && !n32) {
        const bool v1_fits = ws_smem_bytes(64, BK, g.nseg, nkc, HRp) <= 226 * 1024;
        if (epi == 0 || v1_fits || n32_off) return false;
    }
    const int n_slices = g.N / BN;
    if (n_slices > sms) return false;
    const int grid = (sms / n_slices) * n_slices;
    if ((long)m_tiles * n_slices < 4L * grid) return false;
    WSParams wp{};
    KParams& p = wp.k;
    p.M = g.M; p.N = g.N; p.nseg = g.nseg; p.batch = 1;
    for (int s = 0; s < g.nseg; ++s) {
        p.seg[s].row = (short)g.seg[s].row_off;
        p.seg[s].col = 0;
        p.seg[s].nk = (short)nkc;
        p.seg[s].dw = 0;
    }
    p.bias = g.bias; p.res1 = g.res1; p.ldres1 = g.ldres1; p.res2 = g.res2; p.ldres2 = g.ldres2;
    p.alpha = g.alpha; p.act1 = g.act1; p.act1_p = g.act1_p; p.act2 = g.act2; p.act2_p = g.act2_p;
    p.out32 = g.out32; p.ld32 = g.ld32; p.out16 = g.out16; p.ld16 = g.ld16; p.vec_ok = 1;
    wp.rmin = rmin; wp.HRp = HRp; wp.nkc = nkc; wp.n_slices = n_slices; wp.m_tiles = m_tiles; wp.bo_mode = bo_mode;
    wp.has_r2 = has_r2 ? 1 : 0; wp.do16 = do16 ? 1 : 0;

    CUtensorMap tm[6];
    {
        cuuint64_t dims[3] = {(cuuint64_t)g.a_cols, (cuuint64_t)g.a_rows, 1};
        cuuint64_t str[2] = {(cuuint64_t)g.lda * 2, (cuuint64_t)g.lda * 2 * (cuuint64_t)g.a_rows};
        cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)HRp, 1};
        encode_map(&tm[0], g.A, 3, dims, str, box, BK);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)g.b_cols, (cuuint64_t)g.b_rows};
        cuuint64_t str[1] = {(cuuint64_t)g.ldb * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
        encode_map(&tm[1], g.B, 2, dims, str, box, BK);
    }
    auto map32 = [&](CUtensorMap* m, const float* base, long ld) {
        cuuint64_t dims[2] = {(cuuint64_t)g.N, (cuuint64_t)g.M};
        cuuint64_t str[1] = {(cuuint64_t)ld * 4};
        cuuint32_t box[2] = {32u, (cuuint32_t)BM};
        encode_map_ex(m, base, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B);
    };
    tm[2] = tm[0]; tm[3] = tm[0]; tm[4] = tm[0]; tm[5] = tm[0];        // placeholders for operands a variant does not touch
    if (g.res1) map32(&tm[2], g.res1, g.ldres1);
    if (g.res2) map32(&tm[3], g.res2, g.ldres2);
    if (g.out32) map32(&tm[4], g.out32, g.ld32);
    if (g.out16) {
        cuuint64_t dims[2] = {(cuuint64_t)g.N, (cuuint64_t)g.M};
        cuuint64_t str[1] = {(cuuint64_t)g.ld16 * 2};
        cuuint32_t box[2] = {(cuuint32_t)BN, (cuuint32_t)BM};
        encode_map_ex(&tm[5], g.out16, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, 2, dims, str, box,
                      BN == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
    }
    const size_t smem = ws2_smem_bytes(BN, BK, g.nseg, nkc, HRp, epi, has_r2, do16);
    if (gemm_prof_on()) gemm_prof_record_begin(stream);
#define RVCB_WS2_LAUNCH(bn, bk)                                                              \
    if (BN == bn && BK == bk) {                                                              \
        if (epi == 0) ws2_launch_e<bn, bk, 0>(tm, wp, grid, smem, stream);                   \
        else if (epi == 1) ws2_launch_e<bn, bk, 1>(tm, wp, grid, smem, stream);              \
        else ws2_launch_e<bn, bk, 2>(tm, wp, grid, smem, stream);                            \
    }
    RVCB_WS2_LAUNCH(64, 64) RVCB_WS2_LAUNCH(32, 64) RVCB_WS2_LAUNCH(32, 32)
#undef RVCB_WS2_LAUNCH
    KERNEL_CHECK();
    if (gemm_prof_on()) {
        const double mn = (double)g.M * g.N;
        const double bytes = (double)g.M * nkc * BK * 2 + (double)g.N * g.nseg * nkc * BK * 2 + (g.res1 ? mn * 4 : 0) + (g.res2 ? mn * 4 : 0) +
                             (g.out32 ? mn * 4 : 0) + (g.out16 ? mn * 2 : 0);
        ProfInfo info{g.M, g.N, g.nseg * nkc, BK, -BN /*negative = WS kernel*/, 1, g.nseg, m_tiles * n_slices};
        info.bytes = bytes;
        gemm_prof_record_end(stream, info);
    }
    count_launch();
    return true;
}

bool gemm_ws_try(const GemmArgs& g, cudaStream_t stream) {
    static int mode = -1, bo_mode = 0;
    if (mode < 0) {
        const char* e = getenv("RVCB_WS");
        mode = (e && e[0] == '0') ? 0 : 1;
        const char* b = getenv("RVCB_WS_BASEOFF");
        bo_mode = (b && b[0] == '1') ? 1 : 0;
    }
    if (!mode) return false;
    // qualifying launches: plain 1-D stride-1 convolution taps over one activation matrix
    if (g.conv2d_W != 0 || g.batch != 1 || g.nseg < 2 || g.gate || g.up2_C || g.bias_per_row) return false;
    if (g.block_k != 64 && g.block_k != 32) return false;
    const int BK = g.block_k, nkc = g.seg[0].nk;
    int rmin = g.seg[0].row_off, rmax = rmin;
    for (int s = 0; s < g.nseg; ++s) {
        if (g.seg[s].col_off != 0 || g.seg[s].nk != nkc) return false;
        rmin = std::min(rmin, g.seg[s].row_off);
        rmax = std::max(rmax, g.seg[s].row_off);
    }
    const int span = rmax - rmin;
    if (span > 120) return false;
    const int HRp = round_up(BM + span, 8);
    const int m_tiles = ceil_div(g.M, BM);
    int sms = 148;
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        sms = sm_budget(sms);
    }
    if (ws2_try(g, stream, BK, nkc, rmin, HRp, m_tiles, sms, bo_mode)) return true;
    int BN = 0;
    for (int cand : {64, 32, 16}) {
        if (BK == 32 && cand > 32) continue;
        if (cand > round_up(g.N, 16)) continue;
        if (ws_smem_bytes(cand, BK, g.nseg, nkc, HRp) <= 226 * 1024) { BN = cand; break; }
    }
    if (BN == 0) return false;
    // measured (profiles/r1d): slicing N narrower than 64 columns only pays when the whole C_in fits one chunk
    // (one halo tile per M-tile); otherwise the streaming kernel's 128-wide tiles are faster.
    if (BN < 64 && BN < round_up(g.N, 16) && nkc > 1) return false;
    if (BN < 32 && g.N > 16) return false;
    const int n_slices = ceil_div(g.N, BN);
    if (n_slices > sms) return false;
    const int grid = (sms / n_slices) * n_slices;
    // the resident weights must be amortised over enough M tiles per CTA
    if ((long)m_tiles * n_slices < 4L * grid) return false;
    auto al = [](const void* ptr, int a) { return ptr == nullptr || (reinterpret_cast<uintptr_t>(ptr) % a) == 0; };
    WSParams wp{};
    KParams& p = wp.k;
    p.M = g.M; p.N = g.N; p.nseg = g.nseg; p.batch = 1;
    for (int s = 0; s < g.nseg; ++s) {
        p.seg[s].row = (short)g.seg[s].row_off;
        p.seg[s].col = 0;
        p.seg[s].nk = (short)nkc;
        p.seg[s].dw = 0;
    }
    p.bias = g.bias; p.res1 = g.res1; p.ldres1 = g.ldres1; p.res2 = g.res2; p.ldres2 = g.ldres2;
    p.alpha = g.alpha; p.act1 = g.act1; p.act1_p = g.act1_p; p.act2 = g.act2; p.act2_p = g.act2_p;
    p.out32 = g.out32; p.ld32 = g.ld32; p.out16 = g.out16; p.ld16 = g.ld16;
    bool v = al(g.out32, 16) && al(g.out16, 8) && al(g.res1, 16) && al(g.res2, 16);
    if (g.out32) v = v && (g.ld32 % 4 == 0);
    if (g.out16) v = v && (g.ld16 % 4 == 0);
    if (g.res1) v = v && (g.ldres1 % 4 == 0);
    if (g.res2) v = v && (g.ldres2 % 4 == 0);
    p.vec_ok = v ? 1 : 0;
    if (g.bias) RVCB_CHECK(al(g.bias, 16), "gemm_ws: bias must be 16B aligned");
    wp.rmin = rmin; wp.HRp = HRp; wp.nkc = nkc; wp.n_slices = n_slices; wp.m_tiles = m_tiles; wp.bo_mode = bo_mode;

    CUtensorMap ta, tb;
    {
        cuuint64_t dims[3] = {(cuuint64_t)g.a_cols, (cuuint64_t)g.a_rows, 1};
        cuuint64_t str[2] = {(cuuint64_t)g.lda * 2, (cuuint64_t)g.lda * 2 * (cuuint64_t)g.a_rows};
        cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)HRp, 1};
        encode_map(&ta, g.A, 3, dims, str, box, BK);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)g.b_cols, (cuuint64_t)g.b_rows};
        cuuint64_t str[1] = {(cuuint64_t)g.ldb * 2};
        cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BN};
        encode_map(&tb, g.B, 2, dims, str, box, BK);
    }
    const size_t smem = ws_smem_bytes(BN, BK, g.nseg, nkc, HRp);
    int epi = 3;
    const bool common = p.vec_ok && g.bias && g.act1 == ACT_NONE && (g.N % (BN >= 32 ? 32 : 16)) == 0;
    if (common && !g.res1 && !g.res2 && !g.out32 && g.out16 && g.act2 == ACT_LRELU && g.alpha == 1.f) epi = 0;
    else if (common && g.res1 && !g.res2 && g.out32 && g.out16 && g.act2 == ACT_LRELU && g.alpha == 1.f) epi = 1;
    else if (common && g.res1 && g.out32 && (!g.out16 || g.act2 == ACT_LRELU)) epi = 2;
#define RVCB_WS_LAUNCH(bn, bk)                                   \
    if (BN == bn && BK == bk) {                                  \
        ws_launch<bn, bk>(ta, tb, wp, grid, smem, stream, epi);  \
        return true;                                             \
    }
    RVCB_WS_LAUNCH(64, 64) RVCB_WS_LAUNCH(32, 64) RVCB_WS_LAUNCH(16, 64)
    RVCB_WS_LAUNCH(32, 32) RVCB_WS_LAUNCH(16, 32)
#undef RVCB_WS_LAUNCH
    return false;
}

}  // namespace rvcb
