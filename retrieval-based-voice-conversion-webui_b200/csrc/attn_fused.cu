// Fused multi-head self-attention for HuBERT (fairseq MultiheadAttention as called at rvc/hubert.py:60-70; 12 heads x 64):
//     ctx[:, h] = softmax(q_h k_h^T) v_h          (q already carries the 64^-0.5 scaling, folded into the packed weights)
// One launch per layer instead of  QK^T GEMM -> fp32 scores in HBM -> softmax kernel -> fp16 probabilities in HBM -> PV GEMM:
// the scores live in TENSOR MEMORY only, the probabilities in shared memory only.
//
// One CTA = one (head, 128-query block).  Two passes over the key blocks (T <= ~1600, d = 64: QK^T is cheap, and an exact row
// maximum first means the running output never has to be rescaled in TMEM):
//   pass 1   S = Q K_b^T (tcgen05.mma, M = 128, N = 128, K = 64) -> row maxima (thread = query row, tcgen05.ld)
//   pass 2   S again -> p = exp(S - max) (keys >= T masked to 0), row sums, p -> fp16 K-major swizzled tile in shared memory
//            -> O += P V_b (M = 128, N = 64, K = 128) accumulated in TMEM; S is double-buffered so QK^T of block b+1 runs
//            under the exponentials of block b
//   final    O / rowsum -> fp16 -> ctx[T, heads*64]
// Operands arrive by TMA (Q once; K and V^T blocks through 2-slot rings; rows / keys beyond T are zero-filled by the tensor map).
// Warps: 0 TMA producer, 1 MMA issuer (+ TMEM owner), 2..5 softmax / epilogue (one TMEM lane quadrant each).
//
// The same kernel, templated on the head geometry, serves the TextEncoder's relative-position attention
// (rvc/layers/attentions.py:86-142; 2 heads x 96, window 10): q.k over 128 padded columns, V of 96 columns,
//     S[i, j] = (q_i . k_j) / sqrt(96) + [|j - i| <= 10] qrel[i, j - i + 10]          (qrel = (q / sqrt(96)) E_k^T, a tiny GEMM upstream)
//     O[i]    = sum_j P[i, j] v_j + sum_{|j - i| <= 10} P[i, j] E_v[j - i + 10]
// the band logits are added where the scores are read, the band probabilities are kept per row in shared memory and the E_v
// term is a 21 x 96 register FMA in the final step -- no fp32 score matrix, no probability matrix, no separate rel-v GEMM.
#include "attn_fused.cuh"
#include "tc_common.cuh"

namespace rvcb {

namespace {

constexpr int AT_THREADS = 192;
constexpr int AT_CH = 128 * 128;             // one 64-column operand chunk of 128 rows (Q, K, P): 16 KB
constexpr int AT_WIN = 10, AT_NREL = 2 * AT_WIN + 1;

template <int DK, int DV, bool REL>
struct AtCfg {
    static constexpr int NKQ = DK / 64;                        // 64-column chunks of the q.k contraction
    static constexpr int Q_BYTES = NKQ * AT_CH;
    static constexpr int K_BYTES = NKQ * AT_CH;
    static constexpr int VC_BYTES = DV * 128;                  // one V^T chunk: [DV dims, 64 keys] fp16
    static constexpr int REL_BYTES = REL ? (2 * 128 * AT_NREL * 4 + AT_NREL * DV * 4 + 1023) / 1024 * 1024 : 0;
    static constexpr int SMEM = Q_BYTES + 2 * K_BYTES + 4 * VC_BYTES + 2 * AT_CH + REL_BYTES + 256 + 1024;
    static_assert(VC_BYTES % 1024 == 0 && DV % 16 == 0 && DV <= 128 && SMEM <= 232448, "attention tile geometry");
};

struct AttnParams {
    int T, heads, nkv;
    float qscale;
    const float* qrel;       // [heads, T, 32] relative-key logits (REL)
    const float* ev;         // [21, DV] relative-value embedding (REL)
    __half* out;
    long ldo;
};

template <int DK, int DV, bool REL>
__global__ void __launch_bounds__(AT_THREADS, 1)
attn_fused_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                  const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ AttnParams p) {
    using G = AtCfg<DK, DV, REL>;
    constexpr int NKQ = G::NKQ;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + G::Q_BYTES;                 // [2]
    uint8_t* sV = sK + 2 * G::K_BYTES;             // [2][2 chunks]
    uint8_t* sP = sV + 4 * G::VC_BYTES;            // [2 chunks]
    float* s_qrel = reinterpret_cast<float*>(sP + 2 * AT_CH);          // [128][21] band logits of this CTA's query rows
    float* s_pb = s_qrel + 128 * AT_NREL;                               // [128][21] band probabilities
    float* s_ev = s_pb + 128 * AT_NREL;                                 // [21][DV]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * AT_CH + G::REL_BYTES);
    uint64_t* q_full = bars;             // [1]
    uint64_t* k_full = bars + 1;         // [2]
    uint64_t* k_empty = bars + 3;        // [2]
    uint64_t* v_full = bars + 5;         // [2]
    uint64_t* v_empty = bars + 7;        // [2]
    uint64_t* s_full = bars + 9;         // [2]
    uint64_t* s_empty = bars + 11;       // [2]
    uint64_t* p_full = bars + 13;        // [1]
    uint64_t* p_empty = bars + 14;       // [1]
    uint64_t* o_full = bars + 15;        // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int h = blockIdx.x % p.heads;
    const int qb = blockIdx.x / p.heads;
    const int nkv = p.nkv;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_q);
        prefetch_tmap(&tmap_k);
        prefetch_tmap(&tmap_v);
        mbar_init(q_full, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
            mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
        }
        mbar_init(p_full, 4);
        mbar_init(p_empty, 1);
        mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    if (REL) {
        for (int i = threadIdx.x; i < 128 * AT_NREL; i += AT_THREADS) {
            const int r = i / AT_NREL, d = i - r * AT_NREL, row = qb * 128 + r;
            s_qrel[i] = row < p.T ? p.qrel[((long)h * p.T + row) * 32 + d] : 0.f;
            s_pb[i] = 0.f;
        }
        for (int i = threadIdx.x; i < AT_NREL * DV; i += AT_THREADS) s_ev[i] = p.ev[i];
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tm_s0 = tmem_base, tm_o = tmem_base + 256;

    if (warp == 0) {
        // ======================= TMA producer =======================
        if (elect_one()) {
            mbar_expect_tx(q_full, G::Q_BYTES);
#pragma unroll
            for (int c = 0; c < NKQ; ++c) tma_load_2d(sQ + c * AT_CH, &tmap_q, q_full, h * DK + c * 64, qb * 128);
        }
        __syncwarp();
        for (int it = 0; it < 2 * nkv; ++it) {            // K blocks: pass 1 then pass 2
            const int kb = it % nkv, s = it & 1;
            mbar_wait(&k_empty[s], ((it >> 1) & 1) ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&k_full[s], G::K_BYTES);
#pragma unroll
                for (int c = 0; c < NKQ; ++c) tma_load_2d(sK + s * G::K_BYTES + c * AT_CH, &tmap_k, &k_full[s], h * DK + c * 64, kb * 128);
            }
            __syncwarp();
            if (it >= nkv) {                              // V^T block of the same keys (pass 2)
                const int vi = it - nkv, vs = vi & 1;
                mbar_wait(&v_empty[vs], ((vi >> 1) & 1) ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(&v_full[vs], 2 * G::VC_BYTES);
                    tma_load_2d(sV + (vs * 2 + 0) * G::VC_BYTES, &tmap_v, &v_full[vs], kb * 128, h * DV);
                    tma_load_2d(sV + (vs * 2 + 1) * G::VC_BYTES, &tmap_v, &v_full[vs], kb * 128 + 64, h * DV);
                }
                __syncwarp();
            }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        constexpr uint32_t idesc_s = (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t idesc_o = (1u << 4) | ((uint32_t)(DV >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t desc_hi = (uint32_t)(((uint64_t)(1024 >> 4) << 32 | (1ull << 46) | (2ull << 61)) >> 32);      // SW128, SBO = 1024
        auto lo = [](const void* ptr) { return ((smem_u32(ptr) & 0x3FFFF) >> 4) | (1u << 16); };
        mbar_wait(q_full, 0);
        tc_fence_after();
        // S[it & 1] = Q K^T for K block `it` of the 2 * nkv block sequence
        auto issue_s = [&](int it) {
            const int s = it & 1;
            const uint32_t ph = (it >> 1) & 1;
            mbar_wait(&s_empty[s], ph ^ 1);
            mbar_wait(&k_full[s], ph);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < NKQ; ++c) {
                    const uint32_t q_lo = lo(sQ + c * AT_CH), k_lo = lo(sK + s * G::K_BYTES + c * AT_CH);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_f16(tm_s0 + s * 128, ((uint64_t)desc_hi << 32) | (uint64_t)(q_lo + 2 * ks), ((uint64_t)desc_hi << 32) | (uint64_t)(k_lo + 2 * ks),
                                 idesc_s, (uint32_t)((c | ks) != 0));
                }
                umma_commit(&k_empty[s]);
                umma_commit(&s_full[s]);
            }
            __syncwarp();
        };
        for (int it = 0; it < nkv; ++it) issue_s(it);                 // pass 1
        issue_s(nkv);                                                // first block of pass 2
        for (int vi = 0; vi < nkv; ++vi) {
            if (vi + 1 < nkv) issue_s(nkv + vi + 1);                 // next scores under this block's exponentials
            const int vs = vi & 1;
            mbar_wait(p_full, vi & 1);
            mbar_wait(&v_full[vs], (vi >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const uint32_t p_lo = lo(sP + c * AT_CH), v_lo = lo(sV + (vs * 2 + c) * G::VC_BYTES);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)
                        umma_f16(tm_o, ((uint64_t)desc_hi << 32) | (uint64_t)(p_lo + 2 * ks), ((uint64_t)desc_hi << 32) | (uint64_t)(v_lo + 2 * ks),
                                 idesc_o, (uint32_t)((vi | c | ks) != 0));
                }
                umma_commit(&v_empty[vs]);
                umma_commit(p_empty);
                if (vi + 1 == nkv) umma_commit(o_full);
            }
            __syncwarp();
        }
    } else {
        // ======================= softmax / epilogue: 4 warps, thread = query row =======================
        const int q = warp & 3;
        const int r = q * 32 + lane;
        const int row = qb * 128 + r;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const float qscale = p.qscale;
        const float* qr = s_qrel + r * AT_NREL;
        float* pb = s_pb + r * AT_NREL;
        // score of key column `col` for this row: scaling + (REL) the band logit
        auto score = [&](uint32_t raw, int col, bool band_block) {
            float s = __uint_as_float(raw) * qscale;
            if (REL && band_block) {
                const int d = col - row + AT_WIN;
                if ((unsigned)d <= (unsigned)(2 * AT_WIN)) s += qr[d];
            }
            return s;
        };
        float m = -INFINITY;
        // ---- pass 1: row maxima (64 columns per TMEM round trip, four independent partial maxima) ----
        for (int it = 0; it < nkv; ++it) {
            const int s = it & 1;
            mbar_wait(&s_full[s], (it >> 1) & 1);
            tc_fence_after();
            const int col0 = it * 128;
            const bool band_block = REL && (it + 1 >= qb) && (it <= qb + 1);
            const bool full_block = col0 + 128 <= p.T;
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t a[4][16];
#pragma unroll
                for (int j = 0; j < 4; ++j) tmem_ld16(tm_s0 + lane_addr + (uint32_t)(s * 128 + hf * 64 + j * 16), a[j]);
                tmem_ld_wait();
                const int c = col0 + hf * 64;
                float pm[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float v = score(a[j][i], c + j * 16 + i, band_block);
                        pm[j] = (full_block || c + j * 16 + i < p.T) ? fmaxf(pm[j], v) : pm[j];
                    }
                m = fmaxf(m, fmaxf(fmaxf(pm[0], pm[1]), fmaxf(pm[2], pm[3])));
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);
        }
        // ---- pass 2: p = exp(S - m), row sums, P tile (one 64-key chunk per TMEM round trip) ----
        float l = 0.f;
        uint8_t* prow = sP + r * 128;
        const int swz = r & 7;
        for (int vi = 0; vi < nkv; ++vi) {
            const int it = nkv + vi, s = it & 1;
            mbar_wait(&s_full[s], (it >> 1) & 1);
            mbar_wait(p_empty, (vi & 1) ^ 1);               // the previous block's P V has consumed the tile
            tc_fence_after();
            const int col0 = vi * 128;
            const bool band_block = REL && (vi + 1 >= qb) && (vi <= qb + 1);
            const bool full_block = col0 + 128 <= p.T;
#pragma unroll 1
            for (int hf = 0; hf < 2; ++hf) {
                uint32_t a[4][16];
#pragma unroll
                for (int j = 0; j < 4; ++j) tmem_ld16(tm_s0 + lane_addr + (uint32_t)(s * 128 + hf * 64 + j * 16), a[j]);
                tmem_ld_wait();
                const int c = col0 + hf * 64;
                float pl[4] = {0.f, 0.f, 0.f, 0.f};
                uint8_t* dst = prow + hf * AT_CH;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float f[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const float e = __expf(score(a[j][i], c + j * 16 + i, band_block) - m);
                        f[i] = (full_block || c + j * 16 + i < p.T) ? e : 0.f;
                        pl[j] += f[i];
                    }
                    if (REL && band_block) {                // keep the band probabilities of this row (the fp16 value the P V MMA sees)
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int d = c + j * 16 + i - row + AT_WIN;
                            if ((unsigned)d <= (unsigned)(2 * AT_WIN)) pb[d] = __half2float(__float2half_rn(f[i]));
                        }
                    }
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int ch = j * 2 + hh;
                        const __half2 h0 = __floats2half2_rn(f[8 * hh], f[8 * hh + 1]), h1 = __floats2half2_rn(f[8 * hh + 2], f[8 * hh + 3]);
                        const __half2 h2 = __floats2half2_rn(f[8 * hh + 4], f[8 * hh + 5]), h3 = __floats2half2_rn(f[8 * hh + 6], f[8 * hh + 7]);
                        *reinterpret_cast<uint4*>(dst + ((ch ^ swz) << 4)) =
                            make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                       *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
                    }
                }
                l += (pl[0] + pl[1]) + (pl[2] + pl[3]);
            }
            fence_proxy_async();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&s_empty[s]);
                mbar_arrive(p_full);
            }
        }
        // ---- final: (O + band probabilities x E_v) / l -> fp16 ----
        mbar_wait(o_full, 0);
        tc_fence_after();
        const float inv = 1.f / l;
        __half* orow = p.out + (long)row * p.ldo + h * DV;
#pragma unroll 1
        for (int cc = 0; cc < DV / 16; ++cc) {
            uint32_t a[16];
            tmem_ld16(tm_o + lane_addr + (uint32_t)(cc * 16), a);
            tmem_ld_wait();
            float o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(a[i]);
            if (REL) {
                for (int d = 0; d < AT_NREL; ++d) {
                    const float w = pb[d];
                    const float* e = s_ev + d * DV + cc * 16;
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = fmaf(w, e[i], o[i]);
                }
            }
            if (row < p.T) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const __half2 h0 = __floats2half2_rn(o[8 * j] * inv, o[8 * j + 1] * inv), h1 = __floats2half2_rn(o[8 * j + 2] * inv, o[8 * j + 3] * inv);
                    const __half2 h2 = __floats2half2_rn(o[8 * j + 4] * inv, o[8 * j + 5] * inv), h3 = __floats2half2_rn(o[8 * j + 6] * inv, o[8 * j + 7] * inv);
                    *reinterpret_cast<uint4*>(orow + cc * 16 + 8 * j) =
                        make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                   *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int DK, int DV, bool REL>
void attn_launch(const AttnFusedArgs& a, cudaStream_t stream) {
    using G = AtCfg<DK, DV, REL>;
    CUtensorMap tq, tk, tv;
    {
        cuuint64_t dims[2] = {(cuuint64_t)a.heads * DK, (cuuint64_t)a.T};
        cuuint64_t str[1] = {(cuuint64_t)a.ldq * 2};
        cuuint32_t box[2] = {64u, 128u};
        encode_map(&tq, a.q, 2, dims, str, box, 64);
        str[0] = (cuuint64_t)a.ldk * 2;
        encode_map(&tk, a.k, 2, dims, str, box, 64);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)a.T, (cuuint64_t)a.heads * DV};      // only the T valid key columns: the rest reads as zero
        cuuint64_t str[1] = {(cuuint64_t)a.ldv * 2};
        cuuint32_t box[2] = {64u, (cuuint32_t)DV};
        encode_map(&tv, a.vT, 2, dims, str, box, 64);
    }
    AttnParams p{};
    p.T = a.T; p.heads = a.heads; p.nkv = ceil_div(a.T, 128);
    p.qscale = a.qscale; p.qrel = a.qrel; p.ev = a.ev;
    p.out = a.out; p.ldo = a.ldo;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(attn_fused_kernel<DK, DV, REL>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
        configured = true;
    }
    const int grid = a.heads * ceil_div(a.T, 128);
    if (gemm_prof_on()) gemm_prof_record_begin(stream);
    attn_fused_kernel<DK, DV, REL><<<grid, AT_THREADS, G::SMEM, stream>>>(tq, tk, tv, p);
    KERNEL_CHECK();
    if (gemm_prof_on()) {
        // two QK^T passes + one PV per head, reported as one launch of the streaming class (flops = 2 * M * N * kb * BK * batch)
        ProfInfo info{a.T, a.T, 1, 2 * DK + DV, 128, a.heads, 1, grid};
        gemm_prof_record_end(stream, info);
    }
    count_launch();
}

}  // namespace

bool attention_fused_supported(const AttnFusedArgs& a) {
    auto al16 = [](const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    const bool geo = (a.dk == 64 && a.dv == 64 && !a.qrel) || (a.dk == 128 && a.dv == 96 && a.qrel && a.ev);
    return geo && a.T >= 1 && a.heads >= 1 && al16(a.q) && al16(a.k) && al16(a.vT) && al16(a.out) && a.ldq % 8 == 0 && a.ldk % 8 == 0 &&
           a.ldv % 8 == 0 && a.ldo % 8 == 0;
}

void attention_fused(const AttnFusedArgs& a, cudaStream_t stream) {
    RVCB_CHECK(attention_fused_supported(a), "attention_fused: unsupported arguments (head geometry 64/64 or 128/96 + relative positions)");
    if (a.qrel) attn_launch<128, 96, true>(a, stream);
    else attn_launch<64, 64, false>(a, stream);
}

}  // namespace rvcb
