// 3x3 convolution of RMVPE's full-resolution U-Net levels (rvc/f0/deepunet.py:7-45, ConvBlockRes at 16 channels, W = 128 mel bins):
//     y[h, w, :] = relu(bias + sum_{dh,dw} W[dh,dw] x[h+dh-1, w+dw-1, :])  (+ fp32 residual)        x, y: [H, 128, 16] channels last
// 14 of these run per utterance at H = 1632.  On the generic implicit-GEMM kernel each 128-pixel tile fetches nine shifted 4 KB
// copies of the input through 32-byte TMA rows and pays a per-tile latency chain (38 us per launch for 2 us of MMA work).
// Here one image row is one M = 128 MMA tile and its input arrives ONCE:
//   * one TMA box per output row: the 3 x 130 x 16 halo (rows h-1..h+1, columns -1..128; zero fill outside the image) lands as
//     a pixel-major K-major tile of 32-byte rows (SW32);
//   * a tap (dh, dw) is that tile at the row-shifted UMMA descriptor (dh * 130 + dw) -- 9 MMAs of K = 16, N = 16 per row against
//     the 4.6 KB of weights resident in shared memory;
//   * thread = pixel epilogue: TMEM -> +bias -> ReLU -> (+ residual) -> fp32 and / or fp16 rows (64 / 32 contiguous bytes per thread).
// Persistent grid, 2 CTAs per SM (80 KB each), 6-stage halo ring, double-buffered 16-column accumulators.
// Warps: 0 TMA producer, 1 MMA issuer (+ TMEM owner), 2..5 epilogue.  Bound: HBM (6.7 MB in, up to 20 MB out per launch).
#include "conv2d_row.cuh"
#include "tc_common.cuh"

namespace rvcb {

namespace {

constexpr int CR_THREADS = 192;
constexpr int CR_W = 128, CR_C = 16;
constexpr int CR_HALO_TX = 3 * (CR_W + 2) * CR_C * 2;               // 12480 bytes delivered per row
constexpr int CR_HALO = (CR_HALO_TX + 255) / 256 * 256;              // 12544: stage pitch, SW32 atoms are 256 bytes
constexpr int CR_NS = 6;
constexpr int CR_WBYTES = 9 * 16 * 32;                               // nine [16 x 16] fp16 tap blocks
constexpr int CR_SMEM = CR_WBYTES + 512 + CR_NS * CR_HALO + 256 + 1024;

struct CRParams {
    int H;
    const float* bias;
    const float* res2; long ldres2;
    float* out32; long ld32;
    __half* out16; long ld16;
    int relu;
    int n_valid;                 // output channels actually present (bias entries, stored columns); 16 except for the 16 -> 3 head conv
};

__global__ void __launch_bounds__(CR_THREADS, 2)
conv2d_row_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CRParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;                                   // [9][16 rows x 32 B]
    uint8_t* sH = smem + CR_WBYTES + 512;                 // [NS] halo tiles (5120 = 20 * 256: keeps the 256-byte atom alignment)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sH + CR_NS * CR_HALO);
    uint64_t* w_full = bars;             // [1]
    uint64_t* h_full = bars + 1;         // [NS]
    uint64_t* h_empty = h_full + CR_NS;  // [NS]
    uint64_t* t_full = h_empty + CR_NS;  // [2]
    uint64_t* t_empty = t_full + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
        mbar_init(w_full, 1);
        for (int i = 0; i < CR_NS; ++i) { mbar_init(&h_full[i], 1); mbar_init(&h_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<32>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ======================= TMA producer =======================
        if (elect_one()) {
            mbar_expect_tx(w_full, CR_WBYTES);
#pragma unroll
            for (int t = 0; t < 9; ++t) tma_load_2d(sW + t * 512, &tmap_w, w_full, t * 16, 0);
        }
        __syncwarp();
        uint32_t s = 0, ph = 0;
        for (int r = blockIdx.x; r < p.H; r += gridDim.x) {
            mbar_wait(&h_empty[s], ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&h_full[s], CR_HALO_TX);
                tma_load_3d(sH + s * CR_HALO, &tmap_x, &h_full[s], 0, -1, r - 1);      // columns -1..128, rows r-1..r+1, zero fill outside
            }
            __syncwarp();
            if (++s == CR_NS) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ======================= MMA issuer =======================
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t desc_hi = (uint32_t)(((uint64_t)(256 >> 4) << 32 | (1ull << 46) | (6ull << 61)) >> 32);      // SW32, SBO = 8 rows x 32 B
        const uint32_t h_lo0 = ((smem_u32(sH) & 0x3FFFF) >> 4) | (1u << 16);
        const uint32_t w_lo0 = ((smem_u32(sW) & 0x3FFFF) >> 4) | (1u << 16);
        mbar_wait(w_full, 0);
        uint32_t s = 0, ph = 0;
        int it = 0;
        for (int r = blockIdx.x; r < p.H; r += gridDim.x, ++it) {
            const int a = it & 1;
            mbar_wait(&t_empty[a], ((it >> 1) & 1) ^ 1);
            mbar_wait(&h_full[s], ph);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const uint32_t a_lo = h_lo0 + (uint32_t)((s * CR_HALO + ((t / 3) * (CR_W + 2) + (t % 3)) * 32) >> 4);
                    const uint32_t b_lo = w_lo0 + (uint32_t)((t * 512) >> 4);
                    umma_f16(tmem_base + a * 16, ((uint64_t)desc_hi << 32) | (uint64_t)a_lo, ((uint64_t)desc_hi << 32) | (uint64_t)b_lo, idesc, (uint32_t)(t != 0));
                }
                umma_commit(&h_empty[s]);
                umma_commit(&t_full[a]);
            }
            __syncwarp();
            if (++s == CR_NS) { s = 0; ph ^= 1; }
        }
    } else {
        // ======================= epilogue: thread = pixel =======================
        const int q = warp & 3;
        const int w = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        float bias[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) bias[i] = i < p.n_valid ? __ldg(p.bias + i) : 0.f;
        int it = 0;
        for (int r = blockIdx.x; r < p.H; r += gridDim.x, ++it) {
            const int a = it & 1;
            const long px = (long)r * CR_W + w;
            float4 res[4];
            if (p.res2) {                                   // residual rows do not depend on the accumulator: fetch them first
#pragma unroll
                for (int i = 0; i < 4; ++i) res[i] = *reinterpret_cast<const float4*>(p.res2 + px * p.ldres2 + 4 * i);
            }
            mbar_wait(&t_full[a], (it >> 1) & 1);
            tc_fence_after();
            uint32_t raw[16];
            tmem_ld16(tmem_base + lane_addr + (uint32_t)(a * 16), raw);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[a]);
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                v[i] = __uint_as_float(raw[i]) + bias[i];
                if (p.relu) v[i] = fmaxf(v[i], 0.f);
            }
            if (p.res2) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { v[4 * i] += res[i].x; v[4 * i + 1] += res[i].y; v[4 * i + 2] += res[i].z; v[4 * i + 3] += res[i].w; }
            }
            if (p.out32) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4*>(p.out32 + px * p.ld32 + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
            if (p.out16 && p.n_valid < 16) {               // narrow head (16 -> 3): a few scalar stores per pixel
                for (int i = 0; i < p.n_valid; ++i) p.out16[px * p.ld16 + i] = __float2half_rn(v[i]);
            } else if (p.out16) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const __half2 h0 = __floats2half2_rn(v[8 * i], v[8 * i + 1]), h1 = __floats2half2_rn(v[8 * i + 2], v[8 * i + 3]);
                    const __half2 h2 = __floats2half2_rn(v[8 * i + 4], v[8 * i + 5]), h3 = __floats2half2_rn(v[8 * i + 6], v[8 * i + 7]);
                    *reinterpret_cast<uint4*>(p.out16 + px * p.ld16 + 8 * i) =
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
        tmem_dealloc<32>(tmem_base);
    }
}


// ------------------------------------------------------------------------------------------------
// The same idea for the narrower levels (W = 64 / 32, C = 32 / 64) and for C_in != C_out: an M = 128 tile is BH = 128 / W whole
// image rows.  Horizontal zero padding cannot come from a padded pitch any more (the 128 output pixels must be 128 CONSECUTIVE
// operand rows), so the halo arrives as THREE column-shifted copies, one per dw, each a TMA box {C_in, W, BH + 2} starting at
// column dw - 1: the tensor map zero-fills exactly the pixels that fall outside the image for that shift.  Inside a copy the
// three vertical taps are row shifts of dh * W.  Activations are read 3 times instead of 9, weights stay resident.
// ------------------------------------------------------------------------------------------------
template <int CIN, int COUT, int W>
struct CTCfg {
    static constexpr int BH = 128 / W;
    static constexpr int PITCH = CIN * 2;                                   // 32 / 64 / 128 bytes: SW32 / SW64 / SW128
    static constexpr int COPY = (BH + 2) * W * PITCH;                       // one column-shifted halo copy
    static constexpr int STAGE = 3 * COPY;
    static constexpr int WTAP = COUT * PITCH;                               // one tap of the weights: [C_out rows, C_in cols]
    static constexpr int WBYTES = (9 * WTAP + 1023) / 1024 * 1024;
    static constexpr int NS = (WBYTES + 2 * STAGE + 2048 <= 232448) ? 2 : 1;
    static constexpr int TMEM_COLS = 2 * COUT <= 32 ? 32 : (2 * COUT <= 64 ? 64 : 128);
    static constexpr int SMEM = WBYTES + NS * STAGE + 256 + 1024;
    static_assert(COPY % 1024 == 0 && WTAP % 512 == 0 && 128 % W == 0 && CIN % 16 == 0 && COUT % 16 == 0 && CIN <= 64 && COUT <= 64, "tile geometry");
};

template <int CIN, int COUT, int W>
__global__ void __launch_bounds__(CR_THREADS, 1)
conv2d_tile_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CRParams p) {
    using G = CTCfg<CIN, COUT, W>;
    constexpr int BH = G::BH, PITCH = G::PITCH, NS = G::NS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sW = smem;
    uint8_t* sH = smem + G::WBYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sH + NS * G::STAGE);
    uint64_t* w_full = bars;
    uint64_t* h_full = bars + 1;         // [NS]
    uint64_t* h_empty = h_full + NS;     // [NS]
    uint64_t* t_full = h_empty + NS;     // [2]
    uint64_t* t_empty = t_full + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int ntiles = (p.H + BH - 1) / BH;
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_x);
        prefetch_tmap(&tmap_w);
        mbar_init(w_full, 1);
        for (int i = 0; i < NS; ++i) { mbar_init(&h_full[i], 1); mbar_init(&h_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 4); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<G::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(w_full, 9 * G::WTAP);
#pragma unroll
            for (int t = 0; t < 9; ++t) tma_load_2d(sW + t * G::WTAP, &tmap_w, w_full, t * CIN, 0);
        }
        __syncwarp();
        uint32_t s = 0, ph = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            mbar_wait(&h_empty[s], ph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(&h_full[s], G::STAGE);
#pragma unroll
                for (int dw = 0; dw < 3; ++dw) tma_load_3d(sH + s * G::STAGE + dw * G::COPY, &tmap_x, &h_full[s], 0, dw - 1, tile * BH - 1);
            }
            __syncwarp();
            if (++s == (uint32_t)NS) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(COUT >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        constexpr uint32_t desc_hi = (uint32_t)(((uint64_t)((8 * PITCH) >> 4) << 32 | (1ull << 46) |
                                                 ((PITCH == 128 ? 2ull : PITCH == 64 ? 4ull : 6ull) << 61)) >> 32);
        const uint32_t h_lo0 = ((smem_u32(sH) & 0x3FFFF) >> 4) | (1u << 16);
        const uint32_t w_lo0 = ((smem_u32(sW) & 0x3FFFF) >> 4) | (1u << 16);
        mbar_wait(w_full, 0);
        uint32_t s = 0, ph = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int a = it & 1;
            mbar_wait(&t_empty[a], ((it >> 1) & 1) ^ 1);
            mbar_wait(&h_full[s], ph);
            tc_fence_after();
            if (elect_one()) {
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const int dh = t / 3, dw = t % 3;
                    const uint32_t a_lo = h_lo0 + (uint32_t)((s * G::STAGE + dw * G::COPY + dh * W * PITCH) >> 4);
                    const uint32_t b_lo = w_lo0 + (uint32_t)((t * G::WTAP) >> 4);
#pragma unroll
                    for (int ks = 0; ks < CIN / 16; ++ks)
                        umma_f16(tmem_base + a * COUT, ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + 2 * ks), ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + 2 * ks),
                                 idesc, (uint32_t)((t | ks) != 0));
                }
                umma_commit(&h_empty[s]);
                umma_commit(&t_full[a]);
            }
            __syncwarp();
            if (++s == (uint32_t)NS) { s = 0; ph ^= 1; }
        }
    } else {
        const int q = warp & 3;
        const int i = q * 32 + lane;
        const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
        const long npix = (long)p.H * W;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int a = it & 1;
            const long px = (long)tile * 128 + i;
            const bool ok = px < npix;
            mbar_wait(&t_full[a], (it >> 1) & 1);
            tc_fence_after();
#pragma unroll 1
            for (int cc = 0; cc < COUT / 16; ++cc) {
                float4 res[4];
                if (p.res2 && ok) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) res[j] = *reinterpret_cast<const float4*>(p.res2 + px * p.ldres2 + cc * 16 + 4 * j);
                }
                uint32_t raw[16];
                tmem_ld16(tmem_base + lane_addr + (uint32_t)(a * COUT + cc * 16), raw);
                tmem_ld_wait();
                float v[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    v[j] = __uint_as_float(raw[j]) + __ldg(p.bias + cc * 16 + j);
                    if (p.relu) v[j] = fmaxf(v[j], 0.f);
                }
                if (!ok) continue;
                if (p.res2) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { v[4 * j] += res[j].x; v[4 * j + 1] += res[j].y; v[4 * j + 2] += res[j].z; v[4 * j + 3] += res[j].w; }
                }
                if (p.out32) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<float4*>(p.out32 + px * p.ld32 + cc * 16 + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                if (p.out16) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const __half2 h0 = __floats2half2_rn(v[8 * j], v[8 * j + 1]), h1 = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
                        const __half2 h2 = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]), h3 = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
                        *reinterpret_cast<uint4*>(p.out16 + px * p.ld16 + cc * 16 + 8 * j) =
                            make_uint4(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1),
                                       *reinterpret_cast<const uint32_t*>(&h2), *reinterpret_cast<const uint32_t*>(&h3));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_empty[a]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<G::TMEM_COLS>(tmem_base);
    }
}

template <int CIN, int COUT, int W>
void conv2d_tile_launch(const Conv2dRowArgs& a, const CRParams& p, cudaStream_t stream) {
    using G = CTCfg<CIN, COUT, W>;
    CUtensorMap tx, tw;
    {
        cuuint64_t dims[3] = {(cuuint64_t)CIN, (cuuint64_t)W, (cuuint64_t)a.H};
        cuuint64_t str[2] = {(cuuint64_t)a.ldx * 2, (cuuint64_t)a.ldx * 2 * W};
        cuuint32_t box[3] = {(cuuint32_t)CIN, (cuuint32_t)W, (cuuint32_t)(G::BH + 2)};
        encode_map(&tx, a.x, 3, dims, str, box, CIN);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)a.w_cols, (cuuint64_t)a.w_rows};
        cuuint64_t str[1] = {(cuuint64_t)a.w_cols * 2};
        cuuint32_t box[2] = {(cuuint32_t)CIN, (cuuint32_t)COUT};
        encode_map(&tw, a.w, 2, dims, str, box, CIN);
    }
    static bool configured = false;
    static int sms = 148;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(conv2d_tile_kernel<CIN, COUT, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    const int ntiles = (a.H + G::BH - 1) / G::BH;
    const int grid = std::min(ntiles, sm_budget(sms));
    if (gemm_prof_on()) gemm_prof_record_begin(stream);
    conv2d_tile_kernel<CIN, COUT, W><<<grid, CR_THREADS, G::SMEM, stream>>>(tx, tw, p);
    KERNEL_CHECK();
    if (gemm_prof_on()) gemm_prof_record_end(stream, ProfInfo{a.H * W, COUT, 9 * (CIN / 16), 16, COUT, 1, 9, ntiles});
    count_launch();
}

}  // namespace

bool conv2d_row_try(const Conv2dRowArgs& a, cudaStream_t stream) {
    static const bool on = [] { const char* e = getenv("RVCB_CONV_ROW"); return !(e && e[0] == '0'); }();
    static const bool tiles_on = [] { const char* e = getenv("RVCB_CONV_TILE"); return !(e && e[0] == '0'); }();
    auto al16 = [](const void* ptr) { return ptr == nullptr || (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; };
    if (!on || a.H < 1 || a.w_cols != 9 * a.cin || a.w_rows < a.cout) return false;
    if (a.ldx % 8 || !al16(a.x) || !al16(a.w) || !a.bias) return false;
    if (a.res2 && (a.ldres2 % 4 || !al16(a.res2))) return false;
    if (a.out32 && (a.ld32 % 4 || !al16(a.out32))) return false;
    const bool narrow = a.W == CR_W && a.cin == CR_C && a.cout < CR_C && a.cout >= 1 && a.out16 && !a.out32 && !a.res2;
    if (a.out16 && !narrow && (a.ld16 % 8 || !al16(a.out16))) return false;
    if (!a.out32 && !a.out16) return false;
    CRParams p{};
    p.H = a.H; p.bias = a.bias; p.res2 = a.res2; p.ldres2 = a.ldres2; p.out32 = a.out32; p.ld32 = a.ld32; p.out16 = a.out16; p.ld16 = a.ld16;
    p.relu = a.relu ? 1 : 0;
    p.n_valid = narrow ? a.cout : CR_C;
    if (!(a.W == CR_W && a.cin == CR_C && (a.cout == CR_C || narrow))) {
        if (!tiles_on) return false;
        if (a.cin == 32 && a.cout == 32 && a.W == 64) { conv2d_tile_launch<32, 32, 64>(a, p, stream); return true; }
        if (a.cin == 64 && a.cout == 64 && a.W == 32) { conv2d_tile_launch<64, 64, 32>(a, p, stream); return true; }
        if (a.cin == 32 && a.cout == 16 && a.W == 128) { conv2d_tile_launch<32, 16, 128>(a, p, stream); return true; }
        if (a.cin == 64 && a.cout == 32 && a.W == 64) { conv2d_tile_launch<64, 32, 64>(a, p, stream); return true; }
        return false;
    }
    CUtensorMap tx, tw;
    {
        cuuint64_t dims[3] = {(cuuint64_t)CR_C, (cuuint64_t)CR_W, (cuuint64_t)a.H};
        cuuint64_t str[2] = {(cuuint64_t)a.ldx * 2, (cuuint64_t)a.ldx * 2 * CR_W};
        cuuint32_t box[3] = {(cuuint32_t)CR_C, (cuuint32_t)(CR_W + 2), 3u};
        encode_map(&tx, a.x, 3, dims, str, box, 16);
    }
    {
        cuuint64_t dims[2] = {(cuuint64_t)a.w_cols, (cuuint64_t)a.w_rows};
        cuuint64_t str[1] = {(cuuint64_t)a.w_cols * 2};
        cuuint32_t box[2] = {16u, 16u};
        encode_map(&tw, a.w, 2, dims, str, box, 16);
    }
    static bool configured = false;
    static int sms = 148;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(conv2d_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CR_SMEM));
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        configured = true;
    }
    const int grid = std::min(a.H, 2 * sm_budget(sms));
    if (gemm_prof_on()) gemm_prof_record_begin(stream);
    conv2d_row_kernel<<<grid, CR_THREADS, CR_SMEM, stream>>>(tx, tw, p);
    KERNEL_CHECK();
    if (gemm_prof_on()) gemm_prof_record_end(stream, ProfInfo{a.H * CR_W, CR_C, 9, 16, 16, 1, 9, a.H});
    count_launch();
    return true;
}

}  // namespace rvcb
