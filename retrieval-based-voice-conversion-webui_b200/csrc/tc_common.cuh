// Shared pieces of the tcgen05/TMA kernels: parameter block, PTX wrappers, tensor-map encoding.
#pragma once
#include "gemm.cuh"

#include <cudaTypedefs.h>
#include <cstdlib>
#include <mutex>
#include <utility>
#include <vector>

namespace rvcb {

// ------------------------------------------------------------------------------------------------
// kernel parameters
// ------------------------------------------------------------------------------------------------
struct SegPacked {
    short row, col, nk;
    signed char dw, pad;
};

struct KParams {
    int M, N, nseg, batch, num_m_tiles, num_n_tiles, num_tiles, total_kb;
    int conv2d_W, BH;
    int a_row_z, a_col_z, b_row_z, b_col_z, b_col0;
    long c_z, bias_z;
    const float* bias;
    int bias_per_row;
    const float* res1; long ldres1;
    const float* res2; long ldres2;
    float alpha;
    int act1; float act1_p;
    int act2; float act2_p;
    int gate;
    float* out32; long ld32;
    __half* out16; long ld16;
    int up2_C;
    int vec_ok;
    SegPacked seg[GEMM_MAX_SEG];
};

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One lane of a CONVERGED warp.  Role loops run warp-uniform and wrap only the issuing instruction in elect_one(): the
// descriptors then live in uniform registers (an `if (lane == 0)` loop makes the compiler emit an ELECT / R2UR.BROADCAST
// loop in front of every UTCHMMA / UTMALDG, ~27 instructions per MMA -- measured 95 cycles per 32-cycle MMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// non-blocking phase test, warp-uniform: lane 0's answer is broadcast so a converged warp takes one branch
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// shared -> global tile store (bulk async group); rows/columns outside the tensor map's extents are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(map), "r"((uint32_t)__cvta_generic_to_shared(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Programmatic dependent launch: a kernel launched with programmaticStreamSerialization may start while its predecessor
// drains; pdl_wait() blocks until every prerequisite grid has completed and its memory is visible, pdl_trigger() lets the
// next kernel in the stream begin its prologue (barrier init, TMEM alloc, descriptor prefetch) on idle SMs.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_c),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = sm_100).
//   bits [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version | [61,64) layout type
template <int BK>
__device__ __forceinline__ uint64_t make_kmajor_desc(uint32_t smem_addr) {
    constexpr uint64_t layout = (BK == 64) ? 2ull : (BK == 32) ? 4ull : 6ull;     // SW128 / SW64 / SW32
    constexpr uint64_t sbo = (8 * BK * 2) >> 4;                                  // 8 rows of BK fp16
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// ------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------
constexpr int kThreads = 320;          // 2 control warps + 8 epilogue warps
constexpr int kEpiWarps = 8;
constexpr int BM = 128;


// ------------------------------------------------------------------------------------------------
// host side: tensor maps
// ------------------------------------------------------------------------------------------------
inline PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    });
    RVCB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available (no CUDA driver?)");
    return fn;
}

inline CUtensorMapSwizzle swizzle_for(int bk) {
    return bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B;
}

inline void encode_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                       const cuuint32_t* box, int bk) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    RVCB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
    for (int i = 0; i < rank - 1; ++i) RVCB_CHECK(strides_bytes[i] % 16 == 0, "TMA stride must be a multiple of 16 bytes");
    CUresult r = get_encode_fn()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(bk), CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RVCB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}


// general form: explicit element type and swizzle (fp32 residual / output tiles of the weight-stationary kernel)
inline void encode_map_ex(CUtensorMap* map, const void* base, CUtensorMapDataType dtype, int elem_bytes, int rank, const cuuint64_t* dims,
                          const cuuint64_t* strides_bytes, const cuuint32_t* box, CUtensorMapSwizzle swz) {
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    (void)elem_bytes;
    RVCB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
    for (int i = 0; i < rank - 1; ++i) RVCB_CHECK(strides_bytes[i] % 16 == 0, "TMA stride must be a multiple of 16 bytes");
    CUresult r = get_encode_fn()(map, dtype, rank, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                                 CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    RVCB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
}

// launch helper: cudaLaunchKernelEx with the programmatic-stream-serialization attribute (enabled with RVCB_PDL=1)
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args&&... args) {
    static int use_pdl = -1, pdl_max_grid = 100;
    if (use_pdl < 0) {
        const char* e = getenv("RVCB_PDL");
        // measured (profiles/README.md): 1 (every launch) is neutral-to-negative -- early CTAs of big grids park on SMs the other
        // stream could use and 226 KB CTAs cannot co-reside anyway.  2: only launches of at most RVCB_PDL_MAXGRID CTAs, whose
        // prologue (TMEM alloc, barrier init, tensor-map fetch) then overlaps the predecessor's tail on idle SMs.
        use_pdl = (e && e[0] == '1') ? 1 : ((e && e[0] == '2') ? 2 : 0);
        if (const char* m = getenv("RVCB_PDL_MAXGRID")) pdl_max_grid = atoi(m);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (use_pdl == 1 || (use_pdl == 2 && grid <= pdl_max_grid)) ? 1 : 0;
    CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...));
}

// per-launch profiling hooks (defined in gemm_tc.cu)
struct ProfInfo { int M, N, kb, BK, BN, batch, nseg, tiles; double bytes = 0; };   // bytes: algorithmic HBM bytes (WS kernel)
bool gemm_prof_on();
void gemm_prof_record_begin(cudaStream_t stream);
void gemm_prof_record_end(cudaStream_t stream, const ProfInfo& info);

// weight-stationary halo convolution kernel (gemm_ws.cu); returns false if the launch does not qualify
bool gemm_ws_try(const GemmArgs& g, cudaStream_t stream);
// cluster split-K for small-M, long-K launches (gemm_sk.cu); returns false when the launch does not qualify
bool gemm_sk_try(const GemmArgs& g, cudaStream_t stream);

}  // namespace rvcb
