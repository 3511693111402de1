// IVF-Flat (L2, nprobe = 1) retrieval + blend, and the brute-force top-1 sweep kernel.
// Replaces faiss index.search(npy, k=8) / reconstruct_n and the numpy blend at
// infer/modules/vc/pipeline.py:113-138 and infer/lib/rtrvc.py:169-185 (both run on the CPU there).
//
// HBM-bound scan: every database vector is read once per 32-query tile with coalesced float4
// loads (one warp = one 128-float chunk), squared differences are accumulated per lane and
// combined with an xor-butterfly of warp shuffles.  The summation order is FIXED (see
// oracle/ivf.py: lane l owns elements 128c+4l+e, c outer, e inner, separate rn multiply/add,
// butterfly 16,8,4,2,1) so distances and therefore arg-min / top-k indices are bit-exact
// against the oracle, independent of grid shape.
#include "../../include/rvcb200.h"
#include "api_macros.h"
#include "common.cuh"

#include <algorithm>
#include <vector>

using namespace rvcb;


constexpr int QT = 32;                  // queries per tile
constexpr float KNN_FLT_MAX = 3.4028235e38f;

struct rvcb_index {
    float* centroids = nullptr;   // [nlist, d]
    float* vectors = nullptr;     // [ntotal, d], id order (= big_npy / reconstruct_n)
    long long* list_off = nullptr;
    long long* list_ids = nullptr;
    int nlist = 0, d = 0;
    long long ntotal = 0;
    // workspace
    unsigned long long* best = nullptr;
    int best_cap = 0;
    std::vector<unsigned long long*> retired;     // outgrown workspaces (see ensure_ws)
    ~rvcb_index() {
        cudaFree(centroids); cudaFree(vectors); cudaFree(list_off); cudaFree(list_ids); cudaFree(best);
        for (auto* p : retired) cudaFree(p);
    }
};

template <int CH>
__device__ __forceinline__ float lane_order_dist(const float4 (&q)[CH], const float4 (&v)[CH]) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const float4 a = q[c], b = v[c];
        float d;
        d = __fsub_rn(a.x, b.x); acc = __fadd_rn(acc, __fmul_rn(d, d));
        d = __fsub_rn(a.y, b.y); acc = __fadd_rn(acc, __fmul_rn(d, d));
        d = __fsub_rn(a.z, b.z); acc = __fadd_rn(acc, __fmul_rn(d, d));
        d = __fsub_rn(a.w, b.w); acc = __fadd_rn(acc, __fmul_rn(d, d));
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, s));
    return acc;
}

// best[q] = min over db rows of pack(dist, idx).  grid = (query tiles, db splits)
template <int CH>
__global__ void __launch_bounds__(256) knn_top1_kernel(const float* __restrict__ db, long long n, const float* __restrict__ q, int nq,
                                                       unsigned long long* __restrict__ best) {
    extern __shared__ float4 qs[];                 // [QT][d/4]
    constexpr int d = CH * 128;
    constexpr int d4 = d >> 2;
    const int q0 = blockIdx.x * QT;
    const int nq_tile = min(QT, nq - q0);
    for (int i = threadIdx.x; i < QT * d4; i += blockDim.x) {
        const int qi = i / d4;
        qs[i] = qi < nq_tile ? reinterpret_cast<const float4*>(q + (long)(q0 + qi) * d)[i - qi * d4] : make_float4(0, 0, 0, 0);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long per = (n + gridDim.y - 1) / gridDim.y;
    const long long r0 = (long long)blockIdx.y * per;
    const long long r1 = min(n, r0 + per);
    float bestd = INFINITY;
    unsigned int besti = 0xffffffffu;
    for (long long r = r0 + warp; r < r1; r += 8) {
        float4 v[CH];
        const float4* vr = reinterpret_cast<const float4*>(db + r * d);
#pragma unroll
        for (int c = 0; c < CH; ++c) v[c] = __ldg(vr + c * 32 + lane);
        for (int qi = 0; qi < nq_tile; ++qi) {
            float4 qq[CH];
#pragma unroll
            for (int c = 0; c < CH; ++c) qq[c] = qs[qi * d4 + c * 32 + lane];
            const float dist = lane_order_dist<CH>(qq, v);
            if (lane == qi && (dist < bestd || (dist == bestd && (unsigned int)r < besti))) {
                bestd = dist;
                besti = (unsigned int)r;
            }
        }
    }
    if (lane < nq_tile && besti != 0xffffffffu) {
        const unsigned long long key = ((unsigned long long)__float_as_uint(bestd) << 32) | besti;
        atomicMin(&best[q0 + lane], key);
    }
}

__global__ void unpack_best_kernel(const unsigned long long* __restrict__ best, int nq, float* __restrict__ D, long long* __restrict__ I,
                                   int* __restrict__ lists) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const unsigned long long k = best[i];
    const bool ok = k != ~0ull;
    if (D) D[i] = ok ? __uint_as_float((unsigned int)(k >> 32)) : KNN_FLT_MAX;
    if (I) I[i] = ok ? (long long)(k & 0xffffffffu) : -1;
    if (lists) lists[i] = ok ? (int)(k & 0xffffffffu) : -1;
}

// one warp per query: exact scan of the probed list, ascending top-k (ties -> lower list position)
template <int K, int CH>
__global__ void __launch_bounds__(256) ivf_scan_kernel(const float* __restrict__ vectors, const long long* __restrict__ list_off,
                                                       const long long* __restrict__ list_ids, const unsigned long long* __restrict__ best,
                                                       const float* __restrict__ q, int nq, float* __restrict__ D, long long* __restrict__ I) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qi = blockIdx.x * 8 + warp;
    if (qi >= nq) return;
    constexpr int d = CH * 128;
    float4 qq[CH];
    const float4* qr = reinterpret_cast<const float4*>(q + (long)qi * d);
#pragma unroll
    for (int c = 0; c < CH; ++c) qq[c] = qr[c * 32 + lane];
    float bd[K];
    long long bi[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { bd[k] = KNN_FLT_MAX; bi[k] = -1; }
    const unsigned long long key = best[qi];
    if (key != ~0ull) {
        const int l = (int)(key & 0xffffffffu);
        const long long a = list_off[l], b = list_off[l + 1];
        for (long long p = a; p < b; ++p) {
            const long long id = list_ids[p];
            float4 v[CH];
            const float4* vr = reinterpret_cast<const float4*>(vectors + id * d);
#pragma unroll
            for (int c = 0; c < CH; ++c) v[c] = __ldg(vr + c * 32 + lane);
            const float dist = lane_order_dist<CH>(qq, v);
            // sorted insert; strict '<' keeps the earlier list position first on ties.  A slot that
            // still holds the (FLT_MAX, -1) filler is always replaced.
            if (dist < bd[K - 1] || bi[K - 1] < 0) {
                bd[K - 1] = dist;
                bi[K - 1] = id;
#pragma unroll
                for (int k = K - 1; k > 0; --k) {
                    const bool sw = (bi[k - 1] < 0) || (bd[k] < bd[k - 1]);
                    if (sw) {
                        const float td = bd[k]; bd[k] = bd[k - 1]; bd[k - 1] = td;
                        const long long ti = bi[k]; bi[k] = bi[k - 1]; bi[k - 1] = ti;
                    }
                }
            }
        }
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            D[(long)qi * K + k] = bi[k] < 0 ? KNN_FLT_MAX : bd[k];
            I[(long)qi * K + k] = bi[k];
        }
    }
}

// pipeline.py:129-138 in numpy's evaluation order (bit-exact for k = 8)
__global__ void blend_kernel(const float* __restrict__ vectors, long long ntotal, int d, const float* __restrict__ feats, int k,
                             const float* __restrict__ D, const long long* __restrict__ I, float rate, float* __restrict__ out) {
    const int qi = blockIdx.x;
    __shared__ float w[32];
    __shared__ long long ids[32];
    if (threadIdx.x < k) {
        const float s = D[(long)qi * k + threadIdx.x];
        const float r = __fdiv_rn(1.f, s);
        w[threadIdx.x] = __fmul_rn(r, r);
        long long id = I[(long)qi * k + threadIdx.x];
        if (id < 0) id += ntotal;                       // numpy negative index: big_npy[-1]
        ids[threadIdx.x] = id;
    }
    __syncthreads();
    float sum;
    if (k == 8) {   // numpy pairwise sum for n == 8
        sum = __fadd_rn(__fadd_rn(__fadd_rn(w[0], w[1]), __fadd_rn(w[2], w[3])), __fadd_rn(__fadd_rn(w[4], w[5]), __fadd_rn(w[6], w[7])));
    } else {
        sum = 0.f;
        for (int j = 0; j < k; ++j) sum = __fadd_rn(sum, w[j]);
    }
    const float omr = (float)(1.0 - (double)rate);
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        float acc = 0.f;
        for (int j = 0; j < k; ++j) {
            const float wj = __fdiv_rn(w[j], sum);
            const float p = __fmul_rn(vectors[ids[j] * d + c], wj);
            acc = (j == 0) ? p : __fadd_rn(acc, p);
        }
        out[(long)qi * d + c] = __fadd_rn(__fmul_rn(acc, rate), __fmul_rn(omr, feats[(long)qi * d + c]));
    }
}

static void top1(const float* db, long long n, int d, const float* q, int nq, unsigned long long* best, cudaStream_t st) {
    RVCB_CHECK(d % 128 == 0 && d <= 1024, "knn: d must be a multiple of 128 and <= 1024");
    RVCB_CHECK(n < 0xffffffffLL, "knn: database too large");
    CUDA_CHECK(cudaMemsetAsync(best, 0xff, sizeof(unsigned long long) * nq, st));
    const int qtiles = ceil_div(nq, QT);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // enough splits to fill the machine (2 CTAs/SM by shared memory), at least 64 rows per split
    long long splits = (2LL * sms + qtiles - 1) / qtiles;
    const long long max_splits = (n + 63) / 64;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    const size_t smem = (size_t)QT * d * sizeof(float);
#define RVCB_TOP1(CH)                                                                                              \
    case CH: {                                                                                                     \
        static bool attr = false;                                                                                  \
        if (!attr) {                                                                                               \
            CUDA_CHECK(cudaFuncSetAttribute(knn_top1_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            attr = true;                                                                                           \
        }                                                                                                          \
        knn_top1_kernel<CH><<<dim3(qtiles, (unsigned)splits), 256, smem, st>>>(db, n, q, nq, best);                \
        break;                                                                                                     \
    }
    switch (d / 128) {
        RVCB_TOP1(1) RVCB_TOP1(2) RVCB_TOP1(3) RVCB_TOP1(4) RVCB_TOP1(6) RVCB_TOP1(8)
        default: RVCB_CHECK(false, "knn: unsupported dimension (128, 256, 384, 512, 768, 1024)");
    }
#undef RVCB_TOP1
    KERNEL_CHECK();
    count_launch();
}

// The coarse-assignment workspace grows geometrically and the outgrown block is kept until the index dies: a CUDA graph
// captured for a shorter utterance still points at it (see Arena::reserve).
static void ensure_ws(rvcb_index* ix, int nq) {
    if (nq > ix->best_cap) {
        const int want = std::max(nq, ix->best_cap + ix->best_cap / 2);
        unsigned long long* nb = nullptr;
        CUDA_CHECK(cudaMalloc(&nb, sizeof(unsigned long long) * want));
        if (ix->best) ix->retired.push_back(ix->best);
        ix->best = nb;
        ix->best_cap = want;
    }
}

extern "C" {

int rvcb_index_create(const float* centroids, int nlist, const float* vectors, int64_t ntotal, int d, const int64_t* list_off,
                      const int64_t* list_ids, rvcb_index** out) {
    RVCB_API_BEGIN
    RVCB_CHECK(centroids && vectors && list_off && list_ids && out, "null argument");
    RVCB_CHECK(d % 128 == 0 && d <= 1024, "index: d must be a multiple of 128 (<= 1024)");
    auto* ix = new rvcb_index();
    try {
        ix->nlist = nlist; ix->d = d; ix->ntotal = ntotal;
        ix->centroids = dev_upload(centroids, (size_t)nlist * d);
        ix->vectors = dev_upload(vectors, (size_t)ntotal * d);
        ix->list_off = dev_upload((const long long*)list_off, (size_t)nlist + 1);
        ix->list_ids = dev_upload((const long long*)list_ids, (size_t)ntotal);
    } catch (...) {
        delete ix;
        throw;
    }
    *out = ix;
    RVCB_API_END
}

int64_t rvcb_index_ntotal(const rvcb_index* ix) { return ix ? ix->ntotal : -1; }

int rvcb_index_search(rvcb_index* ix, const float* d_q, int nq, int k, float* d_D, int64_t* d_I, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(ix && d_q && d_D && d_I && nq > 0, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    ensure_ws(ix, nq);
    top1(ix->centroids, ix->nlist, ix->d, d_q, nq, ix->best, st);      // coarse quantiser, nprobe = 1
    const int grid = ceil_div(nq, 8);
#define RVCB_SCAN(K, CH) ivf_scan_kernel<K, CH><<<grid, 256, 0, st>>>(ix->vectors, ix->list_off, ix->list_ids, ix->best, d_q, nq, d_D, (long long*)d_I)
#define RVCB_SCAN_K(CH)                                                       \
    if (k == 8) RVCB_SCAN(8, CH);                                             \
    else if (k == 1) RVCB_SCAN(1, CH);                                        \
    else if (k == 4) RVCB_SCAN(4, CH);                                        \
    else if (k == 16) RVCB_SCAN(16, CH);                                      \
    else RVCB_CHECK(false, "index_search: k must be 1, 4, 8 or 16");
    if (ix->d == 768) { RVCB_SCAN_K(6) }
    else if (ix->d == 256) { RVCB_SCAN_K(2) }
    else if (ix->d == 128) { RVCB_SCAN_K(1) }
    else if (ix->d == 512) { RVCB_SCAN_K(4) }
    else if (ix->d == 1024) { RVCB_SCAN_K(8) }
    else RVCB_CHECK(false, "index_search: unsupported dimension");
#undef RVCB_SCAN_K
#undef RVCB_SCAN
    KERNEL_CHECK();
    count_launch();
    RVCB_API_END
}

int rvcb_index_blend(rvcb_index* ix, const float* d_feats_in, int nq, int k, const float* d_D, const int64_t* d_I, float index_rate,
                     float* d_feats_out, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(ix && d_feats_in && d_D && d_I && d_feats_out && k >= 1 && k <= 32, "bad argument");
    blend_kernel<<<nq, 256, 0, (cudaStream_t)stream>>>(ix->vectors, ix->ntotal, ix->d, d_feats_in, k, d_D, (const long long*)d_I, index_rate,
                                                       d_feats_out);
    KERNEL_CHECK();
    count_launch();
    RVCB_API_END
}

int rvcb_knn_bruteforce_top1(const float* d_db, int64_t n, int d, const float* d_q, int nq, float* d_D, int64_t* d_I, void* stream) {
    RVCB_API_BEGIN
    RVCB_CHECK(d_db && d_q && nq > 0, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    static unsigned long long* ws = nullptr;
    static int ws_cap = 0;
    if (nq > ws_cap) {      // outgrown blocks are leaked on purpose (a captured graph may still reference them); geometric growth
        const int want = std::max(nq, ws_cap + ws_cap / 2);
        CUDA_CHECK(cudaMalloc(&ws, sizeof(unsigned long long) * want));
        ws_cap = want;
    }
    top1(d_db, n, d, d_q, nq, ws, st);
    unpack_best_kernel<<<ceil_div(nq, 256), 256, 0, st>>>(ws, nq, d_D, (long long*)d_I, nullptr);
    KERNEL_CHECK();
    count_launch();
    RVCB_API_END
}

void rvcb_index_destroy(rvcb_index* ix) { delete ix; }

}  // extern "C"
