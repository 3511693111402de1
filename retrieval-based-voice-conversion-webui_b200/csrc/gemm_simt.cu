// Plain SIMT restatement of the implicit-GEMM contract in gemm.cuh.
// One thread per output element, fp32 accumulation in the same K order as the tensor-core
// kernel walks it.  Used by the tests to validate gemm_tc (descriptor / swizzle / segment logic)
// and selectable with RVCB_GEMM=simt for debugging.  It is a CUDA kernel, not a CPU fallback.
#include "gemm.cuh"

#include <cstdlib>
#include <cstring>

namespace rvcb {

struct SimtParams {
    const __half* A; long lda; int a_rows, a_cols, W;
    const __half* B; long ldb; int b_rows, b_cols;
    int M, N, bk, nseg, batch;
    long a_row_z, a_col_z, b_row_z, b_col_z, c_z, bias_z; int b_col0;
    const float* bias; int bias_per_row;
    const float* res1; long ldres1; const float* res2; long ldres2;
    float alpha; int act1; float act1_p; int act2; float act2_p; int gate;
    float* out32; long ld32; __half* out16; long ld16; int up2_C;
    GemmSeg seg[GEMM_MAX_SEG];
};

__device__ __forceinline__ float simt_dot(const SimtParams& p, int z, int m, int n) {
    float acc = 0.f;
    int kb = 0;
    const long brow = (long)n + z * p.b_row_z;
    for (int s = 0; s < p.nseg; ++s) {
        const GemmSeg sg = p.seg[s];
        long arow_base;
        bool row_valid;
        if (p.W == 0) {
            const long r = (long)m + sg.row_off + z * p.a_row_z;
            row_valid = r >= 0 && r < p.a_rows;
            arow_base = r * p.lda;
        } else {
            const int h = m / p.W + sg.row_off, w = m % p.W + sg.dw;
            row_valid = h >= 0 && h < p.a_rows && w >= 0 && w < p.W;
            arow_base = ((long)h * p.W + w) * p.lda;
        }
        for (int kc = 0; kc < sg.nk; ++kc, ++kb) {
            for (int k = 0; k < p.bk; ++k) {
                const long ac = z * p.a_col_z + sg.col_off + kc * p.bk + k;
                const long bc = p.b_col0 + z * p.b_col_z + (long)kb * p.bk + k;
                float a = 0.f, b = 0.f;
                if (row_valid && ac >= 0 && ac < p.a_cols) a = __half2float(p.A[arow_base + ac]);
                if (brow >= 0 && brow < p.b_rows && bc >= 0 && bc < p.b_cols) b = __half2float(p.B[brow * p.ldb + bc]);
                acc = fmaf(a, b, acc);
            }
        }
    }
    return acc;
}

__global__ void gemm_simt_kernel(const __grid_constant__ SimtParams p) {
    const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ncols = p.gate ? p.N / 2 : p.N;
    const long per_z = (long)p.M * ncols;
    if (idx >= per_z * p.batch) return;
    const int z = (int)(idx / per_z);
    const long r = idx - (long)z * per_z;
    const int m = (int)(r / ncols), nc = (int)(r % ncols);
    const float* bz = p.bias ? p.bias + z * p.bias_z : nullptr;
    if (p.gate) {
        float t0 = simt_dot(p, z, m, 2 * nc), t1 = simt_dot(p, z, m, 2 * nc + 1);
        if (bz) {
            t0 += p.bias_per_row ? p.bias[m] : bz[2 * nc];
            t1 += p.bias_per_row ? p.bias[m] : bz[2 * nc + 1];
        }
        if (p.res1) {
            t0 += p.res1[z * p.c_z + (long)m * p.ldres1 + 2 * nc];
            t1 += p.res1[z * p.c_z + (long)m * p.ldres1 + 2 * nc + 1];
        }
        const float g = tanhf(t0) * (1.f / (1.f + expf(-t1)));
        if (p.out32) p.out32[z * p.c_z + (long)m * p.ld32 + nc] = g;
        if (p.out16) p.out16[z * p.c_z + (long)m * p.ld16 + nc] = __float2half_rn(g);
        return;
    }
    const int n = nc;
    float t = simt_dot(p, z, m, n);
    if (bz) t += p.bias_per_row ? p.bias[m] : bz[n];
    if (p.res1) t += p.res1[z * p.c_z + (long)m * p.ldres1 + n];
    t = apply_act(t, p.act1, p.act1_p);
    float v = p.alpha * t;
    if (p.res2) v += p.res2[z * p.c_z + (long)m * p.ldres2 + n];
    long orow = m, ocol = n;
    if (p.up2_C) {
        const int ab = n / p.up2_C, co = n - ab * p.up2_C;
        const int ii = m / p.W, jj = m - ii * p.W;
        orow = (long)(2 * ii + (ab >> 1)) * (2 * p.W) + 2 * jj + (ab & 1);
        ocol = co;
    }
    if (p.out32) p.out32[z * p.c_z + orow * p.ld32 + ocol] = v;
    if (p.out16) p.out16[z * p.c_z + orow * p.ld16 + ocol] = __float2half_rn(apply_act(v, p.act2, p.act2_p));
}

void gemm_simt(const GemmArgs& g, cudaStream_t stream) {
    SimtParams p{};
    p.A = g.A; p.lda = g.lda; p.a_rows = g.a_rows; p.a_cols = g.a_cols; p.W = g.conv2d_W;
    p.B = g.B; p.ldb = g.ldb; p.b_rows = g.b_rows; p.b_cols = g.b_cols;
    p.M = g.M; p.N = g.N; p.bk = g.block_k; p.nseg = g.nseg; p.batch = g.batch;
    p.a_row_z = g.a_row_z; p.a_col_z = g.a_col_z; p.b_row_z = g.b_row_z; p.b_col_z = g.b_col_z; p.c_z = g.c_z;
    p.bias_z = g.bias_z; p.b_col0 = g.b_col0;
    p.bias = g.bias; p.bias_per_row = g.bias_per_row; p.res1 = g.res1; p.ldres1 = g.ldres1; p.res2 = g.res2; p.ldres2 = g.ldres2;
    p.alpha = g.alpha; p.act1 = g.act1; p.act1_p = g.act1_p; p.act2 = g.act2; p.act2_p = g.act2_p; p.gate = g.gate;
    p.out32 = g.out32; p.ld32 = g.ld32; p.out16 = g.out16; p.ld16 = g.ld16; p.up2_C = g.up2_C;
    memcpy(p.seg, g.seg, sizeof(GemmSeg) * g.nseg);
    const long total = (long)g.M * (g.gate ? g.N / 2 : g.N) * g.batch;
    gemm_simt_kernel<<<(unsigned)ceil_div_l(total, 256), 256, 0, stream>>>(p);
    KERNEL_CHECK();
    count_launch();
}

void gemm(const GemmArgs& g, cudaStream_t stream) {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("RVCB_GEMM");
        mode = (e && strcmp(e, "simt") == 0) ? 1 : 0;
    }
    if (mode == 1) gemm_simt(g, stream);
    else gemm_tc(g, stream);
}

}  // namespace rvcb
