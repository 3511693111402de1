// Host-side weight container (the rvcb_weights handle) and packers that turn the reference's
// state_dict tensors into the fp16 K-major "B operand" layouts the implicit-GEMM engine reads.
#pragma once
#include "common.cuh"

#include <cmath>
#include <string>
#include <unordered_map>
#include <vector>

struct rvcb_weights {
    struct Tensor {
        std::vector<int64_t> shape;
        std::vector<float> data;
        int64_t numel() const { return (int64_t)data.size(); }
        int64_t dim(int i) const { return shape.at(i); }
    };
    std::unordered_map<std::string, Tensor> t;
    bool has(const std::string& n) const { return t.find(n) != t.end(); }
    const Tensor& get(const std::string& n) const {
        auto it = t.find(n);
        RVCB_CHECK(it != t.end(), "missing weight '" + n + "'");
        return it->second;
    }
};

namespace rvcb {

using WT = rvcb_weights::Tensor;

// device allocations owned by a model handle
struct DevOwner {
    std::vector<void*> ptrs;
    template <typename T>
    T* upload(const std::vector<T>& h) {
        T* d = nullptr;
        CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(T)));
        if (!h.empty()) CUDA_CHECK(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
        ptrs.push_back(d);
        return d;
    }
    template <typename T>
    T* alloc(size_t n) {
        T* d = nullptr;
        CUDA_CHECK(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)));
        CUDA_CHECK(cudaMemset(d, 0, std::max<size_t>(n, 1) * sizeof(T)));
        ptrs.push_back(d);
        return d;
    }
    ~DevOwner() {
        for (void* p : ptrs) cudaFree(p);
    }
};

// fp16 matrix [rows, cols] on the device, K (cols) contiguous
struct PackedB {
    __half* d = nullptr;
    int rows = 0, cols = 0;
};

inline PackedB upload_half(DevOwner& own, const std::vector<float>& h, int rows, int cols) {
    std::vector<__half> hh((size_t)rows * cols);
    for (size_t i = 0; i < hh.size(); ++i) hh[i] = __float2half_rn(h[i]);
    PackedB p;
    p.d = own.upload(hh);
    p.rows = rows;
    p.cols = cols;
    return p;
}

inline int pad_to(int x, int m) { return (x + m - 1) / m * m; }

// Linear / 1x1 conv: W [N, K] -> [N_pad(16), K_pad(bk)]; optional per-row scale
inline PackedB pack_linear(DevOwner& own, const float* W, int N, int K, int bk = 64, const float* row_scale = nullptr) {
    const int Np = pad_to(N, 16), Kp = pad_to(K, bk);
    std::vector<float> h((size_t)Np * Kp, 0.f);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) h[(size_t)n * Kp + k] = W[(size_t)n * K + k] * (row_scale ? row_scale[n] : 1.f);
    return upload_half(own, h, Np, Kp);
}

// Conv1d weight [Cout, Cin, k] -> [Cout_pad, k*Cin_pad], tap-major (B[co, j*Cin_pad + ci])
inline PackedB pack_conv1d(DevOwner& own, const float* W, int Cout, int Cin, int k, int bk = 64, const float* row_scale = nullptr) {
    const int Np = pad_to(Cout, 16), Cp = pad_to(Cin, bk);
    std::vector<float> h((size_t)Np * k * Cp, 0.f);
    for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
            for (int j = 0; j < k; ++j)
                h[((size_t)co * k + j) * Cp + ci] = W[((size_t)co * Cin + ci) * k + j] * (row_scale ? row_scale[co] : 1.f);
    return upload_half(own, h, Np, k * Cp);
}

// ConvTranspose1d weight [Cin, Cout, k], stride s, padding p -> polyphase [s*Cout, 3*Cin_pad]
// row = r*Cout + co; segment index di <-> delta = di-1; tap j = r + p - delta*s
// Optional extra K columns (Mp of them) fold the NSF noise convolution (Conv1d(1, Cout, kn, stride_n) over the harmonic
// source, nsf.py:176-183) into the same GEMM: the A operand carries har[tin*s*stride_n + m - pad] in column m, and the
// weight of output phase r at column m is Wn[co, m - r*stride_n].
// Number of input-frame shifts d = -dm..dm a ConvTranspose1d(k, stride s, padding p) needs: output phase r of frame q reads
// frame q + d with kernel tap j = r + p - d*s, so dm = max(floor((s - 1 + p) / s), floor((k - 1 - p) / s))  (1 when k < 3s + 2).
inline int convT1d_reach(int k, int s, int p) {
    const int hi = (s - 1 + p) / s, lo = (k - 1 - p) / s;
    return hi > lo ? (hi > 1 ? hi : 1) : (lo > 1 ? lo : 1);
}
inline PackedB pack_convT1d(DevOwner& own, const float* W, int Cin, int Cout, int k, int s, int p, int bk = 64,
                            const float* Wn = nullptr, int kn = 0, int stride_n = 0, int Mp = 0) {
    const int dm = convT1d_reach(k, s, p), ntap = 2 * dm + 1;
    const int Cp = pad_to(Cin, bk), Np = pad_to(s * Cout, 16);
    const int Kt = ntap * Cp + Mp;
    std::vector<float> h((size_t)Np * Kt, 0.f);
    for (int r = 0; r < s; ++r) {
        for (int di = 0; di < ntap; ++di) {
            const int j = r + p - (di - dm) * s;
            if (j < 0 || j >= k) continue;
            for (int co = 0; co < Cout; ++co)
                for (int ci = 0; ci < Cin; ++ci)
                    h[(size_t)(r * Cout + co) * Kt + di * Cp + ci] = W[((size_t)ci * Cout + co) * k + j];
        }
        for (int m = 0; m < Mp && Wn; ++m) {
            const int j = m - r * stride_n;
            if (j < 0 || j >= kn) continue;
            for (int co = 0; co < Cout; ++co) h[(size_t)(r * Cout + co) * Kt + ntap * Cp + m] = Wn[(size_t)co * kn + j];
        }
    }
    return upload_half(own, h, Np, Kt);
}

// Conv2d 3x3 weight [Cout, Cin, 3, 3] (+ folded per-output scale) -> [Cout_pad, 9*Cin_pad], (dh, dw, ci) order
inline PackedB pack_conv2d_3x3(DevOwner& own, const float* W, int Cout, int Cin, int bk, const float* row_scale = nullptr) {
    const int Np = pad_to(Cout, 16), Cp = pad_to(Cin, bk);
    std::vector<float> h((size_t)Np * 9 * Cp, 0.f);
    for (int co = 0; co < Cout; ++co)
        for (int ci = 0; ci < Cin; ++ci)
            for (int t = 0; t < 9; ++t)
                h[((size_t)co * 9 + t) * Cp + ci] = W[((size_t)co * Cin + ci) * 9 + t] * (row_scale ? row_scale[co] : 1.f);
    return upload_half(own, h, Np, 9 * Cp);
}

// ConvTranspose2d(3x3, stride 2, pad 1, output_padding 1) weight [Cin, Cout, 3, 3] (+ folded scale per Cout)
// -> [4*Cout, 4*Cin_pad]; row = (a*2+b)*Cout + co; segment si <-> (dh, dw) = (si>>1, si&1); kh = a+1-2dh, kw = b+1-2dw
inline PackedB pack_convT2d_up2(DevOwner& own, const float* W, int Cin, int Cout, int bk, const float* col_scale = nullptr) {
    const int Cp = pad_to(Cin, bk);
    std::vector<float> h((size_t)4 * Cout * 4 * Cp, 0.f);
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b)
            for (int si = 0; si < 4; ++si) {
                const int dh = si >> 1, dw = si & 1;
                const int kh = a + 1 - 2 * dh, kw = b + 1 - 2 * dw;
                if (kh < 0 || kh > 2 || kw < 0 || kw > 2) continue;
                for (int co = 0; co < Cout; ++co)
                    for (int ci = 0; ci < Cin; ++ci)
                        h[((size_t)((a * 2 + b) * Cout + co) * 4 + si) * Cp + ci] =
                            W[(((size_t)ci * Cout + co) * 3 + kh) * 3 + kw] * (col_scale ? col_scale[co] : 1.f);
            }
    return upload_half(own, h, 4 * Cout, 4 * Cp);
}

inline float* upload_vec(DevOwner& own, const std::vector<float>& v) { return own.upload(v); }

// fold legacy weight-norm pairs if the container holds them: w = g * v / ||v|| (norm over all dims but 0)
inline std::vector<float> effective_weight(const rvcb_weights& w, const std::string& base) {
    if (w.has(base + ".weight")) return w.get(base + ".weight").data;
    std::string gk = base + ".weight_g", vk = base + ".weight_v";
    if (!w.has(gk)) {
        gk = base + ".parametrizations.weight.original0";
        vk = base + ".parametrizations.weight.original1";
    }
    const WT& g = w.get(gk);
    const WT& v = w.get(vk);
    const int64_t n0 = v.dim(0), inner = v.numel() / n0;
    RVCB_CHECK(g.numel() == n0, "weight_g shape mismatch for " + base);
    std::vector<float> out(v.data.size());
    for (int64_t i = 0; i < n0; ++i) {
        double nrm = 0;
        for (int64_t j = 0; j < inner; ++j) nrm += (double)v.data[i * inner + j] * v.data[i * inner + j];
        const float sc = (float)(g.data[i] / std::sqrt(nrm));
        for (int64_t j = 0; j < inner; ++j) out[i * inner + j] = v.data[i * inner + j] * sc;
    }
    return out;
}

}  // namespace rvcb
