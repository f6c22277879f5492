// Training-mode elementwise / reduction kernels around the conv GEMMs: BatchNorm batch statistics,
// BN-apply + activation + temporal mask (forward and backward), weight (un)packing, layout casts.
//
// Reference semantics (file:line): nn.BatchNorm1d(momentum=0.1, eps=1e-5) in training mode inside
// ConvBn1d.forward models.py:127-139 -- biased batch variance over ALL B*t positions (padded frames
// included, the mask is applied after the activation), running_var updated with the unbiased one;
// ResidualActivation.forward :357-371 (hardtanh passes gradient only for a < z < b; relu for z > 0;
// leaky_relu slope for z <= 0); temporal mask multiply :136-138 is part of the autograd graph.
// All of these are HBM bound: bf16 activations, 16-byte vector accesses, fp32 math.
#include "common.cuh"
#include "../../include/convasr_b200.h"
#include <atomic>

namespace cab {
extern std::atomic<int64_t> g_launch_count;

__device__ __forceinline__ float act_fwd(float z, int act, float a, float b) {
    switch (act) {
        case CAB_ACT_RELU: return fmaxf(z, 0.f);
        case CAB_ACT_HARDTANH: return fminf(fmaxf(z, a), b);
        case CAB_ACT_LEAKY_RELU: return z > 0.f ? z : z * a;
        default: return z;
    }
}
__device__ __forceinline__ float act_grad(float z, int act, float a, float b) {
    switch (act) {
        case CAB_ACT_RELU: return z > 0.f ? 1.f : 0.f;
        case CAB_ACT_HARDTANH: return (z > a && z < b) ? 1.f : 0.f;
        case CAB_ACT_LEAKY_RELU: return z > 0.f ? 1.f : a;
        default: return 1.f;
    }
}

// ---------------------------------------------------------------------------------------
// per-channel sum / sum of squares of a bf16 [R, ld] matrix (R = B*t rows).
// block (64, 4): x = channel pair, y = row lane; grid (ceil(C/128), row blocks)
// ---------------------------------------------------------------------------------------
constexpr int kRowsPerBlock = 256;
__global__ void __launch_bounds__(256)
colstats_kernel(const __nv_bfloat16* __restrict__ x, int R, int C, int ld, float* __restrict__ out) {
    const int c = (blockIdx.x * 64 + threadIdx.x) * 2;
    const int r0 = blockIdx.y * kRowsPerBlock;
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    if (c < C) {
        const int r1 = min(R, r0 + kRowsPerBlock);
        for (int r = r0 + threadIdx.y; r < r1; r += 4) {
            const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(x + (size_t)r * ld + c));
            s0 += v.x; s1 += v.y;
            q0 = fmaf(v.x, v.x, q0); q1 = fmaf(v.y, v.y, q1);
        }
    }
    __shared__ float sm[4][64][4];
    sm[threadIdx.y][threadIdx.x][0] = s0; sm[threadIdx.y][threadIdx.x][1] = s1;
    sm[threadIdx.y][threadIdx.x][2] = q0; sm[threadIdx.y][threadIdx.x][3] = q1;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        for (int y = 1; y < 4; ++y) {
            s0 += sm[y][threadIdx.x][0]; s1 += sm[y][threadIdx.x][1];
            q0 += sm[y][threadIdx.x][2]; q1 += sm[y][threadIdx.x][3];
        }
        atomicAdd(out + c, s0);
        atomicAdd(out + C + c, q0);
        if (c + 1 < C) { atomicAdd(out + c + 1, s1); atomicAdd(out + C + c + 1, q1); }
    }
}

// mean / invstd / fused scale+shift, running statistics update (momentum, unbiased running_var)
__global__ void bn_finalize_kernel(const float* __restrict__ stats, int C, float n, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum,
                                   float* __restrict__ running_mean, float* __restrict__ running_var,
                                   float* __restrict__ out /* [4][C]: scale, shift, mean, invstd */) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float mean = stats[c] / n;
    const float var = fmaxf(stats[C + c] / n - mean * mean, 0.f);
    const float invstd = rsqrtf(var + eps);
    const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    out[c] = g * invstd;
    out[C + c] = b - mean * g * invstd;
    out[2 * C + c] = mean;
    out[3 * C + c] = invstd;
    if (running_mean) running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    if (running_var) running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (n / fmaxf(n - 1.f, 1.f));
}

// out = act(y * scale + shift) * (t < len_b); 8 channels per thread
__global__ void __launch_bounds__(256)
bn_act_mask_fwd_kernel(const __nv_bfloat16* __restrict__ y, const float* __restrict__ ss, int B, int T, int C, int ld,
                       int act, float a, float bb, const float* __restrict__ xlen, __nv_bfloat16* __restrict__ out) {
    const int vec_per_row = ld / 8;
    const size_t n = (size_t)B * T * vec_per_row;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % vec_per_row);
        const size_t row = i / vec_per_row;
        const int t = (int)(row % T), b = (int)(row / T);
        const int c0 = cv * 8;
        bool keep = true;
        if (xlen != nullptr) keep = t < frac_len(__ldg(xlen + b), T);
        uint4 v = *reinterpret_cast<const uint4*>(y + row * ld + c0);
        uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&w[j]));
            const int c = c0 + 2 * j;
            float z0 = 0.f, z1 = 0.f;
            if (keep && c < C) z0 = act_fwd(fmaf(f.x, __ldg(ss + c), __ldg(ss + C + c)), act, a, bb);
            if (keep && c + 1 < C) z1 = act_fwd(fmaf(f.y, __ldg(ss + c + 1), __ldg(ss + C + c + 1)), act, a, bb);
            w[j] = pack_bf16x2(z0, z1);
        }
        *reinterpret_cast<uint4*>(out + row * ld + c0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

// backward pass 1: per-channel sum(dz), sum(dz * xhat), dz = g * act'(z) * mask
__global__ void __launch_bounds__(256)
bn_act_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ g,
                         const float* __restrict__ ss, int B, int T, int C, int ld, int act, float a, float bb,
                         const float* __restrict__ xlen, float* __restrict__ out) {
    const int c = (blockIdx.x * 64 + threadIdx.x) * 2;
    const int R = B * T;
    const int r0 = blockIdx.y * kRowsPerBlock;
    float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
    if (c < C) {
        const float sc0 = ss[c], sh0 = ss[C + c], m0 = ss[2 * C + c], i0 = ss[3 * C + c];
        const bool has1 = c + 1 < C;
        const float sc1 = has1 ? ss[c + 1] : 0.f, sh1 = has1 ? ss[C + c + 1] : 0.f, m1 = has1 ? ss[2 * C + c + 1] : 0.f, i1 = has1 ? ss[3 * C + c + 1] : 0.f;
        const int r1 = min(R, r0 + kRowsPerBlock);
        for (int r = r0 + threadIdx.y; r < r1; r += 4) {
            const int t = r % T, b = r / T;
            if (xlen != nullptr && t >= frac_len(__ldg(xlen + b), T)) continue;
            const float2 yv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(y + (size_t)r * ld + c));
            const float2 gv = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(g + (size_t)r * ld + c));
            const float dz0 = gv.x * act_grad(fmaf(yv.x, sc0, sh0), act, a, bb);
            s0 += dz0; q0 = fmaf(dz0, (yv.x - m0) * i0, q0);
            if (has1) {
                const float dz1 = gv.y * act_grad(fmaf(yv.y, sc1, sh1), act, a, bb);
                s1 += dz1; q1 = fmaf(dz1, (yv.y - m1) * i1, q1);
            }
        }
    }
    __shared__ float sm[4][64][4];
    sm[threadIdx.y][threadIdx.x][0] = s0; sm[threadIdx.y][threadIdx.x][1] = s1;
    sm[threadIdx.y][threadIdx.x][2] = q0; sm[threadIdx.y][threadIdx.x][3] = q1;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        for (int yy = 1; yy < 4; ++yy) {
            s0 += sm[yy][threadIdx.x][0]; s1 += sm[yy][threadIdx.x][1];
            q0 += sm[yy][threadIdx.x][2]; q1 += sm[yy][threadIdx.x][3];
        }
        atomicAdd(out + c, s0);
        atomicAdd(out + C + c, q0);
        if (c + 1 < C) { atomicAdd(out + c + 1, s1); atomicAdd(out + C + c + 1, q1); }
    }
}

// backward pass 2: dy = scale * (dz - sum_dz/n - xhat * sum_dzx/n)
__global__ void __launch_bounds__(256)
bn_act_bwd_apply_kernel(const __nv_bfloat16* __restrict__ y, const __nv_bfloat16* __restrict__ g,
                        const float* __restrict__ ss, const float* __restrict__ sums, float inv_n, int B, int T, int C,
                        int ld, int act, float a, float bb, const float* __restrict__ xlen,
                        __nv_bfloat16* __restrict__ dy) {
    const int vec_per_row = ld / 8;
    const size_t n = (size_t)B * T * vec_per_row;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int cv = (int)(i % vec_per_row);
        const size_t row = i / vec_per_row;
        const int t = (int)(row % T), b = (int)(row / T);
        const int c0 = cv * 8;
        bool keep = true;
        if (xlen != nullptr) keep = t < frac_len(__ldg(xlen + b), T);
        const uint4 yv4 = *reinterpret_cast<const uint4*>(y + row * ld + c0);
        const uint4 gv4 = *reinterpret_cast<const uint4*>(g + row * ld + c0);
        const uint32_t yw[4] = {yv4.x, yv4.y, yv4.z, yv4.w};
        const uint32_t gw[4] = {gv4.x, gv4.y, gv4.z, gv4.w};
        uint32_t ow[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 yf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yw[j]));
            const float2 gf = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gw[j]));
            float o[2] = {0.f, 0.f};
            const float yy[2] = {yf.x, yf.y}, gg[2] = {gf.x, gf.y};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = c0 + 2 * j + e;
                if (c < C) {
                    const float sc = __ldg(ss + c), sh = __ldg(ss + C + c), mean = __ldg(ss + 2 * C + c), istd = __ldg(ss + 3 * C + c);
                    const float dz = keep ? gg[e] * act_grad(fmaf(yy[e], sc, sh), act, a, bb) : 0.f;
                    const float xhat = (yy[e] - mean) * istd;
                    o[e] = sc * (dz - __ldg(sums + c) * inv_n - xhat * __ldg(sums + C + c) * inv_n);
                }
            }
            ow[j] = pack_bf16x2(o[0], o[1]);
        }
        *reinterpret_cast<uint4*>(dy + row * ld + c0) = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
}

// ---------------------------------------------------------------------------------------
// weight packing: fp32 [Co, Ci, K] -> bf16 tap-major [K, Co, ci_ld] (forward operand) and/or
// bf16 [K, Ci, co_ld] with flipped taps (dgrad operand); inverse for the gradient.
// ---------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, int Co, int Ci, int K, __nv_bfloat16* __restrict__ fwd,
                                   int ci_ld, __nv_bfloat16* __restrict__ dgr, int co_ld) {
    const size_t n = (size_t)Co * Ci * K;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        const int ci = (int)((i / K) % Ci);
        const int co = (int)(i / ((size_t)K * Ci));
        const __nv_bfloat16 v = __float2bfloat16_rn(w[i]);
        if (fwd) fwd[((size_t)k * Co + co) * ci_ld + ci] = v;
        if (dgr) dgr[((size_t)(K - 1 - k) * Ci + ci) * co_ld + co] = v;
    }
}
// packed gradient fp32 [K, M, ld] -> fp32 [Co, Ci, K]; transposed = packed is [K, Ci, Co]
__global__ void unpack_wgrad_kernel(const float* __restrict__ packed, int K, int Co, int Ci, int ld, int transposed,
                                    float* __restrict__ grad, int accumulate) {
    const size_t n = (size_t)Co * Ci * K;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        const int ci = (int)((i / K) % Ci);
        const int co = (int)(i / ((size_t)K * Ci));
        const float v = transposed ? packed[((size_t)k * Ci + ci) * ld + co] : packed[((size_t)k * Co + co) * ld + ci];
        grad[i] = accumulate ? grad[i] + v : v;
    }
}

// fp32 [B, C, T] -> bf16 channels-last [B, T, ld] (zero padded channels) + per-class sums over (b, t)
__global__ void __launch_bounds__(256)
bct_to_btc_kernel(const float* __restrict__ x, int B, int C, int T, int ld, __nv_bfloat16* __restrict__ out,
                  float* __restrict__ csum) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int c = c0 + i, t = t0 + tx;
        float v = 0.f;
        if (c < C && t < T) v = x[((size_t)b * C + c) * T + t];
        tile[i][tx] = v;
        if (csum != nullptr) {
            float s = warp_sum(v);
            if (tx == 0 && c < C) atomicAdd(csum + c, s);
        }
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int t = t0 + i, c = c0 + tx;
        if (t < T && c < ld) out[((size_t)b * T + t) * ld + c] = __float2bfloat16_rn(tile[tx][i]);
    }
}

}  // namespace cab

using namespace cab;

#define GRID_1D(n, per) (int)(((n) / (per) + 255) / 256 > 148 * 8 ? 148 * 8 : (((n) / (per) + 255) / 256 < 1 ? 1 : ((n) / (per) + 255) / 256))

extern "C" int cab_bn_batch_stats(const void* y, int B, int T, int C, int ld, const float* gamma, const float* beta, float eps,
                                  float momentum, float* running_mean, float* running_var, float* ws_sums /*[2][C]*/,
                                  float* out_ss /*[4][C]*/, cab_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CAB_CHECK_ARG(y && ws_sums && out_ss, "null pointer argument");
    CAB_CHECK_ARG(ld % 8 == 0 && ld >= C && C % 2 == 0, "bad channel layout C=%d ld=%d", C, ld);
    CAB_CHECK_CUDA(cudaMemsetAsync(ws_sums, 0, sizeof(float) * 2 * C, stream));
    const int R = B * T;
    dim3 grid((C + 127) / 128, (R + kRowsPerBlock - 1) / kRowsPerBlock), block(64, 4);
    colstats_kernel<<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(y), R, C, ld, ws_sums);
    CAB_CHECK_LAUNCH();
    bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(ws_sums, C, (float)R, gamma, beta, eps, momentum, running_mean, running_var, out_ss);
    CAB_CHECK_LAUNCH();
    g_launch_count.fetch_add(2, std::memory_order_relaxed);
    return 0;
}

extern "C" int cab_bn_act_mask_fwd(const void* y, const float* ss, int B, int T, int C, int ld, int act, float act_a, float act_b,
                                   const float* xlen_frac, void* out, cab_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CAB_CHECK_ARG(y && ss && out, "null pointer argument");
    CAB_CHECK_ARG(ld % 8 == 0 && ld >= C, "bad channel layout");
    const size_t n = (size_t)B * T * (ld / 8);
    bn_act_mask_fwd_kernel<<<GRID_1D(n, 1), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(y), ss, B, T, C, ld, act, act_a, act_b, xlen_frac, static_cast<__nv_bfloat16*>(out));
    CAB_CHECK_LAUNCH();
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

extern "C" int cab_bn_act_mask_bwd(const void* y, const void* grad_out, const float* ss, int B, int T, int C, int ld, int act,
                                   float act_a, float act_b, const float* xlen_frac, float* sums /*[2][C]: dbeta, dgamma*/,
                                   void* grad_y, cab_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CAB_CHECK_ARG(y && grad_out && ss && sums && grad_y, "null pointer argument");
    CAB_CHECK_ARG(ld % 8 == 0 && ld >= C && C % 2 == 0, "bad channel layout");
    CAB_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * C, stream));
    const int R = B * T;
    dim3 grid((C + 127) / 128, (R + kRowsPerBlock - 1) / kRowsPerBlock), block(64, 4);
    bn_act_bwd_reduce_kernel<<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(grad_out), ss, B, T, C, ld, act, act_a, act_b, xlen_frac, sums);
    CAB_CHECK_LAUNCH();
    const size_t n = (size_t)B * T * (ld / 8);
    bn_act_bwd_apply_kernel<<<GRID_1D(n, 1), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(y), static_cast<const __nv_bfloat16*>(grad_out), ss, sums, 1.f / (float)R, B, T, C, ld, act, act_a, act_b, xlen_frac, static_cast<__nv_bfloat16*>(grad_y));
    CAB_CHECK_LAUNCH();
    g_launch_count.fetch_add(2, std::memory_order_relaxed);
    return 0;
}

extern "C" int cab_pack_weight(const float* w, int Co, int Ci, int K, void* fwd, int ci_ld, void* dgrad, int co_ld, cab_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CAB_CHECK_ARG(w && (fwd || dgrad), "null pointer argument");
    CAB_CHECK_ARG((!fwd || ci_ld >= Ci) && (!dgrad || co_ld >= Co), "bad pitch");
    const size_t n = (size_t)Co * Ci * K;
    pack_weight_kernel<<<GRID_1D(n, 1), 256, 0, stream>>>(w, Co, Ci, K, static_cast<__nv_bfloat16*>(fwd), ci_ld, static_cast<__nv_bfloat16*>(dgrad), co_ld);
    CAB_CHECK_LAUNCH();
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

extern "C" int cab_unpack_wgrad(const float* packed, int K, int Co, int Ci, int ld, int transposed, float* grad, int accumulate, cab_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CAB_CHECK_ARG(packed && grad, "null pointer argument");
    const size_t n = (size_t)Co * Ci * K;
    unpack_wgrad_kernel<<<GRID_1D(n, 1), 256, 0, stream>>>(packed, K, Co, Ci, ld, transposed, grad, accumulate);
    CAB_CHECK_LAUNCH();
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return 0;
}

extern "C" int cab_bct_to_btc(const float* x, int B, int C, int T, int ld, void* out, float* class_sums, cab_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    CAB_CHECK_ARG(x && out, "null pointer argument");
    CAB_CHECK_ARG(ld >= C, "bad pitch");
    if (class_sums) CAB_CHECK_CUDA(cudaMemsetAsync(class_sums, 0, sizeof(float) * C, stream));
    dim3 grid((T + 31) / 32, (ld + 31) / 32, B);
    bct_to_btc_kernel<<<grid, 256, 0, stream>>>(x, B, C, T, ld, static_cast<__nv_bfloat16*>(out), class_sums);
    CAB_CHECK_LAUNCH();
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return 0;
}
