// Internal helpers shared by the sm_100a kernels of libeav_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/eav_b200.h"

namespace eav {

void set_error(const char *fmt, ...);
void count_launch();   // bumps the process-wide kernel-launch counter (eav_launch_count)

#define EAV_REQUIRE(cond, code, ...)            \
    do {                                        \
        if (!(cond)) {                          \
            ::eav::set_error(__VA_ARGS__);      \
            return (code);                      \
        }                                       \
    } while (0)

#define EAV_CUDA_LAUNCH_CHECK(name)                                               \
    do {                                                                          \
        ::eav::count_launch();                                                    \
        cudaError_t e__ = cudaGetLastError();                                     \
        if (e__ != cudaSuccess) {                                                 \
            ::eav::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return (int)e__;                                                      \
        }                                                                         \
    } while (0)

// One slot per CUDA device for launcher-side caches.  cudaFuncSetAttribute (opt-in dynamic shared memory) and
// the SM count are properties of a DEVICE, so a process-wide `static bool attr_set` would leave the second GPU
// of a single process ("one process, 8 streams") without its > 48 KB opt-in.
int current_device_slot();     // cudaGetDevice() clamped to [0, 64)
int device_sm_count();         // multiprocessor count of the current device (cached per device)
template <typename T>
struct PerDevice {
    T v[64];
    explicit PerDevice(T init) { for (int i = 0; i < 64; ++i) v[i] = init; }
    T &here() { return v[current_device_slot()]; }
};

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ELU (alpha = 1).  exp(x) is ONE multiply + MUFU.EX2 in its .ftz form (no denormal pre-scale / post-square as in
// __expf; results below 2^-126 flush to zero, harmless for exp(x) - 1 and for a derivative that multiplies a gradient).
__device__ __forceinline__ float exp_fast(float x) {
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 1.4426950408889634f));
    return e;
}
// Forward activation with ABSOLUTE accuracy (~2e-7 near zero): multiply, MUFU.EX2, add, compare, select.  Every ELU of
// the net feeds a convolution or a pooling sum, where the absolute error is what counts; a relative-accuracy expm1
// (libdevice's: ~25 predicated instructions, a Taylor/EX2 hybrid: 14) made the bandwidth-bound kernels issue-bound
// (4 ELUs per 16 loaded bytes in dw_fwd: 0.235 -> 0.193 ms from this change alone).  The backward kernels recompute the
// activation with the same expression, so forward and backward see the same values.
__device__ __forceinline__ float elu_fast(float x) { return x > 0.f ? x : exp_fast(x) - 1.f; }
__device__ __forceinline__ float elu_grad_from_pre(float x) { return x > 0.f ? 1.f : exp_fast(x); }
__device__ __forceinline__ float elu_bwd_act(float x) { return elu_fast(x); }

// ---------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), counter = element index / 4, key = (seed, step).
// Used for on-device dropout (EAV_DROPOUT_PHILOX); forward and backward regenerate the
// same keep decision from (seed, step, layer, element).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}
__device__ __forceinline__ bool philox_keep(uint64_t seed, uint64_t step, uint32_t layer,
                                            uint64_t elem, float p_drop) {
    uint4 c = make_uint4((uint32_t)(elem >> 2), (uint32_t)(elem >> 34), layer, (uint32_t)step);
    uint2 k = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
    uint4 r = philox4x32_10(c, k);
    uint32_t v = (elem & 3) == 0 ? r.x : (elem & 3) == 1 ? r.y : (elem & 3) == 2 ? r.z : r.w;
    // keep with probability 1 - p_drop
    return (float)(v >> 8) * (1.0f / 16777216.0f) >= p_drop;
}

// Keep decisions of the four elements 4*block .. 4*block+3 (bit j = element 4*block + j) from ONE Philox call:
// identical to philox_keep() element by element, at a quarter of the integer work.
__device__ __forceinline__ uint32_t philox_keep4(uint64_t seed, uint64_t step, uint32_t layer, uint64_t block,
                                                 float p_drop) {
    uint4 c = make_uint4((uint32_t)block, (uint32_t)(block >> 32), layer, (uint32_t)step);
    uint2 k = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(step >> 32));
    uint4 r = philox4x32_10(c, k);
    const float sc = 1.0f / 16777216.0f;
    return ((float)(r.x >> 8) * sc >= p_drop ? 1u : 0u) | ((float)(r.y >> 8) * sc >= p_drop ? 2u : 0u) |
           ((float)(r.z >> 8) * sc >= p_drop ? 4u : 0u) | ((float)(r.w >> 8) * sc >= p_drop ? 8u : 0u);
}

// ---------------------------------------------------------------------------------
// EEGNet derived dimensions, parameter offsets and workspace layout.
// ---------------------------------------------------------------------------------
struct NetDims {
    int M, B, N;            // models, batch per model, samples N = M*B
    int C, T, K1, F1, D, G; // G = F1*D
    int F2, K2, P1, P2, T4, T32, NC, FEAT;
    int variant, bn_train, dropout_mode;
    int dp_world;           // data-parallel replicas of one model (1 = single device)
    int pad1l, pad2l;       // 'same' left paddings
    float p_drop, eps, momentum, norm_rate;
    uint64_t seed, step;
    const unsigned long long *step_ptr;   // optional device-resident step counter
    int64_t pstride, bnstride;
    // parameter offsets (floats) inside one model's slice
    int64_t oW1, og1, ob1, oW2, og2, ob2, oW3, oW3p, og3, ob3, oWd, obd, n_params;
    // bn_state offsets
    int64_t orm1, orv1, orm2, orv2, orm3, orv3;
};

int make_dims(const eav_eegnet_cfg *cfg, NetDims *d);

// Per-BN-layer statistics block, one float4-pair per (model, channel):
//   fwd: {mean, invstd, scale = gamma*invstd, shift = beta - mean*scale}
//   bwd: {k = gamma*invstd, c1 = sum(dz)/cnt (0 in eval), c2 = sum(dz*xhat)/cnt (0 in eval), unused}
struct WsLayout {
    // offsets in BYTES from the workspace base
    size_t y1, y2, d1, y3d, y3, feat, probs, dz;      // saved activations
    size_t bnf1, bnf2, bnf3, bnb1, bnb2, bnb3;         // float4 per (m, ch)
    size_t bnsum;                                      // double [6 slots][M][max ch][2]: BN sums for the dp all-reduce
    size_t bnsum_slot;                                 // bytes per slot
    size_t part;                                       // BatchNorm partial sums
    size_t partw;                                      // weight-gradient partials: temporal conv (dW1)
    size_t partw2, partw3;                             // ... depthwise (dW2), block-2 conv (dW3): separate so the
                                                       // dW3 kernels can run on a forked stream
    size_t dz3, dd1, dy3d, dz2, dz1;                   // backward scratch
    size_t tcw;                                        // packed tensor-core operand of the temporal conv weights
    size_t tcw2;                                       // ... of the block-2 conv weights (forward pack, d(input) pack)
    size_t total;
};
WsLayout make_ws_layout(const NetDims &d);

// split plan shared by launcher and workspace sizing
int tconv_dw_ctas_per_model(const NetDims &d);
int sepconv_dw_splits(const NetDims &d);

}  // namespace eav
