// C ABI of libeav_b200.so for the EEGNet path: argument validation, parameter/workspace
// layout and the launch sequences of forward and backward (include/eav_b200.h).
#include <stdarg.h>
#include <string.h>

#include "eegnet_kernels.cuh"

namespace eav {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int current_device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return dev < 64 ? dev : 63;
}

int device_sm_count() {
    static PerDevice<int> sms(0);
    int &n = sms.here();
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

int make_dims(const eav_eegnet_cfg *c, NetDims *d) {
    EAV_REQUIRE(c != nullptr, EAV_ERR_BAD_ARG, "cfg is NULL");
    EAV_REQUIRE(c->n_models > 0 && c->batch > 0, EAV_ERR_BAD_ARG, "n_models=%d batch=%d must be positive", c->n_models, c->batch);
    EAV_REQUIRE(c->chans > 0 && c->samples > 0 && c->kern_len > 0 && c->F1 > 0 && c->D > 0 && c->F2 > 0 &&
                    c->kern_len2 > 0 && c->pool1 > 0 && c->pool2 > 0 && c->n_classes > 0,
                EAV_ERR_BAD_ARG, "all model dimensions must be positive");
    EAV_REQUIRE(c->variant == EAV_VARIANT_TOR || c->variant == EAV_VARIANT_CNN, EAV_ERR_BAD_ARG, "unknown variant %d", c->variant);
    EAV_REQUIRE(c->dropout_mode >= 0 && c->dropout_mode <= 3, EAV_ERR_BAD_ARG, "unknown dropout_mode %d", c->dropout_mode);
    EAV_REQUIRE(c->dropout_p >= 0.f && c->dropout_p <= 1.f, EAV_ERR_BAD_ARG, "dropout_p=%f out of [0,1]", c->dropout_p);
    EAV_REQUIRE((int64_t)c->n_models * c->batch < (1ll << 24), EAV_ERR_UNSUPPORTED, "n_models*batch too large");
    d->M = c->n_models; d->B = c->batch; d->N = d->M * d->B;
    d->C = c->chans; d->T = c->samples; d->K1 = c->kern_len; d->F1 = c->F1; d->D = c->D; d->G = c->F1 * c->D;
    d->F2 = c->F2; d->K2 = c->kern_len2; d->P1 = c->pool1; d->P2 = c->pool2;
    d->T4 = d->T / d->P1; d->T32 = d->T4 / d->P2; d->NC = c->n_classes; d->FEAT = d->F2 * d->T32;
    EAV_REQUIRE(d->T32 > 0, EAV_ERR_BAD_ARG, "Samples=%d too short for the two pooling stages", d->T);
    d->variant = c->variant; d->bn_train = c->bn_train != 0; d->dropout_mode = c->dropout_mode;
    d->pad1l = (d->K1 - 1) / 2; d->pad2l = (d->K2 - 1) / 2;   // torch padding='same': left = total//2
    d->p_drop = c->dropout_p; d->eps = c->bn_eps; d->momentum = c->bn_momentum; d->norm_rate = c->norm_rate;
    d->seed = c->seed; d->step = c->step;
    d->step_ptr = reinterpret_cast<const unsigned long long *>(c->step_device_ptr);
    d->dp_world = c->dp_world > 1 ? c->dp_world : 1;
    if (d->p_drop == 0.f) d->dropout_mode = EAV_DROPOUT_NONE;
    int64_t o = 0;
    d->oW1 = o; o += (int64_t)d->F1 * d->K1;
    d->og1 = o; o += d->F1;
    d->ob1 = o; o += d->F1;
    d->oW2 = o; o += (int64_t)d->G * d->C;
    d->og2 = o; o += d->G;
    d->ob2 = o; o += d->G;
    if (d->variant == EAV_VARIANT_TOR) {
        d->oW3 = o; o += (int64_t)d->F2 * d->G * d->K2;
        d->oW3p = -1;
    } else {
        d->oW3 = o; o += (int64_t)d->G * d->K2;          // depthwise temporal (G,1,1,K2)
        d->oW3p = o; o += (int64_t)d->F2 * d->G;         // pointwise (F2,G,1,1)
    }
    d->og3 = o; o += d->F2;
    d->ob3 = o; o += d->F2;
    d->oWd = o; o += (int64_t)d->NC * d->FEAT;
    d->obd = o; o += d->NC;
    d->n_params = o;
    d->pstride = c->param_stride;
    d->bnstride = c->bn_stride;
    EAV_REQUIRE(d->pstride >= d->n_params, EAV_ERR_BAD_ARG, "param_stride=%lld < n_params=%lld", (long long)d->pstride, (long long)d->n_params);
    int64_t b = 0;
    d->orm1 = b; b += d->F1; d->orv1 = b; b += d->F1;
    d->orm2 = b; b += d->G;  d->orv2 = b; b += d->G;
    d->orm3 = b; b += d->F2; d->orv3 = b; b += d->F2;
    EAV_REQUIRE(d->bnstride >= b, EAV_ERR_BAD_ARG, "bn_stride=%lld < %lld", (long long)d->bnstride, (long long)b);
    return 0;
}

static size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }

WsLayout make_ws_layout(const NetDims &d) {
    WsLayout w;
    size_t o = 0;
    auto take = [&](size_t n_floats) { size_t at = o; o = align_up(o + n_floats * sizeof(float), 256); return at; };
    const size_t N = d.N;
    w.y1 = take(N * d.F1 * d.C * d.T);
    w.y2 = take(N * d.G * d.T);
    w.d1 = take(N * d.G * d.T4);
    w.y3d = take(d.variant == EAV_VARIANT_CNN ? N * d.G * d.T4 : 0);
    w.y3 = take(N * d.F2 * d.T4);
    w.feat = take(N * d.FEAT);
    w.probs = take(N * d.NC);
    w.dz = take(N * d.NC);
    w.bnf1 = take((size_t)d.M * d.F1 * 4); w.bnf2 = take((size_t)d.M * d.G * 4); w.bnf3 = take((size_t)d.M * d.F2 * 4);
    w.bnb1 = take((size_t)d.M * d.F1 * 4); w.bnb2 = take((size_t)d.M * d.G * 4); w.bnb3 = take((size_t)d.M * d.F2 * 4);
    {   // float64 BN sums, 6 slots (fwd 1..3, bwd 1..3), each [M][max channels][2]
        int mc = d.F1 > d.G ? d.F1 : d.G; if (d.F2 > mc) mc = d.F2;
        w.bnsum_slot = align_up((size_t)d.M * mc * 2 * sizeof(double), 256);
        w.bnsum = take(6 * w.bnsum_slot / sizeof(float));
    }
    // BN partial sums (largest user)
    size_t pa = 0;
    pa = max_sz(pa, N * tconv_fwd_rows_per_sample(d) * 2 * d.F1);         // tconv_fwd
    pa = max_sz(pa, N * cdiv(d.T, 128) * 2 * d.G);                        // dw_fwd
    pa = max_sz(pa, (size_t)d.M * d.B * cdiv(d.T4, 128) * 2 * d.F2);       // sepconv_fwd (<= one row per sample and tile)
    if (d.variant == EAV_VARIANT_TOR) pa = max_sz(pa, (size_t)d.M * sepconv_fwd_rows_per_model(d) * 2 * d.F2);
    pa = max_sz(pa, N * 2 * d.F2);                                        // pw_fwd / tail_bwd
    pa = max_sz(pa, N * 2 * d.G);                                         // pool1_bwd
    pa = max_sz(pa, N * 2 * d.F1);                                        // dw_bwd
    pa = max_sz(pa, (size_t)d.M * 128 * 2 * (d.F1 + d.G));                // fused block-1 backward: [M][S <= 128][F1 | G][2]
    w.part = take(pa);
    // weight-gradient partials, one region per layer (the block-2 kernels run on a forked stream)
    w.partw = take((size_t)d.M * tconv_dw_ctas_per_model(d) * d.F1 * d.K1);        // tconv_bwd_dw
    w.partw2 = take(max_sz(N * d.G * d.C, tconv_bwd_fused_shape_ok(d) ? tconv_bwd_fused_partw2_floats(d) : 0));   // dw_bwd / fused block-1 backward
    if (d.variant == EAV_VARIANT_TOR) w.partw3 = take((size_t)d.M * sepconv_dw_splits(d) * d.F2 * d.G * 16);
    else w.partw3 = take(max_sz(N * d.F2 * d.G, N * d.G * d.K2));
    w.dz3 = take(N * d.F2 * d.T4);
    w.dd1 = take(N * d.G * d.T4);
    w.dy3d = take(d.variant == EAV_VARIANT_CNN ? N * d.G * d.T4 : 0);
    w.dz2 = take(N * d.G * d.T);
    w.dz1 = take(N * d.F1 * d.C * d.T);
    w.tcw = take(tconv_fwd_tc_scratch_floats(d));
    w.tcw2 = take(sepconv_tc_scratch_floats(d));
    w.total = o;
    return w;
}

}  // namespace eav

using namespace eav;

extern "C" const char *eav_last_error_string(void) { return g_err; }
extern "C" int eav_abi_version(void) { return EAV_ABI_VERSION; }
extern "C" uint64_t eav_launch_count(void) { return __atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

extern "C" int eav_check_device(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { set_error("no CUDA device: %s", cudaGetErrorString(e)); return (int)e; }
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    EAV_REQUIRE(major == 10, EAV_ERR_UNSUPPORTED, "device is sm_%d%d; libeav_b200 is built for sm_100a only", major, minor);
    return 0;
}

extern "C" int64_t eav_eegnet_param_layout(const eav_eegnet_cfg *cfg, int64_t *offsets) {
    if (cfg == nullptr) { set_error("cfg is NULL"); return EAV_ERR_BAD_ARG; }
    eav_eegnet_cfg c = *cfg;
    c.param_stride = INT32_MAX;   // layout query: strides not checked
    c.bn_stride = INT32_MAX;
    NetDims d;
    int rc = make_dims(&c, &d);
    if (rc) return rc;
    if (offsets) {
        if (d.variant == EAV_VARIANT_TOR) {
            int64_t o[12] = {d.oW1, d.og1, d.ob1, d.oW2, d.og2, d.ob2, d.oW3, d.og3, d.ob3, d.oWd, d.obd, -1};
            memcpy(offsets, o, sizeof(o));
        } else {
            int64_t o[12] = {d.oW1, d.og1, d.ob1, d.oW2, d.og2, d.ob2, d.oW3, d.oW3p, d.og3, d.ob3, d.oWd, d.obd};
            memcpy(offsets, o, sizeof(o));
        }
    }
    return d.n_params;
}

extern "C" size_t eav_eegnet_workspace_bytes(const eav_eegnet_cfg *cfg) {
    NetDims d;
    if (make_dims(cfg, &d)) return 0;
    return make_ws_layout(d).total;
}

extern "C" int eav_eegnet_workspace_offsets(const eav_eegnet_cfg *cfg, size_t *o) {
    NetDims d;
    int rc = make_dims(cfg, &d);
    if (rc) return rc;
    EAV_REQUIRE(o != nullptr, EAV_ERR_BAD_ARG, "workspace_offsets: null pointer");
    const WsLayout w = make_ws_layout(d);
    const size_t v[16] = {w.y1, w.y2, w.d1, w.y3d, w.y3, w.feat, w.probs, w.dz, w.dz3, w.dd1, w.dy3d, w.dz2, w.dz1,
                          w.bnf1, w.bnf2, w.bnf3};
    memcpy(o, v, sizeof(v));
    return 0;
}

#define WS(T, off) reinterpret_cast<T *>(reinterpret_cast<char *>(workspace) + (off))
#define TRY(call) do { int rc__ = (call); if (rc__) return rc__; } while (0)

// ---------------------------------------------------------------------------------
// Stage table.  forward = stages [0, EAV_STAGE_FWD_END), backward = [EAV_STAGE_FWD_END,
// EAV_STAGE_COUNT).  eav_eegnet_run_stage() launches exactly one of them on the buffers
// a previous full pass left in the workspace (per-kernel timing / profiling hook).
// ---------------------------------------------------------------------------------
enum {
    ST_TCONV_FWD = 0, ST_BN1_REDUCE, ST_BN1, ST_DW_FWD, ST_RENORM_W2, ST_BN2_REDUCE, ST_BN2, ST_POOL1_FWD,
    ST_SEPCONV_FWD, ST_BN3_REDUCE, ST_BN3, ST_TAIL_FWD, ST_RENORM_WD,
    ST_TAIL_BWD, ST_DENSE_BWD_W, ST_BN3_BWD_REDUCE, ST_BN3_BWD, ST_BN3_BWD_APPLY, ST_SEPCONV_BWD_DX, ST_SEPCONV_BWD_DW, ST_POOL1_BWD,
    ST_BN2_BWD_REDUCE, ST_BN2_BWD, ST_DW_BWD, ST_BN1_BWD_REDUCE, ST_BN1_BWD, ST_TCONV_BWD_DW, ST_COUNT
};
static const int ST_FWD_END = ST_TAIL_BWD;
static const char *kStageNames[ST_COUNT] = {
    "tconv_fwd", "bn1_reduce", "bn1_finalize", "dw_fwd", "renorm_depthwise", "bn2_reduce", "bn2_finalize",
    "pool1_fwd", "sepconv_fwd", "bn3_reduce", "bn3_finalize", "tail_fwd", "renorm_dense",
    "tail_bwd", "dense_bwd_w", "bn3_bwd_reduce", "bn3_bwd_finalize", "bn3_bwd_apply", "sepconv_bwd_dx", "sepconv_bwd_dw",
    "pool1_bwd", "bn2_bwd_reduce", "bn2_bwd_finalize", "dw_bwd", "bn1_bwd_reduce", "bn1_bwd_finalize",
    "tconv_bwd_dw"};

// which BatchNorm layer (1..3) a *_REDUCE stage belongs to, and whether it is a backward one
static bool reduce_stage(int stage, int *layer, bool *bwd) {
    switch (stage) {
        case ST_BN1_REDUCE: *layer = 1; *bwd = false; return true;
        case ST_BN2_REDUCE: *layer = 2; *bwd = false; return true;
        case ST_BN3_REDUCE: *layer = 3; *bwd = false; return true;
        case ST_BN1_BWD_REDUCE: *layer = 1; *bwd = true; return true;
        case ST_BN2_BWD_REDUCE: *layer = 2; *bwd = true; return true;
        case ST_BN3_BWD_REDUCE: *layer = 3; *bwd = true; return true;
        default: return false;
    }
}

struct StageArgs {
    const float *x; const int32_t *x_index; float *params; float *bn_state; const uint8_t *mask1, *mask2;
    float *out; const float *dout; float *grads; void *workspace;
};

static int run_stage(const NetDims &d, const WsLayout &w, int stage, const StageArgs &a, cudaStream_t st) {
    void *workspace = a.workspace;
    float *part = WS(float, w.part);
    float *pstat = d.bn_train ? part : nullptr;
    float *partw = WS(float, w.partw), *partw2 = WS(float, w.partw2), *partw3 = WS(float, w.partw3);
    const bool tor = d.variant == EAV_VARIANT_TOR;
    // rows of BatchNorm partial sums each producer writes per model
    const int rows1 = d.B * tconv_fwd_rows_per_sample(d);
    const int rows2 = d.B * dw_fwd_tiles(d);
    const int rows3 = tor ? sepconv_fwd_rows_per_model(d) : d.B;
    // data-parallel (dp_world > 1) train-mode BN: statistics go through float64 sums that the
    // caller all-reduces between the *_reduce stage and the *_finalize stage
    const bool dp_bn = d.dp_world > 1 && d.bn_train;
    auto sums = [&](int layer, bool bwd) -> double * {
        return reinterpret_cast<double *>(reinterpret_cast<char *>(workspace) + w.bnsum + ((bwd ? 3 : 0) + layer - 1) * w.bnsum_slot);
    };
    const double W = (double)d.dp_world;
    {
        int layer; bool bwd;
        if (reduce_stage(stage, &layer, &bwd)) {
            if (!dp_bn) return 0;
            const int rows = bwd ? d.B : (layer == 1 ? rows1 : layer == 2 ? rows2 : rows3);
            return launch_bn_reduce(d, layer, part, rows, sums(layer, bwd), st);
        }
    }
    switch (stage) {
        case ST_TCONV_FWD: return launch_tconv_fwd(d, a.x, a.x_index, a.params, WS(float, w.tcw), WS(float, w.y1), pstat, nullptr, st);
        case ST_BN1:
            if (!d.bn_train)     // eval mode: all three layers at once (ST_BN2 / ST_BN3 are then no-ops)
                return launch_bn_eval_finalize_all(d, a.params, a.bn_state, WS(float4, w.bnf1), WS(float4, w.bnf2), WS(float4, w.bnf3), st);
            return launch_bn_finalize(d, 1, part, rows1, W * d.B * d.C * d.T, dp_bn ? sums(1, false) : nullptr, a.params, a.bn_state, WS(float4, w.bnf1), st);
        case ST_DW_FWD:
            if (dw_fwd_fuses_pool(d)) {   // eval mode: BN2's affine is known up front (ST_BN1 wrote it), pool inside dw_fwd
                return launch_dw_fwd(d, WS(float, w.y1), a.params, WS(float4, w.bnf1), WS(float, w.y2), pstat, nullptr,
                                     WS(float4, w.bnf2), WS(float, w.d1), st);
            }
            return launch_dw_fwd(d, WS(float, w.y1), a.params, WS(float4, w.bnf1), WS(float, w.y2), pstat, nullptr, nullptr, nullptr, st);
        case ST_RENORM_W2:   // hook after the layer used W_old (EEGNet_tor.py:33-34): nothing reads W2 again before the
                             // backward pass, so both hooks run as ONE launch at ST_RENORM_WD
            return 0;
        case ST_BN2:
            if (!d.bn_train) return 0;
            return launch_bn_finalize(d, 2, part, rows2, W * d.B * d.T, dp_bn ? sums(2, false) : nullptr, a.params, a.bn_state, WS(float4, w.bnf2), st);
        case ST_POOL1_FWD:
            if (dw_fwd_fuses_pool(d)) return 0;
            return launch_pool1_fwd(d, WS(float, w.y2), WS(float4, w.bnf2), a.mask1, WS(float, w.d1), st);
        case ST_SEPCONV_FWD:
            if (tor) return launch_sepconv_fwd(d, WS(float, w.d1), a.params, WS(float, w.tcw2), WS(float, w.y3), pstat, nullptr, st);
            TRY(launch_dwt_fwd(d, WS(float, w.d1), a.params, WS(float, w.y3d), st));
            return launch_pw_fwd(d, WS(float, w.y3d), a.params, WS(float, w.y3), pstat, nullptr, st);
        case ST_BN3:
            if (!d.bn_train) return 0;
            return launch_bn_finalize(d, 3, part, rows3, W * d.B * d.T4, dp_bn ? sums(3, false) : nullptr, a.params, a.bn_state, WS(float4, w.bnf3), st);
        case ST_TAIL_FWD:
            return launch_tail_fwd(d, WS(float, w.y3), WS(float4, w.bnf3), a.mask2, a.params, WS(float, w.feat), a.out, WS(float, w.probs), st);
        case ST_RENORM_WD:   // hooks on depthwiseConv and dense (EEGNet_tor.py:33-34,47-48)
            if (tor && d.norm_rate > 0.f)
                return launch_renorm_two(a.params + d.oW2, d.G, d.C, a.params + d.oWd, d.NC, d.FEAT, d.M, d.pstride, d.norm_rate, st);
            return 0;
        case ST_TAIL_BWD:
            return launch_tail_bwd(d, a.dout, WS(float, w.probs), a.params, WS(float, w.y3), WS(float4, w.bnf3), a.mask2,
                                   WS(float, w.dz), WS(float, w.dz3), part, st);
        case ST_DENSE_BWD_W: return launch_dense_bwd_w(d, WS(float, w.feat), WS(float, w.dz), a.grads, st);
        case ST_BN3_BWD:
            return launch_bn_bwd_finalize(d, 3, part, d.B, W * d.B * d.T4, dp_bn ? sums(3, true) : nullptr, a.params, WS(float4, w.bnf3), WS(float4, w.bnb3), a.grads, st);
        case ST_BN3_BWD_APPLY:   // dz3 -> dy3 in place (variant 0; variant 1's pointwise kernel applies it itself)
            if (tail_bwd_folds_bn3(d)) return 0;    // eval mode: tail_bwd already wrote dy3 = k * dz3
            if (tor) return launch_bn_bwd_apply(d, WS(float, w.dz3), WS(float, w.y3), WS(float4, w.bnf3), WS(float4, w.bnb3), st);
            return 0;
        case ST_SEPCONV_BWD_DX:
            if (tor)
                return launch_sepconv_bwd_dx(d, WS(float, w.dz3), WS(float, w.y3), WS(float4, w.bnf3), WS(float4, w.bnb3),
                                             a.params, WS(float, w.tcw2), WS(float, w.dd1), st);
            return launch_pw_bwd(d, WS(float, w.dz3), WS(float, w.y3), WS(float4, w.bnf3), WS(float4, w.bnb3), WS(float, w.y3d),
                                 a.params, WS(float, w.dy3d), partw3, a.grads, st);
        case ST_SEPCONV_BWD_DW:
            if (tor)
                return launch_sepconv_bwd_dw(d, WS(float, w.dz3), WS(float, w.y3), WS(float4, w.bnf3), WS(float4, w.bnb3),
                                             WS(float, w.d1), partw3, a.grads, st);
            return launch_dwt_bwd(d, WS(float, w.dy3d), WS(float, w.d1), a.params, WS(float, w.dd1), partw3, a.grads, st);
        case ST_POOL1_BWD:
            return launch_pool1_bwd(d, WS(float, w.dd1), WS(float, w.y2), WS(float4, w.bnf2), a.mask1, WS(float, w.dz2), part, st);
        case ST_BN2_BWD:
            return launch_bn_bwd_finalize(d, 2, part, d.B, W * d.B * d.T, dp_bn ? sums(2, true) : nullptr, a.params, WS(float4, w.bnf2), WS(float4, w.bnb2), a.grads, st);
        case ST_DW_BWD:
            if (tconv_bwd_fused_ok(d)) return 0;     // eval mode: regenerated inside the tconv_bwd_dw stage
            return launch_dw_bwd(d, WS(float, w.dz2), WS(float, w.y2), WS(float4, w.bnf2), WS(float4, w.bnb2), WS(float, w.y1),
                                 WS(float4, w.bnf1), a.params, WS(float, w.dz1), partw2, part, a.grads, st);
        case ST_BN1_BWD:
            if (tconv_bwd_fused_ok(d)) return 0;     // its sums come out of the fused kernel: finalized in that stage
            return launch_bn_bwd_finalize(d, 1, part, d.B, W * d.B * d.C * d.T, dp_bn ? sums(1, true) : nullptr, a.params, WS(float4, w.bnf1), WS(float4, w.bnb1), a.grads, st);
        case ST_TCONV_BWD_DW:
            if (tconv_bwd_fused_ok(d)) {
                // eval-mode BN: ONE pass over y1 produces dW1 (tcgen05), dW2 and the BatchNorm-1 gradient sums
                const int S = tconv_bwd_dw_tc_splits(d);
                TRY(launch_tconv_bwd_fused_tc(d, a.x, a.x_index, WS(float, w.dz2), WS(float, w.y1), WS(float4, w.bnf1),
                                              WS(float4, w.bnf2), a.params, partw, partw2, part, a.grads, S, st));
                return launch_block1_bwd_finalize(d, partw2, partw, part, S, a.params, WS(float4, w.bnf1), WS(float4, w.bnf2),
                                                  WS(float4, w.bnb1), a.grads, st);
            }
            return launch_tconv_bwd_dw(d, a.x, a.x_index, WS(float, w.dz1), WS(float, w.y1), WS(float4, w.bnf1), WS(float4, w.bnb1),
                                       partw, a.grads, st);
        default: break;
    }
    set_error("run_stage: unknown stage %d", stage);
    return EAV_ERR_BAD_ARG;
}

// One helper stream + two events per device, created on first use and kept for the process.
struct SideStream { cudaStream_t stream; cudaEvent_t fork, join; };
static SideStream *side_stream() {
    static SideStream pool[64];
    static bool made[64] = {false};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!made[dev]) {
        if (cudaStreamCreateWithFlags(&pool[dev].stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        cudaEventCreateWithFlags(&pool[dev].fork, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&pool[dev].join, cudaEventDisableTiming);
        made[dev] = true;
    }
    return &pool[dev];
}

static int check_common(const NetDims &d, const WsLayout &w, const StageArgs &a, size_t workspace_bytes, const char *who) {
    EAV_REQUIRE(a.x && a.params && a.workspace, EAV_ERR_BAD_ARG, "%s: null pointer", who);
    EAV_REQUIRE(workspace_bytes >= w.total, EAV_ERR_WORKSPACE, "%s: workspace %zu < required %zu", who, workspace_bytes, w.total);
    EAV_REQUIRE(d.dropout_mode != EAV_DROPOUT_MASK || (a.mask1 && a.mask2), EAV_ERR_BAD_ARG, "%s: dropout masks required", who);
    return 0;
}

extern "C" int eav_eegnet_stage_count(void) { return ST_COUNT; }
extern "C" int eav_eegnet_stage_forward_end(void) { return ST_FWD_END; }
extern "C" const char *eav_eegnet_stage_name(int stage) { return (stage >= 0 && stage < ST_COUNT) ? kStageNames[stage] : ""; }

extern "C" int eav_eegnet_stage_allreduce(const eav_eegnet_cfg *cfg, int stage, size_t *offset_bytes, size_t *n_doubles) {
    NetDims d;
    TRY(make_dims(cfg, &d));
    EAV_REQUIRE(offset_bytes && n_doubles, EAV_ERR_BAD_ARG, "stage_allreduce: null pointer");
    *offset_bytes = 0;
    *n_doubles = 0;
    int layer; bool bwd;
    if (d.dp_world > 1 && d.bn_train && reduce_stage(stage, &layer, &bwd)) {
        const WsLayout w = make_ws_layout(d);
        const int ch = layer == 1 ? d.F1 : layer == 2 ? d.G : d.F2;
        *offset_bytes = w.bnsum + ((bwd ? 3 : 0) + layer - 1) * w.bnsum_slot;
        *n_doubles = (size_t)d.M * ch * 2;
    }
    return 0;
}

extern "C" int eav_eegnet_run_stage(const eav_eegnet_cfg *cfg, int stage, const float *x, const int32_t *x_index,
                                    float *params, float *bn_state, const uint8_t *mask1, const uint8_t *mask2,
                                    float *out, const float *dout, float *grads, void *workspace,
                                    size_t workspace_bytes, void *stream) {
    NetDims d;
    TRY(make_dims(cfg, &d));
    const WsLayout w = make_ws_layout(d);
    StageArgs a = {x, x_index, params, bn_state, mask1, mask2, out, dout, grads, workspace};
    TRY(check_common(d, w, a, workspace_bytes, "eegnet_run_stage"));
    EAV_REQUIRE(stage >= 0 && stage < ST_COUNT, EAV_ERR_BAD_ARG, "eegnet_run_stage: stage %d out of range", stage);
    if (stage < ST_FWD_END) EAV_REQUIRE(bn_state && out, EAV_ERR_BAD_ARG, "eegnet_run_stage: forward stage needs bn_state and out");
    else EAV_REQUIRE(dout && grads, EAV_ERR_BAD_ARG, "eegnet_run_stage: backward stage needs dout and grads");
    return run_stage(d, w, stage, a, (cudaStream_t)stream);
}

extern "C" int eav_eegnet_forward(const eav_eegnet_cfg *cfg, const float *x, const int32_t *x_index, float *params,
                                  float *bn_state, const uint8_t *mask1, const uint8_t *mask2, float *out,
                                  void *workspace, size_t workspace_bytes, void *stream) {
    NetDims d;
    TRY(make_dims(cfg, &d));
    const WsLayout w = make_ws_layout(d);
    StageArgs a = {x, x_index, params, bn_state, mask1, mask2, out, nullptr, nullptr, workspace};
    TRY(check_common(d, w, a, workspace_bytes, "eegnet_forward"));
    EAV_REQUIRE(bn_state && out, EAV_ERR_BAD_ARG, "eegnet_forward: null pointer");
    EAV_REQUIRE(!(d.dp_world > 1 && d.bn_train), EAV_ERR_UNSUPPORTED,
                "eegnet_forward: dp_world > 1 with train-mode BN must be driven stage by stage (eav_eegnet_run_stage + eav_eegnet_stage_allreduce)");
    for (int s = 0; s < ST_FWD_END; ++s) TRY(run_stage(d, w, s, a, (cudaStream_t)stream));
    return 0;
}

extern "C" int eav_eegnet_backward(const eav_eegnet_cfg *cfg, const float *x, const int32_t *x_index,
                                   const float *params, const float *dout, const uint8_t *mask1,
                                   const uint8_t *mask2, float *grads, void *workspace, size_t workspace_bytes,
                                   void *stream) {
    NetDims d;
    TRY(make_dims(cfg, &d));
    const WsLayout w = make_ws_layout(d);
    StageArgs a = {x, x_index, const_cast<float *>(params), nullptr, mask1, mask2, nullptr, dout, grads, workspace};
    TRY(check_common(d, w, a, workspace_bytes, "eegnet_backward"));
    EAV_REQUIRE(dout && grads, EAV_ERR_BAD_ARG, "eegnet_backward: null pointer");
    EAV_REQUIRE(!(d.dp_world > 1 && d.bn_train), EAV_ERR_UNSUPPORTED,
                "eegnet_backward: dp_world > 1 with train-mode BN must be driven stage by stage");
    cudaStream_t st = (cudaStream_t)stream;
    // The weight gradients of the dense layer and of the block-2 conv do not feed the rest of
    // the chain: they run on a forked stream (works under stream capture too) and fill the SM
    // time the latency/HBM-bound kernels of the main chain leave idle.  Joined before return.
    SideStream *side = (d.variant == EAV_VARIANT_TOR) ? side_stream() : nullptr;
    for (int s = ST_FWD_END; s < ST_COUNT; ++s) {
        if (side != nullptr && (s == ST_DENSE_BWD_W || s == ST_SEPCONV_BWD_DW)) continue;   // issued on the fork
        TRY(run_stage(d, w, s, a, st));
        if (side != nullptr && (s == ST_TAIL_BWD || s == ST_BN3_BWD_APPLY)) {
            cudaEventRecord(side->fork, st);
            cudaStreamWaitEvent(side->stream, side->fork, 0);
            TRY(run_stage(d, w, s == ST_TAIL_BWD ? ST_DENSE_BWD_W : ST_SEPCONV_BWD_DW, a, side->stream));
        }
    }
    if (side != nullptr) {
        cudaEventRecord(side->join, side->stream);
        cudaStreamWaitEvent(st, side->join, 0);
    }
    return 0;
}

// The two max-norm forward hooks (EEGNet_tor.py:33-34,47-48) on their own: what a forward pass leaves behind in the
// weights.  Used when the validation forward runs on a SNAPSHOT of the parameters (pipelined validation): the reference's
// validate() forward would have renormed the live weights, so the live copy gets the same treatment.
extern "C" int eav_eegnet_apply_hooks(const eav_eegnet_cfg *cfg, float *params, void *stream) {
    NetDims d;
    TRY(make_dims(cfg, &d));
    EAV_REQUIRE(params != nullptr, EAV_ERR_BAD_ARG, "eegnet_apply_hooks: null pointer");
    if (d.variant != EAV_VARIANT_TOR || d.norm_rate <= 0.f) return 0;
    return launch_renorm_two(params + d.oW2, d.G, d.C, params + d.oWd, d.NC, d.FEAT, d.M, d.pstride, d.norm_rate,
                             (cudaStream_t)stream);
}
