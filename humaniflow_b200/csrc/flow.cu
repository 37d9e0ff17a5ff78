// Ancestor-conditioned SO(3) normalising flow for sm_100a: one launch walks the whole kinematic tree.
//
// Replaces: models/humaniflow_model.py:116-186,286-320;
// models/norm_flows/pyro_conditional_norm_flow.py:21-129; local_diffeo_transformed_distribution.py:72-142;
// transforms/{conditional_spline_coupling,scaled_radial_tanh,so3_exp,to}_transform.py;
// utils/rigid_transform_utils.py:142-314; pyro 1.7.0 SplineCoupling/ConditionalDenseNN/Permute [upstream].
//
// Design: a CTA owns NR rows (samples) and keeps every activation of those rows in shared memory for all
// 23 joints (image-level features, the rotations already drawn for the ancestors, contexts, hidden
// layers).  Dense layers are fp32 register-tiled GEMMs (4 outputs x NR/4 rows per thread, K split across
// the 16 warps and reduced through shared memory); weights (3.4 MB for the whole tree, L2-resident) are
// streamed with 128-bit read-only loads, each element exactly once per CTA.  fp32 everywhere the
// reference is fp32, fp64 for the exp/log maps and the radial-tanh inverse, as in the reference.
#include "common.cuh"
#include "tma.cuh"
#include <vector>
#include <cmath>
#include <cstring>

#define HF_FJ 23          // max body parts
#define HF_FANC 23        // max ancestors per joint
#define HF_NT 512         // threads per CTA
#define HF_NW 16          // warps per CTA

namespace {

constexpr int FEATS = 256, CTX = 64, H1 = 64, H2 = 32, H3 = 32, NRAW = 62, NBINS = 8;
// coupling block layout (floats, k-major matrices [K][O + WPAD]: rows padded by 4 floats so that the four k-slices of a
// dense_layer tile fall into different shared-memory banks)
constexpr int WPAD = 4, CTXP = CTX + WPAD;
constexpr int OFF_W0 = 0, OFF_B0 = 65 * (64 + WPAD), OFF_W1 = OFF_B0 + 64, OFF_B1 = OFF_W1 + 64 * (32 + WPAD),
              OFF_W2 = OFF_B1 + 32, OFF_B2 = OFF_W2 + 32 * (32 + WPAD), OFF_W3 = OFF_B2 + 32, OFF_B3 = OFF_W3 + 32 * (64 + WPAD),
              COUPLING_FLOATS = OFF_B3 + 64;

struct FlowParams {
    const float* pack;
    const float* betaW;     // [nb][FEATS]
    int J, nb, T;
    float radius, base_std;
    int off_ctxW[HF_FJ], off_ctxB[HF_FJ], off_nn[HF_FJ][2];
    int anc_cnt[HF_FJ];
    signed char anc[HF_FJ][HF_FANC];
    // sampling kernel: feature part of every context Linear as one K-major matrix [FEATS][J*CTX]; per joint one
    // contiguous block [ancW (9a x CTX) | ctxB (CTX) | coupling 0 | coupling 1] streamed into shared memory
    const float* wfeat;
    const float* jpack;
    int off_jb[HF_FJ];
};

__device__ __forceinline__ float elu(float x) { return x > 0.f ? x : expm1f(x); }

// ---- register-tiled partial GEMM of one warp: part[32][NR] = sum_{k in [k0,k1)} W[k][ob*32 + :] x X[k][:] ----
// packed fp32 x 2 FMA (Blackwell FFMA2): d = a * b + c on both halves, each half rounded like a scalar fmaf, so results are
// bit-identical to the scalar form at half the issue slots
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n"
        ".reg .b64 ra, rb, rc, rd;\n"
        "mov.b64 ra, {%2, %3};\n"
        "mov.b64 rb, {%4, %5};\n"
        "mov.b64 rc, {%6, %7};\n"
        "fma.rn.f32x2 rd, ra, rb, rc;\n"
        "mov.b64 {%0, %1}, rd;\n"
        "}\n" : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}

template <int NR, bool SMEMW = false, int UNR = 4, bool PART_T = false, typename XRow>
__device__ __forceinline__ void warp_gemm(const float* __restrict__ W, int ldw, int ob, int k0, int k1, XRow xrow,
                                          float* __restrict__ part, int lane) {
    constexpr int SPL = NR / 4, SP2 = SPL / 2;
    const int oi = lane & 7, sq = lane >> 3;
    float2 acc[4][SP2];                    // [output][row pair]
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int s = 0; s < SP2; ++s) acc[i][s] = make_float2(0.f, 0.f);
    const float* wp = W + ob * 32 + oi * 4;
#pragma unroll UNR
    for (int k = k0; k < k1; ++k) {
        const float4 w = SMEMW ? *reinterpret_cast<const float4*>(wp + (size_t)k * ldw)
                               : __ldg(reinterpret_cast<const float4*>(wp + (size_t)k * ldw));
        const float2* xp = reinterpret_cast<const float2*>(xrow(k) + sq * SPL);
        const float2 w0 = make_float2(w.x, w.x), w1 = make_float2(w.y, w.y), w2 = make_float2(w.z, w.z), w3 = make_float2(w.w, w.w);
#pragma unroll
        for (int q = 0; q < SP2; ++q) {
            const float2 x = xp[q];
            acc[0][q] = fma2(w0, x, acc[0][q]);
            acc[1][q] = fma2(w1, x, acc[1][q]);
            acc[2][q] = fma2(w2, x, acc[2][q]);
            acc[3][q] = fma2(w3, x, acc[3][q]);
        }
    }
    if (PART_T) {   // partial tile transposed [row][32 outputs]: one 128-bit store per row, bank-conflict free
#pragma unroll
        for (int q = 0; q < SP2; ++q) {
            *reinterpret_cast<float4*>(part + (sq * SPL + 2 * q) * 32 + oi * 4) = make_float4(acc[0][q].x, acc[1][q].x, acc[2][q].x, acc[3][q].x);
            *reinterpret_cast<float4*>(part + (sq * SPL + 2 * q + 1) * 32 + oi * 4) = make_float4(acc[0][q].y, acc[1][q].y, acc[2][q].y, acc[3][q].y);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float2* pp = reinterpret_cast<float2*>(part + (oi * 4 + i) * NR + sq * SPL);
#pragma unroll
            for (int q = 0; q < SP2; ++q) pp[q] = acc[i][q];
        }
    }
}

// A dense layer over the CTA (the layers of the chain are tiny: 32-64 outputs x NR <= 24 rows x K <= 65): bias + activation
// into dst[O][NR].  ACT: 0 none, 1 ELU, 2 ReLU.  Ends with ONE __syncthreads().
// Mapping: a tile = 4 outputs x 4 rows (16 accumulators); the four lanes of a tile take every fourth k (packed fp32x2 FMAs,
// one 128-bit load of 4 weights + one of 4 row values per k) and combine their partial sums with two xor-shuffles; lane i of
// the tile then finishes output i of the tile.  No shared-memory round trip, no second barrier: the earlier form (K split
// over the 16 warps, partial tiles through shared memory, two barriers) measured ~1.9 k cycles per layer against ~0.4 k of
// FMA work.  W rows are padded by 4 floats (LDW = O + 4) so the four k-slices of a tile hit different banks.
template <int NR, int OT, int ACT, bool SMEMW = false, typename XRow>
__device__ __forceinline__ void dense_layer(const float* __restrict__ W, const float* __restrict__ bias, int K,
                                            XRow xrow, float* scratch, float* dst, const float* add = nullptr) {
    constexpr int O = OT * 32, LDW = O + 4, RG = NR / 4, NTILE = (O / 4) * RG;
    static_assert(NR % 4 == 0 && NTILE * 4 <= HF_NT, "dense_layer tile mapping");
    (void)scratch;
    const int tile = threadIdx.x >> 2, ks = threadIdx.x & 3;
    if (tile < NTILE) {                     // whole warps (NTILE * 4 is a multiple of 32 for NR in {8, 16, 24})
        const int og = tile / RG, rg = tile - og * RG;
        float2 acc[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
        const float* wp = W + og * 4;
#pragma unroll 4
        for (int k = ks; k < K; k += 4) {
            const float4 w = SMEMW ? *reinterpret_cast<const float4*>(wp + (size_t)k * LDW)
                                   : __ldg(reinterpret_cast<const float4*>(wp + (size_t)k * LDW));
            const float4 x = *reinterpret_cast<const float4*>(xrow(k) + rg * 4);
            const float2 xa = make_float2(x.x, x.y), xb = make_float2(x.z, x.w);
            const float2 w0 = make_float2(w.x, w.x), w1 = make_float2(w.y, w.y), w2 = make_float2(w.z, w.z), w3 = make_float2(w.w, w.w);
            acc[0][0] = fma2(w0, xa, acc[0][0]); acc[0][1] = fma2(w0, xb, acc[0][1]);
            acc[1][0] = fma2(w1, xa, acc[1][0]); acc[1][1] = fma2(w1, xb, acc[1][1]);
            acc[2][0] = fma2(w2, xa, acc[2][0]); acc[2][1] = fma2(w2, xb, acc[2][1]);
            acc[3][0] = fma2(w3, xa, acc[3][0]); acc[3][1] = fma2(w3, xb, acc[3][1]);
        }
        // sum the four k-slices (fixed order: (0+1) + (2+3) in every lane), then lane ks finishes output ks of the tile
        float r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float vx = acc[i][h].x, vy = acc[i][h].y;
                vx += __shfl_xor_sync(0xffffffffu, vx, 1); vy += __shfl_xor_sync(0xffffffffu, vy, 1);
                vx += __shfl_xor_sync(0xffffffffu, vx, 2); vy += __shfl_xor_sync(0xffffffffu, vy, 2);
                if (i == ks) { r[2 * h] = vx; r[2 * h + 1] = vy; }
            }
        }
        const int o = og * 4 + ks;
        const float b = SMEMW ? bias[o] : __ldg(bias + o);
        float* dp = dst + o * NR + rg * 4;
        float4 a4 = add ? *reinterpret_cast<const float4*>(add + o * NR + rg * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        float v[4] = {r[0] + b + a4.x, r[1] + b + a4.y, r[2] + b + a4.z, r[3] + b + a4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (ACT == 1) v[q] = elu(v[q]);
            if (ACT == 2) v[q] = fmaxf(v[q], 0.f);
        }
        *reinterpret_cast<float4*>(dp) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __syncthreads();
}

// ---- rational-linear spline (pyro _monotonic_rational_spline, order='linear') for one scalar ----
struct SplineBin {
    float in_w, in_cw, in_ch, in_h, d0, d1, lam;
};

// Cooperative knot computation for all rows of the CTA: 8 lanes per (row, d, kind) group evaluate the softmax
// over the 8 bins and the cumulative knot positions with warp shuffles (width 8), results to smem
// KN[((row*2 + d)*2 + kind)*9 + i], kind 0 = widths, 1 = heights.  Ends with __syncthreads().
template <int NR>
__device__ __forceinline__ void spline_knots(const float* __restrict__ raw, float bound, float* __restrict__ KN) {
    const float lo = -bound, hi = bound;
    for (int t = threadIdx.x; t < NR * 32; t += HF_NT) {      // NR*4 groups of 8 lanes; HF_NT % 32 == 0 keeps groups in a warp
        const int g = t >> 3, b = t & 7;
        const int row = g >> 2, d = (g >> 1) & 1, kind = g & 1;
        const float x = raw[(kind * 2 * NBINS + d * NBINS + b) * NR + row];
        float m = x;
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1, 8));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2, 8));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4, 8));
        const float e = expf(x - m);
        float sum = e;
        sum += __shfl_xor_sync(0xffffffffu, sum, 1, 8);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2, 8);
        sum += __shfl_xor_sync(0xffffffffu, sum, 4, 8);
        float c = 1e-3f + 0.992f * (e / sum);
        // inclusive scan over the 8 lanes
        float u = __shfl_up_sync(0xffffffffu, c, 1, 8); if (b >= 1) c += u;
        u = __shfl_up_sync(0xffffffffu, c, 2, 8); if (b >= 2) c += u;
        u = __shfl_up_sync(0xffffffffu, c, 4, 8); if (b >= 4) c += u;
        float* k = KN + g * 9;
        k[b + 1] = (b == NBINS - 1) ? hi : (hi - lo) * c + lo;
        if (b == 0) k[0] = lo;
    }
    __syncthreads();
}

// raw: pointer to this row's NN outputs with stride `rs` between consecutive outputs; kn: this (row, d)'s knots
// [widths 9 | heights 9]; d in {0,1}
__device__ __forceinline__ SplineBin spline_select(const float* raw, int rs, int d, float input, const float* kn,
                                                   bool inverse) {
    const float* kw = kn;
    const float* kh = kn + 9;
    const float* ks = inverse ? kh : kw;
    int idx = -1;
#pragma unroll
    for (int b = 0; b <= NBINS; ++b) idx += (input >= ks[b] + 1e-6f) ? 1 : 0;
    idx = max(0, min(idx, NBINS - 1));
    SplineBin r;
    r.in_w = kw[idx + 1] - kw[idx]; r.in_cw = kw[idx]; r.in_ch = kh[idx]; r.in_h = kh[idx + 1] - kh[idx];
    // derivatives: softplus + min, padded with 1 - min at both ends; lambdas: sigmoid, affine
    auto deriv = [&](int i) -> float {   // i in [0, NBINS]
        if (i == 0 || i == NBINS) return 0.999f;
        float x = raw[(4 * NBINS + d * (NBINS - 1) + (i - 1)) * rs];
        float sp = (x > 20.f) ? x : log1pf(expf(x));
        return 1e-3f + sp;
    };
    r.d0 = deriv(idx);
    r.d1 = deriv(idx + 1);
    float l = raw[(4 * NBINS + 2 * (NBINS - 1) + d * NBINS + idx) * rs];
    l = 1.f / (1.f + expf(-l));
    r.lam = 0.95f * l + 0.025f;
    return r;
}

__device__ __forceinline__ float spline_forward(const float* raw, int rs, int d, float x, float bound, const float* kn) {
    if (!(x >= -bound && x <= bound)) return x;
    SplineBin b = spline_select(raw, rs, d, x, kn, false);
    const float delta = b.in_h / b.in_w;
    const float wb = sqrtf(b.d0 / b.d1);
    const float wc = (b.lam * b.d0 + (1.f - b.lam) * wb * b.d1) / delta;
    const float ya = b.in_ch, yb = b.in_h + b.in_ch;
    const float yc = ((1.f - b.lam) * ya + b.lam * wb * yb) / ((1.f - b.lam) + b.lam * wb);
    const float theta = (x - b.in_cw) / b.in_w;
    float num, den;
    if (theta <= b.lam) {
        num = ya * (b.lam - theta) + wc * yc * theta;
        den = (b.lam - theta) + wc * theta;
    } else {
        num = wc * yc * (1.f - theta) + wb * yb * (theta - b.lam);
        den = wc * (1.f - theta) + wb * (theta - b.lam);
    }
    return num / den;
}

constexpr int KNF = 36;
// One knot task of the sampling kernels: 4 lanes per (row, dimension), lane q takes bins 2q and 2q + 1 of BOTH softmax / cumulative
// vectors plus their derivative (softplus) and lambda (sigmoid) entries: 8 * NR tasks (one pass of a 192-thread group at NR = 24;
// the 8-lanes-per-vector form needed two), two shuffle steps per reduction instead of three.
//   KN[(row*2 + d)*KNF + ...] = [widths 9 | heights 9 | derivatives 9 | lambdas 8 | pad]
template <int NR>
__device__ __forceinline__ void knot_task(const float* __restrict__ raw, float bound, float* __restrict__ KN, int t) {
    const float lo = -bound, hi = bound;
    const int g = t >> 2, q = t & 3, row = g >> 1, d = g & 1, b0 = 2 * q;
    const float* r0 = raw + row;
    float x[2][2], e[2][2], c[2][2];          // [vector: widths / heights][bin]
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        x[v][0] = r0[(v * 2 * NBINS + d * NBINS + b0) * NR];
        x[v][1] = r0[(v * 2 * NBINS + d * NBINS + b0 + 1) * NR];
    }
    const float xd0 = r0[(4 * NBINS + d * (NBINS - 1) + b0) * NR], xd1 = r0[(4 * NBINS + d * (NBINS - 1) + min(b0 + 1, NBINS - 2)) * NR];
    const float xl0 = r0[(4 * NBINS + 2 * (NBINS - 1) + d * NBINS + b0) * NR], xl1 = r0[(4 * NBINS + 2 * (NBINS - 1) + d * NBINS + b0 + 1) * NR];
    float* k = KN + g * KNF;
#pragma unroll
    for (int v = 0; v < 2; ++v) {
        float m = fmaxf(x[v][0], x[v][1]);
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1, 4));
        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2, 4));
        e[v][0] = expf(x[v][0] - m); e[v][1] = expf(x[v][1] - m);
        float s = e[v][0] + e[v][1];
        s += __shfl_xor_sync(0xffffffffu, s, 1, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 2, 4);
        c[v][0] = 1e-3f + 0.992f * (e[v][0] / s);
        c[v][1] = 1e-3f + 0.992f * (e[v][1] / s);
        float p = c[v][0] + c[v][1];                      // inclusive scan of the pair sums over the 4 lanes
        float u = __shfl_up_sync(0xffffffffu, p, 1, 4); if (q >= 1) p += u;
        u = __shfl_up_sync(0xffffffffu, p, 2, 4); if (q >= 2) p += u;
        u = __shfl_up_sync(0xffffffffu, p, 1, 4);
        const float excl = q >= 1 ? u : 0.f;
        k[v * 9 + b0 + 1] = (hi - lo) * (excl + c[v][0]) + lo;
        k[v * 9 + b0 + 2] = (q == 3) ? hi : (hi - lo) * p + lo;
        if (q == 0) k[v * 9] = lo;
    }
    const float sp0 = (xd0 > 20.f) ? xd0 : log1pf(expf(xd0)), sp1 = (xd1 > 20.f) ? xd1 : log1pf(expf(xd1));
    k[18 + b0 + 1] = 1e-3f + sp0;
    k[18 + b0 + 2] = (q == 3) ? 0.999f : 1e-3f + sp1;
    if (q == 0) k[18] = 0.999f;
    k[27 + b0] = 0.95f * (1.f / (1.f + expf(-xl0))) + 0.025f;
    k[27 + b0 + 1] = 0.95f * (1.f / (1.f + expf(-xl1))) + 0.025f;
}

template <int NR>
__device__ __forceinline__ void spline_knots_full(const float* __restrict__ raw, float bound, float* __restrict__ KN) {
    for (int t = threadIdx.x; t < NR * 8; t += HF_NT) knot_task<NR>(raw, bound, KN, t);      // whole warps: NR * 8 is a multiple of 32
    __syncthreads();
}

__device__ __forceinline__ float spline_forward_pre(float x, float bound, const float* kn) {
    if (!(x >= -bound && x <= bound)) return x;
    int idx = -1;
#pragma unroll
    for (int b = 0; b <= NBINS; ++b) idx += (x >= kn[b] + 1e-6f) ? 1 : 0;
    idx = max(0, min(idx, NBINS - 1));
    const float in_cw = kn[idx], in_w = kn[idx + 1] - in_cw, in_ch = kn[9 + idx], in_h = kn[9 + idx + 1] - in_ch;
    const float d0 = kn[18 + idx], d1 = kn[18 + idx + 1], lam = kn[27 + idx];
    const float delta = in_h / in_w;
    const float wb = sqrtf(d0 / d1);
    const float wc = (lam * d0 + (1.f - lam) * wb * d1) / delta;
    const float ya = in_ch, yb = in_h + in_ch;
    const float yc = ((1.f - lam) * ya + lam * wb * yb) / ((1.f - lam) + lam * wb);
    const float theta = (x - in_cw) / in_w;
    float num, den;
    if (theta <= lam) {
        num = ya * (lam - theta) + wc * yc * theta;
        den = (lam - theta) + wc * theta;
    } else {
        num = wc * yc * (1.f - theta) + wb * yb * (theta - lam);
        den = wc * (1.f - theta) + wb * (theta - lam);
    }
    return num / den;
}

// inverse spline; *fwd_logdet receives the FORWARD log|dy/dx| evaluated through the inverse formulas
// (pyro Spline._inverse caches -logabsdet of the inverse direction).
__device__ __forceinline__ float spline_inverse(const float* raw, int rs, int d, float y, float bound,
                                                const float* kn, float* fwd_logdet) {
    if (!(y >= -bound && y <= bound)) { *fwd_logdet = 0.f; return y; }
    SplineBin b = spline_select(raw, rs, d, y, kn, true);
    const float delta = b.in_h / b.in_w;
    const float wb = sqrtf(b.d0 / b.d1);
    const float wc = (b.lam * b.d0 + (1.f - b.lam) * wb * b.d1) / delta;
    const float ya = b.in_ch, yb = b.in_h + b.in_ch;
    const float yc = ((1.f - b.lam) * ya + b.lam * wb * yb) / ((1.f - b.lam) + b.lam * wb);
    float num, den, dnum;
    if (y <= yc) {
        num = b.lam * (ya - y);
        den = (wc - 1.f) * y + ya - wc * yc;
        dnum = wc * b.lam * (yc - ya) * b.in_w;
    } else {
        num = (wc - b.lam * wb) * y + b.lam * wb * yb - wc * yc;
        den = (wc - wb) * y + wb * yb - wc * yc;
        dnum = wb * wc * (1.f - b.lam) * (yb - yc) * b.in_w;
    }
    const float theta = num / den;
    *fwd_logdet = -(logf(dnum) - 2.f * logf(fabsf(den)));
    return theta * b.in_w + b.in_cw;
}

// ---- exp / log maps ----
__device__ __forceinline__ void so3_exp_f64(double x, double y, double z, float* R) {
    // utils/rigid_transform_utils.py:182-201
    double th = sqrt(x * x + y * y + z * z);
    double alpha, beta;
    if (th > 1e-10) { double sn, cs; sincos(th, &sn, &cs); alpha = sn / th; beta = (1.0 - cs) / (th * th); }
    else { alpha = 1.0 - 1.0 / 6.0; beta = 0.5 - 1.0 / 24.0; }   // reference substitutes theta=1 in the Taylor terms
    double xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
    R[0] = (float)(1.0 + beta * (-(yy + zz))); R[1] = (float)(-alpha * z + beta * xy); R[2] = (float)(alpha * y + beta * xz);
    R[3] = (float)(alpha * z + beta * xy); R[4] = (float)(1.0 + beta * (-(xx + zz))); R[5] = (float)(-alpha * x + beta * yz);
    R[6] = (float)(-alpha * y + beta * xz); R[7] = (float)(alpha * x + beta * yz); R[8] = (float)(1.0 + beta * (-(xx + yy)));
}

__device__ __forceinline__ void so3_exp_f64d(double x, double y, double z, double* R) {
    double th = sqrt(x * x + y * y + z * z);
    double alpha, beta;
    if (th > 1e-10) { alpha = sin(th) / th; beta = (1.0 - cos(th)) / (th * th); }
    else { alpha = 1.0 - 1.0 / 6.0; beta = 0.5 - 1.0 / 24.0; }
    double xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z;
    R[0] = 1.0 + beta * (-(yy + zz)); R[1] = -alpha * z + beta * xy; R[2] = alpha * y + beta * xz;
    R[3] = alpha * z + beta * xy; R[4] = 1.0 + beta * (-(xx + zz)); R[5] = -alpha * x + beta * yz;
    R[6] = -alpha * y + beta * xz; R[7] = alpha * x + beta * yz; R[8] = 1.0 + beta * (-(xx + yy));
}

__device__ __forceinline__ void rodrigues_f32(float x, float y, float z, float* o) {
    float ex = x + 1e-8f, ey = y + 1e-8f, ez = z + 1e-8f;
    float angle = sqrtf(ex * ex + ey * ey + ez * ez);
    float rx = x / angle, ry = y / angle, rz = z / angle;
    float s = sinf(angle), c1 = 1.f - cosf(angle);
    float xx = rx * rx, yy = ry * ry, zz = rz * rz, xy = rx * ry, xz = rx * rz, yz = ry * rz;
    o[0] = 1.f + c1 * (-(yy + zz)); o[1] = -s * rz + c1 * xy;        o[2] = s * ry + c1 * xz;
    o[3] = s * rz + c1 * xy;        o[4] = 1.f + c1 * (-(xx + zz)); o[5] = -s * rx + c1 * yz;
    o[6] = -s * ry + c1 * xz;       o[7] = s * rx + c1 * yz;        o[8] = 1.f + c1 * (-(xx + yy));
}

// utils/rigid_transform_utils.py:204-279 (returns the axis-angle vector)
__device__ void so3_log_f64(const double* r, double* x) {
    const double PI = 3.14159265358979323846;
    double c = 0.5 * (r[0] + r[4] + r[8] - 1.0);
    c = fmin(fmax(c, -1.0), 1.0);
    double th = acos(c);
    double ratio = th / sin(th);
    if (th < 1e-20) ratio = 1.0 + th * th / 6.0;
    // vee(ratio * 0.5 (R - R^T)) = (-m12, m02, -m01)
    x[0] = -ratio * 0.5 * (r[5] - r[7]);
    x[1] = ratio * 0.5 * (r[2] - r[6]);
    x[2] = -ratio * 0.5 * (r[1] - r[3]);
    if (fabs(PI - th) < 1e-2) {
        double k = th * th / (1.0 - cos(th));
        double q1 = k * (r[0] - 1.0), q2 = k * (r[4] - 1.0), q3 = k * (r[8] - 1.0);
        double a1 = sqrt(fmax(q1 - q2 - q3, 1e-8) / 2.0);
        double a2 = sqrt(fmax(-q1 + q2 - q3, 1e-8) / 2.0);
        double a3 = sqrt(fmax(-q1 - q2 + q3, 1e-8) / 2.0);
        double best = 1e300;
        int bi = 0;
        for (int s = 0; s < 8; ++s) {      // order of itertools.product([0,1], repeat=3) * 2 - 1; first minimum wins
            double sx = (s & 4) ? 1.0 : -1.0, sy = (s & 2) ? 1.0 : -1.0, sz = (s & 1) ? 1.0 : -1.0;
            double E[9];
            so3_exp_f64d(sx * a1, sy * a2, sz * a3, E);
            double d = 0.0;
            for (int e = 0; e < 9; ++e) { double t = r[e] - E[e]; d += t * t; }
            if (d < best) { best = d; bi = s; }
        }
        x[0] = ((bi & 4) ? 1.0 : -1.0) * a1;
        x[1] = ((bi & 2) ? 1.0 : -1.0) * a2;
        x[2] = ((bi & 1) ? 1.0 : -1.0) * a3;
    }
}

// log((2 - 2cos t)/t^2) in fp64 (utils/rigid_transform_utils.py:298-314)
__device__ __forceinline__ double so3_log_abs_det(double x, double y, double z) {
    double n = sqrt(x * x + y * y + z * z);
    double ratio = (n > 1e-10) ? (2.0 - 2.0 * cos(n)) / (n * n) : (1.0 - 1.0 / 12.0);
    return log(ratio);
}

// image-level features of the CTA's rows: Fs[o][s] = ELU(img_base[img][o] + sum_l betaW[l][o] beta[s][l])
template <int NR>
__device__ __forceinline__ void image_feats(const FlowParams& P, const float* __restrict__ img_base,
                                            const float* __restrict__ betas, const int* __restrict__ img_index,
                                            int r0, int R, float* Fs) {
    for (int e = threadIdx.x; e < FEATS * NR; e += HF_NT) {
        const int s = e / FEATS, o = e - s * FEATS;
        const int r = r0 + s;
        float a = 0.f;
        if (r < R) {
            a = __ldg(img_base + (size_t)__ldg(img_index + r) * FEATS + o);
            for (int l = 0; l < P.nb; ++l) a = fmaf(__ldg(P.betaW + l * FEATS + o), __ldg(betas + (size_t)r * P.nb + l), a);
            a = elu(a);
        }
        Fs[o * NR + s] = a;
    }
}

// context input row k of joint j: image features then the ancestors' rotations (row-major 3x3 each)
struct CtxRow {
    const float* Fs; const float* Ps; const signed char* anc; int NR;
    __device__ __forceinline__ const float* operator()(int k) const {
        if (k < FEATS) return Fs + k * NR;
        int q = k - FEATS, a = q / 9;
        return Ps + (anc[a] * 9 + (q - a * 9)) * NR;
    }
};
struct PlainRow {
    const float* X; int NR;
    __device__ __forceinline__ const float* operator()(int k) const { return X + k * NR; }
};

template <int NR>
struct SmemLayout {
    static constexpr int Fs = 0;                          // [FEATS][NR]
    static constexpr int Ps = Fs + FEATS * NR;            // [HF_FJ*9][NR]
    static constexpr int Cs = Ps + HF_FJ * 9 * NR;        // [CTX+1][NR]
    static constexpr int Ha = Cs + (CTX + 1) * NR;        // [64][NR]
    static constexpr int Hb = Ha + 64 * NR;               // [32][NR]
    static constexpr int Hc = Hb + 32 * NR;               // [32][NR]
    static constexpr int Raw = Hc + 32 * NR;              // [64][NR]
    static constexpr int Zs = Raw + 64 * NR;              // [3][NR]
    static constexpr int Scratch = Zs + 4 * NR;           // [HF_NW][32][NR]
    static constexpr int Total = Scratch + HF_NW * 32 * NR;
};

// the 4-layer hypernet of one coupling: Cs (context + x1 in row CTX) -> Raw[64][NR]
template <int NR>
__device__ __forceinline__ void coupling_nn(const float* __restrict__ cw, float* sm) {
    using L = SmemLayout<NR>;
    dense_layer<NR, 2, 2>(cw + OFF_W0, cw + OFF_B0, CTX + 1, PlainRow{sm + L::Cs, NR}, sm + L::Scratch, sm + L::Ha);
    dense_layer<NR, 1, 2>(cw + OFF_W1, cw + OFF_B1, H1, PlainRow{sm + L::Ha, NR}, sm + L::Scratch, sm + L::Hb);
    dense_layer<NR, 1, 2>(cw + OFF_W2, cw + OFF_B2, H2, PlainRow{sm + L::Hb, NR}, sm + L::Scratch, sm + L::Hc);
    dense_layer<NR, 2, 0>(cw + OFF_W3, cw + OFF_B3, H3, PlainRow{sm + L::Hc, NR}, sm + L::Scratch, sm + L::Raw);
}

template <int NR>
__device__ __forceinline__ void coupling_nn_smem(const float* cw, float* sm) {
    using L = SmemLayout<NR>;
    dense_layer<NR, 2, 2, true>(cw + OFF_W0, cw + OFF_B0, CTX + 1, PlainRow{sm + L::Cs, NR}, sm + L::Scratch, sm + L::Ha);
    dense_layer<NR, 1, 2, true>(cw + OFF_W1, cw + OFF_B1, H1, PlainRow{sm + L::Ha, NR}, sm + L::Scratch, sm + L::Hb);
    dense_layer<NR, 1, 2, true>(cw + OFF_W2, cw + OFF_B2, H2, PlainRow{sm + L::Hb, NR}, sm + L::Scratch, sm + L::Hc);
    dense_layer<NR, 2, 0, true>(cw + OFF_W3, cw + OFF_B3, H3, PlainRow{sm + L::Hc, NR}, sm + L::Scratch, sm + L::Raw);
}

// ancestors' rotation rows only (the feature part of the context Linear is precomputed)
struct AncRow {
    const float* Ps; const signed char* anc; int NR;
    __device__ __forceinline__ const float* operator()(int k) const {
        const int a = k / 9;
        return Ps + (anc[a] * 9 + (k - a * 9)) * NR;
    }
};

// same rows through a per-k offset table in shared memory (no division in the k loop)
struct TabRow {
    const float* Ps; const int* off;
    __device__ __forceinline__ const float* operator()(int k) const { return Ps + off[k]; }
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// shared-memory layout of the sampling kernel (floats)
template <int NR>
struct SampleSmem {
    static constexpr int ANC_MAX = 7 * 9 * CTXP + CTX;     // deepest SMPL joint has 7 ancestors; checked on the host
    static constexpr int Wanc = 0;
    static constexpr int Wc0 = Wanc + ANC_MAX;
    static constexpr int Wc1 = Wc0 + COUPLING_FLOATS;
    static constexpr int Us = Wc1 + COUPLING_FLOATS;       // [2][CTX*NR]
    static constexpr int Act = Us + 2 * CTX * NR;           // SmemLayout<NR> minus its Fs (aliased onto Scratch)
    static constexpr int Total = Act + (SmemLayout<NR>::Total - SmemLayout<NR>::Ps);
};

// =====================================  context prologue on the tensor cores  =====================================
// U[r][j*64 + o] = sum_k Wfeat_j[o][k] F[r][k] for all rows and joints is one (R x 256) x (256 x J*64) GEMM with no dependency on
// the chain: 1.2 GMAC at B=32, N=100, 22 % of the sampling kernel when every CTA did its own 24 rows on the CUDA cores (each CTA
// re-read the whole 1.5 MB matrix from L2).  Here it runs as split-tf32 on tcgen05: x = hi + lo with hi = tf32(x) (round to
// nearest) and lo = x - hi (exact in fp32; the MMA reads its upper 10 mantissa bits), D = Fhi Whi + Fhi Wlo + Flo Whi with fp32
// accumulation in TMEM.  The dropped terms (lo x lo, the truncated tail of lo) are <= 2^-21 of a product, unbiased.
//   flow_feats_split_kernel : F = ELU(img_base[img] + betaW . beta) -> [R][512] = [hi(256) | lo(256)]
//   flow_ctx_gemm_kernel    : tile 128 rows x 128 outputs (two joints), K in 8 steps of 32 floats (128-byte swizzled rows),
//                             3-stage TMA ring of 64 KB stages; 8 epilogue warps (TMEM lane quarter x joint) write U in the
//                             layout the sampling kernel streams ([row group of NR][joint][o][NR]).
constexpr int CG_STAGES = 3, CG_THREADS = 320, CG_BN = 128;
constexpr int CG_STAGE_BYTES = 2 * 128 * 128 + 2 * CG_BN * 128;

__global__ void flow_feats_split_kernel(const FlowParams P, const float* __restrict__ img_base, const float* __restrict__ betas,
                                        const int* __restrict__ img_index, int R, float* __restrict__ F) {
    HF_PDL_SYNC();
    const int e = blockIdx.x * blockDim.x + threadIdx.x;          // one thread: 4 consecutive features of one row
    if (e >= R * (FEATS / 4)) return;
    const int r = e / (FEATS / 4), o = (e - r * (FEATS / 4)) * 4;
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(img_base + (size_t)__ldg(img_index + r) * FEATS + o));
    float a[4] = {b4.x, b4.y, b4.z, b4.w};
    for (int l = 0; l < P.nb; ++l) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(P.betaW + l * FEATS + o));
        const float bt = __ldg(betas + (size_t)r * P.nb + l);
        a[0] = fmaf(w.x, bt, a[0]); a[1] = fmaf(w.y, bt, a[1]); a[2] = fmaf(w.z, bt, a[2]); a[3] = fmaf(w.w, bt, a[3]);
    }
    float hi[4], lo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float x = elu(a[i]);
        uint32_t hb;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"(x));
        hi[i] = __uint_as_float(hb);
        lo[i] = x - hi[i];
    }
    *reinterpret_cast<float4*>(F + (size_t)r * 2 * FEATS + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(F + (size_t)r * 2 * FEATS + FEATS + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
}

__global__ void __launch_bounds__(CG_THREADS, 1)
flow_ctx_gemm_kernel(const __grid_constant__ CUtensorMap mapF, const __grid_constant__ CUtensorMap mapW, int R, int J, int NR,
                     float* __restrict__ U) {
    extern __shared__ uint8_t cg_smem[];
    __shared__ __align__(8) uint64_t bars[2 * CG_STAGES + 1];
    __shared__ uint32_t tmem_base_s;
    const uint32_t tile_base = (smem_u32(cg_smem) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128, n0 = blockIdx.y * CG_BN;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[CG_STAGES]), tfull = smem_u32(&bars[2 * CG_STAGES]);
    if (threadIdx.x == 0) {
        for (int s = 0; s < CG_STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(tfull, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)CG_BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    HF_PDL_SYNC();
    const uint32_t tmem_base = tmem_base_s;
    constexpr int NIT = FEATS / 32;
    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < NIT; ++it) {
                const int st = it % CG_STAGES;
                const uint32_t ph = (uint32_t)(it / CG_STAGES) & 1u;
                mbar_wait(empty0 + 8 * st, ph ^ 1u);
                const uint32_t sa = tile_base + st * CG_STAGE_BYTES, fb = full0 + 8 * st;
                mbar_expect_tx(fb, CG_STAGE_BYTES);
                tma_load_2d(sa, &mapF, fb, it * 32, m0);
                tma_load_2d(sa + 16384, &mapF, fb, FEATS + it * 32, m0);
                tma_load_2d(sa + 32768, &mapW, fb, it * 32, n0);                      // rows past J*64 are zero-filled
                tma_load_2d(sa + 32768 + CG_BN * 128, &mapW, fb, FEATS + it * 32, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, CG_BN);
            for (int it = 0; it < NIT; ++it) {
                const int st = it % CG_STAGES;
                const uint32_t ph = (uint32_t)(it / CG_STAGES) & 1u;
                mbar_wait(full0 + 8 * st, ph);
                tcgen05_fence_after();
                const uint32_t sa = tile_base + st * CG_STAGE_BYTES;
                const uint64_t a_hi = umma_desc_sw128(sa), a_lo = umma_desc_sw128(sa + 16384);
                const uint64_t b_hi = umma_desc_sw128(sa + 32768), b_lo = umma_desc_sw128(sa + 32768 + CG_BN * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, a_lo + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, (uint32_t)((it | k) != 0));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, a_hi + (uint64_t)(2 * k), b_lo + (uint64_t)(2 * k), idesc, 1u);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_base, a_hi + (uint64_t)(2 * k), b_hi + (uint64_t)(2 * k), idesc, 1u);
                umma_commit(empty0 + 8 * st);
            }
            umma_commit(tfull);
        }
    } else {
        const int q = warp & 3, half = (warp - 2) >> 2;     // TMEM lane quarter (= warp % 4), joint of the pair
        const int r = m0 + q * 32 + lane, j = blockIdx.y * 2 + half;
        mbar_wait_warp(tfull, 0);
        tcgen05_fence_after();
        uint32_t v[64];
        const uint32_t ta = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 64);
        tmem_ld32(ta, v);
        tmem_ld32(ta + 32, v + 32);
        tmem_ld_wait();
        if (r < R && j < J) {
            float* dst = U + ((size_t)(r / NR) * J + j) * CTX * NR + (r % NR);
#pragma unroll
            for (int o = 0; o < CTX; ++o) dst[o * NR] = __uint_as_float(v[o]);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)CG_BN) : "memory");
    }
}

// =====================================  sampling  =====================================
// One CTA owns NR rows for the whole tree.
//  prologue : image features -> U[j] = Wfeat_j . feats for all joints, computed BEFORE this kernel by flow_ctx_gemm_kernel
//             (tensor cores, split tf32) and parked in an L2-resident scratch, laid out so that a joint's 64 x NR block is
//             contiguous; with have_U = 0 the CTA computes its own rows on the CUDA cores (cross-check path)
//  chain    : per joint, weights arrive in shared memory by 1-D bulk copies issued one joint ahead (mbarrier
//             complete_tx), U(j) by cp.async; every dense layer then runs on low-latency LDS operands.
template <int NR>
__global__ void __launch_bounds__(HF_NT, 1)
flow_sample_kernel(const __grid_constant__ FlowParams P, const float* __restrict__ img_base, const float* __restrict__ betas,
                   const int* __restrict__ img_index, const float* __restrict__ base_noise, int R, int Rn,
                   float* __restrict__ rotmats, float* __restrict__ axisangle_pe, float* __restrict__ Uscratch, int have_U, int dbg) {
    long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
    auto lap = [&](int c) { if (dbg) { const long long t = clock64(); tph[c] += t - tlast; tlast = t; } };
    using L = SmemLayout<NR>;
    using S = SampleSmem<NR>;
    extern __shared__ __align__(16) float smraw[];
    __shared__ __align__(8) uint64_t wbar[3];
    float* sm = smraw + S::Act - L::Ps;              // so that sm + L::Ps, sm + L::Cs ... address the activation block
    float* Fs = sm + L::Scratch;                      // image features live in the scratch area during the prologue
    float* Us = smraw + S::Us;
    const int r0 = blockIdx.x * NR;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = smem_u32(&wbar[0]);
    float* Ucta = Uscratch + (size_t)blockIdx.x * P.J * CTX * NR;
    auto joint_bytes_anc = [&](int j) { return (uint32_t)((9 * P.anc_cnt[j] * CTXP + CTX) * 4); };
    auto issue_anc = [&](int j) {
        mbar_expect_tx(bar0, joint_bytes_anc(j));
        bulk_load_1d(smem_u32(smraw + S::Wanc), P.jpack + P.off_jb[j], joint_bytes_anc(j), bar0);
    };
    auto issue_cpl = [&](int j, int t) {
        mbar_expect_tx(bar0 + 8 * (1 + t), COUPLING_FLOATS * 4);
        bulk_load_1d(smem_u32(smraw + (t ? S::Wc1 : S::Wc0)),
                     P.jpack + P.off_jb[j] + 9 * P.anc_cnt[j] * CTXP + CTX + t * COUPLING_FLOATS, COUPLING_FLOATS * 4,
                     bar0 + 8 * (1 + t));
    };
    auto fetch_U = [&](int j) {   // 64 x NR floats, contiguous in the scratch
        const float* src = Ucta + (size_t)j * CTX * NR;
        float* dst = Us + (j & 1) * CTX * NR;
        for (int i = tid; i < CTX * NR / 4; i += HF_NT) cp_async16(dst + i * 4, src + i * 4);
        cp_async_commit();
    };
    if (tid == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(bar0 + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid == 0) { issue_anc(0); issue_cpl(0, 0); if (P.T > 1) issue_cpl(0, 1); }     // weights: constant, no dependency on the predecessor
    HF_PDL_SYNC();
    if (!have_U) {
        // CUDA-core prologue (cross-check path, HF_FLOW_SIMT_PROLOGUE=1): 2*J output tiles of 32, full K = FEATS per warp
        image_feats<NR>(P, img_base, betas, img_index, r0, R, Fs);
        __syncthreads();
        for (int t = warp; t < 2 * P.J; t += HF_NW)
            warp_gemm<NR, false, 8>(P.wfeat + t * 32, P.J * CTX, 0, 0, FEATS, PlainRow{Fs, NR},
                                    Ucta + (size_t)(t >> 1) * CTX * NR + (t & 1) * 32 * NR, lane);
        __threadfence();
        __syncthreads();
    }
    lap(0);
    fetch_U(0);

    for (int j = 0; j < P.J; ++j) {
        const uint32_t par = (uint32_t)j & 1u;
        const int Ka = 9 * P.anc_cnt[j];
        // one thread observes each weight barrier, the next CTA barrier publishes that to the rest: 512 threads polling the same
        // mbarrier cost ~450 cycles per wait
        cp_async_wait_all();
        if (tid == 0) mbar_wait(bar0, par);                // ancestor block
        if (tid == 32) mbar_wait(bar0 + 8, par);           // first coupling (the second: below); a completed wait is still ~150 cycles, so in parallel
        __syncthreads();
        lap(1);
        // base sample (zero for point-estimate rows); first permutation is the identity.  Written here (row CTX of
        // Cs and Zs are not touched by the context layer), ordered by the layer's own barriers.
        if (tid < NR) {
            const int r = r0 + tid;
            float z0 = 0.f, z1 = 0.f, z2 = 0.f;
            if (r < Rn) {
                const float* z = base_noise + ((size_t)r * P.J + j) * 3;
                z0 = z[0]; z1 = z[1]; z2 = z[2];
            }
            sm[L::Zs + tid] = z0; sm[L::Zs + NR + tid] = z1; sm[L::Zs + 2 * NR + tid] = z2;
            sm[L::Cs + CTX * NR + tid] = z0;
        }
        // context = ELU(U_j + b + Wanc . vec(ancestor rotations))
        dense_layer<NR, 2, 1, true>(smraw + S::Wanc, smraw + S::Wanc + Ka * CTXP, Ka, AncRow{sm + L::Ps, P.anc[j], NR},
                                    sm + L::Scratch, sm + L::Cs, Us + (j & 1) * CTX * NR);
        lap(2);
        if (j + 1 < P.J) {
            if (tid == 0) issue_anc(j + 1);
            fetch_U(j + 1);
        }
        for (int t = 0; t < P.T; ++t) {
            coupling_nn_smem<NR>(smraw + (t ? S::Wc1 : S::Wc0), sm);
            lap(3);
            if (tid == 0 && j + 1 < P.J) issue_cpl(j + 1, t);
            spline_knots_full<NR>(sm + L::Raw, P.radius, sm + L::Ha);      // 72 x NR floats: spans Ha and the start of Hb (both free here)
            lap(4);
            // spline on the two trailing coordinates, then rotate the vector for the next Permute
            // (pyro_conditional_norm_flow.py:46-62: with <=2 transforms the only non-identity Permute is [1,2,0],
            //  applied to the running vector before the second coupling).
            if (tid < 2 * NR) {        // 2*NR <= 64: whole warps (NR in {8,16,24} -> 16/32/48 threads: lanes come in (d=0,d=1) pairs)
                const int s = tid >> 1, d = tid & 1;
                const float x = sm[L::Zs + (1 + d) * NR + s];
                const float y0 = sm[L::Zs + s];
                const float mine = spline_forward_pre(x, P.radius, sm + L::Ha + (s * 2 + d) * KNF);
                const unsigned act = __activemask();
                const float other = __shfl_xor_sync(act, mine, 1);
                if (d == 0) {
                    const float y1 = mine, y2 = other;
                    if (t + 1 < P.T) {     // next permutation relative to the current order is always [1,2,0]
                        sm[L::Zs + s] = y1; sm[L::Zs + NR + s] = y2; sm[L::Zs + 2 * NR + s] = y0;
                        sm[L::Cs + CTX * NR + s] = y1;
                    } else {
                        sm[L::Zs + NR + s] = y1; sm[L::Zs + 2 * NR + s] = y2;
                    }
                }
            }
            if (tid == 2 * NR + 32 && t + 1 < P.T) mbar_wait(bar0 + 8 * (2 + t), par);       // next coupling's weights (issued a whole joint ago)
            __syncthreads();
            lap(5);
        }
        // radial tanh -> exp map -> store (shared copy feeds the descendants' contexts)
        if (tid < NR) {
            const int r = r0 + tid;
            float x = sm[L::Zs + tid], y = sm[L::Zs + NR + tid], z = sm[L::Zs + 2 * NR + tid];
            const float n = sqrtf(x * x + y * y + z * z);
            if (n > 1e-7f) {
                const float th = tanhf(n / P.radius);
                x = th * (x / n) * P.radius; y = th * (y / n) * P.radius; z = th * (z / n) * P.radius;
            }
            float Rm[9];
            if (r < Rn) so3_exp_f64((double)x, (double)y, (double)z, Rm);
            else rodrigues_f32(x, y, z, Rm);
#pragma unroll
            for (int e = 0; e < 9; ++e) sm[L::Ps + (j * 9 + e) * NR + tid] = Rm[e];
            if (r < R) {
                float* o = rotmats + ((size_t)r * P.J + j) * 9;
#pragma unroll
                for (int e = 0; e < 9; ++e) o[e] = Rm[e];
                if (r >= Rn && axisangle_pe) {
                    float* a = axisangle_pe + ((size_t)(r - Rn) * P.J + j) * 3;
                    a[0] = x; a[1] = y; a[2] = z;
                }
            }
        }
        __syncthreads();
        lap(6);
    }
    if (dbg && blockIdx.x == 0 && tid == 0)
        printf("flow phases (cycles, CTA 0): prologue %lld | weight waits %lld | context layers %lld | coupling MLPs %lld | knots %lld | spline %lld | exp map + store %lld\n",
               tph[0], tph[1], tph[2], tph[3], tph[4], tph[5], tph[6]);
}

// =====================================  sampling, level-parallel  =====================================
// The chain above walks the joints in index order, but the kinematic tree only has a few dependency LEVELS (SMPL: {1,2,3},
// {4,5,6}, {7,8,9}, {10..14}, ... : 9 rounds of <= 3 joints instead of 23 steps), and every phase of the chain is latency-bound.
// Here the CTA is three independent groups of 192 threads (+ one producer warp each); in a round each group takes one joint of the
// level through its context layer, couplings, spline and exp map with GROUP barriers (bar.sync id, 192) only; rounds end with a
// barrier over the consumer threads that publishes the new rotations.  A joint's weights (~100 KB) do not fit three times next to
// the activations, so each group streams them LAYER BY LAYER: the [W | b] block of the layer after next (<= 17.9 KB) is fetched
// by the group's producer warp with one bulk copy into a two-deep ring (full / empty mbarriers) while the current layer computes.
// U_j (the precomputed image-feature term) is added straight from the L2-resident scratch.  64-output layers use an 8 x 4
// register tile (dense_layer_g84), 32-output layers the 4 x 4 tile; the arithmetic and its order are those of the joint-by-joint
// kernel, so both produce the same bits.
constexpr int LG = 3, LGT = 192, LNT = LG * LGT;
constexpr int LWB = 65 * (64 + WPAD) + 64;        // largest [W | b] block in floats: first coupling layer; the ancestor block is <= 63*68 + 64

struct LevelSched { int count[LG]; signed char joint[LG][HF_FJ]; };      // the joints of every group in execution order (a static list schedule)

template <int NR>
struct LevelSmem {
    static constexpr int ActFloats = (CTX + 1 + 64 + 32 + 32 + 64 + 4) * NR;      // Cs | Ha | Hb | Hc | Raw | Zs of one group
    static constexpr int Ps = 0;                                                  // [HF_FJ*9][NR], shared by the groups
    static constexpr int Act = Ps + HF_FJ * 9 * NR;
    static constexpr int Wb = Act + LG * ActFloats;                               // [LG][2][LWB]
    static constexpr int Total = Wb + LG * 2 * LWB;
};

// dense layer of ONE thread group: same tile mapping and arithmetic as dense_layer (4 outputs x 4 rows per 4 lanes), tiles
// strided over the group's LGT / 4 tile slots; weights and bias in shared memory; `add` (optional) is read from global memory
// first so that its latency hides behind the k loop.  No barrier inside.
template <int NR, int OT, int ACT, typename XRow>
__device__ __forceinline__ void dense_layer_g(const float* __restrict__ W, const float* __restrict__ bias, int K, XRow xrow,
                                              float* dst, const float* __restrict__ add, int lt) {
    constexpr int O = OT * 32, LDW = O + 4, RG = NR / 4, NTILE = (O / 4) * RG;
    const int ks = lt & 3;
    for (int tile = lt >> 2; tile < NTILE; tile += LGT / 4) {        // whole warps: NTILE and LGT / 4 are multiples of 8
        const int og = tile / RG, rg = tile - og * RG;
        const int o = og * 4 + ks;
        float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (add) a4 = __ldg(reinterpret_cast<const float4*>(add + o * NR + rg * 4));
        float2 acc[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = make_float2(0.f, 0.f);
        const float* wp = W + og * 4;
#pragma unroll 4
        for (int k = ks; k < K; k += 4) {
            const float4 w = *reinterpret_cast<const float4*>(wp + (size_t)k * LDW);
            const float4 x = *reinterpret_cast<const float4*>(xrow(k) + rg * 4);
            const float2 xa = make_float2(x.x, x.y), xb = make_float2(x.z, x.w);
            const float2 w0 = make_float2(w.x, w.x), w1 = make_float2(w.y, w.y), w2 = make_float2(w.z, w.z), w3 = make_float2(w.w, w.w);
            acc[0][0] = fma2(w0, xa, acc[0][0]); acc[0][1] = fma2(w0, xb, acc[0][1]);
            acc[1][0] = fma2(w1, xa, acc[1][0]); acc[1][1] = fma2(w1, xb, acc[1][1]);
            acc[2][0] = fma2(w2, xa, acc[2][0]); acc[2][1] = fma2(w2, xb, acc[2][1]);
            acc[3][0] = fma2(w3, xa, acc[3][0]); acc[3][1] = fma2(w3, xb, acc[3][1]);
        }
        float r[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float vx = acc[i][h].x, vy = acc[i][h].y;
                vx += __shfl_xor_sync(0xffffffffu, vx, 1); vy += __shfl_xor_sync(0xffffffffu, vy, 1);
                vx += __shfl_xor_sync(0xffffffffu, vx, 2); vy += __shfl_xor_sync(0xffffffffu, vy, 2);
                if (i == ks) { r[2 * h] = vx; r[2 * h + 1] = vy; }
            }
        }
        const float b = bias[o];
        float v[4] = {r[0] + b + a4.x, r[1] + b + a4.y, r[2] + b + a4.z, r[3] + b + a4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            if (ACT == 1) v[q] = elu(v[q]);
            if (ACT == 2) v[q] = fmaxf(v[q], 0.f);
        }
        *reinterpret_cast<float4*>(dst + o * NR + rg * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// 64-output layers of one thread group: tile = 8 outputs x 4 rows, 4 k-lanes (48 tiles x 4 = the 192 threads at NR = 24: ONE pass).
// Per k a lane loads 8 weights (two LDS.128 that are broadcasts for the six tiles of an output group sharing a warp) and 4 row
// values for 16 FFMA2: 5 shared-memory wavefronts per 16 FFMA2 against 8 per 8 for the 4 x 4 tile, which keeps the shared-memory
// pipe (60 % busy with the small tile) off the critical path.  The 4 k-slices are combined by a transpose-reduction (16 + 8
// shuffles) that leaves every lane with two outputs of the tile's 4 rows.
template <int NR, int ACT, typename XRow>
__device__ __forceinline__ void dense_layer_g84(const float* __restrict__ W, const float* __restrict__ bias, int K, XRow xrow,
                                                float* dst, const float* __restrict__ add, int lt) {
    constexpr int O = 64, LDW = O + 4, RG = NR / 4, NTILE = 8 * RG;
    const int ks = lt & 3;
    for (int tile = lt >> 2; tile < NTILE; tile += LGT / 4) {
        const int og = tile / RG, rg = tile - og * RG;
        float f[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) f[i] = 0.f;
        float2* acc = reinterpret_cast<float2*>(f);          // acc[o * 2 + rp]: output o (0..7), row pair rp
        const float* wp = W + og * 8;
#pragma unroll 4
        for (int k = ks; k < K; k += 4) {
            const float4 wa = *reinterpret_cast<const float4*>(wp + (size_t)k * LDW), wc = *reinterpret_cast<const float4*>(wp + (size_t)k * LDW + 4);
            const float4 x = *reinterpret_cast<const float4*>(xrow(k) + rg * 4);
            const float2 xa = make_float2(x.x, x.y), xb = make_float2(x.z, x.w);
            const float w[8] = {wa.x, wa.y, wa.z, wa.w, wc.x, wc.y, wc.z, wc.w};
#pragma unroll
            for (int o = 0; o < 8; ++o) {
                const float2 ww = make_float2(w[o], w[o]);
                acc[o * 2] = fma2(ww, xa, acc[o * 2]);
                acc[o * 2 + 1] = fma2(ww, xb, acc[o * 2 + 1]);
            }
        }
        // (lanes 0+1 and 2+3 first, then the pairs: the summation order of dense_layer, so both kernels agree bit for bit)
#pragma unroll
        for (int step = 0; step < 2; ++step) {
            const int sft = 1 << step, n = 16 >> step;
            const bool up = (ks & sft) != 0;
#pragma unroll
            for (int i = 0; i < n; ++i) {
                const float keep = up ? f[i + n] : f[i], send = up ? f[i] : f[i + n];
                f[i] = keep + __shfl_xor_sync(0xffffffffu, send, sft);
            }
        }
        // lane ks holds entries [e8*8, +8) of the (output, row) grid with e8 = 2*(ks&1) + (ks>>1): outputs og*8 + 2*e8 and +1, 4 rows each
        const int e8 = 2 * (ks & 1) + (ks >> 1);
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int o = og * 8 + 2 * e8 + q;
            float4 a4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (add) a4 = __ldg(reinterpret_cast<const float4*>(add + o * NR + rg * 4));
            const float b = bias[o];
            float v[4] = {f[q * 4] + b + a4.x, f[q * 4 + 1] + b + a4.y, f[q * 4 + 2] + b + a4.z, f[q * 4 + 3] + b + a4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (ACT == 1) v[e] = elu(v[e]);
                if (ACT == 2) v[e] = fmaxf(v[e], 0.f);
            }
            *reinterpret_cast<float4*>(dst + o * NR + rg * 4) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// spline_knots_full for one thread group (no barrier)
template <int NR>
__device__ __forceinline__ void spline_knots_full_g(const float* __restrict__ raw, float bound, float* __restrict__ KN, int lt) {
    for (int t = lt; t < NR * 8; t += LGT) knot_task<NR>(raw, bound, KN, t);
}

template <int NR>
__global__ void __launch_bounds__(LNT + 32 * LG, 1)
flow_sample_levels_kernel(const __grid_constant__ FlowParams P, const __grid_constant__ LevelSched S, const float* __restrict__ base_noise,
                          int R, int Rn, float* __restrict__ rotmats, float* __restrict__ axisangle_pe, const float* __restrict__ U, int dbg) {
    using LS = LevelSmem<NR>;
    long long tph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
    auto lap = [&](int c) { if (dbg) { const long long tt = clock64(); tph[c] += tt - tlast; tlast = tt; } };
    extern __shared__ __align__(16) float smraw[];
    // weight blocks: one bulk copy each, issued by the group's PRODUCER warp (warps 18..20; issuing a bulk copy costs the issuing
    // thread ~500 cycles, which sat on every layer's critical path when the group's first thread did it) into a two-deep ring
    // guarded by full / empty mbarriers
    __shared__ __align__(8) uint64_t lbar[LG][4];         // full[0..1], empty[0..1]
    __shared__ int anc_off[LG][64];                        // row offset into Ps of every input k of the group's context layer
    __shared__ volatile int jdone[HF_FJ];                  // joint finished: its rotations are in Ps
    const int tid = threadIdx.x;
    const int NL = 1 + 4 * P.T;                            // layers per joint: ancestor block, then 4 per coupling
    if (tid < HF_FJ) jdone[tid] = 0;
    if (tid < LG) {
        const uint32_t bb = smem_u32(&lbar[tid][0]);
        for (int i = 0; i < 4; ++i) mbar_init(bb + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid >= LNT) {
        // ---------------- producer warps ----------------
        const int g = (tid - LNT) >> 5;
        if ((tid & 31) == 0) {
            const uint32_t bar_g = smem_u32(&lbar[g][0]);
            float* wb = smraw + LS::Wb + g * 2 * LWB;
            int blk = 0;                                   // running block counter of the group: buffer = blk & 1, use = blk >> 1
            for (int rd = 0; rd < S.count[g]; ++rd) {
                const int j = S.joint[g][rd];
                const float* jb = P.jpack + P.off_jb[j];
                const int na = 9 * P.anc_cnt[j] * CTXP + CTX;
                for (int l = 0; l < NL; ++l, ++blk) {
                    const float* src = jb;
                    int n = na;
                    if (l > 0) {
                        const int tt = (l - 1) >> 2, q = (l - 1) & 3;
                        const int off = q == 0 ? OFF_W0 : (q == 1 ? OFF_W1 : (q == 2 ? OFF_W2 : OFF_W3));
                        const int end = q == 0 ? OFF_W1 : (q == 1 ? OFF_W2 : (q == 2 ? OFF_W3 : COUPLING_FLOATS));
                        src = jb + na + tt * COUPLING_FLOATS + off;
                        n = end - off;
                    }
                    const int buf = blk & 1;
                    mbar_wait_long(bar_g + 8 * (2 + buf), (uint32_t)(((blk >> 1) & 1) ^ 1));       // buffer released by the consumers
                    mbar_expect_tx(bar_g + 8 * buf, (uint32_t)n * 4u);
                    bulk_load_1d(smem_u32(wb + buf * LWB), src, (uint32_t)n * 4u, bar_g + 8 * buf);
                }
            }
        }
        return;
    }
    // ---------------- consumer groups ----------------
    const int g = tid / LGT, lt = tid - g * LGT;
    float* Ps = smraw + LS::Ps;
    float* Cs = smraw + LS::Act + g * LS::ActFloats;
    float* Ha = Cs + (CTX + 1) * NR;
    float* Hb = Ha + 64 * NR;
    float* Hc = Hb + 32 * NR;
    float* Raw = Hc + 32 * NR;
    float* Zs = Raw + 64 * NR;
    float* wb = smraw + LS::Wb + g * 2 * LWB;
    const int r0 = blockIdx.x * NR;
    const float* Ucta = U + (size_t)blockIdx.x * P.J * CTX * NR;
    const uint32_t bar_g = smem_u32(&lbar[g][0]);
    auto gbar = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "n"(LGT) : "memory"); };
    int blk = 0;
    // ONE group barrier per layer (a bar.sync costs ~300 cycles on this loaded SM): it publishes the layer's outputs and, because
    // the group's first thread has observed the next block's mbarrier just before it, the next layer's weights; the finished
    // buffer goes back to the producer right behind it
    auto first_block = [&]() -> const float* {
        if (lt == 0) mbar_wait(bar_g + 8 * (blk & 1), (uint32_t)((blk >> 1) & 1));
        gbar();
        return wb + (blk & 1) * LWB;
    };
    auto next_block = [&](bool has_next) -> const float* {
        if (lt == 0 && has_next) mbar_wait(bar_g + 8 * ((blk + 1) & 1), (uint32_t)(((blk + 1) >> 1) & 1));
        gbar();
        if (lt == 0) mbar_arrive(bar_g + 8 * (2 + (blk & 1)));
        ++blk;
        return wb + (blk & 1) * LWB;
    };
    HF_PDL_SYNC();
    lap(0);
    for (int rd = 0; rd < S.count[g]; ++rd) {
        const int j = S.joint[g][rd];
        {
            const int Ka = 9 * P.anc_cnt[j];
            // the ancestors' rotations must be in Ps: they were produced earlier by this or another group (the deepest ancestor
            // finishing implies the others); one thread polls, the group barrier in first_block() publishes it
            if (lt == 0 && Ka > 0) {
                for (int q = 0; q < P.anc_cnt[j]; ++q) {
                    // bounded like every wait of this library: a schedule bug must surface as a trapped launch, never as a hung GPU
                    unsigned long long t0 = 0, t1;
                    for (unsigned spin = 0; !jdone[(int)P.anc[j][q]]; ++spin) {
                        if ((spin & 0xffffu) == 0xffffu) {
                            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
                            if (t0 == 0) t0 = t1;
                            if (t1 - t0 > 2000000000ull) __trap();     // 2 s
                        }
                    }
                }
                __threadfence_block();
            }
            if (lt < NR) {       // base sample (zero for point-estimate rows); the first permutation is the identity
                const int r = r0 + lt;
                float z0 = 0.f, z1 = 0.f, z2 = 0.f;
                if (r < Rn) {
                    const float* z = base_noise + ((size_t)r * P.J + j) * 3;
                    z0 = z[0]; z1 = z[1]; z2 = z[2];
                }
                Zs[lt] = z0; Zs[NR + lt] = z1; Zs[2 * NR + lt] = z2;
                Cs[CTX * NR + lt] = z0;
            }
            if (lt >= 64 && lt < 64 + Ka) {          // (another warp than the base-sample one)
                const int k = lt - 64, a = k / 9;
                anc_off[g][k] = (P.anc[j][a] * 9 + (k - a * 9)) * NR;
            }
            const float* w = first_block();
            lap(1);
            // context = ELU(U_j + b + Wanc . vec(ancestor rotations))
            dense_layer_g84<NR, 1>(w, w + Ka * CTXP, Ka, TabRow{Ps, anc_off[g]}, Cs, Ucta + (size_t)j * CTX * NR, lt);
            w = next_block(true);
            lap(2);
            for (int t = 0; t < P.T; ++t) {
                dense_layer_g84<NR, 2>(w, w + (OFF_B0 - OFF_W0), CTX + 1, PlainRow{Cs, NR}, Ha, nullptr, lt);
                w = next_block(true);
                dense_layer_g<NR, 1, 2>(w, w + (OFF_B1 - OFF_W1), H1, PlainRow{Ha, NR}, Hb, nullptr, lt);
                w = next_block(true);
                dense_layer_g<NR, 1, 2>(w, w + (OFF_B2 - OFF_W2), H2, PlainRow{Hb, NR}, Hc, nullptr, lt);
                w = next_block(true);
                dense_layer_g84<NR, 0>(w, w + (OFF_B3 - OFF_W3), H3, PlainRow{Hc, NR}, Raw, nullptr, lt);
                w = next_block(t + 1 < P.T);                 // (also the barrier between the raw outputs and the knot pass)
                lap(3);
                spline_knots_full_g<NR>(Raw, P.radius, Ha, lt);          // 72 x NR floats over Ha and the start of Hb (both free here)
                gbar();
                lap(4);
                if (lt < 2 * NR) {
                    const int s = lt >> 1, d = lt & 1;
                    const float x = Zs[(1 + d) * NR + s];
                    const float y0 = Zs[s];
                    const float mine = spline_forward_pre(x, P.radius, Ha + (s * 2 + d) * KNF);
                    const unsigned act = __activemask();
                    const float other = __shfl_xor_sync(act, mine, 1);
                    if (d == 0) {
                        const float y1 = mine, y2 = other;
                        if (t + 1 < P.T) {     // next permutation relative to the current order is always [1,2,0]
                            Zs[s] = y1; Zs[NR + s] = y2; Zs[2 * NR + s] = y0;
                            Cs[CTX * NR + s] = y1;
                        } else {
                            Zs[NR + s] = y1; Zs[2 * NR + s] = y2;
                        }
                    }
                }
                gbar();
                lap(5);
            }
            // radial tanh -> exp map -> store (the shared copy feeds the descendants' contexts after the round barrier)
            if (lt < NR) {
                const int r = r0 + lt;
                float x = Zs[lt], y = Zs[NR + lt], z = Zs[2 * NR + lt];
                const float n = sqrtf(x * x + y * y + z * z);
                if (n > 1e-7f) {
                    const float th = tanhf(n / P.radius);
                    x = th * (x / n) * P.radius; y = th * (y / n) * P.radius; z = th * (z / n) * P.radius;
                }
                float Rm[9];
                if (r < Rn) so3_exp_f64((double)x, (double)y, (double)z, Rm);
                else rodrigues_f32(x, y, z, Rm);
#pragma unroll
                for (int e = 0; e < 9; ++e) Ps[(j * 9 + e) * NR + lt] = Rm[e];
                if (r < R) {
                    float* o = rotmats + ((size_t)r * P.J + j) * 9;
#pragma unroll
                    for (int e = 0; e < 9; ++e) o[e] = Rm[e];
                    if (r >= Rn && axisangle_pe) {
                        float* a = axisangle_pe + ((size_t)(r - Rn) * P.J + j) * 3;
                        a[0] = x; a[1] = y; a[2] = z;
                    }
                }
            }
        }
        gbar();                                  // the joint's rotations are written by every row thread of the group ...
        if (lt == 0) { __threadfence_block(); jdone[j] = 1; }      // ... before the flag goes up
        lap(6);
    }
    if (dbg && blockIdx.x == 0 && lt == 0)
        printf("flow levels group %d (cycles): start %lld | weight waits %lld | context %lld | coupling layers %lld | knots %lld | spline %lld | exp map %lld | round barrier %lld\n",
               g, tph[0], tph[1], tph[2], tph[3], tph[4], tph[5], tph[6], tph[7]);
}

// =====================================  contexts for teacher forcing  =====================================
template <int NR>
__global__ void __launch_bounds__(HF_NT, 1)
flow_context_kernel(const __grid_constant__ FlowParams P, const float* __restrict__ img_base, const float* __restrict__ betas,
                    const int* __restrict__ img_index, const float* __restrict__ anc_rotmats, int R,
                    float* __restrict__ ctx_out) {
    using L = SmemLayout<NR>;
    extern __shared__ __align__(16) float sm[];
    const int r0 = blockIdx.x * NR;
    const int j = blockIdx.y;
    image_feats<NR>(P, img_base, betas, img_index, r0, R, sm + L::Fs);
    for (int e = threadIdx.x; e < P.J * 9 * NR; e += HF_NT) {
        const int s = e / (P.J * 9), q = e - s * (P.J * 9);
        const int r = r0 + s;
        sm[L::Ps + q * NR + s] = (r < R) ? anc_rotmats[(size_t)r * P.J * 9 + q] : 0.f;
    }
    __syncthreads();
    const int Kc = FEATS + 9 * P.anc_cnt[j];
    dense_layer<NR, 2, 1>(P.pack + P.off_ctxW[j], P.pack + P.off_ctxB[j], Kc,
                          CtxRow{sm + L::Fs, sm + L::Ps, P.anc[j], NR}, sm + L::Scratch, sm + L::Cs);
    for (int e = threadIdx.x; e < CTX * NR; e += HF_NT) {
        const int s = e / CTX, o = e - s * CTX;
        const int r = r0 + s;
        if (r < R) ctx_out[((size_t)r * P.J + j) * CTX + o] = sm[L::Cs + o * NR + s];
    }
}

// =====================================  log_prob  =====================================
// CTA = (joint, NR/3 rows); the three pre-images of each target rotation are 3 rows of the coupling MLPs.
// ALGEBRA = true evaluates the so(3) density of a given vector instead (single candidate per row).
template <int NR, bool ALGEBRA>
__global__ void __launch_bounds__(HF_NT, 1)
flow_logprob_kernel(const __grid_constant__ FlowParams P, const float* __restrict__ ctx, int ctx_row_stride, int joint_first,
                    int joint_count, const double* __restrict__ rot, const float* __restrict__ valg, int R,
                    float* __restrict__ out) {
    using L = SmemLayout<NR>;
    constexpr int NC = ALGEBRA ? 1 : 3;
    constexpr int RT = NR / NC;               // target rows per CTA
    extern __shared__ __align__(16) float sm[];
    __shared__ float s_lp[NR];                // running log_prob per candidate
    __shared__ int s_mask[NR];
    const int jj = blockIdx.y, j = joint_first + jj;
    const int r0 = blockIdx.x * RT;
    const int tid = threadIdx.x;
    const double PI = 3.14159265358979323846;
    // contexts, replicated per candidate: candidate c of row s lives in column c*RT + s
    for (int e = tid; e < CTX * NR; e += HF_NT) {
        const int col = e / CTX, o = e - col * CTX;
        const int r = r0 + (col % RT);
        sm[L::Cs + o * NR + col] = (r < R) ? ctx[(size_t)r * ctx_row_stride + j * CTX + o] : 0.f;
    }
    if (tid < RT) {
        const int r = r0 + tid;
        double x[3] = {0.0, 0.0, 0.0};
        if (r < R) {
            if (ALGEBRA) {
                const float* v = valg + ((size_t)r * joint_count + jj) * 3;
                x[0] = v[0]; x[1] = v[1]; x[2] = v[2];
            } else {
                double rm[9];
                const double* rp = rot + ((size_t)r * joint_count + jj) * 9;
                for (int e = 0; e < 9; ++e) rm[e] = rp[e];
                so3_log_f64(rm, x);
            }
        }
        for (int c = 0; c < NC; ++c) {
            double v[3] = {x[0], x[1], x[2]};
            int mask = 1;
            if (c > 0) {   // so3_xset k = -1 (c=1), +1 (c=2); masked by |.| < radius, masked-out -> 0 (so3_exp_transform.py:36-41)
                const double n = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
                const double k = (c == 1) ? -1.0 : 1.0;
                const double f = (n + 2.0 * PI * k);
                v[0] = x[0] / n * f; v[1] = x[1] / n * f; v[2] = x[2] / n * f;
                const double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
                mask = (vn < (double)P.radius) ? 1 : 0;   // NaN (n == 0) compares false
                if (!mask) { v[0] = v[1] = v[2] = 0.0; }
            }
            const int col = c * RT + tid;
            float lp = ALGEBRA ? 0.f : -(float)so3_log_abs_det(v[0], v[1], v[2]);
            // ToTransform: fp64 -> fp32; then invert the radial tanh in fp64 (scaled_radial_tanh_transform.py:41-59)
            const float y0 = (float)v[0], y1 = (float)v[1], y2 = (float)v[2];
            const double dy0 = y0, dy1 = y1, dy2 = y2;
            const double yn = sqrt(dy0 * dy0 + dy1 * dy1 + dy2 * dy2);
            float x0 = y0, x1 = y1, x2 = y2;
            if (yn > 1e-7) {
                const double a = atanh(yn / (double)P.radius);
                x0 = (float)(a * (dy0 / yn) * (double)P.radius);
                x1 = (float)(a * (dy1 / yn) * (double)P.radius);
                x2 = (float)(a * (dy2 / yn) * (double)P.radius);
            }
            const float xn32 = sqrtf(x0 * x0 + x1 * x1 + x2 * x2), yn32 = sqrtf(y0 * y0 + y1 * y1 + y2 * y2);
            float ld = 0.f;
            if (yn32 > 1e-7f) {
                const float q = yn32 / P.radius;
                ld = 2.f * (logf(yn32) - logf(xn32)) + log1pf(-(q * q));
            }
            lp = lp + (0.0f - ld);
            s_lp[col] = lp;
            s_mask[col] = mask;
            sm[L::Zs + col] = x0; sm[L::Zs + NR + col] = x1; sm[L::Zs + 2 * NR + col] = x2;
            sm[L::Cs + CTX * NR + col] = x0;   // conditioning coordinate of the last coupling
        }
    }
    __syncthreads();
    for (int t = P.T - 1; t >= 0; --t) {
        coupling_nn<NR>(P.pack + P.off_nn[j][t], sm);
        spline_knots<NR>(sm + L::Raw, P.radius, sm + L::Ha);
        if (tid < 2 * NR) {
            const int s = tid >> 1, d = tid & 1;
            float ld;
            const float y = sm[L::Zs + (1 + d) * NR + s];
            sm[L::Hb + d * NR + s] = spline_inverse(sm + L::Raw + s, NR, d, y, P.radius, sm + L::Ha + (s * 2 + d) * 18, &ld);
            sm[L::Hc + d * NR + s] = ld;
        }
        __syncthreads();
        if (tid < NR) {
            s_lp[tid] = s_lp[tid] - (sm[L::Hc + tid] + sm[L::Hc + NR + tid]);
            const float x0 = sm[L::Zs + tid], x1 = sm[L::Hb + tid], x2 = sm[L::Hb + NR + tid];
            if (t > 0) {   // undo Permute [1,2,0]: current = prev[[1,2,0]]  =>  prev = (cur[2], cur[0], cur[1])
                sm[L::Zs + tid] = x2; sm[L::Zs + NR + tid] = x0; sm[L::Zs + 2 * NR + tid] = x1;
                sm[L::Cs + CTX * NR + tid] = x2;
            } else {
                sm[L::Zs + tid] = x0; sm[L::Zs + NR + tid] = x1; sm[L::Zs + 2 * NR + tid] = x2;
            }
        }
        __syncthreads();
    }
    if (tid < RT) {
        const int r = r0 + tid;
        const float sc = P.base_std;
        const float log_scale = logf(sc), half_log_2pi = 0.91893853320467274178f;
        float terms[NC];
        for (int c = 0; c < NC; ++c) {
            const int col = c * RT + tid;
            float b = 0.f;
            for (int e = 0; e < 3; ++e) {
                const float x = sm[L::Zs + e * NR + col];
                b += -(x * x) / (2.f * sc * sc) - log_scale - half_log_2pi;
            }
            terms[c] = s_mask[col] ? (s_lp[col] + b) : -INFINITY;
        }
        float res = terms[0];
        if (!ALGEBRA) {
            float m = fmaxf(terms[0], fmaxf(terms[NC > 1 ? 1 : 0], terms[NC > 2 ? 2 : 0]));
            if (isinf(m)) m = 0.f;     // torch.logsumexp: an infinite maximum is not subtracted (keeps +inf, no NaN)
            float ssum = 0.f;
            for (int c = 0; c < NC; ++c) ssum += expf(terms[c] - m);
            res = m + logf(ssum);
        }
        if (r < R) out[(size_t)r * joint_count + jj] = res;
    }
}

}  // namespace

struct hf_flow {
    hf_flow_config cfg;
    FlowParams P;
    float* pack;
    float* betaW;
    float* wfeat;
    float* jpack;
    LevelSched sched;             // level-parallel sampling kernel: joints per (round, group)
    float* wsplit;                // [J*64][512] = [tf32 hi (256) | lo (256)] of the feature part of every context Linear
    CUtensorMap mapW;
    const void* mapF_ptr; int mapF_R; CUtensorMap mapF;     // cached activation map (workspace pointer and row count)
};

namespace {

template <typename K>
int set_smem(K kernel, size_t bytes) {
    HF_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return HF_OK;
}

size_t flow_ws_u_bytes(const hf_flow* h, int R, int nr) {
    return (((size_t)hf::div_up(R, nr) * h->P.J * CTX * nr * sizeof(float)) + 255) & ~(size_t)255;
}

int pick_rows(int R) {
    if (R <= 148 * 8) return 8;
    if (R <= 148 * 16) return 16;
    return 24;     // larger batches run as several waves of 24-row CTAs (shared memory is full at 24 rows)
}

}  // namespace

extern "C" int hf_flow_create(hf_flow_t** out, const hf_flow_config* cfg, const int* ancestors,
                              const int* anc_offsets, const float* beta_weight, const float* const* ctx_weight,
                              const float* const* ctx_bias, const float* const* nn_weight,
                              const float* const* nn_bias) {
    if (!out || !cfg) return hf::fail(HF_ERR_INVALID, "hf_flow_create: null argument");
    if (cfg->num_joints < 1 || cfg->num_joints > HF_FJ || cfg->feats_dim != FEATS || cfg->context_dim != CTX ||
        cfg->num_transforms < 1 || cfg->num_transforms > 2 || cfg->hidden[0] != H1 || cfg->hidden[1] != H2 ||
        cfg->hidden[2] != H3 || cfg->num_bins != NBINS || cfg->num_betas < 1 || cfg->num_betas > 16)
        return hf::fail(HF_ERR_UNSUPPORTED,
                        "hf_flow_create: kernels are specialised to feats 256, context 64, hidden [64,32,32], 8 bins, "
                        "<=2 spline couplings, <=23 joints (got feats %d ctx %d T %d hidden [%d,%d,%d] bins %d J %d)",
                        cfg->feats_dim, cfg->context_dim, cfg->num_transforms, cfg->hidden[0], cfg->hidden[1],
                        cfg->hidden[2], cfg->num_bins, cfg->num_joints);
    hf_flow* h = new hf_flow();
    h->cfg = *cfg;
    FlowParams& P = h->P;
    P.J = cfg->num_joints; P.nb = cfg->num_betas; P.T = cfg->num_transforms;
    P.radius = cfg->radius; P.base_std = cfg->base_std;
    std::vector<float> pack;
    for (int j = 0; j < P.J; ++j) {
        const int a = anc_offsets[j + 1] - anc_offsets[j];
        if (a < 0 || a > HF_FANC) { delete h; return hf::fail(HF_ERR_INVALID, "hf_flow_create: joint %d has %d ancestors", j, a); }
        P.anc_cnt[j] = a;
        for (int q = 0; q < a; ++q) {
            int an = ancestors[anc_offsets[j] + q];
            if (an < 0 || an >= j) { delete h; return hf::fail(HF_ERR_INVALID, "hf_flow_create: ancestor %d of joint %d is not an earlier joint", an, j); }
            P.anc[j][q] = (signed char)an;
        }
        const int Kc = FEATS + 9 * a;
        P.off_ctxW[j] = (int)pack.size();
        pack.resize(pack.size() + (size_t)Kc * CTXP, 0.f);
        for (int k = 0; k < Kc; ++k)
            for (int o = 0; o < CTX; ++o) pack[P.off_ctxW[j] + k * CTXP + o] = ctx_weight[j][(size_t)o * Kc + k];
        P.off_ctxB[j] = (int)pack.size();
        pack.insert(pack.end(), ctx_bias[j], ctx_bias[j] + CTX);
        for (int t = 0; t < P.T; ++t) {
            const int base = (int)pack.size();
            P.off_nn[j][t] = base;
            pack.resize(pack.size() + COUPLING_FLOATS, 0.f);
            const int li = (j * P.T + t) * 4;
            const int Ks[4] = {CTX + 1, H1, H2, H3}, Os[4] = {H1, H2, H3, NRAW}, Op[4] = {64, 32, 32, 64};
            const int offW[4] = {OFF_W0, OFF_W1, OFF_W2, OFF_W3}, offB[4] = {OFF_B0, OFF_B1, OFF_B2, OFF_B3};
            for (int l = 0; l < 4; ++l) {
                for (int k = 0; k < Ks[l]; ++k)
                    for (int o = 0; o < Os[l]; ++o)
                        pack[base + offW[l] + k * (Op[l] + WPAD) + o] = nn_weight[li + l][(size_t)o * Ks[l] + k];
                for (int o = 0; o < Os[l]; ++o) pack[base + offB[l] + o] = nn_bias[li + l][o];
            }
        }
    }
    // sampling-kernel packing: Wfeat [FEATS][J*CTX] and per-joint blocks [ancW | ctxB | coupling 0 | coupling 1]
    std::vector<float> wfeat((size_t)FEATS * P.J * CTX), jpack;
    for (int j = 0; j < P.J; ++j) {
        const int a = P.anc_cnt[j], Kc = FEATS + 9 * a;
        if (a > 7) { delete h; return hf::fail(HF_ERR_UNSUPPORTED, "hf_flow_create: joint %d has %d ancestors (sampling kernel stages at most 7)", j, a); }
        for (int k = 0; k < FEATS; ++k)
            for (int o = 0; o < CTX; ++o) wfeat[(size_t)k * P.J * CTX + j * CTX + o] = ctx_weight[j][(size_t)o * Kc + k];
        P.off_jb[j] = (int)jpack.size();
        for (int k = 0; k < 9 * a; ++k)
            for (int o = 0; o < CTXP; ++o) jpack.push_back(o < CTX ? ctx_weight[j][(size_t)o * Kc + FEATS + k] : 0.f);
        jpack.insert(jpack.end(), ctx_bias[j], ctx_bias[j] + CTX);
        for (int t = 0; t < 2; ++t) {
            if (t < P.T) jpack.insert(jpack.end(), pack.begin() + P.off_nn[j][t], pack.begin() + P.off_nn[j][t] + COUPLING_FLOATS);
            else jpack.resize(jpack.size() + COUPLING_FLOATS, 0.f);
        }
    }
    {   // static list schedule of the joints on LG workers: a joint is ready when all its ancestors are done (its context reads
        // their rotations); among the ready joints the one with the longest chain of descendants goes first.  Unit joint cost.
        // (SMPL: 8 steps for 23 joints on 3 workers; the kernel's groups follow their lists and wait on per-joint flags.)
        int height[HF_FJ], done_at[HF_FJ];
        for (int j = P.J - 1; j >= 0; --j) {
            height[j] = 1;
            for (int c = j + 1; c < P.J; ++c)
                for (int q = 0; q < P.anc_cnt[c]; ++q)
                    if (P.anc[c][q] == j) height[j] = std::max(height[j], height[c] + 1);
            done_at[j] = -1;
        }
        LevelSched& S = h->sched;
        for (int g = 0; g < LG; ++g) S.count[g] = 0;
        int left = P.J;
        for (int step = 0; left > 0; ++step) {
            bool taken[HF_FJ] = {};
            for (int g = 0; g < LG; ++g) {
                int best = -1;
                for (int j = 0; j < P.J; ++j) {
                    if (done_at[j] >= 0 || taken[j]) continue;
                    bool ready = true;
                    for (int q = 0; q < P.anc_cnt[j]; ++q) { const int a = P.anc[j][q]; if (done_at[a] < 0 || done_at[a] > step) ready = false; }
                    if (ready && (best < 0 || height[j] > height[best])) best = j;
                }
                if (best < 0) continue;
                taken[best] = true;
                S.joint[g][S.count[g]++] = (signed char)best;
            }
            for (int j = 0; j < P.J; ++j) if (taken[j]) { done_at[j] = step + 1; --left; }
        }
    }
    std::vector<float> wsplit((size_t)P.J * CTX * 2 * FEATS);
    for (int j = 0; j < P.J; ++j) {
        const int Kc = FEATS + 9 * P.anc_cnt[j];
        for (int o = 0; o < CTX; ++o)
            for (int k = 0; k < FEATS; ++k) {
                const float w = ctx_weight[j][(size_t)o * Kc + k];
                uint32_t u;
                memcpy(&u, &w, 4);
                u = (u + 0x1000u) & 0xffffe000u;            // tf32: round to nearest (ties away), as cvt.rna.tf32.f32
                float hi;
                memcpy(&hi, &u, 4);
                if (!std::isfinite(hi)) hi = w;
                wsplit[((size_t)j * CTX + o) * 2 * FEATS + k] = hi;
                wsplit[((size_t)j * CTX + o) * 2 * FEATS + FEATS + k] = w - hi;
            }
    }
    std::vector<float> bw((size_t)P.nb * FEATS);
    for (int l = 0; l < P.nb; ++l)
        for (int o = 0; o < FEATS; ++o) bw[l * FEATS + o] = beta_weight[(size_t)o * P.nb + l];
    int rc;
    if ((rc = hf::upload(&h->pack, pack.data(), pack.size()))) return rc;
    if ((rc = hf::upload(&h->betaW, bw.data(), bw.size()))) return rc;
    if ((rc = hf::upload(&h->wfeat, wfeat.data(), wfeat.size()))) return rc;
    if ((rc = hf::upload(&h->jpack, jpack.data(), jpack.size()))) return rc;
    if ((rc = hf::upload(&h->wsplit, wsplit.data(), wsplit.size()))) return rc;
    {
        const uint64_t dims[2] = {(uint64_t)(2 * FEATS), (uint64_t)(P.J * CTX)}, strides[1] = {(uint64_t)(2 * FEATS * 4)};
        const uint32_t box[2] = {32, (uint32_t)CG_BN};
        if ((rc = encode_map(&h->mapW, h->wsplit, 2, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32))) return rc;
    }
    h->mapF_ptr = nullptr; h->mapF_R = 0;
    P.pack = h->pack; P.betaW = h->betaW; P.wfeat = h->wfeat; P.jpack = h->jpack;
    *out = h;
    return HF_OK;
}

extern "C" void hf_flow_destroy(hf_flow_t* h) {
    if (!h) return;
    cudaFree(h->pack); cudaFree(h->betaW); cudaFree(h->wfeat); cudaFree(h->jpack); cudaFree(h->wsplit);
    delete h;
}

#define HF_DISPATCH_ROWS(NRv, ...)                            \
    switch (NRv) {                                            \
        case 8:  { constexpr int NR = 8;  __VA_ARGS__; } break; \
        case 16: { constexpr int NR = 16; __VA_ARGS__; } break; \
        default: { constexpr int NR = 24; __VA_ARGS__; } break; \
    }

extern "C" size_t hf_flow_workspace_bytes(const hf_flow_t* h, int R) {
    if (!h || R <= 0) return 0;
    const int nr = pick_rows(R);
    return flow_ws_u_bytes(h, R, nr) + (size_t)R * 2 * FEATS * sizeof(float);     // U | F = [R][hi | lo]
}

extern "C" int hf_flow_sample(const hf_flow_t* h, const float* img_base, const float* betas, const int* img_index,
                              const float* base_noise, int R, int Rn, float* rotmats, float* axisangle_pe,
                              void* workspace, size_t workspace_bytes, void* stream) {
    if (!h || !img_base || !betas || !img_index || !rotmats) return hf::fail(HF_ERR_INVALID, "hf_flow_sample: null argument");
    if (Rn < 0 || Rn > R || (Rn > 0 && !base_noise)) return hf::fail(HF_ERR_INVALID, "hf_flow_sample: bad row split R=%d Rn=%d", R, Rn);
    if (R <= 0) return HF_OK;
    if (!workspace || workspace_bytes < hf_flow_workspace_bytes(h, R) || ((uintptr_t)workspace & 15))
        return hf::fail(HF_ERR_INVALID, "hf_flow_sample: workspace too small or unaligned (%zu < %zu)", workspace_bytes, hf_flow_workspace_bytes(h, R));
    const int nr = pick_rows(R);
    const int have_U = getenv("HF_FLOW_SIMT_PROLOGUE") ? 0 : 1;
    if (have_U) {
        // context prologue on the tensor cores: features (split tf32) -> U for all joints
        float* F = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + flow_ws_u_bytes(h, R, nr));
        hf_flow* hm = const_cast<hf_flow*>(h);
        if (hm->mapF_ptr != F || hm->mapF_R != R) {
            const uint64_t dims[2] = {(uint64_t)(2 * FEATS), (uint64_t)R}, strides[1] = {(uint64_t)(2 * FEATS * 4)};
            const uint32_t box[2] = {32, 128};
            int rc = encode_map(&hm->mapF, F, 2, dims, strides, box, CU_TENSOR_MAP_DATA_TYPE_FLOAT32);
            if (rc) return rc;
            hm->mapF_ptr = F; hm->mapF_R = R;
        }
        HF_CUDA(hf::launch_pdl(flow_feats_split_kernel, dim3(hf::div_up(R * (FEATS / 4), 256)), dim3(256), 0, (cudaStream_t)stream, h->P, img_base, betas,
                               img_index, R, F));
        HF_LAUNCH_CHECK();
        const size_t gsmem = CG_STAGES * CG_STAGE_BYTES + 1024;
        int rc = set_smem(flow_ctx_gemm_kernel, gsmem);
        if (rc) return rc;
        HF_CUDA(hf::launch_pdl(flow_ctx_gemm_kernel, dim3(hf::div_up(R, 128), hf::div_up(h->P.J * CTX, CG_BN)), dim3(CG_THREADS), gsmem, (cudaStream_t)stream, h->mapF, h->mapW, R,
                               h->P.J, nr, (float*)workspace));
        HF_LAUNCH_CHECK();
    }
    const bool levels = have_U && getenv("HF_FLOW_CHAIN") == nullptr;      // HF_FLOW_CHAIN=1: the joint-by-joint kernel (cross-check)
    if (levels) {
        HF_DISPATCH_ROWS(nr, {
            const size_t smem = LevelSmem<NR>::Total * sizeof(float);
            int rc = set_smem(flow_sample_levels_kernel<NR>, smem);
            if (rc) return rc;
            HF_CUDA(hf::launch_pdl(flow_sample_levels_kernel<NR>, dim3(hf::div_up(R, NR)), dim3(LNT + 32 * LG), smem, (cudaStream_t)stream, h->P, h->sched, base_noise,
                                   R, Rn, rotmats, axisangle_pe, (const float*)workspace, getenv("HF_FLOW_DBG") ? atoi(getenv("HF_FLOW_DBG")) : 0));
        });
        HF_LAUNCH_CHECK();
        return HF_OK;
    }
    HF_DISPATCH_ROWS(nr, {
        const size_t smem = SampleSmem<NR>::Total * sizeof(float);
        int rc = set_smem(flow_sample_kernel<NR>, smem);
        if (rc) return rc;
        HF_CUDA(hf::launch_pdl(flow_sample_kernel<NR>, dim3(hf::div_up(R, NR)), dim3(HF_NT), smem, (cudaStream_t)stream, h->P, img_base, betas,
                               img_index, base_noise, R, Rn, rotmats, axisangle_pe, (float*)workspace, have_U, getenv("HF_FLOW_DBG") ? 1 : 0));
    });
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" int hf_flow_context(const hf_flow_t* h, const float* img_base, const float* betas, const int* img_index,
                               const float* anc_rotmats, int R, float* ctx_out, void* stream) {
    if (!h || !img_base || !betas || !img_index || !anc_rotmats || !ctx_out) return hf::fail(HF_ERR_INVALID, "hf_flow_context: null argument");
    if (R <= 0) return HF_OK;
    constexpr int NR = 8;
    const size_t smem = SmemLayout<NR>::Total * sizeof(float);
    int rc = set_smem(flow_context_kernel<NR>, smem);
    if (rc) return rc;
    dim3 grid(hf::div_up(R, NR), h->P.J);
    flow_context_kernel<NR><<<grid, HF_NT, smem, (cudaStream_t)stream>>>(h->P, img_base, betas, img_index, anc_rotmats, R, ctx_out);
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" int hf_flow_log_prob(const hf_flow_t* h, const float* ctx, int ctx_row_stride, int joint_first,
                                int joint_count, const double* rot_f64, int R, float* out, void* stream) {
    if (!h || !ctx || !rot_f64 || !out) return hf::fail(HF_ERR_INVALID, "hf_flow_log_prob: null argument");
    if (joint_first < 0 || joint_count < 1 || joint_first + joint_count > h->P.J)
        return hf::fail(HF_ERR_INVALID, "hf_flow_log_prob: joints [%d,%d) out of range", joint_first, joint_first + joint_count);
    if (R <= 0) return HF_OK;
    constexpr int NR = 24;
    const size_t smem = SmemLayout<NR>::Total * sizeof(float);
    int rc = set_smem(flow_logprob_kernel<NR, false>, smem);
    if (rc) return rc;
    dim3 grid(hf::div_up(R, NR / 3), joint_count);
    flow_logprob_kernel<NR, false><<<grid, HF_NT, smem, (cudaStream_t)stream>>>(h->P, ctx, ctx_row_stride, joint_first,
                                                                              joint_count, rot_f64, nullptr, R, out);
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" int hf_flow_algebra_log_prob(const hf_flow_t* h, const float* ctx, int ctx_row_stride, int joint_first,
                                        int joint_count, const float* v, int R, float* out, void* stream) {
    if (!h || !ctx || !v || !out) return hf::fail(HF_ERR_INVALID, "hf_flow_algebra_log_prob: null argument");
    if (joint_first < 0 || joint_count < 1 || joint_first + joint_count > h->P.J)
        return hf::fail(HF_ERR_INVALID, "hf_flow_algebra_log_prob: joints [%d,%d) out of range", joint_first, joint_first + joint_count);
    if (R <= 0) return HF_OK;
    constexpr int NR = 8;
    const size_t smem = SmemLayout<NR>::Total * sizeof(float);
    int rc = set_smem(flow_logprob_kernel<NR, true>, smem);
    if (rc) return rc;
    dim3 grid(hf::div_up(R, NR), joint_count);
    flow_logprob_kernel<NR, true><<<grid, HF_NT, smem, (cudaStream_t)stream>>>(h->P, ctx, ctx_row_stride, joint_first,
                                                                             joint_count, nullptr, v, R, out);
    HF_LAUNCH_CHECK();
    return HF_OK;
}


// =====================================================================================================================
// Backward of the log-densities w.r.t. their INPUTS: d log_prob / d (context row, target rotation or algebra
// vector), weights fixed -- what a fitting loop needs from the pose prior (optimise/optimise_humaniflow.py:96-114: the loss is
// -sum_j log p_j(R_j | ctx_j), and ctx_j depends on the ancestors' rotations, the shape and the global rotation).  The reference gets
// it from torch.autograd through pyro's SplineCoupling / ConditionalDenseNN and local_diffeo_transformed_distribution.py:84-142.
// Correctness-first form: ONE THREAD per (row, joint) recomputes the forward chain with all activations in local memory and
// walks it backwards (these calls see a few dozen rows, not thousands); every formula is the derivative of the forward code in
// flow_logprob_kernel / spline_inverse above.  fp32 like the forward, fp64 for the exp / log-map parts.
namespace {

struct CplActs { float in[CTX + 1], h1[64], h2[32], h3[32], raw[64]; };

template <int O>
__device__ void ref_dense_fwd(const float* __restrict__ W, const float* __restrict__ b, int K, const float* x, float* y, bool relu) {
    constexpr int LDW = O + WPAD;
    float acc[O];
#pragma unroll
    for (int o = 0; o < O; ++o) acc[o] = __ldg(b + o);
    for (int k = 0; k < K; ++k) {
        const float xv = x[k];
        const float4* wr = reinterpret_cast<const float4*>(W + (size_t)k * LDW);
#pragma unroll
        for (int q = 0; q < O / 4; ++q) {
            const float4 w = __ldg(wr + q);
            acc[q * 4] = fmaf(w.x, xv, acc[q * 4]); acc[q * 4 + 1] = fmaf(w.y, xv, acc[q * 4 + 1]);
            acc[q * 4 + 2] = fmaf(w.z, xv, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(w.w, xv, acc[q * 4 + 3]);
        }
    }
#pragma unroll
    for (int o = 0; o < O; ++o) y[o] = relu ? fmaxf(acc[o], 0.f) : acc[o];
}

// gx[k] = sum_o W[k][o] gy[o]   (gy = gradient w.r.t. the layer's pre-activation)
template <int O>
__device__ void ref_dense_bwd(const float* __restrict__ W, int K, const float* gy, float* gx) {
    constexpr int LDW = O + WPAD;
    for (int k = 0; k < K; ++k) {
        const float4* wr = reinterpret_cast<const float4*>(W + (size_t)k * LDW);
        float a = 0.f;
#pragma unroll
        for (int q = 0; q < O / 4; ++q) {
            const float4 w = __ldg(wr + q);
            a = fmaf(w.x, gy[q * 4], a); a = fmaf(w.y, gy[q * 4 + 1], a); a = fmaf(w.z, gy[q * 4 + 2], a); a = fmaf(w.w, gy[q * 4 + 3], a);
        }
        gx[k] = a;
    }
}

__device__ void ref_mlp_fwd(const float* __restrict__ cw, CplActs& A) {
    ref_dense_fwd<64>(cw + OFF_W0, cw + OFF_B0, CTX + 1, A.in, A.h1, true);
    ref_dense_fwd<32>(cw + OFF_W1, cw + OFF_B1, H1, A.h1, A.h2, true);
    ref_dense_fwd<32>(cw + OFF_W2, cw + OFF_B2, H2, A.h2, A.h3, true);
    ref_dense_fwd<64>(cw + OFF_W3, cw + OFF_B3, H3, A.h3, A.raw, false);
}

// g_raw[64] (entries >= NRAW ignored) -> g_in[CTX + 1]
__device__ void ref_mlp_bwd(const float* __restrict__ cw, const CplActs& A, float* g_raw, float* g_in) {
    float g3[32], g2[32], g1[64];
    g_raw[62] = 0.f; g_raw[63] = 0.f;
    ref_dense_bwd<64>(cw + OFF_W3, H3, g_raw, g3);
    for (int k = 0; k < 32; ++k) g3[k] = A.h3[k] > 0.f ? g3[k] : 0.f;
    ref_dense_bwd<32>(cw + OFF_W2, H2, g3, g2);
    for (int k = 0; k < 32; ++k) g2[k] = A.h2[k] > 0.f ? g2[k] : 0.f;
    ref_dense_bwd<32>(cw + OFF_W1, H1, g2, g1);
    for (int k = 0; k < 64; ++k) g1[k] = A.h1[k] > 0.f ? g1[k] : 0.f;
    ref_dense_bwd<64>(cw + OFF_W0, CTX + 1, g1, g_in);
}

// Inverse rational-linear spline of dimension d for one scalar, from the raw NN outputs (the arithmetic of spline_knots +
// spline_inverse); with g_raw != nullptr also the backward: upstream gx (on the returned x) and gld (on *ld) -> *gy += d/dy,
// g_raw[...] += d/d(raw outputs).
__device__ float ref_spline_inverse(const float* raw, int d, float y, float bound, float* ld, float* g_raw, float gx, float gld, float* gy) {
    const float lo = -bound, hi = bound;
    if (!(y >= -bound && y <= bound)) { *ld = 0.f; if (g_raw) *gy += gx; return y; }
    float sm[2][NBINS], kn[2][NBINS + 1];            // [widths / heights]
    for (int v = 0; v < 2; ++v) {
        const float* r = raw + v * 2 * NBINS + d * NBINS;
        float m = r[0];
        for (int b = 1; b < NBINS; ++b) m = fmaxf(m, r[b]);
        float ssum = 0.f;
        for (int b = 0; b < NBINS; ++b) { sm[v][b] = expf(r[b] - m); ssum += sm[v][b]; }
        float cum = 0.f;
        kn[v][0] = lo;
        for (int b = 0; b < NBINS; ++b) {
            sm[v][b] /= ssum;
            cum += 1e-3f + 0.992f * sm[v][b];
            kn[v][b + 1] = (b == NBINS - 1) ? hi : (hi - lo) * cum + lo;
        }
    }
    int idx = -1;
    for (int b = 0; b <= NBINS; ++b) idx += (y >= kn[1][b] + 1e-6f) ? 1 : 0;
    idx = max(0, min(idx, NBINS - 1));
    const float in_cw = kn[0][idx], in_w = kn[0][idx + 1] - kn[0][idx], in_ch = kn[1][idx], in_h = kn[1][idx + 1] - kn[1][idx];
    auto deriv_raw = [&](int i) { return raw[4 * NBINS + d * (NBINS - 1) + (i - 1)]; };
    auto deriv = [&](int i) -> float {
        if (i == 0 || i == NBINS) return 0.999f;
        const float u = deriv_raw(i);
        return 1e-3f + ((u > 20.f) ? u : log1pf(expf(u)));
    };
    const float d0 = deriv(idx), d1 = deriv(idx + 1);
    const float lraw = raw[4 * NBINS + 2 * (NBINS - 1) + d * NBINS + idx];
    const float sg = 1.f / (1.f + expf(-lraw));
    const float lam = 0.95f * sg + 0.025f;
    const float delta = in_h / in_w;
    const float wb = sqrtf(d0 / d1);
    const float C = lam * d0 + (1.f - lam) * wb * d1;
    const float wc = C / delta;
    const float ya = in_ch, yb = in_h + in_ch;
    const float Aq = (1.f - lam) * ya + lam * wb * yb, Bq = (1.f - lam) + lam * wb;
    const float yc = Aq / Bq;
    const bool left = y <= yc;
    float num, den, dnum;
    if (left) {
        num = lam * (ya - y);
        den = (wc - 1.f) * y + ya - wc * yc;
        dnum = wc * lam * (yc - ya) * in_w;
    } else {
        num = (wc - lam * wb) * y + lam * wb * yb - wc * yc;
        den = (wc - wb) * y + wb * yb - wc * yc;
        dnum = wb * wc * (1.f - lam) * (yb - yc) * in_w;
    }
    const float theta = num / den;
    *ld = -(logf(dnum) - 2.f * logf(fabsf(den)));
    const float x = theta * in_w + in_cw;
    if (!g_raw) return x;
    // ---- backward ----
    float g_in_w = gx * theta, g_in_cw = gx, g_in_h = 0.f, g_in_ch = 0.f;
    const float g_theta = gx * in_w;
    const float g_num = g_theta / den;
    const float g_den = -g_theta * num / (den * den) + gld * 2.f / den;
    const float g_dnum = -gld / dnum;
    float g_lam = 0.f, g_ya = 0.f, g_yb = 0.f, g_yc = 0.f, g_wc = 0.f, g_wb = 0.f, g_y = 0.f;
    if (left) {
        g_lam += g_num * (ya - y); g_ya += g_num * lam; g_y -= g_num * lam;
        g_wc += g_den * (y - yc); g_y += g_den * (wc - 1.f); g_ya += g_den; g_yc -= g_den * wc;
        const float e = (yc - ya) * in_w;
        g_wc += g_dnum * lam * e; g_lam += g_dnum * wc * e;
        g_yc += g_dnum * wc * lam * in_w; g_ya -= g_dnum * wc * lam * in_w; g_in_w += g_dnum * wc * lam * (yc - ya);
    } else {
        g_wc += g_num * (y - yc); g_lam += g_num * wb * (yb - y); g_wb += g_num * lam * (yb - y);
        g_y += g_num * (wc - lam * wb); g_yb += g_num * lam * wb; g_yc -= g_num * wc;
        g_wc += g_den * (y - yc); g_wb += g_den * (yb - y); g_y += g_den * (wc - wb); g_yb += g_den * wb; g_yc -= g_den * wc;
        const float e = (yb - yc) * in_w;
        g_wb += g_dnum * wc * (1.f - lam) * e; g_wc += g_dnum * wb * (1.f - lam) * e; g_lam -= g_dnum * wb * wc * e;
        g_yb += g_dnum * wb * wc * (1.f - lam) * in_w; g_yc -= g_dnum * wb * wc * (1.f - lam) * in_w;
        g_in_w += g_dnum * wb * wc * (1.f - lam) * (yb - yc);
    }
    {   // yc = Aq / Bq
        const float g_A = g_yc / Bq, g_B = -g_yc * Aq / (Bq * Bq);
        g_lam += g_A * (wb * yb - ya) + g_B * (wb - 1.f);
        g_ya += g_A * (1.f - lam);
        g_wb += g_A * lam * yb + g_B * lam;
        g_yb += g_A * lam * wb;
    }
    g_in_ch += g_ya + g_yb; g_in_h += g_yb;
    float g_d0 = 0.f, g_d1 = 0.f;
    {   // wc = C / delta, C = lam d0 + (1 - lam) wb d1
        const float g_C = g_wc / delta, g_delta = -g_wc * C / (delta * delta);
        g_lam += g_C * (d0 - wb * d1); g_d0 += g_C * lam; g_wb += g_C * (1.f - lam) * d1; g_d1 += g_C * (1.f - lam) * wb;
        g_in_h += g_delta / in_w; g_in_w -= g_delta * in_h / (in_w * in_w);
    }
    g_d0 += g_wb * 0.5f * wb / d0; g_d1 -= g_wb * 0.5f * wb / d1;
    *gy += g_y;
    // lambda, derivatives
    g_raw[4 * NBINS + 2 * (NBINS - 1) + d * NBINS + idx] += g_lam * 0.95f * sg * (1.f - sg);
    if (idx >= 1) { const float u = deriv_raw(idx); g_raw[4 * NBINS + d * (NBINS - 1) + idx - 1] += g_d0 * ((u > 20.f) ? 1.f : 1.f / (1.f + expf(-u))); }
    if (idx + 1 <= NBINS - 1) { const float u = deriv_raw(idx + 1); g_raw[4 * NBINS + d * (NBINS - 1) + idx] += g_d1 * ((u > 20.f) ? 1.f : 1.f / (1.f + expf(-u))); }
    // knots: in_cw = kw[idx], in_w = kw[idx+1] - kw[idx] (heights alike); kn[i] = (hi - lo) cum_i + lo for 1 <= i <= 7, ends pinned
    for (int v = 0; v < 2; ++v) {
        const float g_lo_knot = v == 0 ? (g_in_cw - g_in_w) : (g_in_ch - g_in_h), g_hi_knot = v == 0 ? g_in_w : g_in_h;
        float g_c[NBINS];                // gradient w.r.t. the bin lengths c_b = 1e-3 + 0.992 softmax_b
        for (int b = 0; b < NBINS; ++b) {
            float a = 0.f;               // c_b enters every knot i > b (i <= 7)
            if (idx >= 1 && idx <= NBINS - 1 && idx > b) a += g_lo_knot;
            if (idx + 1 <= NBINS - 1 && idx + 1 > b) a += g_hi_knot;
            g_c[b] = a * (hi - lo) * 0.992f;
        }
        float dot = 0.f;
        for (int b = 0; b < NBINS; ++b) dot += sm[v][b] * g_c[b];
        for (int b = 0; b < NBINS; ++b) g_raw[v * 2 * NBINS + d * NBINS + b] += sm[v][b] * (g_c[b] - dot);
    }
    return x;
}

// Algebra-space log-density of joint j for one vector y (fp32), context ctx[CTX]; returns lp.  With g != 0 also accumulates
// g * d lp / d ctx into g_ctx[CTX] and writes g * d lp / d y into g_y[3].
__device__ float ref_algebra_lp(const FlowParams& P, int j, const float* ctx, const float* y, float g, float* g_ctx, float* g_y) {
    const float rad = P.radius;
    // radial tanh inverse (fp64 like the forward)
    const double dy0 = y[0], dy1 = y[1], dy2 = y[2];
    const double yn = sqrt(dy0 * dy0 + dy1 * dy1 + dy2 * dy2);
    float z[3] = {y[0], y[1], y[2]};
    double a_th = 0.0;
    if (yn > 1e-7) {
        a_th = atanh(yn / (double)rad);
        z[0] = (float)(a_th * (dy0 / yn) * (double)rad); z[1] = (float)(a_th * (dy1 / yn) * (double)rad); z[2] = (float)(a_th * (dy2 / yn) * (double)rad);
    }
    const float xn32 = sqrtf(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]), yn32 = sqrtf(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
    float lp = 0.f;
    if (yn32 > 1e-7f) {
        const float q = yn32 / rad;
        lp -= 2.f * (logf(yn32) - logf(xn32)) + log1pf(-(q * q));
    }
    CplActs A[2];
    float zin[2][3];                  // state entering stage i (stage i handles coupling t = T-1-i)
    for (int i = 0; i < P.T; ++i) {
        const int t = P.T - 1 - i;
        zin[i][0] = z[0]; zin[i][1] = z[1]; zin[i][2] = z[2];
        for (int k = 0; k < CTX; ++k) A[i].in[k] = ctx[k];
        A[i].in[CTX] = z[0];
        ref_mlp_fwd(P.pack + P.off_nn[j][t], A[i]);
        float ld0, ld1;
        const float x1 = ref_spline_inverse(A[i].raw, 0, z[1], rad, &ld0, nullptr, 0.f, 0.f, nullptr);
        const float x2 = ref_spline_inverse(A[i].raw, 1, z[2], rad, &ld1, nullptr, 0.f, 0.f, nullptr);
        lp -= ld0 + ld1;
        const float x0 = z[0];
        if (t > 0) { z[0] = x2; z[1] = x0; z[2] = x1; } else { z[0] = x0; z[1] = x1; z[2] = x2; }
    }
    const float sc = P.base_std;
    for (int e = 0; e < 3; ++e) lp += -(z[e] * z[e]) / (2.f * sc * sc) - logf(sc) - 0.91893853320467274178f;
    if (g == 0.f) return lp;
    // ---- backward ----
    float gz[3];
    for (int e = 0; e < 3; ++e) gz[e] = g * (-z[e] / (sc * sc));
    for (int i = P.T - 1; i >= 0; --i) {
        const int t = P.T - 1 - i;
        float gx0, gx1, gx2;
        if (t > 0) { gx2 = gz[0]; gx0 = gz[1]; gx1 = gz[2]; } else { gx0 = gz[0]; gx1 = gz[1]; gx2 = gz[2]; }
        float g_raw[64];
        for (int k = 0; k < 64; ++k) g_raw[k] = 0.f;
        float gz1 = 0.f, gz2 = 0.f, ldd;
        ref_spline_inverse(A[i].raw, 0, zin[i][1], rad, &ldd, g_raw, gx1, -g, &gz1);
        ref_spline_inverse(A[i].raw, 1, zin[i][2], rad, &ldd, g_raw, gx2, -g, &gz2);
        float g_in[CTX + 1];
        ref_mlp_bwd(P.pack + P.off_nn[j][t], A[i], g_raw, g_in);
        for (int k = 0; k < CTX; ++k) g_ctx[k] += g_in[k];
        gz[0] = gx0 + g_in[CTX]; gz[1] = gz1; gz[2] = gz2;
    }
    // radial tanh inverse: z = f(n) y / n, f = r atanh(n / r); lp term -ld(n), ld = 2 log n - 2 log f + log(1 - (n/r)^2)
    if (yn > 1e-7) {
        const double r = rad, s = yn / r, f = r * a_th, fp = 1.0 / (1.0 - s * s);
        const double yh[3] = {dy0 / yn, dy1 / yn, dy2 / yn};
        const double dotg = yh[0] * gz[0] + yh[1] * gz[1] + yh[2] * gz[2];
        const double dld = 2.0 / yn - 2.0 * fp / f - 2.0 * s * fp / r;
        for (int e = 0; e < 3; ++e) g_y[e] = (float)((f / yn) * gz[e] + (fp - f / yn) * dotg * yh[e] - (double)g * dld * yh[e]);
    } else {
        for (int e = 0; e < 3; ++e) g_y[e] = gz[e];
    }
    return lp;
}

template <bool ALGEBRA>
__global__ void __launch_bounds__(64)
flow_logprob_bwd_kernel(const __grid_constant__ FlowParams P, const float* __restrict__ ctx, int ctx_row_stride, int joint_first, int joint_count,
                        const double* __restrict__ rot, const float* __restrict__ valg, const float* __restrict__ g_out, int R,
                        float* __restrict__ g_ctx_out, double* __restrict__ g_rot, float* __restrict__ g_valg) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x, jj = blockIdx.y, j = joint_first + jj;
    if (r >= R) return;
    const double PI = 3.14159265358979323846;
    float c[CTX], gc[CTX];
    for (int k = 0; k < CTX; ++k) { c[k] = ctx[(size_t)r * ctx_row_stride + j * CTX + k]; gc[k] = 0.f; }
    const float g = g_out[(size_t)r * joint_count + jj];
    if (ALGEBRA) {
        const float* v = valg + ((size_t)r * joint_count + jj) * 3;
        const float y[3] = {v[0], v[1], v[2]};
        float gy[3] = {0.f, 0.f, 0.f};
        if (g != 0.f) ref_algebra_lp(P, j, c, y, g, gc, gy);
        for (int e = 0; e < 3; ++e) g_valg[((size_t)r * joint_count + jj) * 3 + e] = gy[e];
    } else {
        double rm[9], x[3];
        const double* rp = rot + ((size_t)r * joint_count + jj) * 9;
        for (int e = 0; e < 9; ++e) rm[e] = rp[e];
        so3_log_f64(rm, x);
        const double n = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
        double vc[3][3];
        int mask[3] = {1, 0, 0};
        float term[3];
        for (int cnd = 0; cnd < 3; ++cnd) {
            if (cnd > 0) {   // the arithmetic of flow_logprob_kernel (the mask is a comparison at the edge of the support)
                const double fk = n + 2.0 * PI * (cnd == 1 ? -1.0 : 1.0);
                double v[3] = {x[0] / n * fk, x[1] / n * fk, x[2] / n * fk};
                const double vn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
                mask[cnd] = (vn < (double)P.radius) ? 1 : 0;          // NaN (n == 0) compares false
                for (int e = 0; e < 3; ++e) vc[cnd][e] = mask[cnd] ? v[e] : 0.0;
            } else {
                for (int e = 0; e < 3; ++e) vc[cnd][e] = x[e];
            }
            term[cnd] = -INFINITY;
            if (mask[cnd]) {
                const float y[3] = {(float)vc[cnd][0], (float)vc[cnd][1], (float)vc[cnd][2]};
                term[cnd] = -(float)so3_log_abs_det(vc[cnd][0], vc[cnd][1], vc[cnd][2]) + ref_algebra_lp(P, j, c, y, 0.f, nullptr, nullptr);
            }
        }
        float m = fmaxf(term[0], fmaxf(term[1], term[2]));
        if (isinf(m)) m = 0.f;
        float ssum = 0.f;
        for (int cnd = 0; cnd < 3; ++cnd) ssum += expf(term[cnd] - m);
        double gx[3] = {0.0, 0.0, 0.0};
        for (int cnd = 0; cnd < 3; ++cnd) {
            if (!mask[cnd]) continue;
            const float gcnd = g * expf(term[cnd] - m) / ssum;
            if (gcnd == 0.f) continue;
            const float y[3] = {(float)vc[cnd][0], (float)vc[cnd][1], (float)vc[cnd][2]};
            float gy[3] = {0.f, 0.f, 0.f};
            ref_algebra_lp(P, j, c, y, gcnd, gc, gy);
            // -log|det J_exp|(v): L(n) = log((2 - 2 cos n) / n^2), dL/dn = sin n / (1 - cos n) - 2 / n
            const double vn = sqrt(vc[cnd][0] * vc[cnd][0] + vc[cnd][1] * vc[cnd][1] + vc[cnd][2] * vc[cnd][2]);
            double gv[3] = {gy[0], gy[1], gy[2]};
            if (vn > 1e-10) {
                const double dL = sin(vn) / (1.0 - cos(vn)) - 2.0 / vn;
                for (int e = 0; e < 3; ++e) gv[e] -= (double)gcnd * dL * vc[cnd][e] / vn;
            }
            if (cnd == 0) { for (int e = 0; e < 3; ++e) gx[e] += gv[e]; }
            else {   // v = f(n) x, f = 1 + 2 pi k / n
                const double k2 = 2.0 * PI * (cnd == 1 ? -1.0 : 1.0), f = 1.0 + k2 / n, fp = -k2 / (n * n);
                const double dotg = (x[0] * gv[0] + x[1] * gv[1] + x[2] * gv[2]) / n;
                for (int e = 0; e < 3; ++e) gx[e] += f * gv[e] + fp * dotg * x[e];
            }
        }
        // x = so3_log(R): generic branch x = (theta / sin theta) w, w = 0.5 vee(R - R^T), theta = acos((tr R - 1) / 2)
        double gr[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        double cth = 0.5 * (rm[0] + rm[4] + rm[8] - 1.0);
        const bool clamped = cth < -1.0 || cth > 1.0;
        cth = fmin(fmax(cth, -1.0), 1.0);
        const double th = acos(cth);
        if (fabs(PI - th) < 1e-2) {
            // near pi the forward takes a_i = sqrt(max(q_i', 1e-8) / 2) with the sign pattern that reproduces R best
            const double k = th * th / (1.0 - cos(th));
            const double q1 = k * (rm[0] - 1.0), q2 = k * (rm[4] - 1.0), q3 = k * (rm[8] - 1.0);
            const double s1 = q1 - q2 - q3, s2 = -q1 + q2 - q3, s3 = -q1 - q2 + q3;
            const double ga1 = (x[0] != 0.0 && s1 > 1e-8) ? gx[0] * ((x[0] > 0) ? 1.0 : -1.0) / (4.0 * sqrt(s1 / 2.0)) : 0.0;
            const double ga2 = (x[1] != 0.0 && s2 > 1e-8) ? gx[1] * ((x[1] > 0) ? 1.0 : -1.0) / (4.0 * sqrt(s2 / 2.0)) : 0.0;
            const double ga3 = (x[2] != 0.0 && s3 > 1e-8) ? gx[2] * ((x[2] > 0) ? 1.0 : -1.0) / (4.0 * sqrt(s3 / 2.0)) : 0.0;
            const double gq1 = ga1 - ga2 - ga3, gq2 = -ga1 + ga2 - ga3, gq3 = -ga1 - ga2 + ga3;
            gr[0] += gq1 * k; gr[4] += gq2 * k; gr[8] += gq3 * k;
            const double gk = gq1 * (rm[0] - 1.0) + gq2 * (rm[4] - 1.0) + gq3 * (rm[8] - 1.0);
            // k(theta) = theta^2 / (1 - cos theta)
            const double dk = (2.0 * th * (1.0 - cos(th)) - th * th * sin(th)) / ((1.0 - cos(th)) * (1.0 - cos(th)));
            if (!clamped) { const double gcth = gk * dk * (-1.0 / sin(th)); gr[0] += 0.5 * gcth; gr[4] += 0.5 * gcth; gr[8] += 0.5 * gcth; }
        } else {
            const double sn = sin(th);
            const double ratio = (th < 1e-20) ? 1.0 : th / sn;
            const double w[3] = {-0.5 * (rm[5] - rm[7]), 0.5 * (rm[2] - rm[6]), -0.5 * (rm[1] - rm[3])};
            const double gw[3] = {ratio * gx[0], ratio * gx[1], ratio * gx[2]};
            gr[5] += -0.5 * gw[0]; gr[7] += 0.5 * gw[0]; gr[2] += 0.5 * gw[1]; gr[6] += -0.5 * gw[1]; gr[1] += -0.5 * gw[2]; gr[3] += 0.5 * gw[2];
            if (!clamped && th > 1e-6) {
                // d ratio / d cos(theta) = (d ratio / d theta) (-1 / sin theta), d ratio / d theta = (sin - theta cos) / sin^2
                const double dr = (sn - th * cos(th)) / (sn * sn);
                const double gcth = (w[0] * gx[0] + w[1] * gx[1] + w[2] * gx[2]) * dr * (-1.0 / sn);
                gr[0] += 0.5 * gcth; gr[4] += 0.5 * gcth; gr[8] += 0.5 * gcth;
            }
        }
        for (int e = 0; e < 9; ++e) g_rot[((size_t)r * joint_count + jj) * 9 + e] = gr[e];
    }
    for (int k = 0; k < CTX; ++k) g_ctx_out[((size_t)r * joint_count + jj) * CTX + k] = gc[k];
}

}  // namespace

extern "C" int hf_flow_log_prob_backward(const hf_flow_t* h, const float* ctx, int ctx_row_stride, int joint_first, int joint_count,
                                         const double* rot_f64, const float* grad_out, int R, float* grad_ctx, double* grad_rot, void* stream) {
    if (!h || !ctx || !rot_f64 || !grad_out || !grad_ctx || !grad_rot) return hf::fail(HF_ERR_INVALID, "hf_flow_log_prob_backward: null argument");
    if (joint_first < 0 || joint_count < 1 || joint_first + joint_count > h->P.J)
        return hf::fail(HF_ERR_INVALID, "hf_flow_log_prob_backward: joints [%d,%d) out of range", joint_first, joint_first + joint_count);
    if (R <= 0) return HF_OK;
    HF_CUDA(cudaFuncSetAttribute(flow_logprob_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 0));
    flow_logprob_bwd_kernel<false><<<dim3(hf::div_up(R, 64), joint_count), 64, 0, (cudaStream_t)stream>>>(h->P, ctx, ctx_row_stride, joint_first, joint_count, rot_f64,
                                                                                                     nullptr, grad_out, R, grad_ctx, grad_rot, nullptr);
    HF_LAUNCH_CHECK();
    return HF_OK;
}

extern "C" int hf_flow_algebra_log_prob_backward(const hf_flow_t* h, const float* ctx, int ctx_row_stride, int joint_first, int joint_count,
                                                 const float* v, const float* grad_out, int R, float* grad_ctx, float* grad_v, void* stream) {
    if (!h || !ctx || !v || !grad_out || !grad_ctx || !grad_v) return hf::fail(HF_ERR_INVALID, "hf_flow_algebra_log_prob_backward: null argument");
    if (joint_first < 0 || joint_count < 1 || joint_first + joint_count > h->P.J)
        return hf::fail(HF_ERR_INVALID, "hf_flow_algebra_log_prob_backward: joints [%d,%d) out of range", joint_first, joint_first + joint_count);
    if (R <= 0) return HF_OK;
    flow_logprob_bwd_kernel<true><<<dim3(hf::div_up(R, 64), joint_count), 64, 0, (cudaStream_t)stream>>>(h->P, ctx, ctx_row_stride, joint_first, joint_count, nullptr,
                                                                                                    v, grad_out, R, grad_ctx, nullptr, grad_v);
    HF_LAUNCH_CHECK();
    return HF_OK;
}
