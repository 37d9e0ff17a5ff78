// ResNet encoder for sm_100a: BatchNorm-folded conv (+bias, +residual, +ReLU) as an implicit GEMM on the
// 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM), operands staged by TMA.
//
// Replaces models/resnet.py:202-217.  Activations are bf16 NHWC, weights bf16
// (Cout, kh, kw, Cin) with the eval-mode BatchNorm scale folded in, accumulation fp32.
//
// im2col is never materialised: the A tile of one k-block is ONE 4-D TMA box {64 channels, TW, TH, TB}
// of the NHWC activation tensor, shifted by the filter tap; out-of-bounds box elements are zero-filled by
// the TMA unit, which is exactly the convolution's zero padding.  Stride-2 convolutions use one tensor
// map per input parity (pixel stride 2 in the map), so every load stays a plain tiled box.  The 7x7/2 stem
// (Cin = 18 -> 32) reads a physically padded input as "pixel pairs" (2 taps x 32 ch = one 128-byte row).
//
// Kernel = 6 warps: warp 0 TMA producer, warp 1 MMA issuer (one elected thread) + TMEM owner, warps 2-5
// epilogue (TMEM -> registers -> bias/residual/ReLU -> bf16 NHWC).  smem ring of STAGES x (A 16 KB + B BN*128 B),
// 128-byte swizzle on both the TMA and the UMMA descriptor side.
#include "common.cuh"
#include "tma.cuh"
#include <vector>
#include <map>
#include <cstring>

namespace {

// ------------------------------------------------------------------ tcgen05 implicit-GEMM conv
struct ConvGeom {
    int kind;        // 0 = generic (Cin % 64 == 0), 1 = stem pixel-pair layout (32 ch), 2 = compact stem (24 ch, sliding window map)
    int ksize, stride, pad;
    int cin;         // channels per pixel in the A tensor
    int Ho, Wo, B, cout;
    int TW, TH, TB;  // output tile (TW*TH*TB == 128)
    int tiles_w, tiles_h, tiles_b;
    int nkb;         // number of 64-wide k-blocks (primary convolution + fused 1x1 branch)
    int nkb1;        // k-blocks of the primary convolution; k-blocks >= nkb1 read the second input (1x1, own stride)
    int relu;
};

struct ConvMaps {
    CUtensorMap a[4];   // activation maps (stride 1: [0]; stride 2: [ph*2+pw]; stem: [ph])
    CUtensorMap a2;     // second input of a fused op (the block's 1x1 downsample branch): strided 1x1 view of its NHWC tensor
    CUtensorMap b;      // weights [Cout][K]
    CUtensorMap o;      // output (B,Ho,Wo,Cout), box {64, TW, TH, TB}
    CUtensorMap r;      // residual, same shape as the output
};

// timing aid (HF_CONV_DBG bit 3): %globaltimer stamps of CTA 0 at the milestones of one launch, fetched by hf_debug_conv_stamps
__device__ unsigned long long g_conv_ts[16];
__device__ __forceinline__ void conv_stamp(int dbg, int slot) {
    if ((dbg & 8) && blockIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_conv_ts[slot] = t;
    }
}

constexpr int CONV_MAX_KB = 96;     // k-blocks per tile (3x3 x 512 channels = 72; checked on the host)
constexpr int CONV_THREADS = 416;   // warp 0 operand TMA, warp 1 MMA issuer + TMEM owner, warps 2-9 epilogue, warp 10 residual TMA,
                                    // warps 11 / 12: second operand-TMA / MMA issuer (odd k-blocks)

// Persistent implicit-GEMM convolution: one CTA per SM walks the output tiles (n-tile fastest, so the activation
// tile stays hot in L2 across the Cout tiles).  Every byte that crosses the SM boundary moves by TMA:
//   operands      global -> smem ring   (full/empty mbarriers, STAGES deep)
//   residual      global -> staging tile (rfull), prefetched one tile ahead
//   result        staging tile -> global (TMA store, bulk groups; sfree recycles the staging tile)
// MMA accumulators are double-buffered in TMEM (tfull/tempty), so the epilogue of tile t (TMEM -> +bias
// +residual -> ReLU -> bf16, in place in the 128B-swizzled staging tile) overlaps the loads/MMAs of tile t+1.
// PAIR: the two CTAs of a (2,1,1) cluster work on two vertically adjacent M-tiles of the same N-tile as ONE 256 x BN
// tcgen05.mma.cta_group::2 issued by the leader (rank 0): each CTA stages its own 128 activation rows but only HALF of the
// weight tile (BN/2 rows), which cuts the operand bytes every SM pulls through the L2->SM fabric (the measured limiter)
// by a quarter (BN = 128) to a third (BN = 256).  num_tiles then counts pair tiles.
template <int BN, int STAGES, int SR, bool PAIR>
__global__ void __launch_bounds__(CONV_THREADS, 1)
conv_tcgen05_kernel(const __grid_constant__ ConvMaps maps, const ConvGeom g, const float* __restrict__ bias,
                    int has_res, int num_tiles, int ntn, int pdl_early, int dbg) {
    // dbg (HF_CONV_DBG, timing experiments only, results are garbage): bit 0 skip the MMAs, bit 1 skip the activation loads,
    // bit 2 skip the weight loads
    constexpr int BROWS = PAIR ? BN / 2 : BN;     // weight rows staged by this CTA
    constexpr int A_BYTES = 128 * 128, B_BYTES = BROWS * 128, STAGE_BYTES = A_BYTES + B_BYTES;
    if (dbg & 256) num_tiles = 0;                  // timing experiment: prologue + teardown only
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
    const bool leader = rank == 0;
    const int cta = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;           // persistent worker id (a CTA or a CTA pair)
    const int nworkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    if (threadIdx.x == 0) conv_stamp(dbg, 0);                       // kernel entry
    constexpr int SBUF_BYTES = 128 * BN * 2;     // staging tile: BN/64 boxes of [128 rows][128 B], swizzled
    // DUAL: two producer warps (0, 11) and two MMA issuers (1, 12) take alternate k-blocks; the issuers accumulate into
    // SEPARATE TMEM accumulators that the epilogue adds.  Each single-thread role spends ~700 cycles per k-block on the issue
    // latencies of its barrier / TMA / MMA / commit instructions against 256 cycles of tensor-pipe work; two
    // interleaved instances of each role halve that.  Needs 2 x 2 x BN TMEM columns (BN <= 128) and at least two k-blocks.
    // The split is by the parity of the GLOBAL k-block counter and the ring has an even number of slots, so every slot (and
    // its two mbarriers) is always handled by the same producer and the same issuer: an mbarrier wait only tells phases
    // apart by parity, which is safe only for a waiter that has observed every earlier phase of that barrier.
    constexpr bool DUAL_OK = !PAIR && BN <= 128 && (STAGES % 2 == 0);
    constexpr int NACC = DUAL_OK ? 2 : 1;
    constexpr uint32_t TMEM_COLS = (2 * NACC * BN <= 128) ? 128 : ((2 * NACC * BN <= 256) ? 256 : 512);
    const bool dual = DUAL_OK && g.nkb >= 2 && !(dbg & 16);
    constexpr int HALF = BN / 2;                 // columns per epilogue warp
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * STAGES + 4 + 2 * SR];
    __shared__ uint32_t tmem_base_s;
    // per-k-block load coordinates {tensor-map index (-1 = fused branch input), channel offset, dw, dh}: they do not depend on
    // the tile, and computing them in the single-thread producer loop (two integer divisions per k-block) measured ~110 cycles
    // of the ~350 a producer iteration takes (tools/ubench/tma_ring.cu)
    __shared__ int4 ktab[CONV_MAX_KB];
    for (int kb = threadIdx.x; kb < g.nkb; kb += CONV_THREADS) {
        int mi, c0, d1, d2;
        if (kb >= g.nkb1) { mi = -1; c0 = (kb - g.nkb1) * 64; d1 = 0; d2 = 0; }
        else if (g.kind == 0) {
            const int cpb = g.cin >> 6;
            const int tap = kb / cpb, cb = kb - tap * cpb;
            const int kh = tap / g.ksize, kw = tap - kh * g.ksize;
            const int dw = kw - g.pad, dh = kh - g.pad;
            c0 = cb * 64;
            if (g.stride == 1) { mi = 0; d1 = dw; d2 = dh; }
            else { const int pw = dw & 1, phh = dh & 1; mi = phh * 2 + pw; d1 = (dw - pw) >> 1; d2 = (dh - phh) >> 1; }
        } else if (g.kind == 1) { const int kh1 = kb >> 2, q4 = kb & 3; mi = kh1 & 1; c0 = 0; d1 = q4; d2 = kh1 >> 1; }
        else { const int kh = kb / 3, cb = kb - kh * 3; mi = kh & 1; c0 = cb * 64; d1 = 0; d2 = kh >> 1; }
        ktab[kb] = make_int4(mi, c0, d1, d2);
    }
    const uint32_t tile_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sbuf_base = tile_base + STAGES * STAGE_BYTES;
    uint8_t* sbuf_ptr = smem_raw + (sbuf_base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]);
    const uint32_t tfull0 = smem_u32(&bars[2 * STAGES]), tempty0 = smem_u32(&bars[2 * STAGES + 2]);
    const uint32_t rfull0 = smem_u32(&bars[2 * STAGES + 4]), sfree0 = smem_u32(&bars[2 * STAGES + 4 + SR]);
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(tfull0 + 8 * b, dual ? 2 : 1); mbar_init(tempty0 + 8 * b, PAIR ? 16 : 8); }
        for (int b = 0; b < SR; ++b) { mbar_init(rfull0 + 8 * b, 1); mbar_init(sfree0 + 8 * b, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tcgen05_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();       // barrier inits + TMEM base visible (pair: in both CTAs)
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    if (threadIdx.x == 0) conv_stamp(dbg, 1);                       // barriers + TMEM ready
    if (pdl_early) {
        // Programmatic dependent launch: release the next kernel of the stream now (its CTAs take over each SM as soon as this
        // grid's CTA leaves it and run their own prologue up to this point), then wait until the predecessor grid has
        // completed and its writes are visible before touching any activation buffer.
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }
    if (threadIdx.x == 0) conv_stamp(dbg, 2);                       // dependency wait over

    if (warp == 0 || warp == 11) {
        const int t = warp == 0 ? 0 : 1;                             // producer id = parity of the k-blocks it loads
        if ((t == 0 || dual) && lane == 0) {
            // ONE thread walks the loop: an mbarrier parity wait is only safe for a waiter that goes on to the next use of
            // the barrier itself (31 more lanes polling would race with the phases lane 0 starts), and a lane-0 poll with the
            // other lanes parked at a warp barrier inside the loop measured 2x slower than leaving them out altogether
            const uint32_t full_lead = PAIR ? mapa_rank(full0, 0) : full0;   // pair: both CTAs' loads complete on the leader's barrier
            const uint32_t my_bytes = ((dbg & 2) ? 0u : (uint32_t)A_BYTES) + ((dbg & 4) ? 0u : (uint32_t)B_BYTES);
            const uint32_t tx_bytes = PAIR ? 2 * my_bytes : my_bytes;
            const int nkb = g.nkb, kstep = dual ? 2 : 1;
            uint32_t base = 0;                                        // k-blocks of all previous tiles (ring position)
            int nt = cta % ntn, mt = cta / ntn;                      // tile = mt * ntn + nt, advanced by nworkers per step
            const int step_nt = nworkers % ntn, step_mt = nworkers / ntn;
            for (int tile = cta; tile < num_tiles; tile += nworkers) {
                int tt = PAIR ? 2 * mt + (int)rank : mt;             // this CTA's M-tile (past the end: all-zero boxes)
                const int tw = tt % g.tiles_w; tt /= g.tiles_w;
                const int th = tt % g.tiles_h;
                const int tb = tt / g.tiles_h;
                const int wo0 = tw * g.TW, ho0 = th * g.TH, b0 = tb * g.TB, n0 = nt * BN + (int)rank * BROWS;
                for (int kb = dual ? (int)((base ^ (uint32_t)t) & 1u) : 0; kb < nkb; kb += kstep) {
                    const uint32_t kbc = base + (uint32_t)kb, st = kbc % STAGES, ph = (kbc / STAGES) & 1u;
                    mbar_wait_long(empty0 + 8 * st, ph ^ 1u);
                    const uint32_t sa = tile_base + st * STAGE_BYTES, sb = sa + A_BYTES;
                    const uint32_t fb = full_lead + 8 * st;
                    if (leader) mbar_expect_tx(full0 + 8 * st, tx_bytes);
                    const int4 kt = ktab[kb];
                    const int mi = kt.x, c0 = kt.y, c1 = wo0 + kt.z, c2 = ho0 + kt.w;
                    const CUtensorMap* amap;
                    amap = mi < 0 ? &maps.a2 : &maps.a[mi];
                    if (PAIR) {
                        if (!(dbg & 2)) tma_load_4d_2cta(sa, amap, fb, c0, c1, c2, b0);
                        if (!(dbg & 4)) tma_load_2d_2cta(sb, &maps.b, fb, kb * 64, n0);
                    } else {
                        if (!(dbg & 2)) tma_load_4d(sa, amap, fb, c0, c1, c2, b0);
                        if (!(dbg & 4)) tma_load_2d(sb, &maps.b, fb, kb * 64, n0);
                    }
                }
                base += (uint32_t)nkb;
                nt += step_nt; mt += step_mt;
                if (nt >= ntn) { nt -= ntn; ++mt; }
            }
        }
    } else if (warp == 1 || warp == 12) {
        const int t = warp == 1 ? 0 : 1;                             // issuer id = parity of the k-blocks it takes
        if (leader && (t == 0 || dual) && lane == 0) {              // one thread (see the producer)
            const uint32_t idesc = umma_idesc_bf16(PAIR ? 256 : 128, BN);
            const int nkb = g.nkb, kstep = dual ? 2 : 1;
            uint32_t lt = 0, base = 0;                               // base = k-blocks of all previous tiles (ring position)
            for (int tile = cta; tile < num_tiles; tile += nworkers, ++lt) {
                const uint32_t buf = lt & 1u;
                mbar_wait_long(tempty0 + 8 * buf, ((lt >> 1) & 1u) ^ 1u);
                tcgen05_fence_after();
                const uint32_t d = tmem_base + (buf * NACC + (uint32_t)t) * BN;
                const int kb0 = dual ? (int)((base ^ (uint32_t)t) & 1u) : 0;     // this issuer's first k-block of the tile
                for (int kb = kb0; kb < nkb; kb += kstep) {
                    const uint32_t kbc = base + (uint32_t)kb, st = kbc % STAGES, ph = (kbc / STAGES) & 1u;
                    mbar_wait_long(full0 + 8 * st, ph);
                    tcgen05_fence_after();
                    if (lt == 0 && kb == 0) conv_stamp(dbg, 3);                  // first operand stage landed
                    const uint32_t sa = tile_base + st * STAGE_BYTES, sb = sa + A_BYTES;
                    const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sb);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        if (dbg & 1) continue;
                        const uint32_t acc = (uint32_t)((kb > kb0) | (k != 0));        // this issuer's first MMA of the tile overwrites
                        if (PAIR) umma_bf16_2cta(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, acc);
                        else umma_bf16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, acc);
                    }
                    if (PAIR) umma_commit_2cta(empty0 + 8 * st); else umma_commit(empty0 + 8 * st);
                }
                if (PAIR) umma_commit_2cta(tfull0 + 8 * buf); else umma_commit(tfull0 + 8 * buf);
                if (lt == 0 && t == 0) conv_stamp(dbg, 4);          // all MMAs of the first tile issued
                base += (uint32_t)nkb;
            }
        }
    } else if (warp == 10) {
        if (lane == 0) {
            // staging tiles: wait until the previous store of the buffer has left smem, then fetch the residual
            uint32_t lt = 0;
            for (int tile = cta; tile < num_tiles; tile += nworkers, ++lt) {
                const int nt = tile % ntn;
                int t = PAIR ? 2 * (tile / ntn) + (int)rank : tile / ntn;
                const int tw = t % g.tiles_w; t /= g.tiles_w;
                const int th = t % g.tiles_h;
                const int tb = t / g.tiles_h;
                const uint32_t sb = lt % SR, k = lt / SR;
                mbar_wait_long(sfree0 + 8 * sb, (k & 1u) ^ 1u);
                const uint32_t rb = rfull0 + 8 * sb;
                if (has_res && !(dbg & 128)) {
                    mbar_expect_tx(rb, SBUF_BYTES);
#pragma unroll
                    for (int x = 0; x < BN / 64; ++x)
                        tma_load_4d(sbuf_base + sb * SBUF_BYTES + x * 16384, &maps.r, rb, nt * BN + x * 64, tw * g.TW, th * g.TH, tb * g.TB);
                } else {
                    mbar_arrive(rb);
                }
            }
        }
    } else {
        // epilogue: warp e owns TMEM lanes [32*(warp%4), +32) (= tile rows) and the column half e/4
        const int e = warp - 2;
        const int q = warp & 3, colhalf = e >> 2;
        const int row = q * 32 + lane;
        uint32_t lt = 0;
        for (int tile = cta; tile < num_tiles; tile += nworkers, ++lt) {
            const int nt = tile % ntn;
            int t = PAIR ? 2 * (tile / ntn) + (int)rank : tile / ntn;
            const int tw = t % g.tiles_w; t /= g.tiles_w;
            const int th = t % g.tiles_h;
            const int tb = t / g.tiles_h;
            const int n0 = nt * BN;
            const uint32_t buf = lt & 1u, par = (lt >> 1) & 1u;
            const uint32_t sb = lt % SR;
            mbar_wait_warp(rfull0 + 8 * sb, (lt / SR) & 1u);  // staging tile free (+ residual landed)
            mbar_wait_warp(tfull0 + 8 * buf, par);            // accumulator complete
            tcgen05_fence_after();
            if (lt == 0 && threadIdx.x == 64) conv_stamp(dbg, 5);   // first accumulator complete
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + buf * (NACC * BN) + (uint32_t)(colhalf * HALF);
            uint8_t* srow = sbuf_ptr + sb * SBUF_BYTES + row * 128;
            // 32-column chunks, double-buffered in registers: the TMEM load of chunk c+1 is in flight while chunk c is
            // converted (tcgen05.wait::ld waits for every outstanding load, so the next one is issued right after it)
            constexpr int NCH = HALF / 32;
            uint32_t v[2][32];
            if (!dual && !(dbg & 32)) tmem_ld32(trow, v[0]);
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                if (dbg & 32) continue;
                const int col = colhalf * HALF + ch * 32;                  // first column (within the BN tile) of this chunk
                float4 bb[8];
                const float4* bp = reinterpret_cast<const float4*>(bias + n0 + col);
#pragma unroll
                for (int i = 0; i < 8; ++i) bb[i] = __ldg(bp + i);
                uint8_t* sbox = srow + (col >> 6) * 16384;                 // 64-channel box
                const int j0 = (col & 63) >> 3;                            // first 16-byte piece within the 128-byte row
                uint4 rr[4];
                if (has_res) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) rr[i] = *reinterpret_cast<const uint4*>(sbox + (((j0 + i) ^ (row & 7)) << 4));
                }
                if (dual) {          // the two issuers' partial sums (even / odd k-blocks) are added here
                    tmem_ld32(trow + (uint32_t)(ch * 32), v[ch & 1]);
                    tmem_ld32(trow + (uint32_t)(BN + ch * 32), v[(ch & 1) ^ 1]);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[ch & 1][i] = __float_as_uint(__uint_as_float(v[ch & 1][i]) + __uint_as_float(v[(ch & 1) ^ 1][i]));
                } else {
                    tmem_ld_wait();
                    if (ch + 1 < NCH) tmem_ld32(trow + (uint32_t)((ch + 1) * 32), v[(ch + 1) & 1]);
                }
                const uint32_t* vv = v[ch & 1];
#pragma unroll
                for (int i = 0; i < 4; ++i) {                              // one 16-byte piece = 8 channels
                    float f[8];
                    const float4 b0 = bb[2 * i], b1 = bb[2 * i + 1];
                    f[0] = __uint_as_float(vv[8 * i + 0]) + b0.x; f[1] = __uint_as_float(vv[8 * i + 1]) + b0.y;
                    f[2] = __uint_as_float(vv[8 * i + 2]) + b0.z; f[3] = __uint_as_float(vv[8 * i + 3]) + b0.w;
                    f[4] = __uint_as_float(vv[8 * i + 4]) + b1.x; f[5] = __uint_as_float(vv[8 * i + 5]) + b1.y;
                    f[6] = __uint_as_float(vv[8 * i + 6]) + b1.z; f[7] = __uint_as_float(vv[8 * i + 7]) + b1.w;
                    if (has_res) {
                        const uint32_t r4[4] = {rr[i].x, rr[i].y, rr[i].z, rr[i].w};
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&r4[t]);
                            f[2 * t] += __bfloat162float(h.x);
                            f[2 * t + 1] += __bfloat162float(h.y);
                        }
                    }
                    if (g.relu) {
#pragma unroll
                        for (int t = 0; t < 8; ++t) f[t] = fmaxf(f[t], 0.f);
                    }
                    uint32_t o[4];
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * t], f[2 * t + 1]);
                        o[t] = *reinterpret_cast<uint32_t*>(&h);
                    }
                    *reinterpret_cast<uint4*>(sbox + (((j0 + i) ^ (row & 7)) << 4)) = make_uint4(o[0], o[1], o[2], o[3]);
                }
            }
            // accumulator buffer can be refilled; staging tile goes out by TMA
            tcgen05_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {                                   // the accumulator buffer of BOTH CTAs is released to the leader's MMA thread
                if (PAIR && !leader) mbar_arrive_cluster(mapa_rank(tempty0 + 8 * buf, 0)); else mbar_arrive(tempty0 + 8 * buf);
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) {
                if (lt == 0) conv_stamp(dbg, 6);                    // first tile converted into the staging buffer
                if (!(dbg & 64)) {
#pragma unroll
                for (int x = 0; x < BN / 64; ++x)
                    tma_store_4d(&maps.o, sbuf_base + sb * SBUF_BYTES + x * 16384, n0 + x * 64, tw * g.TW, th * g.TH, tb * g.TB);
                }
                bulk_commit();
                if (SR == 1) {                            // single staging tile: recycle it as soon as this store has read it
                    bulk_wait_read<0>();
                    mbar_arrive(sfree0);
                } else {
                    bulk_wait_read<1>();                  // every group but the newest has finished reading smem
                    if (lt >= 1) mbar_arrive(sfree0 + 8 * ((lt - 1) % SR));
                }
            }
        }
        if (threadIdx.x == 64) { conv_stamp(dbg, 7); bulk_wait<0>(); conv_stamp(dbg, 8); }   // last store issued / landed
    }
    tcgen05_fence_before();
    if (!pdl_early) __threadfence();
    __syncthreads();
    // late-trigger mode: every thread of this CTA is past its last global write (thread 64 has waited for the TMA stores)
    if (!pdl_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (threadIdx.x == 0) conv_stamp(dbg, 9);                       // all roles done
    if (PAIR) cluster_sync_all();       // neither CTA may leave (or free TMEM) while its peer can still signal it
    if (warp == 1) {
        tcgen05_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ SIMT direct convolution (debug cross-check)
__global__ void conv_simt_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                                 const float* __restrict__ bias, const __nv_bfloat16* __restrict__ res,
                                 __nv_bfloat16* __restrict__ y, int B, int H, int W, int cin, int cout, int ks,
                                 int stride, int pad, int Ho, int Wo, int relu, const __nv_bfloat16* __restrict__ x2,
                                 int H2, int W2, int cin2, int stride2) {
    const int ldw = ks * ks * cin + (x2 ? cin2 : 0);      // weight row = [primary taps | fused 1x1 branch]
    const size_t total = (size_t)B * Ho * Wo * cout;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int co = (int)(e % cout);
        size_t p = e / cout;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int b = (int)(p / Ho);
        float acc = 0.f;
        for (int kh = 0; kh < ks; ++kh) {
            const int ih = ho * stride + kh - pad;
            if (ih < 0 || ih >= H) continue;
            for (int kw = 0; kw < ks; ++kw) {
                const int iw = wo * stride + kw - pad;
                if (iw < 0 || iw >= W) continue;
                const __nv_bfloat16* xp = x + (((size_t)b * H + ih) * W + iw) * cin;
                const __nv_bfloat16* wp = w + (size_t)co * ldw + ((size_t)kh * ks + kw) * cin;
                for (int c = 0; c < cin; ++c) acc = fmaf(__bfloat162float(xp[c]), __bfloat162float(wp[c]), acc);
            }
        }
        if (x2) {
            const __nv_bfloat16* xp = x2 + (((size_t)b * H2 + (size_t)ho * stride2) * W2 + (size_t)wo * stride2) * cin2;
            const __nv_bfloat16* wp = w + (size_t)co * ldw + (size_t)ks * ks * cin;
            for (int c = 0; c < cin2; ++c) acc = fmaf(__bfloat162float(xp[c]), __bfloat162float(wp[c]), acc);
        }
        acc += bias[co];
        if (res) acc += __bfloat162float(res[e]);
        if (relu) acc = fmaxf(acc, 0.f);
        y[e] = __float2bfloat16_rn(acc);
    }
}

// ------------------------------------------------------------------ layout / pooling kernels
// fp32 NCHW -> bf16 NHWC with channel padding and spatial zero border (stem input).
__device__ __forceinline__ float ld_as_float(const float* p) { return __ldg(p); }
__device__ __forceinline__ float ld_as_float(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename TIn>
__global__ void nchw_to_nhwc_pad_kernel(const TIn* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int C,
                                        int H, int W, int Cp, int Hp, int Wp, int top, int left) {
    // one thread per (b, h, w): reads C strided planes (coalesced over w), writes Cp contiguous bf16
    HF_PDL_SYNC();
    const size_t total = (size_t)B * H * W;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int w = (int)(e % W);
        size_t p = e / W;
        const int h = (int)(p % H);
        const int b = (int)(p / H);
        __nv_bfloat16* o = y + (((size_t)b * Hp + h + top) * Wp + w + left) * Cp;
        const TIn* xi = x + ((size_t)b * C * H + h) * W + w;
        for (int c0 = 0; c0 < Cp; c0 += 8) {
            uint32_t pk[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int ca = c0 + 2 * i, cb = ca + 1;
                const float fa = ca < C ? ld_as_float(xi + (size_t)ca * H * W) : 0.f;
                const float fb = cb < C ? ld_as_float(xi + (size_t)cb * H * W) : 0.f;
                __nv_bfloat162 hh = __floats2bfloat162_rn(fa, fb);
                pk[i] = *reinterpret_cast<uint32_t*>(&hh);
            }
            *reinterpret_cast<uint4*>(o + c0) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
    }
}

// 3x3 stride-2 pad-1 max pooling on bf16 NHWC, 8 channels per thread.
__global__ void maxpool_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int H, int W,
                               int C, int Ho, int Wo) {
    HF_PDL_SYNC();
    const int C8 = C / 8;
    const size_t total = (size_t)B * Ho * Wo * C8;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int c8 = (int)(e % C8);
        size_t p = e / C8;
        const int wo = (int)(p % Wo); p /= Wo;
        const int ho = (int)(p % Ho);
        const int b = (int)(p / Ho);
        // all nine taps are requested before any is used (out-of-range taps re-read the centre, which is always in range)
        uint4 tap[9];
        const int ihc = ho * 2, iwc = wo * 2;
#pragma unroll
        for (int kh = 0; kh < 3; ++kh)
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                int ih = ihc + kh - 1, iw = iwc + kw - 1;
                if (ih < 0 || ih >= H || iw < 0 || iw >= W) { ih = ihc; iw = iwc; }
                tap[kh * 3 + kw] = __ldg(reinterpret_cast<const uint4*>(x + (((size_t)b * H + ih) * W + iw) * C + c8 * 8));
            }
        float m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m[i] = -INFINITY;
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const uint32_t vv[4] = {tap[t].x, tap[t].y, tap[t].z, tap[t].w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&vv[i]);
                m[2 * i] = fmaxf(m[2 * i], __bfloat162float(h.x));
                m[2 * i + 1] = fmaxf(m[2 * i + 1], __bfloat162float(h.y));
            }
        }
        uint32_t o[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(m[2 * i], m[2 * i + 1]);
            o[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(y + (((size_t)b * Ho + ho) * Wo + wo) * C + c8 * 8) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

// global average pool bf16 NHWC -> fp32 (B, C)
__global__ void avgpool_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ y, int B, int HW, int C) {
    HF_PDL_SYNC();
    const int b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float s = 0.f;
    const __nv_bfloat16* p = x + (size_t)b * HW * C + c;
    for (int i = 0; i < HW; ++i) s += __bfloat162float(p[(size_t)i * C]);
    y[(size_t)b * C + c] = s / (float)HW;
}

// ------------------------------------------------------------------ host side
int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

int pow2_ceil(int v) { int p = 1; while (p < v) p <<= 1; return p; }

struct ConvPlan {
    ConvGeom g;
    ConvMaps maps;
    int bn, ntn, num_tiles, has_res, sr, pair;
    dim3 grid;
    size_t smem;
    int stages;
};

// x: NHWC activation (generic) or the padded stem input; returns launch plan.
// Generic: x (B,H,W,cin), cin % 64 == 0.  Stem (kind 1): x is (B,Hp,Wp,32) with 3 zero rows/cols before the
// image, ksize 7, stride 2, pad 3; H, W are the UNPADDED sizes.
int plan_conv(ConvPlan* p, int kind, const void* x, const void* w, int B, int H, int W, int cin, int cout, int ks,
              int stride, int pad, int relu, int Hp, int Wp, const void* out, const void* res,
              const void* x2 = nullptr, int H2 = 0, int W2 = 0, int cin2 = 0, int stride2 = 1) {
    ConvGeom& g = p->g;
    g.kind = kind; g.ksize = ks; g.stride = stride; g.pad = pad; g.cin = cin; g.B = B; g.cout = cout; g.relu = relu;
    g.Ho = (H + 2 * pad - ks) / stride + 1;
    g.Wo = (W + 2 * pad - ks) / stride + 1;
    g.TW = std::min(16, pow2_ceil(g.Wo));
    g.TH = std::min(128 / g.TW, pow2_ceil(g.Ho));
    g.TB = 128 / (g.TW * g.TH);
    g.tiles_w = hf::div_up(g.Wo, g.TW); g.tiles_h = hf::div_up(g.Ho, g.TH); g.tiles_b = hf::div_up(B, g.TB);
    if (cout % 64) return hf::fail(HF_ERR_UNSUPPORTED, "conv: cout %d is not a multiple of 64", cout);
    const uint32_t box[4] = {64, (uint32_t)g.TW, (uint32_t)g.TH, (uint32_t)g.TB};
    int ktot;
    if (kind == 0) {
        if (cin % 64) return hf::fail(HF_ERR_UNSUPPORTED, "conv: cin %d is not a multiple of 64", cin);
        if (stride != 1 && stride != 2) return hf::fail(HF_ERR_UNSUPPORTED, "conv: stride %d", stride);
        g.nkb = ks * ks * (cin / 64);
        ktot = ks * ks * cin;
        const uint64_t eb = 2;
        if (stride == 1) {
            const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)W, (uint64_t)H, (uint64_t)B};
            const uint64_t st[3] = {cin * eb, (uint64_t)W * cin * eb, (uint64_t)H * W * cin * eb};
            int rc = encode_map(&p->maps.a[0], x, 4, dims, st, box);
            if (rc) return rc;
        } else {
            for (int ph = 0; ph < 2; ++ph)
                for (int pw = 0; pw < 2; ++pw) {
                    if (ph >= H || pw >= W) continue;
                    const uint64_t dims[4] = {(uint64_t)cin, (uint64_t)((W - pw + 1) / 2), (uint64_t)((H - ph + 1) / 2), (uint64_t)B};
                    const uint64_t st[3] = {2 * cin * eb, 2 * (uint64_t)W * cin * eb, (uint64_t)H * W * cin * eb};
                    const uint8_t* base = (const uint8_t*)x + ((size_t)ph * W + pw) * cin * eb;
                    int rc = encode_map(&p->maps.a[ph * 2 + pw], base, 4, dims, st, box);
                    if (rc) return rc;
                }
        }
    } else if (kind == 2) {
        // Compact stem: the 7 horizontal taps x 24 channels of one output pixel are 168 contiguous bf16 of the padded NHWC
        // row (padded to 192 = three 64-wide k-blocks, the 8th tap has zero weights).  Consecutive output pixels start
        // 2 pixels = 96 bytes apart, so dimension 1 of the map is a SLIDING WINDOW over dimension 0 (stride 96 B < 384 B).
        if (ks != 7 || stride != 2 || pad != 3 || cin != 24 || (Wp & 1)) return hf::fail(HF_ERR_UNSUPPORTED, "stem conv: unsupported geometry");
        g.nkb = 7 * 3;
        ktot = 7 * 8 * 24;
        for (int ph = 0; ph < 2; ++ph) {
            const uint64_t dims[4] = {192, (uint64_t)g.Wo, (uint64_t)((Hp - ph + 1) / 2), (uint64_t)B};
            const uint64_t st[3] = {96, 2 * (uint64_t)Wp * 48, (uint64_t)Hp * Wp * 48};
            const uint8_t* base = (const uint8_t*)x + (size_t)ph * Wp * 48;
            int rc = encode_map(&p->maps.a[ph], base, 4, dims, st, box);
            if (rc) return rc;
        }
    } else {
        if (ks != 7 || stride != 2 || pad != 3 || cin != 32 || (Wp & 1)) return hf::fail(HF_ERR_UNSUPPORTED, "stem conv: unsupported geometry");
        g.nkb = 7 * 4;
        ktot = 7 * 8 * 32;
        for (int ph = 0; ph < 2; ++ph) {
            const uint64_t dims[4] = {64, (uint64_t)(Wp / 2), (uint64_t)((Hp - ph + 1) / 2), (uint64_t)B};
            const uint64_t st[3] = {128, 2 * (uint64_t)Wp * 64, (uint64_t)Hp * Wp * 64};
            const uint8_t* base = (const uint8_t*)x + (size_t)ph * Wp * 64;
            int rc = encode_map(&p->maps.a[ph], base, 4, dims, st, box);
            if (rc) return rc;
        }
    }
    g.nkb1 = g.nkb;
    if (g.nkb + (x2 ? cin2 / 64 : 0) > CONV_MAX_KB) return hf::fail(HF_ERR_UNSUPPORTED, "conv: %d k-blocks per tile exceed %d", g.nkb + (x2 ? cin2 / 64 : 0), CONV_MAX_KB);
    if (x2) {   // fused 1x1 branch: second input (B,H2,W2,cin2) sampled with stride2 at the output pixels, K appended
        if (kind != 0 || cin2 % 64 || stride2 < 1 || (H2 - 1) / stride2 + 1 != g.Ho || (W2 - 1) / stride2 + 1 != g.Wo)
            return hf::fail(HF_ERR_UNSUPPORTED, "conv: fused 1x1 branch does not match the output geometry");
        const uint64_t dims[4] = {(uint64_t)cin2, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)B};
        const uint64_t st[3] = {(uint64_t)stride2 * cin2 * 2, (uint64_t)stride2 * W2 * cin2 * 2, (uint64_t)H2 * W2 * cin2 * 2};
        int rc = encode_map(&p->maps.a2, x2, 4, dims, st, box);
        if (rc) return rc;
        g.nkb += cin2 / 64;
        ktot += cin2;
    } else {
        p->maps.a2 = p->maps.a[0];
    }
    const int tiles_m = g.tiles_w * g.tiles_h * g.tiles_b;
    p->bn = (cout % 128 == 0) ? 128 : 64;
    // 256-wide tiles for wide layers with enough tiles: 512 tensor-pipe cycles per k-block behind every 16 KB activation tile
    // and every barrier round trip (a single issuer then keeps the pipe > 90 % busy, tools/ubench/tma_ring.cu), half the tiles.
    // The rule looks at the layer only (>= 8 such tiles per image), never at the batch size: the accumulation order of a
    // layer must not depend on how many images share the launch, or shards would not reproduce the unsharded result.
    static const int bn256_env = getenv("HF_CONV_BN256") ? atoi(getenv("HF_CONV_BN256")) : 1;     // bit 0: layers without residual, bit 1: with
    if (cout % 256 == 0 && (g.Ho * g.Wo / 128) * (cout / 256) >= 8 && ((res == nullptr) ? (bn256_env & 1) : (bn256_env & 2))) p->bn = 256;
    {
        const uint64_t dims[4] = {(uint64_t)cout, (uint64_t)g.Wo, (uint64_t)g.Ho, (uint64_t)B};
        const uint64_t st[3] = {(uint64_t)cout * 2, (uint64_t)g.Wo * cout * 2, (uint64_t)g.Ho * g.Wo * cout * 2};
        int rc = encode_map(&p->maps.o, out, 4, dims, st, box);
        if (rc) return rc;
        p->has_res = res != nullptr;
        rc = encode_map(&p->maps.r, res ? res : out, 4, dims, st, box);
        if (rc) return rc;
    }
    // CTA pairs (one 256-row cta_group::2 MMA per two M-tiles, each CTA staging half of the weight tile): bit-identical
    // results and 25 % fewer operand bytes per SM, but measured 6-10 % SLOWER on this network (the k-block loop is bound by
    // the ~700 cycles of issue latency of its barrier / TMA / MMA / commit instructions, not by operand bytes),
    // so it is opt-in: HF_CONV_PAIR=1
    static const int pair_env = getenv("HF_CONV_PAIR") ? atoi(getenv("HF_CONV_PAIR")) : 0;
    p->pair = (pair_env && tiles_m >= 2) ? 1 : 0;
    {
        const uint64_t dims[2] = {(uint64_t)ktot, (uint64_t)cout};
        const uint64_t st[1] = {(uint64_t)ktot * 2};
        const uint32_t bx[2] = {64, (uint32_t)(p->pair ? p->bn / 2 : p->bn)};
        int rc = encode_map(&p->maps.b, w, 2, dims, st, bx);
        if (rc) return rc;
    }
    // operand ring depth vs staging tiles: residual layers prefetch the residual several tiles ahead (DRAM latency),
    // layers without residual spend the shared memory on a deeper operand ring
    // (one staging tile serialises the epilogue of tile t+1 behind the TMA store of tile t: only the 64 KB tiles of BN = 256)
    if (p->pair) {      // stage = 16 KB of activations + bn/2 weight rows
        if (p->bn == 256)      { p->stages = 4; p->sr = 1; }
        else if (p->bn == 128) { p->stages = p->has_res ? 4 : 5; p->sr = p->has_res ? 3 : 2; }
        else                   { p->stages = p->has_res ? 6 : 7; p->sr = p->has_res ? 4 : 2; }
    } else {
        // layers without residual: even ring -> dual producers / issuers; residual layers (short K, epilogue-bound) keep the
        // deeper residual prefetch (three / four staging tiles) and an odd ring, i.e. one producer and one issuer
        // HF_CONV_CFG bits (A/B switches, default all on): 1 = four-slot ring for residual layers, 2 = eight stages for the
        // 64-wide layers, 4 = six stages + one staging tile also for layers with two tiles per CTA
        static const int cfg_env = getenv("HF_CONV_CFG") ? atoi(getenv("HF_CONV_CFG")) : 7;
        if (p->bn == 256)      { p->stages = p->has_res ? 2 : 3; p->sr = p->has_res ? 2 : 1; }
        else if (p->bn == 128) { p->stages = p->has_res ? ((cfg_env & 1) ? 4 : 3) : 4; p->sr = p->has_res ? 3 : 2; }
        else                   { p->stages = p->has_res ? 5 : ((cfg_env & 2) ? 8 : 6); p->sr = p->has_res ? 4 : 2; }
    }
    static const int stages_env = getenv("HF_CONV_STAGES") ? atoi(getenv("HF_CONV_STAGES")) : 0;     // timing experiment: BN = 128 without residual
    if (stages_env && !p->pair && p->bn == 128 && !p->has_res) { p->stages = stages_env; p->sr = stages_env >= 5 ? 1 : 2; }
    p->ntn = cout / p->bn;
    // one tile per CTA (the 16x16 and 8x8 layers at B = 32): a second staging tile is useless, and in the network these
    // layers are bound by operand latency (weights and activations arrive from HBM / a cold L2), i.e. by the bytes the
    // ring keeps in flight -> six 32 KB stages + one staging tile = 224 KB
    static const int cfg_env2 = getenv("HF_CONV_CFG") ? atoi(getenv("HF_CONV_CFG")) : 7;
    if (!p->pair && p->bn == 128 && !p->has_res && tiles_m * p->ntn <= ((cfg_env2 & 4) ? 2 : 1) * num_sms() && stages_env == 0) { p->stages = 6; p->sr = 1; }
    p->smem = (size_t)p->stages * (128 * 128 + (p->pair ? p->bn / 2 : p->bn) * 128) + (size_t)p->sr * 128 * p->bn * 2 + 1024;
    if (p->pair) {
        p->num_tiles = hf::div_up(tiles_m, 2) * p->ntn;                      // pair tiles
        p->grid = dim3(2 * std::min(p->num_tiles, num_sms() / 2));
    } else {
        p->num_tiles = tiles_m * p->ntn;
        p->grid = dim3(std::min(p->num_tiles, num_sms()));
    }
    return HF_OK;
}

template <int BN, int STAGES, int SR, bool PAIR>
int launch_conv_t(const ConvPlan& p, const float* bias, cudaStream_t s) {
    static size_t attr_bytes[32] = {};   // the attribute is per device: raise it (to what this plan needs) once per device
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_bytes[dev & 31] < p.smem) {
        HF_CUDA(cudaFuncSetAttribute(conv_tcgen05_kernel<BN, STAGES, SR, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
        attr_bytes[dev & 31] = p.smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = p.grid; cfg.blockDim = dim3(CONV_THREADS); cfg.dynamicSmemBytes = p.smem; cfg.stream = s;
    cudaLaunchAttribute at[2];
    int na = 0;
    static const bool no_pdl = getenv("HF_NO_PDL") != nullptr;
    if (!no_pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (PAIR) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    // early trigger + griddepcontrol.wait is the default (digest-identical to serialized launches, tools/enc_pdl_check.py);
    // HF_PDL_EARLY=0 falls back to triggering at the end of the kernel
    static const int pdl_early = no_pdl ? 0 : (getenv("HF_PDL_EARLY") ? atoi(getenv("HF_PDL_EARLY")) : 1);
    static const int dbg = getenv("HF_CONV_DBG") ? atoi(getenv("HF_CONV_DBG")) : 0;
    HF_CUDA(cudaLaunchKernelEx(&cfg, conv_tcgen05_kernel<BN, STAGES, SR, PAIR>, p.maps, p.g, bias, p.has_res, p.num_tiles, p.ntn, pdl_early, dbg));
    HF_LAUNCH_CHECK();
    return HF_OK;
}

int launch_conv(const ConvPlan& p, const float* bias, cudaStream_t s) {
    if (p.pair) {
        if (p.bn == 256) return launch_conv_t<256, 4, 1, true>(p, bias, s);
        if (p.bn == 128) return p.has_res ? launch_conv_t<128, 4, 3, true>(p, bias, s) : launch_conv_t<128, 5, 2, true>(p, bias, s);
        return p.has_res ? launch_conv_t<64, 6, 4, true>(p, bias, s) : launch_conv_t<64, 7, 2, true>(p, bias, s);
    }
    if (p.bn == 256) return p.has_res ? launch_conv_t<256, 2, 2, false>(p, bias, s) : launch_conv_t<256, 3, 1, false>(p, bias, s);
    if (p.bn == 128 && !p.has_res && p.stages == 6) return launch_conv_t<128, 6, 1, false>(p, bias, s);
    if (p.bn == 128 && !p.has_res && p.stages == 2) return launch_conv_t<128, 2, 2, false>(p, bias, s);
    if (p.bn == 128 && !p.has_res && p.stages == 3) return launch_conv_t<128, 3, 2, false>(p, bias, s);
    if (p.bn == 128 && !p.has_res && p.stages == 5) return launch_conv_t<128, 5, 1, false>(p, bias, s);
    if (p.bn == 128 && p.has_res && p.stages == 4) return launch_conv_t<128, 4, 3, false>(p, bias, s);
    if (p.bn == 128) return p.has_res ? launch_conv_t<128, 3, 3, false>(p, bias, s) : launch_conv_t<128, 4, 2, false>(p, bias, s);
    if (p.bn == 64 && !p.has_res && p.stages == 8) return launch_conv_t<64, 8, 2, false>(p, bias, s);
    return p.has_res ? launch_conv_t<64, 5, 4, false>(p, bias, s) : launch_conv_t<64, 6, 2, false>(p, bias, s);
}

int launch_simt(const __nv_bfloat16* x, const __nv_bfloat16* w, const float* bias, const __nv_bfloat16* res,
                __nv_bfloat16* y, int B, int H, int W, int cin, int cout, int ks, int stride, int pad, int relu,
                cudaStream_t s, int Ho_override = 0, int Wo_override = 0, const __nv_bfloat16* x2 = nullptr, int H2 = 0, int W2 = 0,
                int cin2 = 0, int stride2 = 1) {
    const int Ho = Ho_override ? Ho_override : (H + 2 * pad - ks) / stride + 1;
    const int Wo = Wo_override ? Wo_override : (W + 2 * pad - ks) / stride + 1;
    const size_t total = (size_t)B * Ho * Wo * cout;
    int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 64);
    conv_simt_kernel<<<blocks, 256, 0, s>>>(x, w, bias, res, y, B, H, W, cin, cout, ks, stride, pad, Ho, Wo, relu, x2, H2, W2, cin2, stride2);
    HF_LAUNCH_CHECK();
    return HF_OK;
}

}  // namespace

struct hf_encoder {
    std::vector<hf_enc_op> ops;
    std::vector<__nv_bfloat16*> w;        // device weights (tcgen05 packing: stem repacked to pixel pairs)
    std::vector<__nv_bfloat16*> w_plain;  // device weights in plain (cout,k,k,cin) order for the SIMT path
    std::vector<float*> bias;
    std::vector<int> w_cin;               // cin of each weight as given
    int in_channels, stem_cin, feat_dim, impl;
    int stem_cp;                          // channels per pixel of the staged stem input actually in use (24 or 32)
    __nv_bfloat16 *w_stem24, *w_plain24;  // compact-stem weight packings (cout,7,8,24) / (cout,7,7,24)
    int debug_stop;                       // >= 0: forward stops after this op (layer-by-layer bring-up)
    // plan cache for one (B,H,W,workspace)
    int pB, pH, pW;
    void* pws;
    std::vector<ConvPlan> plans;
};

namespace {
constexpr int STEM_CP = 32;

struct BufShape { int H, W, C; };

// walk the op list to derive every op's input / output activation shape (buffer ids are reused along the way)
struct OpShapes { std::vector<BufShape> in, out, in2; };

int infer_shapes(const hf_encoder* h, int B, int H, int W, OpShapes& os, size_t* max_act) {
    std::map<int, BufShape> cur;
    os.in.assign(h->ops.size(), BufShape{0, 0, 0});
    os.out.assign(h->ops.size(), BufShape{0, 0, 0});
    os.in2.assign(h->ops.size(), BufShape{0, 0, 0});
    size_t mx = 0;
    for (size_t i = 0; i < h->ops.size(); ++i) {
        const hf_enc_op& op = h->ops[i];
        BufShape in;
        if (i == 0) in = {H, W, h->stem_cin};
        else {
            if (!cur.count(op.src)) return hf::fail(HF_ERR_INVALID, "encoder program: op %zu reads unwritten buffer %d", i, op.src);
            in = cur[op.src];
        }
        BufShape out = in;
        if (op.kind == HF_OP_CONV) {
            out.H = (in.H + 2 * op.pad - op.ksize) / op.stride + 1;
            out.W = (in.W + 2 * op.pad - op.ksize) / op.stride + 1;
            out.C = op.cout;
            if (i > 0 && op.cin != in.C) return hf::fail(HF_ERR_INVALID, "encoder program: op %zu cin %d != buffer channels %d", i, op.cin, in.C);
            if (op.res >= 0) {
                if (!cur.count(op.res)) return hf::fail(HF_ERR_INVALID, "encoder program: op %zu adds unwritten buffer %d", i, op.res);
                const BufShape r = cur[op.res];
                if (r.H != out.H || r.W != out.W || r.C != out.C) return hf::fail(HF_ERR_INVALID, "encoder program: op %zu residual shape mismatch", i);
                if (op.res == op.dst) return hf::fail(HF_ERR_INVALID, "encoder program: op %zu writes over its residual", i);
            }
            if (op.src == op.dst) return hf::fail(HF_ERR_INVALID, "encoder program: op %zu is in-place", i);
            if (op.src2 >= 0) {
                if (!cur.count(op.src2)) return hf::fail(HF_ERR_INVALID, "encoder program: op %zu reads unwritten buffer %d", i, op.src2);
                const BufShape s2 = cur[op.src2];
                if (op.stride2 < 1 || s2.C != op.cin2 || (s2.H - 1) / op.stride2 + 1 != out.H || (s2.W - 1) / op.stride2 + 1 != out.W)
                    return hf::fail(HF_ERR_INVALID, "encoder program: op %zu fused 1x1 branch shape mismatch", i);
                if (op.src2 == op.dst) return hf::fail(HF_ERR_INVALID, "encoder program: op %zu writes over its second input", i);
                os.in2[i] = s2;
            }
        } else if (op.kind == HF_OP_MAXPOOL3x3S2) {
            out.H = (in.H + 2 - 3) / 2 + 1; out.W = (in.W + 2 - 3) / 2 + 1;
        }
        os.in[i] = in; os.out[i] = out;
        if (op.kind == HF_OP_GLOBAL_AVGPOOL) continue;
        cur[op.dst] = out;
        mx = std::max(mx, (size_t)B * out.H * out.W * out.C * 2);
    }
    *max_act = (mx + 1023) & ~(size_t)1023;
    return HF_OK;
}

int num_buffers(const hf_encoder* h) {
    int n = 0;
    for (auto& op : h->ops) { n = std::max(n, op.dst + 1); n = std::max(n, op.src + 1); n = std::max(n, op.res + 1); n = std::max(n, op.src2 + 1); }
    return n;
}

size_t stem_in_bytes(int B, int H, int W) { return (((size_t)B * (H + 6) * (W + 8) * STEM_CP * 2) + 1023) & ~(size_t)1023; }

}  // namespace

extern "C" int hf_encoder_create(hf_encoder_t** out, const hf_enc_op* ops, int num_ops, const uint16_t* const* weights,
                                 const float* const* bias, int num_weights, int in_channels, int stem_cin, int feat_dim) {
    if (!out || !ops || num_ops < 1) return hf::fail(HF_ERR_INVALID, "hf_encoder_create: null argument");
    if (ops[0].kind != HF_OP_CONV || ops[0].ksize != 7 || ops[0].stride != 2 || ops[0].pad != 3 || stem_cin != STEM_CP ||
        in_channels > STEM_CP)
        return hf::fail(HF_ERR_UNSUPPORTED, "hf_encoder_create: first op must be the 7x7/2 stem with <=32 input channels padded to 32");
    hf_encoder* h = new hf_encoder();
    h->ops.assign(ops, ops + num_ops);
    h->in_channels = in_channels; h->stem_cin = stem_cin; h->feat_dim = feat_dim; h->impl = 0; h->debug_stop = -1;
    h->stem_cp = STEM_CP; h->w_stem24 = nullptr; h->w_plain24 = nullptr;
    h->pB = h->pH = h->pW = 0; h->pws = nullptr;
    h->w.resize(num_weights); h->w_plain.resize(num_weights); h->bias.resize(num_weights); h->w_cin.resize(num_weights);
    for (int i = 0; i < num_ops; ++i) {
        const hf_enc_op& op = ops[i];
        if (op.kind != HF_OP_CONV) continue;
        const int wi = op.weight_index;
        if (wi < 0 || wi >= num_weights) { hf_encoder_destroy(h); return hf::fail(HF_ERR_INVALID, "hf_encoder_create: weight index %d", wi); }
        const size_t n = (size_t)op.cout * ((size_t)op.ksize * op.ksize * op.cin + (op.src2 >= 0 ? op.cin2 : 0));
        int rc;
        if ((rc = hf::upload((uint16_t**)&h->w_plain[wi], weights[wi], n))) { hf_encoder_destroy(h); return rc; }
        if ((rc = hf::upload(&h->bias[wi], bias[wi], (size_t)op.cout))) { hf_encoder_destroy(h); return rc; }
        h->w_cin[wi] = op.cin;
        if (i == 0) {
            // stem: (cout,7,7,32) -> (cout,7,8,32) with a zero 8th tap so that two taps x 32 ch form one 128-byte k-block
            std::vector<uint16_t> pk((size_t)op.cout * 7 * 8 * STEM_CP, 0);
            for (int co = 0; co < op.cout; ++co)
                for (int kh = 0; kh < 7; ++kh)
                    for (int kw = 0; kw < 7; ++kw)
                        memcpy(&pk[(((size_t)co * 7 + kh) * 8 + kw) * STEM_CP], &weights[wi][(((size_t)co * 7 + kh) * 7 + kw) * STEM_CP], STEM_CP * 2);
            if ((rc = hf::upload((uint16_t**)&h->w[wi], pk.data(), pk.size()))) { hf_encoder_destroy(h); return rc; }
            if (in_channels <= 24) {   // compact variants: drop the (all-zero) channels 24..31
                std::vector<uint16_t> p24((size_t)op.cout * 7 * 8 * 24, 0), q24((size_t)op.cout * 7 * 7 * 24, 0);
                for (int co = 0; co < op.cout; ++co)
                    for (int kh = 0; kh < 7; ++kh)
                        for (int kw = 0; kw < 7; ++kw) {
                            const uint16_t* src = &weights[wi][(((size_t)co * 7 + kh) * 7 + kw) * STEM_CP];
                            memcpy(&p24[(((size_t)co * 7 + kh) * 8 + kw) * 24], src, 24 * 2);
                            memcpy(&q24[(((size_t)co * 7 + kh) * 7 + kw) * 24], src, 24 * 2);
                        }
                if ((rc = hf::upload((uint16_t**)&h->w_stem24, p24.data(), p24.size()))) { hf_encoder_destroy(h); return rc; }
                if ((rc = hf::upload((uint16_t**)&h->w_plain24, q24.data(), q24.size()))) { hf_encoder_destroy(h); return rc; }
            }
        } else {
            h->w[wi] = h->w_plain[wi];
        }
    }
    *out = h;
    return HF_OK;
}

extern "C" void hf_encoder_destroy(hf_encoder_t* h) {
    if (!h) return;
    cudaFree(h->w_stem24); cudaFree(h->w_plain24);
    for (size_t i = 0; i < h->w.size(); ++i) {
        if (h->w[i] && h->w[i] != h->w_plain[i]) cudaFree(h->w[i]);
        if (h->w_plain[i]) cudaFree(h->w_plain[i]);
        if (h->bias[i]) cudaFree(h->bias[i]);
    }
    delete h;
}

extern "C" int hf_encoder_stem_channels(const hf_encoder_t* h) { return h ? h->stem_cp : 0; }

extern "C" int hf_encoder_set_impl(hf_encoder_t* h, int impl) {
    if (!h || impl < 0 || impl > 1) return hf::fail(HF_ERR_INVALID, "hf_encoder_set_impl: bad argument");
    h->impl = impl;
    return HF_OK;
}

extern "C" size_t hf_encoder_workspace_bytes(const hf_encoder_t* h, int B, int H, int W) {
    OpShapes shp;
    size_t max_act = 0;
    if (infer_shapes(h, B, H, W, shp, &max_act)) return 0;
    return stem_in_bytes(B, H, W) + (size_t)num_buffers(h) * max_act + 1024;
}

namespace {
// Geometry of one call: shapes of every op, workspace carve-up, launch plans (re-built when the geometry or the workspace
// changes).  The padded stem input lives at the start of the workspace; its border and channel padding are zeroed when the
// plans are built and must stay zero afterwards: the workspace contents belong to the encoder between calls with the same
// geometry (hf_encoder_invalidate forces a rebuild + re-zeroing, e.g. after the caller recycled the memory).
struct EncCall {
    OpShapes shp;
    size_t max_act, sib;
    uint8_t* ws;
    __nv_bfloat16* stem_in;
    int Hp, Wp;
};

int enc_prepare(hf_encoder* h, int B, int H, int W, void* workspace, size_t workspace_bytes, cudaStream_t stream, EncCall& c) {
    int rc = infer_shapes(h, B, H, W, c.shp, &c.max_act);
    if (rc) return rc;
    const size_t need = hf_encoder_workspace_bytes(h, B, H, W);
    if (workspace_bytes < need) return hf::fail(HF_ERR_INVALID, "hf_encoder: workspace too small (%zu < %zu)", workspace_bytes, need);
    c.ws = (uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    c.stem_in = (__nv_bfloat16*)c.ws;
    c.sib = stem_in_bytes(B, H, W);
    c.Hp = H + 6; c.Wp = W + 8;
    auto buf = [&](int id) { return (__nv_bfloat16*)(c.ws + c.sib + (size_t)id * c.max_act); };
    if (h->pB != B || h->pH != H || h->pW != W || h->pws != workspace) {
        h->plans.assign(h->ops.size(), ConvPlan());
        for (size_t i = 0; i < h->ops.size(); ++i) {
            const hf_enc_op& op = h->ops[i];
            if (op.kind != HF_OP_CONV) continue;
            if (i == 0) {
                rc = HF_ERR_UNSUPPORTED;
                if (h->w_stem24) {   // compact stem first; its sliding-window tensor map may be refused by the driver
                    rc = plan_conv(&h->plans[i], 2, c.stem_in, h->w_stem24, B, H, W, 24, op.cout, 7, 2, 3, op.relu, c.Hp, c.Wp, buf(op.dst), nullptr);
                    if (rc == HF_OK) h->stem_cp = 24;
                }
                if (rc != HF_OK) {
                    rc = plan_conv(&h->plans[i], 1, c.stem_in, h->w[op.weight_index], B, H, W, STEM_CP, op.cout, 7, 2, 3, op.relu, c.Hp, c.Wp, buf(op.dst), nullptr);
                    h->stem_cp = STEM_CP;
                }
            } else {
                const BufShape in = c.shp.in[i];
                if (op.cin % 64 != 0) return hf::fail(HF_ERR_UNSUPPORTED, "encoder: conv %zu has cin %d (must be a multiple of 64)", i, op.cin);
                const BufShape in2 = c.shp.in2[i];
                rc = plan_conv(&h->plans[i], 0, buf(op.src), h->w[op.weight_index], B, in.H, in.W, op.cin, op.cout, op.ksize, op.stride, op.pad, op.relu, 0, 0,
                               buf(op.dst), op.res >= 0 ? buf(op.res) : nullptr, op.src2 >= 0 ? buf(op.src2) : nullptr, in2.H, in2.W, op.cin2, op.stride2);
            }
            if (rc) return rc;
        }
        // zero the padded stem input once: borders and channel padding stay zero, the interior is rewritten per call
        HF_CUDA(cudaMemsetAsync(c.stem_in, 0, c.sib, stream));
        h->pB = B; h->pH = H; h->pW = W; h->pws = workspace;
    }
    return HF_OK;
}

int enc_run_ops(hf_encoder* h, const EncCall& c, int B, int H, int W, float* feats, cudaStream_t stream);
}  // namespace

extern "C" int hf_encoder_invalidate(hf_encoder_t* h) {
    if (!h) return hf::fail(HF_ERR_INVALID, "hf_encoder_invalidate: null handle");
    h->pB = h->pH = h->pW = 0; h->pws = nullptr;
    return HF_OK;
}

// input_kind: 0 = fp32 NCHW, 1 = bf16 NCHW, 2 = already staged (the caller filled hf_encoder_stem_input's buffer)
static int enc_forward_any(hf_encoder_t* h, const void* input, int input_kind, int B, int H, int W, float* feats, void* workspace,
                           size_t workspace_bytes, void* stream_) {
    if (!h || (!input && input_kind != 2) || !feats || !workspace) return hf::fail(HF_ERR_INVALID, "hf_encoder_forward: null argument");
    cudaStream_t stream = (cudaStream_t)stream_;
    EncCall c;
    int rc = enc_prepare(h, B, H, W, workspace, workspace_bytes, stream, c);
    if (rc) return rc;
    if (input_kind != 2) {
        const size_t total = (size_t)B * H * W;
        int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 32);
        // follows a memset on the first call: plain launch
        if (input_kind == 0) nchw_to_nhwc_pad_kernel<float><<<blocks, 256, 0, stream>>>((const float*)input, c.stem_in, B, h->in_channels, H, W, h->stem_cp, c.Hp, c.Wp, 3, 3);
        else nchw_to_nhwc_pad_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((const __nv_bfloat16*)input, c.stem_in, B, h->in_channels, H, W, h->stem_cp, c.Hp, c.Wp, 3, 3);
        HF_LAUNCH_CHECK();
    }
    return enc_run_ops(h, c, B, H, W, feats, stream);
}

extern "C" int hf_encoder_forward(hf_encoder_t* h, const float* input, int B, int H, int W, float* feats, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    return enc_forward_any(h, input, 0, B, H, W, feats, workspace, workspace_bytes, stream);
}
extern "C" int hf_encoder_forward_bf16(hf_encoder_t* h, const uint16_t* input, int B, int H, int W, float* feats, void* workspace,
                                       size_t workspace_bytes, void* stream) {
    return enc_forward_any(h, input, 1, B, H, W, feats, workspace, workspace_bytes, stream);
}
extern "C" int hf_encoder_forward_staged(hf_encoder_t* h, int B, int H, int W, float* feats, void* workspace, size_t workspace_bytes,
                                         void* stream) {
    return enc_forward_any(h, nullptr, 2, B, H, W, feats, workspace, workspace_bytes, stream);
}
extern "C" int hf_encoder_stem_input(hf_encoder_t* h, int B, int H, int W, void* workspace, size_t workspace_bytes, void* stream,
                                     void** staged, int* dims) {
    if (!h || !workspace || !staged || !dims) return hf::fail(HF_ERR_INVALID, "hf_encoder_stem_input: null argument");
    EncCall c;
    int rc = enc_prepare(h, B, H, W, workspace, workspace_bytes, (cudaStream_t)stream, c);
    if (rc) return rc;
    *staged = c.stem_in;
    dims[0] = c.Hp; dims[1] = c.Wp; dims[2] = h->stem_cp; dims[3] = 3; dims[4] = 3;      // {Hp, Wp, Cp, top, left}
    return HF_OK;
}

namespace {
int enc_run_ops(hf_encoder* h, const EncCall& c, int B, int H, int W, float* feats, cudaStream_t stream) {
    const OpShapes& shp = c.shp;
    const size_t max_act = c.max_act, sib = c.sib;
    uint8_t* ws = c.ws;
    __nv_bfloat16* stem_in = c.stem_in;
    const int Hp = c.Hp, Wp = c.Wp;
    int rc = HF_OK;
    auto buf = [&](int id) { return (__nv_bfloat16*)(ws + sib + (size_t)id * max_act); };
    (void)H; (void)W;
    // profiling aid (HF_ENC_TIMING=1): CUDA events around every op of this call, per-op times printed to stderr
    static const bool op_timing = getenv("HF_ENC_TIMING") != nullptr;
    std::vector<cudaEvent_t> evs;
    auto mark = [&]() { if (op_timing) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, stream); evs.push_back(e); } };
    mark();
    for (size_t i = 0; i < h->ops.size(); ++i) {
        const hf_enc_op& op = h->ops[i];
        if (op.kind == HF_OP_CONV) {
            const __nv_bfloat16* res = op.res >= 0 ? buf(op.res) : nullptr;
            if (h->impl == 0) {
                rc = launch_conv(h->plans[i], h->bias[op.weight_index], stream);
            } else if (i == 0) {
                // SIMT path reads the physically padded input as a pad-0 convolution with the true output size
                rc = launch_simt(stem_in, h->stem_cp == 24 ? h->w_plain24 : h->w_plain[op.weight_index], h->bias[op.weight_index], res,
                                 buf(op.dst), B, Hp, Wp, h->stem_cp, op.cout, 7, 2, 0, op.relu, stream, h->plans[i].g.Ho, h->plans[i].g.Wo);
            } else {
                const BufShape in = shp.in[i];
                const BufShape in2 = shp.in2[i];
                rc = launch_simt(buf(op.src), h->w_plain[op.weight_index], h->bias[op.weight_index], res, buf(op.dst), B, in.H, in.W,
                                 op.cin, op.cout, op.ksize, op.stride, op.pad, op.relu, stream, 0, 0, op.src2 >= 0 ? buf(op.src2) : nullptr, in2.H,
                                 in2.W, op.cin2, op.stride2);
            }
            if (rc) return rc;
        } else if (op.kind == HF_OP_MAXPOOL3x3S2) {
            const BufShape in = shp.in[i], o = shp.out[i];
            const size_t total = (size_t)B * o.H * o.W * (in.C / 8);
            int blocks = (int)std::min<size_t>((total + 255) / 256, 148 * 32);
            maxpool_kernel<<<blocks, 256, 0, stream>>>(buf(op.src), buf(op.dst), B, in.H, in.W, in.C, o.H, o.W);
            HF_LAUNCH_CHECK();
        } else if (op.kind == HF_OP_GLOBAL_AVGPOOL) {
            const BufShape in = shp.in[i];
            dim3 grid(hf::div_up(in.C, 128), B);
            avgpool_kernel<<<grid, 128, 0, stream>>>(buf(op.src), feats, B, in.H * in.W, in.C);
            HF_LAUNCH_CHECK();
        }
        mark();
        if (h->debug_stop == (int)i) break;
    }
    if (op_timing) {
        cudaStreamSynchronize(stream);
        float tot = 0.f;
        for (size_t i = 0; i + 1 < evs.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, evs[i], evs[i + 1]);
            tot += ms;
            const hf_enc_op& op = h->ops[i];
            if (op.kind == HF_OP_CONV) {
                const ConvPlan& p = h->plans[i];
                const double gf = 2.0 * B * p.g.Ho * p.g.Wo * (double)op.cout * ((double)op.ksize * op.ksize * op.cin + (op.src2 >= 0 ? op.cin2 : 0)) * 1e-9;
                fprintf(stderr, "op %2zu conv k%d s%d %4d->%4d @%3dx%-3d bn %3d st %d sr %d res %d tiles %5d nkb %2d grid %3d : %7.2f us  %6.1f TF/s\n", i, op.ksize, op.stride,
                        op.cin + (op.src2 >= 0 ? op.cin2 : 0), op.cout, p.g.Ho, p.g.Wo, p.bn, p.stages, p.sr, p.has_res, p.num_tiles, p.g.nkb, p.grid.x, ms * 1e3, gf / ms);
            } else fprintf(stderr, "op %2zu kind %d : %7.2f us\n", i, op.kind, ms * 1e3);
        }
        fprintf(stderr, "encoder ops total %.1f us\n", tot * 1e3);
        for (auto e : evs) cudaEventDestroy(e);
    }
    return HF_OK;
}
}  // namespace

// Bring-up aid: run the program up to and including op `op_index`, then copy that op's output activation
// (bf16 NHWC) to `out` and report its dims {H, W, C}.
extern "C" int hf_encoder_debug_op_output(hf_encoder_t* h, int op_index, const float* input, int B, int H, int W,
                                          void* workspace, size_t workspace_bytes, uint16_t* out, size_t out_bytes,
                                          int* dims, float* feats, void* stream_) {
    if (!h || op_index < 0 || op_index >= (int)h->ops.size()) return hf::fail(HF_ERR_INVALID, "debug: bad op index");
    OpShapes shp;
    size_t max_act = 0;
    int rc = infer_shapes(h, B, H, W, shp, &max_act);
    if (rc) return rc;
    h->debug_stop = op_index;
    rc = hf_encoder_forward(h, input, B, H, W, feats, workspace, workspace_bytes, stream_);
    h->debug_stop = -1;
    if (rc) return rc;
    const hf_enc_op& op = h->ops[op_index];
    if (op.kind == HF_OP_GLOBAL_AVGPOOL) { dims[0] = dims[1] = 1; dims[2] = h->feat_dim; return HF_OK; }
    const BufShape o = shp.out[op_index];
    dims[0] = o.H; dims[1] = o.W; dims[2] = o.C;
    const size_t bytes = (size_t)B * o.H * o.W * o.C * 2;
    if (out_bytes < bytes) return hf::fail(HF_ERR_INVALID, "debug: output buffer too small");
    uint8_t* ws = (uint8_t*)(((uintptr_t)workspace + 1023) & ~(uintptr_t)1023);
    HF_CUDA(cudaMemcpyAsync(out, ws + stem_in_bytes(B, H, W) + (size_t)op.dst * max_act, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream_));
    return HF_OK;
}

extern "C" int hf_debug_conv_stamps(unsigned long long* out16) {
    HF_CUDA(cudaDeviceSynchronize());
    HF_CUDA(cudaMemcpyFromSymbol(out16, g_conv_ts, sizeof(unsigned long long) * 16));
    return HF_OK;
}

extern "C" int hf_conv2d_nhwc(const uint16_t* x, const uint16_t* w, const float* bias, const uint16_t* res, uint16_t* y,
                              int B, int H, int W, int cin, int cout, int ksize, int stride, int pad, int relu, int impl,
                              void* stream) {
    if (!x || !w || !bias || !y) return hf::fail(HF_ERR_INVALID, "hf_conv2d_nhwc: null argument");
    if (impl == 1)
        return launch_simt((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias, (const __nv_bfloat16*)res,
                           (__nv_bfloat16*)y, B, H, W, cin, cout, ksize, stride, pad, relu, (cudaStream_t)stream);
    ConvPlan p;
    int rc = plan_conv(&p, 0, x, w, B, H, W, cin, cout, ksize, stride, pad, relu, 0, 0, y, res);
    if (rc) return rc;
    return launch_conv(p, bias, (cudaStream_t)stream);
}

extern "C" int hf_conv2d_nhwc_branch(const uint16_t* x, const uint16_t* w, const float* bias, const uint16_t* res, uint16_t* y,
                                     int B, int H, int W, int cin, int cout, int ksize, int stride, int pad, int relu,
                                     const uint16_t* x2, int H2, int W2, int cin2, int stride2, int impl, void* stream) {
    if (!x || !w || !bias || !y || !x2) return hf::fail(HF_ERR_INVALID, "hf_conv2d_nhwc_branch: null argument");
    if (impl == 1)
        return launch_simt((const __nv_bfloat16*)x, (const __nv_bfloat16*)w, bias, (const __nv_bfloat16*)res,
                           (__nv_bfloat16*)y, B, H, W, cin, cout, ksize, stride, pad, relu, (cudaStream_t)stream, 0, 0,
                           (const __nv_bfloat16*)x2, H2, W2, cin2, stride2);
    ConvPlan p;
    int rc = plan_conv(&p, 0, x, w, B, H, W, cin, cout, ksize, stride, pad, relu, 0, 0, y, res, x2, H2, W2, cin2, stride2);
    if (rc) return rc;
    return launch_conv(p, bias, (cudaStream_t)stream);
}
