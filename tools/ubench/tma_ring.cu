// TMA -> smem ring -> tcgen05.mma throughput of ONE role pair per CTA (the conv mainloop skeleton), B200 sm_100a.
//   producer thread: wait(empty[s]); expect_tx(full[s]); TMA A box (4-D NHWC, 128 px x 64 ch = 16 KB) + TMA B box (2-D weights, BN x 64)
//   consumer thread: wait(full[s]); NMMA x tcgen05.mma (128 x BN x 16); tcgen05.commit(empty[s])
// Reports cycles per k-block for grid = 1 and grid = 148 and the aggregate L2->SM bandwidth.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/tma_ring tools/ubench/tma_ring.cu
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void umma_commit(uint32_t bar) { asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}

// C = channels of the activation tensor (B,64,64,C); weights (256 rows, K = 9*C).  KS = 64-wide k-blocks per stage.
template <int BN, int S, int KS, bool DO_A, bool DO_B, bool DO_MMA>
__global__ void __launch_bounds__(128, 1)
ring(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, long long* out, int iters, int C, int nimg) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * S + 1];
    __shared__ uint32_t tmem_base_s;
    constexpr int A_BYTES = 16384 * KS, B_BYTES = BN * 128 * KS, STAGE = A_BYTES + B_BYTES;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[S]), done = smem_u32(&bars[2 * S]);
    if (threadIdx.x == 0) { for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); } mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x >= 32 && threadIdx.x < 64) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int cpb = C / 64;
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        int tile = blockIdx.x;
        for (int i = 0; i < iters; ++i) {
            const uint32_t s = i % S, ph = (i / S) & 1;
            wait(empty0 + 8 * s, ph ^ 1);
            const uint32_t sa = base + s * STAGE, fb = full0 + 8 * s;
            mbar_expect_tx(fb, (DO_A ? A_BYTES : 0) + (DO_B ? B_BYTES : 0));
#pragma unroll
            for (int q = 0; q < KS; ++q) {
                const int kb = (i * KS + q) % (9 * cpb);                 // walk the 3x3 taps / channel blocks like the conv
                if (kb == 0 && q == 0 && i) tile += gridDim.x;
                const int tap = kb / cpb, cb = kb - tap * cpb;
                const int t = tile % (32 * nimg);                        // 32 tiles of 16 x 8 pixels per 64 x 64 image
                const int tw = t & 3, th = (t >> 2) & 7, tb = t >> 5;
                if (DO_A) tma_load_4d(sa + q * 16384, &mapA, fb, cb * 64, tw * 16 + tap % 3 - 1, th * 8 + tap / 3 - 1, tb);
                if (DO_B) tma_load_2d(sa + A_BYTES + q * BN * 128, &mapB, fb, kb * 64, 0);
            }
        }
        out[blockIdx.x * 2] = clock64() - t0;
    } else if (threadIdx.x == 32) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t d = tmem_base_s;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t s = i % S, ph = (i / S) & 1;
            wait(full0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (DO_MMA) {
                const uint32_t sa = base + s * STAGE;
#pragma unroll
                for (int q = 0; q < KS; ++q) {
                    const uint64_t da = umma_desc_sw128(sa + q * 16384), db = umma_desc_sw128(sa + A_BYTES + q * BN * 128);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, 1u);
                }
            }
            umma_commit(empty0 + 8 * s);
        }
        umma_commit(done);
        wait(done, 0);
        out[blockIdx.x * 2 + 1] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x >= 32 && threadIdx.x < 64) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(256u) : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_load_im2col_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int n, uint16_t ow, uint16_t oh) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2], {%7, %8};"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(n), "h"(ow), "h"(oh) : "memory");
}
// A-tile loads only (16 KB per k-block), VAR selects the box / mode; NSPLIT loads of 16 KB / NSPLIT each
template <int VAR, int S>
__global__ void __launch_bounds__(128, 1)
aload(const __grid_constant__ CUtensorMap map, long long* out, int iters, int C) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * S];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[S]);
    if (threadIdx.x == 0) { for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    const int cpb = C / 64;
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        int tile = blockIdx.x, cb = -1, dw = -1, dh = -1;
        for (int i = 0; i < iters; ++i) {
            const uint32_t s = i % S, ph = (i / S) & 1;
            wait(empty0 + 8 * s, ph ^ 1);
            const uint32_t sa = base + s * 16384, fb = full0 + 8 * s;
            mbar_expect_tx(fb, 16384);
            // incremental (tap, channel block) walk: no integer divisions in the loop
            if (++cb == cpb) { cb = 0; ++dw; if (dw == 2) { dw = -1; ++dh; if (dh == 2) { dh = -1; tile += gridDim.x; if (tile >= 1024) tile -= 1024; } } }
            const int t = tile;
            const int tw = t & 3, th = (t >> 2) & 7, tb = t >> 5;
            const int tap = (dh + 1) * 3 + dw + 1;
            if (VAR == 0) tma_load_4d(sa, &map, fb, cb * 64, tw * 16 + dw, th * 8 + dh, tb);                 // {64,16,8,1}
            if (VAR == 1) tma_load_3d(sa, &map, fb, cb * 64, tw * 16 + dw, tb * 64 + th * 8 + dh);           // 3-D (C, W, H*B) {64,16,8}
            if (VAR == 2) tma_load_2d(sa, &map, fb, cb * 64, (tb * 4096 + th * 512 + tw * 128) + dw + 64 * dh);   // 2-D (C, pixels) {64,128}
            if (VAR == 3) tma_load_4d(sa, &map, fb, cb * 64, dw, (t & 31) * 2 + dh, tb);                     // {64,64,2,1} full-width rows
            if (VAR == 4) { tma_load_4d(sa, &map, fb, cb * 64, tw * 16 + dw, th * 8 + dh, tb); tma_load_4d(sa + 8192, &map, fb, cb * 64, tw * 16 + dw, th * 8 + 4 + dh, tb); }  // 2 x {64,16,4,1}
            if (VAR == 5) tma_load_im2col_4d(sa, &map, fb, cb * 64, tw * 16 - 1, th * 8 - 1, tb, (uint16_t)(tap % 3), (uint16_t)(tap / 3));   // im2col, 128 pixels
        }
        out[blockIdx.x * 2] = clock64() - t0;
    } else if (threadIdx.x == 32) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t s = i % S, ph = (i / S) & 1;
            wait(full0 + 8 * s, ph);
            mbar_arrive(empty0 + 8 * s);
        }
        out[blockIdx.x * 2 + 1] = clock64() - t0;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int BN, int S, int KS, bool DO_A, bool DO_B, bool DO_MMA>
void run(const char* name, const CUtensorMap& mA, const CUtensorMap& mB, long long* d, int grid, int C, int nimg) {
    const int iters = 2000 / KS;
    const size_t smem = (size_t)S * (16384 * KS + BN * 128 * KS) + 1024;
    auto k = ring<BN, S, KS, DO_A, DO_B, DO_MMA>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<<<grid, 128, smem>>>(mA, mB, d, iters, C, nimg);      // warm
    cudaEventRecord(a);
    k<<<grid, 128, smem>>>(mA, mB, d, iters, C, nimg);
    cudaEventRecord(b);
    cudaError_t e = cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[2 * 148]; cudaMemcpy(h, d, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
    double mx = 0; for (int i = 0; i < grid; ++i) mx = h[2 * i + 1] > mx ? h[2 * i + 1] : mx;
    const double bytes = (double)grid * iters * ((DO_A ? 16384.0 * KS : 0) + (DO_B ? BN * 128.0 * KS : 0));
    printf("%-44s grid %3d: %7.1f cyc/k-block (64)  kernel %7.1f us  L2->SM %6.2f TB/s  tensor util %4.0f%%  %s\n", name, grid, mx / iters / KS, ms * 1e3,
           bytes / (ms * 1e-3) / 1e12, DO_MMA ? 100.0 * (128.0 * BN / 256 * 4) / (mx / iters / KS) : 0.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    const int C = 256, NIMG = 32, H = 64, W = 64, COUT = 256;
    void *x, *w; long long* d;
    cudaMalloc(&x, (size_t)NIMG * H * W * C * 2); cudaMalloc(&w, (size_t)COUT * 9 * C * 2); cudaMalloc(&d, 8 * 2 * 148);
    cudaMemset(x, 0, (size_t)NIMG * H * W * C * 2); cudaMemset(w, 0, (size_t)COUT * 9 * C * 2);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    CUtensorMap mA, mB64, mB128, mB256;
    {
        cuuint64_t gd[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NIMG}, gs[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t bx[4] = {64, 16, 8, 1}, es[4] = {1, 1, 1, 1};
        enc(&mA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    auto encB = [&](CUtensorMap* m, int bn) {
        cuuint64_t gd[2] = {(cuuint64_t)9 * C, (cuuint64_t)COUT}, gs[1] = {(cuuint64_t)9 * C * 2};
        cuuint32_t bx[2] = {64, (cuuint32_t)bn}, es[2] = {1, 1};
        enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, w, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    encB(&mB64, 64); encB(&mB128, 128); encB(&mB256, 256);
    for (int grid : {1, 148}) {
        run<128, 6, 1, false, false, false>("empty handshake S=6", mA, mB128, d, grid, C, NIMG);
        run<128, 6, 1, true, true, false>("loads only BN=128 S=6", mA, mB128, d, grid, C, NIMG);
        run<128, 6, 1, true, false, false>("A loads only S=6", mA, mB128, d, grid, C, NIMG);
        run<128, 6, 1, false, true, false>("B loads only BN=128 S=6", mA, mB128, d, grid, C, NIMG);
        run<128, 6, 1, false, false, true>("MMA only BN=128 S=6", mA, mB128, d, grid, C, NIMG);
        run<128, 6, 1, true, true, true>("loads + MMA BN=128 S=6 KS=1", mA, mB128, d, grid, C, NIMG);
        run<128, 3, 2, true, true, true>("loads + MMA BN=128 S=3 KS=2", mA, mB128, d, grid, C, NIMG);
        run<256, 4, 1, true, true, true>("loads + MMA BN=256 S=4 KS=1", mA, mB256, d, grid, C, NIMG);
        run<256, 2, 2, true, true, true>("loads + MMA BN=256 S=2 KS=2", mA, mB256, d, grid, C, NIMG);
        run<64, 8, 1, true, true, true>("loads + MMA BN=64 S=8 KS=1", mA, mB64, d, grid, C, NIMG);
        run<64, 4, 2, true, true, true>("loads + MMA BN=64 S=4 KS=2", mA, mB64, d, grid, C, NIMG);
    }
    {   // A-tile load variants
        typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const int*, const int*,
                                           cuuint32_t, cuuint32_t, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* p2 = nullptr;
        cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &p2, cudaEnableDefault, &q);
        EncodeIm2colFn enc2 = (EncodeIm2colFn)p2;
        CUtensorMap m3, m2, m4w, m4h, mi;
        cuuint32_t es[4] = {1, 1, 1, 1};
        { cuuint64_t gd[3] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H * NIMG}, gs[2] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2}; cuuint32_t bx[3] = {64, 16, 8};
          enc(&m3, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, x, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
        { cuuint64_t gd[2] = {(cuuint64_t)C, (cuuint64_t)W * H * NIMG}, gs[1] = {(cuuint64_t)C * 2}; cuuint32_t bx[2] = {64, 128};
          enc(&m2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, x, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
        cuuint64_t gd4[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)NIMG}, gs4[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        { cuuint32_t bx[4] = {64, 64, 2, 1};
          enc(&m4w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, gd4, gs4, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
        { cuuint32_t bx[4] = {64, 16, 4, 1};
          enc(&m4h, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, gd4, gs4, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
        CUresult ri = CUDA_ERROR_UNKNOWN;
        if (enc2) { int lo[2] = {-1, -1}, hi[2] = {-1, -1};     // 3x3, pad 1: bounding box corners of the filter's top-left positions
          ri = enc2(&mi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, gd4, gs4, lo, hi, 64, 128, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); }
        printf("im2col encode: %d\n", (int)ri);
        auto runA = [&](const char* name, auto kern, const CUtensorMap& m, int grid) {
            const int iters = 2000; const size_t smem = 13 * 16384 + 1024;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            kern<<<grid, 128, smem>>>(m, d, iters, C);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2 * 148]; cudaMemcpy(h, d, sizeof(long long) * 2 * grid, cudaMemcpyDeviceToHost);
            double mx = 0; for (int i = 0; i < grid; ++i) mx = h[2 * i + 1] > mx ? h[2 * i + 1] : mx;
            printf("%-44s grid %3d: %7.1f cyc / 16 KB A tile  %s\n", name, grid, mx / iters, e == cudaSuccess ? "" : cudaGetErrorString(e));
        };
        for (int grid : {1, 148}) {
            runA("A 4-D {64,16,8,1}", aload<0, 6>, mA, grid);
            runA("A 3-D (C,W,H*B) {64,16,8}", aload<1, 6>, m3, grid);
            runA("A 2-D (C,pixels) {64,128}", aload<2, 6>, m2, grid);
            runA("A 4-D {64,64,2,1}", aload<3, 6>, m4w, grid);
            runA("A 2 x 4-D {64,16,4,1}", aload<4, 6>, m4h, grid);
            if (ri == CUDA_SUCCESS) runA("A im2col 128 px", aload<5, 6>, mi, grid);
            runA("A 4-D {64,16,8,1} S=2", aload<0, 2>, mA, grid);
            runA("A 4-D {64,16,8,1} S=3", aload<0, 3>, mA, grid);
            runA("A 4-D {64,16,8,1} S=9", aload<0, 9>, mA, grid);
            runA("A 4-D {64,16,8,1} S=12", aload<0, 12>, mA, grid);
            runA("A 2-D S=12", aload<2, 12>, m2, grid);
        }
    }
    printf("status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
