// Micro-benchmarks of the synchronisation primitives the tensor-core kernels are built from (B200, sm_100a):
//   1. mbarrier ping-pong between two single threads of different warps: try_wait (suspending) vs test_wait (spinning)
//   2. tcgen05.commit -> mbarrier wait latency with no MMA outstanding / after one 128x64x16 MMA
//   3. producer/consumer ring (S slots) of empty handshakes: cycles per iteration
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/sync_lat tools/ubench/sync_lat.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
template <int MODE>
__device__ __forceinline__ void wait(uint32_t bar, uint32_t parity) {
    if (MODE == 0) { while (!try_wait(bar, parity)) {} }
    else { while (!test_wait(bar, parity)) {} }
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int MODE>
__global__ void pingpong(long long* out, int iters) {
    __shared__ __align__(8) uint64_t bars[2];
    const uint32_t b0 = smem_u32(&bars[0]), b1 = smem_u32(&bars[1]);
    if (threadIdx.x == 0) { mbar_init(b0, 1); mbar_init(b1, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) { mbar_arrive(b0); wait<MODE>(b1, i & 1); }
        out[0] = clock64() - t0;
    } else if (threadIdx.x == 32) {
        for (int i = 0; i < iters; ++i) { wait<MODE>(b0, i & 1); mbar_arrive(b1); }
    }
}

// one thread: commit -> wait, serial
template <int MODE>
__global__ void commit_lat(long long* out, int iters) {
    __shared__ __align__(8) uint64_t bars[1];
    __shared__ uint32_t tmem_base_s;
    const uint32_t b0 = smem_u32(&bars[0]);
    if (threadIdx.x == 0) { mbar_init(b0, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) { umma_commit(b0); wait<MODE>(b0, i & 1); }
        out[0] = clock64() - t0;
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(64u) : "memory");
}

// producer (thread 0) / consumer (thread 32) ring of S slots, empty handshakes: producer waits empty[s], arrives full[s];
// consumer waits full[s], COMMIT ? tcgen05.commit(empty[s]) : arrive(empty[s])
template <int MODE, int S, bool COMMIT>
__global__ void ring(long long* out, int iters) {
    __shared__ __align__(8) uint64_t bars[2 * S];
    __shared__ uint32_t tmem_base_s;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[S]);
    if (threadIdx.x == 0) { for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x >= 32 && threadIdx.x < 64) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) { const uint32_t s = i % S, ph = (i / S) & 1; wait<MODE>(empty0 + 8 * s, ph ^ 1); mbar_arrive(full0 + 8 * s); }
        out[0] = clock64() - t0;
    } else if (threadIdx.x == 32) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t s = i % S, ph = (i / S) & 1;
            wait<MODE>(full0 + 8 * s, ph);
            if (COMMIT) umma_commit(empty0 + 8 * s); else mbar_arrive(empty0 + 8 * s);
        }
        out[1] = clock64() - t0;
    }
    __syncthreads();
    if (threadIdx.x >= 32 && threadIdx.x < 64) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(64u) : "memory");
}

// ring of empty handshakes with NP extra single-lane pollers (lane 0 of warps 2..2+NP) parked in try_wait on a barrier that
// completes only at the end (like epilogue warps waiting for an accumulator), ALL lanes polling if WARPPOLL
template <int MODE, int S, int NP, bool WARPPOLL, bool PADBARS>
__global__ void ring_pollers(long long* out, int iters) {
    __shared__ __align__(128) uint64_t bars[(2 * S + 1) * (PADBARS ? 16 : 1)];
    constexpr int ST = PADBARS ? 128 : 8;
    __shared__ uint32_t tmem_base_s;
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = full0 + S * ST, done = full0 + 2 * S * ST;
    if (threadIdx.x == 0) { for (int s = 0; s < S; ++s) { mbar_init(full0 + ST * s, 1); mbar_init(empty0 + ST * s, 1); } mbar_init(done, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x >= 32 && threadIdx.x < 64) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) { const uint32_t s = i % S, ph = (i / S) & 1; wait<MODE>(empty0 + ST * s, ph ^ 1); mbar_arrive(full0 + ST * s); }
        out[0] = clock64() - t0;
    } else if (threadIdx.x == 32) {
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
            const uint32_t s = i % S, ph = (i / S) & 1;
            wait<MODE>(full0 + ST * s, ph);
            umma_commit(empty0 + ST * s);
        }
        out[1] = clock64() - t0;
        mbar_arrive(done);
    } else if (warp >= 2 && warp < 2 + NP && (WARPPOLL || lane == 0)) {
        while (!try_wait(done, 0)) {}
    }
    __syncthreads();
    if (threadIdx.x >= 32 && threadIdx.x < 64) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(64u) : "memory");
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
// one thread issues `iters` x PER MMAs (128 x N x 16, bf16) back to back on a fixed smem tile, one commit per PER MMAs
template <int N, int PER>
__global__ void mma_issue(long long* out, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[1];
    __shared__ uint32_t tmem_base_s;
    const uint32_t b0 = smem_u32(&bars[0]);
    const uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
    for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(b0, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = umma_desc_sw128(base), db = umma_desc_sw128(base + 16384);
        const uint32_t d = tmem_base_s;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int k = 0; k < PER; ++k) umma_bf16(d, da + (uint64_t)(2 * (k & 3)), db + (uint64_t)(2 * (k & 3)), idesc, 1u);
            umma_commit(b0);
        }
        long long t1 = clock64();
        wait<0>(b0, (iters - 1) & 1);
        out[0] = t1 - t0; out[1] = clock64() - t0;
    }
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base_s), "r"(256u) : "memory");
}

int main() {
    long long* d; cudaMalloc(&d, 64); long long h[2];
    const int N = 2000;
    auto rep = [&](const char* name) { cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("%-48s %8.1f cycles/iter (consumer %8.1f)\n", name, (double)h[0] / N, (double)h[1] / N); h[1] = 0; cudaMemset(d, 0, 16); };
    cudaMemset(d, 0, 16);
    pingpong<0><<<1, 64>>>(d, N); rep("pingpong try_wait (round trip)");
    pingpong<1><<<1, 64>>>(d, N); rep("pingpong test_wait spin (round trip)");
    commit_lat<0><<<1, 64>>>(d, N); rep("tcgen05.commit -> try_wait");
    commit_lat<1><<<1, 64>>>(d, N); rep("tcgen05.commit -> test_wait spin");
    ring<0, 4, false><<<1, 64>>>(d, N); rep("ring S=4 arrive/arrive try_wait");
    ring<1, 4, false><<<1, 64>>>(d, N); rep("ring S=4 arrive/arrive test_wait");
    ring<0, 4, true><<<1, 64>>>(d, N); rep("ring S=4 arrive/commit try_wait");
    ring<1, 4, true><<<1, 64>>>(d, N); rep("ring S=4 arrive/commit test_wait");
    ring<0, 8, true><<<1, 64>>>(d, N); rep("ring S=8 arrive/commit try_wait");
    ring<1, 8, true><<<1, 64>>>(d, N); rep("ring S=8 arrive/commit test_wait");
    ring_pollers<0, 6, 0, false, false><<<1, 416>>>(d, N); rep("ring S=6 commit, 416 thr, 0 pollers");
    ring_pollers<0, 6, 1, false, false><<<1, 416>>>(d, N); rep("ring S=6 commit, 1 lane-0 poller");
    ring_pollers<0, 6, 9, false, false><<<1, 416>>>(d, N); rep("ring S=6 commit, 9 lane-0 pollers");
    ring_pollers<0, 6, 9, true, false><<<1, 416>>>(d, N); rep("ring S=6 commit, 9 full-warp pollers");
    ring_pollers<0, 6, 9, false, true><<<1, 416>>>(d, N); rep("ring S=6 commit, 9 lane-0 pollers, padded bars");
    ring_pollers<1, 6, 9, false, false><<<1, 416>>>(d, N); rep("ring S=6 commit test_wait, 9 lane-0 try pollers");
    auto rep2 = [&](const char* name, int per) { cudaDeviceSynchronize(); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); printf("%-48s issue %7.1f  total %7.1f cycles per MMA\n", name, (double)h[0] / N / per, (double)h[1] / N / per); cudaMemset(d, 0, 16); };
    cudaFuncSetAttribute(mma_issue<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(mma_issue<64, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(mma_issue<128, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(mma_issue<256, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(mma_issue<128, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    mma_issue<8, 4><<<1, 128, 50 * 1024>>>(d, N); rep2("MMA 128x8x16, 4 per commit", 4);
    mma_issue<64, 4><<<1, 128, 50 * 1024>>>(d, N); rep2("MMA 128x64x16, 4 per commit", 4);
    mma_issue<128, 4><<<1, 128, 50 * 1024>>>(d, N); rep2("MMA 128x128x16, 4 per commit", 4);
    mma_issue<256, 4><<<1, 128, 50 * 1024>>>(d, N); rep2("MMA 128x256x16, 4 per commit", 4);
    mma_issue<128, 8><<<1, 128, 50 * 1024>>>(d, N); rep2("MMA 128x128x16, 8 per commit", 8);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
