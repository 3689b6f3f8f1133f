// HBM read-bandwidth microbenchmark: what can a pure read stream reach on this B200, for (a) a flat 16B/lane
// grid-stride sweep and (b) the megakernel's tile pattern (8 rows x 2 KB pieces, rows 8 KB apart) via TMA bulk
// copies into a smem ring? Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_read hbm_read.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void flat_read(const uint4* __restrict__ p, size_t n, unsigned long long* out) {
    unsigned long long acc = 0;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        uint4 a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p + i));
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p + i + stride));
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(p + i + 2 * stride));
        asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "l"(p + i + 3 * stride));
        acc += a.x ^ b.y ^ c.z ^ d.w;
    }
    if (acc == 0x123456789ull) *out = acc;
}

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// each CTA streams its contiguous share in `piece`-byte bulk copies, 8 per tile, ring of `stages` tiles
__global__ void __launch_bounds__(160) tma_read(const uint8_t* __restrict__ p, size_t bytes_per_cta, int piece, int row_stride,
                                                int stages, unsigned long long* out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * 8 * piece);
    uint64_t* empty = full + stages;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint8_t* base = p + (size_t)blockIdx.x * bytes_per_cta;
    // tile t covers 8 "rows": row r at base + (t / tiles_per_group) * 8 * row_stride + r * row_stride + (t % tiles_per_group) * piece
    const int tiles_per_group = row_stride / piece;
    const size_t ntiles = bytes_per_cta / (8 * (size_t)piece);
    if (warp >= 1) {  // 4 producer warps
        const int pw = warp - 1;
        for (size_t t = 0; t < ntiles; ++t) {
            if ((int)(t & 3) != pw) continue;
            const int st = t % stages;
            const uint32_t par = (t / stages) & 1;
            if (lane == 0) {
                uint32_t done = 0;
                while (!done) asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.u32 %0,1,0,P;}" : "=r"(done) : "r"(s32(&empty[st])), "r"(par ^ 1) : "memory");
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[st])), "r"(8 * piece) : "memory");
            }
            __syncwarp();
            if (lane < 8) {
                const uint8_t* src = base + (t / tiles_per_group) * 8 * (size_t)row_stride + (size_t)lane * row_stride + (t % tiles_per_group) * (size_t)piece;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(smem + ((size_t)st * 8 + lane) * piece)), "l"(src), "r"(piece), "r"(s32(&full[st])) : "memory");
            }
        }
    } else {  // consumer warp 0: wait + release immediately
        unsigned long long acc = 0;
        for (size_t t = 0; t < ntiles; ++t) {
            const int st = t % stages;
            const uint32_t par = (t / stages) & 1;
            uint32_t done = 0;
            while (!done) asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.u32 %0,1,0,P;}" : "=r"(done) : "r"(s32(&full[st])), "r"(par) : "memory");
            acc += smem[(size_t)st * 8 * piece + lane * 16];
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[st])) : "memory");
        }
        if (acc == 0x123456789ull) *out = acc;
    }
}

int main() {
    const size_t bytes = 8ull << 30;
    uint8_t* d;
    unsigned long long* out;
    cudaMalloc(&d, bytes);
    cudaMalloc(&out, 8);
    cudaMemset(d, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int thr : {256, 512, 1024}) for (int mult : {1, 2, 4}) {
        const int grid = 148 * mult * (1024 / thr);
        flat_read<<<grid, thr>>>((const uint4*)d, bytes / 16, out);
        cudaEventRecord(e0);
        flat_read<<<grid, thr>>>((const uint4*)d, bytes / 16, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("flat_read threads=%d grid=%d : %.1f GB/s\n", thr, grid, bytes / ms / 1e6);
    }
    cudaFuncSetAttribute(tma_read, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for (int piece : {2048, 4096, 8192}) for (int stages : {4, 8, 11}) {
        if ((size_t)stages * 8 * piece + 256 > 200 * 1024) continue;
        const int row_stride = 8192;
        const size_t per_cta = (bytes / 148) / (8 * (size_t)row_stride) * (8 * (size_t)row_stride);
        const size_t smem = (size_t)stages * 8 * piece + 256;
        tma_read<<<148, 160, smem>>>(d, per_cta, piece, row_stride, stages, out);
        cudaEventRecord(e0);
        tma_read<<<148, 160, smem>>>(d, per_cta, piece, row_stride, stages, out);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("tma_read piece=%d stages=%d (%.0f KB in flight/SM): %.1f GB/s  err=%s\n", piece, stages, stages * 8 * piece / 1024.0, per_cta * 148.0 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
