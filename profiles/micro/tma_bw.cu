// Micro-benchmark: what does ONE SM's copy engine sustain for 1-D bulk copies (cp.async.bulk global -> shared)?
// Every CTA streams its own contiguous region of a large buffer through a ring of K stages, each stage filled by
// `tiles` copies of `bytes` bytes whose global addresses are offset by `mis` bytes from 128-byte alignment; one
// thread issues, nobody reads the data.   nvcc -arch=sm_100a -O3 -o build/micro/tma_bw profiles/micro/tma_bw.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128) k(const char *src, size_t per_cta, int bytes, int tiles, int stages, int mis, int iters) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(sm);
    unsigned char *buf = sm + 128;
    const int slot = (bytes + 127) / 128 * 128 + 128;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + s)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const char *base = src + (size_t)blockIdx.x * per_cta + mis;
    size_t pos = 0;
    for (int it = 0; it < iters + stages; ++it) {
        const int st = it % stages;
        if (it >= stages) {  // wait for the copy issued `stages` iterations ago
            uint32_t ok = 0, par = ((it / stages) - 1) & 1;
            while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(s32(bar + st)), "r"(par) : "memory");
        }
        if (it < iters) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + st)), "r"(bytes * tiles) : "memory");
            for (int t = 0; t < tiles; ++t) {
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                                 s32(buf + (size_t)(st * tiles + t) * slot)), "l"(base + pos), "r"(bytes), "r"(s32(bar + st)) : "memory");
                pos += (size_t)bytes + 4096 * 3;  // separate streams-ish: skip ahead so tiles are not contiguous
                if (pos + bytes + 256 > per_cta) pos = 0;
            }
        }
    }
}
int main() {
    const size_t per_cta = 48u << 20;  // 48 MB per CTA slice: far larger than L2 overall
    int nsm = 148;
    char *d; cudaMalloc(&d, per_cta * 2 * nsm + 4096); cudaMemset(d, 1, per_cta * 2 * nsm + 4096);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("ctas/SM bytes tiles stages misalign  GB/s   B/clk/SM(1.965GHz)\n");
    for (int cps = 1; cps <= 2; ++cps)
    for (int mis : {0, 16})
    for (int bytes : {2048, 4096, 8192, 16384})
    for (int tiles : {1, 8})
    for (int stages : {2, 4}) {
        size_t smem = 128 + (size_t)stages * tiles * ((bytes + 127) / 128 * 128 + 128);
        if (smem * cps > 220 * 1024 || smem > 227 * 1024) continue;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        int iters = (int)((64u << 20) / ((size_t)bytes * tiles)); if (iters > 4096) iters = 4096;
        k<<<nsm * cps, 128, smem>>>(d, per_cta, bytes, tiles, stages, mis, 8);
        cudaEventRecord(e0);
        k<<<nsm * cps, 128, smem>>>(d, per_cta, bytes, tiles, stages, mis, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double gb = (double)nsm * cps * iters * tiles * bytes / 1e9;
        cudaError_t err = cudaGetLastError();
        printf("%d %6d %2d %d %2d  %8.1f  %6.2f %s\n", cps, bytes, tiles, stages, mis, gb / (ms * 1e-3), gb * 1e9 / (ms * 1e-3) / nsm / 1.965e9, err ? cudaGetErrorString(err) : "");
    }
    return 0;
}
