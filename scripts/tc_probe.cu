// scripts/tc_probe.cu -- developer probe (not product): which shared-memory layout / descriptor fields does
// tcgen05.mma.kind::tf32 expect?  One CTA, one MMA (M=128, N=16, K=8) per hypothesis, exact small-integer data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/tc_probe scripts/tc_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type = 0) {
    uint64_t d = (uint64_t)(layout_type & 7) << 61;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

struct Hyp {
    int a_major, b_major;       // 0 = K-major, 1 = MN-major
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    uint32_t swz;               // descriptor layout_type: 0 none, 2 = SWIZZLE_128B
};

// A: 128 x 8, B: 16 x 8 (N x K), values given in plain row-major arrays; `layout` decides the smem placement
__global__ void probe_kernel(const float *A, const float *B, float *D, Hyp pl, Hyp h, int layout_a, int layout_b, int *status) {
    extern __shared__ __align__(1024) char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    char *sa = smem, *sb = smem + 16384;
    for (int i = tid; i < 32768 / 4; i += blockDim.x) ((float *)smem)[i] = 0.f;
    __syncthreads();
    // element (mn, k) -> byte offset
    auto off = [](int layout, int mn, int k, uint32_t lbo, uint32_t sbo) -> uint32_t {
        if (layout == 2) { // MN-major, 128B swizzle: atom = 8 k-rows x 128 B (32 mn); 16B chunk index ^= k%8
            const uint32_t c = (uint32_t)(mn % 32) / 4, r = (uint32_t)(k % 8);
            return (uint32_t)(mn / 32) * lbo + (uint32_t)(k / 8) * sbo + r * 128u + ((c ^ r) * 16u) + (uint32_t)(mn % 4) * 4u;
        }
        if (layout == 1)   // MN-major canonical, no swizzle: 16B chunk = 4 consecutive mn, 8 k-rows per core matrix
            return (uint32_t)(mn / 4) * sbo + (uint32_t)(k % 8) * 16u + (uint32_t)(mn % 4) * 4u + (uint32_t)(k / 8) * lbo;
        // K-major canonical, no swizzle: core matrix = 8 mn-rows x 16 B (4 k)
        return (uint32_t)(mn / 8) * sbo + (uint32_t)(mn % 8) * 16u + (uint32_t)(k / 4) * lbo + (uint32_t)(k % 4) * 4u;
    };
    for (int i = tid; i < 128 * 8; i += blockDim.x) {
        const int m = i / 8, k = i % 8;
        *(float *)(sa + off(layout_a, m, k, pl.a_lbo, pl.a_sbo)) = A[i];
    }
    for (int i = tid; i < 16 * 8; i += blockDim.x) {
        const int n = i / 8, k = i % 8;
        *(float *)(sb + off(layout_b, n, k, pl.b_lbo, pl.b_sbo)) = B[i];
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    if (tid == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)h.a_major << 15) | ((uint32_t)h.b_major << 16) |
                               ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint64_t da = make_desc(smem_u32(sa), h.a_lbo, h.a_sbo, h.swz), db = make_desc(smem_u32(sb), h.b_lbo, h.b_sbo, h.swz);
        asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                     ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(0u) : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    if (warp < 4) {
        uint32_t ok = 0;
        for (int it = 0; it < (1 << 20) && !ok; it++)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.b32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
        if (!ok && lane == 0) atomicExch(status, 1);
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 16; i++) D[(warp * 32 + lane) * 16 + i] = __uint_as_float(r[i]);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
    }
}

int main() {
    float hA[128 * 8], hB[16 * 8], ref[128 * 16], hD[128 * 16];
    srand(1);
    for (int i = 0; i < 128 * 8; i++) hA[i] = (float)(rand() % 7 - 3);
    for (int i = 0; i < 16 * 8; i++) hB[i] = (float)(rand() % 5 - 2);
    for (int m = 0; m < 128; m++)
        for (int n = 0; n < 16; n++) {
            float s = 0;
            for (int k = 0; k < 8; k++) s += hA[m * 8 + k] * hB[n * 8 + k];
            ref[m * 16 + n] = s;
        }
    float *dA, *dB, *dD;
    int *dS;
    cudaMalloc(&dA, sizeof hA); cudaMalloc(&dB, sizeof hB); cudaMalloc(&dD, sizeof hD); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, hA, sizeof hA, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, sizeof hB, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    struct { const char *name; int la, lb; Hyp h; } T[] = {
        // MN-major: chunks of 4 mn; mn-block stride X, k-block stride Y
        {"MN-major  LBO=kblk(128)  SBO=mnblk(128)", 1, 1, {1, 1, 128, 128, 128, 128}},
        {"MN-major  LBO=kblk(4096) SBO=mnblk(128)", 1, 1, {1, 1, 4096, 128, 4096, 128}},
        {"MN-major  fields swapped (LBO=mnblk 128, SBO=kblk 4096) data as row above", 1, 1, {1, 1, 128, 4096, 128, 4096}},
        {"MN-major  LBO=kblk(128)  SBO=mnblk(1024)", 1, 1, {1, 1, 128, 1024, 128, 1024}},
        // K-major: core matrix 8 mn x 16B; mn-group stride SBO, k-chunk stride LBO
        {"K-major   SBO=mngrp(128) LBO=kchunk(2048)", 0, 0, {0, 0, 2048, 128, 2048, 128}},
        {"K-major   SBO=mngrp(256) LBO=kchunk(128)", 0, 0, {0, 0, 128, 256, 128, 256}},
        {"K-major   fields swapped vs row above", 0, 0, {0, 0, 256, 128, 256, 128}},
        // MN-major 128B swizzle: LBO = stride between 32-element mn blocks, SBO = stride between 8-row k blocks
        {"MN-major SW128  LBO=mnblk(1024) SBO=kblk(4096)", 2, 2, {1, 1, 1024, 4096, 1024, 4096, 2}},
        {"MN-major SW128  fields swapped (desc LBO=kblk SBO=mnblk)", 2, 2, {1, 1, 4096, 1024, 4096, 1024, 2}},
        {"MN-major SW128  LBO=mnblk(2048) SBO=kblk(1024)", 2, 2, {1, 1, 2048, 1024, 2048, 1024, 2}},
        {"MN-major SW128  LBO=1 (ignored?) SBO=kblk(1024), mnblk stride 1024 assumed", 2, 2, {1, 1, 16, 1024, 16, 1024, 2}},
    };
    for (auto &t : T) {
        cudaMemset(dD, 0xff, sizeof hD);
        cudaMemset(dS, 0, 4);
        // placement always uses (lbo = k-block/chunk stride, sbo = mn-block stride) of the row's geometry;
        // "swapped" rows hand the two fields to the descriptor the other way round
        Hyp place = t.h, desc = t.h;
        if (strstr(t.name, "assumed")) { place.a_lbo = 1024; place.b_lbo = 1024; place.a_sbo = 4096; place.b_sbo = 4096; desc.a_sbo = 4096; desc.b_sbo = 4096; }
        if (strstr(t.name, "swapped")) { place.a_lbo = t.h.a_sbo; place.a_sbo = t.h.a_lbo; place.b_lbo = t.h.b_sbo; place.b_sbo = t.h.b_lbo; }
        probe_kernel<<<1, 256, 65536>>>(dA, dB, dD, place, desc, t.la, t.lb, dS);
        cudaError_t e = cudaDeviceSynchronize();
        int st = 0;
        cudaMemcpy(hD, dD, sizeof hD, cudaMemcpyDeviceToHost);
        cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
        int bad = 0, zeros = 0;
        for (int i = 0; i < 128 * 16; i++) { bad += hD[i] != ref[i]; zeros += hD[i] == 0.f; }
        printf("%-80s err=%s timeout=%d mismatches=%d/2048 zeros=%d  D[0][0..3]=%g %g %g %g  ref=%g %g %g %g\n", t.name,
               cudaGetErrorString(e), st, bad, zeros, hD[0], hD[1], hD[2], hD[3], ref[0], ref[1], ref[2], ref[3]);
        if (e != cudaSuccess) { printf("sticky error, stopping\n"); break; }
    }
    return 0;
}
