// Probe of the tensor-memory paths an attention kernel leans on (measured on the GPU box, numbers quoted in DESIGN.md):
//   A  tcgen05.ld.32x32b.x32 throughput per SM with 4 / 8 / 16 warps (bytes per clock)
//   B  MUFU.EX2 throughput per SM (8 / 16 warps)
//   C  tcgen05.st.32x32b.x32 throughput per SM
//   D  tcgen05.mma with the A operand IN TENSOR MEMORY (P of P V written by tcgen05.st as packed bf16 pairs, row = lane,
//      16 K elements = 8 columns): checked against a host reference
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I azula_b200/csrc scripts/tmem_probe.cu
//        -o build/tmem_probe   (run on the GPU box: build/tmem_probe)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "tc.cuh"

__device__ __forceinline__ void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ void tmem_ld_32x32b_x64(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
        "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
        "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
          "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
          "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
          "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
          "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
        : "r"(taddr)
        : "memory");
}
// 16 lanes x 256 bits, repeated 8 times along the columns: 32 registers per thread (the accumulator-fragment shape)
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// mode 0: ld, 1: ex2, 2: st, 3: four ld.x32 per wait, 4: ld.x64, 5: ld.16x256b.x8 (2 KiB per instruction)
__global__ void __launch_bounds__(512, 1) rate_kernel(int mode, int iters, long long* cycles, float* sink) {
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        tc::tmem_alloc(tc::smem_u32(&tmem_slot), 512);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot + ((uint32_t)((warp & 3) * 32) << 16);
    float acc = 0.f;
    uint32_t r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
    __syncthreads();
    const long long t0 = clock64();
    if (mode == 0) {
        for (int it = 0; it < iters; ++it) {
            uint32_t v[32];
            tc::tmem_ld_32x32b_x32(tmem + (uint32_t)((it * 32 + (warp >> 2) * 64) & 511), v);
            tc::tmem_ld_wait();
            acc += __uint_as_float(v[0]) + __uint_as_float(v[13]) + __uint_as_float(v[31]);
        }
    } else if (mode == 1) {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = -0.001f * (threadIdx.x + i);
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc += x[i];
    } else if (mode == 3) {
        for (int it = 0; it < iters; it += 4) {
            uint32_t v0[32], v1[32], v2[32], v3[32];
            const uint32_t c = (uint32_t)((warp >> 2) * 128) & 511;
            tc::tmem_ld_32x32b_x32(tmem + ((c + 0) & 511), v0);
            tc::tmem_ld_32x32b_x32(tmem + ((c + 32) & 511), v1);
            tc::tmem_ld_32x32b_x32(tmem + ((c + 64) & 511), v2);
            tc::tmem_ld_32x32b_x32(tmem + ((c + 96) & 511), v3);
            tc::tmem_ld_wait();
            acc += __uint_as_float(v0[0]) + __uint_as_float(v1[7]) + __uint_as_float(v2[21]) + __uint_as_float(v3[31]) + __uint_as_float(v0[31]) + __uint_as_float(v1[0]);
        }
    } else if (mode == 4) {
        for (int it = 0; it < iters; it += 2) {
            uint32_t v[64];
            tmem_ld_32x32b_x64(tmem + (uint32_t)((it * 32 + (warp >> 2) * 64) & 511 & ~63), v);
            tc::tmem_ld_wait();
            acc += __uint_as_float(v[0]) + __uint_as_float(v[33]) + __uint_as_float(v[63]);
        }
    } else if (mode == 5) {
        for (int it = 0; it < iters; it += 1) {
            uint32_t v[32];
            tmem_ld_16x256b_x8(tmem + (uint32_t)((it * 64) & 511 & ~63), v);  // 16 lanes x 64 columns = 4 KiB
            tc::tmem_ld_wait();
            acc += __uint_as_float(v[0]) + __uint_as_float(v[13]) + __uint_as_float(v[31]);
        }
    } else {
        for (int it = 0; it < iters; ++it) {
            tmem_st_32x32b_x32(tmem + (uint32_t)((it * 32 + (warp >> 2) * 64) & 511), r);
        }
        tmem_st_wait();
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem_slot, 512);
}

// tcgen05.mma issue / execution rate: `iters` back-to-back M = 128, K = 16 MMAs with N columns (SS operands in shared memory,
// contents irrelevant), one commit at the end; clocks per MMA
template <int N>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int iters, long long* cycles) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        tc::mbar_init(tc::smem_u32(&bar_done), 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) {
        tc::tmem_alloc(tc::smem_u32(&tmem_slot), 256);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = tc::idesc_bf16_f32(128, N);
        const uint64_t da = tc::smem_desc_sw128(base), db = tc::smem_desc_sw128(base + 16384);
        const long long t0 = clock64();
        for (int i = 0; i < iters; ++i) tc::mma_f16_ss(tmem, da + (uint64_t)(2 * (i & 3)), db + (uint64_t)(2 * (i & 3)), idesc, 1);
        const long long t1 = clock64();
        tc::mma_commit(tc::smem_u32(&bar_done));
        tc::mbar_wait(tc::smem_u32(&bar_done), 0);
        const long long t2 = clock64();
        cycles[0] = t1 - t0, cycles[1] = t2 - t0;
    }
    __syncthreads();
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

template <int N>
void mma_rate(long long* dc) {
    const int iters = 2048;
    cudaFuncSetAttribute(mma_rate_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    mma_rate_kernel<N><<<1, 128, 64 * 1024>>>(iters, dc);
    mma_rate_kernel<N><<<1, 128, 64 * 1024>>>(iters, dc);
    cudaDeviceSynchronize();
    long long c[2];
    cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
    printf("tcgen05.mma M 128 N %3d K 16 (SS): issue %6.1f clk / MMA, complete %6.1f clk / MMA  (%.0f MAC / clk)\n", N, (double)c[0] / iters,
           (double)c[1] / iters, 128.0 * N * 16 * iters / (double)c[1]);
}

// D[128 x 64] = A[128 x 64] B[64 x 64]^T: A in tensor memory (written by tcgen05.st), B K-major in swizzled shared memory
__global__ void __launch_bounds__(128, 1) ts_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, float* d) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* bp = smem_raw + (base - tc::smem_u32(smem_raw));
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tc::mbar_init(tc::smem_u32(&bar_done), 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) {
        tc::tmem_alloc(tc::smem_u32(&tmem_slot), 128);
        tc::tmem_relinquish();
    }
    // B: row n (128 bytes), 16-byte chunk c at position c ^ (n & 7), 8-row groups 1024 bytes apart
    for (int i = threadIdx.x; i < 64 * 8; i += 128) {
        const int n = i >> 3, c = i & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(b + n * 64 + c * 8);
        *reinterpret_cast<uint4*>(bp + n * 128 + ((c ^ (n & 7)) << 4)) = v;
    }
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    // A: thread = row; 64 bf16 = 32 packed columns [0, 32)
    {
        uint32_t r[32];
        const int row = threadIdx.x;
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = *reinterpret_cast<const uint32_t*>(a + row * 64 + 2 * j);
        tmem_st_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16), r);
        tmem_st_wait();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = tc::idesc_bf16_f32(128, 64);
        const uint64_t db = tc::smem_desc_sw128(base);
        for (int k = 0; k < 4; ++k) mma_f16_ts(tmem + 64, tmem + 8 * k, db + (uint64_t)(2 * k), idesc, k != 0);
        tc::mma_commit(tc::smem_u32(&bar_done));
    }
    tc::mbar_wait(tc::smem_u32(&bar_done), 0);
    tc::fence_after_sync();
    uint32_t v[32];
    for (int h = 0; h < 2; ++h) {
        tc::tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + 64 + 32 * h, v);
        tc::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) d[(warp * 32 + lane) * 64 + 32 * h + j] = __uint_as_float(v[j]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 128);
}

int main() {
    long long* dc;
    float* sink;
    cudaMalloc(&dc, 148 * 8), cudaMalloc(&sink, 4);
    const char* names[6] = {"tcgen05.ld x32", "ex2.approx", "tcgen05.st x32", "4 x ld x32 / wait", "ld 32x32b.x64", "ld 16x256b.x8"};
    for (int mode = 0; mode < 6; ++mode) {
        for (int threads : {128, 256, 512}) {
            for (int grid : {1}) {
                const int iters = 4096;
                rate_kernel<<<grid, threads>>>(mode, iters, dc, sink);
                rate_kernel<<<grid, threads>>>(mode, iters, dc, sink);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) {
                    printf("%s: %s\n", names[mode], cudaGetErrorString(e));
                    return 1;
                }
                std::vector<long long> c(grid);
                cudaMemcpy(c.data(), dc, grid * 8, cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (auto v : c) mx = v > mx ? v : mx;
                const double per = mode == 1 ? 8.0 * threads * iters : 128.0 * threads * iters;  // ops or bytes per CTA
                printf("%-16s threads %3d grid %3d: %9lld clk  %8.1f %s / clk / SM\n", names[mode], threads, grid, mx,
                       per / (double)mx, mode == 1 ? "ops" : "bytes");
            }
        }
    }
    mma_rate<32>(dc), mma_rate<64>(dc), mma_rate<128>(dc), mma_rate<256>(dc);
    // D
    std::vector<__nv_bfloat16> ha(128 * 64), hb(64 * 64);
    std::vector<float> fa(128 * 64), fb(64 * 64);
    srand(1);
    for (size_t i = 0; i < ha.size(); ++i) ha[i] = __float2bfloat16((rand() % 17 - 8) / 8.0f), fa[i] = __bfloat162float(ha[i]);
    for (size_t i = 0; i < hb.size(); ++i) hb[i] = __float2bfloat16((rand() % 13 - 6) / 4.0f), fb[i] = __bfloat162float(hb[i]);
    __nv_bfloat16 *da, *db;
    float* dd;
    cudaMalloc(&da, ha.size() * 2), cudaMalloc(&db, hb.size() * 2), cudaMalloc(&dd, 128 * 64 * 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024);
    ts_kernel<<<1, 128, 16 * 1024>>>(da, db, dd);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("ts mma: %s\n", cudaGetErrorString(e));
        return 1;
    }
    std::vector<float> hd(128 * 64);
    cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            double s = 0;
            for (int k = 0; k < 64; ++k) s += (double)fa[m * 64 + k] * fb[n * 64 + k];
            worst = fmax(worst, fabs(s - hd[m * 64 + n]));
        }
    printf("A-in-TMEM mma (128 x 64 x 64): max |d - ref| = %g %s\n", worst, worst < 1e-3 ? "OK" : "MISMATCH");
    return 0;
}
