// Probe: can tcgen05.mma read the nine taps of a 3x3 convolution as SHIFTED VIEWS of one TMA-loaded halo tile?
//
//   halo tile  = (16 + 2) rows x PITCH columns of pixels x 64 channels (128-byte swizzled rows), one TMA box
//   A operand  = 128 pixels = 8 wide x 16 tall patch: core-matrix group g = patch row g (8 pixels, 128 B apart),
//                stride-byte-offset = PITCH * 128, start address = tile + ((kh * PITCH) + kw) * 128 (+ 32 per K step)
//
// Variants: PITCH 10 (dense halo, SBO = 1280, groups change swizzle phase) and PITCH 16 (SBO = 2048, phase constant),
// each with descriptor base_offset = 0 and base_offset = (start >> 7) & 7.  Prints the max error of each against a host
// reference.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I azula_b200/csrc scripts/halo_probe.cu
//                    -o build/halo_probe -lcuda   (run on the GPU box: build/halo_probe)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_bf16.h>
#include "tc.cuh"

constexpr int PH = 16, PW = 8, CH = 64, NOUT = 64;

struct Params {
    int pitch;        // halo row pitch in pixels
    int base_mode;    // 0: base_offset 0, 1: (start >> 7) & 7
    float* out;       // [128][NOUT]
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr, uint32_t sbo_bytes, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(sbo_bytes >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7u) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

__global__ void __launch_bounds__(128, 1) probe_kernel(const __grid_constant__ CUtensorMap tmap_a,
                                                       const __grid_constant__ CUtensorMap tmap_b, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t a_tile = base;                 // halo tile, up to 18 * 16 * 128 = 36 KiB
    const uint32_t b_tile = base + 40 * 1024;     // 9 taps x (64 rows x 128 B) = 72 KiB
    __shared__ __align__(8) uint64_t bar_full, bar_done;
    __shared__ uint32_t tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tc::mbar_init(tc::smem_u32(&bar_full), 1);
        tc::mbar_init(tc::smem_u32(&bar_done), 1);
        tc::fence_barrier_init();
    }
    if (warp == 0) {
        tc::tmem_alloc(tc::smem_u32(&tmem_slot), 64);
        tc::tmem_relinquish();
    }
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t a_bytes = (uint32_t)(PH + 2) * p.pitch * 128;
        tc::mbar_expect_tx(tc::smem_u32(&bar_full), a_bytes + 9 * NOUT * 128);
        tc::tma_load_4d(a_tile, &tmap_a, tc::smem_u32(&bar_full), 0, -1, -1, 0);  // halo starts at pixel (-1, -1): zero fill
        for (int t = 0; t < 9; ++t) tc::tma_load_2d(b_tile + t * NOUT * 128, &tmap_b, tc::smem_u32(&bar_full), t * CH, 0);
        tc::mbar_wait(tc::smem_u32(&bar_full), 0);
        tc::fence_after_sync();
        constexpr uint32_t idesc = tc::idesc_bf16_f32(128, NOUT);
        const uint32_t sbo = (uint32_t)p.pitch * 128;
        for (int t = 0; t < 9; ++t) {
            const int kh = t / 3, kw = t % 3;
            const uint32_t a0 = a_tile + (uint32_t)(kh * p.pitch + kw) * 128;
            for (int k = 0; k < 4; ++k) {
                const uint32_t a_addr = a0 + k * 32;
                const uint64_t da = desc_sw128(a_addr, sbo, p.base_mode ? (a_addr >> 7) & 7u : 0u);
                const uint64_t db = desc_sw128(b_tile + t * NOUT * 128 + k * 32, 1024, 0);
                tc::mma_f16_ss(tmem, da, db, idesc, (t | k) != 0);
            }
        }
        tc::mma_commit(tc::smem_u32(&bar_done));
    }
    __syncwarp();
    tc::mbar_wait(tc::smem_u32(&bar_done), 0);
    tc::fence_after_sync();
    for (int c0 = 0; c0 < NOUT; c0 += 32) {
        uint32_t acc[32];
        tc::tmem_ld_32x32b_x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, acc);
        tc::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) p.out[(warp * 32 + lane) * NOUT + c0 + j] = __uint_as_float(acc[j]);
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

static float bf(float v) { return __bfloat162float(__float2bfloat16(v)); }

int main() {
    // image: PH x PW pixels x CH channels (no halo in memory: the TMA unit zero-fills rows / columns -1 and PH / PW)
    std::vector<float> x(PH * PW * CH), w(NOUT * 9 * CH);
    srand(1);
    for (auto& v : x) v = bf((rand() % 2001 - 1000) / 500.0f);
    for (auto& v : w) v = bf((rand() % 2001 - 1000) / 4000.0f);
    std::vector<__nv_bfloat16> xb(x.size()), wb(w.size());
    for (size_t i = 0; i < x.size(); ++i) xb[i] = __float2bfloat16(x[i]);
    for (size_t i = 0; i < w.size(); ++i) wb[i] = __float2bfloat16(w[i]);  // [NOUT][tap][CH]
    std::vector<float> ref(128 * NOUT, 0.f);
    for (int h = 0; h < PH; ++h)
        for (int ww = 0; ww < PW; ++ww)
            for (int co = 0; co < NOUT; ++co) {
                double s = 0;
                for (int kh = 0; kh < 3; ++kh)
                    for (int kw = 0; kw < 3; ++kw) {
                        const int ih = h + kh - 1, iw = ww + kw - 1;
                        if (ih < 0 || ih >= PH || iw < 0 || iw >= PW) continue;
                        for (int c = 0; c < CH; ++c) s += (double)x[(ih * PW + iw) * CH + c] * w[(co * 9 + kh * 3 + kw) * CH + c];
                    }
                ref[(h * PW + ww) * NOUT + co] = (float)s;
            }
    __nv_bfloat16 *dx, *dw;
    float* dout;
    cudaMalloc(&dx, xb.size() * 2), cudaMalloc(&dw, wb.size() * 2), cudaMalloc(&dout, 128 * NOUT * 4);
    cudaMemcpy(dx, xb.data(), xb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dw, wb.data(), wb.size() * 2, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024);
    for (int pitch : {10, 16})
        for (int mode : {0, 1}) {
            CUtensorMap ta, tb;
            uint64_t dims[4] = {CH, PW, PH, 1}, str[3] = {CH * 2, CH * 2 * PW, CH * 2 * PW * PH};
            uint32_t box[4] = {CH, (uint32_t)pitch, PH + 2, 1};
            if (tc::make_map_bf16(&ta, dx, 4, dims, str, box)) { printf("tmap a failed\n"); return 1; }
            uint64_t dimb[2] = {9 * CH, NOUT}, strb[1] = {9 * CH * 2};
            uint32_t boxb[2] = {CH, NOUT};
            if (tc::make_map_bf16(&tb, dw, 2, dimb, strb, boxb)) { printf("tmap b failed\n"); return 1; }
            cudaMemset(dout, 0, 128 * NOUT * 4);
            Params p{pitch, mode, dout};
            probe_kernel<<<1, 128, 120 * 1024>>>(ta, tb, p);
            cudaError_t e = cudaDeviceSynchronize();
            std::vector<float> out(128 * NOUT);
            cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
            double err = 0, mag = 0;
            for (size_t i = 0; i < out.size(); ++i) err = fmax(err, fabs(out[i] - ref[i])), mag = fmax(mag, fabs(ref[i]));
            printf("pitch %2d base_mode %d: %s max|err| = %.3e (max|ref| = %.3f) %s\n", pitch, mode, cudaGetErrorString(e), err, mag,
                   err < 1e-3 * mag ? "OK" : "MISMATCH");
        }
    return 0;
}
