// Second-stage microbenchmark: candidate inner-loop formulations of one packed (s16x2) DP cell.
// Each variant keeps KR packed rows in registers and iterates "columns"; reports packed-cells/clk/SM
// (x2 = cells/clk/SM).  Run under ncu with pipe metrics to see which pipe each opcode lands on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ubench_cell tools/ubench_cell.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

enum V { V_STD = 0, V_STD_PRMT, V_SHIFT, V_SHIFT_PRMT, V_SHIFT_LDS, V_VIADD_ONLY, V_DPX_VIADD_1_1, V_DPX_VIADD_2_1, V_DPX_PRMT_1_1, V_DPX_IMAD_2_1, V_COUNT };
static const char* names[V_COUNT] = {"std 5.5dpx (score in reg)", "std + PRMT score", "shifted 4.5dpx + VIADD", "shifted + PRMT score", "shifted + VIADD addr + LDS.32 score",
   "VIADD only (32-bit add)", "mix VIADDMNMX:VIADD 1:1", "mix VIADDMNMX:VIADD 2:1", "mix VIADDMNMX:PRMT 1:1", "mix VIADDMNMX:IMAD 2:1"};

__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned c) { unsigned d; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

template <int VAR, int KR>
__global__ void __launch_bounds__(512) cellbench(unsigned* out, long long* cyc, int iters, unsigned a, unsigned b, unsigned c, const unsigned* __restrict__ gsel)
{
    extern __shared__ unsigned tab[];   // lane-replicated table: entry e at tab[e*32 + lane]
    for (int i = threadIdx.x; i < 640 * 32; i += blockDim.x) tab[i] = (i * 2654435761u) & 0x00070007u;
    __syncthreads();
    unsigned H[KR], E[KR], sel[KR];
#pragma unroll
    for (int j = 0; j < KR; ++j) { H[j] = 0; E[j] = 0; sel[j] = gsel[((threadIdx.x * KR + j) & 1023) + (VAR == V_SHIFT_LDS ? 1024 : 0)]; }
    const unsigned mgapE = 0xfffefffeu, mgapO = 0xfff8fff8u;
    unsigned srcA = a + threadIdx.x, srcB = b ^ (threadIdx.x * 3), cm = 0, Ftop = 0, Hdtop = 0, tb = 0;
    const unsigned lane4 = (threadIdx.x & 31) * 4;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        // per-column boundary values change every iteration (stand-ins for the shuffled-in values)
        srcA = srcA * 1664525u + 1013904223u; srcB = srcB ^ (srcA >> 3);
        tb = (srcA >> 28) * 25u * 128u;
        unsigned F = Ftop, hd = Hdtop;
        if (VAR == V_VIADD_ONLY || VAR == V_DPX_VIADD_1_1 || VAR == V_DPX_VIADD_2_1 || VAR == V_DPX_PRMT_1_1 || VAR == V_DPX_IMAD_2_1) {
#pragma unroll
            for (int j = 0; j < KR; ++j) {
                if (VAR == V_VIADD_ONLY) { asm volatile("add.s32 %0, %0, %1;" : "+r"(H[j]) : "r"(srcA)); }
                if (VAR == V_DPX_VIADD_1_1) { H[j] = __viaddmax_s16x2(H[j], mgapE, srcA); asm volatile("add.s32 %0, %0, %1;" : "+r"(E[j]) : "r"(srcB)); }
                if (VAR == V_DPX_VIADD_2_1) { H[j] = __viaddmax_s16x2(H[j], mgapE, srcA); sel[j] = __viaddmax_s16x2(sel[j], mgapE, srcB); asm volatile("add.s32 %0, %0, %1;" : "+r"(E[j]) : "r"(srcB)); }
                if (VAR == V_DPX_PRMT_1_1) { H[j] = __viaddmax_s16x2(H[j], mgapE, srcA); E[j] = prmt(srcA, srcB, E[j]); }
                if (VAR == V_DPX_IMAD_2_1) { H[j] = __viaddmax_s16x2(H[j], mgapE, srcA); sel[j] = __viaddmax_s16x2(sel[j], mgapE, srcB); asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(E[j]) : "r"(srcB), "r"(srcA)); }
            }
        } else {
#pragma unroll
            for (int j = 0; j < KR; ++j) {
                unsigned s;
                if (VAR == V_STD_PRMT || VAR == V_SHIFT_PRMT) s = prmt(srcA, srcB, sel[j]);
                else if (VAR == V_SHIFT_LDS) { unsigned ad; asm volatile("add.s32 %0, %1, %2;" : "=r"(ad) : "r"(sel[j]), "r"(tb)); s = *reinterpret_cast<const unsigned*>(reinterpret_cast<const char*>(tab) + (min(ad, 79000u) + lane4)); }
                else s = sel[j];
                unsigned Hn;
                if (VAR == V_STD || VAR == V_STD_PRMT) {
                    unsigned h = __vadd2(hd, s);
                    Hn = __vimax3_s16x2_relu(h, E[j], F);
                    unsigned Hg = __vadd2(Hn, mgapO);
                    E[j] = __viaddmax_s16x2(E[j], mgapE, Hg);
                    F = __viaddmax_s16x2(F, mgapE, Hg);
                } else {
                    // shifted form: E,F stored +gapO; score word pre-biased by +gapO, 32-bit add is carry-safe (both halves >= 0)
                    unsigned t; asm volatile("add.s32 %0, %1, %2;" : "=r"(t) : "r"(hd), "r"(s));
                    unsigned U = __vimax3_s16x2(t, E[j], F);
                    Hn = __viaddmax_s16x2_relu(U, mgapO, mgapO);
                    E[j] = __viaddmax_s16x2(E[j], mgapE, Hn);
                    F = __viaddmax_s16x2(F, mgapE, Hn);
                }
                hd = H[j]; H[j] = Hn;
                if (j & 1) cm = __vimax3_s16x2(cm, H[j - 1], Hn);
            }
        }
        Ftop = F ^ (srcB & 0x00010001u); Hdtop = hd;
    }
    long long t1 = clock64();
    unsigned acc = cm ^ Ftop ^ Hdtop ^ srcA;
#pragma unroll
    for (int j = 0; j < KR; ++j) acc ^= H[j] ^ E[j] ^ sel[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) cyc[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0;
}

template <int VAR, int KR>
void run(int nsm, unsigned* gsel, int only)
{
    if (only >= 0 && only != VAR) return;
    auto kern = cellbench<VAR, KR>;
    size_t smem = 640 * 32 * 4;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int iters = 2048;
    printf("%-40s KR=%2d:", names[VAR], KR);
    for (int warps : {4, 8, 12, 16}) {
        int threads = warps * 32;
        unsigned* out; long long* cyc;
        CK(cudaMalloc(&out, sizeof(unsigned) * nsm * threads));
        CK(cudaMalloc(&cyc, sizeof(long long) * nsm * warps));
        kern<<<nsm, threads, smem>>>(out, cyc, 8, 0x04fafafau, 0xfa04fafau, 3u, gsel);
        CK(cudaDeviceSynchronize());
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        kern<<<nsm, threads, smem>>>(out, cyc, iters, 0x04fafafau, 0xfa04fafau, 3u, gsel);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        std::vector<long long> h(nsm * warps);
        CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
        double mx = 0; for (auto v : h) mx = v > mx ? v : mx;
        double units = (double)warps * iters * KR * 32.0;
        printf("  w%-2d %6.2f/clk/SM (%.3f ms)", warps, units / mx, ms);
        CK(cudaFree(out)); CK(cudaFree(cyc));
    }
    printf("   [units = packed cells or op-groups per clk per SM; uses MAX warp cycles]\n");
}

int main(int argc, char** argv)
{
    int only = argc > 1 ? atoi(argv[1]) : -1;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    std::vector<unsigned> hs(2048);
    for (int i = 0; i < 1024; ++i) { unsigned q0 = (i * 7) & 3, q1 = (i * 13 + 1) & 3; hs[i + 1024] = (q0 * 5 + q1) * 128; hs[i] = (q0 | ((q0 | 8) << 4) | ((4 + q1) << 8) | ((4 + q1 | 8) << 12)); }
    unsigned* gsel; CK(cudaMalloc(&gsel, 8192)); CK(cudaMemcpy(gsel, hs.data(), 8192, cudaMemcpyHostToDevice));
    printf("device %s, %d SMs\n", p.name, nsm);
    run<V_STD, 16>(nsm, gsel, only);
    run<V_STD_PRMT, 16>(nsm, gsel, only);
    run<V_SHIFT, 16>(nsm, gsel, only);
    run<V_SHIFT_PRMT, 16>(nsm, gsel, only);
    run<V_SHIFT_LDS, 16>(nsm, gsel, only);
    run<V_VIADD_ONLY, 16>(nsm, gsel, only);
    run<V_DPX_VIADD_1_1, 16>(nsm, gsel, only);
    run<V_DPX_VIADD_2_1, 16>(nsm, gsel, only);
    run<V_DPX_PRMT_1_1, 16>(nsm, gsel, only);
    run<V_DPX_IMAD_2_1, 16>(nsm, gsel, only);
    return 0;
}
