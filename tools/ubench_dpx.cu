// Issue-rate / latency microbenchmark for the integer + DPX opcodes the SW kernels use.
// Measures lane-ops per clock per SM for each opcode on the attached GPU so that the
// "integer-ALU / DPX roofline" in DESIGN.md rests on a measured denominator, not on a nominal one.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o ubench_dpx tools/ubench_dpx.cu
//   ./ubench_dpx [json_out]
//
// Method: every warp runs ITER iterations of U independent dependency chains of one opcode; each warp
// records clock64() before/after; throughput = warps_per_SM * ITER * U * 32 / mean(cycles).  With one
// warp per SM and U=1 the same loop gives the dependent-issue latency.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

enum Op { OP_VIADDMAX16, OP_VIADDMAX16_RELU, OP_VIMAX3_16_RELU, OP_VMAX16, OP_VADD16, OP_IADD, OP_IMAD, OP_PRMT, OP_LOP3,
          OP_VIADDMAX32_RELU, OP_VIMAX3_32_RELU, OP_VIADDMIN32, OP_MIX_ALU_IMAD, OP_CELL6, OP_CELL6_PRMT, OP_CELL_IMADADD,
          OP_SHFL, OP_LDS32, OP_LDS128, OP_CELL6_LDS, OP_COUNT };
static const char* op_names[OP_COUNT] = { "VIADDMNMX.S16x2", "VIADDMNMX.S16x2.RELU", "VIMNMX3.S16x2.RELU", "VIMNMX.S16x2", "VIADD.16x2", "IADD3", "IMAD",
          "PRMT", "LOP3", "VIADDMNMX.RELU(s32)", "VIMNMX3.RELU(s32)", "VIADDMNMX(min,s32)", "mix 1 VIADDMNMX16 + 1 IMAD", "cell6 (2add+2viaddmax+vimax3relu+vmax)",
          "cell6 + PRMT score", "cell (adds as IMAD.IADD) 4dpx+2imad", "SHFL.UP", "LDS.32", "LDS.128", "cell6 + LDS.128/4" };
// lane-ops counted per "unit" of each op (for cell ops: instructions per packed cell)
static const int op_instr[OP_COUNT] = {1,1,1,1,1,1,1,1,1,1,1,1,2,6,7,6,1,1,1,6};

template <int OP, int U>
__global__ void __launch_bounds__(1024) bench(unsigned* out, long long* cyc, int iters, unsigned a, unsigned b, unsigned c)
{
    __shared__ __align__(16) unsigned sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i * a;
    __syncthreads();
    unsigned x[U], y[U], z[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { x[u] = threadIdx.x * 7 + u; y[u] = threadIdx.x + 3 * u; z[u] = u; }
    unsigned ga = a, gb = b, gc = c, fch = b, hd = a ^ b, cm = 0; uint4 sv = make_uint4(0, 0, 0, 0);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (OP == OP_VIADDMAX16) x[u] = __viaddmax_s16x2(x[u], ga, gb);
            else if (OP == OP_VIADDMAX16_RELU) x[u] = __viaddmax_s16x2_relu(x[u], ga, gb);
            else if (OP == OP_VIMAX3_16_RELU) x[u] = __vimax3_s16x2_relu(x[u], ga, y[u]);
            else if (OP == OP_VMAX16) x[u] = __vmaxs2(x[u], y[u]);
            else if (OP == OP_VADD16) x[u] = __vadd2(x[u], ga);
            else if (OP == OP_IADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[u]) : "r"(ga));
            else if (OP == OP_IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[u]) : "r"(ga), "r"(gb));
            else if (OP == OP_PRMT) x[u] = __byte_perm(x[u], ga, y[u]);
            else if (OP == OP_LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x6a;" : "+r"(x[u]) : "r"(ga), "r"(gb));
            else if (OP == OP_VIADDMAX32_RELU) x[u] = __viaddmax_s32_relu(x[u], ga, gb);
            else if (OP == OP_VIMAX3_32_RELU) x[u] = __vimax3_s32_relu(x[u], ga, y[u]);
            else if (OP == OP_VIADDMIN32) x[u] = __viaddmin_s32(x[u], ga, gb);
            else if (OP == OP_MIX_ALU_IMAD) { x[u] = __viaddmax_s16x2(x[u], ga, gb); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[u]) : "r"(ga), "r"(gb)); }
            else if (OP == OP_CELL6 || OP == OP_CELL6_PRMT || OP == OP_CELL_IMADADD || OP == OP_CELL6_LDS) {
                // one packed DP cell: x = H of this row (previous column), y = E, fch = F chain down the rows, hd = diagonal carrier, cm = column max
                unsigned s;
                if (OP == OP_CELL6_PRMT) s = __byte_perm(ga, gb, z[u]);
                else if (OP == OP_CELL6_LDS) { if ((u & 3) == 0) sv = reinterpret_cast<uint4*>(sm)[(threadIdx.x & 31) + 32 * ((u / 4 + it) & 31)]; s = (u & 3) == 0 ? sv.x : (u & 3) == 1 ? sv.y : (u & 3) == 2 ? sv.z : sv.w; }
                else s = z[u];
                unsigned h;
                if (OP == OP_CELL_IMADADD) asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(h) : "r"(hd), "r"(s)); else h = __vadd2(hd, s);
                hd = x[u];
                unsigned H = __vimax3_s16x2_relu(h, y[u], fch);
                unsigned Hg;
                if (OP == OP_CELL_IMADADD) asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(Hg) : "r"(H), "r"(ga)); else Hg = __vadd2(H, ga);
                y[u] = __viaddmax_s16x2(y[u], gc, Hg);
                fch = __viaddmax_s16x2(fch, gc, Hg);
                x[u] = H;
                cm = __vmaxs2(cm, H);
            }
            else if (OP == OP_SHFL) x[u] = __shfl_up_sync(0xffffffffu, x[u], 1) + 1;
            else if (OP == OP_LDS32) x[u] = sm[(x[u] + threadIdx.x) & 4095];
            else if (OP == OP_LDS128) { uint4 v = reinterpret_cast<uint4*>(sm)[(x[u] + threadIdx.x) & 1023]; x[u] = v.x ^ v.y ^ v.z ^ v.w; }
        }
    }
    long long t1 = clock64();
    unsigned acc = gb ^ fch ^ hd ^ cm ^ sv.x;
#pragma unroll
    for (int u = 0; u < U; ++u) acc ^= x[u] ^ y[u] ^ z[u];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if ((threadIdx.x & 31) == 0) cyc[(blockIdx.x * blockDim.x + threadIdx.x) >> 5] = t1 - t0;
}

struct Res { double lanes_per_clk_sm; double cyc_per_iter; };

template <int OP, int U>
Res run(int nsm, int warps_per_sm, int iters)
{
    int threads = warps_per_sm * 32;
    unsigned* out; long long* cyc;
    CK(cudaMalloc(&out, sizeof(unsigned) * nsm * threads));
    CK(cudaMalloc(&cyc, sizeof(long long) * nsm * warps_per_sm));
    bench<OP, U><<<nsm, threads>>>(out, cyc, 16, 0xfffefffeu, 3u, 0xfffefffeu);
    CK(cudaDeviceSynchronize());
    bench<OP, U><<<nsm, threads>>>(out, cyc, iters, 0xfffefffeu, 3u, 0xfffefffeu);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(nsm * warps_per_sm);
    CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost));
    double mean = 0; for (auto v : h) mean += (double)v; mean /= h.size();
    CK(cudaFree(out)); CK(cudaFree(cyc));
    Res r;
    r.cyc_per_iter = mean / iters;
    r.lanes_per_clk_sm = (double)warps_per_sm * iters * U * 32.0 * op_instr[OP] / mean;
    return r;
}

template <int OP>
void one(FILE* js, int nsm, bool& first)
{
    const int iters = 4096;
    Res lat = run<OP, 1>(nsm, 1, iters);          // 1 warp/SM, 1 chain: dependent-issue latency per iteration
    Res t8 = run<OP, 8>(nsm, 8, iters);           // 8 warps (2/SMSP) x 8 chains
    Res t16 = run<OP, 8>(nsm, 16, iters);         // 16 warps (4/SMSP) x 8 chains
    Res t32 = run<OP, 8>(nsm, 32, iters);         // 32 warps (8/SMSP) x 8 chains
    double best = std::max(t8.lanes_per_clk_sm, std::max(t16.lanes_per_clk_sm, t32.lanes_per_clk_sm));
    printf("%-46s lat/iter %7.2f cyc | lane-instr/clk/SM: w8 %7.1f  w16 %7.1f  w32 %7.1f  best %7.1f\n", op_names[OP], lat.cyc_per_iter,
           t8.lanes_per_clk_sm, t16.lanes_per_clk_sm, t32.lanes_per_clk_sm, best);
    if (js) fprintf(js, "%s\n  {\"op\": \"%s\", \"instr_per_unit\": %d, \"latency_cyc_per_iter\": %.3f, \"lane_instr_per_clk_sm\": {\"w8\": %.2f, \"w16\": %.2f, \"w32\": %.2f}, \"best\": %.2f}",
                    first ? "" : ",", op_names[OP], op_instr[OP], lat.cyc_per_iter, t8.lanes_per_clk_sm, t16.lanes_per_clk_sm, t32.lanes_per_clk_sm, best);
    first = false;
}

int main(int argc, char** argv)
{
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    printf("device %s, %d SMs, cc %d.%d, clock %.0f MHz\n", p.name, nsm, p.major, p.minor, p.clockRate / 1000.0);
    FILE* js = argc > 1 ? fopen(argv[1], "w") : nullptr;
    if (js) fprintf(js, "{\"device\": \"%s\", \"sms\": %d, \"ops\": [", p.name, nsm);
    bool first = true;
    one<OP_VIADDMAX16>(js, nsm, first);
    one<OP_VIADDMAX16_RELU>(js, nsm, first);
    one<OP_VIMAX3_16_RELU>(js, nsm, first);
    one<OP_VMAX16>(js, nsm, first);
    one<OP_VADD16>(js, nsm, first);
    one<OP_IADD>(js, nsm, first);
    one<OP_IMAD>(js, nsm, first);
    one<OP_PRMT>(js, nsm, first);
    one<OP_LOP3>(js, nsm, first);
    one<OP_VIADDMAX32_RELU>(js, nsm, first);
    one<OP_VIMAX3_32_RELU>(js, nsm, first);
    one<OP_VIADDMIN32>(js, nsm, first);
    one<OP_MIX_ALU_IMAD>(js, nsm, first);
    one<OP_CELL6>(js, nsm, first);
    one<OP_CELL6_PRMT>(js, nsm, first);
    one<OP_CELL_IMADADD>(js, nsm, first);
    one<OP_CELL6_LDS>(js, nsm, first);
    one<OP_SHFL>(js, nsm, first);
    one<OP_LDS32>(js, nsm, first);
    one<OP_LDS128>(js, nsm, first);
    if (js) { fprintf(js, "\n]}\n"); fclose(js); }
    return 0;
}
