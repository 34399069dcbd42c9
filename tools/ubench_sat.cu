// Does VIADDMNMX.S16x2 wrap its intermediate sum at 16 bits?  (decides whether the int16 clamp of ssw.c:425 is free in packed form)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned* out, unsigned a, unsigned b, unsigned c)
{
    out[0] = __viaddmin_s16x2(a, b, c);            // min(a + b, c) per half, signed
    out[1] = __viaddmin_u16x2(a, b, c);            // unsigned
    out[2] = __viaddmax_s16x2(a, b, c);
    out[3] = __viaddmin_s16x2_relu(a, b, c);
    out[4] = __vaddss2(a, b);                      // saturating signed add
    int r = __viaddmin_s32((int)a, (int)b, (int)c);
    out[5] = (unsigned)r;
}
int main()
{
    unsigned* d; cudaMalloc(&d, 64);
    unsigned h[6];
    // halves: low = 32767 + 4, high = 32760 + 10 ; c = 32767 in both
    k<<<1, 1>>>(d, 0x7ff87fffu, 0x000a0004u, 0x7fff7fffu);
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("a=7ff8|7fff b=000a|0004 c=7fff|7fff: viaddmin_s16x2 %08x  viaddmin_u16x2 %08x  viaddmax_s16x2 %08x  viaddmin_s16x2_relu %08x  vaddss2 %08x  s32 %08x\n", h[0], h[1], h[2], h[3], h[4], h[5]);
    // negative addend near the top and near zero: low = 32767 - 6, high = 3 - 6
    k<<<1, 1>>>(d, 0x00037fffu, 0xfffafffau, 0x7fff7fffu);
    cudaMemcpy(h, d, 24, cudaMemcpyDeviceToHost);
    printf("a=0003|7fff b=fffa|fffa c=7fff|7fff: viaddmin_s16x2 %08x  viaddmin_u16x2 %08x  viaddmax_s16x2 %08x  viaddmin_s16x2_relu %08x  vaddss2 %08x\n", h[0], h[1], h[2], h[3], h[4]);
    return 0;
}
