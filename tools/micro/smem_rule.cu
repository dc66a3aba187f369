// Micro-experiment: how does sm_100 split 64-/128-bit shared-memory accesses into wavefronts?
// Every pattern touches 32 distinct 8-/16-byte units (so 2 / 4 wavefronts is the floor); the patterns differ in whether
// fixed half-warps / quarter-warps see distinct banks.  Prints cycles per access (dependent chain => latency + wavefronts).
#include <cstdio>
#include <cuda_runtime.h>
template <int W>  // W = words per access (2 or 4)
__global__ void k(const int* perm, long long* out, int iters) {
    __shared__ __align__(16) float buf[32 * 4 * 4];
    for (int i = threadIdx.x; i < 32 * 4 * 4; i += 32) buf[i] = 0.f;
    __syncwarp();
    int u = perm[threadIdx.x];
    float acc = 0.f;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        if (W == 4) { float4 v = *reinterpret_cast<float4*>(buf + 4 * u); acc += v.x + v.y + v.z + v.w; u = (u + (int)v.x) & 31; }
        else        { float2 v = *reinterpret_cast<float2*>(buf + 2 * u); acc += v.x + v.y; u = (u + (int)v.x) & 31; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    if (acc == 123.f) out[1] = 1;
}
// throughput version: 8 independent loads in flight
template <int W>
__global__ void kt(const int* perm, long long* out, int iters) {
    __shared__ __align__(16) float buf[32 * 4 * 4];
    for (int i = threadIdx.x; i < 32 * 4 * 4; i += 32) buf[i] = 0.f;
    __syncwarp();
    const int u = perm[threadIdx.x];
    float acc[8] = {0};
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            unsigned a = (unsigned)__cvta_generic_to_shared(buf + W * u);
            if (W == 4) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); acc[j] += v.x + v.w; }
            else        { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); acc[j] += v.x + v.y; }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    float s = 0; for (int j = 0; j < 8; ++j) s += acc[j];
    if (s == 123.f) out[1] = 1;
}
int main() {
    int *d; long long *o; cudaMalloc(&d, 128); cudaMalloc(&o, 16);
    const int iters = 20000;
    struct P { const char* name; int perm[32]; } pats[5];
    pats[0].name = "identity (unit = lane)";
    for (int l = 0; l < 32; ++l) pats[0].perm[l] = l;
    pats[1].name = "128b: quarter-warps see 2 bank groups only (u = 8*(l%4) + l/4)";
    for (int l = 0; l < 32; ++l) pats[1].perm[l] = 8 * (l % 4) + l / 4;
    pats[2].name = "64b: half-warps see 8 bank pairs only (u = 16*(l%2) + l/2)";
    for (int l = 0; l < 32; ++l) pats[2].perm[l] = 16 * (l % 2) + l / 2;
    pats[3].name = "rotation within the warp (u = (l + 5) % 32)";
    for (int l = 0; l < 32; ++l) pats[3].perm[l] = (l + 5) % 32;
    pats[4].name = "lane 8 and lane 13 swap bank group with other quarter (u: 8<->21)";
    for (int l = 0; l < 32; ++l) pats[4].perm[l] = l; pats[4].perm[8] = 21; pats[4].perm[21] = 8;
    for (int w = 2; w <= 4; w += 2)
        for (auto& p : pats) {
            cudaMemcpy(d, p.perm, 128, cudaMemcpyHostToDevice);
            long long h[2];
            if (w == 2) k<2><<<1, 32>>>(d, o, iters); else k<4><<<1, 32>>>(d, o, iters);
            cudaMemcpy(h, o, 16, cudaMemcpyDeviceToHost);
            double lat = (double)h[0] / iters;
            if (w == 2) kt<2><<<1, 32>>>(d, o, iters); else kt<4><<<1, 32>>>(d, o, iters);
            cudaMemcpy(h, o, 16, cudaMemcpyDeviceToHost);
            printf("LDS.%d  %-70s chain %.1f cyc/load   throughput %.2f cyc/load\n", w * 32, p.name, lat, (double)h[0] / iters / 8);
        }
    return 0;
}
