// tools/chainbench.cu -- what can ONE warp do for ONE chain?  Cycles per 64-byte block of a lone warp running
//   (a) SHA-256 rounds only, W+K rows read from shared memory (the schedule expanded elsewhere),
//   (b) full SHA-256 (schedule + rounds), (c) MD5, (d) fused SHA-256 + MD5 (the lane kernel's compress<>),
// message words read from shared memory.  One CTA, NW warps (one per SMSP when NW <= 4).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/chainbench tools/chainbench.cu
#include <cstdio>
#include "../modal_client_b200/csrc/b200hash_kernels.cu"

namespace b200h {

template <int MODE>
__global__ void __launch_bounds__(128) chain_bench(uint32_t* out, long long* cyc, int nblocks, uint32_t one, uint32_t seed) {
    __shared__ __align__(16) uint32_t rows[32][68];  // 32 blocks of 64 W+K words (or 16 message words), padded
    for (int i = threadIdx.x; i < 32 * 68; i += blockDim.x) (&rows[0][0])[i] = seed * 2654435761u + i * 40503u;
    __syncthreads();
    uint32_t hs[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
    uint32_t hm[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
    const long long t0 = clock64();
#pragma unroll 1
    for (int b = 0; b < nblocks; ++b) {
        const uint4* row = reinterpret_cast<const uint4*>(rows[b & 31]);
        if (MODE == 0) {
            uint32_t k[64];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const uint4 v = row[i];
                k[4 * i] = v.x; k[4 * i + 1] = v.y; k[4 * i + 2] = v.z; k[4 * i + 3] = v.w;
            }
            uint32_t a = hs[0], bb = hs[1], c = hs[2], d = hs[3], e = hs[4], f = hs[5], g = hs[6], h = hs[7];
#pragma unroll
            for (int i = 0; i < 64; i += 8) {
                CH_RND(a, bb, c, d, e, f, g, h, k[i]);
                CH_RND(h, a, bb, c, d, e, f, g, k[i + 1]);
                CH_RND(g, h, a, bb, c, d, e, f, k[i + 2]);
                CH_RND(f, g, h, a, bb, c, d, e, k[i + 3]);
                CH_RND(e, f, g, h, a, bb, c, d, k[i + 4]);
                CH_RND(d, e, f, g, h, a, bb, c, k[i + 5]);
                CH_RND(c, d, e, f, g, h, a, bb, k[i + 6]);
                CH_RND(bb, c, d, e, f, g, h, a, k[i + 7]);
            }
            hs[0] += a; hs[1] += bb; hs[2] += c; hs[3] += d; hs[4] += e; hs[5] += f; hs[6] += g; hs[7] += h;
        } else {
            uint32_t x[16];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint4 v = row[i];
                x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
            }
            if (MODE == 1) compress<true, false>(hs, hm, x, false, 0u, 0u, one);
            if (MODE == 2) compress<false, true>(hs, hm, x, false, 0u, 0u, one);
            if (MODE == 3) compress<true, true>(hs, hm, x, false, 0u, 0u, one);
        }
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= hs[i];
#pragma unroll
    for (int i = 0; i < 4; ++i) s ^= hm[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}

}  // namespace b200h

template <int MODE>
static void run(const char* name, int warps) {
    uint32_t* out;
    long long* cyc;
    cudaMalloc(&out, 4096);
    cudaMalloc(&cyc, 8);
    const int nb = 4096;
    for (int rep = 0; rep < 2; ++rep) b200h::chain_bench<MODE><<<1, 32 * warps>>>(out, cyc, nb, 1u, 7u);
    cudaDeviceSynchronize();
    long long c = 0;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    cudaError_t e = cudaGetLastError();
    const double cpb = (double)c / nb;
    printf("%-34s warps=%d  %8.1f cycles/block  -> %6.1f MB/s per chain at 1.965 GHz  (%s)\n", name, warps, cpb,
           64.0 * 1965.0 / cpb, cudaGetErrorString(e));
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    for (int w : {1, 2, 4}) {
        run<0>("SHA-256 rounds only (W+K from smem)", w);
        run<1>("SHA-256 full (schedule + rounds)", w);
        run<2>("MD5", w);
        run<3>("fused SHA-256 + MD5", w);
    }
    return 0;
}
