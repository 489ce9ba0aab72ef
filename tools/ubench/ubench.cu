// Instruction-throughput micro-benchmarks for sm_100a (B200): which integer ops the fused
// per-pixel kernels should lean on.  One block per SM, NW warps, ILP independent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    if (OP == 0) { asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
    else if (OP == 1) { asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
    else if (OP == 2) { asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
    else if (OP == 3) { asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c & 0x7777)); }
    else if (OP == 4) { asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
    else if (OP == 5) { asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
    else if (OP == 6) { asm volatile("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
    else if (OP == 7) { r = max(a, max(b, c)); }
    else if (OP == 8) { asm volatile("vadd4.u32.u32.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
    else if (OP == 9) { asm volatile("bfe.u32 %0, %1, 8, 8;" : "=r"(r) : "r"(a)); r += b; }
    else if (OP == 10) { float f = __uint_as_float(a); f = __fmaf_rn(f, 1.0001f, __uint_as_float(b)); r = __float_as_uint(f); }
    else if (OP == 11) { r = __popc(a) + b; }
    else if (OP == 12) { asm volatile("mul.wide.u16 %0, %1, %2;" : "=r"(r) : "h"((uint16_t)a), "h"((uint16_t)b)); r += c; }
    else if (OP == 13) { asm volatile("mad24.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); }
    else if (OP == 14) { r = __vadd2(a, b); }
    else if (OP == 15) { r = a * b + c; r = (r >> 12); }
    return r;
}

template <int OP>
__global__ void k_alu(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed + threadIdx.x * 7 + i;
    uint32_t b = seed | 3, c = seed ^ 0x1234;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) v[i] = op<OP>(v[i], b, c);
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// mixed: imad + lop3 alternating (different pipes?)
__global__ void k_mix(uint32_t* out, long long* cyc, uint32_t seed) {
    uint32_t v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = seed + threadIdx.x * 7 + i;
    uint32_t b = seed | 3, c = seed ^ 0x1234;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; i += 2) {
            v[i] = op<0>(v[i], b, c);
            v[i + 1] = op<4>(v[i + 1], b, c);
        }
    }
    long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s ^= v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// shared-memory ops.  MODE 0: LDS.32 conflict-free; 1: ATOMS.ADD distinct addresses per lane (bin = lane + 32*k);
// 2: ATOMS.ADD all lanes same address; 3: 4 distinct addresses per warp (8 lanes each); 4: random bins (256) ;
// 5: LDS.U8 stride 3; 6: ATOMS random bins with 8 lane-class replicas; 7: LDS.128
template <int MODE>
__global__ void k_smem(uint32_t* out, long long* cyc, uint32_t seed) {
    __shared__ uint32_t sm[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = i * seed;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    uint32_t acc = 0, rnd = seed + threadIdx.x * 2654435761u;
    long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            rnd = rnd * 1664525u + 1013904223u;
            if (MODE == 0) acc += sm[(lane + 32 * ((it + i) & 63))];
            if (MODE == 1) atomicAdd(&sm[lane + 32 * ((it + i) & 63)], 1u);
            if (MODE == 2) atomicAdd(&sm[(it + i) & 255], 1u);
            if (MODE == 3) atomicAdd(&sm[((it + i) & 63) * 4 + (lane >> 3)], 1u);
            if (MODE == 4) atomicAdd(&sm[(rnd >> 24)], 1u);
            if (MODE == 5) acc += reinterpret_cast<uint8_t*>(sm)[lane * 3 + 96 * ((it + i) & 63)];
            if (MODE == 6) atomicAdd(&sm[((rnd >> 24) << 3) + (lane & 7)], 1u);
            if (MODE == 7) { uint4 q = reinterpret_cast<uint4*>(sm)[lane + 32 * ((it + i) & 31)]; acc += q.x ^ q.y ^ q.z ^ q.w; }
            if (MODE == 8) atomicAdd(&sm[((rnd >> 27) + 100)], 1u);   // 32 hot bins (leaf-like clustering)
            if (MODE == 9) atomicAdd(&sm[(((rnd >> 27) + 100) << 3) + (lane & 7)], 1u);  // hot bins, 8 replicas
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + sm[threadIdx.x];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <typename F>
void run(const char* name, F launch, int nw) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 8);
    launch(out, cyc); launch(out, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double c = 0; for (int i = 0; i < 148; ++i) c += h[i]; c /= 148;
    double winst = (double)ITERS * ILP * nw;
    printf("%-34s warps=%2d  cycles=%9.0f  warp-inst/clk/SM=%6.3f  (lane-ops/clk/SM=%7.1f)\n", name, nw, c, winst / c, winst * 32 / c);
    cudaFree(out); cudaFree(cyc);
}

#define ALU(OP, NAME) for (int nw : {8, 16, 32}) run(NAME, [&](uint32_t* o, long long* c) { k_alu<OP><<<148, nw * 32>>>(o, c, 12345u); }, nw);
#define SMEM(M, NAME) for (int nw : {8, 16}) run(NAME, [&](uint32_t* o, long long* c) { k_smem<M><<<148, nw * 32>>>(o, c, 12345u); }, nw);

int main() {
    ALU(0, "imad"); ALU(1, "dp2a"); ALU(2, "dp4a"); ALU(3, "prmt"); ALU(4, "lop3"); ALU(5, "shf"); ALU(6, "iadd");
    ALU(7, "max3"); ALU(8, "vadd4"); ALU(9, "bfe+add"); ALU(10, "ffma"); ALU(11, "popc+add"); ALU(12, "mul.wide.u16+add");
    ALU(13, "mad24"); ALU(14, "vadd2"); ALU(15, "imad+shr");
    for (int nw : {8, 16, 32}) run("mix imad/lop3", [&](uint32_t* o, long long* c) { k_mix<<<148, nw * 32>>>(o, c, 12345u); }, nw);
    SMEM(0, "lds32"); SMEM(7, "lds128"); SMEM(5, "lds.u8 stride3"); SMEM(1, "atoms distinct"); SMEM(2, "atoms same addr");
    SMEM(3, "atoms 4 addrs/warp"); SMEM(4, "atoms random 256 bins"); SMEM(6, "atoms random 256 bins x8 replicas");
    SMEM(8, "atoms 32 hot bins"); SMEM(9, "atoms 32 hot bins x8 replicas");
    return 0;
}
