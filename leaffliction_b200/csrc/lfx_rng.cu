// NumPy's LEGACY normal stream on the GPU: what `np.random.seed(s); np.random.normal(loc, scale, n).astype(np.uint8)`
// produces (srcs/preprocessing/image_augmenter.py:121-123 after ImageAugmenter(seed) seeded the global stream,
// :16-18).  MT19937 (init_genrand seeding, in-place twist, tempering), genrand_res53 doubles, polar-method
// gauss with its one-value cache -- restated in parallel form: attempt t consumes
// words 4t..4t+3; the k-th ACCEPTED attempt yields normals 2k = f*x2 and 2k+1 = f*x1.
//
// One warp per stream (image): the 624-word state lives in shared memory; the twist runs as one ascending pass in
// chunks of 32 (loads, __syncwarp, stores -- every operand is then either still old or already new exactly as
// in the sequential generator); each 624-word block is exactly 156 attempts, compacted with warp ballots.
// All floating point is explicit round-to-nearest fp64 without FMA contraction (the host libraries are built
// without FMA); CUDA's log() is within 1 ulp of glibc's, which can move a result only when 5*g lies within an
// ulp of an integer (probability ~1e-15 per sample) -- the uint8 stream is otherwise identical.
#include "lfx_common.cuh"

namespace {

constexpr int RNG_WARPS = 8;

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= (y >> 18);
    return y;
}

__global__ void __launch_bounds__(RNG_WARPS * 32) k_legacy_normal_u8(const uint32_t* __restrict__ seeds, uint8_t* __restrict__ out, int B,
                                                                    int n, double loc, double scale) {
    __shared__ uint32_t s_mt[RNG_WARPS][624];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int b = blockIdx.x * RNG_WARPS + wid;
    if (b >= B) return;
    uint32_t* mt = s_mt[wid];
    if (lane == 0) {  // init_genrand (sequential recurrence)
        uint32_t s = seeds[b];
        for (int i = 0; i < 624; ++i) {
            mt[i] = s;
            s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)(i + 1);
        }
    }
    __syncwarp();
    uint8_t* o = out + (size_t)b * n;
    int produced = 0;  // normals written so far (warp-uniform)
    while (produced < n) {
        // ---- twist: next 624 words, in place, ascending chunks of 32
        for (int c0 = 0; c0 < 624; c0 += 32) {
            const int i = c0 + lane;
            uint32_t v = 0;
            if (i < 624) {
                const int i1 = (i + 1 == 624) ? 0 : i + 1;
                const int im = (i + 397 >= 624) ? i + 397 - 624 : i + 397;
                const uint32_t y = (mt[i] & 0x80000000u) | (mt[i1] & 0x7FFFFFFFu);
                v = mt[im] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
            }
            __syncwarp();
            if (i < 624) mt[i] = v;
            __syncwarp();
        }
        // Note on i = 623: it needs the NEW mt[0] and mt[396]; both were stored by earlier chunks.  Lanes of the
        // last chunk read mt[i+1] of their right neighbour before anyone stores (loads precede the __syncwarp).
        // ---- 156 attempts of this block
        for (int t0 = 0; t0 < 156 && produced < n; t0 += 32) {
            const int t = t0 + lane;
            bool acc = false;
            double g0 = 0.0, g1 = 0.0;
            if (t < 156) {
                const uint32_t a = mt_temper(mt[4 * t]) >> 5, bq = mt_temper(mt[4 * t + 1]) >> 6;
                const uint32_t c = mt_temper(mt[4 * t + 2]) >> 5, d = mt_temper(mt[4 * t + 3]) >> 6;
                const double d1 = __ddiv_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)bq), 9007199254740992.0);
                const double d2 = __ddiv_rn(__dadd_rn(__dmul_rn((double)c, 67108864.0), (double)d), 9007199254740992.0);
                const double x1 = __dadd_rn(__dmul_rn(2.0, d1), -1.0), x2 = __dadd_rn(__dmul_rn(2.0, d2), -1.0);
                const double r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
                acc = (r2 < 1.0) && (r2 != 0.0);
                if (acc) {
                    const double f = __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2));
                    g0 = __dadd_rn(loc, __dmul_rn(scale, __dmul_rn(f, x2)));  // returned first
                    g1 = __dadd_rn(loc, __dmul_rn(scale, __dmul_rn(f, x1)));  // the cached value, returned next
                }
            }
            const uint32_t bal = __ballot_sync(0xffffffffu, acc);
            if (acc) {
                const int k = produced + 2 * __popc(bal & ((1u << lane) - 1u));
                if (k < n) o[k] = (uint8_t)(int)g0;          // .astype(np.uint8): truncate toward zero, wrap mod 256
                if (k + 1 < n) o[k + 1] = (uint8_t)(int)g1;
            }
            produced += 2 * __popc(bal);
        }
    }
}

}  // namespace

extern "C" int lfx_legacy_normal_u8(const uint32_t* seeds, uint8_t* out, int B, int n, double loc, double scale, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0 || n == 0) return LFX_OK;
    LFX_REQUIRE(seeds && out && B > 0 && n > 0, LFX_ERR_ARG, "legacy_normal_u8: bad arguments");
    k_legacy_normal_u8<<<lfx_div_up(B, RNG_WARPS), RNG_WARPS * 32, 0, (cudaStream_t)stream>>>(seeds, out, B, n, loc, scale);
    return lfx_check_launch("legacy_normal_u8");
}
