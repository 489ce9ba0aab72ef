// NumPy's LEGACY normal stream on the GPU: what `np.random.seed(s); np.random.normal(loc, scale, n).astype(np.uint8)`
// produces (srcs/preprocessing/image_augmenter.py:121-123 after ImageAugmenter(seed) seeded the global stream,
// :16-18).  MT19937 (init_genrand seeding, in-place twist, tempering), genrand_res53 doubles, polar-method
// gauss with its one-value cache -- restated in parallel form: attempt t consumes
// words 4t..4t+3; the k-th ACCEPTED attempt yields normals 2k = f*x2 and 2k+1 = f*x1.
//
// One warp per stream (image): the 624-word state lives in shared memory.  A 624-word block is exactly 156 attempts
// and attempt t consumes words 4t..4t+3, so lane j of round c twists the four words of attempt t = 32c + j itself
// (ascending quads, loads - __syncwarp - stores: every operand is then either still old or already new exactly as in
// the sequential generator: word i reads i+1 -- old, or new[0] for i = 623 -- and i+397, which is old for i < 227 and
// was stored by an earlier round otherwise), tempers them in registers and evaluates its attempt; accepted attempts
// are compacted with a warp ballot.
//
// Arithmetic.  The reference is fp64 (explicit round-to-nearest without FMA contraction: the host libraries are built
// without FMA); CUDA's log() is within 1 ulp of glibc's, which can move a result only when 5*g lies within an ulp of
// an integer (probability ~1e-15 per sample).  An fp32 evaluation on the fast units decides first.  With u = 2^-24:
// x1, x2 carry 2.25u (the fast path drops the low 26 bits of each double: < 2^-26), r2 5u, L = __logf(r2) an absolute 5u + 4e-7 (its bound on [0.5, 2]; 2 ulp elsewhere), the reciprocal
// and the reciprocal square root a few ulp, and g = scale * x * sqrt(-2L / r2) the absolute error
// |g| (0.5 dL/|L| + ~12u).  A value with |g| >= 0.5 has |L| >= 0.005 (x^2 <= r2), so for |scale| <= 8 the error stays
// below 1.5e-4; below 0.5 the byte is 0 whatever the error.  Attempts whose fp32 radius lies within 1e-6 of 1 (the
// acceptance test) or whose value lies within 2^-11 of an integer are re-evaluated in fp64 (0.15 % of the attempts).
// q > 0 always (r2 < 1), and q * rsqrt(q) = sqrt(q).
#include "lfx_common.cuh"

namespace {

constexpr int RNG_WARPS = 8;
#define LFX_RNG_EPS 0.00048828125f

__device__ __forceinline__ uint32_t mt_temper(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    y ^= (y >> 18);
    return y;
}
__device__ __forceinline__ uint32_t mt_twist(uint32_t cur, uint32_t nxt, uint32_t far) {
    const uint32_t y = (cur & 0x80000000u) | (nxt & 0x7FFFFFFFu);
    return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
}

// fp64 reference evaluation of one attempt -> accepted?, the two bytes
__device__ __noinline__ bool attempt64(uint32_t a, uint32_t bq, uint32_t c, uint32_t d, double loc, double scale, uint8_t& o0, uint8_t& o1) {
    const double d1 = __ddiv_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)bq), 9007199254740992.0);
    const double d2 = __ddiv_rn(__dadd_rn(__dmul_rn((double)c, 67108864.0), (double)d), 9007199254740992.0);
    const double x1 = __dadd_rn(__dmul_rn(2.0, d1), -1.0), x2 = __dadd_rn(__dmul_rn(2.0, d2), -1.0);
    const double r2 = __dadd_rn(__dmul_rn(x1, x1), __dmul_rn(x2, x2));
    if (!((r2 < 1.0) && (r2 != 0.0))) return false;
    const double f = __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2));
    const double g0 = __dadd_rn(loc, __dmul_rn(scale, __dmul_rn(f, x2)));  // returned first
    const double g1 = __dadd_rn(loc, __dmul_rn(scale, __dmul_rn(f, x1)));  // the cached value, returned next
    o0 = (uint8_t)(int)g0;  // .astype(np.uint8): truncate toward zero, wrap mod 256
    o1 = (uint8_t)(int)g1;
    return true;
}

// byte of an fp32 value known to within 7e-5; false when fp64 has to decide
__device__ __forceinline__ bool byte32(float g, uint8_t& o) {
    if (fabsf(g) < 0.5f) {
        o = 0;
        return true;
    }
    o = (uint8_t)(int)g;
    return fabsf(g - rintf(g)) >= LFX_RNG_EPS;
}

__global__ void __launch_bounds__(RNG_WARPS * 32) k_legacy_normal_u8(const uint32_t* __restrict__ seeds, uint8_t* __restrict__ out, int B,
                                                                    int n, double loc, double scale, int fast) {
    __shared__ __align__(16) uint32_t s_mt[RNG_WARPS][624];
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
    const bool pair_ok = ((reinterpret_cast<uintptr_t>(o) & 1) == 0);
    const float fscale = (float)scale;
    int produced = 0;  // normals written so far (warp-uniform, always even)
    while (produced < n) {
        // ---- phase 1: twist the whole 624-word block in place, five rounds of one quad per lane
#pragma unroll 1
        for (int t0 = 0; t0 < 156; t0 += 32) {
            const int t = t0 + lane;
            const bool act = t < 156;
            uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
            if (act) {
                const int i0 = 4 * t;
                const uint4 q = *reinterpret_cast<const uint4*>(mt + i0);
                const uint32_t nx = mt[i0 + 4 == 624 ? 0 : i0 + 4];
                int j = i0 + 396;  // quad holding words i0+397..i0+399 in .y .z .w (i0 + 397 = 1 mod 4)
                if (j >= 624) j -= 624;
                const uint4 m = *reinterpret_cast<const uint4*>(mt + j);
                const uint32_t m3 = mt[j + 4 == 624 ? 0 : j + 4];
                w0 = mt_twist(q.x, q.y, m.y);
                w1 = mt_twist(q.y, q.z, m.z);
                w2 = mt_twist(q.z, q.w, m.w);
                w3 = mt_twist(q.w, nx, m3);
            }
            __syncwarp();
            if (act) *reinterpret_cast<uint4*>(mt + 4 * t) = make_uint4(w0, w1, w2, w3);
            __syncwarp();
        }
        // ---- phase 2: the 156 attempts of the block.  The five rounds are independent of each other (no barrier, no
        // shared-memory write), so they are unrolled: the long fp32 chains (log, divide, square root) of one round hide
        // behind the others.  Result of a round, one word per lane: byte 0 / 1 = the two output bytes, bit 16 = accepted,
        // bit 17 = fp64 has to decide (kept in registers: arrays of flags and bytes ended up in local memory).
        constexpr uint32_t ACC = 1u << 16, SLOW = 1u << 17;
        uint32_t res[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const int t = c * 32 + lane;
            uint32_t r = 0u;
            if (t < 156) {
                if (!fast) {
                    r = SLOW;
                } else {
                    // the low words (26 of the 53 bits of each double) move x by less than 2^-26 = u/4: the fast path
                    // leaves them untempered (budgeted above); the fp64 re-evaluation tempers all four words
                    const uint2 q = *reinterpret_cast<const uint2*>(mt + 4 * t);
                    const uint2 q2 = *reinterpret_cast<const uint2*>(mt + 4 * t + 2);
                    const uint32_t a = mt_temper(q.x) >> 5, cc = mt_temper(q2.x) >> 5;
                    const float x1 = (float)((int)a - (1 << 26)) * 0x1p-26f;
                    const float x2 = (float)((int)cc - (1 << 26)) * 0x1p-26f;
                    const float r2 = fmaf(x1, x1, x2 * x2);
                    if (fabsf(r2 - 1.f) <= 1e-6f || r2 < 1e-8f) {
                        r = SLOW;   // the acceptance test (or, 2^-54 of the attempts, whether r2 is zero) needs the exact radius
                    } else if (r2 < 1.f) {
                        // fast units (MUFU lg2 / rcp / rsq: a few ulp each, |L| abs 4e-7 near 1) -- inside the error budget above
                        const float q3 = __fdividef(-2.f * __logf(r2), r2);
                        float rs;
                        asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(q3));
                        const float f = fscale * (q3 * rs);
                        uint8_t b0, b1;
                        const bool k0 = byte32(f * x2, b0), k1 = byte32(f * x1, b1);
                        r = (uint32_t)b0 | ((uint32_t)b1 << 8) | ACC | ((k0 && k1) ? 0u : SLOW);
                    }
                }
            }
            res[c] = r;
        }
        // the rare fp64 re-evaluations after all five fast rounds (0.15 % of the attempts: one warp in five has any)
        if (__any_sync(0xffffffffu, ((res[0] | res[1] | res[2] | res[3] | res[4]) & SLOW) != 0u)) {
#pragma unroll 1
            for (int c = 0; c < 5; ++c) {
                if (res[c] & SLOW) {
                    const uint4 q = *reinterpret_cast<const uint4*>(mt + 4 * (c * 32 + lane));
                    uint8_t b0 = 0, b1 = 0;
                    const bool ok = attempt64(mt_temper(q.x) >> 5, mt_temper(q.y) >> 6, mt_temper(q.z) >> 5, mt_temper(q.w) >> 6, loc, scale, b0, b1);
                    res[c] = ok ? ((uint32_t)b0 | ((uint32_t)b1 << 8) | ACC) : 0u;
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const bool acc = (res[c] & ACC) != 0u;
            const uint32_t bal = __ballot_sync(0xffffffffu, acc);
            if (acc) {
                const int k = produced + 2 * __popc(bal & ((1u << lane) - 1u));
                if (k + 1 < n && pair_ok) {
                    *reinterpret_cast<uint16_t*>(o + k) = (uint16_t)res[c];
                } else {
                    if (k < n) o[k] = (uint8_t)res[c];
                    if (k + 1 < n) o[k + 1] = (uint8_t)(res[c] >> 8);
                }
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
    const int fast = (loc == 0.0 && fabs(scale) <= 8.0) ? 1 : 0;   // the fp32 error bound above assumes |g| <= 8 * 12
    k_legacy_normal_u8<<<lfx_div_up(B, RNG_WARPS), RNG_WARPS * 32, 0, (cudaStream_t)stream>>>(seeds, out, B, n, loc, scale, fast);
    return lfx_check_launch("legacy_normal_u8");
}
