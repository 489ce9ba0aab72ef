// NumPy's LEGACY normal stream on the GPU: what `np.random.seed(s); np.random.normal(loc, scale, n).astype(np.uint8)`
// produces (srcs/preprocessing/image_augmenter.py:121-123 after ImageAugmenter(seed) seeded the global stream,
// :16-18).  MT19937 (init_genrand seeding, in-place twist, tempering), genrand_res53 doubles, polar-method
// gauss with its one-value cache -- restated in parallel form: attempt t consumes
// words 4t..4t+3; the k-th ACCEPTED attempt yields normals 2k = f*x2 and 2k+1 = f*x1.
//
// One warp per stream (image): the 624-word state lives in shared memory.  A 624-word block is exactly 156 attempts
// and attempt t consumes words 4t..4t+3, so lane j of round c twists the four words of attempt t = 32c + j itself
// (ascending quads; every operand is either still old or already new exactly as in the sequential generator: word i
// reads i+1 -- old, passed between neighbouring lanes by shuffle, or new[0] for i = 623 -- and i+397, which is old for
// i < 227 and was stored by an earlier round otherwise), tempers the two words the fast path needs straight from its
// registers and evaluates its attempt; accepted attempts are compacted with a warp ballot.  ~105 warp instructions per
// round (the first version, which twisted the block, re-read it and branched per attempt, needed 195).
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

// (cur & 0x80000000) | (nxt & 0x7FFFFFFF) as ONE bit-select (LOP3 0xE4), then the twist: 5 instructions per word
__device__ __forceinline__ uint32_t mt_twist5(uint32_t cur, uint32_t nxt, uint32_t far) {
    uint32_t y;
    asm("lop3.b32 %0, %1, %2, 0x80000000, 0xE4;" : "=r"(y) : "r"(cur), "r"(nxt));
    uint32_t mag;   // (nxt & 1) * 0x9908B0DF as an IMAD: the multiply pipe is idle, the logic pipe is the busy one
    asm("mul.lo.u32 %0, %1, 0x9908B0DF;" : "=r"(mag) : "r"(nxt & 1u));
    return far ^ (y >> 1) ^ mag;
}
// the top 27 bits of the tempered word (genrand_res53's a = genrand() >> 5), last temper step folded into the shift
__device__ __forceinline__ uint32_t mt_temper_hi27(uint32_t y) {
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9D2C5680u;
    y ^= (y << 15) & 0xEFC60000u;
    return (y >> 5) ^ (y >> 23);
}

// FAST: loc == 0 and |scale| <= 8 (the fp32 error bound in the header); otherwise every attempt is evaluated in fp64.
template <bool FAST>
__global__ void __launch_bounds__(RNG_WARPS * 32) k_legacy_normal_u8(const uint32_t* __restrict__ seeds, uint8_t* __restrict__ out, int B,
                                                                    int n, double loc, double scale) {
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
    asm volatile("" : "+l"(o));   // keep the stream's base in registers (the compiler re-derived it from the parameters per store)
    const bool pair_ok = ((reinterpret_cast<uintptr_t>(o) & 1) == 0);
    const float fscale = (float)scale;
    const uint32_t lt_mask = (1u << lane) - 1u;
    // Round c (0..4) gives lane j the quad of attempt t = 32c + j: words i0 = 128c + 4j .. i0 + 3.  Word i needs word i + 1
    // (old; the first word of the NEXT lane's quad comes by shuffle, so no lane reads a word another lane stores in the
    // same round) and word i + 397 mod 624: quad i0 + 396 (.y .z .w) and the word after it.  i0 + 396 < 624 in round 0
    // and for lanes < 25 of round 1 (old values, stored by rounds 3 / 4 only); it wraps to i0 - 228 otherwise (new
    // values, stored by earlier rounds).  Everything but round 1's lane-dependent wrap folds into immediate offsets.
    uint32_t* const mq = mt + 4 * lane;
    const uint32_t* const far1 = mq + 128 + (lane < 25 ? 396 : -228);
    const uint32_t* const far1n = (lane == 24) ? mt : far1 + 4;   // word 624 = word 0 (new)
    constexpr uint32_t ACC = 1u << 31, SLOW = 1u << 30;   // ACC in the sign bit: one compare
    int produced = 0;  // normals written so far (warp-uniform, always even)
    while (produced < n) {
        // Result of a round, one word per lane: byte 0 / 1 = the two output bytes, bit 31 = accepted, bit 30 = fp64 has
        // to decide.  The rounds are unrolled: the fp32 chain (log, reciprocal, square root) of one hides behind the
        // twist of the next.
        uint32_t res[5];
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const bool act = (c < 4) || (lane < 28);   // 156 = 4 * 32 + 28 attempts per block
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            if (act) q = *reinterpret_cast<const uint4*>(mq + 128 * c);
            uint32_t nx = __shfl_down_sync(0xffffffffu, q.x, 1);
            if (c < 4) {
                if (lane == 31) nx = mq[128 * c + 4];   // first word of the next round: still old
            } else {
                if (lane == 27) nx = mt[0];             // word 624 = word 0, new since round 0
            }
            uint32_t w0 = 0u, w2 = 0u;
            if (act) {
                const uint32_t* fp = (c == 0) ? mq + 396 : (c == 1) ? far1 : mq + (128 * c - 228);
                const uint32_t* fn = (c == 1) ? far1n : fp + 4;
                const uint4 m = *reinterpret_cast<const uint4*>(fp);
                const uint32_t m3 = *fn;
                w0 = mt_twist5(q.x, q.y, m.y);
                const uint32_t w1 = mt_twist5(q.y, q.z, m.z);
                w2 = mt_twist5(q.z, q.w, m.w);
                const uint32_t w3 = mt_twist5(q.w, nx, m3);
                *reinterpret_cast<uint4*>(mq + 128 * c) = make_uint4(w0, w1, w2, w3);
            }
            __syncwarp();   // the next rounds read these words
            uint32_t r = 0u;
            if (!FAST) {
                r = act ? SLOW : 0u;
            } else {
                // the low words (26 of the 53 bits of each double) move x by less than 2^-26 = u/4: the fast path leaves
                // them untempered (budgeted above); the fp64 re-evaluation tempers all four words.  Branch-free: a warp
                // always holds accepted attempts, so rejected lanes just compute a value nobody selects.
                const uint32_t a = mt_temper_hi27(w0), cc = mt_temper_hi27(w2);
                const float x1 = (float)((int)a - (1 << 26)) * 0x1p-26f;
                const float x2 = (float)((int)cc - (1 << 26)) * 0x1p-26f;
                const float r2 = fmaf(x1, x1, x2 * x2);
                // fast units (MUFU lg2 / rcp / rsq: a few ulp each, |L| abs 4e-7 near 1) -- inside the error budget above
                float l2, rc, rs;
                asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(r2));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"(r2));
                const float q3 = (l2 * -1.3862943611198906f) * rc;   // -2 ln(r2) / r2
                asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(q3));
                const float f = fscale * (q3 * rs);
                const float g0 = f * x2, g1 = f * x1;
                const int i0 = (int)g0, i1 = (int)g1;   // |g| < 0.5 truncates to 0 whatever the error
                const bool k0 = (fabsf(g0) < 0.5f) || (fabsf(g0 - rintf(g0)) >= LFX_RNG_EPS);
                const bool k1 = (fabsf(g1) < 0.5f) || (fabsf(g1 - rintf(g1)) >= LFX_RNG_EPS);
                // 1e-8 <= r2 < 1 - 1e-6: decided here; r2 > 1 + 1e-6: rejected; between (the acceptance test itself, or
                // -- 2^-54 of the attempts -- whether r2 is zero): the exact radius is needed
                const bool in_fast = (r2 >= 1e-8f) && (r2 < 1.f - 1e-6f);
                const bool out = r2 > 1.f + 1e-6f;
                const uint32_t bytes = __byte_perm((uint32_t)i0 & 0xFFu, (uint32_t)i1, 0x3340);
                r = in_fast ? (bytes | ACC | ((k0 && k1) ? 0u : SLOW)) : (out ? 0u : SLOW);
                if (!act) r = 0u;
            }
            res[c] = r;
        }
        // the rare fp64 re-evaluations after all five fast rounds (0.15 % of the attempts: one warp in five has any)
        if (__any_sync(0xffffffffu, ((res[0] | res[1] | res[2] | res[3] | res[4]) & SLOW) != 0u)) {
#pragma unroll   // (res[] indexed by a run-time c would live in local memory)
            for (int c = 0; c < 5; ++c) {
                if (res[c] & SLOW) {
                    const uint4 q = *reinterpret_cast<const uint4*>(mt + 4 * (c * 32 + lane));
                    uint8_t b0 = 0, b1 = 0;
                    const bool ok = attempt64(mt_temper(q.x) >> 5, mt_temper(q.y) >> 6, mt_temper(q.z) >> 5, mt_temper(q.w) >> 6, loc, scale, b0, b1);
                    res[c] = ok ? ((uint32_t)b0 | ((uint32_t)b1 << 8) | ACC) : 0u;
                }
            }
        }
        if (pair_ok && produced + 2 * 156 <= n) {   // the whole block fits: 16-bit stores, no per-lane bound checks
            uint32_t pairs = (uint32_t)produced >> 1;
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                const bool acc = (int)res[c] < 0;
                const uint32_t bal = __ballot_sync(0xffffffffu, acc);
                uint64_t dst;   // o + 2 * (pairs + rank among the accepted lanes): one IMAD.WIDE
                asm("mad.wide.u32 %0, %1, 2, %2;" : "=l"(dst) : "r"(pairs + __popc(bal & lt_mask)), "l"(o));
                if (acc) *reinterpret_cast<uint16_t*>(dst) = (uint16_t)res[c];
                pairs += __popc(bal);
            }
            produced = (int)(pairs << 1);
        } else {
#pragma unroll
            for (int c = 0; c < 5; ++c) {
                const bool acc = (int)res[c] < 0;
                const uint32_t bal = __ballot_sync(0xffffffffu, acc);
                if (acc) {
                    const int k = produced + 2 * __popc(bal & lt_mask);
                    if (k < n) o[k] = (uint8_t)res[c];
                    if (k + 1 < n) o[k + 1] = (uint8_t)(res[c] >> 8);
                }
                produced += 2 * __popc(bal);
            }
        }
    }
}

// CPython's random.seed(int) for 0 <= int < 2^32 (MT19937ar init_genrand(19650218) + init_by_array({key}, 1)) and the
// first `nwords` tempered 32-bit outputs, one thread per task seed: the 1247-step dependent seeding chain that costs a host
// core ~1.8 us per task (lfx_params.cu) is ~25 us of latency here for ANY number of tasks.  The state lives in local memory
// (word i of every thread of a warp is one coalesced line); the first outputs only read words n, n+1 and n+397 of the
// seeded state (n < 227), so no block twist is needed.
__global__ void __launch_bounds__(128) k_seed_words(const uint32_t* __restrict__ seeds, int B, int nwords, uint32_t* __restrict__ words) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= B) return;
    uint32_t mt[624];
    const uint32_t key = seeds[t];
    uint32_t base = 19650218u, prev = base;
    mt[0] = prev;
#pragma unroll 8
    for (int i = 1; i < 624; ++i) {          // init_genrand word i, then the first init_by_array loop on it
        base = 1812433253u * (base ^ (base >> 30)) + (uint32_t)i;
        prev = (base ^ ((prev ^ (prev >> 30)) * 1664525u)) + key;
        mt[i] = prev;
    }
    uint32_t m0 = prev;                      // the wrap: mt[0] = mt[623], then word 1 again (624th step)
    uint32_t m1 = (mt[1] ^ ((m0 ^ (m0 >> 30)) * 1664525u)) + key;
    prev = m1;
#pragma unroll 8
    for (int i = 2; i < 624; ++i) {          // second loop: 623 steps from word 2
        prev = (mt[i] ^ ((prev ^ (prev >> 30)) * 1566083941u)) - (uint32_t)i;
        mt[i] = prev;
    }
    m0 = prev;
    m1 = (m1 ^ ((m0 ^ (m0 >> 30)) * 1566083941u)) - 1u;
    mt[1] = m1;
    mt[0] = 0x80000000u;
    for (int n = 0; n < nwords; ++n) {
        const uint32_t y = (mt[n] & 0x80000000u) | (mt[n + 1] & 0x7FFFFFFFu);
        uint32_t v = mt[n + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
        v ^= (v >> 11);
        v ^= (v << 7) & 0x9D2C5680u;
        v ^= (v << 15) & 0xEFC60000u;
        v ^= (v >> 18);
        words[(size_t)t * nwords + n] = v;
    }
}

}  // namespace

extern "C" int lfx_seed_words(const uint32_t* seeds, int B, int nwords, uint32_t* words, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(seeds && words && B > 0 && nwords >= 1 && nwords <= 226, LFX_ERR_ARG, "seed_words: bad arguments (1 <= nwords <= 226)");
    k_seed_words<<<lfx_div_up(B, 128), 128, 0, (cudaStream_t)stream>>>(seeds, B, nwords, words);
    return lfx_check_launch("seed_words");
}

extern "C" int lfx_legacy_normal_u8(const uint32_t* seeds, uint8_t* out, int B, int n, double loc, double scale, lfx_stream_t stream) {
    LFX_REQUIRE_READY();
    if (B == 0 || n == 0) return LFX_OK;
    LFX_REQUIRE(seeds && out && B > 0 && n > 0, LFX_ERR_ARG, "legacy_normal_u8: bad arguments");
    const bool fast = (loc == 0.0 && fabs(scale) <= 8.0);   // the fp32 error bound above assumes |g| <= 8 * 12
    if (fast)
        k_legacy_normal_u8<true><<<lfx_div_up(B, RNG_WARPS), RNG_WARPS * 32, 0, (cudaStream_t)stream>>>(seeds, out, B, n, loc, scale);
    else
        k_legacy_normal_u8<false><<<lfx_div_up(B, RNG_WARPS), RNG_WARPS * 32, 0, (cudaStream_t)stream>>>(seeds, out, B, n, loc, scale);
    return lfx_check_launch("legacy_normal_u8");
}
