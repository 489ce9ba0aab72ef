// Host-side (no CUDA) batch parameter drawing for the class-balancing augment tasks: what
// `_process_single_transformation` (srcs/preprocessing/dataset_balancer.py:201-207) draws for one task --
// a fresh ImageAugmenter(seed) seeds Python's `random` (image_augmenter.py:16-18, only `if seed:`) and the
// method then consumes the stream in the reference's order (image_augmenter.py:22,35,50,79-80,100-105,126).
// At 36,864 tasks the interpreter's `random.seed()` + draws cost ~13 us per task and were 95 % of the
// balancing wall time; this restatement runs ~1.5 us per task per host thread.
//
// CPython's `random` is restated from its published algorithm (MT19937ar `init_by_array` seeding with the
// 32-bit chunks of abs(seed); `random()` = genrand_res53; `_randbelow_with_getrandbits` rejection sampling for
// `choice` / `randint`); PIL's rotate geometry (Image.rotate with expand=True) and libImaging's 16.16 affine
// coefficients follow SURVEY.md A.1.  tests/test_params_cpu.py checks every output against the interpreter.
#include <math.h>
#include <sched.h>
#include <stdlib.h>
#include <string.h>

#include <thread>
#include <vector>

#include "lfx_common.cuh"

namespace {

constexpr int MT_N = 624, MT_M = 397;

constexpr int SEED_LANES = 8;   // tasks seeded side by side (see seed_lanes)

// State words are addressed through a stride so that one generator can live in its own array (stride 1) or in one lane
// of the interleaved block that seed_lanes fills (stride SEED_LANES).
struct PyRandom {
    uint32_t own[MT_N];
    uint32_t* mt = own;
    int stride = 1;
    int idx;  // next word of the current block; words < idx are already twisted in place
    uint32_t& w(int i) { return mt[(size_t)i * stride]; }

    static const uint32_t* base_state() {  // init_genrand(19650218), shared by every init_by_array
        static uint32_t base[MT_N];
        static bool done = [] {
            uint32_t s = 19650218u;
            for (int i = 0; i < MT_N; ++i) {
                base[i] = s;
                s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)(i + 1);
            }
            return true;
        }();
        (void)done;
        return base;
    }

    // init_by_array({key}, 1) for SEED_LANES keys at once.  The recurrence is a 1247-step dependent chain per key
    // (multiply - xor - add, ~6 cycles a step): one key at a time leaves the core idle most of the time, eight
    // interleaved chains fill its pipelines (and vectorise).  st[i][l] = word i of lane l.
    // (AVX2 clone picked at load time where the CPU has it: eight 32-bit lanes = one vector, native 32-bit multiply)
    __attribute__((target_clones("avx2", "default"))) static void seed_lanes(const uint32_t* keys, uint32_t (*st)[SEED_LANES]) {
        const uint32_t* base = base_state();
        for (int i = 0; i < MT_N; ++i)
            for (int l = 0; l < SEED_LANES; ++l) st[i][l] = base[i];
        // first loop of init_by_array (k = 624 steps from i = 1): words 1..623, then the wrap (mt[0] = mt[623]) and word 1 again;
        // second loop (623 steps): words 2..623, the wrap, word 1 -- written as straight loops so that they vectorise
        for (int i = 1; i < MT_N; ++i)
            for (int l = 0; l < SEED_LANES; ++l) st[i][l] = (st[i][l] ^ ((st[i - 1][l] ^ (st[i - 1][l] >> 30)) * 1664525u)) + keys[l];
        for (int l = 0; l < SEED_LANES; ++l) st[0][l] = st[MT_N - 1][l];
        for (int l = 0; l < SEED_LANES; ++l) st[1][l] = (st[1][l] ^ ((st[0][l] ^ (st[0][l] >> 30)) * 1664525u)) + keys[l];
        for (int i = 2; i < MT_N; ++i)
            for (int l = 0; l < SEED_LANES; ++l) st[i][l] = (st[i][l] ^ ((st[i - 1][l] ^ (st[i - 1][l] >> 30)) * 1566083941u)) - (uint32_t)i;
        for (int l = 0; l < SEED_LANES; ++l) st[0][l] = st[MT_N - 1][l];
        for (int l = 0; l < SEED_LANES; ++l) st[1][l] = (st[1][l] ^ ((st[0][l] ^ (st[0][l] >> 30)) * 1566083941u)) - 1u;
        for (int l = 0; l < SEED_LANES; ++l) st[0][l] = 0x80000000u;
    }
    void attach(uint32_t (*st)[SEED_LANES], int lane) {   // this generator = lane `lane` of a seeded block
        mt = &st[0][lane];
        stride = SEED_LANES;
        idx = 0;
    }

    void seed(uint32_t key) {  // random.seed(int) for 0 <= int < 2^32: init_by_array({key}, 1)
        mt = own;
        stride = 1;
        memcpy(mt, base_state(), sizeof(own));
        int i = 1;
        for (int k = MT_N; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key;  // + init_key[0] + j, j == 0
            if (++i >= MT_N) {
                mt[0] = mt[MT_N - 1];
                i = 1;
            }
        }
        for (int k = MT_N - 1; k; --k) {
            mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
            if (++i >= MT_N) {
                mt[0] = mt[MT_N - 1];
                i = 1;
            }
        }
        mt[0] = 0x80000000u;
        idx = 0;
    }

    // The reference generator twists a whole block in place in ascending order; doing the same word by word on
    // demand is identical (word i only reads words that are still old, or already new, exactly as there).
    // words drawn elsewhere (lfx_seed_words): the first n_pre tempered outputs of this seed; `starved` is set when a draw
    // needs more than that (the caller then seeds on the host and draws again)
    const uint32_t* pre = nullptr;
    int n_pre = 0, i_pre = 0;
    bool starved = false;
    void attach_words(const uint32_t* wds, int n) {
        pre = wds;
        n_pre = n;
        i_pre = 0;
        starved = false;
    }

    uint32_t next_u32() {
        if (pre) {
            if (i_pre < n_pre) return pre[i_pre++];
            starved = true;
            return 0u;          // randbelow(n) accepts 0: every rejection loop ends
        }
        if (idx == MT_N) idx = 0;
        const int i = idx++;
        const int i1 = (i + 1 == MT_N) ? 0 : i + 1;
        const int im = (i + MT_M >= MT_N) ? i + MT_M - MT_N : i + MT_M;
        const uint32_t y = (w(i) & 0x80000000u) | (w(i1) & 0x7FFFFFFFu);
        uint32_t v = w(im) ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
        w(i) = v;
        v ^= (v >> 11);
        v ^= (v << 7) & 0x9D2C5680u;
        v ^= (v << 15) & 0xEFC60000u;
        v ^= (v >> 18);
        return v;
    }
    double random() {
        const uint32_t a = next_u32() >> 5, b = next_u32() >> 6;
        return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
    }
    double uniform(double a, double b) { return a + (b - a) * random(); }
    uint32_t randbelow(uint32_t n) {  // n >= 1; getrandbits(k) with k = n.bit_length() <= 32
        int k = 0;
        for (uint32_t t = n; t; t >>= 1) ++k;
        uint32_t r = next_u32() >> (32 - k);
        while (r >= n) r = next_u32() >> (32 - k);
        return r;
    }
};

// round(v, 15) = the nearest double to the decimal that is v correctly rounded (half-even on its exact binary value) to 15
// places.  Exact integer form: |v| = m / 2^sh, N = round_half_even(m * 10^15 / 2^sh) in 128-bit arithmetic, result N / 10^15
// (N < 2^53 and 10^15 are exact doubles, so the one division is the correctly rounded strtod of "0.ddd...").  Equal to
// snprintf("%.15f") + strtod on 5 M random sines / cosines and the half-way cases; ~20x cheaper.
double py_round15(double v) {
    if (v == 0.0 || !(fabs(v) < 4.0)) {
        char buf[64];
        snprintf(buf, sizeof buf, "%.15f", v);
        return strtod(buf, nullptr);
    }
    int e;
    const double fr = frexp(fabs(v), &e);
    const uint64_t m = (uint64_t)ldexp(fr, 53);
    const int sh = 53 - e;
    if (sh >= 120) return copysign(0.0, v);
    const unsigned __int128 P = (unsigned __int128)m * 1000000000000000ULL;
    unsigned __int128 q = P >> sh;
    const unsigned __int128 rem = P & ((((unsigned __int128)1) << sh) - 1), half = ((unsigned __int128)1) << (sh - 1);
    if (rem > half || (rem == half && (q & 1))) ++q;
    return copysign((double)(uint64_t)q / 1e15, v);
}

double py_floordiv(double vx, double wx) {  // float.__floordiv__
    double mod = fmod(vx, wx);
    double div = (vx - mod) / wx;
    if (mod != 0.0 && ((wx < 0) != (mod < 0))) div -= 1.0;
    if (div == 0.0) return copysign(0.0, vx / wx);
    double fl = floor(div);
    if (div - fl > 0.5) fl += 1.0;
    return fl;
}

inline int32_t fix16(double v) { return (int32_t)floor(v * 65536.0 + 0.5); }

// PIL Image.rotate(angle, resample=NEAREST, expand=True) -> 16.16 inverse-affine coefficients + output size.
void rotate_params(double angle, int w, int h, int32_t* ip) {
    angle = fmod(angle, 360.0);
    if (angle < 0) angle += 360.0;
    double m[6] = {1.0, 0.0, 0.0, 0.0, 1.0, 0.0};
    int nw = w, nh = h;
    if (angle != 0.0) {  // 90/180/270 cannot come out of uniform(-30, 30)
        const double cx = w / 2.0, cy = h / 2.0;
        const double a = -(angle * (M_PI / 180.0));
        const double c = py_round15(cos(a)), s = py_round15(sin(a)), ns = py_round15(-sin(a));
        m[0] = c, m[1] = s, m[2] = 0.0, m[3] = ns, m[4] = c, m[5] = 0.0;
        const double t2 = m[0] * -cx + m[1] * -cy + m[2], t5 = m[3] * -cx + m[4] * -cy + m[5];
        m[2] = t2 + cx;
        m[5] = t5 + cy;
        const double px[4] = {0.0, (double)w, (double)w, 0.0}, py[4] = {0.0, 0.0, (double)h, (double)h};
        double xmin = 0, xmax = 0, ymin = 0, ymax = 0;
        for (int k = 0; k < 4; ++k) {
            const double X = m[0] * px[k] + m[1] * py[k] + m[2], Y = m[3] * px[k] + m[4] * py[k] + m[5];
            if (k == 0 || X < xmin) xmin = X;
            if (k == 0 || X > xmax) xmax = X;
            if (k == 0 || Y < ymin) ymin = Y;
            if (k == 0 || Y > ymax) ymax = Y;
        }
        nw = (int)(ceil(xmax) - floor(xmin));
        nh = (int)(ceil(ymax) - floor(ymin));
        const double ox = -(nw - w) / 2.0, oy = -(nh - h) / 2.0;
        const double n2 = m[0] * ox + m[1] * oy + m[2], n5 = m[3] * ox + m[4] * oy + m[5];
        m[2] = n2;
        m[5] = n5;
    }
    ip[0] = fix16(m[0]);
    ip[1] = fix16(m[1]);
    ip[2] = fix16(m[2] + m[0] * 0.5 + m[1] * 0.5);
    ip[3] = fix16(m[3]);
    ip[4] = fix16(m[4]);
    ip[5] = fix16(m[5] + m[3] * 0.5 + m[4] * 0.5);
    ip[6] = nw;
    ip[7] = nh;
}

void draw_one(PyRandom& r, int transform, int H, int W, int32_t* ip, double* dp) {   // r: freshly seeded with the task seed
    memset(ip, 0, 8 * sizeof(int32_t));
    for (int k = 0; k < 8; ++k) dp[k] = 0.0;
    switch (transform) {
        case LFX_AUG_FLIP:  // random.choice([True, False]): index 0 -> FLIP_LEFT_RIGHT
            ip[0] = r.randbelow(2) == 0 ? 0 : 1;
            break;
        case LFX_AUG_ROTATE:
            dp[0] = r.uniform(-30.0, 30.0);
            rotate_params(dp[0], W, H, ip);
            break;
        case LFX_AUG_SKEW: {
            const double s = r.uniform(0.05, 0.15);
            dp[0] = 1 + s, dp[2] = -s * W, dp[4] = 1 + s, dp[5] = -s * H;
            ip[0] = 1;  // PERSPECTIVE
            break;
        }
        case LFX_AUG_SHEAR: {
            const double k = r.uniform(-0.2, 0.2);
            dp[0] = 1.0, dp[4] = 1.0;
            if (r.randbelow(2) == 0) dp[1] = k; else dp[3] = k;
            ip[0] = 0;  // AFFINE
            break;
        }
        case LFX_AUG_CROP: {
            const double ratio = r.uniform(0.8, 0.95);
            const int nw = (int)(W * ratio), nh = (int)(H * ratio);
            ip[2] = nw, ip[3] = nh;
            ip[0] = (int32_t)r.randbelow((uint32_t)(W - nw + 1));
            ip[1] = (int32_t)r.randbelow((uint32_t)(H - nh + 1));
            break;
        }
        case LFX_AUG_DISTORTION:
            dp[0] = r.uniform(0.0, 2.0);
            ip[0] = (int32_t)py_floordiv((double)((long long)H * W) * dp[0], 100.0);
            break;
    }
}

}  // namespace

extern "C" int lfx_draw_augment_params(const int32_t* transform, const uint32_t* seed, int B, int H, int W, int32_t* iparams,
                                       double* dparams, int threads) {
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(transform && seed && iparams && dparams && B > 0 && H > 0 && W > 0, LFX_ERR_ARG, "draw_augment_params: bad arguments");
    for (int i = 0; i < B; ++i)
        LFX_REQUIRE(transform[i] >= LFX_AUG_FLIP && transform[i] <= LFX_AUG_DISTORTION, LFX_ERR_ARG, "draw_augment_params: unknown transform %d at task %d",
                    transform[i], i);
    PyRandom::base_state();
    int nt = threads;
    if (nt <= 0) {   // the cores this process may run on (a rank bound to a slice of the box must not spawn a thread per box core)
        cpu_set_t set;
        nt = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
    }
    nt = nt < 1 ? 1 : (nt > 64 ? 64 : nt);
    if (B < 2048) nt = 1;
    auto work = [&](int lo, int hi) {
        uint32_t st[MT_N][SEED_LANES];
        uint32_t keys[SEED_LANES];
        PyRandom r;
        for (int i0 = lo; i0 < hi; i0 += SEED_LANES) {
            const int n = hi - i0 < SEED_LANES ? hi - i0 : SEED_LANES;
            for (int l = 0; l < SEED_LANES; ++l) keys[l] = seed[i0 + (l < n ? l : 0)];
            PyRandom::seed_lanes(keys, st);
            for (int l = 0; l < n; ++l) {
                r.attach(st, l);
                draw_one(r, transform[i0 + l], H, W, iparams + (size_t)(i0 + l) * 8, dparams + (size_t)(i0 + l) * 8);
            }
        }
    };
    if (nt == 1) {
        work(0, B);
    } else {
        std::vector<std::thread> pool;
        const int per = (B + nt - 1) / nt;
        for (int t = 0; t < nt; ++t) {
            const int lo = t * per, hi = lo + per < B ? lo + per : B;
            if (lo < hi) pool.emplace_back(work, lo, hi);
        }
        for (auto& th : pool) th.join();
    }
    return LFX_OK;
}

// lfx_draw_augment_params with the seeding done on the device: words[B][nwords] (HOST copy of lfx_seed_words' output) = the
// first outputs of each task's stream.  A task whose draws need more words (rejection sampling: ~2^-(nwords-3) of the tasks)
// is seeded here as lfx_draw_augment_params would.
extern "C" int lfx_draw_augment_params_words(const int32_t* transform, const uint32_t* seed, const uint32_t* words, int nwords, int B,
                                             int H, int W, int32_t* iparams, double* dparams) {
    if (B == 0) return LFX_OK;
    LFX_REQUIRE(transform && seed && words && iparams && dparams && B > 0 && H > 0 && W > 0 && nwords >= 4, LFX_ERR_ARG,
                "draw_augment_params_words: bad arguments");
    PyRandom r;
    for (int i = 0; i < B; ++i) {
        LFX_REQUIRE(transform[i] >= LFX_AUG_FLIP && transform[i] <= LFX_AUG_DISTORTION, LFX_ERR_ARG,
                    "draw_augment_params_words: unknown transform %d at task %d", transform[i], i);
        r.attach_words(words + (size_t)i * nwords, nwords);
        draw_one(r, transform[i], H, W, iparams + (size_t)i * 8, dparams + (size_t)i * 8);
        if (r.starved) {
            r.pre = nullptr;
            r.seed(seed[i]);
            draw_one(r, transform[i], H, W, iparams + (size_t)i * 8, dparams + (size_t)i * 8);
        }
    }
    return LFX_OK;
}

// The task list of the balancing pass (dataset_balancer.py:115-129): per class, per transform, per copy the reference draws
// `random.choice(source_images)` and `random.randint(0, 1000000)` from ONE stream seeded by `random.seed(seed)` (:31).  For
// in-memory datasets the choice is an index: group g (one class x transform pair, in plan order) has group_count[g] tasks
// drawn from a class of group_class_size[g] images -> local_index[t] = _randbelow(size), task_seed[t] = _randbelow(1000001).
extern "C" int lfx_draw_balance_tasks(uint32_t seed, int ngroups, const int32_t* group_count, const int32_t* group_class_size,
                                      int32_t* local_index, int32_t* task_seed) {
    LFX_REQUIRE(ngroups >= 0 && (ngroups == 0 || (group_count && group_class_size && local_index && task_seed)), LFX_ERR_ARG,
                "draw_balance_tasks: bad arguments");
    PyRandom r;
    r.seed(seed);
    size_t t = 0;
    for (int g = 0; g < ngroups; ++g) {
        LFX_REQUIRE(group_count[g] >= 0 && group_class_size[g] > 0, LFX_ERR_ARG, "draw_balance_tasks: group %d has count %d, class size %d", g,
                    group_count[g], group_class_size[g]);
        for (int i = 0; i < group_count[g]; ++i, ++t) {
            local_index[t] = (int32_t)r.randbelow((uint32_t)group_class_size[g]);
            task_seed[t] = (int32_t)r.randbelow(1000001u);
        }
    }
    return LFX_OK;
}
