/*
 * ORACLE (test infrastructure only -- never linked into the product path).
 *
 * Plain-C restatement of the NumPy *legacy* RandomState stream that the reference's
 * latent sampler depends on:
 *   /root/reference/models/wrappers.py:167-175   StyleGAN2.sample_latent
 *       seed = np.random.randint(np.iinfo(np.int32).max); RandomState(seed).standard_normal(512*n)
 *   /root/reference/decomposition.py:226-227     np.random.seed(config.seed or SEED_SAMPLING)
 *   /root/reference/models/biggan/pytorch_biggan/pytorch_pretrained_biggan/utils.py:21-33
 *       truncnorm.rvs(..., random_state=RandomState(seed))  -> RandomState.uniform -> random_sample
 *
 * The arithmetic lives in a third-party dependency (NumPy, un-pinned by the reference,
 * environment.yml:14; 2.3.5 in this image).  NumPy freezes the legacy stream by policy (NEP 19), so
 * the published algorithm is restated here:
 *   - MT19937 (Matsumoto & Nishimura 1998), Knuth-style init_genrand seeding for integer seeds
 *   - random_sample(): 53-bit double from two consecutive outputs, (a>>5, b>>6)
 *   - legacy_gauss(): Marsaglia polar method, pair emitted as [f*x2, f*x1]
 *   - randint(high) for high-1 <= 0xFFFFFFFF: masked rejection on 32-bit outputs
 * Pinned against numpy itself in tests/test_oracle_rng.py (bit-exact).
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o oracle/_build/libmt19937_legacy.so oracle/mt19937_legacy.c -lm
 */
#include <stdint.h>
#include <math.h>
#include <stddef.h>

#define MT_N 624
#define MT_M 397

typedef struct {
    uint32_t mt[MT_N];
    int pos;
    int has_gauss;
    double gauss;
} gso_mt_t;

void gso_mt_seed(gso_mt_t *s, uint32_t seed) {
    s->mt[0] = seed;
    for (int i = 1; i < MT_N; ++i)
        s->mt[i] = 1812433253u * (s->mt[i - 1] ^ (s->mt[i - 1] >> 30)) + (uint32_t)i;
    s->pos = MT_N;
    s->has_gauss = 0;
    s->gauss = 0.0;
}

static void gso_mt_twist(gso_mt_t *s) {
    uint32_t *mt = s->mt;
    int i;
    for (i = 0; i < MT_N - MT_M; ++i) {
        uint32_t y = (mt[i] & 0x80000000u) | (mt[i + 1] & 0x7fffffffu);
        mt[i] = mt[i + MT_M] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; i < MT_N - 1; ++i) {
        uint32_t y = (mt[i] & 0x80000000u) | (mt[i + 1] & 0x7fffffffu);
        mt[i] = mt[i + (MT_M - MT_N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    uint32_t y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    s->pos = 0;
}

uint32_t gso_mt_next_u32(gso_mt_t *s) {
    if (s->pos == MT_N) gso_mt_twist(s);
    uint32_t y = s->mt[s->pos++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

double gso_mt_next_double(gso_mt_t *s) {
    uint32_t a = gso_mt_next_u32(s) >> 5, b = gso_mt_next_u32(s) >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
}

double gso_legacy_gauss(gso_mt_t *s) {
    if (s->has_gauss) {
        s->has_gauss = 0;
        double t = s->gauss;
        s->gauss = 0.0;
        return t;
    }
    double f, x1, x2, r2;
    do {
        x1 = 2.0 * gso_mt_next_double(s) - 1.0;
        x2 = 2.0 * gso_mt_next_double(s) - 1.0;
        r2 = x1 * x1 + x2 * x2;
    } while (r2 >= 1.0 || r2 == 0.0);
    f = sqrt(-2.0 * log(r2) / r2);
    s->gauss = f * x1;
    s->has_gauss = 1;
    return f * x2;
}

/* RandomState.randint(high) with 0 < high-1 < 0xFFFFFFFF (legacy masked rejection). */
uint32_t gso_legacy_randint(gso_mt_t *s, uint32_t high) {
    uint32_t rng = high - 1u;
    if (rng == 0u) return 0u;
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    do { v = gso_mt_next_u32(s) & mask; } while (v > rng);
    return v;
}

/* ---- flat entry points for ctypes ---- */

size_t gso_state_size(void) { return sizeof(gso_mt_t); }

/* RandomState(seed).standard_normal(n).astype(float32) */
void gso_standard_normal_f32(uint32_t seed, int64_t n, float *out) {
    gso_mt_t s;
    gso_mt_seed(&s, seed);
    for (int64_t i = 0; i < n; ++i) out[i] = (float)gso_legacy_gauss(&s);
}

void gso_standard_normal_f64(uint32_t seed, int64_t n, double *out) {
    gso_mt_t s;
    gso_mt_seed(&s, seed);
    for (int64_t i = 0; i < n; ++i) out[i] = gso_legacy_gauss(&s);
}

/* RandomState(seed).random_sample(n) */
void gso_random_sample_f64(uint32_t seed, int64_t n, double *out) {
    gso_mt_t s;
    gso_mt_seed(&s, seed);
    for (int64_t i = 0; i < n; ++i) out[i] = gso_mt_next_double(&s);
}

/* raw tempered 32-bit outputs */
void gso_raw_u32(uint32_t seed, int64_t n, uint32_t *out) {
    gso_mt_t s;
    gso_mt_seed(&s, seed);
    for (int64_t i = 0; i < n; ++i) out[i] = gso_mt_next_u32(&s);
}

/* np.random.seed(seed0); [np.random.randint(high) for _ in range(count)] */
void gso_randint_sequence(uint32_t seed0, uint32_t high, int64_t count, uint32_t *out) {
    gso_mt_t s;
    gso_mt_seed(&s, seed0);
    for (int64_t i = 0; i < count; ++i) out[i] = gso_legacy_randint(&s, high);
}
