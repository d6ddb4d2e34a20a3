// randomizer.cpp -- host-side restatement of BEAST 2's beast.base.util.Randomizer stream.
//
// [EXT] BEAST 2 core is not under the reference tree; this follows the published
// MersenneTwisterFast (Sean Luke) variant BEAST ships, as recalled in SURVEY.md Appendix F.2:
// old Knuth seeding mt[i] = 69069 * mt[i-1], nextDouble from two tempered words
// ((y >>> 6) << 27 + (z >>> 5)) / 2^53, nextInt with the power-of-two fast path.  Call sites
// in the reference: operators/UniformizationRetypeOperator.java:79,156,165 (nextDouble),
// operators/NodeRetype.java:37,48 (nextInt).  PARITY UNPINNED against a real JVM.
//
// The uniformization kernels take a window of uniforms per job and report how many they
// consumed, so the host peeks a window and then skips what was used.
#include <stdint.h>
#include <string.h>

namespace {

struct MTFast {
    static constexpr int N = 624, M = 397;
    uint32_t mt[N];
    int mti;

    void seed(int64_t s)
    {
        mt[0] = (uint32_t)(int32_t)s;
        for (mti = 1; mti < N; mti++)
            mt[mti] = 69069u * mt[mti - 1];
    }
    uint32_t gen()
    {
        static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
        uint32_t y;
        if (mti >= N) {
            int kk;
            for (kk = 0; kk < N - M; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + M] ^ (y >> 1) ^ mag01[y & 1];
            }
            for (; kk < N - 1; kk++) {
                y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
                mt[kk] = mt[kk + (M - N)] ^ (y >> 1) ^ mag01[y & 1];
            }
            y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
            mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ mag01[y & 1];
            mti = 0;
        }
        y = mt[mti++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
    double nextDouble()
    {
        uint32_t y = gen(), z = gen();
        return (double)((((uint64_t)(y >> 6)) << 27) + (z >> 5)) / 9007199254740992.0;
    }
    int32_t nextInt(int32_t n)
    {
        if ((n & -n) == n) {
            uint32_t y = gen();
            return (int32_t)(((int64_t)n * (int64_t)(y >> 1)) >> 31);
        }
        int32_t bits, val;
        do {
            uint32_t y = gen();
            bits = (int32_t)(y >> 1);
            val = bits % n;
        } while (bits - val + (n - 1) < 0);
        return val;
    }
};

}  // namespace

extern "C" {

// A seed whose low 32 bits are zero would leave the generator in the all-zero state (every draw
// 0 forever, and the operators' rejection loops would never end); BEAST's Randomizer refuses a
// zero seed, so do these entries: mtt_rng_new returns NULL, mtt_rng_set_seed returns -1.
void *mtt_rng_new(int64_t seed)
{
    if ((uint32_t)seed == 0u)
        return nullptr;
    MTFast *r = new MTFast();
    r->seed(seed);
    return r;
}
void mtt_rng_free(void *h) { delete (MTFast *)h; }
int mtt_rng_set_seed(void *h, int64_t seed)
{
    if (!h || (uint32_t)seed == 0u)
        return -1;
    ((MTFast *)h)->seed(seed);
    return 0;
}
double mtt_rng_next_double(void *h) { return ((MTFast *)h)->nextDouble(); }
int32_t mtt_rng_next_int(void *h, int32_t n) { return n > 0 ? ((MTFast *)h)->nextInt(n) : -1; }
void mtt_rng_next_doubles(void *h, int64_t n, double *out)
{
    MTFast *r = (MTFast *)h;
    for (int64_t i = 0; i < n; i++)
        out[i] = r->nextDouble();
}
// the next n doubles WITHOUT advancing the stream
void mtt_rng_peek_doubles(const void *h, int64_t n, double *out)
{
    MTFast tmp;
    memcpy(&tmp, h, sizeof(MTFast));
    for (int64_t i = 0; i < n; i++)
        out[i] = tmp.nextDouble();
}
void mtt_rng_skip_doubles(void *h, int64_t n)
{
    MTFast *r = (MTFast *)h;
    for (int64_t i = 0; i < n; i++)
        r->nextDouble();
}

}  // extern "C"
