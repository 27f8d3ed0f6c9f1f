#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
static uint64_t s[2] = {0x9E3779B97F4A7C15ull, 0xBF58476D1CE4E5B9ull};
static inline uint64_t rnd(void) { uint64_t s1 = s[0], s0 = s[1]; s[0] = s0; s1 ^= s1 << 23; s[1] = s1 ^ s0 ^ (s1 >> 18) ^ (s0 >> 5); return s[1] + s0; }
int main(int argc, char** argv) {
    long long n = argc > 1 ? atoll(argv[1]) : 200000000LL, bad3 = 0, bad5 = 0;
    for (long long i = 0; i < n; ++i) {
        uint64_t r1 = rnd(), r2 = rnd();
        // a: a float (or bf16 on odd i) value with random sign/exponent in a modest range
        uint32_t fb = (uint32_t)(r1 & 0x807FFFFFu) | ((uint32_t)(100 + (r1 >> 40) % 56) << 23);
        if (i & 1) fb &= 0xFFFF0000u;
        float af; memcpy(&af, &fb, 4);
        double a = (double)af;
        // b: a positive double with random mantissa (sometimes all ones / near powers of two), exponent in [-20, 20]
        uint64_t mant = r2 & 0xFFFFFFFFFFFFFull;
        int sel = (int)((r2 >> 52) & 15);
        if (sel == 0) mant = 0xFFFFFFFFFFFFFull;
        else if (sel == 1) mant = 0xFFFFFFFFFFFFFull - (r2 >> 60);
        else if (sel == 2) mant = (r2 >> 60);
        uint64_t bb = ((uint64_t)(1023 - 20 + (r1 >> 56) % 41) << 52) | mant;
        double b; memcpy(&b, &bb, 8);
        double want = a / b;
        double y = 1.0 / b;
        double q0 = a * y;
        double e0 = fma(-b, q0, a);
        double q1 = fma(e0, y, q0);
        if (q1 != want) ++bad3;
        double e1 = fma(-b, q1, a);
        double q2 = fma(e1, y, q1);
        if (q2 != want) { if (bad5 < 5) printf("5-op mismatch a=%a b=%a want=%a got=%a\n", a, b, want, q2); ++bad5; }
    }
    printf("trials %lld: 3-op mismatches %lld, 5-op mismatches %lld\n", n, bad3, bad5);
    return 0;
}
