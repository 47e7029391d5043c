// CPU proof-by-search for the FMA division sequence of the select kernel (mcts.cu: div_by_count).
//
//   q0 = RN(a * y)            y = RN(1 / b)
//   r0 = RN(a - b * q0)       (fma, exact)
//   q1 = RN(q0 + r0 * y)      (fma)
//
// must equal the IEEE quotient RN(a / b) bit for bit for every integer divisor 1..65535 (visit counts): a / b with a
// small integer b is never closer than 2^-17 ulp to a rounding boundary, the error of q0 + r0 * y before its rounding
// is ~2^-53 ulp.  This program searches for counterexamples anyway (random numerators of many significand patterns
// inside the kernel's +-2^400 exponent window, and small integer numerators).
//
//   gcc -O2 -mfma -ffp-contract=off -o /tmp/fastdiv_check tools/fastdiv_check.c -lm && /tmp/fastdiv_check [percent]
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static uint64_t s[2] = {0x9E3779B97F4A7C15ull, 0xD1B54A32D192ED03ull};
static inline uint64_t rnd(void) {
  uint64_t a = s[0], b = s[1];
  s[0] = b;
  a ^= a << 23;
  s[1] = a ^ b ^ (a >> 17) ^ (b >> 26);
  return s[1] + b;
}
static inline double bits2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t d2bits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

static inline double div1(double a, double b, double y) {
  const double q0 = a * y;
  const double r0 = __builtin_fma(-b, q0, a);
  return __builtin_fma(r0, y, q0);
}
// random double with biased exponent in [elo, ehi], random sign, significand drawn from a mix of patterns
static double rand_double(int elo, int ehi) {
  uint64_t m = rnd() & 0xFFFFFFFFFFFFFull;
  switch (rnd() & 7) {
    case 0: m &= ~((1ull << (rnd() % 52)) - 1); break;            // trailing zeros (short significands)
    case 1: m |= (1ull << (rnd() % 52)) - 1; break;               // trailing ones
    case 2: m = 0xFFFFFFFFFFFFFull - (rnd() & 0xFF); break;       // near all ones
    case 3: m = rnd() & 0xFF; break;                              // near a power of two
    default: break;
  }
  const uint64_t e = (uint64_t)(elo + (int)(rnd() % (uint64_t)(ehi - elo + 1)));
  return bits2d(((rnd() & 1) << 63) | (e << 52) | m);
}

int main(int argc, char** argv) {
  // optional argument: scale in percent of the full search (tests run 5)
  const long pct = argc > 1 ? atol(argv[1]) : 100;
  unsigned long long bad1 = 0, n1 = 0;
  // (1) integer divisors
  for (int b = 1; b <= 65535; ++b) {
    const double db = (double)b, y = 1.0 / db;
    for (long i = 0; i < 30 * pct; ++i) {
      const double a = rand_double(1023 - 400, 1023 + 400);
      const double want = a / db, got = div1(a, db, y);
      ++n1;
      if (d2bits(want) != d2bits(got)) {
        if (bad1 < 10) printf("int divisor mismatch: %a / %d: %a vs %a\n", a, b, want, got);
        ++bad1;
      }
    }
    // integer numerators as well (sums of +-1 outcomes, small table values)
    for (int a = -300; a <= 300; ++a) {
      if (a == 0) continue;
      const double want = (double)a / db, got = div1((double)a, db, y);
      ++n1;
      if (d2bits(want) != d2bits(got)) { if (bad1 < 10) printf("int/int mismatch %d / %d\n", a, b); ++bad1; }
    }
  }
  printf("integer divisors: %llu checked, %llu mismatches\n", n1, bad1);
  return bad1 ? 1 : 0;
}
