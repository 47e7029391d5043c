// numpy's legacy MT19937 stream driven by a whole warp (shared by the tree kernels and the replay sampler).
#pragma once
#include <stdint.h>

namespace mz {

// ---------------------------------------------------------------------------
// numpy legacy MT19937, one stream per tree, driven by a whole warp
// ---------------------------------------------------------------------------
struct WarpRng {
  uint32_t* key;
  int pos;
  int lane;
  unsigned long long draws, twists;

  __device__ void load(uint32_t* k, const int* pos_ptr, int ln) {
    key = k; pos = *pos_ptr; lane = ln; draws = 0; twists = 0;
  }
  __device__ void store(int* pos_ptr) const {
    if (lane == 0) *pos_ptr = pos;
  }
  // genrand regeneration: chunks of 32 consecutive words in ascending order;
  // inside a chunk every lane reads its three inputs before any lane writes, which
  // reproduces the sequential recurrence (k[i+1] old, k[i+397 mod 624] new iff < i).
  __device__ void twist() {
    __syncwarp();
    for (int c = 0; c < 20; ++c) {
      const int i = c * 32 + lane;
      uint32_t a = 0, b = 0, s = 0;
      if (i < 624) {
        a = key[i];
        b = key[(i + 1) % 624];
        s = key[(i + 397) % 624];
      }
      __syncwarp();
      if (i < 624) {
        const uint32_t y = (a & 0x80000000u) | (b & 0x7fffffffu);
        key[i] = s ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      __syncwarp();
    }
    pos = 0;
    ++twists;
  }
  __device__ uint32_t next_u32() {
    if (pos >= 624) twist();
    uint32_t y = key[pos];
    ++pos;
    ++draws;
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  // random_sample(): (a >> 5, b >> 6) -> 53-bit double
  __device__ double next_double() {
    const uint32_t a = next_u32() >> 5, b = next_u32() >> 6;
    return __ddiv_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 9007199254740992.0);
  }
  // randint(0, k) of RandomState.choice: masked rejection on 32-bit draws, no draw for k == 1
  __device__ uint32_t bounded(uint32_t k) {
    const uint32_t rng = k - 1;
    if (rng == 0) return 0;
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    do { v = next_u32() & mask; } while (v > rng);
    return v;
  }
};


// The same stream driven by ONE thread (thread-per-tree kernels for tiny action spaces): sequential genrand twist.
struct ThreadRng {
  uint32_t* key;
  int pos;
  unsigned long long draws, twists;

  __device__ void load(uint32_t* k, const int* pos_ptr) { key = k; pos = *pos_ptr; draws = 0; twists = 0; }
  __device__ void store(int* pos_ptr) const { *pos_ptr = pos; }
  __device__ void twist() {
    const uint32_t kUp = 0x80000000u, kLo = 0x7fffffffu, kMat = 0x9908b0dfu;
    int kk = 0;
    for (; kk < 624 - 397; ++kk) {
      const uint32_t y = (key[kk] & kUp) | (key[kk + 1] & kLo);
      key[kk] = key[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? kMat : 0u);
    }
    for (; kk < 623; ++kk) {
      const uint32_t y = (key[kk] & kUp) | (key[kk + 1] & kLo);
      key[kk] = key[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? kMat : 0u);
    }
    const uint32_t y = (key[623] & kUp) | (key[0] & kLo);
    key[623] = key[396] ^ (y >> 1) ^ ((y & 1u) ? kMat : 0u);
    pos = 0;
    ++twists;
  }
  __device__ uint32_t next_u32() {
    if (pos >= 624) twist();
    uint32_t y = key[pos];
    ++pos;
    ++draws;
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  __device__ double next_double() {
    const uint32_t a = next_u32() >> 5, b = next_u32() >> 6;
    return __ddiv_rn(__dadd_rn(__dmul_rn((double)a, 67108864.0), (double)b), 9007199254740992.0);
  }
  __device__ uint32_t bounded(uint32_t k) {
    const uint32_t rng = k - 1;
    if (rng == 0) return 0;
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t v;
    do { v = next_u32() & mask; } while (v > rng);
    return v;
  }
};

}  // namespace mz
