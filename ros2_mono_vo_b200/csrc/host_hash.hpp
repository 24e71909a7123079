// host_hash.hpp -- 64-bit content hash of host buffers (the key of the device-resident caches).
// Four independent multiply-rotate lanes over 32-byte blocks (memory-bound on the host: ~10+ GB/s), folded with the
// length.  Not cryptographic: a cache key; a collision needs two different frames / descriptor blocks of the same size
// with the same 64-bit hash (2^-64 per pair).
#pragma once
#include <stdint.h>
#include <string.h>

namespace mvo {

struct ContentHasher {
  uint64_t h[4] = {0x243F6A8885A308D3ull, 0x13198A2E03707344ull, 0xA4093822299F31D0ull, 0x082EFA98EC4E6C89ull};
  uint64_t total = 0;
  static inline uint64_t rotl(uint64_t v, int r) { return (v << r) | (v >> (64 - r)); }
  inline void block(const uint8_t* p) {
    uint64_t w[4];
    memcpy(w, p, 32);
    h[0] = rotl(h[0] ^ w[0], 29) * 0x9E3779B97F4A7C15ull;
    h[1] = rotl(h[1] ^ w[1], 31) * 0xC2B2AE3D27D4EB4Full;
    h[2] = rotl(h[2] ^ w[2], 27) * 0x165667B19E3779F9ull;
    h[3] = rotl(h[3] ^ w[3], 33) * 0x85EBCA77C2B2AE63ull;
  }
  inline void update(const uint8_t* p, size_t n) {
    total += n;
    while (n >= 32) {
      block(p);
      p += 32;
      n -= 32;
    }
    if (n) {
      uint8_t tail[32] = {0};
      memcpy(tail, p, n);
      tail[31] = (uint8_t)n;
      block(tail);
    }
  }
  inline uint64_t finish() const {
    uint64_t v = total * 0xD6E8FEB86659FD93ull;
    for (int i = 0; i < 4; ++i) {
      v ^= h[i];
      v = rotl(v, 23) * 0x9FB21C651E98DF25ull;
      v ^= v >> 29;
    }
    return v ? v : 1;   // 0 is "no hash"
  }
};

inline uint64_t content_hash(const uint8_t* p, size_t n) {
  ContentHasher c;
  c.update(p, n);
  return c.finish();
}
// rows x row_bytes bytes of a strided image
inline uint64_t content_hash_rows(const uint8_t* p, int rows, size_t row_bytes, size_t stride) {
  if (stride == row_bytes) return content_hash(p, (size_t)rows * row_bytes);
  ContentHasher c;
  for (int r = 0; r < rows; ++r) c.update(p + (size_t)r * stride, row_bytes);
  return c.finish();
}

}  // namespace mvo
