// Device-side helpers shared by the scan / join / aggregate kernels: streaming vector loads, segment decoders for
// ValueSegment / DictionarySegment / FrameOfReferenceSegment, warp scans, decoupled look-back.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "internal.hpp"

namespace hyb {

constexpr unsigned kFullMask = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------------------------------
// Streaming loads/stores: every column byte is read once, so do not let it displace dictionaries in L1.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ld_stream_v4(const void* ptr) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(ptr));
  return r;
}

__device__ __forceinline__ uint2 ld_stream_v2(const void* ptr) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(ptr));
  return r;
}

__device__ __forceinline__ uint32_t ld_stream_u32(const void* ptr) {
  uint32_t r;
  asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(ptr));
  return r;
}

__device__ __forceinline__ void st_stream_v2(void* ptr, uint32_t x, uint32_t y) {
  asm volatile("st.global.L1::no_allocate.v2.u32 [%0], {%1, %2};" ::"l"(ptr), "r"(x), "r"(y) : "memory");
}

__device__ __forceinline__ void st_stream_v4(void* ptr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(ptr), "r"(x), "r"(y), "r"(z), "r"(w)
               : "memory");
}

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* ptr) {
  unsigned long long r;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(r) : "l"(ptr) : "memory");
  return r;
}

__device__ __forceinline__ void st_volatile_u64(unsigned long long* ptr, unsigned long long value) {
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(ptr), "l"(value) : "memory");
}

// Bytes per row of the vector a scan streams (codes of compressed segments, values of unencoded ones); 0 = bit-packed.
__device__ __forceinline__ uint32_t segment_stream(const DevSegment& segment, const char*& base) {
  if (segment.encoding == HYB_ENC_UNENCODED) {
    base = static_cast<const char*>(segment.values);
    return (segment.data_type == HYB_TYPE_INT32 || segment.data_type == HYB_TYPE_FLOAT32) ? 4u : 8u;
  }
  base = static_cast<const char*>(segment.av);
  switch (segment.vector_type) {
    case HYB_VEC_FIXED_1B:
      return 1u;
    case HYB_VEC_FIXED_2B:
      return 2u;
    case HYB_VEC_FIXED_4B:
      return 4u;
    default:
      return 0u;
  }
}

// L2 prefetch hint for the line holding `row` of a segment's streamed vector (no registers held, no fault on misuse of
// the hint: callers still keep `row` inside the segment).
__device__ __forceinline__ void prefetch_codes(const DevSegment& segment, uint32_t row) {
  const char* base;
  const uint32_t bytes_per_row = segment_stream(segment, base);
  if (bytes_per_row == 0) return;
  asm volatile("prefetch.global.L2 [%0];" ::"l"(base + static_cast<size_t>(row) * bytes_per_row));
}


// ---------------------------------------------------------------------------------------------------------------------
// Compressed-vector decode: 8 consecutive entries starting at `row0` (row0 % 8 == 0) of a FixedWidthIntegerVector
// (fixed_width_integer_vector.hpp:29-33) or BitPackingVector (compact_iterator.hpp:218-252: entry i occupies bits
// [i*b, (i+1)*b) of a little-endian uint64 word stream, LSB first, spilling into the next word).
// Entries past `rows` are returned as garbage/zero; callers mask them. Buffers carry a 64-byte tail pad (Arena).
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_codes8(const void* base, uint32_t vector_type, uint32_t bit_width, uint32_t row0,
                                            uint32_t rows, uint32_t (&codes)[8]) {
  switch (vector_type) {
    case HYB_VEC_FIXED_1B: {
      const uint2 v = ld_stream_v2(static_cast<const uint8_t*>(base) + row0);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        codes[j] = (v.x >> (8 * j)) & 0xFFu;
        codes[4 + j] = (v.y >> (8 * j)) & 0xFFu;
      }
      break;
    }
    case HYB_VEC_FIXED_2B: {
      const uint4 v = ld_stream_v4(static_cast<const uint16_t*>(base) + row0);
      codes[0] = v.x & 0xFFFFu;
      codes[1] = v.x >> 16;
      codes[2] = v.y & 0xFFFFu;
      codes[3] = v.y >> 16;
      codes[4] = v.z & 0xFFFFu;
      codes[5] = v.z >> 16;
      codes[6] = v.w & 0xFFFFu;
      codes[7] = v.w >> 16;
      break;
    }
    case HYB_VEC_FIXED_4B: {
      const uint4 a = ld_stream_v4(static_cast<const uint32_t*>(base) + row0);
      const uint4 b = ld_stream_v4(static_cast<const uint32_t*>(base) + row0 + 4);
      codes[0] = a.x;
      codes[1] = a.y;
      codes[2] = a.z;
      codes[3] = a.w;
      codes[4] = b.x;
      codes[5] = b.y;
      codes[6] = b.z;
      codes[7] = b.w;
      break;
    }
    default: {  // HYB_VEC_BITPACKED
      const auto* words = static_cast<const unsigned long long*>(base);
      const unsigned long long mask = (bit_width >= 64) ? ~0ull : ((1ull << bit_width) - 1ull);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        codes[j] = 0;
        if (row0 + j < rows) {
          const unsigned long long bit = static_cast<unsigned long long>(row0 + j) * bit_width;
          const unsigned long long word = bit >> 6;
          const uint32_t shift = static_cast<uint32_t>(bit & 63);
          unsigned long long value = __ldg(words + word) >> shift;
          if (shift + bit_width > 64) value |= __ldg(words + word + 1) << (64 - shift);
          codes[j] = static_cast<uint32_t>(value & mask);
        }
      }
      break;
    }
  }
}

// ---- early loads ---------------------------------------------------------------------------------------------------
// A kernel that reads several columns per row wants ALL of an iteration's loads in flight before the first use; decoders
// that load inside per-encoding branches serialise one DRAM round trip per column instead. The code vectors that dominate
// (FixedWidthIntegerVector of 1 or 2 bytes: dictionary value-IDs, FoR offsets) are therefore fetched up front as raw
// 16 bytes per 8 rows and unpacked later; every other layout keeps loading at its point of use.
__device__ __forceinline__ bool raw_codes_loadable(const DevSegment& segment) {
  return segment.encoding != HYB_ENC_UNENCODED &&
         (segment.vector_type == HYB_VEC_FIXED_1B || segment.vector_type == HYB_VEC_FIXED_2B);
}

__device__ __forceinline__ uint4 load_raw_codes8(const DevSegment& segment, uint32_t row0) {
  if (segment.vector_type == HYB_VEC_FIXED_1B) {
    const uint2 v = ld_stream_v2(static_cast<const uint8_t*>(segment.av) + row0);
    return make_uint4(v.x, v.y, 0u, 0u);
  }
  return ld_stream_v4(static_cast<const uint16_t*>(segment.av) + row0);
}

__device__ __forceinline__ void unpack_raw_codes8(const uint4& raw, uint32_t vector_type, uint32_t (&codes)[8]) {
  if (vector_type == HYB_VEC_FIXED_1B) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      codes[j] = (raw.x >> (8 * j)) & 0xFFu;
      codes[4 + j] = (raw.y >> (8 * j)) & 0xFFu;
    }
  } else {
    codes[0] = raw.x & 0xFFFFu;
    codes[1] = raw.x >> 16;
    codes[2] = raw.y & 0xFFFFu;
    codes[3] = raw.y >> 16;
    codes[4] = raw.z & 0xFFFFu;
    codes[5] = raw.z >> 16;
    codes[6] = raw.w & 0xFFFFu;
    codes[7] = raw.w >> 16;
  }
}

// load_codes8, or the unpacking of an early load (by value: taking the address of a register array element would move it
// to local memory).
__device__ __forceinline__ void codes8(const DevSegment& segment, uint32_t row0, bool have_raw, uint4 raw,
                                       uint32_t (&codes)[8]) {
  if (have_raw) {
    unpack_raw_codes8(raw, segment.vector_type, codes);
  } else {
    load_codes8(segment.av, segment.vector_type, segment.bit_width, row0, segment.row_count, codes);
  }
}

// Single-entry decode (gather paths: position-filtered scans, join build side through a filter).
__device__ __forceinline__ uint32_t load_code1(const void* base, uint32_t vector_type, uint32_t bit_width, uint32_t row) {
  switch (vector_type) {
    case HYB_VEC_FIXED_1B:
      return __ldg(static_cast<const uint8_t*>(base) + row);
    case HYB_VEC_FIXED_2B:
      return __ldg(static_cast<const uint16_t*>(base) + row);
    case HYB_VEC_FIXED_4B:
      return __ldg(static_cast<const uint32_t*>(base) + row);
    default: {
      const auto* words = static_cast<const unsigned long long*>(base);
      const unsigned long long mask = (1ull << bit_width) - 1ull;
      const unsigned long long bit = static_cast<unsigned long long>(row) * bit_width;
      const unsigned long long word = bit >> 6;
      const uint32_t shift = static_cast<uint32_t>(bit & 63);
      unsigned long long value = __ldg(words + word) >> shift;
      if (shift + bit_width > 64) value |= __ldg(words + word + 1) << (64 - shift);
      return static_cast<uint32_t>(value & mask);
    }
  }
}

// 8 null flags (one byte per row) as a bit mask; 0 when the segment has no null vector.
__device__ __forceinline__ uint32_t load_nulls8(const uint8_t* nulls, uint32_t row0) {
  if (!nulls) return 0;
  const uint2 v = ld_stream_v2(nulls + row0);
  uint32_t mask = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mask |= ((v.x >> (8 * j)) & 0xFFu) ? (1u << j) : 0u;
    mask |= ((v.y >> (8 * j)) & 0xFFu) ? (1u << (4 + j)) : 0u;
  }
  return mask;
}

// ---------------------------------------------------------------------------------------------------------------------
// Warp primitives
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t value, uint32_t lane) {
#pragma unroll
  for (int delta = 1; delta < 32; delta <<= 1) {
    const uint32_t other = __shfl_up_sync(kFullMask, value, delta);
    if (lane >= static_cast<uint32_t>(delta)) value += other;
  }
  return value;
}

__device__ __forceinline__ unsigned long long warp_reduce_sum_u64(unsigned long long value) {
#pragma unroll
  for (int delta = 16; delta > 0; delta >>= 1) value += __shfl_xor_sync(kFullMask, value, delta);
  return value;
}

// ---------------------------------------------------------------------------------------------------------------------
// Decoupled look-back over tile status words (single-pass ordered compaction). Status word: bits 63..62 = state
// (0 invalid, 1 tile aggregate, 2 inclusive prefix), bits 61..0 = value. Must be called by all 32 lanes of one warp.
// Tiles are claimed through an atomic ticket, so every predecessor is already running: the spin cannot deadlock.
// Returns the exclusive prefix of `tile` and publishes its inclusive prefix.
// ---------------------------------------------------------------------------------------------------------------------
constexpr unsigned long long kStatusAggregate = 1ull << 62;
constexpr unsigned long long kStatusPrefix = 2ull << 62;
constexpr unsigned long long kStatusValueMask = (1ull << 62) - 1ull;

__device__ __forceinline__ unsigned long long lookback_exclusive_prefix(unsigned long long* status, uint32_t tile,
                                                                        unsigned long long tile_total, uint32_t lane) {
  if (tile == 0) {
    if (lane == 0) st_volatile_u64(status, kStatusPrefix | tile_total);
    return 0;
  }
  if (lane == 0) st_volatile_u64(status + tile, kStatusAggregate | tile_total);
  unsigned long long exclusive = 0;
  long long look = static_cast<long long>(tile) - 1;
  while (true) {
    const long long index = look - static_cast<long long>(lane);
    unsigned long long word;
    do {
      word = index >= 0 ? ld_volatile_u64(status + index) : kStatusPrefix;
    } while (__any_sync(kFullMask, (word >> 62) == 0));
    const uint32_t prefix_lanes = __ballot_sync(kFullMask, (word >> 62) == 2);
    if (prefix_lanes) {
      const uint32_t first = __ffs(prefix_lanes) - 1;
      exclusive += warp_reduce_sum_u64(lane <= first ? (word & kStatusValueMask) : 0ull);
      break;
    }
    exclusive += warp_reduce_sum_u64(word & kStatusValueMask);
    look -= 32;
  }
  if (lane == 0) st_volatile_u64(status + tile, kStatusPrefix | (exclusive + tile_total));
  return exclusive;
}

// upper_bound(starts, value) - 1 over a small device array: the chunk that owns tile `value`.
__device__ __forceinline__ uint32_t find_owner(const uint32_t* starts, uint32_t count, uint32_t value) {
  uint32_t lo = 0, hi = count;  // invariant: starts[lo] <= value < starts[hi]
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (__ldg(starts + mid) <= value) {
      lo = mid;
    } else {
      hi = mid;
    }
  }
  return lo;
}

}  // namespace hyb
