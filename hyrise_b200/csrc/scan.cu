// TableScan on the device: predicate evaluation on encoded segments + order-preserving compaction into a RowIDPosList.
//
// Replaces the hot loop AbstractTableScanImpl::_scan_with_iterators / _simd_scan_with_iterators
// (src/lib/operators/table_scan/abstract_table_scan_impl.hpp:56-242) and the per-chunk dispatch of
// ColumnVsValueTableScanImpl (column_vs_value_table_scan_impl.cpp:43-272), ColumnBetweenTableScanImpl
// (column_between_table_scan_impl.cpp:42-226) and ColumnIsNullTableScanImpl for Value / Dictionary / FrameOfReference
// segments. One launch scans every chunk of the column:
//
//   scan_prepare_kernel   one thread per chunk: turns the predicate into a per-chunk test. Dictionary segments get the
//                         value-ID range the reference derives from lower_bound/upper_bound (binary search on the
//                         device-resident dictionary); "no row can match" chunks are marked so the scan skips their
//                         bytes (the reference's early-outs, column_vs_value_table_scan_impl.cpp:228-272).
//   scan_kernel           persistent CTAs claim 4096-row tiles through an atomic ticket. Per tile: 128-bit streaming
//                         loads of value-IDs / values, in-register decode, 8 predicates per thread -> bit mask, warp
//                         prefix sums, one decoupled look-back per tile for the global output offset, matches staged in
//                         shared memory and written as coalesced 8-byte RowIDs. Output order == reference order
//                         (chunk by chunk, ascending ChunkOffset), in a single pass over the input.
//
// HBM traffic per launch = N * (bytes per row of the scanned column) + M * 8 (RowIDs) — the compulsory bytes.
#include <algorithm>
#include <cstdlib>
#include <cmath>
#include <limits>

#include "bulk_copy.cuh"
#include "device_utils.cuh"
#include "internal.hpp"
#include "predicate.cuh"

namespace hyb {

constexpr int kScanThreads = 512;  // 16 warps x 1024 rows: 16384-row tiles (see the look-back note at scan_kernel)
constexpr int kScanWarps = kScanThreads / 32;
constexpr int kScanIterations = 4;                       // 8 rows per thread per iteration
constexpr int kScanWarpRows = 32 * 8 * kScanIterations;  // rows owned by one warp in a tile (contiguous)
constexpr int kScanTileRows = kScanWarps * kScanWarpRows;  // 8192
constexpr int kFilteredTileRows = 4096;                  // position-filtered scan: inputs per tile

struct ScanPredicateDevice {
  int32_t condition;
  int32_t data_type;
  hyb_value lower;
  hyb_value upper;
  const uint32_t* value_id_bounds;  // device copy or nullptr
};

// ---------------------------------------------------------------------------------------------------------------------
// Per-chunk predicate preparation
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__device__ uint32_t device_lower_bound(const T* dictionary, uint32_t size, T value) {
  uint32_t lo = 0, hi = size;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (dictionary[mid] < value) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return lo;
}

template <typename T>
__device__ uint32_t device_upper_bound(const T* dictionary, uint32_t size, T value) {
  uint32_t lo = 0, hi = size;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (!(value < dictionary[mid])) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  return lo;
}

__device__ void dictionary_bounds(const DevSegment& segment, hyb_value value, uint32_t& lower, uint32_t& upper) {
  const uint32_t size = segment.dict_size;
  switch (segment.data_type) {
    case HYB_TYPE_INT32:
      lower = device_lower_bound(static_cast<const int32_t*>(segment.values), size, value.i32);
      upper = device_upper_bound(static_cast<const int32_t*>(segment.values), size, value.i32);
      break;
    case HYB_TYPE_INT64:
      lower = device_lower_bound(static_cast<const long long*>(segment.values), size, static_cast<long long>(value.i64));
      upper = device_upper_bound(static_cast<const long long*>(segment.values), size, static_cast<long long>(value.i64));
      break;
    case HYB_TYPE_FLOAT32:
      lower = device_lower_bound(static_cast<const float*>(segment.values), size, value.f32);
      upper = device_upper_bound(static_cast<const float*>(segment.values), size, value.f32);
      break;
    default:
      lower = device_lower_bound(static_cast<const double*>(segment.values), size, value.f64);
      upper = device_upper_bound(static_cast<const double*>(segment.values), size, value.f64);
      break;
  }
}

__device__ __forceinline__ bool is_between(int32_t condition) {
  return condition >= HYB_PRED_BETWEEN_INCLUSIVE && condition <= HYB_PRED_BETWEEN_EXCLUSIVE;
}
__device__ __forceinline__ bool lower_inclusive(int32_t condition) {
  return condition == HYB_PRED_BETWEEN_INCLUSIVE || condition == HYB_PRED_BETWEEN_UPPER_EXCLUSIVE;
}
__device__ __forceinline__ bool upper_inclusive(int32_t condition) {
  return condition == HYB_PRED_BETWEEN_INCLUSIVE || condition == HYB_PRED_BETWEEN_LOWER_EXCLUSIVE;
}

__device__ long long value_as_int(hyb_value value, int32_t data_type) {
  return data_type == HYB_TYPE_INT32 ? static_cast<long long>(value.i32) : static_cast<long long>(value.i64);
}
__device__ double value_as_float(hyb_value value, int32_t data_type) {
  return data_type == HYB_TYPE_FLOAT32 ? static_cast<double>(value.f32) : value.f64;
}

__global__ void scan_prepare_kernel(const DevSegment* __restrict__ segments, uint32_t chunk_count,
                                    ScanPredicateDevice predicate, ChunkTest* __restrict__ tests) {
  const uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
  if (chunk >= chunk_count) return;
  const DevSegment segment = segments[chunk];
  ChunkTest test{};
  test.mode = kTestNone;
  const int32_t condition = predicate.condition;

  if (segment.row_count == 0) {
    tests[chunk] = test;
    return;
  }

  if (segment.encoding == HYB_ENC_DICTIONARY) {
    // Value-ID range [lo, hi) following column_vs_value_table_scan_impl.cpp:96-110 and
    // column_between_table_scan_impl.cpp:112-125. INVALID_VALUE_ID ("past the end") is represented as dict_size.
    const uint32_t size = segment.dict_size;
    uint32_t lo = 0, hi = 0;
    bool negate = false;
    if (condition == HYB_PRED_IS_NULL) {
      lo = size;
      hi = size + 1;
    } else if (condition == HYB_PRED_IS_NOT_NULL) {
      lo = 0;
      hi = size;
    } else if (is_between(condition)) {
      uint32_t lower_lb, lower_ub, upper_lb, upper_ub;
      if (predicate.value_id_bounds) {
        lower_lb = min(predicate.value_id_bounds[4 * chunk + 0], size);
        lower_ub = min(predicate.value_id_bounds[4 * chunk + 1], size);
        upper_lb = min(predicate.value_id_bounds[4 * chunk + 2], size);
        upper_ub = min(predicate.value_id_bounds[4 * chunk + 3], size);
      } else {
        dictionary_bounds(segment, predicate.lower, lower_lb, lower_ub);
        dictionary_bounds(segment, predicate.upper, upper_lb, upper_ub);
      }
      lo = lower_inclusive(condition) ? lower_lb : lower_ub;
      hi = upper_inclusive(condition) ? upper_ub : upper_lb;
    } else {
      uint32_t lb, ub;
      if (predicate.value_id_bounds) {
        lb = min(predicate.value_id_bounds[2 * chunk + 0], size);
        ub = min(predicate.value_id_bounds[2 * chunk + 1], size);
      } else {
        dictionary_bounds(segment, predicate.lower, lb, ub);
      }
      switch (condition) {
        case HYB_PRED_EQUALS:
          lo = lb;
          hi = ub;
          break;
        case HYB_PRED_NOT_EQUALS:
          lo = lb;
          hi = ub;
          negate = true;
          break;
        case HYB_PRED_LESS_THAN:
          lo = 0;
          hi = lb;
          break;
        case HYB_PRED_LESS_THAN_EQUALS:
          lo = 0;
          hi = ub;
          break;
        case HYB_PRED_GREATER_THAN:
          lo = ub;
          hi = size;
          break;
        default:  // HYB_PRED_GREATER_THAN_EQUALS
          lo = lb;
          hi = size;
          break;
      }
    }
    if (negate && lo >= hi) {
      // value not in the dictionary: every non-NULL row matches
      negate = false;
      lo = 0;
      hi = size;
    }
    if (!negate && lo >= hi) {
      test.mode = kTestNone;
    } else {
      test.mode = kTestIdRange;
      test.negate = negate;
      test.id_lo = lo;
      test.id_span = hi - lo;
    }
    tests[chunk] = test;
    return;
  }

  // ValueSegment / FrameOfReferenceSegment: typed comparison (type_comparison.hpp:87-180).
  if (condition == HYB_PRED_IS_NULL || condition == HYB_PRED_IS_NOT_NULL) {
    if (!segment.nulls) {
      if (condition == HYB_PRED_IS_NULL) {
        test.mode = kTestNone;
      } else {
        test.mode = kTestNull;  // no null vector: every row matches
        test.want_null = 0;
      }
    } else {
      test.mode = kTestNull;
      test.want_null = condition == HYB_PRED_IS_NULL;
    }
    tests[chunk] = test;
    return;
  }

  const bool integral = segment.data_type == HYB_TYPE_INT32 || segment.data_type == HYB_TYPE_INT64;
  if (integral) {
    const long long type_min = segment.data_type == HYB_TYPE_INT32 ? INT_MIN : LLONG_MIN;
    const long long type_max = segment.data_type == HYB_TYPE_INT32 ? INT_MAX : LLONG_MAX;
    long long lo = type_min, hi = type_max;
    bool empty = false, negate = false;
    const long long a = value_as_int(predicate.lower, segment.data_type);
    const long long b = value_as_int(predicate.upper, segment.data_type);
    switch (condition) {
      case HYB_PRED_EQUALS:
        lo = hi = a;
        break;
      case HYB_PRED_NOT_EQUALS:
        lo = hi = a;
        negate = true;
        break;
      case HYB_PRED_LESS_THAN:
        if (a == type_min) empty = true;
        hi = a - 1;
        break;
      case HYB_PRED_LESS_THAN_EQUALS:
        hi = a;
        break;
      case HYB_PRED_GREATER_THAN:
        if (a == type_max) empty = true;
        lo = a + 1;
        break;
      case HYB_PRED_GREATER_THAN_EQUALS:
        lo = a;
        break;
      default: {  // BETWEEN (column_between_table_scan_impl.cpp:88-97: empty integer ranges produce no output)
        lo = a;
        hi = b;
        if (!lower_inclusive(condition)) {
          if (a == type_max) empty = true;
          lo = a + 1;
        }
        if (!upper_inclusive(condition)) {
          if (b == type_min) empty = true;
          hi = b - 1;
        }
        break;
      }
    }
    if (empty || lo > hi) {
      test.mode = kTestNone;
    } else {
      test.mode = kTestInt;
      test.negate = negate;
      test.int_lo = lo;
      test.int_hi = hi;
    }
  } else {
    const double a = value_as_float(predicate.lower, segment.data_type);
    const double b = value_as_float(predicate.upper, segment.data_type);
    test.mode = kTestFloat;
    test.float_lo = -INFINITY;
    test.float_hi = INFINITY;
    test.float_lo_inclusive = 1;
    test.float_hi_inclusive = 1;
    switch (condition) {
      case HYB_PRED_EQUALS:
        test.float_lo = test.float_hi = a;
        break;
      case HYB_PRED_NOT_EQUALS:
        test.float_lo = test.float_hi = a;
        test.negate = 1;
        break;
      case HYB_PRED_LESS_THAN:
        test.float_hi = a;
        test.float_hi_inclusive = 0;
        break;
      case HYB_PRED_LESS_THAN_EQUALS:
        test.float_hi = a;
        break;
      case HYB_PRED_GREATER_THAN:
        test.float_lo = a;
        test.float_lo_inclusive = 0;
        break;
      case HYB_PRED_GREATER_THAN_EQUALS:
        test.float_lo = a;
        break;
      default:
        test.float_lo = a;
        test.float_hi = b;
        test.float_lo_inclusive = lower_inclusive(condition);
        test.float_hi_inclusive = upper_inclusive(condition);
        break;
    }
  }
  tests[chunk] = test;
}

struct ScanParams {
  const DevSegment* segments;     // descriptors of the scanned column, one per chunk
  const ChunkTest* tests;         // per chunk
  const uint2* tile_map;          // per tile: {chunk, first row | last-tile-of-chunk << 31}
  uint32_t chunk_count;
  uint32_t tile_count;
  unsigned long long* tile_status;  // tile_count, zero-initialised
  uint32_t* ticket;               // zero-initialised
  hyb_row_id* out;                // capacity = table rows
  unsigned long long* chunk_end;  // [chunk] = inclusive prefix after the chunk's last tile; [chunk_count] = total
  uint32_t prefetch;              // 1: hint the next tile's codes into L2 during the write-out
};

// Per-tile schedule, all of it latency: (1) ticket + tile map entry — prefetched one tile ahead by thread 0;
// (2) column loads, 4 x 128 bit in flight per thread; (3) barrier A: warp totals; (4) warp 0: look-back for the global
// offset, other warps: stage matches in shared memory; (5) barrier B; (6) coalesced RowID stores. Two barriers per
// 8192-row tile; shared-memory reuse across iterations is ordered by those same two barriers (see comments inline).
__global__ void __launch_bounds__(kScanThreads) scan_kernel(const ScanParams params) {
  __shared__ uint16_t s_offsets[kScanTileRows];  // tile-relative offsets of the matches, in row order
  __shared__ uint32_t s_warp_totals[kScanWarps];
  __shared__ uint32_t s_next[2][3];              // double-buffered {tile, chunk, row0|last}
  __shared__ unsigned long long s_base;

  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    const uint32_t tile = atomicAdd(params.ticket, 1u);
    const uint2 info = tile < params.tile_count ? __ldg(params.tile_map + tile) : make_uint2(0, 0);
    s_next[0][0] = tile;
    s_next[0][1] = info.x;
    s_next[0][2] = info.y;
  }
  __syncthreads();

  for (uint32_t iteration = 0;; ++iteration) {
    const uint32_t slot = iteration & 1;
    const uint32_t tile = s_next[slot][0];
    if (tile >= params.tile_count) return;
    const uint32_t chunk = s_next[slot][1];
    const uint32_t tile_row0 = s_next[slot][2] & 0x7FFFFFFFu;
    const bool last_tile_of_chunk = (s_next[slot][2] >> 31) != 0;

    // Claim the next tile now; the answer is only needed after barrier B.
    uint32_t next_tile = 0;
    uint2 next_info = make_uint2(0, 0);
    if (threadIdx.x == 0) {
      next_tile = atomicAdd(params.ticket, 1u);
      if (next_tile < params.tile_count) next_info = __ldg(params.tile_map + next_tile);
    }

    const DevSegment segment = params.segments[chunk];
    const ChunkTest test = params.tests[chunk];

    // 1. predicate masks for this thread's kScanIterations x 8 rows (the warp owns kScanWarpRows contiguous rows)
    uint32_t masks[kScanIterations];
    unsigned long long packed_counts = 0;  // iteration i's count in bits [16i, 16i+16)
#pragma unroll
    for (int it = 0; it < kScanIterations; ++it) {
      const uint32_t row0 = tile_row0 + warp * kScanWarpRows + it * 256 + lane * 8;
      masks[it] = (test.mode != kTestNone && row0 < segment.row_count) ? evaluate8(segment, test, row0) : 0u;
      packed_counts |= static_cast<unsigned long long>(__popc(masks[it])) << (16 * it);
    }

    // 2. warp scan of all iterations at once (each field <= 256 fits 16 bits)
    unsigned long long inclusive = packed_counts;
#pragma unroll
    for (int delta = 1; delta < 32; delta <<= 1) {
      const unsigned long long other = __shfl_up_sync(kFullMask, inclusive, delta);
      if (lane >= static_cast<uint32_t>(delta)) inclusive += other;
    }
    const unsigned long long warp_sums = __shfl_sync(kFullMask, inclusive, 31);
    const unsigned long long exclusive = inclusive - packed_counts;
    uint32_t warp_total = 0;
#pragma unroll
    for (int it = 0; it < kScanIterations; ++it) warp_total += static_cast<uint32_t>((warp_sums >> (16 * it)) & 0xFFFFu);
    if (lane == 31) s_warp_totals[warp] = warp_total;
    __syncthreads();  // barrier A. (s_warp_totals of the previous tile was last read before its barrier B.)

    uint32_t warp_base = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kScanWarps; ++w) {
      const uint32_t total = s_warp_totals[w];
      if (w < static_cast<int>(warp)) warp_base += total;
      tile_total += total;
    }

    // 3. warp 0 resolves the global offset while the other warps stage their matches. s_offsets / s_base of the
    //    previous tile are free: every warp finished its write-out before arriving at barrier A.
    if (warp == 0) {
      const unsigned long long base = lookback_exclusive_prefix(params.tile_status, tile, tile_total, lane);
      if (lane == 0) {
        s_base = base;
        if (last_tile_of_chunk) params.chunk_end[chunk] = base + tile_total;
        if (tile + 1 == params.tile_count) params.chunk_end[params.chunk_count] = base + tile_total;
        s_next[slot ^ 1][0] = next_tile;
        s_next[slot ^ 1][1] = next_info.x;
        s_next[slot ^ 1][2] = next_info.y;
      }
    }
    {
      uint32_t position = warp_base;
#pragma unroll
      for (int it = 0; it < kScanIterations; ++it) {
        // rows of iteration `it` come after all rows of earlier iterations of this warp
        uint32_t at = position + static_cast<uint32_t>((exclusive >> (16 * it)) & 0xFFFFu);
        const uint32_t relative = warp * kScanWarpRows + it * 256 + lane * 8;
        const uint32_t mask = masks[it];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (mask & (1u << j)) s_offsets[at++] = static_cast<uint16_t>(relative + j);
        }
        position += static_cast<uint32_t>((warp_sums >> (16 * it)) & 0xFFFFu);
      }
    }
    __syncthreads();  // barrier B: staging, s_base and s_next[slot ^ 1] are visible

    // 4. coalesced RowID write-out, with the next tile's codes on their way into L2 meanwhile
    if (params.prefetch) {
      const uint32_t following = s_next[slot ^ 1][0];
      if (following < params.tile_count) {
        const DevSegment& next_segment = params.segments[s_next[slot ^ 1][1]];
        const uint32_t next_row = (s_next[slot ^ 1][2] & 0x7FFFFFFFu) + threadIdx.x * (kScanTileRows / kScanThreads);
        if (next_row < next_segment.row_count) prefetch_codes(next_segment, next_row);
      }
    }
    hyb_row_id* out = params.out + s_base;
    for (uint32_t i = threadIdx.x; i < tile_total; i += kScanThreads) {
      st_stream_v2(out + i, chunk, tile_row0 + s_offsets[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Two-pass scan (options.scan_two_pass): no CTA ever waits for another one.
//   scan_mask_kernel    reads the column once, evaluates the predicate and keeps one match BIT per row (1/16 of a 2-byte
//                       vector) plus the match count of every tile;
//   scan_bases_kernel   exclusive prefix of the tile counts (one CTA; the counts are a few thousand words);
//   scan_expand_kernel  turns the bits into RowIDs at the tile's known output offset: warp scan, staging in shared memory,
//                       coalesced stores — scan_kernel without its loads, its predicate and its look-back chain.
// Traffic: column + 2 x rows / 8 + RowIDs, i.e. ~5 % more than the single pass; the kernels are plain streaming kernels.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScanThreads) scan_mask_kernel(const ScanParams params, uint32_t* __restrict__ masks_out,
                                                                 uint32_t* __restrict__ tile_counts) {
  __shared__ uint32_t s_warp_totals[kScanWarps];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // Two tiles per iteration: their 2 x 4 vector loads per thread are independent and all in flight before the first
  // predicate is evaluated; the two match counts (<= 16 384 each) share one reduction as the halves of a 32-bit word.
  for (uint32_t tile = blockIdx.x; tile < params.tile_count; tile += 2 * gridDim.x) {
    uint32_t packed[2] = {0u, 0u}, counts = 0;
    const uint32_t tiles[2] = {tile, tile + gridDim.x};
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (tiles[t] >= params.tile_count) continue;  // uniform
      const uint2 info = __ldg(params.tile_map + tiles[t]);
      const uint32_t chunk = info.x, tile_row0 = info.y & 0x7FFFFFFFu;
      const DevSegment segment = params.segments[chunk];
      const ChunkTest test = params.tests[chunk];
#pragma unroll
      for (int it = 0; it < kScanIterations; ++it) {
        const uint32_t row0 = tile_row0 + warp * kScanWarpRows + it * 256 + lane * 8;
        const uint32_t mask = (test.mode != kTestNone && row0 < segment.row_count) ? evaluate8(segment, test, row0) : 0u;
        packed[t] |= mask << (8 * it);
        counts += static_cast<uint32_t>(__popc(mask)) << (16 * t);
      }
    }
    masks_out[static_cast<size_t>(tiles[0]) * kScanThreads + threadIdx.x] = packed[0];
    if (tiles[1] < params.tile_count) masks_out[static_cast<size_t>(tiles[1]) * kScanThreads + threadIdx.x] = packed[1];
    counts = __reduce_add_sync(kFullMask, counts);  // no carry between the halves: a warp holds <= 1024 rows per tile
    if (lane == 0) s_warp_totals[warp] = counts;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t total = 0;
#pragma unroll
      for (int w = 0; w < kScanWarps; ++w) total += s_warp_totals[w];
      tile_counts[tiles[0]] = total & 0xFFFFu;
      if (tiles[1] < params.tile_count) tile_counts[tiles[1]] = total >> 16;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(1024) scan_bases_kernel(const uint32_t* __restrict__ tile_counts, uint32_t tile_count,
                                                          unsigned long long* __restrict__ tile_bases) {
  __shared__ unsigned long long s_partial[1024];
  const uint32_t per_thread = (tile_count + 1023) / 1024;
  const uint32_t begin = min(tile_count, threadIdx.x * per_thread), end = min(tile_count, begin + per_thread);
  unsigned long long sum = 0;
  for (uint32_t tile = begin; tile < end; ++tile) sum += tile_counts[tile];
  s_partial[threadIdx.x] = sum;
  __syncthreads();
  for (uint32_t stride = 1; stride < 1024; stride <<= 1) {  // Hillis-Steele inclusive scan of the thread sums
    const unsigned long long other = threadIdx.x >= stride ? s_partial[threadIdx.x - stride] : 0ull;
    __syncthreads();
    s_partial[threadIdx.x] += other;
    __syncthreads();
  }
  unsigned long long running = s_partial[threadIdx.x] - sum;
  for (uint32_t tile = begin; tile < end; ++tile) {
    tile_bases[tile] = running;
    running += tile_counts[tile];
  }
  if (threadIdx.x == 1023) tile_bases[tile_count] = s_partial[1023];
}

__global__ void __launch_bounds__(kScanThreads) scan_expand_kernel(const ScanParams params, const uint32_t* __restrict__ masks_in,
                                                                   const unsigned long long* __restrict__ tile_bases) {
  __shared__ uint16_t s_offsets[kScanTileRows];
  __shared__ uint32_t s_warp_totals[kScanWarps];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (uint32_t tile = blockIdx.x; tile < params.tile_count; tile += gridDim.x) {
    const uint2 info = __ldg(params.tile_map + tile);
    const uint32_t chunk = info.x, tile_row0 = info.y & 0x7FFFFFFFu;
    const unsigned long long base = __ldg(tile_bases + tile);
    const uint32_t packed = ld_stream_u32(masks_in + static_cast<size_t>(tile) * kScanThreads + threadIdx.x);
    unsigned long long packed_counts = 0;
#pragma unroll
    for (int it = 0; it < kScanIterations; ++it) {
      packed_counts |= static_cast<unsigned long long>(__popc((packed >> (8 * it)) & 0xFFu)) << (16 * it);
    }
    unsigned long long inclusive = packed_counts;
#pragma unroll
    for (int delta = 1; delta < 32; delta <<= 1) {
      const unsigned long long other = __shfl_up_sync(kFullMask, inclusive, delta);
      if (lane >= static_cast<uint32_t>(delta)) inclusive += other;
    }
    const unsigned long long warp_sums = __shfl_sync(kFullMask, inclusive, 31);
    const unsigned long long exclusive = inclusive - packed_counts;
    uint32_t warp_total = 0;
#pragma unroll
    for (int it = 0; it < kScanIterations; ++it) warp_total += static_cast<uint32_t>((warp_sums >> (16 * it)) & 0xFFFFu);
    if (lane == 31) s_warp_totals[warp] = warp_total;
    __syncthreads();  // also: every warp has finished the previous tile's write-out
    uint32_t warp_base = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kScanWarps; ++w) {
      const uint32_t total = s_warp_totals[w];
      if (w < static_cast<int>(warp)) warp_base += total;
      tile_total += total;
    }
    if (threadIdx.x == 0) {
      if (info.y >> 31) params.chunk_end[chunk] = base + tile_total;
      if (tile + 1 == params.tile_count) params.chunk_end[params.chunk_count] = base + tile_total;
    }
    uint32_t position = warp_base;
#pragma unroll
    for (int it = 0; it < kScanIterations; ++it) {
      uint32_t at = position + static_cast<uint32_t>((exclusive >> (16 * it)) & 0xFFFFu);
      const uint32_t relative = warp * kScanWarpRows + it * 256 + lane * 8;
      const uint32_t mask = (packed >> (8 * it)) & 0xFFu;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (mask & (1u << j)) s_offsets[at++] = static_cast<uint16_t>(relative + j);
      }
      position += static_cast<uint32_t>((warp_sums >> (16 * it)) & 0xFFFFu);
    }
    __syncthreads();
    hyb_row_id* out = params.out + base;
    for (uint32_t i = threadIdx.x; i < tile_total; i += kScanThreads) {
      st_stream_v2(out + i, chunk, tile_row0 + s_offsets[i]);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// scan_bulk_kernel: the same single-pass ordered compaction, restructured around the Blackwell asynchronous machinery.
//
//   producer warp   one lane claims tiles through the atomic ticket (so every predecessor of a claimed tile is owned by a
//                   running CTA: the look-back cannot deadlock), reads the tile-map entry and the segment descriptor and
//                   issues ONE cp.async.bulk per tile (the TMA unit moves the tile's 8192 codes, <= 32 KB, global ->
//                   shared) that completes on the stage's `full` mbarrier; it runs kBulkStages - 1 tiles ahead of the
//                   consumers and is throttled by the `empty` mbarriers. No load latency is left on the consumers' path.
//   8 consumer warps  wait on `full`, evaluate their 1024 rows from shared memory (conflict-free 128-bit LDS), release
//                   the stage, scan their match counts and publish the warp total. There is no CTA-wide barrier: the
//                   warp that arrives LAST (shared-memory counter) owns the tile's decoupled look-back and publishes the
//                   global offset through the `ready` mbarrier, while the other warps compact their matches into
//                   warp-private staging; each warp then writes its own run of RowIDs (contiguous, 256 bytes per store).
//
// Eligible columns (checked on the host per call): every chunk streams a fixed-width vector of 1, 2 or 4 bytes per row
// (dictionary value-IDs, FrameOfReference offsets, int32 / float values) and has no NULL vector; anything else takes
// scan_kernel.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kBulkStages = 2;
constexpr int kBulkConsumerWarps = kScanWarps;                   // 8 x 1024 rows = one 8192-row tile
constexpr int kBulkThreads = (kBulkConsumerWarps + 1) * 32;      // + the producer warp
constexpr uint32_t kBulkEndOfTiles = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t evaluate8_staged(const DevSegment& segment, const ChunkTest& test, uint32_t row0,
                                                     const unsigned char* staged, uint32_t width) {
  const uint32_t rows = segment.row_count;
  const uint32_t valid = rows - row0 >= 8 ? 0xFFu : ((1u << (rows - row0)) - 1u);
  uint32_t codes[8];
  if (width == 2) {
    const uint4 v = *reinterpret_cast<const uint4*>(staged);
    codes[0] = v.x & 0xFFFFu;
    codes[1] = v.x >> 16;
    codes[2] = v.y & 0xFFFFu;
    codes[3] = v.y >> 16;
    codes[4] = v.z & 0xFFFFu;
    codes[5] = v.z >> 16;
    codes[6] = v.w & 0xFFFFu;
    codes[7] = v.w >> 16;
  } else if (width == 1) {
    const uint2 v = *reinterpret_cast<const uint2*>(staged);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      codes[j] = (v.x >> (8 * j)) & 0xFFu;
      codes[4 + j] = (v.y >> (8 * j)) & 0xFFu;
    }
  } else {
    const uint4 a = *reinterpret_cast<const uint4*>(staged);
    const uint4 b = *reinterpret_cast<const uint4*>(staged + 16);
    codes[0] = a.x;
    codes[1] = a.y;
    codes[2] = a.z;
    codes[3] = a.w;
    codes[4] = b.x;
    codes[5] = b.y;
    codes[6] = b.z;
    codes[7] = b.w;
  }
  uint32_t mask = 0;
  switch (test.mode) {
    case kTestIdRange:
      if (!test.negate) {
#pragma unroll
        for (int j = 0; j < 8; ++j) mask |= ((codes[j] - test.id_lo) < test.id_span) ? (1u << j) : 0u;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const bool inside = (codes[j] - test.id_lo) < test.id_span;
          mask |= (!inside && codes[j] < segment.dict_size) ? (1u << j) : 0u;
        }
      }
      break;
    case kTestInt: {
      uint32_t minimum = 0;
      if (segment.encoding == HYB_ENC_FRAME_OF_REFERENCE) {
        minimum = static_cast<uint32_t>(__ldg(static_cast<const int32_t*>(segment.values) + row0 / HYB_FOR_BLOCK_SIZE));
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const long long value = static_cast<int32_t>(minimum + codes[j]);
        const bool inside = value >= test.int_lo && value <= test.int_hi;
        mask |= (inside != static_cast<bool>(test.negate)) ? (1u << j) : 0u;
      }
      break;
    }
    case kTestFloat:
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const double value = __uint_as_float(codes[j]);
        const bool above = test.float_lo_inclusive ? value >= test.float_lo : value > test.float_lo;
        const bool below = test.float_hi_inclusive ? value <= test.float_hi : value < test.float_hi;
        mask |= ((above && below) != static_cast<bool>(test.negate)) ? (1u << j) : 0u;
      }
      break;
    case kTestNull:  // no NULL vector on this path: IS NOT NULL matches every row
      mask = test.want_null ? 0u : 0xFFu;
      break;
    default:
      break;
  }
  return mask & valid;
}

__global__ void __launch_bounds__(kBulkThreads, 2) scan_bulk_kernel(const ScanParams params, const uint32_t stage_bytes) {
  extern __shared__ __align__(128) unsigned char s_dynamic[];  // kBulkStages input stages, then the match staging
  __shared__ __align__(8) unsigned long long s_full[kBulkStages], s_empty[kBulkStages], s_ready[2];
  __shared__ uint4 s_info[kBulkStages];  // {tile, chunk, row0 | last-tile-of-chunk << 31, bytes per row}
  __shared__ __align__(16) DevSegment s_segment[kBulkStages];  // the tile's chunk: descriptor and predicate test
  __shared__ __align__(16) ChunkTest s_test[kBulkStages];
  __shared__ uint32_t s_totals[2][kBulkConsumerWarps];
  __shared__ uint32_t s_arrived[2];
  __shared__ unsigned long long s_base[2];

  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int stage = 0; stage < kBulkStages; ++stage) {
      mbarrier_init(&s_full[stage], 1);                    // the producer's arrive (+ the copy's bytes)
      mbarrier_init(&s_empty[stage], kBulkConsumerWarps);  // one arrival per consumer warp
    }
    mbarrier_init(&s_ready[0], 1);
    mbarrier_init(&s_ready[1], 1);
    s_arrived[0] = 0;
    s_arrived[1] = 0;
    mbarrier_init_fence();
  }
  __syncthreads();  // the only CTA-wide barrier of the kernel

  if (warp == kBulkConsumerWarps) {
    // ---- producer warp: lane 0 claims tiles and issues the copies; lanes 0-2 / 3-6 fetch the chunk's segment descriptor /
    //      predicate test (16 bytes each, in parallel) and hand them to the consumers through the stage header ------------
    for (uint32_t fill = 0;; ++fill) {
      const uint32_t stage = fill % kBulkStages;
      uint32_t tile = 0;
      if (lane == 0) {
        mbarrier_wait(&s_empty[stage], ((fill / kBulkStages) & 1u) ^ 1u);  // passes at once for the first kBulkStages fills
        tile = atomicAdd(params.ticket, 1u);
      }
      tile = __shfl_sync(kFullMask, tile, 0);
      if (tile >= params.tile_count) {
        if (lane == 0) {
          s_info[stage] = make_uint4(kBulkEndOfTiles, 0u, 0u, 0u);
          mbarrier_arrive(&s_full[stage]);
        }
        return;
      }
      const uint2 info = __ldg(params.tile_map + tile);
      static_assert(sizeof(DevSegment) == 48 && sizeof(ChunkTest) == 64, "copied as 3 + 4 uint4");
      if (lane < 3) {
        reinterpret_cast<uint4*>(&s_segment[stage])[lane] = __ldg(reinterpret_cast<const uint4*>(params.segments + info.x) + lane);
      } else if (lane < 7) {
        reinterpret_cast<uint4*>(&s_test[stage])[lane - 3] = __ldg(reinterpret_cast<const uint4*>(params.tests + info.x) + (lane - 3));
      }
      __syncwarp();
      if (lane == 0) {
        const DevSegment& segment = s_segment[stage];
        const uint32_t row0 = info.y & 0x7FFFFFFFu;
        const char* base;
        const uint32_t width = segment_stream(segment, base);
        s_info[stage] = make_uint4(tile, info.x, info.y, width);
        if (s_test[stage].mode != kTestNone) {
          const uint32_t rows = min(static_cast<uint32_t>(kScanTileRows), segment.row_count - row0);
          const uint32_t bytes = (rows * width + 15u) & ~15u;  // the readable tail pad covers the round-up
          mbarrier_arrive_expect_tx(&s_full[stage], bytes);
          bulk_copy_to_shared(s_dynamic + size_t{stage} * stage_bytes, base + size_t{row0} * width, bytes, &s_full[stage]);
        } else {
          mbarrier_arrive(&s_full[stage]);  // "no row can match": the chunk's bytes are never read
        }
      }
    }
  }

  // ---- consumers ----------------------------------------------------------------------------------------------------
  uint16_t* staging = reinterpret_cast<uint16_t*>(s_dynamic + size_t{kBulkStages} * stage_bytes) + warp * kScanWarpRows;
  for (uint32_t iteration = 0;; ++iteration) {
    const uint32_t stage = iteration % kBulkStages;
    const uint32_t slot = iteration & 1u;
    mbarrier_wait(&s_full[stage], (iteration / kBulkStages) & 1u);
    const uint4 info = s_info[stage];
    if (info.x == kBulkEndOfTiles) return;
    const uint32_t tile = info.x, chunk = info.y, width = info.w;
    const uint32_t tile_row0 = info.z & 0x7FFFFFFFu;
    const bool last_tile_of_chunk = (info.z >> 31) != 0;
    const DevSegment& segment = s_segment[stage];  // shared memory: no global access on the consumers' path
    const ChunkTest& test = s_test[stage];
    const unsigned char* staged = s_dynamic + size_t{stage} * stage_bytes;

    // 1. predicate masks of this thread's kScanIterations x 8 rows, read from the staged tile
    uint32_t masks[kScanIterations];
    unsigned long long packed_counts = 0;  // iteration i's count in bits [16i, 16i+16)
#pragma unroll
    for (int it = 0; it < kScanIterations; ++it) {
      const uint32_t relative = warp * kScanWarpRows + it * 256 + lane * 8;
      const uint32_t row0 = tile_row0 + relative;
      masks[it] = (test.mode != kTestNone && row0 < segment.row_count)
                      ? evaluate8_staged(segment, test, row0, staged + size_t{relative} * width, width)
                      : 0u;
      packed_counts |= static_cast<unsigned long long>(__popc(masks[it])) << (16 * it);
    }
    __syncwarp();
    if (lane == 0) mbarrier_arrive(&s_empty[stage]);  // the producer may refill the stage

    // 2. warp scan of all iterations at once
    unsigned long long inclusive = packed_counts;
#pragma unroll
    for (int delta = 1; delta < 32; delta <<= 1) {
      const unsigned long long other = __shfl_up_sync(kFullMask, inclusive, delta);
      if (lane >= static_cast<uint32_t>(delta)) inclusive += other;
    }
    const unsigned long long warp_sums = __shfl_sync(kFullMask, inclusive, 31);
    const unsigned long long exclusive = inclusive - packed_counts;
    uint32_t warp_total = 0;
#pragma unroll
    for (int it = 0; it < kScanIterations; ++it) warp_total += static_cast<uint32_t>((warp_sums >> (16 * it)) & 0xFFFFu);

    // 3. publish the warp total; the last warp to arrive resolves the tile's global offset (decoupled look-back)
    uint32_t arrived = 0;
    if (lane == 0) {
      s_totals[slot][warp] = warp_total;
      __threadfence_block();
      arrived = atomicAdd(&s_arrived[slot], 1u);
    }
    arrived = __shfl_sync(kFullMask, arrived, 0);
    if (arrived == kBulkConsumerWarps - 1) {
      __threadfence_block();
      uint32_t tile_total = lane < kBulkConsumerWarps ? *reinterpret_cast<volatile uint32_t*>(&s_totals[slot][lane]) : 0u;
#pragma unroll
      for (int delta = 16; delta > 0; delta >>= 1) tile_total += __shfl_xor_sync(kFullMask, tile_total, delta);
      const unsigned long long base = lookback_exclusive_prefix(params.tile_status, tile, tile_total, lane);
      if (lane == 0) {
        s_base[slot] = base;
        if (last_tile_of_chunk) params.chunk_end[chunk] = base + tile_total;
        if (tile + 1 == params.tile_count) params.chunk_end[params.chunk_count] = base + tile_total;
        s_arrived[slot] = 0;  // next use: iteration + 2, after every warp has passed this tile's `ready`
        mbarrier_arrive(&s_ready[slot]);
      }
    }

    // 4. compact this warp's matches (tile-relative offsets, row order) into its private staging
    {
      uint32_t position = 0;
#pragma unroll
      for (int it = 0; it < kScanIterations; ++it) {
        uint32_t at = position + static_cast<uint32_t>((exclusive >> (16 * it)) & 0xFFFFu);
        const uint32_t relative = warp * kScanWarpRows + it * 256 + lane * 8;
        const uint32_t mask = masks[it];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (mask & (1u << j)) staging[at++] = static_cast<uint16_t>(relative + j);
        }
        position += static_cast<uint32_t>((warp_sums >> (16 * it)) & 0xFFFFu);
      }
    }
    __syncwarp();

    // 5. this warp's run of RowIDs: contiguous in the output, 256 bytes per warp store
    mbarrier_wait(&s_ready[slot], (iteration >> 1) & 1u);
    unsigned long long at = s_base[slot];
#pragma unroll
    for (int w = 0; w < kBulkConsumerWarps; ++w) at += w < static_cast<int>(warp) ? s_totals[slot][w] : 0u;
    hyb_row_id* out = params.out + at;
    for (uint32_t i = lane; i < warp_total; i += 32) st_stream_v2(out + i, chunk, tile_row0 + staging[i]);
    __syncwarp();  // staging is rewritten by the next tile
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Position-filtered scan (reference-table input whose pos lists reference a single chunk each — what a previous
// TableScan on the same table produced; abstract_dereferenced_column_table_scan_impl.cpp:38-46 +
// table_scan.cpp:150-197): gather the referenced rows, test, and emit the referenced RowIDs in input order.
// ---------------------------------------------------------------------------------------------------------------------
struct FilteredScanParams {
  const DevSegment* segments;
  const ChunkTest* tests;
  const hyb_row_id* input;            // flat input pos list
  unsigned long long input_count;
  const unsigned long long* input_chunk_end;  // inclusive prefix per chunk of the input list (chunk_count entries)
  uint32_t chunk_count;
  uint32_t tile_count;                // ceil(input_count / kFilteredTileRows)
  unsigned long long* tile_status;
  uint32_t* ticket;
  hyb_row_id* out;
  unsigned long long* out_total;      // [0] = total matches
  uint32_t null_row_matches;          // 1: the predicate is IS NULL — a NULL_ROW_ID (outer join) is a NULL value and matches
};

__global__ void __launch_bounds__(kScanThreads) filtered_scan_kernel(const FilteredScanParams params) {
  __shared__ hyb_row_id s_rows[kFilteredTileRows];
  __shared__ uint32_t s_warp_totals[kScanWarps];
  __shared__ uint32_t s_tile;
  __shared__ unsigned long long s_base;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = threadIdx.x >> 5;
  constexpr int kPerThread = kFilteredTileRows / kScanThreads;  // 16 consecutive inputs per thread

  while (true) {
    if (threadIdx.x == 0) s_tile = atomicAdd(params.ticket, 1u);
    __syncthreads();
    const uint32_t tile = s_tile;
    if (tile >= params.tile_count) return;
    const unsigned long long first = static_cast<unsigned long long>(tile) * kFilteredTileRows +
                                     static_cast<unsigned long long>(threadIdx.x) * kPerThread;
    uint32_t mask = 0;
    hyb_row_id rows[kPerThread];
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      const unsigned long long index = first + j;
      if (index < params.input_count) {
        rows[j] = params.input[index];
        if (rows[j].chunk_id != HYB_INVALID_CHUNK_ID) {
          const DevSegment& segment = params.segments[rows[j].chunk_id];
          const ChunkTest& test = params.tests[rows[j].chunk_id];
          if (evaluate1(segment, test, rows[j].chunk_offset)) mask |= 1u << j;
        } else if (params.null_row_matches) {  // NULL_ROW_ID (outer join): the value is NULL; only IS NULL holds
          mask |= 1u << j;
        }
      }
    }
    const uint32_t count = __popc(mask);
    const uint32_t inclusive = warp_inclusive_scan(count, lane);
    if (lane == 31) s_warp_totals[warp] = inclusive;
    __syncthreads();
    uint32_t warp_base = 0, tile_total = 0;
#pragma unroll
    for (int w = 0; w < kScanWarps; ++w) {
      const uint32_t total = s_warp_totals[w];
      if (w < static_cast<int>(warp)) warp_base += total;
      tile_total += total;
    }
    if (warp == 0) {
      const unsigned long long base = lookback_exclusive_prefix(params.tile_status, tile, tile_total, lane);
      if (lane == 0) {
        s_base = base;
        if (tile + 1 == params.tile_count) params.out_total[0] = base + tile_total;
      }
    }
    uint32_t position = warp_base + inclusive - count;
#pragma unroll
    for (int j = 0; j < kPerThread; ++j) {
      if (mask & (1u << j)) s_rows[position++] = rows[j];
    }
    __syncthreads();
    hyb_row_id* out = params.out + s_base;
    for (uint32_t i = threadIdx.x; i < tile_total; i += kScanThreads) {
      st_stream_v2(out + i, s_rows[i].chunk_id, s_rows[i].chunk_offset);
    }
  }
}

// Per-chunk boundaries of a filtered scan's output: the output is ordered by chunk (it is a subsequence of an input that
// is), so chunk c ends at upper_bound(out, c).
__global__ void pos_list_chunk_ends_kernel(const hyb_row_id* __restrict__ rows, const unsigned long long* total_ptr,
                                           uint32_t chunk_count, unsigned long long* __restrict__ chunk_end) {
  const uint32_t chunk = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long total = *total_ptr;
  if (chunk > chunk_count) return;
  if (chunk == chunk_count) {
    chunk_end[chunk] = total;
    return;
  }
  unsigned long long lo = 0, hi = total;
  while (lo < hi) {
    const unsigned long long mid = (lo + hi) >> 1;
    if (rows[mid].chunk_id <= chunk) {
      lo = mid + 1;
    } else {
      hi = mid;
    }
  }
  chunk_end[chunk] = lo;
}

static bool supported_condition(int32_t condition) {
  return (condition >= HYB_PRED_EQUALS && condition <= HYB_PRED_BETWEEN_EXCLUSIVE) || condition == HYB_PRED_IS_NULL ||
         condition == HYB_PRED_IS_NOT_NULL;
}

// Shared by hyb_table_scan and the fused predicates of hyb_aggregate_hash: builds the per-chunk tests on the device.
int prepare_chunk_tests(hyb_context* context, Table* table, const hyb_scan_predicate* predicate, ChunkTest** out_tests,
                        void** out_bounds_scratch) {
  *out_tests = nullptr;
  *out_bounds_scratch = nullptr;
  HYB_CHECK(predicate->column_id < table->column_count, HYB_ERR_INVALID, "predicate column out of range");
  HYB_CHECK(supported_condition(predicate->condition), HYB_ERR_UNSUPPORTED,
            "predicate condition " + std::to_string(predicate->condition) + " is not on the GPU path");
  const uint32_t chunk_count = table->chunk_count();
  const int32_t data_type = table->column_types[predicate->column_id];
  const bool needs_value =
      predicate->condition != HYB_PRED_IS_NULL && predicate->condition != HYB_PRED_IS_NOT_NULL;
  if (data_type == HYB_TYPE_STRING && needs_value) {
    HYB_CHECK(predicate->value_id_bounds, HYB_ERR_INVALID,
              "string dictionary scans need host-computed value_id_bounds (dictionary_segment.cpp:94-119)");
  }
  ScanPredicateDevice device_predicate{};
  device_predicate.condition = predicate->condition;
  device_predicate.data_type = data_type;
  device_predicate.lower = predicate->lower;
  device_predicate.upper = predicate->upper;
  if (predicate->value_id_bounds && needs_value && chunk_count) {
    const bool between = predicate->condition >= HYB_PRED_BETWEEN_INCLUSIVE &&
                         predicate->condition <= HYB_PRED_BETWEEN_EXCLUSIVE;
    const size_t bytes = sizeof(uint32_t) * size_t{chunk_count} * (between ? 4 : 2);
    HYB_TRY(device_alloc(context, bytes, out_bounds_scratch));
    HYB_CUDA(cudaMemcpyAsync(*out_bounds_scratch, predicate->value_id_bounds, bytes, cudaMemcpyHostToDevice,
                             context->stream));
    // value_id_bounds is borrowed pageable memory: the copy is staged before cudaMemcpyAsync returns.
    device_predicate.value_id_bounds = static_cast<const uint32_t*>(*out_bounds_scratch);
  }
  void* tests = nullptr;
  HYB_TRY(device_alloc(context, sizeof(ChunkTest) * std::max<uint32_t>(chunk_count, 1), &tests));
  if (chunk_count) {
    const DevSegment* column_segments = table->d_segments + size_t{predicate->column_id} * chunk_count;
    scan_prepare_kernel<<<(chunk_count + 127) / 128, 128, 0, context->stream>>>(
        column_segments, chunk_count, device_predicate, static_cast<ChunkTest*>(tests));
    HYB_CUDA(cudaGetLastError());
  }
  *out_tests = static_cast<ChunkTest*>(tests);
  return HYB_OK;
}

static uint64_t column_bytes_per_launch(const Table* table, uint32_t column_id) {
  uint64_t bytes = 0;
  for (uint32_t chunk = 0; chunk < table->chunk_count(); ++chunk) {
    const auto& segment = table->segments[size_t{chunk} * table->column_count + column_id];
    if (segment.encoding == HYB_ENC_UNENCODED) {
      bytes += data_type_size(segment.data_type) * segment.row_count;
    } else {
      bytes += vector_bytes(segment.vector_type, segment.bit_width, segment.row_count);
      if (segment.encoding == HYB_ENC_FRAME_OF_REFERENCE) {
        bytes += sizeof(int32_t) * ((segment.row_count + HYB_FOR_BLOCK_SIZE - 1) / HYB_FOR_BLOCK_SIZE);
      }
    }
    if (segment.nulls) bytes += segment.row_count;
  }
  return bytes;
}

}  // namespace hyb

using namespace hyb;

extern "C" {

int hyb_table_scan(hyb_context* context, hyb_table_t table_handle, const hyb_scan_predicate* predicate,
                   hyb_pos_list_t input_filter, hyb_pos_list_t* out_pos_list) {
  HYB_CHECK(context && predicate && out_pos_list, HYB_ERR_INVALID, "NULL argument");
  *out_pos_list = 0;
  DeviceGuard guard(context->device);
  std::lock_guard<std::mutex> lock(context->mutex);
  auto* table = find_table(context, table_handle);
  HYB_CHECK(table, HYB_ERR_NOT_FOUND, "unknown table handle");
  PosList* filter = nullptr;
  if (input_filter) {
    filter = find_pos_list(context, input_filter);
    HYB_CHECK(filter, HYB_ERR_NOT_FOUND, "unknown input_filter handle");
    HYB_CHECK(filter->table == table_handle, HYB_ERR_INVALID, "input_filter belongs to a different table");
  }
  HYB_TRY(sync_table_descriptors(context, table));

  timing_begin(context);
  ChunkTest* tests = nullptr;
  void* bounds_scratch = nullptr;
  HYB_TRY(prepare_chunk_tests(context, table, predicate, &tests, &bounds_scratch));

  const uint32_t chunk_count = table->chunk_count();
  auto result = std::make_unique<PosList>();
  result->table = table_handle;
  result->chunk_count = chunk_count;
  result->stream = context->stream;
  result->owner = context;
  const DevSegment* column_segments = table->d_segments + size_t{predicate->column_id} * chunk_count;
  uint32_t launches = 1;
  uint64_t input_rows = 0, input_bytes = 0;

  void* chunk_end = nullptr;
  HYB_TRY(device_alloc(context, sizeof(uint64_t) * (size_t{chunk_count} + 1), &chunk_end));
  result->d_chunk_end = static_cast<uint64_t*>(chunk_end);

  if (!filter) {
    const uint2* tile_map = nullptr;
    uint32_t tile_count = 0;
    HYB_TRY(get_tile_map(context, table, kScanTileRows, &tile_map, &tile_count));
    input_rows = table->row_count();
    input_bytes = column_bytes_per_launch(table, predicate->column_id);
    result->capacity = table->row_count();
    void* out = nullptr;
    HYB_TRY(device_alloc(context, sizeof(hyb_row_id) * result->capacity, &out));
    result->d_row_ids = static_cast<hyb_row_id*>(out);
    // chunk_end is pre-filled with "unset" so chunks without rows can be patched on the host.
    HYB_CUDA(cudaMemsetAsync(chunk_end, 0xFF, sizeof(uint64_t) * (size_t{chunk_count} + 1), context->stream));
    if (tile_count > 0) {
      void* status = nullptr;
      HYB_TRY(device_alloc(context, sizeof(uint64_t) * (size_t{tile_count} + 1), &status));
      HYB_CUDA(cudaMemsetAsync(status, 0, sizeof(uint64_t) * (size_t{tile_count} + 1), context->stream));
      ScanParams params{};
      params.segments = column_segments;
      params.tests = tests;
      params.tile_map = tile_map;
      params.chunk_count = chunk_count;
      params.tile_count = tile_count;
      params.tile_status = static_cast<unsigned long long*>(status);
      params.ticket = reinterpret_cast<uint32_t*>(static_cast<unsigned long long*>(status) + tile_count);
      params.out = result->d_row_ids;
      params.chunk_end = reinterpret_cast<unsigned long long*>(result->d_chunk_end);
      params.prefetch = 1;  // measured -8 % kernel time at SF 10
      // Bulk (TMA-staged) path: every chunk streams a fixed-width vector of <= 4 bytes per row without a NULL vector.
      uint32_t stream_width = 0;
      bool bulk = context->options.scan_bulk;
      for (uint32_t chunk = 0; chunk < chunk_count && bulk; ++chunk) {
        const DevSegment& segment = table->segments[size_t{chunk} * table->column_count + predicate->column_id];
        uint32_t width = 0;
        if (segment.encoding == HYB_ENC_UNENCODED) {
          width = static_cast<uint32_t>(data_type_size(segment.data_type));
        } else {
          width = segment.vector_type == HYB_VEC_FIXED_1B ? 1u : segment.vector_type == HYB_VEC_FIXED_2B ? 2u
                  : segment.vector_type == HYB_VEC_FIXED_4B ? 4u : 0u;
        }
        if (width == 0 || width > 4 || segment.nulls) bulk = false;
        stream_width = std::max(stream_width, width);
      }
      timing_kernel_begin(context);
      if (context->options.scan_two_pass) {
        uint32_t* masks = nullptr;
        uint32_t* tile_counts = nullptr;
        unsigned long long* tile_bases = nullptr;
        HYB_TRY(device_alloc(context, sizeof(uint32_t) * size_t{tile_count} * kScanThreads, reinterpret_cast<void**>(&masks)));
        HYB_TRY(device_alloc(context, sizeof(uint32_t) * size_t{tile_count}, reinterpret_cast<void**>(&tile_counts)));
        HYB_TRY(device_alloc(context, sizeof(unsigned long long) * (size_t{tile_count} + 1), reinterpret_cast<void**>(&tile_bases)));
        int mask_blocks = 1, expand_blocks = 1;
        HYB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&mask_blocks, scan_mask_kernel, kScanThreads, 0));
        HYB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&expand_blocks, scan_expand_kernel, kScanThreads, 0));
        scan_mask_kernel<<<std::min<uint32_t>(tile_count, context->sm_count * std::max(mask_blocks, 1)), kScanThreads, 0,
                           context->stream>>>(params, masks, tile_counts);
        scan_bases_kernel<<<1, 1024, 0, context->stream>>>(tile_counts, tile_count, tile_bases);
        scan_expand_kernel<<<std::min<uint32_t>(tile_count, context->sm_count * std::max(expand_blocks, 1)), kScanThreads, 0,
                             context->stream>>>(params, masks, tile_bases);
        device_free(context, masks);
        device_free(context, tile_counts);
        device_free(context, tile_bases);
        launches = 4;
      } else if (bulk) {
        const uint32_t stage_bytes = kScanTileRows * stream_width;
        const size_t dynamic_bytes = size_t{kBulkStages} * stage_bytes + sizeof(uint16_t) * kScanTileRows;
        HYB_CUDA(cudaFuncSetAttribute(scan_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(dynamic_bytes)));
        int blocks_per_sm = 0;
        HYB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, scan_bulk_kernel, kBulkThreads, dynamic_bytes));
        const uint32_t grid = std::min<uint32_t>(tile_count, context->sm_count * std::max(blocks_per_sm, 1));
        scan_bulk_kernel<<<grid, kBulkThreads, dynamic_bytes, context->stream>>>(params, stage_bytes);
      } else {
        int blocks_per_sm = 0;
        HYB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, scan_kernel, kScanThreads, 0));
        const uint32_t grid = std::min<uint32_t>(tile_count, context->sm_count * std::max(blocks_per_sm, 1));
        scan_kernel<<<grid, kScanThreads, 0, context->stream>>>(params);
      }
      timing_kernel_end(context);
      HYB_CUDA(cudaGetLastError());
      device_free(context, status);
      launches = std::max<uint32_t>(launches, 2);
    } else {
      timing_kernel_begin(context);
      timing_kernel_end(context);
    }
  } else {
    // Input = a previous scan's output on the same table.
    HYB_CUDA(cudaStreamSynchronize(context->stream));
    uint64_t input_count = 0;
    HYB_CUDA(cudaMemcpy(&input_count, filter->d_chunk_end + filter->chunk_count, sizeof(uint64_t),
                        cudaMemcpyDeviceToHost));
    if (input_count == ~uint64_t{0}) input_count = 0;
    input_rows = input_count;
    input_bytes = input_count * sizeof(hyb_row_id);
    result->capacity = input_count;
    void* out = nullptr;
    HYB_TRY(device_alloc(context, sizeof(hyb_row_id) * std::max<uint64_t>(input_count, 1), &out));
    result->d_row_ids = static_cast<hyb_row_id*>(out);
    const uint32_t tile_count = static_cast<uint32_t>((input_count + kFilteredTileRows - 1) / kFilteredTileRows);
    void* status = nullptr;
    HYB_TRY(device_alloc(context, sizeof(uint64_t) * (size_t{tile_count} + 2), &status));
    HYB_CUDA(cudaMemsetAsync(status, 0, sizeof(uint64_t) * (size_t{tile_count} + 2), context->stream));
    auto* total = static_cast<unsigned long long*>(status) + tile_count + 1;
    timing_kernel_begin(context);
    if (tile_count > 0) {
      FilteredScanParams params{};
      params.segments = column_segments;
      params.tests = tests;
      params.input = filter->d_row_ids;
      params.input_count = input_count;
      params.chunk_count = chunk_count;
      params.tile_count = tile_count;
      params.tile_status = static_cast<unsigned long long*>(status);
      params.ticket = reinterpret_cast<uint32_t*>(static_cast<unsigned long long*>(status) + tile_count);
      params.out = result->d_row_ids;
      params.out_total = total;
      params.null_row_matches = predicate->condition == HYB_PRED_IS_NULL;
      int blocks_per_sm = 0;
      HYB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, filtered_scan_kernel, kScanThreads, 0));
      const uint32_t grid = std::min<uint32_t>(tile_count, context->sm_count * std::max(blocks_per_sm, 1));
      filtered_scan_kernel<<<grid, kScanThreads, 0, context->stream>>>(params);
      HYB_CUDA(cudaGetLastError());
    }
    timing_kernel_end(context);
    if (filter->ascending) {
      pos_list_chunk_ends_kernel<<<(chunk_count + 1 + 127) / 128, 128, 0, context->stream>>>(
          result->d_row_ids, total, chunk_count, reinterpret_cast<unsigned long long*>(result->d_chunk_end));
    } else {
      // input in another order than the table's (a join's output, abstract_dereferenced_column_table_scan_impl.cpp:49-86):
      // the matches keep the input order and form ONE list
      result->chunk_count = 1;
      result->ascending = false;
      result->may_hold_null_rows = filter->may_hold_null_rows && predicate->condition == HYB_PRED_IS_NULL;
      pos_list_chunk_ends_kernel<<<1, 128, 0, context->stream>>>(result->d_row_ids, total, 0u,
                                                                 reinterpret_cast<unsigned long long*>(result->d_chunk_end));
      HYB_CUDA(cudaMemcpyAsync(result->d_chunk_end + 1, result->d_chunk_end, sizeof(uint64_t), cudaMemcpyDeviceToDevice, context->stream));
    }
    HYB_CUDA(cudaGetLastError());
    device_free(context, status);
    launches = 3;
  }
  device_free(context, tests);
  device_free(context, bounds_scratch);
  timing_end(context, launches, input_bytes, input_rows, 0);
  timing_output_count(context, result->d_chunk_end + chunk_count, sizeof(hyb_row_id));

  const auto handle = context->next_handle++;
  context->pos_lists.emplace(handle, std::move(result));
  *out_pos_list = handle;
  return HYB_OK;
}

static int ensure_pos_list_host(hyb_context* context, PosList* list) {
  if (list->host_valid) return HYB_OK;
  list->h_chunk_offsets.assign(size_t{list->chunk_count} + 1, 0);
  std::vector<uint64_t> ends(size_t{list->chunk_count} + 1);
  HYB_CUDA(cudaMemcpyAsync(ends.data(), list->d_chunk_end, sizeof(uint64_t) * ends.size(), cudaMemcpyDeviceToHost,
                           context->stream));
  HYB_CUDA(cudaStreamSynchronize(context->stream));
  uint64_t running = 0;
  for (uint32_t chunk = 0; chunk < list->chunk_count; ++chunk) {
    list->h_chunk_offsets[chunk] = running;
    if (ends[chunk] != ~uint64_t{0}) running = ends[chunk];
  }
  list->h_chunk_offsets[list->chunk_count] = running;
  list->host_valid = true;
  return HYB_OK;
}

int hyb_pos_list_info(hyb_context* context, hyb_pos_list_t handle, uint64_t* out_total, uint32_t* out_chunk_count) {
  HYB_CHECK(context, HYB_ERR_INVALID, "context is NULL");
  DeviceGuard guard(context->device);
  std::lock_guard<std::mutex> lock(context->mutex);
  auto* list = find_pos_list(context, handle);
  HYB_CHECK(list, HYB_ERR_NOT_FOUND, "unknown pos list handle");
  HYB_TRY(ensure_pos_list_host(context, list));
  if (out_total) *out_total = list->h_chunk_offsets.back();
  if (out_chunk_count) *out_chunk_count = list->chunk_count;
  return HYB_OK;
}

int hyb_pos_list_chunk_offsets(hyb_context* context, hyb_pos_list_t handle, uint64_t* out_chunk_offsets) {
  HYB_CHECK(context && out_chunk_offsets, HYB_ERR_INVALID, "NULL argument");
  DeviceGuard guard(context->device);
  std::lock_guard<std::mutex> lock(context->mutex);
  auto* list = find_pos_list(context, handle);
  HYB_CHECK(list, HYB_ERR_NOT_FOUND, "unknown pos list handle");
  HYB_TRY(ensure_pos_list_host(context, list));
  std::copy(list->h_chunk_offsets.begin(), list->h_chunk_offsets.end(), out_chunk_offsets);
  return HYB_OK;
}

int hyb_pos_list_copy(hyb_context* context, hyb_pos_list_t handle, uint64_t begin, uint64_t count,
                      hyb_row_id* out_row_ids) {
  HYB_CHECK(context && (out_row_ids || count == 0), HYB_ERR_INVALID, "NULL argument");
  DeviceGuard guard(context->device);
  std::lock_guard<std::mutex> lock(context->mutex);
  auto* list = find_pos_list(context, handle);
  HYB_CHECK(list, HYB_ERR_NOT_FOUND, "unknown pos list handle");
  HYB_TRY(ensure_pos_list_host(context, list));
  HYB_CHECK(begin + count <= list->h_chunk_offsets.back(), HYB_ERR_INVALID, "range exceeds the pos list");
  if (count) {
    HYB_CUDA(cudaMemcpyAsync(out_row_ids, list->d_row_ids + begin, sizeof(hyb_row_id) * count, cudaMemcpyDeviceToHost,
                             context->stream));
    HYB_CUDA(cudaStreamSynchronize(context->stream));
  }
  return HYB_OK;
}

int hyb_pos_list_device_ptr(hyb_context* context, hyb_pos_list_t handle, void** out_device_row_ids) {
  HYB_CHECK(context && out_device_row_ids, HYB_ERR_INVALID, "NULL argument");
  std::lock_guard<std::mutex> lock(context->mutex);
  auto* list = find_pos_list(context, handle);
  HYB_CHECK(list, HYB_ERR_NOT_FOUND, "unknown pos list handle");
  *out_device_row_ids = list->d_row_ids;
  return HYB_OK;
}

int hyb_pos_list_free(hyb_context* context, hyb_pos_list_t handle) {
  HYB_CHECK(context, HYB_ERR_INVALID, "context is NULL");
  DeviceGuard guard(context->device);
  std::lock_guard<std::mutex> lock(context->mutex);
  auto it = context->pos_lists.find(handle);
  HYB_CHECK(it != context->pos_lists.end(), HYB_ERR_NOT_FOUND, "unknown pos list handle");
  context->pos_lists.erase(it);
  return HYB_OK;
}

}  // extern "C"
