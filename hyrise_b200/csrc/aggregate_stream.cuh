// aggregate_stream_kernel — the low-cardinality AggregateHash path (Q1 / Q6 shapes) rebuilt around the TMA unit.
//
// Included by aggregate.cu (uses FastPlan, WorkType, apply_affine, multiply, add_where, key entries and the per-CTA
// partials layout of aggregate_fast_kernel; the host merge of the partials is shared). What changed and why:
//
//   * aggregate_fast_kernel was instruction-bound, not bandwidth-bound: 410 M warp instructions for 60 M rows (vector
//     unpacking of 8 rows per thread, per-iteration hit masks, an instruction footprint that overflowed the I-cache).
//     Here one row is handled per lane and step, with scalar code that is ~4x shorter per row.
//   * no load latency on the compute warps: a producer warp walks the (static, reproducible) tile schedule ahead of the
//     consumers and has the TMA unit copy — per tile, with one cp.async.bulk per array — the tile's slice of every
//     referenced column, the small dictionaries of the value columns and the key words of the group-by dictionaries
//     into a 3-stage shared-memory ring (mbarrier full/empty pairs). Segment descriptors and the chunk's predicate
//     tests are fetched by the producer's lanes in parallel and handed over through the stage header, so consumers
//     never wait on global memory except for dictionaries too large to stage (gathered through L1).
//   * there is no CTA-wide barrier in the row loop: consumer warps run tile after tile on their own; the value-ID
//     combination -> group table is private to a warp and rebuilt from the staged key words when the chunk changes.
//
// Eligibility is decided on the host per call (stream_plan_for): every referenced column streams a fixed-width vector
// of <= 4 bytes per row, carries no NULLs, group-by columns are dictionary segments with 1-byte value-IDs whose
// combinations fit kMaxCombos; everything else keeps aggregate_fast_kernel / aggregate_general_kernel.
#pragma once

namespace hyb {

constexpr int kStreamStages = 3;
constexpr int kStreamConsumerWarps = 8;
constexpr int kStreamConsumerThreads = kStreamConsumerWarps * 32;
constexpr int kStreamThreads = kStreamConsumerThreads + 32;  // + the producer warp
constexpr int kStreamRowsPerWarp = 256;
constexpr int kStreamTileRows = kStreamConsumerWarps * kStreamRowsPerWarp;  // 3072
constexpr int kStreamLaneRows = 4;                                          // consecutive rows per lane and step
constexpr int kStreamSteps = kStreamRowsPerWarp / (32 * kStreamLaneRows);   // 2
constexpr int kStreamMaxColumns = 12;   // distinct staged columns (predicates + group-by + values)
constexpr int kStreamUnitTiles = 4;
constexpr uint32_t kStreamEnd = 0xFFFFFFFFu;
enum : uint32_t { kValueBits = 0, kValueStagedDictionary = 1, kValueGlobalDictionary = 2 };

// Header of a stage: what the producer learned about the tile's chunk, handed to the consumers with the data.
struct StreamStageInfo {
  uint32_t tile;   // kStreamEnd: no more tiles
  uint32_t chunk;
  uint32_t rows;   // valid rows of the tile (0: a predicate rules the whole chunk out, nothing was copied)
  uint32_t row0;   // first row of the tile inside its chunk
  uint32_t first_position;  // table position of the tile's first row
  uint32_t regular;         // 1: every staged slice has the vector width the plan's launch constants name
  uint32_t column_width[kStreamMaxColumns];  // bytes per row of the staged columns in this chunk
  const int32_t* predicate_minima[HYB_MAX_FUSED_PREDICATES];  // FrameOfReference block minima
  ChunkTest tests[HYB_MAX_FUSED_PREDICATES];
  uint32_t group_dict_size[HYB_MAX_GROUPBY_COLUMNS];
  uint32_t group_entry_type[HYB_MAX_GROUPBY_COLUMNS];  // 0xFF = staged uint64 key words, else hyb_data_type of staged dictionary values
  const void* dictionary[kFastMaxColumns];  // value column: dictionary in global memory
};

struct StreamColumn {
  const DevSegment* segments;  // descriptors of the column, one per chunk
  uint32_t slot_offset;        // byte offset of the column's tile slice inside a stage
};

// Passed as a kernel parameter: every field the row loop reads sits in the constant bank (uniform registers, uniform
// branches). The launch constants describe the COMMON case (taken from the first chunk); a tile of a chunk that deviates —
// another vector width because its dictionary is smaller, a dictionary that does not fit the stage — is flagged by the
// producer and takes the same code with the values read from the stage header instead.
struct StreamPlan {
  FastPlan fast;
  uint32_t column_count;
  StreamColumn columns[kStreamMaxColumns];
  uint32_t column_width[kStreamMaxColumns];              // launch constant: the widest vector of the column over all chunks
  uint32_t predicate_offset[HYB_MAX_FUSED_PREDICATES];   // stage offsets of the roles' column slices
  uint32_t group_offset[HYB_MAX_GROUPBY_COLUMNS];
  uint32_t value_offset[kFastMaxColumns];
  uint32_t predicate_width[HYB_MAX_FUSED_PREDICATES];    // launch constants
  uint32_t predicate_mode[HYB_MAX_FUSED_PREDICATES];
  uint32_t predicate_encoding[HYB_MAX_FUSED_PREDICATES];
  uint32_t value_width[kFastMaxColumns];
  uint32_t value_kind[kFastMaxColumns];
  uint32_t dictionary_offset[kFastMaxColumns];           // stage offset of value column c's staged dictionary
  uint32_t group_words_offset[HYB_MAX_GROUPBY_COLUMNS];  // stage offset of group-by column q's per-entry key data
  uint32_t info_offset;
  uint32_t stage_bytes;
};

__device__ __forceinline__ void stream_consumer_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(kStreamConsumerThreads) : "memory");
}

// Entries local0 .. local0 + 3 (local0 % 4 == 0) of a staged vector of `width` bytes per entry: one 4 / 8 / 16-byte LDS.
__device__ __forceinline__ void stream_codes4(const unsigned char* slot, uint32_t width, uint32_t local0, uint32_t (&codes)[4]) {
  if (width == 1) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(slot + local0);
    codes[0] = v & 0xFFu;
    codes[1] = (v >> 8) & 0xFFu;
    codes[2] = (v >> 16) & 0xFFu;
    codes[3] = v >> 24;
  } else if (width == 2) {
    const uint2 v = *reinterpret_cast<const uint2*>(slot + size_t{local0} * 2);
    codes[0] = v.x & 0xFFFFu;
    codes[1] = v.x >> 16;
    codes[2] = v.y & 0xFFFFu;
    codes[3] = v.y >> 16;
  } else {
    const uint4 v = *reinterpret_cast<const uint4*>(slot + size_t{local0} * 4);
    codes[0] = v.x;
    codes[1] = v.y;
    codes[2] = v.z;
    codes[3] = v.w;
  }
}

// One row against one predicate; `code` is the row's entry of the streamed vector. No NULLs on this path.
__device__ __forceinline__ bool stream_test(uint32_t mode, const ChunkTest& test, uint32_t encoding, const int32_t* minima,
                                            uint32_t code, uint32_t row) {
  switch (mode) {
    case kTestIdRange: {
      const bool inside = (code - test.id_lo) < test.id_span;
      return inside != static_cast<bool>(test.negate);
    }
    case kTestInt: {
      uint32_t minimum = 0;
      if (encoding == HYB_ENC_FRAME_OF_REFERENCE) minimum = static_cast<uint32_t>(__ldg(minima + row / HYB_FOR_BLOCK_SIZE));
      const long long value = static_cast<int32_t>(minimum + code);
      const bool inside = value >= test.int_lo && value <= test.int_hi;
      return inside != static_cast<bool>(test.negate);
    }
    case kTestFloat: {
      const double value = __uint_as_float(code);
      const bool above = test.float_lo_inclusive ? value >= test.float_lo : value > test.float_lo;
      const bool below = test.float_hi_inclusive ? value <= test.float_hi : value < test.float_hi;
      return (above && below) != static_cast<bool>(test.negate);
    }
    case kTestNull:
      return !test.want_null;
    default:
      return false;
  }
}

// AggregateKeyEntry of dictionary entry `value_id` of a group-by column from the staged per-entry data (see
// key_entry_from_code for the scheme).
__device__ __forceinline__ unsigned long long stream_key_entry(const unsigned char* words, uint32_t entry_type, uint32_t value_id) {
  switch (entry_type) {
    case HYB_TYPE_INT32:
      return static_cast<unsigned long long>(static_cast<long long>(reinterpret_cast<const int32_t*>(words)[value_id]) + 2147483648ll) + 1ull;
    case HYB_TYPE_INT64:
      return reinterpret_cast<const unsigned long long*>(words)[value_id];
    case HYB_TYPE_FLOAT32: {
      const float v = reinterpret_cast<const float*>(words)[value_id];
      return __float_as_uint(v == 0.0f ? 0.0f : v);
    }
    case HYB_TYPE_FLOAT64: {
      const double v = reinterpret_cast<const double*>(words)[value_id];
      return static_cast<unsigned long long>(__double_as_longlong(v == 0.0 ? 0.0 : v));
    }
    default:
      return reinterpret_cast<const unsigned long long*>(words)[value_id];
  }
}

// A tile whose chunk stores a column in a narrower vector than the launch constant (a dictionary that happens to be small,
// typically in the table's last chunk): all consumer warps widen the staged slice in place — read into registers,
// barrier, write — so that the row loop exists in one variant only. Rare; CTA-uniform (every consumer thread calls it).
__device__ __noinline__ void stream_widen_slice(unsigned char* slot, uint32_t from, uint32_t to, uint32_t rows) {
  constexpr int kPerThread = kStreamTileRows / kStreamConsumerThreads;  // 8
  uint32_t held[kPerThread];
#pragma unroll
  for (int i = 0; i < kPerThread; ++i) {
    const uint32_t row = threadIdx.x + i * kStreamConsumerThreads;
    held[i] = 0;
    if (row < rows) held[i] = from == 1 ? slot[row] : reinterpret_cast<const uint16_t*>(slot)[row];
  }
  stream_consumer_barrier();
#pragma unroll
  for (int i = 0; i < kPerThread; ++i) {
    const uint32_t row = threadIdx.x + i * kStreamConsumerThreads;
    if (row < rows) {
      if (to == 2) {
        reinterpret_cast<uint16_t*>(slot)[row] = static_cast<uint16_t>(held[i]);
      } else {
        reinterpret_cast<uint32_t*>(slot)[row] = held[i];
      }
    }
  }
  stream_consumer_barrier();
}

__device__ __forceinline__ void stream_widen_tile(const StreamPlan& plan, const StreamStageInfo* info, unsigned char* stage_base) {
  // slices shared by several roles are widened once: the header records the width per staged column
  for (uint32_t column = 0; column < plan.column_count; ++column) {
    const uint32_t from = info->column_width[column], to = plan.column_width[column];
    if (from < to) stream_widen_slice(stage_base + plan.columns[column].slot_offset, from, to, info->rows);
  }
}

// Per-thread aggregation state of a consumer (registers; static indexes only).
template <int W, int G, int C>
struct StreamState {
  typename WorkType<W>::Accumulator raw_sum[G][C], product_sum[G][C];
  uint32_t rows_seen[G];
  uint32_t first_position[G];
  uint32_t packed_rows;  // byte g = rows of group g since the last flush (< 256 by construction)
  uint32_t seen_groups;  // bit g: this thread has recorded first_position[g]
};

// The rows of one warp in one tile. Vector widths, value kinds and test modes are the plan's launch constants (tiles of a
// chunk with narrower vectors have been widened in place by stream_widen_tile).
template <int W, int G, int C>
__device__ __forceinline__ void stream_warp_rows(const StreamPlan& plan, const StreamStageInfo* info, const unsigned char* stage_base,
                                                 uint32_t warp, uint32_t lane, uint8_t* my_combos, unsigned long long* s_hash,
                                                 unsigned long long (*s_keys)[kMaxKeyWords],
                                                 const uint32_t (&combo_stride)[G == 1 ? 1 : HYB_MAX_GROUPBY_COLUMNS],
                                                 const typename WorkType<W>::Value (&affine_a)[C],
                                                 const typename WorkType<W>::Value (&affine_b)[C], StreamState<W, G, C>& state) {
  using Value = typename WorkType<W>::Value;
  using Accumulator = typename WorkType<W>::Accumulator;
  const FastPlan& fast = plan.fast;
  const uint32_t rows = info->rows;
  const uint32_t groupby_count = fast.groupby_count;
  const uint32_t predicate_count = fast.predicate_count;
  const uint32_t need_raw_mask = fast.need_raw_mask, need_product_mask = fast.need_product_mask;
#pragma unroll 1
  for (int step = 0; step < kStreamSteps; ++step) {
    const uint32_t local0 = warp * kStreamRowsPerWarp + step * (32 * kStreamLaneRows) + lane * kStreamLaneRows;
    if (warp * kStreamRowsPerWarp + step * (32 * kStreamLaneRows) >= rows) break;  // uniform
    const uint32_t valid = local0 >= rows ? 0u : (rows - local0 >= kStreamLaneRows ? 0xFu : ((1u << (rows - local0)) - 1u));
    uint32_t pass = valid;
    for (uint32_t p = 0; p < predicate_count; ++p) {
      const uint32_t width = plan.predicate_width[p];
      const uint32_t mode = plan.predicate_mode[p];
      uint32_t codes[kStreamLaneRows];
      stream_codes4(stage_base + plan.predicate_offset[p], width, local0, codes);
      uint32_t matches = 0;
      if (mode == kTestIdRange) {
        // the common case (dictionary value-ID range): three broadcast reads, two instructions per row
        const uint32_t id_lo = info->tests[p].id_lo, id_span = info->tests[p].id_span;
        const uint32_t flip = info->tests[p].negate ? 0xFu : 0u;
#pragma unroll
        for (int j = 0; j < kStreamLaneRows; ++j) matches |= ((codes[j] - id_lo) < id_span) ? (1u << j) : 0u;
        matches ^= flip;
      } else {
        const uint32_t encoding = plan.predicate_encoding[p];
#pragma unroll
        for (int j = 0; j < kStreamLaneRows; ++j) {
          matches |= stream_test(mode, info->tests[p], encoding, info->predicate_minima[p], codes[j], info->row0 + local0 + j) ? (1u << j) : 0u;
        }
      }
      pass &= matches;
    }
    if (!__any_sync(kFullMask, pass != 0)) continue;

    // ---- values of all columns first: every dictionary lookup of the step (shared-memory reads and, for dictionaries too
    //      large to stage, gathers through L1/L2) is in flight while the rows' groups are resolved and counted ----------
    Value values[C][kStreamLaneRows];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (fast.value_segments[c] == nullptr) continue;
      const uint32_t width = plan.value_width[c];
      const uint32_t kind = plan.value_kind[c];
      uint32_t codes[kStreamLaneRows];
      stream_codes4(stage_base + plan.value_offset[c], width, local0, codes);
      if (kind == kValueStagedDictionary) {
        // 1-byte codes cannot leave the 256-entry staged dictionary, whatever stale bytes sit past the tile's end
        const Value* dictionary = reinterpret_cast<const Value*>(stage_base + plan.dictionary_offset[c]);
        if (width == 1) {
#pragma unroll
          for (int j = 0; j < kStreamLaneRows; ++j) values[c][j] = dictionary[codes[j]];
        } else {
#pragma unroll
          for (int j = 0; j < kStreamLaneRows; ++j) values[c][j] = dictionary[((valid >> j) & 1u) ? codes[j] : 0u];
        }
      } else if (kind == kValueGlobalDictionary) {
        const void* dictionary = info->dictionary[c];
#pragma unroll
        for (int j = 0; j < kStreamLaneRows; ++j) values[c][j] = ((pass >> j) & 1u) ? typed_load<W>(dictionary, 0, codes[j]) : Value{};
      } else {
#pragma unroll
        for (int j = 0; j < kStreamLaneRows; ++j) {
          if constexpr (W == 0) {
            values[c][j] = __uint_as_float(codes[j]);
          } else {
            values[c][j] = Value{};  // unencoded 8-byte values are not streamed (host eligibility)
          }
        }
      }
    }

    // ---- group of every row (G == 1: the single group; else through the warp's combination table) -------------------
    uint32_t group[kStreamLaneRows];
    if constexpr (G == 1) {
#pragma unroll
      for (int j = 0; j < kStreamLaneRows; ++j) group[j] = ((pass >> j) & 1u) ? 0u : static_cast<uint32_t>(G);
    } else {
      uint32_t combination[kStreamLaneRows] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int q = 0; q < HYB_MAX_GROUPBY_COLUMNS; ++q) {
        if (static_cast<uint32_t>(q) < groupby_count) {
          uint32_t codes[kStreamLaneRows];
          stream_codes4(stage_base + plan.group_offset[q], 1u, local0, codes);
#pragma unroll
          for (int j = 0; j < kStreamLaneRows; ++j) combination[j] += codes[j] * combo_stride[q];
        }
      }
      bool unresolved = false;
#pragma unroll
      for (int j = 0; j < kStreamLaneRows; ++j) {
        const bool row_passes = (pass >> j) & 1u;
        combination[j] &= kMaxCombos - 1;  // identity for real rows; rows past the tile's end carry stale codes
        const uint32_t known = my_combos[combination[j]];
        group[j] = row_passes ? known : static_cast<uint32_t>(G);
        unresolved = unresolved || (row_passes && known >= kComboOverflow);
      }
      if (__any_sync(kFullMask, unresolved)) {
        // first sighting of a combination in this warp and chunk (rare): resolve / insert into the CTA's group table
#pragma unroll 1
        for (int j = 0; j < kStreamLaneRows; ++j) {
          uint32_t mine = j == 0 ? group[0] : j == 1 ? group[1] : j == 2 ? group[2] : group[3];
          const uint32_t my_combination = j == 0 ? combination[0] : j == 1 ? combination[1] : j == 2 ? combination[2] : combination[3];
          if (mine == kComboUnresolved) {
            unsigned long long entries[kMaxKeyWords];
            unsigned long long hash = 0x9E3779B97F4A7C15ull;
            uint32_t rest = my_combination;
            for (uint32_t q = 0; q < groupby_count; ++q) {
              const uint32_t size = info->group_dict_size[q];
              entries[q] = stream_key_entry(stage_base + plan.group_words_offset[q], info->group_entry_type[q], rest % size);
              rest /= size;
              hash = mix64(hash ^ entries[q]);
            }
            hash = mix64(hash) | 1ull;
            int32_t found = -1;
            for (int g = 0; g < G && found < 0; ++g) {
              unsigned long long current = *reinterpret_cast<volatile unsigned long long*>(&s_hash[g]);
              if (current == 0ull) {
                current = atomicCAS(&s_hash[g], 0ull, hash);
                if (current == 0ull) {
                  for (uint32_t q = 0; q < groupby_count; ++q) s_keys[g][q] = entries[q];
                  found = g;
                }
              }
              if (current == hash) found = g;
            }
            if (found < 0) {
              *fast.overflow = 1;  // more than G groups: the host falls back to a kernel with more group slots
              my_combos[my_combination] = kComboOverflow;
              mine = G;
            } else {
              my_combos[my_combination] = static_cast<uint8_t>(found);
              mine = static_cast<uint32_t>(found);
            }
          } else if (mine == kComboOverflow) {
            mine = G;
          }
          if (j == 0) group[0] = mine;
          if (j == 1) group[1] = mine;
          if (j == 2) group[2] = mine;
          if (j == 3) group[3] = mine;
        }
      }
    }

    // ---- row counts (byte-packed, flushed by the caller) and the first position of every group ----------------------
    {
      uint32_t step_groups = 0;
#pragma unroll
      for (int j = 0; j < kStreamLaneRows; ++j) {
        // group == G for rows that do not count: shifts of 32 and more yield 0 (shl.b32 clamps)
        uint32_t increment, bit;
        asm("shl.b32 %0, 1, %1;" : "=r"(increment) : "r"(group[j] * (G == 1 ? 32u : 8u)));
        asm("shl.b32 %0, 1, %1;" : "=r"(bit) : "r"(group[j] + (group[j] >= static_cast<uint32_t>(G) ? 32u : 0u)));
        state.packed_rows += increment;
        step_groups |= bit;
      }
      if (step_groups & ~state.seen_groups) {  // rare after the first tiles: a lane meets a group for the first time
#pragma unroll
        for (int j = kStreamLaneRows - 1; j >= 0; --j) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (group[j] == static_cast<uint32_t>(g) && !((state.seen_groups >> g) & 1u)) {
              state.first_position[g] = info->first_position + local0 + j;  // descending j: the smallest j wins
            }
          }
        }
        state.seen_groups |= step_groups;
      }
    }

    // ---- value columns: raw sums and the running product ------------------------------------------------------------
    Value product[kStreamLaneRows];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (fast.value_segments[c] == nullptr) continue;
      const bool in_chain = (need_product_mask >> c) != 0;  // some product at or after this column
      if (in_chain) {
#pragma unroll
        for (int j = 0; j < kStreamLaneRows; ++j) {
          const Value factor = apply_affine<W>(affine_a[c], affine_b[c], values[c][j]);
          product[j] = c == 0 ? factor : multiply<W>(product[j], factor);
        }
      }
      if ((need_raw_mask >> c) & 1u) {
#pragma unroll
        for (int j = 0; j < kStreamLaneRows; ++j) {
          const Accumulator widened = static_cast<Accumulator>(values[c][j]);
#pragma unroll
          for (int g = 0; g < G; ++g) add_where(state.raw_sum[g][c], widened, static_cast<int>(group[j]), g);
        }
      }
      if ((need_product_mask >> c) & 1u) {
#pragma unroll
        for (int j = 0; j < kStreamLaneRows; ++j) {
          const Accumulator widened = static_cast<Accumulator>(product[j]);
#pragma unroll
          for (int g = 0; g < G; ++g) add_where(state.product_sum[g][c], widened, static_cast<int>(group[j]), g);
        }
      }
    }
  }
}

template <int W, int G, int C>
__global__ void __launch_bounds__(kStreamThreads, 2) aggregate_stream_kernel(const __grid_constant__ StreamPlan plan) {
  using Value = typename WorkType<W>::Value;
  using Accumulator = typename WorkType<W>::Accumulator;
  static_assert(G == 1 || G == 4, "row counts are packed as one byte per group");
  const FastPlan& fast = plan.fast;  // kernel parameter: every plan field is a uniform constant-bank read

  extern __shared__ __align__(128) unsigned char s_stages[];  // kStreamStages x stage_bytes
  __shared__ __align__(8) unsigned long long s_full[kStreamStages], s_empty[kStreamStages];
  // CTA-wide group table (<= G distinct keys), filled on first sight
  __shared__ unsigned long long s_hash[G];
  __shared__ unsigned long long s_keys[G][kMaxKeyWords];
  __shared__ uint8_t s_combo_group[G == 1 ? 1 : kStreamConsumerWarps][G == 1 ? 4 : kMaxCombos];  // per warp
  __shared__ Accumulator s_reduce[kStreamConsumerWarps];
  __shared__ unsigned long long s_reduce_u64[kStreamConsumerWarps];

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int stage = 0; stage < kStreamStages; ++stage) {
      mbarrier_init(&s_full[stage], 1);
      mbarrier_init(&s_empty[stage], kStreamConsumerWarps);
    }
    mbarrier_init_fence();
  }
  if (threadIdx.x < G) {
    s_hash[threadIdx.x] = G == 1 ? 1ull : 0ull;  // without group-by columns the single group always exists
    for (int w = 0; w < kMaxKeyWords; ++w) s_keys[threadIdx.x][w] = 0;
  }
  __syncthreads();

  const uint32_t unit_count = (fast.tile_count + kStreamUnitTiles - 1) / kStreamUnitTiles;
  const uint32_t groupby_count = fast.groupby_count;
  const uint32_t predicate_count = fast.predicate_count;

  if (warp == kStreamConsumerWarps) {
    // ---- producer warp ----------------------------------------------------------------------------------------------
    // Lane roles: [0, 12) one staged column each; [12, 16) value column lane - 12 (header + small dictionary);
    // [16, 24) group-by column lane - 16 (header + key data of its dictionary); [24, 32) predicate lane - 24 (header).
    uint32_t fill = 0;
    for (uint32_t unit = blockIdx.x; unit < unit_count; unit += gridDim.x) {
      for (uint32_t tile = unit * kStreamUnitTiles; tile < min(fast.tile_count, (unit + 1) * kStreamUnitTiles); ++tile, ++fill) {
        const uint32_t stage = fill % kStreamStages;
        if (lane == 0) mbarrier_wait(&s_empty[stage], ((fill / kStreamStages) & 1u) ^ 1u);
        __syncwarp();
        unsigned char* stage_base = s_stages + size_t{stage} * plan.stage_bytes;
        auto* info = reinterpret_cast<StreamStageInfo*>(stage_base + plan.info_offset);
        const uint2 where = __ldg(fast.tile_map + tile);
        const uint32_t chunk = where.x;
        const uint32_t row0 = where.y & 0x7FFFFFFFu;
        const uint32_t chunk_rows = fast.size_segments[chunk].row_count;
        const uint32_t rows = min(static_cast<uint32_t>(kStreamTileRows), chunk_rows - row0);
        const void* source = nullptr;
        void* destination = nullptr;
        uint32_t bytes = 0;
        bool ruled_out = false, regular = true;
        if (lane < plan.column_count) {
          const DevSegment& segment = plan.columns[lane].segments[chunk];
          const char* base;
          const uint32_t width = segment_stream(segment, base);
          info->column_width[lane] = width;
          regular = width == plan.column_width[lane];  // narrower: the consumers widen the slice before the row loop
          source = base + size_t{row0} * width;
          destination = stage_base + plan.columns[lane].slot_offset;
          bytes = (rows * width + 15u) & ~15u;
        } else if (lane >= 12 && lane < 12 + C) {
          const int c = lane - 12;
          if (fast.value_segments[c] != nullptr) {
            const DevSegment& segment = fast.value_segments[c][chunk];
            info->dictionary[c] = segment.values;
            if (plan.value_kind[c] == kValueStagedDictionary) {  // the host checked: <= kStagedDictionary entries in every chunk
              source = segment.values;
              destination = stage_base + plan.dictionary_offset[c];
              bytes = (segment.dict_size * static_cast<uint32_t>(sizeof(Value)) + 15u) & ~15u;
            }
          }
        } else if (lane >= 16 && lane < 16 + groupby_count) {
          const int q = lane - 16;
          const DevSegment& segment = fast.group_segments[q][chunk];
          const uint32_t entry_bytes =
              (segment.dict_codes || segment.data_type == HYB_TYPE_INT64 || segment.data_type == HYB_TYPE_FLOAT64) ? 8u : 4u;
          info->group_dict_size[q] = segment.dict_size;
          info->group_entry_type[q] = segment.dict_codes ? 0xFFu : segment.data_type;
          source = segment.dict_codes ? static_cast<const void*>(segment.dict_codes) : segment.values;
          destination = stage_base + plan.group_words_offset[q];
          bytes = (segment.dict_size * entry_bytes + 15u) & ~15u;
        } else if (lane >= 24 && lane < 24 + predicate_count) {
          const int p = lane - 24;
          const ChunkTest test = fast.predicate_tests[p][chunk];
          info->tests[p] = test;
          info->predicate_minima[p] = static_cast<const int32_t*>(fast.predicate_segments[p][chunk].values);
          ruled_out = test.mode == kTestNone;
        }
        const bool skip = __any_sync(kFullMask, ruled_out);
        const bool all_regular = __all_sync(kFullMask, regular);
        if (skip) bytes = 0;
        uint32_t total = bytes;
#pragma unroll
        for (int delta = 16; delta > 0; delta >>= 1) total += __shfl_xor_sync(kFullMask, total, delta);
        if (lane == 0) {
          info->tile = tile;
          info->chunk = chunk;
          info->rows = skip ? 0u : rows;
          info->row0 = row0;
          info->first_position = static_cast<uint32_t>(__ldg(fast.chunk_row_start + chunk)) + row0;
          info->regular = all_regular ? 1u : 0u;
        }
        __syncwarp();  // header complete before the arrive publishes it
        if (lane == 0) {
          if (total) {
            mbarrier_arrive_expect_tx(&s_full[stage], total);
          } else {
            mbarrier_arrive(&s_full[stage]);
          }
        }
        __syncwarp();  // the expected byte count is registered before any copy can complete
        if (bytes) bulk_copy_to_shared(destination, source, bytes, &s_full[stage]);
      }
    }
    if (lane == 0) {
      const uint32_t stage = fill % kStreamStages;
      mbarrier_wait(&s_empty[stage], ((fill / kStreamStages) & 1u) ^ 1u);
      reinterpret_cast<StreamStageInfo*>(s_stages + size_t{stage} * plan.stage_bytes + plan.info_offset)->tile = kStreamEnd;
      mbarrier_arrive(&s_full[stage]);
    }
    return;
  }

  // ---- consumer warps ---------------------------------------------------------------------------------------------------
  StreamState<W, G, C> state;
  state.packed_rows = 0;
  state.seen_groups = 0;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    state.rows_seen[g] = 0;
    state.first_position[g] = 0xFFFFFFFFu;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      state.raw_sum[g][c] = Accumulator{};
      state.product_sum[g][c] = Accumulator{};
    }
  }
  const auto flush_row_counts = [&]() {
#pragma unroll
    for (int g = 0; g < G; ++g) state.rows_seen[g] += (state.packed_rows >> (8 * g)) & 0xFFu;
    state.packed_rows = 0;
  };
  Value affine_a[C], affine_b[C];
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const Value literal = static_cast<Value>(fast.literal[c]);
    const int32_t kind = fast.affine_kind[c];
    affine_a[c] = (kind == kLiteralMinusColumn || kind == kLiteralPlusColumn || kind == kColumnPlusLiteral) ? literal
                  : kind == kColumnMinusLiteral                                                             ? -literal
                                                                                                            : Value{};
    affine_b[c] = kind == kLiteralMinusColumn ? Value(-1) : Value(1);
  }
  uint32_t combo_chunk = 0xFFFFFFFFu;   // chunk the warp's combination table was built for
  uint32_t combo_stride[G == 1 ? 1 : HYB_MAX_GROUPBY_COLUMNS] = {};
  uint8_t* my_combos = s_combo_group[G == 1 ? 0 : warp];

  for (uint32_t iteration = 0;; ++iteration) {
    const uint32_t stage = iteration % kStreamStages;
    mbarrier_wait(&s_full[stage], (iteration / kStreamStages) & 1u);
    const unsigned char* stage_base = s_stages + size_t{stage} * plan.stage_bytes;
    const auto* info = reinterpret_cast<const StreamStageInfo*>(stage_base + plan.info_offset);
    if (info->tile == kStreamEnd) break;
    if (info->rows != 0) {
      if constexpr (G > 1) {
        if (info->chunk != combo_chunk) {
          // New chunk, new dictionaries: rebuild this warp's value-ID combination -> group table from the staged key words
          // (lookup only: a combination that never occurs must not claim a group slot).
          combo_chunk = info->chunk;
          uint32_t combos = 1;
#pragma unroll
          for (int q = 0; q < HYB_MAX_GROUPBY_COLUMNS; ++q) {
            combo_stride[q] = combos;
            if (static_cast<uint32_t>(q) < groupby_count) combos *= info->group_dict_size[q];
          }
          __syncwarp();
          for (uint32_t combination = lane; combination < combos; combination += 32) {
            unsigned long long hash = 0x9E3779B97F4A7C15ull;
            uint32_t rest = combination;
            for (uint32_t q = 0; q < groupby_count; ++q) {
              const uint32_t size = info->group_dict_size[q];
              const unsigned long long entry =
                  stream_key_entry(stage_base + plan.group_words_offset[q], info->group_entry_type[q], rest % size);
              rest /= size;
              hash = mix64(hash ^ entry);
            }
            hash = mix64(hash) | 1ull;
            uint8_t group = kComboUnresolved;
            for (int g = 0; g < G; ++g) {
              if (*reinterpret_cast<volatile unsigned long long*>(&s_hash[g]) == hash) group = static_cast<uint8_t>(g);
            }
            my_combos[combination] = group;
          }
          __syncwarp();
        }
      }
      if (!info->regular) stream_widen_tile(plan, info, s_stages + size_t{stage} * plan.stage_bytes);  // rare, CTA-uniform
      stream_warp_rows<W, G, C>(plan, info, stage_base, warp, lane, my_combos, s_hash, s_keys, combo_stride, affine_a, affine_b,
                                state);
      if ((iteration & 15u) == 15u) flush_row_counts();  // <= 8 rows per lane and tile: the bytes stay below 256
    }
    __syncwarp();
    if (lane == 0) mbarrier_arrive(&s_empty[stage]);
  }
  flush_row_counts();

  // ---- CTA reduction in a fixed order: lanes (butterfly), then warps; partials in the layout of aggregate_fast_kernel -------
  stream_consumer_barrier();
  const size_t cta = blockIdx.x;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const unsigned long long total_rows = warp_reduce_add(static_cast<unsigned long long>(state.rows_seen[g]));
    uint32_t low = state.first_position[g];
#pragma unroll
    for (int delta = 16; delta > 0; delta >>= 1) low = min(low, __shfl_xor_sync(kFullMask, low, delta));
    if (lane == 0) s_reduce_u64[warp] = total_rows;
    stream_consumer_barrier();
    if (threadIdx.x == 0) {
      unsigned long long sum = 0;
      for (int w = 0; w < kStreamConsumerWarps; ++w) sum += s_reduce_u64[w];
      fast.partial_rows[cta * G + g] = sum;
    }
    stream_consumer_barrier();
    if (lane == 0) s_reduce_u64[warp] = low;
    stream_consumer_barrier();
    if (threadIdx.x == 0) {
      unsigned long long value = 0xFFFFFFFFull;
      for (int w = 0; w < kStreamConsumerWarps; ++w) value = min(value, s_reduce_u64[w]);
      fast.partial_min_position[cta * G + g] = value == 0xFFFFFFFFull ? ~0ull : value;
      fast.partial_max_position[cta * G + g] = 0;  // only the immediate-key order needs it; such queries are not streamed
    }
    stream_consumer_barrier();
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const Accumulator lane_sum = warp_reduce_add(which == 0 ? state.raw_sum[g][c] : state.product_sum[g][c]);
        if (lane == 0) s_reduce[warp] = lane_sum;
        stream_consumer_barrier();
        if (threadIdx.x == 0) {
          Accumulator sum{};
          for (int w = 0; w < kStreamConsumerWarps; ++w) sum += s_reduce[w];
          unsigned long long bits;
          memcpy(&bits, &sum, sizeof(bits));
          (which == 0 ? fast.partial_raw : fast.partial_product)[(cta * G + g) * C + c] = bits;
        }
        stream_consumer_barrier();
      }
    }
  }
  if (threadIdx.x < G) {
    const int g = threadIdx.x;
    fast.partial_hash[cta * G + g] = s_hash[g];
    fast.partial_null_mask[cta * G + g] = 0;
    for (int w = 0; w < kMaxKeyWords; ++w) fast.partial_keys[(cta * G + g) * kMaxKeyWords + w] = s_keys[g][w];
    for (int c = 0; c < C; ++c) {
      fast.partial_raw_nulls[(cta * G + g) * C + c] = 0;
      fast.partial_product_nulls[(cta * G + g) * C + c] = 0;
    }
  }
}

}  // namespace hyb
