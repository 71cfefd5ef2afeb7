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
constexpr int kStreamConsumerWarps = 12;
constexpr int kStreamConsumerThreads = kStreamConsumerWarps * 32;
constexpr int kStreamThreads = kStreamConsumerThreads + 32;  // + the producer warp
constexpr int kStreamRowsPerWarp = 256;
constexpr int kStreamTileRows = kStreamConsumerWarps * kStreamRowsPerWarp;  // 4096
constexpr int kStreamSteps = kStreamRowsPerWarp / 32;
constexpr int kStreamMaxColumns = 12;   // distinct staged columns (predicates + group-by + values)
constexpr int kStreamUnitTiles = 4;
constexpr uint32_t kStreamEnd = 0xFFFFFFFFu;

struct StreamColumn {
  const DevSegment* segments;  // descriptors of the column, one per chunk
  uint32_t slot_offset;        // byte offset of the column's tile slice inside a stage
};

// Header of a stage: what the producer learned about the tile's chunk, handed to the consumers with the data.
struct StreamStageInfo {
  uint32_t tile;   // kStreamEnd: no more tiles
  uint32_t chunk;
  uint32_t rows;   // valid rows of the tile (0: a predicate rules the whole chunk out, nothing was copied)
  uint32_t row0;   // first row of the tile inside its chunk
  uint32_t first_position;  // table position of the tile's first row
  uint32_t pad;
  uint32_t width[kStreamMaxColumns];       // bytes per row of a staged column in this chunk
  uint32_t dict_size[kStreamMaxColumns];
  const void* dictionary[kFastMaxColumns];       // value column: dictionary in global memory
  uint32_t value_kind[kFastMaxColumns];          // 0 = unencoded value bits, 1 = dictionary in the stage, 2 = dictionary gathered from global
  uint32_t group_entry_type[HYB_MAX_GROUPBY_COLUMNS];  // 0xFF = staged uint64 key words, else hyb_data_type of staged dictionary values
  const int32_t* predicate_minima[HYB_MAX_FUSED_PREDICATES];  // FrameOfReference block minima
  uint32_t predicate_encoding[HYB_MAX_FUSED_PREDICATES];
  ChunkTest tests[HYB_MAX_FUSED_PREDICATES];
};

struct StreamPlan {
  FastPlan fast;
  uint32_t column_count;
  StreamColumn columns[kStreamMaxColumns];
  uint32_t predicate_slot[HYB_MAX_FUSED_PREDICATES];
  uint32_t group_slot[HYB_MAX_GROUPBY_COLUMNS];
  uint32_t value_slot[kFastMaxColumns];
  uint32_t dictionary_offset[kFastMaxColumns];           // stage offset of value column c's staged dictionary
  uint32_t group_words_offset[HYB_MAX_GROUPBY_COLUMNS];  // stage offset of group-by column q's per-entry key data
  uint32_t info_offset;
  uint32_t stage_bytes;
};

__device__ __forceinline__ void stream_consumer_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(kStreamConsumerThreads) : "memory");
}

__device__ __forceinline__ uint32_t stream_code(const unsigned char* slot, uint32_t width, uint32_t local) {
  if (width == 1) return slot[local];
  if (width == 2) return reinterpret_cast<const uint16_t*>(slot)[local];
  return reinterpret_cast<const uint32_t*>(slot)[local];
}

// One row against one predicate; `code` is the row's entry of the streamed vector. No NULLs on this path.
__device__ __forceinline__ bool stream_test(const ChunkTest& test, uint32_t encoding, const int32_t* minima, uint32_t code,
                                            uint32_t row) {
  switch (test.mode) {
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

template <int W, int G, int C>
__global__ void __launch_bounds__(kStreamThreads, 1) aggregate_stream_kernel(const __grid_constant__ StreamPlan plan) {
  using Value = typename WorkType<W>::Value;
  using Accumulator = typename WorkType<W>::Accumulator;
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
    s_hash[threadIdx.x] = 0;
    for (int w = 0; w < kMaxKeyWords; ++w) s_keys[threadIdx.x][w] = 0;
  }
  __syncthreads();

  const uint32_t unit_count = (fast.tile_count + kStreamUnitTiles - 1) / kStreamUnitTiles;
  const uint32_t groupby_count = fast.groupby_count;
  const uint32_t predicate_count = fast.predicate_count;

  if (warp == kStreamConsumerWarps) {
    // ---- producer warp ----------------------------------------------------------------------------------------------
    // Lane roles: [0, 12) one staged column each; [12, 16) small dictionary of value column lane - 12;
    // [16, 24) key data of group-by column lane - 16; [24, 32) predicate lane - 24 (test + minima into the header).
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
        bool ruled_out = false;
        if (lane < plan.column_count) {
          const DevSegment& segment = plan.columns[lane].segments[chunk];
          const char* base;
          const uint32_t width = segment_stream(segment, base);
          info->width[lane] = width;
          info->dict_size[lane] = segment.dict_size;
          source = base + size_t{row0} * width;
          destination = stage_base + plan.columns[lane].slot_offset;
          bytes = (rows * width + 15u) & ~15u;
        } else if (lane >= 12 && lane < 12 + C) {
          const int c = lane - 12;
          if (fast.value_segments[c] != nullptr) {
            const DevSegment& segment = fast.value_segments[c][chunk];
            info->dictionary[c] = segment.values;
            if (segment.encoding != HYB_ENC_DICTIONARY) {
              info->value_kind[c] = 0;
            } else if (segment.dict_size <= kStagedDictionary) {
              info->value_kind[c] = 1;
              source = segment.values;
              destination = stage_base + plan.dictionary_offset[c];
              bytes = (segment.dict_size * static_cast<uint32_t>(sizeof(Value)) + 15u) & ~15u;
            } else {
              info->value_kind[c] = 2;
            }
          }
        } else if (lane >= 16 && lane < 16 + groupby_count) {
          const int q = lane - 16;
          const DevSegment& segment = fast.group_segments[q][chunk];
          const uint32_t entry_bytes =
              (segment.dict_codes || segment.data_type == HYB_TYPE_INT64 || segment.data_type == HYB_TYPE_FLOAT64) ? 8u : 4u;
          info->group_entry_type[q] = segment.dict_codes ? 0xFFu : segment.data_type;
          source = segment.dict_codes ? static_cast<const void*>(segment.dict_codes) : segment.values;
          destination = stage_base + plan.group_words_offset[q];
          bytes = (segment.dict_size * entry_bytes + 15u) & ~15u;
        } else if (lane >= 24 && lane < 24 + predicate_count) {
          const int p = lane - 24;
          const ChunkTest test = fast.predicate_tests[p][chunk];
          const DevSegment& segment = fast.predicate_segments[p][chunk];
          info->tests[p] = test;
          info->predicate_minima[p] = static_cast<const int32_t*>(segment.values);
          info->predicate_encoding[p] = segment.encoding;
          ruled_out = test.mode == kTestNone;
        }
        const bool skip = __any_sync(kFullMask, ruled_out);
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
  Accumulator raw_sum[G][C], product_sum[G][C];
  uint32_t rows_seen[G], first_position[G], last_position[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    rows_seen[g] = 0;
    first_position[g] = 0xFFFFFFFFu;
    last_position[g] = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      raw_sum[g][c] = Accumulator{};
      product_sum[g][c] = Accumulator{};
    }
  }
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
  const uint32_t need_raw_mask = fast.need_raw_mask, need_product_mask = fast.need_product_mask;
  uint32_t combo_chunk = 0xFFFFFFFFu;   // chunk the warp's combination table was built for
  uint32_t combo_stride[G == 1 ? 1 : HYB_MAX_GROUPBY_COLUMNS];
  uint32_t combos = 0;
  uint8_t* my_combos = s_combo_group[G == 1 ? 0 : warp];
  if (G == 1 && threadIdx.x == 0) s_hash[0] = 1;  // the single group always exists

  for (uint32_t iteration = 0;; ++iteration) {
    const uint32_t stage = iteration % kStreamStages;
    mbarrier_wait(&s_full[stage], (iteration / kStreamStages) & 1u);
    const unsigned char* stage_base = s_stages + size_t{stage} * plan.stage_bytes;
    const auto* info = reinterpret_cast<const StreamStageInfo*>(stage_base + plan.info_offset);
    if (info->tile == kStreamEnd) break;
    const uint32_t rows = info->rows;
    if (rows != 0) {
      if constexpr (G > 1) {
        if (info->chunk != combo_chunk) {
          // New chunk, new dictionaries: rebuild this warp's value-ID combination -> group table from the staged key words
          // (lookup only: a combination that never occurs must not claim a group slot).
          combo_chunk = info->chunk;
          combos = 1;
#pragma unroll
          for (int q = 0; q < HYB_MAX_GROUPBY_COLUMNS; ++q) {
            combo_stride[q] = combos;
            if (static_cast<uint32_t>(q) < groupby_count) combos *= info->dict_size[plan.group_slot[q]];
          }
          __syncwarp();
          for (uint32_t combination = lane; combination < combos; combination += 32) {
            unsigned long long hash = 0x9E3779B97F4A7C15ull;
            uint32_t rest = combination;
            for (uint32_t q = 0; q < groupby_count; ++q) {
              const uint32_t size = info->dict_size[plan.group_slot[q]];
              const uint32_t value_id = rest % size;
              rest /= size;
              const unsigned long long entry = stream_key_entry(stage_base + plan.group_words_offset[q], info->group_entry_type[q], value_id);
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
      const uint32_t warp_row0 = warp * kStreamRowsPerWarp;
#pragma unroll 1
      for (int step = 0; step < kStreamSteps; ++step) {
        const uint32_t local = warp_row0 + step * 32 + lane;
        if (warp_row0 + step * 32 >= rows) break;  // uniform
        const bool valid = local < rows;
        bool pass = valid;
        for (uint32_t p = 0; p < predicate_count; ++p) {
          const uint32_t slot = plan.predicate_slot[p];
          const uint32_t code = stream_code(stage_base + plan.columns[slot].slot_offset, info->width[slot], local);
          pass = pass && stream_test(info->tests[p], info->predicate_encoding[p], info->predicate_minima[p], code,
                                     info->row0 + local);
        }
        if (!__any_sync(kFullMask, pass)) continue;

        // ---- group of the row ---------------------------------------------------------------------------------------
        int32_t group = pass ? 0 : -1;
        if constexpr (G > 1) {
          uint32_t combination = 0;
#pragma unroll
          for (int q = 0; q < HYB_MAX_GROUPBY_COLUMNS; ++q) {
            if (static_cast<uint32_t>(q) < groupby_count) {
              const uint32_t slot = plan.group_slot[q];
              combination += stage_base[plan.columns[slot].slot_offset + local] * combo_stride[q];
            }
          }
          combination = valid ? combination : 0u;
          group = pass ? static_cast<int32_t>(my_combos[combination]) : -1;
          if (__any_sync(kFullMask, group >= static_cast<int32_t>(kComboOverflow))) {
            // first sighting of a combination in this warp and chunk (rare): resolve / insert into the CTA's group table
            if (group == kComboUnresolved) {
              unsigned long long entries[kMaxKeyWords];
              unsigned long long hash = 0x9E3779B97F4A7C15ull;
              uint32_t rest = combination;
              for (uint32_t q = 0; q < groupby_count; ++q) {
                const uint32_t size = info->dict_size[plan.group_slot[q]];
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
                my_combos[combination] = kComboOverflow;
              } else {
                my_combos[combination] = static_cast<uint8_t>(found);
              }
              group = found;
            } else if (group == kComboOverflow) {
              group = -1;
            }
          }
        }

        // ---- row count, first / last position per group ----------------------------------------------------------------
        {
          const uint32_t position = info->first_position + local;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            if (group == g) {
              ++rows_seen[g];
              first_position[g] = min(first_position[g], position);
              last_position[g] = position;
            }
          }
        }

        // ---- value columns: raw sums and the running product ------------------------------------------------------------
        Value product = Value{};
#pragma unroll
        for (int c = 0; c < C; ++c) {
          if (fast.value_segments[c] == nullptr) continue;
          const uint32_t slot = plan.value_slot[c];
          uint32_t code = stream_code(stage_base + plan.columns[slot].slot_offset, info->width[slot], local);
          code = valid ? code : 0u;
          const uint32_t kind = info->value_kind[c];
          Value value;
          if (kind == 1) {
            value = reinterpret_cast<const Value*>(stage_base + plan.dictionary_offset[c])[code];
          } else if (kind == 2) {
            value = pass ? typed_load<W>(info->dictionary[c], 0, code) : Value{};
          } else {
            if constexpr (W == 0) {
              value = __uint_as_float(code);
            } else {
              value = Value{};  // unencoded 8-byte values are not streamed (host eligibility)
            }
          }
          const bool in_chain = (need_product_mask >> c) != 0;  // some product at or after this column
          if (in_chain) {
            const Value factor = apply_affine<W>(affine_a[c], affine_b[c], value);
            product = c == 0 ? factor : multiply<W>(product, factor);
          }
          if ((need_raw_mask >> c) & 1u) {
            const Accumulator widened = static_cast<Accumulator>(value);
#pragma unroll
            for (int g = 0; g < G; ++g) add_where(raw_sum[g][c], widened, group, g);
          }
          if ((need_product_mask >> c) & 1u) {
            const Accumulator widened = static_cast<Accumulator>(product);
#pragma unroll
            for (int g = 0; g < G; ++g) add_where(product_sum[g][c], widened, group, g);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbarrier_arrive(&s_empty[stage]);
  }

  // ---- CTA reduction in a fixed order: lanes (butterfly), then warps; partials in the layout of aggregate_fast_kernel -------
  stream_consumer_barrier();
  const size_t cta = blockIdx.x;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const unsigned long long total_rows = warp_reduce_add(static_cast<unsigned long long>(rows_seen[g]));
    uint32_t low = first_position[g], high = last_position[g];
#pragma unroll
    for (int delta = 16; delta > 0; delta >>= 1) {
      low = min(low, __shfl_xor_sync(kFullMask, low, delta));
      high = max(high, __shfl_xor_sync(kFullMask, high, delta));
    }
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
    }
    stream_consumer_barrier();
    if (lane == 0) s_reduce_u64[warp] = high;
    stream_consumer_barrier();
    if (threadIdx.x == 0) {
      unsigned long long value = 0;
      for (int w = 0; w < kStreamConsumerWarps; ++w) value = max(value, s_reduce_u64[w]);
      fast.partial_max_position[cta * G + g] = value;
    }
    stream_consumer_barrier();
#pragma unroll
    for (int c = 0; c < C; ++c) {
#pragma unroll
      for (int which = 0; which < 2; ++which) {
        const Accumulator lane_sum = warp_reduce_add(which == 0 ? raw_sum[g][c] : product_sum[g][c]);
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
