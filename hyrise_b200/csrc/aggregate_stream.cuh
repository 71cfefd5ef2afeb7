// aggregate_stream_kernel / aggregate_stream_static_kernel — the low-cardinality AggregateHash path (Q1 / Q6 shapes) built
// around the TMA unit.
//
// Included by aggregate.cu (uses FastPlan, WorkType, apply_affine, multiply, add_where, key entries and the per-CTA partials
// layout of aggregate_fast_kernel; the host merge of the partials is shared). Design, with the measurement behind each point
// (profiles/README.md):
//
//   * aggregate_fast_kernel was instruction-bound, not bandwidth-bound (410 M warp instructions for 60 M rows). Here a lane
//     handles 4 consecutive rows per step with scalar code; for the query shapes in HYB_STREAM_SHAPES the row loop is compiled
//     with widths, kinds and the set of sums folded (~95 thread instructions per row for Q1).
//   * one persistent CTA per SM, 15 consumer warps + 1 producer warp, up to 128 registers per thread.
//   * no load latency on the compute warps: the producer walks a static tile schedule ahead of the consumers and has the TMA
//     unit copy — per tile, one cp.async.bulk per array — the tile's slice of every referenced column into a shared-memory
//     ring of StreamPlan::stage_count stages (mbarrier full/empty pairs). What it derives from segment descriptors is cached in
//     registers per chunk; small dictionaries and group-key words are copied once per stage and chunk. Consumers touch global
//     memory only for dictionaries too large to stage (gathered through L1).
//   * no CTA-wide barrier in the row loop: consumer warps run tile after tile on their own; the value-ID combination -> group
//     table is private to a warp and rebuilt from the staged key words when the chunk changes.
//   * sums: one DFMA per (row, group, sum) with a 0/1 group mask; steps with non-finite operands take the select form.
//
// Eligibility is decided on the host per call (stream_layout_for): every referenced column streams a fixed-width vector of
// <= 4 bytes per row, carries no NULLs, group-by columns are dictionary segments with 1-byte value-IDs whose combinations fit
// kMaxCombos; everything else keeps aggregate_fast_kernel / aggregate_general_kernel.
#pragma once

namespace hyb {

// One persistent CTA per SM: 15 consumer warps + the producer warp = 16 warps, which may use 128 registers each (a second
// CTA of 9 warps would cap the row loop at 80 registers: the register file is allocated in units of 4 warps).
constexpr int kStreamMaxStages = 8;  // ring depth is a launch parameter (StreamPlan::stage_count): fewer stages leave more L1 for gathers
constexpr int kStreamConsumerWarps = 15;
constexpr int kStreamConsumerThreads = kStreamConsumerWarps * 32;
constexpr int kStreamThreads = kStreamConsumerThreads + 32;  // + the producer warp
constexpr int kStreamRowsPerWarp = 128;
constexpr int kStreamTileRows = kStreamConsumerWarps * kStreamRowsPerWarp;  // 1920
constexpr int kStreamLaneRows = 4;                                          // consecutive rows per lane and step
constexpr int kStreamSteps = kStreamRowsPerWarp / (32 * kStreamLaneRows);   // 1
constexpr int kStreamMaxColumns = 12;   // distinct staged columns (predicates + group-by + values)
constexpr int kStreamRounds = 4;        // a CTA's tiles: kStreamRounds contiguous runs (chunk locality) spread over the table
constexpr size_t kStreamMaxDynamicBytes = 216 * 1024;
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
  uint32_t array_chunk;     // producer's note: the chunk whose dictionaries / key words this stage holds (survives refills)
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
  uint32_t unit_tiles;  // tiles per contiguous run; run u belongs to CTA u % gridDim.x
  uint32_t stage_count; // depth of the shared-memory ring (2 .. kStreamMaxStages)
};

// ---- compile-time row-loop shapes --------------------------------------------------------------------------------------
// The row loop exists in a layout-generic form (kShape == 0: vector widths, value kinds, test modes and the set of sums are
// uniform constant-bank reads and uniform branches) and in instantiations for the shapes listed in HYB_STREAM_SHAPES, where all
// of that is folded at compile time: no dead sums in registers, no width / kind / mode switches, one SETP per (row, group)
// shared by all of the row's predicated adds. The host derives the shape word of a plan (stream_shape_of) and launches the
// matching instantiation when the registry holds one. Layout of the word:
//   bits [8c, 8c + 8)   value column c: [0,2) width code (0 absent, 1: 1 B, 2: 2 B, 3: 4 B)  [2,4) value kind
//                                        [4,7) AffineKind  [7] raw sum needed
//   bits [32, 36)       product mask          bits [36, 39) predicate count (<= 4)
//   bits [39 + 4p, +4)  predicate p: [0,2) width code  [2,4) mode (0 value-ID range, 1 unencoded int, 2 float)
//   bits [55, 58)       group-by column count  bit 63: static
constexpr uint64_t kShapeStatic = 1ull << 63;
constexpr int kShapeMaxPredicates = 4;
__host__ __device__ constexpr uint32_t shape_width_code(uint32_t width) { return width == 4 ? 3u : width; }
__host__ __device__ constexpr uint32_t shape_code_width(uint32_t code) { return code == 3 ? 4u : code; }
__host__ __device__ constexpr uint64_t shape_value(int c, uint32_t width, uint32_t kind, int32_t affine, bool raw) {
  return static_cast<uint64_t>(shape_width_code(width) | (kind << 2) | (static_cast<uint32_t>(affine) << 4) | (raw ? 0x80u : 0u)) << (8 * c);
}
__host__ __device__ constexpr uint64_t shape_products(uint32_t mask) { return static_cast<uint64_t>(mask & 0xFu) << 32; }
__host__ __device__ constexpr uint64_t shape_predicate(int p, uint32_t width, uint32_t mode_code) {
  return static_cast<uint64_t>(shape_width_code(width) | (mode_code << 2)) << (39 + 4 * p);
}
__host__ __device__ constexpr uint64_t shape_counts(uint32_t predicates, uint32_t groupby) {
  return (static_cast<uint64_t>(predicates) << 36) | (static_cast<uint64_t>(groupby) << 55) | kShapeStatic;
}

template <uint64_t S>
struct StreamShape {
  static constexpr bool kStatic = S != 0;
  __device__ static __forceinline__ uint32_t value_width(const StreamPlan& plan, int c) {
    return kStatic ? shape_code_width(static_cast<uint32_t>(S >> (8 * c)) & 3u) : plan.value_width[c];
  }
  __device__ static __forceinline__ bool value_present(const StreamPlan& plan, int c) {
    return kStatic ? ((S >> (8 * c)) & 3u) != 0 : plan.fast.value_segments[c] != nullptr;
  }
  __device__ static __forceinline__ uint32_t value_kind(const StreamPlan& plan, int c) {
    return kStatic ? static_cast<uint32_t>(S >> (8 * c + 2)) & 3u : plan.value_kind[c];
  }
  __host__ __device__ static constexpr int32_t affine_kind(int c) { return static_cast<int32_t>((S >> (8 * c + 4)) & 7u); }
  __host__ __device__ static constexpr uint32_t raw_mask() {
    return static_cast<uint32_t>(((S >> 7) & 1u) | ((S >> 14) & 2u) | ((S >> 21) & 4u) | ((S >> 28) & 8u));
  }
  __host__ __device__ static constexpr uint32_t product_mask() { return static_cast<uint32_t>(S >> 32) & 0xFu; }
  __device__ static __forceinline__ uint32_t predicate_count(const StreamPlan& plan) {
    return kStatic ? static_cast<uint32_t>(S >> 36) & 7u : plan.fast.predicate_count;
  }
  __device__ static __forceinline__ uint32_t predicate_width(const StreamPlan& plan, int p) {
    return kStatic ? shape_code_width(static_cast<uint32_t>(S >> (39 + 4 * p)) & 3u) : plan.predicate_width[p];
  }
  __device__ static __forceinline__ uint32_t predicate_mode(const StreamPlan& plan, int p) {
    if (!kStatic) return plan.predicate_mode[p];
    const uint32_t code = static_cast<uint32_t>(S >> (41 + 4 * p)) & 3u;
    return code == 0 ? kTestIdRange : code == 1 ? kTestInt : kTestFloat;
  }
  __device__ static __forceinline__ uint32_t groupby_count(const StreamPlan& plan) {
    return kStatic ? static_cast<uint32_t>(S >> 55) & 7u : plan.fast.groupby_count;
  }
  // static shapes: the sums a row feeds, in the order (column 0 raw, column 0 product, column 1 raw, ...)
  __host__ __device__ static constexpr int sum_count() {
    int count = 0;
    for (int c = 0; c < 4; ++c) count += ((raw_mask() >> c) & 1) + ((product_mask() >> c) & 1);
    return count;
  }
  __host__ __device__ static constexpr int sum_source(int k) {  // 2 * column + (1: product)
    int seen = 0;
    for (int c = 0; c < 4; ++c) {
      if ((raw_mask() >> c) & 1) {
        if (seen == k) return 2 * c;
        ++seen;
      }
      if ((product_mask() >> c) & 1) {
        if (seen == k) return 2 * c + 1;
        ++seen;
      }
    }
    return 0;
  }
};

// acc[k] += v[k] for every k if group == g: one SETP and K predicated adds (the compiler turns the C++ form into a
// DADD and two selects per sum; separate asm statements per sum would each carry their own SETP).
template <int K>
__device__ __forceinline__ void add_where_all(double (&acc)[K], const double (&v)[K], int group, int g) {
  static_assert(K >= 1 && K <= 8, "at most two sums per value column");
  if constexpr (K == 1) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %2, %3;\n\t@p add.rn.f64 %0, %0, %1;\n\t}"
        : "+d"(acc[0])
        : "d"(v[0]), "r"(group), "r"(g));
  } else if constexpr (K == 2) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %4, %5;\n\t@p add.rn.f64 %0, %0, %2;\n\t@p add.rn.f64 %1, %1, %3;\n\t}"
        : "+d"(acc[0]), "+d"(acc[1])
        : "d"(v[0]), "d"(v[1]), "r"(group), "r"(g));
  } else if constexpr (K == 3) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %6, %7;\n\t@p add.rn.f64 %0, %0, %3;\n\t@p add.rn.f64 %1, %1, %4;\n\t@p add.rn.f64 %2, %2, %5;\n\t}"
        : "+d"(acc[0]), "+d"(acc[1]), "+d"(acc[2])
        : "d"(v[0]), "d"(v[1]), "d"(v[2]), "r"(group), "r"(g));
  } else if constexpr (K == 4) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %8, %9;\n\t@p add.rn.f64 %0, %0, %4;\n\t@p add.rn.f64 %1, %1, %5;\n\t@p add.rn.f64 %2, %2, %6;\n\t@p add.rn.f64 %3, %3, %7;\n\t}"
        : "+d"(acc[0]), "+d"(acc[1]), "+d"(acc[2]), "+d"(acc[3])
        : "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "r"(group), "r"(g));
  } else if constexpr (K == 5) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %10, %11;\n\t@p add.rn.f64 %0, %0, %5;\n\t@p add.rn.f64 %1, %1, %6;\n\t@p add.rn.f64 %2, %2, %7;\n\t@p add.rn.f64 %3, %3, %8;\n\t@p add.rn.f64 %4, %4, %9;\n\t}"
        : "+d"(acc[0]), "+d"(acc[1]), "+d"(acc[2]), "+d"(acc[3]), "+d"(acc[4])
        : "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "d"(v[4]), "r"(group), "r"(g));
  } else if constexpr (K == 6) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %12, %13;\n\t@p add.rn.f64 %0, %0, %6;\n\t@p add.rn.f64 %1, %1, %7;\n\t@p add.rn.f64 %2, %2, %8;\n\t@p add.rn.f64 %3, %3, %9;\n\t@p add.rn.f64 %4, %4, %10;\n\t@p add.rn.f64 %5, %5, %11;\n\t}"
        : "+d"(acc[0]), "+d"(acc[1]), "+d"(acc[2]), "+d"(acc[3]), "+d"(acc[4]), "+d"(acc[5])
        : "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "d"(v[4]), "d"(v[5]), "r"(group), "r"(g));
  } else if constexpr (K == 7) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %14, %15;\n\t@p add.rn.f64 %0, %0, %7;\n\t@p add.rn.f64 %1, %1, %8;\n\t@p add.rn.f64 %2, %2, %9;\n\t@p add.rn.f64 %3, %3, %10;\n\t@p add.rn.f64 %4, %4, %11;\n\t@p add.rn.f64 %5, %5, %12;\n\t@p add.rn.f64 %6, %6, %13;\n\t}"
        : "+d"(acc[0]), "+d"(acc[1]), "+d"(acc[2]), "+d"(acc[3]), "+d"(acc[4]), "+d"(acc[5]), "+d"(acc[6])
        : "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "d"(v[4]), "d"(v[5]), "d"(v[6]), "r"(group), "r"(g));
  } else if constexpr (K == 8) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %16, %17;\n\t@p add.rn.f64 %0, %0, %8;\n\t@p add.rn.f64 %1, %1, %9;\n\t@p add.rn.f64 %2, %2, %10;\n\t@p add.rn.f64 %3, %3, %11;\n\t@p add.rn.f64 %4, %4, %12;\n\t@p add.rn.f64 %5, %5, %13;\n\t@p add.rn.f64 %6, %6, %14;\n\t@p add.rn.f64 %7, %7, %15;\n\t}"
        : "+d"(acc[0]), "+d"(acc[1]), "+d"(acc[2]), "+d"(acc[3]), "+d"(acc[4]), "+d"(acc[5]), "+d"(acc[6]), "+d"(acc[7])
        : "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]), "d"(v[4]), "d"(v[5]), "d"(v[6]), "d"(v[7]), "r"(group), "r"(g));
  }
}


__device__ __forceinline__ void stream_consumer_barrier() {
  asm volatile("bar.sync 1, %0;" ::"n"(kStreamConsumerThreads) : "memory");
}

// Four consecutive entries of a staged vector of `width` bytes per entry, `at` = address of the first one (16-byte aligned for
// 4-byte entries, 8 for 2, 4 for 1): one 4 / 8 / 16-byte LDS.
__device__ __forceinline__ void stream_codes4(const unsigned char* at, uint32_t width, uint32_t (&codes)[4]) {
  if (width == 1) {
    const uint32_t v = *reinterpret_cast<const uint32_t*>(at);
    codes[0] = __byte_perm(v, 0u, 0x4440u);  // PRMT: the scaled dictionary address then is one LEA
    codes[1] = __byte_perm(v, 0u, 0x4441u);
    codes[2] = __byte_perm(v, 0u, 0x4442u);
    codes[3] = __byte_perm(v, 0u, 0x4443u);
  } else if (width == 2) {
    const uint2 v = *reinterpret_cast<const uint2*>(at);
    codes[0] = v.x & 0xFFFFu;
    codes[1] = v.x >> 16;
    codes[2] = v.y & 0xFFFFu;
    codes[3] = v.y >> 16;
  } else {
    const uint4 v = *reinterpret_cast<const uint4*>(at);
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
  constexpr int kPerThread = kStreamTileRows / kStreamConsumerThreads;  // 4
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
template <int W, int G, int C, uint64_t S>
struct StreamState {
  // static shapes: exactly the sums the query needs, in StreamShape::sum_source order; generic: slot 2c + (1: product)
  static constexpr int K = S != 0 ? (StreamShape<S>::sum_count() > 0 ? StreamShape<S>::sum_count() : 1) : 2 * C;
  typename WorkType<W>::Accumulator sums[G][K];
  uint32_t rows_seen[G];
  uint32_t first_position[G];
  uint32_t packed_rows;  // byte g = rows of group g since the last flush (< 256 by construction)
  uint32_t seen_groups;  // bit g: this thread has recorded first_position[g]
};

// The rows of one warp in one tile. Vector widths, value kinds and test modes are the plan's launch constants (tiles of a
// chunk with narrower vectors have been widened in place by stream_widen_tile), or compile-time constants of shape S.
template <int W, int G, int C, uint64_t S>
__device__ __forceinline__ void stream_warp_rows(const StreamPlan& plan, const StreamStageInfo* info, const unsigned char* stage_base,
                                                 uint32_t warp, uint32_t lane, uint8_t* my_combos, unsigned long long* s_hash,
                                                 unsigned long long (*s_keys)[kMaxKeyWords],
                                                 const uint32_t (&combo_stride)[G == 1 ? 1 : HYB_MAX_GROUPBY_COLUMNS],
                                                 const typename WorkType<W>::Value (&affine_a)[C],
                                                 const typename WorkType<W>::Value (&affine_b)[C], StreamState<W, G, C, S>& state) {
  using Value = typename WorkType<W>::Value;
  using Accumulator = typename WorkType<W>::Accumulator;
  using Shape = StreamShape<S>;
  constexpr bool kStatic = Shape::kStatic;
  const FastPlan& fast = plan.fast;
  const uint32_t rows = info->rows;
  const uint32_t groupby_count = Shape::groupby_count(plan);
  const uint32_t predicate_count = Shape::predicate_count(plan);
  const uint32_t need_raw_mask = kStatic ? Shape::raw_mask() : fast.need_raw_mask;
  const uint32_t need_product_mask = kStatic ? Shape::product_mask() : fast.need_product_mask;
#pragma unroll 1
  for (int step = 0; step < kStreamSteps; ++step) {
    const uint32_t local0 = warp * kStreamRowsPerWarp + step * (32 * kStreamLaneRows) + lane * kStreamLaneRows;
    if (warp * kStreamRowsPerWarp + step * (32 * kStreamLaneRows) >= rows) break;  // uniform
    uint32_t valid = 0xFu;
    if (rows != kStreamTileRows) {  // uniform: the last tile of a chunk
      valid = local0 >= rows ? 0u : (rows - local0 >= kStreamLaneRows ? 0xFu : ((1u << (rows - local0)) - 1u));
    }
    // the lane's first row in a staged slice of 1 / 2 / 4 bytes per row (the slice's stage offset is a uniform operand)
    const unsigned char* rows1 = stage_base + local0;
    const unsigned char* rows2 = stage_base + 2 * local0;
    const unsigned char* rows4 = stage_base + 4 * local0;
    const auto rows_at = [&](uint32_t width) {
      if constexpr (kStatic) {
        return width == 1 ? rows1 : width == 2 ? rows2 : rows4;  // folded: the widths in use keep a register each
      } else {
        return stage_base + local0 * width;
      }
    };
    uint32_t pass = valid;
#pragma unroll
    for (int p = 0; p < (kStatic ? kShapeMaxPredicates : HYB_MAX_FUSED_PREDICATES); ++p) {
      if (static_cast<uint32_t>(p) >= predicate_count) break;
      const uint32_t width = Shape::predicate_width(plan, p);
      const uint32_t mode = Shape::predicate_mode(plan, p);
      uint32_t codes[kStreamLaneRows];
      stream_codes4(rows_at(width) + plan.predicate_offset[p], width, codes);
      uint32_t matches = 0;
      if (mode == kTestIdRange) {
        // the common case (dictionary value-ID range): three broadcast reads, two instructions per row
        const uint32_t id_lo = info->tests[p].id_lo, id_span = info->tests[p].id_span;
        const uint32_t flip = info->tests[p].negate ? 0xFu : 0u;
#pragma unroll
        for (int j = 0; j < kStreamLaneRows; ++j) matches |= ((codes[j] - id_lo) < id_span) ? (1u << j) : 0u;
        matches ^= flip;
      } else {
        const uint32_t encoding = kStatic ? static_cast<uint32_t>(HYB_ENC_UNENCODED) : plan.predicate_encoding[p];
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
      if (!Shape::value_present(plan, c)) continue;
      const uint32_t width = Shape::value_width(plan, c);
      const uint32_t kind = Shape::value_kind(plan, c);
      uint32_t codes[kStreamLaneRows];
      stream_codes4(rows_at(width) + plan.value_offset[c], width, codes);
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
          stream_codes4(rows1 + plan.group_offset[q], 1u, codes);
#pragma unroll
          for (int j = 0; j < kStreamLaneRows; ++j) combination[j] = q == 0 ? codes[j] : combination[j] + codes[j] * combo_stride[q];
        }
      }
      uint32_t highest = 0;
#pragma unroll
      for (int j = 0; j < kStreamLaneRows; ++j) {
        // the mask is the identity for real rows; rows past the tile's end carry stale codes (one column: a byte as it is)
        if (!(kStatic && Shape::groupby_count(plan) == 1)) combination[j] &= kMaxCombos - 1;
        group[j] = ((pass >> j) & 1u) ? my_combos[combination[j]] : static_cast<uint32_t>(G);
        highest = max(highest, group[j]);
      }
      if (__any_sync(kFullMask, highest >= kComboOverflow)) {
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
        // group == G for rows that do not count: its shift is 32 (shl.b32 clamps: 0), its bit lies outside the G low ones
        uint32_t increment, bit;
        asm("shl.b32 %0, 1, %1;" : "=r"(increment) : "r"(group[j] * (G == 1 ? 32u : 8u)));
        asm("shl.b32 %0, 1, %1;" : "=r"(bit) : "r"(group[j]));
        state.packed_rows += increment;
        step_groups |= bit;
      }
      step_groups &= (1u << G) - 1u;
      if (step_groups & ~state.seen_groups) {  // rare after the first tiles: a lane meets a group for the first time
        // positions only grow, so the minimum is the first one (static register indexes: no select chain on `group`)
#pragma unroll
        for (int j = 0; j < kStreamLaneRows; ++j) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            state.first_position[g] =
                min(state.first_position[g], group[j] == static_cast<uint32_t>(g) ? info->first_position + local0 + j : 0xFFFFFFFFu);
          }
        }
        state.seen_groups |= step_groups;
      }
    }

    // ---- value columns: raw sums and the running product ------------------------------------------------------------
    if constexpr (kStatic) {
      constexpr int K = Shape::sum_count();
      if constexpr (K > 0) {
        // A conditional FP64 add costs three issue slots (ptxas turns `@p add.f64` into DADD + two FSEL). Instead every sum
        // takes every row through one DFMA with a 0.0 / 1.0 group mask: fma(v, 1, acc) == acc + v exactly and fma(v, 0, acc)
        // == acc — as long as v is finite (0 * inf = NaN would leak into the other groups). A step with a non-finite
        // operand anywhere in the warp takes the select form below; the last product of the chain being finite implies that
        // every factor before it is.
        Value operands[kStreamLaneRows][K];  // widened where they are consumed: half the registers
        bool finite = true;
#pragma unroll
        for (int j = 0; j < kStreamLaneRows; ++j) {
          Value chain[C];  // chain[c] = f0(col0) * ... * fc(colc)
          Value magnitude{};
          bool first = true;  // folded: which addends exist is a property of the shape
          const auto add_magnitude = [&](Value value) {
            Value absolute;
            if constexpr (W == 0) {
              absolute = fabsf(value);
            } else {
              absolute = fabs(value);
            }
            magnitude = first ? absolute : magnitude + absolute;
            first = false;
          };
#pragma unroll
          for (int c = 0; c < C; ++c) {
            if ((Shape::product_mask() >> c) == 0) {  // no product at or after this column
              if ((Shape::raw_mask() >> c) & 1u) add_magnitude(values[c][j]);
              continue;
            }
            const Value factor =
                Shape::affine_kind(c) == kIdentity ? values[c][j] : apply_affine<W>(affine_a[c], affine_b[c], values[c][j]);
            chain[c] = c == 0 ? factor : multiply<W>(chain[c == 0 ? 0 : c - 1], factor);
            if ((Shape::product_mask() >> c) == 1u) add_magnitude(chain[c]);  // the last product
          }
          finite = finite && magnitude <= (W == 0 ? Value(3.402823466e+38f) : Value(1.7976931348623157e+308));  // false for NaN
#pragma unroll
          for (int k = 0; k < K; ++k) {
            const int source = Shape::sum_source(k);
            operands[j][k] = (source & 1) ? chain[source >> 1] : values[source >> 1][j];
          }
        }
        if (__all_sync(kFullMask, finite)) {
#pragma unroll
          for (int j = 0; j < kStreamLaneRows; ++j) {
            Accumulator widened[K];
#pragma unroll
            for (int k = 0; k < K; ++k) widened[k] = static_cast<Accumulator>(operands[j][k]);
#pragma unroll
            for (int g = 0; g < G; ++g) {
              const Accumulator mask = group[j] == static_cast<uint32_t>(g) ? Accumulator(1) : Accumulator(0);
#pragma unroll
              for (int k = 0; k < K; ++k) state.sums[g][k] = __fma_rn(widened[k], mask, state.sums[g][k]);
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < kStreamLaneRows; ++j) {
            Accumulator widened[K];
#pragma unroll
            for (int k = 0; k < K; ++k) widened[k] = static_cast<Accumulator>(operands[j][k]);
#pragma unroll
            for (int g = 0; g < G; ++g) add_where_all<K>(state.sums[g], widened, static_cast<int>(group[j]), g);
          }
        }
      }
    } else {
      // layout-generic: predicated adds (DADD + two selects each). The masked-FMA form of the static shapes was measured
      // slower here (1.27 vs 0.97 ms for the Q1 plan): with all 2 G C accumulators live it spills.
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
            for (int g = 0; g < G; ++g) add_where(state.sums[g][2 * c], widened, static_cast<int>(group[j]), g);
          }
        }
        if ((need_product_mask >> c) & 1u) {
#pragma unroll
          for (int j = 0; j < kStreamLaneRows; ++j) {
            const Accumulator widened = static_cast<Accumulator>(product[j]);
#pragma unroll
            for (int g = 0; g < G; ++g) add_where(state.sums[g][2 * c + 1], widened, static_cast<int>(group[j]), g);
          }
        }
      }
    }
  }
}

template <int W, int G, int C, uint64_t S>
__device__ __forceinline__ void aggregate_stream_body(const StreamPlan& plan) {
  using Value = typename WorkType<W>::Value;
  using Accumulator = typename WorkType<W>::Accumulator;
  static_assert(G == 1 || G == 4, "row counts are packed as one byte per group");
  const FastPlan& fast = plan.fast;  // kernel parameter: every plan field is a uniform constant-bank read

  extern __shared__ __align__(128) unsigned char s_stages[];  // stage_count x stage_bytes
  __shared__ __align__(8) unsigned long long s_full[kStreamMaxStages], s_empty[kStreamMaxStages];
  // CTA-wide group table (<= G distinct keys), filled on first sight
  __shared__ unsigned long long s_hash[G];
  __shared__ unsigned long long s_keys[G][kMaxKeyWords];
  __shared__ uint8_t s_combo_group[G == 1 ? 1 : kStreamConsumerWarps][G == 1 ? 4 : kMaxCombos];  // per warp
  __shared__ Accumulator s_reduce[kStreamConsumerWarps];
  __shared__ unsigned long long s_reduce_u64[kStreamConsumerWarps];

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int stage = 0; stage < kStreamMaxStages; ++stage) {
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

  const uint32_t unit_count = (fast.tile_count + plan.unit_tiles - 1) / plan.unit_tiles;
  const uint32_t groupby_count = fast.groupby_count;
  const uint32_t predicate_count = fast.predicate_count;

  if (warp == kStreamConsumerWarps) {
    // ---- producer warp ----------------------------------------------------------------------------------------------
    // Lane roles: [0, 12) one staged column each; [12, 16) value column lane - 12 (header + small dictionary);
    // [16, 24) group-by column lane - 16 (header + key data of its dictionary); [24, 32) predicate lane - 24 (header).
    // Everything a lane derives from the chunk's segment descriptors is cached in its registers and reloaded only when the
    // chunk changes (a CTA's tiles are contiguous runs): per tile the producer touches global memory for the tile-map entry
    // alone, so the copies of a freed stage are issued within a few hundred cycles of its release. The per-chunk arrays
    // (staged dictionaries, key words) are not copied again into a stage that still holds this chunk's.
    uint32_t cached_chunk = 0xFFFFFFFFu, chunk_rows = 0, chunk_first_position = 0;
    const char* lane_base = nullptr;      // column lanes: first row of the chunk's vector
    uint32_t lane_width = 0;
    const void* lane_array = nullptr;     // value / group lanes: the per-chunk array to stage
    uint32_t lane_array_bytes = 0;
    uint32_t lane_dict_size = 0, lane_entry_type = 0;
    const void* lane_dictionary = nullptr;
    ChunkTest lane_test{};
    const int32_t* lane_minima = nullptr;
    bool lane_regular = true;
    for (uint32_t stage = lane; stage < plan.stage_count; stage += 32) {
      reinterpret_cast<StreamStageInfo*>(s_stages + size_t{stage} * plan.stage_bytes + plan.info_offset)->array_chunk = 0xFFFFFFFFu;
    }
    __syncwarp();
    uint32_t stage = 0, empty_parity = 1;
    for (uint32_t unit = blockIdx.x; unit < unit_count; unit += gridDim.x) {
      const uint32_t unit_end = min(fast.tile_count, (unit + 1) * plan.unit_tiles);
      uint2 where = __ldg(fast.tile_map + unit * plan.unit_tiles);
      for (uint32_t tile = unit * plan.unit_tiles; tile < unit_end; ++tile) {
        const uint32_t chunk = where.x;
        const uint32_t row0 = where.y & 0x7FFFFFFFu;
        if (tile + 1 < unit_end) where = __ldg(fast.tile_map + tile + 1);  // in flight while this tile is set up
        if (chunk != cached_chunk) {
          cached_chunk = chunk;
          chunk_rows = fast.size_segments[chunk].row_count;
          chunk_first_position = static_cast<uint32_t>(__ldg(fast.chunk_row_start + chunk));
          if (lane < plan.column_count) {
            const DevSegment& segment = plan.columns[lane].segments[chunk];
            lane_width = segment_stream(segment, lane_base);
            lane_regular = lane_width == plan.column_width[lane];  // narrower: the consumers widen the slice before the row loop
          } else if (lane >= 12 && lane < 12 + C) {
            const int c = lane - 12;
            lane_array_bytes = 0;
            if (fast.value_segments[c] != nullptr) {
              const DevSegment& segment = fast.value_segments[c][chunk];
              lane_dictionary = segment.values;
              if (plan.value_kind[c] == kValueStagedDictionary) {  // the host checked: <= kStagedDictionary entries in every chunk
                lane_array = segment.values;
                lane_array_bytes = (segment.dict_size * static_cast<uint32_t>(sizeof(Value)) + 15u) & ~15u;
              }
            }
          } else if (lane >= 16 && lane < 16 + groupby_count) {
            const DevSegment& segment = fast.group_segments[lane - 16][chunk];
            const uint32_t entry_bytes =
                (segment.dict_codes || segment.data_type == HYB_TYPE_INT64 || segment.data_type == HYB_TYPE_FLOAT64) ? 8u : 4u;
            lane_dict_size = segment.dict_size;
            lane_entry_type = segment.dict_codes ? 0xFFu : segment.data_type;
            lane_array = segment.dict_codes ? static_cast<const void*>(segment.dict_codes) : segment.values;
            lane_array_bytes = (segment.dict_size * entry_bytes + 15u) & ~15u;
          } else if (lane >= 24 && lane < 24 + predicate_count) {
            lane_test = fast.predicate_tests[lane - 24][chunk];
            lane_minima = static_cast<const int32_t*>(fast.predicate_segments[lane - 24][chunk].values);
          }
        }
        const uint32_t rows = min(static_cast<uint32_t>(kStreamTileRows), chunk_rows - row0);

        if (lane == 0) mbarrier_wait(&s_empty[stage], empty_parity);
        __syncwarp();
        unsigned char* stage_base = s_stages + size_t{stage} * plan.stage_bytes;
        auto* info = reinterpret_cast<StreamStageInfo*>(stage_base + plan.info_offset);
        const bool arrays_staged = info->array_chunk == chunk;  // this stage's previous tile came from the same chunk
        const void* source = nullptr;
        void* destination = nullptr;
        uint32_t bytes = 0;
        bool ruled_out = false;
        if (lane < plan.column_count) {
          info->column_width[lane] = lane_width;
          source = lane_base + size_t{row0} * lane_width;
          destination = stage_base + plan.columns[lane].slot_offset;
          bytes = (rows * lane_width + 15u) & ~15u;
        } else if (lane >= 12 && lane < 12 + C) {
          const int c = lane - 12;
          info->dictionary[c] = lane_dictionary;
          if (!arrays_staged && lane_array_bytes) {
            source = lane_array;
            destination = stage_base + plan.dictionary_offset[c];
            bytes = lane_array_bytes;
          }
        } else if (lane >= 16 && lane < 16 + groupby_count) {
          const int q = lane - 16;
          info->group_dict_size[q] = lane_dict_size;
          info->group_entry_type[q] = lane_entry_type;
          if (!arrays_staged) {
            source = lane_array;
            destination = stage_base + plan.group_words_offset[q];
            bytes = lane_array_bytes;
          }
        } else if (lane >= 24 && lane < 24 + predicate_count) {
          const int p = lane - 24;
          info->tests[p] = lane_test;
          info->predicate_minima[p] = lane_minima;
          ruled_out = lane_test.mode == kTestNone;
        }
        const bool skip = __any_sync(kFullMask, ruled_out);
        const bool all_regular = __all_sync(kFullMask, lane_regular);
        if (skip) bytes = 0;
        uint32_t total = bytes;
#pragma unroll
        for (int delta = 16; delta > 0; delta >>= 1) total += __shfl_xor_sync(kFullMask, total, delta);
        __syncwarp();  // every lane has read array_chunk
        if (lane == 0) {
          info->tile = tile;
          info->chunk = chunk;
          info->rows = skip ? 0u : rows;
          info->row0 = row0;
          info->first_position = chunk_first_position + row0;
          info->regular = all_regular ? 1u : 0u;
          if (!skip) info->array_chunk = chunk;
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
        if (++stage == plan.stage_count) {
          stage = 0;
          empty_parity ^= 1u;
        }
      }
    }
    if (lane == 0) {
      mbarrier_wait(&s_empty[stage], empty_parity);
      reinterpret_cast<StreamStageInfo*>(s_stages + size_t{stage} * plan.stage_bytes + plan.info_offset)->tile = kStreamEnd;
      mbarrier_arrive(&s_full[stage]);
    }
    return;
  }

  // ---- consumer warps ---------------------------------------------------------------------------------------------------
  using State = StreamState<W, G, C, S>;
  State state;
  state.packed_rows = 0;
  state.seen_groups = 0;
#pragma unroll
  for (int g = 0; g < G; ++g) {
    state.rows_seen[g] = 0;
    state.first_position[g] = 0xFFFFFFFFu;
#pragma unroll
    for (int k = 0; k < State::K; ++k) state.sums[g][k] = Accumulator{};
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

  uint32_t stage = 0, phase = 0, since_flush = 0;
  for (;; stage = stage + 1 == plan.stage_count ? 0 : stage + 1, phase ^= stage == 0 ? 1u : 0u) {
    mbarrier_wait(&s_full[stage], phase);
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
      stream_warp_rows<W, G, C, S>(plan, info, stage_base, warp, lane, my_combos, s_hash, s_keys, combo_stride, affine_a,
                                   affine_b, state);
      if (++since_flush == 16) {  // 4 rows per lane and tile: 64 per flush, the bytes stay below 256
        flush_row_counts();
        since_flush = 0;
      }
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
        // the accumulator of (column c, raw | product): static shapes hold the needed ones only
        int slot = 2 * c + which;
        bool held = true;
        if constexpr (S != 0) {
          slot = -1;
#pragma unroll
          for (int k = 0; k < StreamShape<S>::sum_count(); ++k) {
            if (StreamShape<S>::sum_source(k) == 2 * c + which) slot = k;
          }
          held = slot >= 0;
        }
        unsigned long long* destination = (which == 0 ? fast.partial_raw : fast.partial_product) + (cta * G + g) * C + c;
        if (!held) {
          if (threadIdx.x == 0) *destination = 0;
          continue;
        }
        const Accumulator lane_sum = warp_reduce_add(state.sums[g][slot < 0 ? 0 : slot]);
        if (lane == 0) s_reduce[warp] = lane_sum;
        stream_consumer_barrier();
        if (threadIdx.x == 0) {
          Accumulator sum{};
          for (int w = 0; w < kStreamConsumerWarps; ++w) sum += s_reduce[w];
          unsigned long long bits;
          memcpy(&bits, &sum, sizeof(bits));
          *destination = bits;
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

template <int W, int G, int C>
__global__ void __launch_bounds__(kStreamThreads, 1) aggregate_stream_kernel(const __grid_constant__ StreamPlan plan) {
  aggregate_stream_body<W, G, C, 0>(plan);
}

template <int W, int G, int C, uint64_t S>
__global__ void __launch_bounds__(kStreamThreads, 1) aggregate_stream_static_kernel(const __grid_constant__ StreamPlan plan) {
  static_assert(S != 0, "shape 0 is the layout-generic kernel");
  aggregate_stream_body<W, G, C, S>(plan);
}

}  // namespace hyb
