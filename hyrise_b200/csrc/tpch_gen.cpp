// Seeded TPC-H-shaped generator for lineitem / orders, emitting already-encoded segments (see
// include/hyrise_b200_tpch.h). Host-only tooling for bench.py and the tests; no CUDA in this file.
//
// Encoders follow the reference: DictionaryEncoder (dictionary_encoder.hpp:33-103: sorted unique dictionary per
// segment, value-ID = lower_bound, FixedWidthInteger width from the NULL value-ID = dictionary size),
// FrameOfReferenceEncoder (frame_of_reference_encoder.hpp:25-122: minimum per 2048-row block, offsets compressed by
// the maximum offset of the segment).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hyrise_b200_tpch.h"

namespace {

constexpr uint32_t kChunkSize = HYB_DEFAULT_CHUNK_SIZE;
constexpr uint32_t kOrderBlock = 4096;
constexpr int32_t kOrderDateRange = 2406;   // 1992-01-01 .. 1998-08-02 (dss.h:331-333)
constexpr int32_t kCutoffDay = 1263;        // 1995-06-17: returnflag / linestatus switch (dss.h:380)
constexpr size_t kAlign = 256;
constexpr size_t kTailPad = 64;
constexpr size_t kBlockBytes = size_t{256} << 20;

inline uint64_t mix64(uint64_t x) {  // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct Rng {
  uint64_t seed;
  inline uint64_t at(uint64_t order, uint32_t line, uint32_t field) const {
    return mix64(mix64(seed ^ (order * 8 + line)) + field * 0xD6E8FEB86659FD93ull);
  }
};

struct BlockAllocator {
  hyb_tpch_alloc_fn alloc;
  hyb_tpch_free_fn free_fn;
  std::mutex mutex;
  std::vector<std::pair<char*, size_t>> blocks;
  std::vector<size_t> used;  // bytes handed out per block
  size_t offset = 0;
  size_t total = 0;

  void* allocate(size_t bytes) {
    const size_t need = ((bytes + kTailPad + kAlign - 1) / kAlign) * kAlign;
    std::lock_guard<std::mutex> lock(mutex);
    if (blocks.empty() || offset + need > blocks.back().second) {
      const size_t size = std::max(need, kBlockBytes);
      char* base = static_cast<char*>(alloc(size));
      if (!base) return nullptr;
      std::memset(base, 0, size);
      blocks.emplace_back(base, size);
      used.push_back(0);
      offset = 0;
    }
    void* ptr = blocks.back().first + offset;
    offset += need;
    used.back() = offset;
    total += need;
    return ptr;
  }
  ~BlockAllocator() {
    for (auto& block : blocks) free_fn(block.first);
  }
};

struct TableStore {
  uint32_t column_count = 0;
  std::vector<hyb_segment_desc> segments;                // chunk-major
  std::vector<std::vector<std::vector<int32_t>>> dates;  // [column][chunk] sorted day numbers (date columns only)
  std::vector<std::vector<std::string>> chars;           // [column][chunk] characters (1-char string columns)
  uint64_t rows = 0;
  uint64_t bytes = 0;
};

void run_parallel(size_t count, int threads, const std::function<void(size_t)>& body) {
  if (threads <= 1 || count <= 1) {
    for (size_t i = 0; i < count; ++i) body(i);
    return;
  }
  std::atomic<size_t> next{0};
  const auto worker = [&]() {
    while (true) {
      const size_t i = next.fetch_add(1);
      if (i >= count) return;
      body(i);
    }
  };
  std::vector<std::thread> pool;
  for (int t = 1; t < std::min<int>(threads, static_cast<int>(count)); ++t) pool.emplace_back(worker);
  worker();
  for (auto& thread : pool) thread.join();
}

// FixedWidthIntegerCompressor::_compress_using_max_value (fixed_width_integer_compressor.cpp:33-44)
int32_t compress_fixed_width(const uint32_t* in, uint32_t n, uint32_t max_value, BlockAllocator& memory, const void** out,
                             uint64_t* bytes) {
  if (max_value <= 0xFFu) {
    auto* data = static_cast<uint8_t*>(memory.allocate(n));
    for (uint32_t i = 0; i < n; ++i) data[i] = static_cast<uint8_t>(in[i]);
    *out = data;
    *bytes += n;
    return HYB_VEC_FIXED_1B;
  }
  if (max_value <= 0xFFFFu) {
    auto* data = static_cast<uint16_t*>(memory.allocate(size_t{n} * 2));
    for (uint32_t i = 0; i < n; ++i) data[i] = static_cast<uint16_t>(in[i]);
    *out = data;
    *bytes += size_t{n} * 2;
    return HYB_VEC_FIXED_2B;
  }
  auto* data = static_cast<uint32_t*>(memory.allocate(size_t{n} * 4));
  std::memcpy(data, in, size_t{n} * 4);
  *out = data;
  *bytes += size_t{n} * 4;
  return HYB_VEC_FIXED_4B;
}

// DictionaryEncoder for a float column.
hyb_segment_desc encode_float_dictionary(const float* values, uint32_t n, BlockAllocator& memory, uint64_t* bytes) {
  std::vector<float> dictionary(values, values + n);
  std::sort(dictionary.begin(), dictionary.end());
  dictionary.erase(std::unique(dictionary.begin(), dictionary.end()), dictionary.end());
  std::vector<uint32_t> ids(n);
  for (uint32_t i = 0; i < n; ++i) {
    ids[i] = static_cast<uint32_t>(std::lower_bound(dictionary.begin(), dictionary.end(), values[i]) - dictionary.begin());
  }
  hyb_segment_desc desc{};
  desc.encoding = HYB_ENC_DICTIONARY;
  desc.data_type = HYB_TYPE_FLOAT32;
  desc.row_count = n;
  desc.dictionary_size = static_cast<uint32_t>(dictionary.size());
  auto* stored = static_cast<float*>(memory.allocate(sizeof(float) * dictionary.size()));
  std::memcpy(stored, dictionary.data(), sizeof(float) * dictionary.size());
  *bytes += sizeof(float) * dictionary.size();
  desc.values = stored;
  desc.vector_type = compress_fixed_width(ids.data(), n, desc.dictionary_size, memory, &desc.attribute_vector, bytes);
  return desc;
}

// DictionaryEncoder for a small-domain column whose values are given as integer codes in [0, domain): the dictionary is
// the sorted set of present codes (dates as day numbers, chars as their byte).
hyb_segment_desc encode_code_dictionary(const uint16_t* codes, uint32_t n, uint32_t domain, BlockAllocator& memory,
                                        std::vector<int32_t>& dictionary_out, uint64_t* bytes) {
  std::vector<uint8_t> present(domain, 0);
  for (uint32_t i = 0; i < n; ++i) present[codes[i]] = 1;
  std::vector<uint32_t> id_of(domain, 0);
  dictionary_out.clear();
  for (uint32_t code = 0; code < domain; ++code) {
    if (present[code]) {
      id_of[code] = static_cast<uint32_t>(dictionary_out.size());
      dictionary_out.push_back(static_cast<int32_t>(code));
    }
  }
  std::vector<uint32_t> ids(n);
  for (uint32_t i = 0; i < n; ++i) ids[i] = id_of[codes[i]];
  hyb_segment_desc desc{};
  desc.encoding = HYB_ENC_DICTIONARY;
  desc.data_type = HYB_TYPE_STRING;
  desc.row_count = n;
  desc.dictionary_size = static_cast<uint32_t>(dictionary_out.size());
  desc.vector_type = compress_fixed_width(ids.data(), n, desc.dictionary_size, memory, &desc.attribute_vector, bytes);
  return desc;
}

// FrameOfReferenceEncoder (no NULLs).
hyb_segment_desc encode_frame_of_reference(const int32_t* values, uint32_t n, BlockAllocator& memory, uint64_t* bytes) {
  const uint32_t blocks = (n + HYB_FOR_BLOCK_SIZE - 1) / HYB_FOR_BLOCK_SIZE;
  auto* minima = static_cast<int32_t*>(memory.allocate(sizeof(int32_t) * blocks));
  std::vector<uint32_t> offsets(n);
  uint32_t max_offset = 0;
  for (uint32_t block = 0; block < blocks; ++block) {
    const uint32_t begin = block * HYB_FOR_BLOCK_SIZE, end = std::min(n, begin + HYB_FOR_BLOCK_SIZE);
    int32_t minimum = values[begin];
    for (uint32_t i = begin; i < end; ++i) minimum = std::min(minimum, values[i]);
    minima[block] = minimum;
    for (uint32_t i = begin; i < end; ++i) {
      offsets[i] = static_cast<uint32_t>(values[i]) - static_cast<uint32_t>(minimum);
      max_offset = std::max(max_offset, offsets[i]);
    }
  }
  *bytes += sizeof(int32_t) * blocks;
  hyb_segment_desc desc{};
  desc.encoding = HYB_ENC_FRAME_OF_REFERENCE;
  desc.data_type = HYB_TYPE_INT32;
  desc.row_count = n;
  desc.values = minima;
  desc.vector_type = compress_fixed_width(offsets.data(), n, max_offset, memory, &desc.attribute_vector, bytes);
  return desc;
}

inline int32_t order_key(uint64_t order_index) {  // dbgen sparse keys (build.c:135-146): keep 3 low bits, insert 2 zero bits
  return static_cast<int32_t>(((order_index >> 3) << 5) | (order_index & 7));
}

}  // namespace

struct hyb_tpch {
  BlockAllocator memory;
  TableStore lineitem;
  TableStore orders;
};

extern "C" {

int32_t hyb_tpch_day_number(int32_t year, int32_t month, int32_t day) {
  static const int32_t kDaysBefore[12] = {0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334};
  int32_t days = 0;
  for (int32_t y = 1992; y < year; ++y) days += (y % 4 == 0) ? 366 : 365;
  days += kDaysBefore[month - 1];
  if (month > 2 && year % 4 == 0) days += 1;
  return days + day - 1;
}

int hyb_tpch_generate(double scale_factor, uint64_t seed, int32_t threads, hyb_tpch_alloc_fn alloc,
                      hyb_tpch_free_fn free_fn, hyb_tpch** out) {
  return hyb_tpch_generate_shard(scale_factor, seed, 0, threads, alloc, free_fn, out);
}

int hyb_tpch_generate_shard(double scale_factor, uint64_t seed, uint64_t first_order, int32_t threads,
                            hyb_tpch_alloc_fn alloc, hyb_tpch_free_fn free_fn, hyb_tpch** out) {
  if (!out || scale_factor <= 0) return HYB_ERR_INVALID;
  auto* tables = new hyb_tpch{};
  tables->memory.alloc = alloc ? alloc : +[](size_t bytes) { return std::malloc(bytes); };
  tables->memory.free_fn = free_fn ? free_fn : +[](void* ptr) { std::free(ptr); };
  if (threads <= 0) threads = static_cast<int32_t>(std::max(1u, std::thread::hardware_concurrency()));
  const Rng rng{seed};
  const uint64_t order_count = static_cast<uint64_t>(std::llround(1'500'000.0 * scale_factor));
  const uint64_t part_count = std::max<uint64_t>(1, static_cast<uint64_t>(std::llround(200'000.0 * scale_factor)));

  // Lines per order and their prefix sums at block granularity.
  const auto lines_of = [&](uint64_t order) { return static_cast<uint32_t>(1 + rng.at(order + first_order, 7, 0) % 7); };
  const uint64_t block_count = (order_count + kOrderBlock - 1) / kOrderBlock;
  std::vector<uint64_t> block_rows(block_count + 1, 0);
  run_parallel(block_count, threads, [&](size_t block) {
    uint64_t rows = 0;
    const uint64_t begin = block * kOrderBlock + 1, end = std::min<uint64_t>(order_count, begin + kOrderBlock - 1);
    for (uint64_t order = begin; order <= end; ++order) rows += lines_of(order);
    block_rows[block + 1] = rows;
  });
  for (uint64_t block = 0; block < block_count; ++block) block_rows[block + 1] += block_rows[block];
  const uint64_t lineitem_rows = block_rows[block_count];

  // ---- lineitem ---------------------------------------------------------------------------------------------------
  auto& lineitem = tables->lineitem;
  lineitem.column_count = HYB_L_COLUMN_COUNT;
  lineitem.rows = lineitem_rows;
  const uint32_t lineitem_chunks = static_cast<uint32_t>((lineitem_rows + kChunkSize - 1) / kChunkSize);
  lineitem.segments.resize(size_t{lineitem_chunks} * HYB_L_COLUMN_COUNT);
  lineitem.dates.assign(HYB_L_COLUMN_COUNT, {});
  lineitem.dates[HYB_L_SHIPDATE].resize(lineitem_chunks);
  lineitem.chars.assign(HYB_L_COLUMN_COUNT, {});
  lineitem.chars[HYB_L_RETURNFLAG].resize(lineitem_chunks);
  lineitem.chars[HYB_L_LINESTATUS].resize(lineitem_chunks);
  std::atomic<uint64_t> lineitem_bytes{0};
  std::atomic<bool> failed{false};

  run_parallel(lineitem_chunks, threads, [&](size_t chunk) {
    const uint64_t row_begin = chunk * uint64_t{kChunkSize};
    const uint32_t n = static_cast<uint32_t>(std::min<uint64_t>(kChunkSize, lineitem_rows - row_begin));
    std::vector<int32_t> orderkey(n);
    std::vector<float> quantity(n), extendedprice(n), discount(n), tax(n);
    std::vector<uint16_t> returnflag(n), linestatus(n), shipdate(n);
    // Locate the order that owns row_begin.
    const uint64_t block = static_cast<uint64_t>(std::upper_bound(block_rows.begin(), block_rows.end(), row_begin) -
                                                 block_rows.begin()) - 1;
    uint64_t order = block * kOrderBlock + 1;
    uint64_t row = block_rows[block];
    while (row + lines_of(order) <= row_begin) {
      row += lines_of(order);
      ++order;
    }
    uint32_t filled = 0;
    while (filled < n) {
      const uint32_t lines = lines_of(order);
      const uint64_t global_order = order + first_order;
      const int32_t order_date = static_cast<int32_t>(rng.at(global_order, 7, 1) % kOrderDateRange);
      for (uint32_t line = 0; line < lines && filled < n; ++line, ++row) {
        if (row < row_begin) continue;
        const int32_t ship = order_date + 1 + static_cast<int32_t>(rng.at(global_order, line, 2) % 121);
        const uint32_t qty = 1 + static_cast<uint32_t>(rng.at(global_order, line, 3) % 50);
        const uint32_t disc = static_cast<uint32_t>(rng.at(global_order, line, 4) % 11);
        const uint32_t tx = static_cast<uint32_t>(rng.at(global_order, line, 5) % 9);
        const uint64_t partkey = 1 + rng.at(global_order, line, 6) % part_count;
        const int32_t receipt = ship + 1 + static_cast<int32_t>(rng.at(global_order, line, 8) % 30);
        const uint64_t retail_cents = 90000 + (partkey / 10) % 20001 + 100 * (partkey % 1000);  // dss.h retail price
        orderkey[filled] = order_key(global_order);
        quantity[filled] = static_cast<float>(qty);
        extendedprice[filled] = static_cast<float>(static_cast<double>(qty * retail_cents) / 100.0);
        discount[filled] = static_cast<float>(static_cast<double>(disc) / 100.0);
        tax[filled] = static_cast<float>(static_cast<double>(tx) / 100.0);
        returnflag[filled] = receipt <= kCutoffDay ? ((rng.at(global_order, line, 9) & 1) ? 'R' : 'A') : 'N';
        linestatus[filled] = ship <= kCutoffDay ? 'F' : 'O';
        shipdate[filled] = static_cast<uint16_t>(ship);
        ++filled;
      }
      ++order;
    }
    uint64_t bytes = 0;
    auto* descs = &lineitem.segments[chunk * HYB_L_COLUMN_COUNT];
    auto& memory = tables->memory;
    descs[HYB_L_ORDERKEY] = encode_frame_of_reference(orderkey.data(), n, memory, &bytes);
    descs[HYB_L_QUANTITY] = encode_float_dictionary(quantity.data(), n, memory, &bytes);
    descs[HYB_L_EXTENDEDPRICE] = encode_float_dictionary(extendedprice.data(), n, memory, &bytes);
    descs[HYB_L_DISCOUNT] = encode_float_dictionary(discount.data(), n, memory, &bytes);
    descs[HYB_L_TAX] = encode_float_dictionary(tax.data(), n, memory, &bytes);
    std::vector<int32_t> dictionary;
    for (const uint32_t column : {uint32_t{HYB_L_RETURNFLAG}, uint32_t{HYB_L_LINESTATUS}}) {
      const auto& codes = column == HYB_L_RETURNFLAG ? returnflag : linestatus;
      descs[column] = encode_code_dictionary(codes.data(), n, 256, memory, dictionary, &bytes);
      auto* group_codes = static_cast<uint64_t*>(memory.allocate(sizeof(uint64_t) * dictionary.size()));
      std::string characters;
      for (size_t i = 0; i < dictionary.size(); ++i) {
        group_codes[i] = 2 + static_cast<uint64_t>(dictionary[i]);  // aggregate_hash.cpp:876-878: 1-char string
        characters.push_back(static_cast<char>(dictionary[i]));
      }
      bytes += sizeof(uint64_t) * dictionary.size();
      descs[column].dictionary_codes = group_codes;
      lineitem.chars[column][chunk] = characters;
    }
    descs[HYB_L_SHIPDATE] = encode_code_dictionary(shipdate.data(), n, kOrderDateRange + 122, memory,
                                                   lineitem.dates[HYB_L_SHIPDATE][chunk], &bytes);
    for (uint32_t column = 0; column < HYB_L_COLUMN_COUNT; ++column) {
      if (!descs[column].attribute_vector) failed = true;
    }
    lineitem_bytes += bytes;
  });
  lineitem.bytes = lineitem_bytes;

  // ---- orders -----------------------------------------------------------------------------------------------------
  auto& orders = tables->orders;
  orders.column_count = HYB_O_COLUMN_COUNT;
  orders.rows = order_count;
  const uint32_t order_chunks = static_cast<uint32_t>((order_count + kChunkSize - 1) / kChunkSize);
  orders.segments.resize(size_t{order_chunks} * HYB_O_COLUMN_COUNT);
  orders.dates.assign(HYB_O_COLUMN_COUNT, {});
  orders.dates[HYB_O_ORDERDATE].resize(order_chunks);
  orders.chars.assign(HYB_O_COLUMN_COUNT, {});
  std::atomic<uint64_t> orders_bytes{0};
  run_parallel(order_chunks, threads, [&](size_t chunk) {
    const uint64_t first = chunk * uint64_t{kChunkSize} + 1;
    const uint32_t n = static_cast<uint32_t>(std::min<uint64_t>(kChunkSize, order_count - (first - 1)));
    auto* keys = static_cast<int32_t*>(tables->memory.allocate(sizeof(int32_t) * n));
    std::vector<uint16_t> dates(n);
    if (!keys) {
      failed = true;
      return;
    }
    for (uint32_t i = 0; i < n; ++i) {
      keys[i] = order_key(first + i + first_order);
      dates[i] = static_cast<uint16_t>(rng.at(first + i + first_order, 7, 1) % kOrderDateRange);
    }
    uint64_t bytes = sizeof(int32_t) * n;
    auto* descs = &orders.segments[chunk * HYB_O_COLUMN_COUNT];
    descs[HYB_O_ORDERKEY] = hyb_segment_desc{};
    descs[HYB_O_ORDERKEY].encoding = HYB_ENC_UNENCODED;
    descs[HYB_O_ORDERKEY].data_type = HYB_TYPE_INT32;
    descs[HYB_O_ORDERKEY].row_count = n;
    descs[HYB_O_ORDERKEY].values = keys;
    descs[HYB_O_ORDERDATE] =
        encode_code_dictionary(dates.data(), n, kOrderDateRange, tables->memory, orders.dates[HYB_O_ORDERDATE][chunk], &bytes);
    orders_bytes += bytes;
  });
  orders.bytes = orders_bytes;

  if (failed) {
    delete tables;
    return HYB_ERR_OOM;
  }
  *out = tables;
  return HYB_OK;
}

void hyb_tpch_free(hyb_tpch* tables) { delete tables; }

static int fill_view(const TableStore& store, hyb_table_view* out_view, uint64_t* out_rows) {
  if (!out_view) return HYB_ERR_INVALID;
  out_view->column_count = store.column_count;
  out_view->chunk_count = static_cast<uint32_t>(store.segments.size() / std::max<uint32_t>(store.column_count, 1));
  out_view->segments = store.segments.data();
  if (out_rows) *out_rows = store.rows;
  return HYB_OK;
}

int hyb_tpch_lineitem(const hyb_tpch* tables, hyb_table_view* out_view, uint64_t* out_rows) {
  return tables ? fill_view(tables->lineitem, out_view, out_rows) : HYB_ERR_INVALID;
}

int hyb_tpch_orders(const hyb_tpch* tables, hyb_table_view* out_view, uint64_t* out_rows) {
  return tables ? fill_view(tables->orders, out_view, out_rows) : HYB_ERR_INVALID;
}

int hyb_tpch_date_dictionary(const hyb_tpch* tables, int32_t table, uint32_t column, uint32_t chunk,
                             const int32_t** out_days, uint32_t* out_size) {
  if (!tables || !out_days || !out_size) return HYB_ERR_INVALID;
  const auto& store = table == 0 ? tables->lineitem : tables->orders;
  if (column >= store.dates.size() || chunk >= store.dates[column].size()) return HYB_ERR_INVALID;
  *out_days = store.dates[column][chunk].data();
  *out_size = static_cast<uint32_t>(store.dates[column][chunk].size());
  return HYB_OK;
}

int hyb_tpch_value_id_bounds(const hyb_tpch* tables, int32_t table, uint32_t column, const int32_t* day_numbers,
                             uint32_t value_count, uint32_t* out_bounds) {
  if (!tables || !day_numbers || !out_bounds) return HYB_ERR_INVALID;
  const auto& store = table == 0 ? tables->lineitem : tables->orders;
  if (column >= store.dates.size() || store.dates[column].empty()) return HYB_ERR_INVALID;
  const auto& dictionaries = store.dates[column];
  const auto search = [&](size_t begin, size_t end) {
    for (size_t chunk = begin; chunk < end; ++chunk) {
      const auto& dictionary = dictionaries[chunk];
      for (uint32_t v = 0; v < value_count; ++v) {
        const auto lower = std::lower_bound(dictionary.begin(), dictionary.end(), day_numbers[v]);
        const auto upper = std::upper_bound(dictionary.begin(), dictionary.end(), day_numbers[v]);
        uint32_t* out = out_bounds + (chunk * value_count + v) * 2;
        out[0] = lower == dictionary.end() ? HYB_INVALID_VALUE_ID : static_cast<uint32_t>(lower - dictionary.begin());
        out[1] = upper == dictionary.end() ? HYB_INVALID_VALUE_ID : static_cast<uint32_t>(upper - dictionary.begin());
      }
    }
  };
  // one dictionary search per chunk, like the reference's per-chunk scan jobs: spread over a few threads for big tables
  const size_t chunks = dictionaries.size();
  const size_t workers = chunks >= 2048 ? std::min<size_t>(8, std::max<size_t>(1, std::thread::hardware_concurrency())) : 1;
  if (workers == 1) {
    search(0, chunks);
  } else {
    std::vector<std::thread> threads;
    for (size_t w = 1; w < workers; ++w) threads.emplace_back(search, chunks * w / workers, chunks * (w + 1) / workers);
    search(0, chunks / workers);
    for (auto& thread : threads) thread.join();
  }
  return HYB_OK;
}

int hyb_tpch_char_dictionary(const hyb_tpch* tables, uint32_t column, uint32_t chunk, const char** out_chars,
                             uint32_t* out_size) {
  if (!tables || !out_chars || !out_size) return HYB_ERR_INVALID;
  const auto& store = tables->lineitem;
  if (column >= store.chars.size() || chunk >= store.chars[column].size()) return HYB_ERR_INVALID;
  *out_chars = store.chars[column][chunk].data();
  *out_size = static_cast<uint32_t>(store.chars[column][chunk].size());
  return HYB_OK;
}

int hyb_tpch_host_blocks(const hyb_tpch* tables, hyb_host_block* out_blocks, uint32_t* out_count) {
  if (!tables || !out_count) return HYB_ERR_INVALID;
  const auto& memory = tables->memory;
  if (out_blocks) {
    for (size_t index = 0; index < memory.blocks.size(); ++index) {
      out_blocks[index].base = memory.blocks[index].first;
      out_blocks[index].bytes = memory.used[index];
    }
  }
  *out_count = static_cast<uint32_t>(memory.blocks.size());
  return HYB_OK;
}

uint64_t hyb_tpch_table_bytes(const hyb_tpch* tables, int32_t table) {
  if (!tables) return 0;
  return table == 0 ? tables->lineitem.bytes : tables->orders.bytes;
}

}  // extern "C"
