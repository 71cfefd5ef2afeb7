#pragma once

#include <memory>
#include <optional>
#include <string>

#include "hyrise_b200.h"
#include "utils/abstract_plugin.hpp"

namespace hyrise {

class DeviceColumnPool;

// hyriseBenchmarkTPCH --plugins libhyriseB200Plugin.so (utils/plugin_manager.cpp:60-108 loads the library and calls factory()).
class HyriseB200Plugin : public AbstractPlugin {
 public:
  std::string description() const final;
  void start() final;  // hyb_context on the GPU named by HYB_DEVICE (default 0), empty column pool
  void stop() final;   // drops the device tables, destroys the context

  // After the benchmark's tables were generated and encoded (benchmarklib/benchmark_runner.cpp:175-182): upload every stored
  // table into the device column pool. From here on TableScan / JoinHash / AggregateHash find their inputs on the device.
  std::optional<PreBenchmarkHook> pre_benchmark_hook() final;
  // Adds {"hyrise_b200": {"device": ..., "tables_on_device": ...}} to the JSON report.
  std::optional<PostBenchmarkHook> post_benchmark_hook() final;

 private:
  hyb_context* _context{nullptr};
  std::unique_ptr<DeviceColumnPool> _pool;
  size_t _uploaded_tables{0};
};

}  // namespace hyrise
