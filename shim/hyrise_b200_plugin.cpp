#include "hyrise_b200_plugin.hpp"

#include <cstdlib>

#include "device_column_pool.hpp"
#include "gpu_operators.hpp"
#include "hyrise.hpp"
#include "utils/assert.hpp"

namespace hyrise {

std::string HyriseB200Plugin::description() const {
  return "B200 operators: TableScan / JoinHash / AggregateHash on libhyrise_b200.so";
}

void HyriseB200Plugin::start() {
  const auto* device = std::getenv("HYB_DEVICE");
  const auto status = hyb_context_create(device ? std::atoi(device) : 0, &_context);
  Assert(status == HYB_OK, hyb_last_error());  // no sm_100a GPU: fail loudly, there is no CPU fallback behind the ABI
  _pool = std::make_unique<DeviceColumnPool>(_context);
  set_gpu_column_pool(_pool.get());
}

void HyriseB200Plugin::stop() {
  set_gpu_column_pool(nullptr);
  _pool.reset();  // hyb_table_drop for every uploaded table
  if (_context) {
    hyb_context_destroy(_context);
    _context = nullptr;
  }
}

std::optional<PreBenchmarkHook> HyriseB200Plugin::pre_benchmark_hook() {
  return [this](AbstractBenchmarkItemRunner& /*benchmark_item_runner*/) {
    for (const auto& [name, table] : Hyrise::get().storage_manager.tables()) {
      if (_pool->upload(table)) {
        ++_uploaded_tables;
      }
    }
  };
}

std::optional<PostBenchmarkHook> HyriseB200Plugin::post_benchmark_hook() {
  return [this](nlohmann::json& report) {
    auto stats = hyb_operator_stats{};
    hyb_last_operator_stats(_context, &stats);
    report["hyrise_b200"] = {{"abi_version", hyb_abi_version()},
                             {"tables_on_device", _uploaded_tables},
                             {"last_operator_device_ms", stats.device_ms}};
  };
}

EXPORT_PLUGIN(HyriseB200Plugin);

}  // namespace hyrise
