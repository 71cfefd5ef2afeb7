// AggregateHash on the device (placeholder until the kernels land in this round).
#include "internal.hpp"
using namespace hyb;
extern "C" {
int hyb_aggregate_hash(hyb_context*, const hyb_aggregate_query*, hyb_aggregate_result_t*) {
  return fail(HYB_ERR_UNSUPPORTED, "hyb_aggregate_hash: not implemented yet");
}
int hyb_aggregate_result_info(hyb_context*, hyb_aggregate_result_t, uint64_t*, int32_t*) {
  return fail(HYB_ERR_NOT_FOUND, "unknown aggregate result");
}
int hyb_aggregate_result_row_ids(hyb_context*, hyb_aggregate_result_t, hyb_row_id*) {
  return fail(HYB_ERR_NOT_FOUND, "unknown aggregate result");
}
int hyb_aggregate_result_values(hyb_context*, hyb_aggregate_result_t, uint32_t, void*, uint8_t*, int32_t*) {
  return fail(HYB_ERR_NOT_FOUND, "unknown aggregate result");
}
int hyb_aggregate_result_free(hyb_context*, hyb_aggregate_result_t) {
  return fail(HYB_ERR_NOT_FOUND, "unknown aggregate result");
}
}
