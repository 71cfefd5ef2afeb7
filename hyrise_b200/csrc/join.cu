// JoinHash on the device (placeholder until the kernels land in this round).
#include "internal.hpp"
using namespace hyb;
extern "C" {
int hyb_join_hash(hyb_context*, const hyb_join_side*, const hyb_join_side*, int32_t, int32_t, hyb_join_result_t*) {
  return fail(HYB_ERR_UNSUPPORTED, "hyb_join_hash: not implemented yet");
}
int hyb_join_result_info(hyb_context*, hyb_join_result_t, uint64_t*, uint32_t*, int32_t*) {
  return fail(HYB_ERR_NOT_FOUND, "unknown join result");
}
int hyb_join_result_partition_offsets(hyb_context*, hyb_join_result_t, uint64_t*) {
  return fail(HYB_ERR_NOT_FOUND, "unknown join result");
}
int hyb_join_result_copy(hyb_context*, hyb_join_result_t, uint64_t, uint64_t, hyb_row_id*, hyb_row_id*) {
  return fail(HYB_ERR_NOT_FOUND, "unknown join result");
}
int hyb_join_result_free(hyb_context*, hyb_join_result_t) { return fail(HYB_ERR_NOT_FOUND, "unknown join result"); }
int hyb_join_result_device_ptrs(hyb_context*, hyb_join_result_t, void**, void**) {
  return fail(HYB_ERR_NOT_FOUND, "unknown join result");
}
}
