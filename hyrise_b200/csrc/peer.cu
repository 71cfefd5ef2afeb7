// Peer groups (see peer.hpp): arena set-up over CUDA IPC, the device-side flag wait, phase statistics.
#include <cstring>

#include "device_utils.cuh"
#include "peer.hpp"

namespace hyb {

PeerGroup* find_peer_group(hyb_context* context, hyb_peer_group_t handle) {
  auto it = context->peer_groups.find(handle);
  return it == context->peer_groups.end() ? nullptr : it->second.get();
}

unsigned long long peer_next_epoch(PeerGroup* group) { return ++group->epoch; }

PeerGroup::~PeerGroup() {
  // hyb_peer_group_destroy is the orderly path (ranks synchronised by the caller); this covers context teardown
  for (uint32_t peer = 0; peer < world; ++peer) {
    if (peer != rank && peers[peer]) cudaIpcCloseMemHandle(peers[peer]);
  }
  if (own) cudaFree(own);
  if (d_arrivals) cudaFree(d_arrivals);
  if (h_counts) cudaFreeHost(h_counts);
  for (auto& event : events) {
    if (event) cudaEventDestroy(event);
  }
}

// One warp: lane s polls flag[s] (written by rank s with a system-scope fence in front) until it reaches `epoch`.
__global__ void peer_wait_kernel(const unsigned long long* flags, uint32_t world, unsigned long long epoch) {
  const uint32_t lane = threadIdx.x;
  if (lane < world) {
    while (ld_volatile_u64(flags + lane) < epoch) __nanosleep(100);
  }
  __syncwarp();
  __threadfence_system();  // what the peers stored before raising their flags is visible to the kernels that follow
}

int peer_wait(hyb_context* context, const unsigned long long* flags, uint32_t world, unsigned long long epoch) {
  peer_wait_kernel<<<1, 32, 0, context->stream>>>(flags, world, epoch);
  HYB_CUDA(cudaGetLastError());
  return HYB_OK;
}

}  // namespace hyb

using namespace hyb;

extern "C" {

int hyb_peer_group_create(hyb_context* context, uint32_t rank, uint32_t world, uint64_t tuple_capacity, void* out_ipc_handle,
                          hyb_peer_group_t* out_group) {
  HYB_CHECK(context && out_ipc_handle && out_group, HYB_ERR_INVALID, "NULL argument");
  HYB_CHECK(world >= 1 && world <= kPeerMax && (world & (world - 1)) == 0 && rank < world, HYB_ERR_INVALID,
            "world must be a power of two <= 16 and rank < world");
  DeviceGuard guard(context->device);
  std::lock_guard<std::mutex> lock(context->mutex);
  auto group = std::make_shared<PeerGroup>();
  group->owner = context;
  group->rank = rank;
  group->world = world;
  group->capacity = (tuple_capacity + 15) / 16 * 16;
  void* base = nullptr;
  HYB_CUDA(device_malloc_retry(context, reinterpret_cast<void**>(&base), group->arena_bytes()));
  group->own = static_cast<char*>(base);
  group->peers[rank] = group->own;
  cudaError_t error = cudaMemset(base, 0, kPeerControlBytes);
  cudaIpcMemHandle_t handle;
  if (error == cudaSuccess) error = cudaIpcGetMemHandle(&handle, base);
  if (error == cudaSuccess) error = cudaMalloc(reinterpret_cast<void**>(&group->d_arrivals), 64);
  if (error == cudaSuccess) error = cudaHostAlloc(reinterpret_cast<void**>(&group->h_counts), sizeof(PeerControl::counts) + sizeof(PeerControl::key_info), cudaHostAllocPortable);
  for (auto& event : group->events) {
    if (error == cudaSuccess) error = cudaEventCreate(&event);
  }
  if (error != cudaSuccess) {
    cudaFree(base);
    if (group->d_arrivals) cudaFree(group->d_arrivals);
    if (group->h_counts) cudaFreeHost(group->h_counts);
    HYB_CUDA(error);
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == HYB_IPC_HANDLE_BYTES, "IPC handle size");
  std::memcpy(out_ipc_handle, &handle, sizeof(handle));
  const auto id = context->next_handle++;
  context->peer_groups.emplace(id, std::move(group));
  *out_group = id;
  return HYB_OK;
}

int hyb_peer_group_connect(hyb_context* context, hyb_peer_group_t handle, const void* all_ipc_handles) {
  HYB_CHECK(context && all_ipc_handles, HYB_ERR_INVALID, "NULL argument");
  DeviceGuard guard(context->device);
  std::lock_guard<std::mutex> lock(context->mutex);
  auto* group = find_peer_group(context, handle);
  HYB_CHECK(group, HYB_ERR_NOT_FOUND, "unknown peer group handle");
  HYB_CHECK(!group->connected, HYB_ERR_INVALID, "peer group is already connected");
  for (uint32_t peer = 0; peer < group->world; ++peer) {
    if (peer == group->rank) continue;
    cudaIpcMemHandle_t ipc;
    std::memcpy(&ipc, static_cast<const char*>(all_ipc_handles) + size_t{peer} * HYB_IPC_HANDLE_BYTES, sizeof(ipc));
    void* mapped = nullptr;
    HYB_CUDA(cudaIpcOpenMemHandle(&mapped, ipc, cudaIpcMemLazyEnablePeerAccess));
    group->peers[peer] = static_cast<char*>(mapped);
  }
  group->connected = true;
  return HYB_OK;
}

int hyb_peer_group_destroy(hyb_context* context, hyb_peer_group_t handle) {
  HYB_CHECK(context, HYB_ERR_INVALID, "context is NULL");
  DeviceGuard guard(context->device);
  std::lock_guard<std::mutex> lock(context->mutex);
  auto it = context->peer_groups.find(handle);
  HYB_CHECK(it != context->peer_groups.end(), HYB_ERR_NOT_FOUND, "unknown peer group handle");
  PeerGroup* group = it->second.get();
  cudaStreamSynchronize(context->stream);
  for (int side = 0; side < 2; ++side) {
    if (group->received[side]) context->tables.erase(group->received[side]);
  }
  // The caller synchronises the ranks before destroying (nobody may still map or write this arena); ~PeerGroup unmaps
  // the peers and frees the arena.
  context->peer_groups.erase(it);
  return HYB_OK;
}

int hyb_peer_group_stats(hyb_context* context, hyb_peer_group_t handle, hyb_distributed_stats* out_stats) {
  HYB_CHECK(context && out_stats, HYB_ERR_INVALID, "NULL argument");
  DeviceGuard guard(context->device);
  std::lock_guard<std::mutex> lock(context->mutex);
  auto* group = find_peer_group(context, handle);
  HYB_CHECK(group, HYB_ERR_NOT_FOUND, "unknown peer group handle");
  HYB_CUDA(cudaStreamSynchronize(context->stream));
  const auto span = [&](int from, int to) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, group->events[from], group->events[to]) != cudaSuccess) {
      cudaGetLastError();
      ms = 0.f;
    }
    return ms;
  };
  hyb_distributed_stats stats = group->stats;
  stats.split_count_ms = span(0, 1);
  stats.count_wait_ms = span(1, 2);
  stats.push_ms = span(2, 3);
  stats.done_wait_ms = span(3, 4);
  stats.local_ms = span(4, 5);
  stats.finish_ms = span(5, 6);
  *out_stats = stats;
  return HYB_OK;
}

}  // extern "C"
