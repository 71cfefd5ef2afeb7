// Peer group: the ranks of one NVSwitch domain (one process per GPU) as seen from inside the library. Every rank owns an
// exchange arena — plain device memory exported with CUDA IPC and mapped by all peers — made of a control block (counts,
// key statistics, flags, partial-group blocks) and four tuple regions (build keys / RowIDs, probe keys / RowIDs). All
// cross-rank communication of a distributed operator goes through these arenas with NVLink peer stores issued by our own
// kernels; ordering between ranks is done with epoch flags (a monotonically increasing 64-bit counter per group, bumped
// once per distributed call — every rank makes the same calls in the same order), written after a system-scope fence and
// polled by a one-warp wait kernel on the consumer's stream. No NCCL collective and no host round trip other than the one
// that reads the count matrix is on the step's path.
#pragma once

#include "internal.hpp"

namespace hyb {

constexpr int kPeerMax = 16;
constexpr size_t kPeerAggregateBlockBytes = 32 * 1024;

struct PeerControl {
  unsigned long long counts[2][kPeerMax][kPeerMax];  // [side][source][destination] tuples; rank s writes row s everywhere
  long long key_info[kPeerMax][4];                   // per source, build side: {min, max, count of non-NULL keys, unused}
  unsigned long long count_flag[kPeerMax];           // epoch: source's counts / key_info are complete in THIS arena
  unsigned long long done_flag[kPeerMax];            // epoch: source's tuple stores into THIS arena have completed
  unsigned long long aggregate_flag[kPeerMax];       // epoch: source's partial-group block is complete in THIS arena
  unsigned long long bounds_flag[kPeerMax];          // epoch: source's side_bounds row is complete in THIS arena
  long long side_bounds[kPeerMax][8];                // per source: {build min, max, non-NULL keys, rows, probe min, max, keys, rows}
  unsigned char aggregate_block[kPeerMax][kPeerAggregateBlockBytes];
};
constexpr size_t kPeerControlBytes = (sizeof(PeerControl) + 4095) / 4096 * 4096;

struct PeerGroup {
  hyb_context* owner = nullptr;
  uint32_t rank = 0, world = 1;
  uint64_t capacity = 0;        // tuples per region
  char* own = nullptr;          // this rank's arena
  char* peers[kPeerMax] = {};   // arena of every rank as mapped into this process (peers[rank] == own)
  bool connected = false;
  unsigned long long epoch = 0;
  // persistent one-chunk tables over the receive regions (ValueSegment<int64> of keys), adopted once
  hyb_table_t received[2] = {0, 0};
  // device scratch that lives with the group
  unsigned int* d_arrivals = nullptr;      // CTA arrival counter of the push kernels
  unsigned long long* h_counts = nullptr;  // pinned copy of the local count matrix + key info
  hyb_distributed_stats stats{};
  cudaEvent_t events[8] = {};

  ~PeerGroup();
  PeerControl* control(uint32_t peer) const { return reinterpret_cast<PeerControl*>(peers[peer]); }
  char* region(uint32_t peer, int index) const { return peers[peer] + kPeerControlBytes + size_t(index) * region_bytes(); }
  size_t region_bytes() const { return (capacity * 8 + 64 + 255) / 256 * 256; }
  size_t arena_bytes() const { return kPeerControlBytes + 4 * region_bytes(); }
};

PeerGroup* find_peer_group(hyb_context* context, hyb_peer_group_t handle);
// Bumps the epoch (call once per distributed operator, with the context lock held).
unsigned long long peer_next_epoch(PeerGroup* group);
// Enqueue: spin (one warp) until flag[s] >= epoch for every source rank s.
int peer_wait(hyb_context* context, const unsigned long long* flags, uint32_t world, unsigned long long epoch);

}  // namespace hyb
