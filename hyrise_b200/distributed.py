"""Multi-GPU host layer: one process per GPU (`torchrun`), `torch.distributed` for the plumbing (NCCL on GPUs, gloo in
the CPU tests). SURVEY.md §8e:

  * TableScan      chunks are independent (one JobTask per chunk in the reference, table_scan.cpp:119-232): every rank
                   scans the chunks it owns; NO collective. Global chunk ids = rank's chunk base + local chunk id.
  * AggregateHash  every rank pre-aggregates its shard (the device kernels), then the partial groups are exchanged by
                   group-key hash — ONE all-to-all — and merged by the owning rank. AVG is decomposed into SUM and COUNT
                   (the usual distributed rewrite); MIN/MAX/SUM/COUNT merge associatively; group order is restored from
                   the global first-row position of every group.
  * JoinHash       the reference already partitions both sides by hash(key) & mask and joins partitions 1:1
                   (join_hash_steps.hpp:509-617): partition p is owned by rank p % world. Both sides' {key, RowID}
                   tuples are exchanged with ONE all-to-all per side; each rank joins what it received.

The compute on each rank goes through a small engine interface so that the same exchange logic runs on the GPU (C-ABI
kernels) and, in the world_size-2 gloo tests, on the CPU oracle.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Sequence

import numpy as np
import torch
import torch.distributed as dist

from . import capi


def _world() -> tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def all_to_all_variable(parts: Sequence[torch.Tensor]) -> list[torch.Tensor]:
    """parts[d] goes to rank d (1-D or 2-D tensors with equal trailing shape/dtype). Returns what every rank sent us, in
    rank order. One count exchange + one payload all_to_all_single — the single exchange step of the path."""
    rank, world = _world()
    if world == 1:
        return [parts[0]]
    device = parts[0].device
    trailing = parts[0].shape[1:]
    send_counts = torch.tensor([part.shape[0] for part in parts], dtype=torch.int64, device=device)
    recv_counts = torch.empty_like(send_counts)
    dist.all_to_all_single(recv_counts, send_counts)
    send = torch.cat(list(parts), dim=0).contiguous()
    recv_list = [int(count) for count in recv_counts.tolist()]
    recv = torch.empty((sum(recv_list),) + tuple(trailing), dtype=send.dtype, device=device)
    width = int(np.prod(trailing)) if trailing else 1
    dist.all_to_all_single(recv.view(-1), send.view(-1), [count * width for count in recv_list],
                           [int(part.shape[0]) * width for part in parts])
    return list(torch.split(recv, recv_list, dim=0))


def splitmix64(values: np.ndarray) -> np.ndarray:
    x = values.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return x ^ (x >> np.uint64(31))


# ---------------------------------------------------------------------------------------------------------------------
# Scan
# ---------------------------------------------------------------------------------------------------------------------
def globalize_row_ids(row_ids: np.ndarray, chunk_base: int) -> np.ndarray:
    """Local RowIDs -> RowIDs of the global (all ranks) table: chunk ids are shifted by the rank's first chunk id."""
    out = row_ids.copy()
    valid = out["chunk_id"] != 0xFFFFFFFF
    out["chunk_id"][valid] += np.uint32(chunk_base)
    return out


def chunk_bases(local_chunk_count: int, device: torch.device) -> list[int]:
    """First global chunk id of every rank (ranks own consecutive chunk ranges)."""
    rank, world = _world()
    if world == 1:
        return [0]
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = local_chunk_count
    dist.all_reduce(counts)
    bases = np.concatenate([[0], np.cumsum(counts.cpu().numpy())[:-1]])
    return [int(b) for b in bases]


# ---------------------------------------------------------------------------------------------------------------------
# Aggregate
# ---------------------------------------------------------------------------------------------------------------------
@dataclass
class PartialGroups:
    """One rank's pre-aggregation result in mergeable form."""

    keys: np.ndarray            # int64 [groups, key columns]: chunk-independent encodings of the group-by values
    key_nulls: np.ndarray       # bool  [groups, key columns]
    first_position: np.ndarray  # int64 [groups]: global position of the first row of the group
    functions: list[int]        # decomposed functions (SUM / MIN / MAX / COUNT / COUNT_STAR only)
    values: list[np.ndarray]    # per decomposed aggregate: float64 or int64 [groups]
    counts: list[np.ndarray]    # per decomposed aggregate: non-NULL inputs, int64 [groups]


def decompose_aggregates(functions: Sequence[int]) -> tuple[list[tuple[int, int]], list[tuple[str, int, int]]]:
    """AVG(x) -> SUM(x), COUNT(x). Returns (local list of (function, original index), recipe per original aggregate:
    ("copy", i, -1) or ("avg", sum index, count index))."""
    local, recipe = [], []
    for index, function in enumerate(functions):
        if function == capi.AGG_AVG:
            recipe.append(("avg", len(local), len(local) + 1))
            local.append((capi.AGG_SUM, index))
            local.append((capi.AGG_COUNT, index))
        else:
            recipe.append(("copy", len(local), -1))
            local.append((function, index))
    return local, recipe


def merge_partial_groups(parts: Sequence[PartialGroups]) -> PartialGroups:
    """Merge partial groups with equal keys (what the owning rank does after the exchange)."""
    parts = [p for p in parts if p is not None]
    first = parts[0]
    keys = np.concatenate([p.keys for p in parts], axis=0)
    key_nulls = np.concatenate([p.key_nulls for p in parts], axis=0)
    positions = np.concatenate([p.first_position for p in parts])
    if len(keys) == 0:
        return first
    combined = np.concatenate([np.where(key_nulls, 0, keys), key_nulls.astype(np.int64)], axis=1)
    unique, inverse = np.unique(combined, axis=0, return_inverse=True)
    inverse = inverse.reshape(-1)
    groups = len(unique)
    width = keys.shape[1]
    out_positions = np.full(groups, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(out_positions, inverse, positions)
    values, counts = [], []
    for index, function in enumerate(first.functions):
        part_values = np.concatenate([p.values[index] for p in parts])
        part_counts = np.concatenate([p.counts[index] for p in parts])
        merged_counts = np.zeros(groups, dtype=np.int64)
        np.add.at(merged_counts, inverse, part_counts)
        if function in (capi.AGG_SUM, capi.AGG_COUNT, capi.AGG_COUNT_STAR):
            merged = np.zeros(groups, dtype=part_values.dtype)
            np.add.at(merged, inverse, np.where(part_counts > 0, part_values, 0) if function == capi.AGG_SUM else part_values)
        elif function == capi.AGG_MIN:
            big = np.inf if part_values.dtype.kind == "f" else np.iinfo(part_values.dtype).max
            merged = np.full(groups, big, dtype=part_values.dtype)
            np.minimum.at(merged, inverse, np.where(part_counts > 0, part_values, big))
        else:
            small = -np.inf if part_values.dtype.kind == "f" else np.iinfo(part_values.dtype).min
            merged = np.full(groups, small, dtype=part_values.dtype)
            np.maximum.at(merged, inverse, np.where(part_counts > 0, part_values, small))
        values.append(merged)
        counts.append(merged_counts)
    return PartialGroups(unique[:, :width], unique[:, width:].astype(bool), out_positions, list(first.functions), values, counts)


def _pack(partial: PartialGroups) -> torch.Tensor:
    """[groups, columns] float64-viewable int64 matrix for the exchange."""
    columns = [partial.keys, partial.key_nulls.astype(np.int64), partial.first_position[:, None]]
    for values, counts in zip(partial.values, partial.counts):
        columns.append(values.view(np.int64)[:, None] if values.dtype == np.float64 else values.astype(np.int64)[:, None])
        columns.append(counts[:, None])
    matrix = np.concatenate(columns, axis=1) if len(partial.keys) else np.zeros((0, sum(c.shape[1] for c in columns)), dtype=np.int64)
    return torch.from_numpy(np.ascontiguousarray(matrix, dtype=np.int64))


def _unpack(matrix: np.ndarray, template: PartialGroups) -> PartialGroups:
    width = template.keys.shape[1]
    keys = matrix[:, :width]
    nulls = matrix[:, width:2 * width].astype(bool)
    positions = matrix[:, 2 * width]
    values, counts = [], []
    cursor = 2 * width + 1
    for original in template.values:
        column = matrix[:, cursor]
        values.append(column.view(np.float64).copy() if original.dtype == np.float64 else column.copy())
        counts.append(matrix[:, cursor + 1].copy())
        cursor += 2
    return PartialGroups(keys.copy(), nulls, positions.copy(), list(template.functions), values, counts)


def exchange_and_merge(partial: PartialGroups, device: torch.device) -> PartialGroups:
    """The aggregate's exchange step: partial groups go to rank hash(key) % world; the owner merges them."""
    rank, world = _world()
    if world == 1:
        return merge_partial_groups([partial])
    matrix = _pack(partial)
    width = partial.keys.shape[1]
    if len(partial.keys):
        mixed = np.zeros(len(partial.keys), dtype=np.uint64)
        for column in range(width):
            mixed = splitmix64(mixed ^ np.where(partial.key_nulls[:, column], 0, partial.keys[:, column]).astype(np.uint64))
        owner = (mixed % np.uint64(world)).astype(np.int64)
    else:
        owner = np.zeros(0, dtype=np.int64)
    parts = [matrix[torch.from_numpy(np.flatnonzero(owner == d))].to(device) for d in range(world)]
    received = all_to_all_variable(parts)
    merged = torch.cat(received, dim=0).cpu().numpy()
    return merge_partial_groups([_unpack(merged, partial)])


def gather_groups(partial: PartialGroups, device: torch.device) -> PartialGroups | None:
    """All owners' (disjoint) groups on rank 0, ordered by first appearance in the global row order."""
    rank, world = _world()
    matrix = _pack(partial).to(device)
    if world > 1:
        parts = [matrix if d == 0 else matrix[:0] for d in range(world)]
        received = all_to_all_variable(parts)
        if rank != 0:
            return None
        matrix = torch.cat(received, dim=0)
    merged = _unpack(matrix.cpu().numpy(), partial)
    order = np.argsort(merged.first_position, kind="stable")
    return PartialGroups(merged.keys[order], merged.key_nulls[order], merged.first_position[order], merged.functions,
                         [v[order] for v in merged.values], [c[order] for c in merged.counts])


def finalize(partial: PartialGroups, recipe) -> tuple[list[np.ndarray], list[np.ndarray]]:
    """Values and NULL masks of the ORIGINAL aggregates (AVG = SUM / COUNT, NULL when no non-NULL input)."""
    values, nulls = [], []
    for kind, a, b in recipe:
        if kind == "avg":
            count = partial.values[b].astype(np.float64)
            with np.errstate(divide="ignore", invalid="ignore"):
                values.append(partial.values[a].astype(np.float64) / count)
            nulls.append(partial.values[b] == 0)
        else:
            values.append(partial.values[a])
            function = partial.functions[a]
            nulls.append(np.zeros(len(partial.keys), dtype=bool) if function in (capi.AGG_COUNT, capi.AGG_COUNT_STAR)
                         else partial.counts[a] == 0)
    return values, nulls


def distributed_aggregate(local_aggregate: Callable[[list[int]], PartialGroups], functions: Sequence[int],
                          device: torch.device):
    """`local_aggregate(decomposed functions + original indexes)` runs the rank's pre-aggregation; returns on rank 0 the
    merged groups (PartialGroups) and the finalized (values, nulls) of the original aggregates, None elsewhere."""
    local, recipe = decompose_aggregates(functions)
    partial = local_aggregate(local)
    owned = exchange_and_merge(partial, device)
    merged = gather_groups(owned, device)
    if merged is None:
        return None
    return merged, finalize(merged, recipe)


# ---------------------------------------------------------------------------------------------------------------------
# Join: radix exchange of {key, RowID} tuples
# ---------------------------------------------------------------------------------------------------------------------
def partition_owner(keys: torch.Tensor, world: int) -> torch.Tensor:
    """Rank that owns a key's radix partition: hash(key) & mask is the identity hash's low bits (join_hash_steps.hpp
    :345-395), partition p belongs to rank p % world (world a power of two <= 2^radix_bits)."""
    return keys & (world - 1)


def exchange_tuples(keys: torch.Tensor, row_ids: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """keys int64 [n], row_ids int64 [n] (packed chunk_id | chunk_offset << 32, global chunk ids). Stable partition by
    owner, one all-to-all, concatenation in source-rank order — which keeps global row order inside every partition."""
    rank, world = _world()
    if world == 1:
        return keys, row_ids
    assert world & (world - 1) == 0, "the radix exchange needs a power-of-two world size"
    owner = partition_owner(keys, world)
    order = torch.argsort(owner, stable=True)
    counts = torch.bincount(owner, minlength=world).tolist()
    payload = torch.stack([keys[order], row_ids[order]], dim=1)
    parts = list(torch.split(payload, counts, dim=0))
    received = torch.cat(all_to_all_variable(parts), dim=0)
    return received[:, 0].contiguous(), received[:, 1].contiguous()


def pack_row_ids(row_ids: np.ndarray) -> np.ndarray:
    return row_ids["chunk_id"].astype(np.int64) | (row_ids["chunk_offset"].astype(np.int64) << 32)


def unpack_row_ids(packed: np.ndarray) -> np.ndarray:
    from .device import ROW_ID_DTYPE
    out = np.empty(len(packed), dtype=ROW_ID_DTYPE)
    out["chunk_id"] = (packed & 0xFFFFFFFF).astype(np.uint32)
    out["chunk_offset"] = ((packed >> 32) & 0xFFFFFFFF).astype(np.uint32)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# GPU engine glue: the per-rank compute through the C-ABI (torch tensors only as device buffers for the exchange)
# ---------------------------------------------------------------------------------------------------------------------
def device_materialize_side(device_context, table, column_id: int, chunk_id_base: int, torch_device: torch.device):
    """hyb_join_materialize into torch-allocated device buffers; NULL keys dropped. Returns (keys, packed RowIDs)."""
    import ctypes as C

    side = capi.JoinSide(table.handle, column_id, 0)
    positions = C.c_uint64()
    capi.check(device_context.lib.hyb_join_side_positions(device_context.ptr, C.byref(side), C.byref(positions)))
    n = positions.value
    keys = torch.empty(n + 8, dtype=torch.int64, device=torch_device)
    row_ids = torch.empty(n + 8, dtype=torch.int64, device=torch_device)
    torch.cuda.synchronize(torch_device)
    capi.check(device_context.lib.hyb_join_materialize(device_context.ptr, C.byref(side), chunk_id_base, keys.data_ptr(),
                                                       row_ids.data_ptr()))
    keys, row_ids = keys[:n], row_ids[:n]
    keep = row_ids >= 0
    return keys[keep], row_ids[keep]


def device_partition_side(device_context, table, column_id: int, chunk_id_base: int, world: int, torch_device: torch.device,
                          filter_handle: int = 0):
    """hyb_join_partition: the side's non-NULL {key, global RowID} tuples grouped by owner rank (key & (world - 1)), stable.
    Returns (keys, packed RowIDs, offsets) — device tensors and world + 1 host offsets; group d is [offsets[d], offsets[d+1])."""
    import ctypes as C

    side = capi.JoinSide(table.handle, column_id, filter_handle)
    positions = C.c_uint64()
    capi.check(device_context.lib.hyb_join_side_positions(device_context.ptr, C.byref(side), C.byref(positions)))
    n = positions.value
    keys = torch.empty(n + 8, dtype=torch.int64, device=torch_device)
    row_ids = torch.empty(n + 8, dtype=torch.int64, device=torch_device)
    torch.cuda.current_stream(torch_device).synchronize()  # the buffers' previous users ran on torch's stream
    offsets = (C.c_uint64 * (world + 1))()
    capi.check(device_context.lib.hyb_join_partition(device_context.ptr, C.byref(side), world, chunk_id_base, keys.data_ptr(),
                                                     row_ids.data_ptr(), offsets))
    return keys, row_ids, [int(v) for v in offsets]


def exchange_partitioned(sides: Sequence[tuple[torch.Tensor, torch.Tensor, list[int]]]):
    """The exchange step of the radix join: for every side (keys, RowIDs, offsets) group d goes to rank d. ONE count
    all-to-all for all sides, then one payload all-to-all per array, written straight into buffers that
    hyb_table_append_chunk_device can adopt (8 spare elements after the last key). Returns [(keys, RowIDs, count)]."""
    rank, world = _world()
    device = sides[0][0].device
    send_counts = [[offsets[d + 1] - offsets[d] for d in range(world)] for _, _, offsets in sides]
    if world == 1:
        return [(keys, row_ids, offsets[-1]) for keys, row_ids, offsets in sides]
    send = torch.tensor(send_counts, dtype=torch.int64, device=device).t().contiguous()  # [destination][side]
    recv = torch.empty_like(send)
    dist.all_to_all_single(recv, send)
    recv_counts = recv.t().tolist()  # [side][source]
    out = []
    for index, (keys, row_ids, offsets) in enumerate(sides):
        total = sum(recv_counts[index])
        recv_keys = torch.empty(total + 8, dtype=torch.int64, device=device)
        recv_rows = torch.empty(total + 8, dtype=torch.int64, device=device)
        dist.all_to_all_single(recv_keys[:total], keys[:offsets[-1]], recv_counts[index], send_counts[index])
        dist.all_to_all_single(recv_rows[:total], row_ids[:offsets[-1]], recv_counts[index], send_counts[index])
        out.append((recv_keys, recv_rows, total))
    return out


class DeviceTupleTable:
    """Received {key, RowID} tuples as a one-chunk device table (ValueSegment<int64> of keys) for hyb_join_hash. `keys`
    must have 8 spare elements after the first `count` (Arena::kTailPad contract of hyb_table_append_chunk_device)."""

    def __init__(self, device_context, keys: torch.Tensor, row_ids: torch.Tensor, count: int | None = None):
        import ctypes as C
        from .device import DeviceTable

        self.device_context = device_context
        if count is None:  # unpadded input: copy into a padded buffer
            count = int(keys.shape[0])
            padded = torch.empty(count + 8, dtype=torch.int64, device=keys.device)
            padded[:count] = keys
            keys = padded
        self.keys = keys
        self.row_ids = row_ids[:count]
        self.count = count
        handle = C.c_uint64()
        capi.check(device_context.lib.hyb_table_create(device_context.ptr, 1, C.byref(handle)))
        if count:
            desc = capi.SegmentDesc()
            desc.encoding = capi.ENC_UNENCODED
            desc.data_type = capi.TYPE_INT64
            desc.row_count = count
            desc.values = self.keys.data_ptr()
            torch.cuda.current_stream(keys.device).synchronize()  # NCCL wrote the buffer on torch's stream
            capi.check(device_context.lib.hyb_table_append_chunk_device(device_context.ptr, handle.value, C.byref(desc)))
        self.table = DeviceTable(device_context, handle.value, None, None)

    def drop(self) -> None:
        self.table.drop()


class _DeviceArray:
    """Zero-copy torch view of library-owned device memory (torch.as_tensor reads __cuda_array_interface__)."""

    def __init__(self, pointer: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<i8", "data": (pointer, False), "version": 2}


class PeerExchange:
    """The radix exchange without a send buffer: every rank owns a receive arena (plain device memory exported through
    CUDA IPC and mapped by all peers); hyb_join_partition_push's ranked-write kernel stores each group straight into its
    owner's arena over NVLink. Collectives left: one world^2-integer all-gather of counts per side and one barrier.
    Arena layout: four regions of `capacity` int64 — build keys, build RowIDs, probe keys, probe RowIDs."""

    REGIONS = 4

    def __init__(self, device_context, torch_device: torch.device, capacity: int):
        import ctypes as C

        self.device_context = device_context
        self.torch_device = torch_device
        self.rank, self.world = _world()
        assert self.world & (self.world - 1) == 0 and self.world <= 16
        # every rank must lay its arena out identically: peers compute region addresses inside OUR arena with this value
        agreed = torch.tensor([capacity], dtype=torch.int64, device=torch_device)
        dist.all_reduce(agreed, op=dist.ReduceOp.MAX)
        self.capacity = (int(agreed.item()) + 8 + 15) // 16 * 16
        lib = device_context.lib
        # Every rank takes part in every collective below even when a local step failed: a rank that raised early would
        # leave the others waiting forever. Failures are agreed on first, then raised everywhere.
        own = C.c_void_p()
        handle = (C.c_ubyte * capi.IPC_HANDLE_BYTES)()
        status = lib.hyb_exchange_arena_create(device_context.ptr, self.REGIONS * self.capacity * 8, C.byref(own), handle)
        self.own = own.value if status == capi.HYB_OK else None
        exported: list = [None] * self.world
        dist.all_gather_object(exported, bytes(handle) if status == capi.HYB_OK else None)
        self.bases = []
        opened = True
        if all(raw is not None for raw in exported):
            for peer, raw in enumerate(exported):
                if peer == self.rank:
                    self.bases.append(self.own)
                    continue
                mapped = C.c_void_p()
                buffer = (C.c_ubyte * capi.IPC_HANDLE_BYTES).from_buffer_copy(raw)
                if lib.hyb_exchange_arena_open(device_context.ptr, buffer, C.byref(mapped)) != capi.HYB_OK:
                    opened = False
                    break
                self.bases.append(mapped.value)
        else:
            opened = False
        agreed_open: list = [None] * self.world
        dist.all_gather_object(agreed_open, opened)
        if not all(agreed_open):
            message = lib.hyb_last_error().decode("utf-8", "replace")
            self._release()
            raise RuntimeError(f"peer exchange unavailable on at least one rank (this rank: {message or 'ok'})")
        self._counts = torch.empty(self.world, dtype=torch.int64, device=torch_device)
        self._matrix = torch.empty(self.world * self.world, dtype=torch.int64, device=torch_device)
        self._error = None
        self._received = 0

    def region(self, peer: int, region: int) -> int:
        return self.bases[peer] + region * self.capacity * 8

    def push_side(self, table, column_id: int, chunk_id_base: int, first_region: int, filter_handle: int = 0) -> int:
        """Split one join side and store it into the owners' arenas (regions first_region = keys, first_region + 1 =
        RowIDs). Returns the number of tuples THIS rank receives for the side (complete after `barrier`)."""
        import ctypes as C

        def exchange(_user, counts, dest_keys, dest_rows):
            try:
                self._counts.copy_(torch.tensor([int(counts[d]) for d in range(self.world)], dtype=torch.int64))
                dist.all_gather_into_tensor(self._matrix, self._counts)
                matrix = self._matrix.view(self.world, self.world).tolist()  # [source][destination]
                for d in range(self.world):
                    before = sum(matrix[s][d] for s in range(self.rank))
                    total = sum(matrix[s][d] for s in range(self.world))
                    if total + 8 > self.capacity:
                        raise RuntimeError(f"receive arena of rank {d} too small: {total} tuples > capacity {self.capacity - 8}")
                    dest_keys[d] = self.region(d, first_region) + before * 8
                    dest_rows[d] = self.region(d, first_region + 1) + before * 8
                self._received = sum(matrix[s][self.rank] for s in range(self.world))
                return 0
            except BaseException as error:  # noqa: BLE001 - must not unwind through the C frame
                self._error = error
                return 1

        callback = capi.EXCHANGE_FN(exchange)
        side = capi.JoinSide(table.handle, column_id, filter_handle)
        self._error = None
        status = self.device_context.lib.hyb_join_partition_push(self.device_context.ptr, C.byref(side), self.world,
                                                                 chunk_id_base, callback, None)
        if self._error is not None:
            raise self._error
        capi.check(status)
        return self._received

    def barrier(self) -> None:
        dist.barrier(device_ids=[self.torch_device.index])

    def received(self, first_region: int, count: int) -> tuple[torch.Tensor, torch.Tensor]:
        """(keys with 8 spare elements, RowIDs) of what this rank received, as zero-copy torch views of its arena."""
        keys = torch.as_tensor(_DeviceArray(self.region(self.rank, first_region), count + 8), device=self.torch_device)
        rows = torch.as_tensor(_DeviceArray(self.region(self.rank, first_region + 1), max(count, 1)), device=self.torch_device)
        return keys, rows[:count]

    def _release(self) -> None:
        lib = self.device_context.lib
        for peer, base in enumerate(self.bases):
            if peer != self.rank and base:
                lib.hyb_exchange_arena_close(self.device_context.ptr, base)
        self.bases = []
        dist.barrier(device_ids=[self.torch_device.index])  # nobody still maps the arena we are about to free
        if self.own:
            lib.hyb_exchange_arena_destroy(self.device_context.ptr, self.own)
            self.own = None

    def close(self) -> None:
        self._release()


def connect_peer_group(device_context, tuple_capacity: int):
    """Creates this rank's hyb_peer_group, all-gathers the CUDA IPC handles (the one collective of the native multi-GPU
    path) and maps every peer's arena. Every rank takes part in every collective even when a local step failed, so that
    nobody hangs; the failure is agreed on and raised everywhere."""
    from .device import DevicePeerGroup

    rank, world = _world()
    agreed = [None] * world
    if world > 1:
        dist.all_gather_object(agreed, int(tuple_capacity))
        tuple_capacity = max(agreed)
    group, error = None, None
    try:
        group = DevicePeerGroup(device_context, rank, world, tuple_capacity)
    except Exception as exception:  # noqa: BLE001
        error = exception
    handles = [group.ipc_handle if group else None]
    if world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, group.ipc_handle if group else None)
    if all(handle is not None for handle in handles):
        try:
            if world > 1:
                group.connect(handles)
        except Exception as exception:  # noqa: BLE001
            error = exception
    elif error is None:
        error = RuntimeError("a peer could not create its exchange arena")
    if world > 1:
        states = [None] * world
        dist.all_gather_object(states, error is None)
        if not all(states) and error is None:
            error = RuntimeError("a peer could not map the exchange arenas")
    if error is not None:
        raise error
    return group


def device_distributed_join(device_context, build_table, build_column: int, probe_table, probe_column: int,
                            radix_bits: int, build_chunk_base: int, probe_chunk_base: int, torch_device: torch.device,
                            peers: "PeerExchange | None" = None):
    """Inner JoinHash across ranks, host-orchestrated (the legacy paths next to DevicePeerGroup.join_hash): materialise both
    sides, ONE all-to-all per side (or the P2P push of `peers`), join the received tuples locally.
    Returns the 5-tuple (pair count on this rank, partition offsets, build RowIDs, probe RowIDs, join result or None): the
    rank's part of the reference-ordered result — partitions p with p % world == rank, hence world <= 2^radix_bits. The
    join result's RowIDs index the two GLOBAL-RowID arrays. With `peers` those arrays are zero-copy views of this rank's
    receive arena: they are valid until the next push into the arena (clone them to keep them longer)."""
    _, world = _world()
    assert world & (world - 1) == 0, "the radix exchange needs a power-of-two world size"
    assert world <= (1 << radix_bits), "every rank must own a partition: world <= 2^radix_bits"
    if peers is not None and world > 1:
        # fused split + NVLink stores into the owners' arenas; collectives: 2 count all-gathers + 1 barrier
        build_count = peers.push_side(build_table, build_column, build_chunk_base, 0)
        probe_count = peers.push_side(probe_table, probe_column, probe_chunk_base, 2)
        peers.barrier()
        build_keys, build_rows = peers.received(0, build_count)
        probe_keys, probe_rows = peers.received(2, probe_count)
    else:
        sides = [device_partition_side(device_context, build_table, build_column, build_chunk_base, world, torch_device),
                 device_partition_side(device_context, probe_table, probe_column, probe_chunk_base, world, torch_device)]
        (build_keys, build_rows, build_count), (probe_keys, probe_rows, probe_count) = exchange_partitioned(sides)
    build = DeviceTupleTable(device_context, build_keys, build_rows, build_count)
    probe = DeviceTupleTable(device_context, probe_keys, probe_rows, probe_count)
    try:
        if probe.count == 0 or build.count == 0:
            empty = np.zeros(0, dtype=np.int64)
            return 0, np.zeros((1 << radix_bits) + 1, dtype=np.uint64), empty, empty, None
        result = device_context.join_hash(build.table, 0, probe.table, 0, capi.JOIN_INNER, radix_bits)
        return result.info()[0], result.partition_offsets(), build.row_ids, probe.row_ids, result
    finally:
        build.drop()
        probe.drop()
