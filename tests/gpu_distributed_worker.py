"""torchrun worker: multi-GPU JoinHash / AggregateHash / TableScan through the C-ABI + NCCL, checked on rank 0 against
the oracle run over the union of all shards. Launched by tests/test_gpu_distributed.py."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))

import oracle_lib as orc  # noqa: E402
from helpers import row_ids_equal  # noqa: E402
from hyrise_b200 import capi  # noqa: E402
from hyrise_b200 import distributed as hd  # noqa: E402
from hyrise_b200.device import Aggregate, DeviceContext, Predicate  # noqa: E402
from hyrise_b200.tpch import (L_LINESTATUS, L_ORDERKEY, L_RETURNFLAG, L_SHIPDATE, O_ORDERKEY, TpchTables, _ViewHolder)  # noqa: E402
from test_oracle_aggregate import Q1_AGGREGATES  # noqa: E402

SF = 0.2


class UnionTable:
    """All ranks' shards as one table view (shards are deterministic, so every rank can rebuild all of them)."""

    def __init__(self, shards, which):
        tables = [getattr(shard, which) for shard in shards]
        self.column_definitions = tables[0].column_definitions
        self.column_count = tables[0].column_count
        self.chunk_count = sum(t.chunk_count for t in tables)
        self._descs = (capi.SegmentDesc * (self.chunk_count * self.column_count))()
        cursor = 0
        self._tables = tables
        self._chunk_owner = []
        for table in tables:
            for chunk in range(table.chunk_count):
                for column in range(self.column_count):
                    self._descs[cursor] = table.segment_desc(chunk, column)
                    cursor += 1
                self._chunk_owner.append((table, chunk))
        self._view = capi.TableView(self.chunk_count, self.column_count, C.cast(self._descs, C.POINTER(capi.SegmentDesc)))
        self.row_count = sum(t.row_count for t in tables)

    def view(self):
        return _ViewHolder(self._view)

    def string_value_id_bounds(self, predicate):
        return np.concatenate([t.string_value_id_bounds(predicate) for t in self._tables], axis=0)


def run_distributed_checks(device, rank, world, torch_device, sf=SF, legacy_paths=True):
    """Scan / JoinHash / Q1 AggregateHash over `world` ranks on shards of one SF-`sf` data set, checked on rank 0 against the
    oracle run over the union of all shards (bit-exact RowIDs in reference order; sums within 1e-6). Returns a summary line on
    rank 0; raises AssertionError on any rank that sees a mismatch (every rank takes part in every collective first)."""
    orders_per_rank = int(round(1_500_000 * sf))
    shard = TpchTables(sf, seed=42, first_order=rank * orders_per_rank)
    lineitem, orders = device.upload(shard.lineitem), device.upload(shard.orders)
    lineitem_base = hd.chunk_bases(shard.lineitem.chunk_count, torch_device)[rank]
    orders_base = hd.chunk_bases(shard.orders.chunk_count, torch_device)[rank]
    rows_per_rank = [None] * world
    dist.all_gather_object(rows_per_rank, shard.lineitem.row_count)
    position_base = sum(rows_per_rank[:rank])
    failures = []

    def expect(condition, message):
        if not condition:
            failures.append(message)

    if rank == 0:
        all_shards = [shard] + [TpchTables(sf, seed=42, first_order=r * orders_per_rank) for r in range(1, world)]
        global_lineitem, global_orders = UnionTable(all_shards, "lineitem"), UnionTable(all_shards, "orders")

    # ---- scan: no collective -----------------------------------------------------------------------------------------
    predicate = Predicate(L_SHIPDATE, capi.PRED_LESS_THAN, "1995-01-01")
    scan = device.table_scan(lineitem, predicate)
    gathered = [None] * world
    dist.all_gather_object(gathered, hd.globalize_row_ids(scan.to_host(), lineitem_base))
    scan.free()
    scan_rows = sum(len(g) for g in gathered)
    if rank == 0:
        expected = orc.table_scan(global_lineitem, predicate, threads=8)
        expect(row_ids_equal(np.concatenate(gathered), expected.row_ids), "distributed scan differs")

    # ---- join: native peer-group path (fused split + NVLink stores, device-side flags), then the legacy host paths --------
    radix_bits = 4
    expected = orc.join_hash(global_orders, O_ORDERKEY, global_lineitem, L_ORDERKEY, capi.JOIN_INNER, radix_bits, threads=8) \
        if rank == 0 else None
    group = hd.connect_peer_group(device, 2 * max(rows_per_rank) + 65_536)

    def check_join(kind, out_build, out_probe, offsets, colocated=False, want=None):
        want = want or expected
        gathered = [None] * world
        dist.all_gather_object(gathered, (out_build, out_probe, offsets))
        if rank != 0:
            return 0
        build_parts, probe_parts = [], []
        for partition in range(1 << radix_bits):
            if colocated:   # every rank holds its slice of every partition; rank order is global probe order
                owners = range(world)
            else:           # a partition lives on exactly one rank
                owners = [partition % world]
                for other in range(world):
                    if other != partition % world:
                        expect(gathered[other][2][partition] == gathered[other][2][partition + 1], f"{kind}: partition on a foreign rank")
            for owner in owners:
                owner_build, owner_probe, owner_offsets = gathered[owner]
                build_parts.append(owner_build[int(owner_offsets[partition]):int(owner_offsets[partition + 1])])
                probe_parts.append(owner_probe[int(owner_offsets[partition]):int(owner_offsets[partition + 1])])
        expect(row_ids_equal(np.concatenate(probe_parts), want.probe), f"distributed join ({kind}): probe RowIDs differ")
        expect(row_ids_equal(np.concatenate(build_parts), want.build), f"distributed join ({kind}): build RowIDs differ")
        return sum(len(p) for p in probe_parts)

    join_pairs = 0
    layouts = []
    for colocated_allowed in ("1", "0"):   # co-located shards: local joins, no exchange; then the radix exchange forced
        device.set_option("join_colocated", colocated_allowed)
        for repeat in range(2):   # twice: the arenas, the received tables and the epoch flags are reused
            result = group.join_hash(orders, O_ORDERKEY, lineitem, L_ORDERKEY, orders_base, lineitem_base, radix_bits)
            got_build, got_probe = result.to_host()
            offsets = result.partition_offsets()
            result.free()
            stats = group.stats()
            layouts.append(int(stats.colocated))
            join_pairs = check_join(f"peer group, colocated={stats.colocated}", got_build, got_probe, offsets, bool(stats.colocated))
    expect(layouts == [1, 1, 0, 0], f"co-location detection: {layouts}")
    device.set_option("join_colocated", "1")

    # shards that are NOT co-located: rank r probes with the lineitem shard of rank r + 1 -> the key ranges cross ranks, the
    # bounds exchange must send every rank down the radix-exchange path
    if world > 1:
        neighbour = TpchTables(sf, seed=42, first_order=((rank + 1) % world) * orders_per_rank)
        crossed = device.upload(neighbour.lineitem)
        crossed_base = hd.chunk_bases(neighbour.lineitem.chunk_count, torch_device)[rank]
        crossed_expected = None
        if rank == 0:
            rotated = [all_shards[(r + 1) % world] for r in range(world)]
            crossed_expected = orc.join_hash(global_orders, O_ORDERKEY, UnionTable(rotated, "lineitem"), L_ORDERKEY, capi.JOIN_INNER,
                                             radix_bits, threads=8)
        result = group.join_hash(orders, O_ORDERKEY, crossed, L_ORDERKEY, orders_base, crossed_base, radix_bits)
        got_build, got_probe = result.to_host()
        offsets = result.partition_offsets()
        result.free()
        crossed_stats = group.stats()
        expect(crossed_stats.colocated == 0, "crossed shards were taken for co-located ones")
        check_join("peer group, crossed shards", got_build, got_probe, offsets, False, crossed_expected)
        crossed.drop()
        neighbour.close()

    if legacy_paths and world > 1:
        peers = hd.PeerExchange(device, torch_device, capacity=2 * shard.lineitem.row_count + 65_536)
        for exchange_kind, exchange in (("nccl all-to-all", None), ("host-orchestrated P2P stores", peers)):
            pairs, offsets, build_rows, probe_rows, result = hd.device_distributed_join(
                device, orders, O_ORDERKEY, lineitem, L_ORDERKEY, radix_bits, orders_base, lineitem_base, torch_device, peers=exchange)
            if result is not None:
                got_build, got_probe = result.to_host()
                out_build = hd.unpack_row_ids(build_rows.cpu().numpy()[got_build["chunk_offset"]])
                out_probe = hd.unpack_row_ids(probe_rows.cpu().numpy()[got_probe["chunk_offset"]])
                result.free()
            else:
                out_build = out_probe = hd.unpack_row_ids(np.zeros(0, dtype=np.int64))
            check_join(exchange_kind, out_build, out_probe, offsets)
        peers.close()

    # ---- aggregate: Q1, partial groups through the peer arenas, every rank ends up with the complete result -----------------
    q1_predicates = [Predicate(L_SHIPDATE, capi.PRED_LESS_THAN_EQUALS, "1998-09-02")]
    output = group.aggregate_hash(lineitem, [L_RETURNFLAG, L_LINESTATUS], Q1_AGGREGATES, q1_predicates, lineitem_base, position_base)
    everyone = [None] * world
    dist.all_gather_object(everyone, (output.row_ids, output.values))
    if rank == 0:
        expected = orc.aggregate_hash(global_lineitem, [L_RETURNFLAG, L_LINESTATUS], Q1_AGGREGATES, predicates=q1_predicates)
        expect(output.group_count == expected.group_count == 4, f"Q1 groups {output.group_count}")
        expect(row_ids_equal(output.row_ids, expected.row_ids), "Q1: group order / representative RowIDs differ")
        for index in range(len(Q1_AGGREGATES)):
            if expected.values[index].dtype.kind == "f":
                expect(np.allclose(output.values[index], expected.values[index], rtol=1e-6, atol=0.0), f"Q1 aggregate {index}")
            else:
                expect(np.array_equal(output.values[index], expected.values[index]), f"Q1 aggregate {index}")
        for other_rows, other_values in everyone[1:]:   # replicated result: bit-identical on every rank
            expect(row_ids_equal(other_rows, output.row_ids), "Q1 result differs between ranks")
            expect(all(np.array_equal(a, b) for a, b in zip(other_values, output.values)), "Q1 values differ between ranks")
    # ---- aggregate, high cardinality (Q3 / Q18 shape): partial groups partitioned by key hash, every group merged by one rank --
    from hyrise_b200.device import Expression
    wide = [Aggregate(capi.AGG_SUM, Expression.column(1)), Aggregate(capi.AGG_AVG, Expression.column(2)),
            Aggregate(capi.AGG_MIN, Expression.column(2)), Aggregate(capi.AGG_MAX, Expression.column(3)),
            Aggregate(capi.AGG_COUNT_STAR)]
    owned = group.aggregate_hash(lineitem, [L_ORDERKEY, L_LINESTATUS], wide, [], lineitem_base, position_base)
    wide_stats = group.stats()
    parts = [None] * world
    dist.all_gather_object(parts, (owned.row_ids, owned.values, owned.nulls, int(wide_stats.aggregate_partitioned)))
    if rank == 0:
        expected = orc.aggregate_hash(global_lineitem, [L_ORDERKEY, L_LINESTATUS], wide)
        expect(all(part[3] == 1 for part in parts), "high-cardinality aggregate did not take the partitioned exchange")
        rows = np.concatenate([part[0] for part in parts])
        expect(len(rows) == expected.group_count, f"partitioned aggregate: {len(rows)} groups, expected {expected.group_count}")
        if len(rows) == expected.group_count:
            def ordered(row_ids):   # first-appearance order = ascending representative RowID
                return np.lexsort((row_ids["chunk_offset"], row_ids["chunk_id"]))
            got_order, want_order = ordered(rows), ordered(expected.row_ids)
            expect(row_ids_equal(rows[got_order], expected.row_ids[want_order]), "partitioned aggregate: representative RowIDs differ")
            expect(np.array_equal(want_order, np.arange(len(want_order))), "oracle groups are not in first-appearance order")
            for index in range(len(wide)):
                got = np.concatenate([part[1][index] for part in parts])[got_order]
                want = expected.values[index][want_order]
                if want.dtype.kind == "f":
                    expect(np.allclose(got, want, rtol=1e-6, atol=0.0), f"partitioned aggregate {index}")
                else:
                    expect(np.array_equal(got, want), f"partitioned aggregate {index}")
            # every group on exactly one rank, each rank's groups in first-appearance order
            for part in parts:
                expect(np.array_equal(ordered(part[0]), np.arange(len(part[0]))), "a rank's groups are not in first-appearance order")
    dist.barrier()
    group.destroy()
    lineitem.drop()
    orders.drop()
    shard.close()
    ok = [None] * world
    dist.all_gather_object(ok, failures)
    all_failures = [message for per_rank in ok for message in per_rank]
    assert not all_failures, "; ".join(all_failures)
    return (f"distributed OK on {world} GPUs: scan {scan_rows} RowIDs, join pairs {join_pairs} (co-located, forced exchange and "
            f"crossed shards; exchange: push {stats.push_ms:.3f} ms, local join {stats.local_ms:.3f} ms, {stats.nvlink_bytes} bytes "
            f"over NVLink from rank 0), Q1 groups {output.group_count}")


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch_device = torch.device("cuda", local_rank)
    device = DeviceContext(local_rank)
    summary = run_distributed_checks(device, rank, world, torch_device)
    if rank == 0:
        print(summary)
    dist.barrier()
    device.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
