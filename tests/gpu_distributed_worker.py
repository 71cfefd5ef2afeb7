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


def main():
    rank, world, local_rank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch_device = torch.device("cuda", local_rank)
    orders_per_rank = int(round(1_500_000 * SF))
    shard = TpchTables(SF, seed=42, first_order=rank * orders_per_rank)
    device = DeviceContext(local_rank)
    lineitem, orders = device.upload(shard.lineitem), device.upload(shard.orders)
    lineitem_base = hd.chunk_bases(shard.lineitem.chunk_count, torch_device)[rank]
    orders_base = hd.chunk_bases(shard.orders.chunk_count, torch_device)[rank]

    if rank == 0:
        all_shards = [shard] + [TpchTables(SF, seed=42, first_order=r * orders_per_rank) for r in range(1, world)]
        global_lineitem, global_orders = UnionTable(all_shards, "lineitem"), UnionTable(all_shards, "orders")

    # ---- scan: no collective -----------------------------------------------------------------------------------------
    predicate = Predicate(L_SHIPDATE, capi.PRED_LESS_THAN, "1995-01-01")
    scan = device.table_scan(lineitem, predicate)
    gathered = [None] * world
    dist.all_gather_object(gathered, hd.globalize_row_ids(scan.to_host(), lineitem_base))
    if rank == 0:
        expected = orc.table_scan(global_lineitem, predicate, threads=8)
        assert row_ids_equal(np.concatenate(gathered), expected.row_ids), "distributed scan differs"

    # ---- join: radix all-to-all ----------------------------------------------------------------------------------------
    radix_bits = 4
    expected = orc.join_hash(global_orders, O_ORDERKEY, global_lineitem, L_ORDERKEY, capi.JOIN_INNER, radix_bits, threads=8) \
        if rank == 0 else None
    peers = hd.PeerExchange(device, torch_device, capacity=2 * shard.lineitem.row_count + 65_536)
    for exchange_kind, exchange in (("nccl all-to-all", None), ("fused split + NVLink P2P stores", peers)):
        pairs, offsets, build_rows, probe_rows, result = hd.device_distributed_join(
            device, orders, O_ORDERKEY, lineitem, L_ORDERKEY, radix_bits, orders_base, lineitem_base, torch_device, peers=exchange)
        if result is not None:
            got_build, got_probe = result.to_host()
            out_build = hd.unpack_row_ids(build_rows.cpu().numpy()[got_build["chunk_offset"]])
            out_probe = hd.unpack_row_ids(probe_rows.cpu().numpy()[got_probe["chunk_offset"]])
            result.free()
        else:
            out_build = out_probe = hd.unpack_row_ids(np.zeros(0, dtype=np.int64))
        gathered = [None] * world
        dist.all_gather_object(gathered, (out_build, out_probe, offsets))
        if rank == 0:
            build_parts, probe_parts = [], []
            for partition in range(1 << radix_bits):
                owner_build, owner_probe, owner_offsets = gathered[partition % world]
                build_parts.append(owner_build[int(owner_offsets[partition]):int(owner_offsets[partition + 1])])
                probe_parts.append(owner_probe[int(owner_offsets[partition]):int(owner_offsets[partition + 1])])
            assert row_ids_equal(np.concatenate(probe_parts), expected.probe), f"distributed join ({exchange_kind}): probe RowIDs differ"
            assert row_ids_equal(np.concatenate(build_parts), expected.build), f"distributed join ({exchange_kind}): build RowIDs differ"
    peers.close()

    # ---- aggregate: Q1 with one all-to-all of partial groups -----------------------------------------------------------
    q1_predicates = [Predicate(L_SHIPDATE, capi.PRED_LESS_THAN_EQUALS, "1998-09-02")]
    functions = [a.function for a in Q1_AGGREGATES]

    def local(decomposed):
        aggregates = [Aggregate(function, None if function == capi.AGG_COUNT_STAR else Q1_AGGREGATES[original].expression)
                      for function, original in decomposed]
        output = device.aggregate_hash(lineitem, [L_RETURNFLAG, L_LINESTATUS], aggregates, predicates=q1_predicates)
        keys = np.zeros((output.group_count, 2), dtype=np.int64)
        for g, row in enumerate(output.row_ids):
            keys[g, 0] = shard.lineitem.char_at(L_RETURNFLAG, int(row["chunk_id"]), int(row["chunk_offset"]))
            keys[g, 1] = shard.lineitem.char_at(L_LINESTATUS, int(row["chunk_id"]), int(row["chunk_offset"]))
        positions = (np.int64(rank) << 40) + output.row_ids["chunk_id"].astype(np.int64) * capi.DEFAULT_CHUNK_SIZE + \
            output.row_ids["chunk_offset"].astype(np.int64)
        values = [v.astype(np.float64) if v.dtype.kind == "f" else v.astype(np.int64) for v in output.values]
        counts = [v.astype(np.int64) if f in (capi.AGG_COUNT, capi.AGG_COUNT_STAR) else (~n).astype(np.int64)
                  for v, n, (f, _) in zip(output.values, output.nulls, decomposed)]
        return hd.PartialGroups(keys, np.zeros_like(keys, dtype=bool), positions, [f for f, _ in decomposed], values, counts)

    outcome = hd.distributed_aggregate(local, functions, torch_device)
    if rank == 0:
        merged, (values, nulls) = outcome
        expected = orc.aggregate_hash(global_lineitem, [L_RETURNFLAG, L_LINESTATUS], Q1_AGGREGATES, predicates=q1_predicates)
        assert len(merged.keys) == expected.group_count == 4
        # same group order: first appearance in global row order
        for index in range(len(functions)):
            if expected.values[index].dtype.kind == "f":
                assert np.allclose(values[index], expected.values[index], rtol=1e-6), index
            else:
                assert np.array_equal(values[index].astype(np.int64), expected.values[index].astype(np.int64)), index
        print(f"distributed OK on {world} GPUs: scan {sum(len(g) for g in gathered)} tuples gathered, join pairs "
              f"{sum(len(p) for p in probe_parts)}, Q1 groups {len(merged.keys)}")
    dist.barrier()
    device.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
