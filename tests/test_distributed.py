"""World-size-2 gloo tests of the multi-GPU host logic (hyrise_b200/distributed.py). The per-rank compute is the CPU
oracle here (on the GPU box it is the C-ABI kernels); what is under test is the sharding, the single all-to-all exchange
per operator and the merge, checked against the oracle run on the whole table in one process."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as orc
from helpers import column_values_at, random_table, row_ids_equal
from hyrise_b200 import capi
from hyrise_b200 import distributed as hd
from hyrise_b200.device import ROW_ID_DTYPE, Aggregate, Expression, Predicate
from hyrise_b200.storage import ColumnDefinition, Table

WORLD = 2


def free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def full_table():
    rng = np.random.default_rng(4242)
    return random_table(rng, 20_000, 1_000).encode("Dictionary")


def shard_of(table: Table, rank: int, world: int):
    """Consecutive chunk range of rank `rank` + its first global chunk id and first global row."""
    per_rank = (table.chunk_count + world - 1) // world
    begin, end = rank * per_rank, min(table.chunk_count, (rank + 1) * per_rank)
    shard = Table(table.column_definitions, table.chunks[begin:end], table.target_chunk_size)
    row_base = sum(chunk.size for chunk in table.chunks[:begin])
    return shard, begin, row_base


def encode_keys(table: Table, groupby, output):
    keys = np.zeros((output.group_count, len(groupby)), dtype=np.int64)
    nulls = np.zeros((output.group_count, len(groupby)), dtype=bool)
    for column_index, column_id in enumerate(groupby):
        values = column_values_at(table, column_id, output.row_ids)
        for g, value in enumerate(values):
            if value is None:
                nulls[g, column_index] = True
            elif isinstance(value, float):
                keys[g, column_index] = np.array([value], dtype=np.float64).view(np.int64)[0]
            else:
                keys[g, column_index] = int(value)
    return keys, nulls


def local_partial(shard: Table, row_base: int, groupby, expressions, decomposed):
    aggregates = [Aggregate(function, None if function == capi.AGG_COUNT_STAR else expressions[original])
                  for function, original in decomposed]
    output = orc.aggregate_hash(shard, groupby, aggregates)
    keys, nulls = encode_keys(shard, groupby, output)
    chunk_starts = np.concatenate([[0], np.cumsum([chunk.size for chunk in shard.chunks])])
    positions = np.array([row_base + chunk_starts[int(r["chunk_id"])] + int(r["chunk_offset"]) for r in output.row_ids],
                         dtype=np.int64)
    values, counts = [], []
    for index, (function, _) in enumerate(decomposed):
        raw = output.values[index]
        values.append(raw.astype(np.float64) if raw.dtype.kind == "f" else raw.astype(np.int64))
        # non-NULL input count: COUNT gives it directly; for SUM/MIN/MAX "is NULL" is all the merge needs
        counts.append(raw.astype(np.int64) if function in (capi.AGG_COUNT, capi.AGG_COUNT_STAR)
                      else (~output.nulls[index]).astype(np.int64))
    return hd.PartialGroups(keys, nulls, positions, [f for f, _ in decomposed], values, counts)


def worker(rank: int, port: int, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    cpu = torch.device("cpu")
    try:
        table = full_table()
        shard, chunk_base, row_base = shard_of(table, rank, WORLD)

        # ---- scan: chunk-partitioned, no collective; only the test gathers the per-rank PosLists ------------------
        predicate = Predicate(4, capi.PRED_LESS_THAN, 100)
        local = orc.table_scan(shard, predicate)
        assert hd.chunk_bases(shard.chunk_count, cpu)[rank] == chunk_base
        gathered = [None] * WORLD
        dist.all_gather_object(gathered, hd.globalize_row_ids(local.row_ids, chunk_base))
        if rank == 0:
            expected = orc.table_scan(table, predicate)
            assert row_ids_equal(np.concatenate(gathered), expected.row_ids)

        # ---- aggregate: pre-aggregate, exchange by key hash, merge ------------------------------------------------
        functions = [capi.AGG_SUM, capi.AGG_AVG, capi.AGG_MIN, capi.AGG_MAX, capi.AGG_COUNT, capi.AGG_COUNT_STAR, capi.AGG_AVG]
        expressions = [Expression.column(0), Expression.column(2), Expression.column(1), Expression.column(3),
                       Expression.column(0), None, Expression.column(1)]
        for groupby in ([4], [], [4, 0], [2]):
            outcome = hd.distributed_aggregate(
                lambda decomposed: local_partial(shard, row_base, groupby, expressions, decomposed), functions, cpu)
            if rank == 0:
                merged, (values, nulls) = outcome
                aggregates = [Aggregate(f, e) for f, e in zip(functions, expressions)]
                expected = orc.aggregate_hash(table, groupby, aggregates)
                # distributed order = first appearance in global row order = the reference's order (non-immediate keys)
                expected_keys, expected_nulls = encode_keys(table, groupby, expected)
                order = np.arange(expected.group_count)
                if expected.used_immediate_keys:  # ascending key order in the reference: compare as sets of groups
                    chunk_starts = np.concatenate([[0], np.cumsum([chunk.size for chunk in table.chunks])])
                    order = np.lexsort(np.concatenate([np.where(expected_nulls, 0, expected_keys), expected_nulls], axis=1).T[::-1])
                    mine = np.lexsort(np.concatenate([np.where(merged.key_nulls, 0, merged.keys), merged.key_nulls], axis=1).T[::-1])
                else:
                    mine = np.arange(len(merged.keys))
                assert len(merged.keys) == expected.group_count, groupby
                assert np.array_equal(np.where(merged.key_nulls, 0, merged.keys)[mine],
                                      np.where(expected_nulls, 0, expected_keys)[order]), groupby
                for index in range(len(functions)):
                    want, want_null = expected.values[index][order], expected.nulls[index][order]
                    got, got_null = values[index][mine], nulls[index][mine]
                    assert np.array_equal(got_null, want_null), (groupby, index)
                    if want.dtype.kind == "f":
                        assert np.allclose(got[~want_null], want[~want_null], rtol=1e-9), (groupby, index)
                    else:
                        assert np.array_equal(got[~want_null].astype(np.int64), want[~want_null].astype(np.int64)), (groupby, index)

        # ---- join: radix exchange of {key, RowID} tuples, local join, partition-major reassembly -----------------------
        rng = np.random.default_rng(77)
        build_full = Table.from_columns([ColumnDefinition("k", capi.TYPE_INT32, True)],
                                        [rng.integers(0, 3000, 6_000, dtype=np.int32)], [rng.random(6_000) < 0.05], 500)
        probe_full = Table.from_columns([ColumnDefinition("k", capi.TYPE_INT32, True)],
                                        [rng.integers(0, 4000, 16_000, dtype=np.int32)], [rng.random(16_000) < 0.05], 1_000)
        radix_bits = 3
        sides = []
        for side_table in (build_full, probe_full):
            side_shard, side_chunk_base, _ = shard_of(side_table, rank, WORLD)
            keys, nulls = side_shard.column_values(0)
            row_ids = np.zeros(side_shard.row_count, dtype=ROW_ID_DTYPE)
            cursor = 0
            for chunk_id, chunk in enumerate(side_shard.chunks):
                row_ids["chunk_id"][cursor:cursor + chunk.size] = side_chunk_base + chunk_id
                row_ids["chunk_offset"][cursor:cursor + chunk.size] = np.arange(chunk.size)
                cursor += chunk.size
            keep = ~nulls                                                # Inner join: NULL keys are discarded on both sides
            received_keys, received_rows = hd.exchange_tuples(torch.from_numpy(keys[keep].astype(np.int64)),
                                                              torch.from_numpy(hd.pack_row_ids(row_ids[keep])))
            # the GPU path's exchange (hyb_join_partition's stable split, then exchange_partitioned) must deliver the same
            # tuples in the same order; the split is emulated with numpy here
            kept_keys, kept_rows = keys[keep].astype(np.int64), hd.pack_row_ids(row_ids[keep])
            owner = kept_keys & (WORLD - 1)
            order = np.argsort(owner, kind="stable")
            offsets = [0] + np.cumsum(np.bincount(owner, minlength=WORLD)).tolist()
            padding = np.zeros(8, dtype=np.int64)
            (again_keys, again_rows, again_count), = hd.exchange_partitioned(
                [(torch.from_numpy(np.concatenate([kept_keys[order], padding])),
                  torch.from_numpy(np.concatenate([kept_rows[order], padding])), offsets)])
            assert again_count == len(received_keys) and len(again_keys) == again_count + 8
            assert np.array_equal(again_keys[:again_count].numpy(), received_keys.numpy())
            assert np.array_equal(again_rows[:again_count].numpy(), received_rows.numpy())
            sides.append((received_keys.numpy(), hd.unpack_row_ids(received_rows.numpy())))
        (build_keys, build_rows), (probe_keys, probe_rows) = sides
        assert ((build_keys & (WORLD - 1)) == rank).all() and ((probe_keys & (WORLD - 1)) == rank).all()
        local_build = Table.from_columns([ColumnDefinition("k", capi.TYPE_INT64)], [build_keys], chunk_size=max(len(build_keys), 1))
        local_probe = Table.from_columns([ColumnDefinition("k", capi.TYPE_INT64)], [probe_keys], chunk_size=max(len(probe_keys), 1))
        joined = orc.join_hash(local_build, 0, local_probe, 0, capi.JOIN_INNER, radix_bits)
        out_build = build_rows[joined.build["chunk_offset"]] if joined.pair_count else np.zeros(0, dtype=ROW_ID_DTYPE)
        out_probe = probe_rows[joined.probe["chunk_offset"]] if joined.pair_count else np.zeros(0, dtype=ROW_ID_DTYPE)
        gathered = [None] * WORLD
        dist.all_gather_object(gathered, (out_build, out_probe, joined.partition_offsets))
        if rank == 0:
            expected = orc.join_hash(build_full, 0, probe_full, 0, capi.JOIN_INNER, radix_bits)
            build_parts, probe_parts = [], []
            for partition in range(1 << radix_bits):
                owner_build, owner_probe, offsets = gathered[partition % WORLD]
                build_parts.append(owner_build[int(offsets[partition]):int(offsets[partition + 1])])
                probe_parts.append(owner_probe[int(offsets[partition]):int(offsets[partition + 1])])
            assert row_ids_equal(np.concatenate(probe_parts), expected.probe)
            assert row_ids_equal(np.concatenate(build_parts), expected.build)
        results.put((rank, "ok"))
    except Exception as error:  # noqa: BLE001
        import traceback
        results.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_world_size_two_gloo():
    context = mp.get_context("spawn")
    results = context.Queue()
    port = free_port()
    processes = [context.Process(target=worker, args=(rank, port, results)) for rank in range(WORLD)]
    for process in processes:
        process.start()
    outcomes = [results.get(timeout=300) for _ in range(WORLD)]
    for process in processes:
        process.join(timeout=60)
    for rank, outcome in outcomes:
        assert outcome == "ok", f"rank {rank}:\n{outcome}"


def test_merge_partial_groups_single_process():
    keys = np.array([[1], [2], [1], [3]], dtype=np.int64)
    nulls = np.zeros((4, 1), dtype=bool)
    partial = hd.PartialGroups(keys, nulls, np.array([5, 1, 9, 7]), [capi.AGG_SUM, capi.AGG_MIN, capi.AGG_COUNT_STAR],
                               [np.array([1.5, 2.0, 2.5, 4.0]), np.array([7, 8, 3, 9]), np.array([2, 1, 1, 1])],
                               [np.array([2, 1, 1, 0]), np.array([2, 1, 1, 1]), np.array([2, 1, 1, 1])])
    merged = hd.merge_partial_groups([partial])
    assert merged.keys[:, 0].tolist() == [1, 2, 3]
    assert merged.values[0].tolist() == [4.0, 2.0, 0.0] and merged.counts[0].tolist() == [3, 1, 0]
    assert merged.values[1].tolist() == [3, 8, 9] and merged.values[2].tolist() == [3, 1, 1]
    assert merged.first_position.tolist() == [5, 1, 7]
    local, recipe = hd.decompose_aggregates([capi.AGG_AVG, capi.AGG_MAX])
    assert local == [(capi.AGG_SUM, 0), (capi.AGG_COUNT, 0), (capi.AGG_MAX, 1)] and recipe[0] == ("avg", 0, 1)
