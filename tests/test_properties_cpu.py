"""Property tests (hypothesis) of pure host-side index logic: the sequence-parallel token layout, the optimizer's multi-tensor
chunk map, padding-mask -> kv range conversion, packed-weight detection."""
import types

import torch
from hypothesis import given, settings
from hypothesis import strategies as st


def _state(world, chunks, rank):
    from transformers_b200.parallel import SequenceParallelState

    s = types.SimpleNamespace(world=world, chunks=chunks, rank=rank)
    s.chunk_rows = types.MethodType(SequenceParallelState.chunk_rows, s)
    return s


@settings(max_examples=60, deadline=None)
@given(world=st.sampled_from([1, 2, 4, 8]), chunks=st.integers(1, 4), piece=st.integers(1, 5), H=st.integers(1, 3))
def test_sequence_parallel_layout_round_trips(world, chunks, piece, H):
    """shard (sp_take_local on every rank) followed by the chunk-wise all-gather (rank-order concatenation inside every row
    block, which is what all_gather_into_tensor does) reproduces the tokens in their original order."""
    from transformers_b200.parallel import sp_take_local

    T = world * chunks * piece
    full = torch.arange(T * H, dtype=torch.float32).view(T, H)
    shards = [sp_take_local(full, _state(world, chunks, r)) for r in range(world)]
    assert all(s.shape == (T // world, H) for s in shards)
    rebuilt = torch.empty_like(full)
    s0 = _state(world, chunks, 0)
    for fr, lr in zip(s0.chunk_rows(T), s0.chunk_rows(T // world)):
        rebuilt[fr] = torch.cat([s[lr] for s in shards], dim=0)
    assert torch.equal(rebuilt, full)
    # every token is owned by exactly one rank
    owned = torch.cat(shards)[:, 0]
    assert sorted(owned.tolist()) == full[:, 0].tolist()


@settings(max_examples=40, deadline=None)
@given(numels=st.lists(st.integers(1, 200000), min_size=1, max_size=6))
def test_optimizer_chunk_map_covers_every_element_once(numels):
    import transformers_b200.optim as optim
    from transformers_b200 import ops

    chunk = 32768
    orig = ops.optim_chunk_elems
    ops.optim_chunk_elems = lambda: chunk
    try:
        table, cmap = optim._tables([(1000 + i, 2000 + i, 3000 + i, 4000 + i, n) for i, n in enumerate(numels)], torch.device("cpu"))
    finally:
        ops.optim_chunk_elems = orig
    assert table.shape == (len(numels), 6) and table[:, 4].tolist() == numels
    covered = [0] * len(numels)
    seen = set()
    for ti, ci in cmap.tolist():
        assert (ti, ci) not in seen
        seen.add((ti, ci))
        covered[ti] += min(chunk, numels[ti] - ci * chunk)
        assert ci * chunk < numels[ti]
    assert covered == numels


@settings(max_examples=60, deadline=None)
@given(L=st.integers(1, 40), data=st.data())
def test_mask_to_kv_ranges_matches_definition(L, data):
    from transformers_b200.modules import mask_to_kv_ranges

    rows = []
    for _ in range(3):
        a = data.draw(st.integers(0, L - 1))
        b = data.draw(st.integers(a + 1, L))
        rows.append([1 if a <= i < b else 0 for i in range(L)])
    m = torch.tensor(rows)
    s, e = mask_to_kv_ranges(m)
    for r, row in enumerate(rows):
        ones = [i for i, x in enumerate(row) if x]
        assert s[r].item() == ones[0] and e[r].item() == ones[-1] + 1


@settings(max_examples=30, deadline=None)
@given(sizes=st.lists(st.integers(1, 5), min_size=2, max_size=4), K=st.integers(1, 6), shift=st.integers(0, 1))
def test_packed_detection_only_accepts_adjacent_row_views(sizes, K, shift):
    from transformers_b200.modules import _is_packed

    buf = torch.arange(sum(sizes) * K, dtype=torch.float32).view(sum(sizes), K)
    views, off = [], 0
    for n in sizes:
        views.append(buf[off:off + n])
        off += n
    assert _is_packed(buf, views)
    assert not _is_packed(buf, [v.clone() for v in views])
    if shift and len(views) > 1 and sizes[0] != sizes[1]:
        assert not _is_packed(buf, [views[1], views[0]] + views[2:])  # same rows, wrong order


@settings(max_examples=60, deadline=None)
@given(world=st.sampled_from([1, 2, 4, 8]), blocks=st.integers(1, 3), num_n=st.integers(1, 9), group_m=st.sampled_from([1, 4, 8, 16]),
       rank=st.integers(0, 7))
def test_scatter_tile_order_visits_every_tile_once_and_own_block_last(world, blocks, num_n, group_m, rank):
    """Restatement of the tile walk of gemm2.cu's SCATTER mode (tile_coords + m_rot rotation, owner = row0 / rows_per_owner):
    every (m-tile, n-tile) is produced exactly once, the destination strip index stays inside the owner's buffer, and the
    rotation makes the own row block the LAST one in tile order (so remote blocks travel while later tiles compute)."""
    rank %= world
    tiles_per_owner = blocks                      # rows_per_owner = blocks * 256
    num_m = world * tiles_per_owner
    rows_per_owner = tiles_per_owner * 256
    m_rot = ((rank + 1) % world) * tiles_per_owner
    seen, owners_in_order = set(), []
    for tile in range(num_m * num_n):
        group_size = group_m * num_n
        group = tile // group_size
        first_m = group * group_m
        gsz = min(group_m, num_m - first_m)
        in_group = tile - group * group_size
        tm = first_m + in_group % gsz
        tn = in_group // gsz
        tm += m_rot
        if tm >= num_m:
            tm -= num_m
        assert 0 <= tm < num_m and 0 <= tn < num_n and (tm, tn) not in seen
        seen.add((tm, tn))
        for cta_rank in range(2):
            for q in range(4):
                row0 = tm * 256 + cta_rank * 128 + q * 32
                owner = row0 // rows_per_owner
                local = row0 - owner * rows_per_owner
                assert 0 <= owner < world and 0 <= local and local + 32 <= rows_per_owner
        owners_in_order.append(tm * 256 // rows_per_owner)
    assert len(seen) == num_m * num_n
    if world > 1 and group_m <= tiles_per_owner * world:
        first_own = owners_in_order.index(rank)
        assert all(o == rank for o in owners_in_order[first_own:]) or group_m > 1  # strict "own block last" holds for group_m == 1
        assert owners_in_order[-1] == rank or group_m > 1


@settings(max_examples=60, deadline=None)
@given(rows=st.lists(st.lists(st.integers(1, 9), min_size=1, max_size=6), min_size=1, max_size=3))
def test_packed_sequence_ranges_partition_every_row(rows):
    """modules.SegmentIds.ranges(): from the sequence indices the reference derives for a packed batch
    (find_packed_sequence_indices, masking_utils.py:728-757) to the (start, end) token ranges the attention Functions launch
    on: every row is partitioned exactly at the places where the index changes."""
    from transformers_b200.modules import SegmentIds

    S = max(sum(r) for r in rows)
    lens = [r + ([S - sum(r)] if sum(r) < S else []) for r in rows]
    pos = torch.stack([torch.cat([torch.arange(n) for n in r]) for r in lens])
    first = pos[:, :1] - 1
    ids = (torch.diff(pos, prepend=first, dim=-1) != 1).cumsum(-1)  # the reference's formula
    got = SegmentIds(ids).ranges()
    assert len(got) == len(lens)
    for r, segs in zip(lens, got):
        ends = torch.tensor(r).cumsum(0).tolist()
        assert segs == list(zip([0] + ends[:-1], ends))


@settings(max_examples=40, deadline=None)
@given(blocks=st.integers(1, 6), K=st.integers(1, 5))
def test_gate_up_interleave_round_trips(blocks, K):
    """ops.interleave_gate_up / deinterleave_gate_up: 128-row blocks alternate gate / up; the inverse returns the inputs; row j
    of gate sits at (j // 128) * 256 + j % 128, the matching up row 128 further."""
    from transformers_b200 import ops

    I = 128 * blocks
    wg = torch.arange(I * K, dtype=torch.float32).view(I, K)
    wu = -wg - 1
    ilv = ops.interleave_gate_up(wg, wu)
    assert ilv.shape == (2 * I, K)
    j = torch.arange(I)
    at = (j // 128) * 256 + j % 128
    assert torch.equal(ilv[at], wg) and torch.equal(ilv[at + 128], wu)
    g, u = ops.deinterleave_gate_up(ilv)
    assert torch.equal(g, wg) and torch.equal(u, wu)


@settings(max_examples=40, deadline=None)
@given(cu=st.lists(st.integers(1, 50), min_size=1, max_size=8))
def test_cu_seq_lens_become_ranges_or_are_refused(cu):
    """SegmentIds.from_cu_seqlens: cumulative lengths of a flattened batch (modeling_flash_attention_utils.py:570-590) -> one row
    of ranges; lists that do not partition the batch raise."""
    import pytest

    from transformers_b200 import B200Error
    from transformers_b200.modules import SegmentIds

    ends = torch.tensor(cu).cumsum(0).tolist()
    total = ends[-1]
    seg = SegmentIds.from_cu_seqlens(torch.tensor([0] + ends, dtype=torch.int32), total)
    assert seg.ranges() == [list(zip([0] + ends[:-1], ends))]
    with pytest.raises(B200Error):
        SegmentIds.from_cu_seqlens([0] + ends, total + 1)
    with pytest.raises(B200Error):
        SegmentIds.from_cu_seqlens([1] + ends, total)
