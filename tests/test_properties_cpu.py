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
