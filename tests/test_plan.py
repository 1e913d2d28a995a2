"""Host-side work decomposition of the attention kernel (fa_plan_segments, no device work): the stream-K cut must
cover every (item, key tile) exactly once, give every cut CTA an equal span, and describe the parts of a cut item
consistently for the writer and the merging reader (csrc/fa_fwd.cuh: FaSegment).

The reference's grid is one CTA per (query block, b*h) (FlashAttention/flashattn_forward_wmma/flashattn_forward_wmma_v4.cu:452).
Measured on B200 (scripts/tail_split_ab.py), cutting pays only for grids that fill less than half the SMs (small
batch x long sequence); multi-wave grids -- all BASELINE configs -- stay whole (csrc/api.cu: fa_plan)."""
import ctypes

import numpy as np
import pytest

import flashattention_kernel_project_b200 as fa


def segments(B, H, Nq, Nkv, sms=148, mode=1):
    """mode 1: the shipped policy; mode 2: the general stream-K plan (partial last wave of any grid cut along K/V);
    mode 3: only the row cut (two single-tile CTAs per item of a small last wave)."""
    lib = fa.load_library()
    before = lib.fa_set_tail_split(mode)
    try:
        n = lib.fa_plan_segments(B, H, Nq, Nkv, sms, None, 0)
        assert n > 0
        buf = (ctypes.c_int * (6 * n))()
        assert lib.fa_plan_segments(B, H, Nq, Nkv, sms, buf, n) == n
    finally:
        lib.fa_set_tail_split(before)
    return np.array(buf, dtype=np.int64).reshape(n, 6)


SHAPES = [(8, 32, 4096, 4096), (1, 32, 4096, 4096), (4, 16, 1024, 1024), (1, 1, 128, 128), (1, 1, 4096, 4096),
          (1, 3, 2000, 2000), (2, 5, 1500, 1500), (1, 2, 777, 777), (3, 5, 512, 512), (1, 2, 300, 300), (1, 7, 1000, 5000),
          (2, 3, 257, 129), (1, 1, 65536, 65536), (1, 5, 8192, 9999), (1, 1, 300, 70000)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("sms", [148, 132, 7])
@pytest.mark.parametrize("mode", [1, 2, 3])
def test_every_key_tile_of_every_item_is_covered_exactly_once(shape, sms, mode):
    B, H, Nq, Nkv = shape
    seg = segments(B, H, Nq, Nkv, sms, mode)
    q_blocks = (Nq + 255) // 256
    items, tiles = B * H * q_blocks, (Nkv + 127) // 128
    cover = np.zeros((items, 2, tiles), dtype=np.int64)  # [item, 128-row tile, key tile]
    for cta, item, k0, k1, slot, tail_item in seg:
        assert 0 <= k0 < k1 <= tiles
        if slot <= -2:  # row cut: one query tile against every key tile
            assert (k0, k1) == (0, tiles)
            cover[item, -2 - slot] += 1
        else:
            cover[item, :, k0:k1] += 1
            assert (slot == -1) == (k0 == 0 and k1 == tiles)  # whole items take the ordinary epilogue
    for item in range(items):  # a ragged last query block has no second tile: nobody may be sent to compute it
        two = (item % q_blocks) * 256 + 128 < Nq
        assert (cover[item, 0] == 1).all() and (cover[item, 1] == (1 if two else cover[item, 1, 0])).all()
        halves = seg[(seg[:, 1] == item) & (seg[:, 4] <= -2)]
        assert len(halves) in (0, 2 if two else 1)
    rows = seg[seg[:, 4] <= -2]
    if len(rows):
        # the row cut applies to a last wave that fills at most half the SMs, after whole-item full waves
        first = rows[:, 0].min()
        assert first % sms == 0 and len(np.unique(rows[:, 1])) * 2 <= sms and (seg[seg[:, 0] < first][:, 4] == -1).all()
        assert not (seg[:, 4] >= 0).any()  # never both cuts in one launch
    # CTAs are numbered densely; whole-item CTAs come first, each with exactly one segment
    ctas = seg[:, 0]
    idle = set(range(ctas.max() + 1)) - set(ctas)
    assert ctas.min() == 0
    if idle:  # only the would-be second tile of a ragged last query block (row cut) leaves a CTA without work
        assert Nq % 256 != 0 and Nq % 256 <= 128 and len(idle) <= B * H
        assert all((c - 1) in set(seg[seg[:, 4] == -2][:, 0]) for c in idle)
    tail = seg[seg[:, 4] >= 0]
    if len(tail) and mode == 1:
        # a cut grid is cut as a whole: equal spans (they differ by at most one key tile), at least 4 key tiles each,
        # no more CTAs than SMs, at most 8 parts per item
        assert items * 2 <= sms
        work = np.bincount(seg[:, 0], weights=seg[:, 3] - seg[:, 2])
        assert work.max() - work.min() <= 1 and work.min() >= 4 and len(work) <= sms
        assert np.bincount(tail[:, 1]).max() <= 9
    if len(tail) and mode == 2:
        # whole-item CTAs first (one segment each, full waves), then the cut CTAs with equal spans of >= 2 key tiles
        first_tail_cta = tail[:, 0].min()
        full = seg[seg[:, 0] < first_tail_cta]
        assert (np.bincount(full[:, 0]) == 1).all() and (full[:, 4] == -1).all() and first_tail_cta % sms == 0
        sel = seg[:, 0] >= first_tail_cta
        work = np.bincount(seg[sel, 0] - first_tail_cta, weights=(seg[:, 3] - seg[:, 2])[sel])
        assert work.max() - work.min() <= 1 and work.min() >= 2 and len(work) <= sms


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("mode", [1, 2])
def test_parts_of_a_cut_item_use_distinct_slots_in_cta_order(shape, mode):
    """The merging CTA reads part t of an item from slot 1 of tail CTA t only if that CTA's span began in an earlier
    item (then it is the item's first part), else from slot 0: no two parts may share a (cta, slot) pair, and the
    parts of an item must be consecutive tail CTAs in key order."""
    seg = segments(*shape, mode=mode)
    tail = seg[seg[:, 4] >= 0]
    assert len({(c, s) for c, _, _, _, s, _ in tail}) == len(tail)
    for item in np.unique(tail[:, 1]):
        parts = tail[tail[:, 1] == item]
        parts = parts[np.argsort(parts[:, 2])]
        assert (np.diff(parts[:, 0]) == 1).all()            # consecutive CTAs ...
        assert (parts[1:, 2] == parts[:-1, 3]).all()        # ... over consecutive key-tile ranges
        assert (parts[1:, 4] == 0).all()                    # later parts start their span inside the item
        assert parts[0, 2] == 0 and parts[-1, 3] == (shape[3] + 127) // 128
        assert len(np.unique(parts[:, 5])) == 1


def test_policy_follows_the_measurements():
    """Multi-wave grids are never cut along K/V (BASELINE configs 2, 3 and 4); a last wave of at most half the SMs is
    cut along query rows (config 3 as the 32-(b,h) shard of 8 GPUs: 512 items = 3 waves + 68 items -> 136 single-tile
    CTAs); one (b,h) of N = 4096 (16 items x 32 key tiles) is cut along K/V into 4 parts per item
    (sqrt(32 / 2.2) ~ 3.8); short sequences on few SMs take the row cut."""
    for shape in [(4, 16, 1024, 1024), (8, 32, 4096, 4096), (1, 32, 8192, 8192), (1, 4, 16384, 16384)]:
        seg = segments(*shape)  # last waves of 108, 100, 136 and 108 items: more than half the SMs
        assert (seg[:, 4] == -1).all() and len(seg) == shape[0] * shape[1] * ((shape[2] + 255) // 256)
    seg = segments(1, 32, 4096, 4096)
    assert len(seg) == 444 + 136 and (seg[:444, 4] == -1).all() and (seg[444:, 4] <= -2).all()
    assert (seg[444:, 1] == 444 + np.arange(136) // 2).all() and (seg[444:, 4] == -2 - np.arange(136) % 2).all()
    seg = segments(1, 32, 65536, 65536)  # config 4 on one GPU: 8192 items = 55 waves + 52
    assert len(seg) == 8192 + 52 and (seg[:8140, 4] == -1).all() and (seg[8140:, 4] <= -2).all()
    seg = segments(1, 1, 4096, 4096)
    assert len(seg) == 64 and (seg[:, 3] - seg[:, 2] == 8).all() and (seg[:, 4] >= 0).all()
    seg = segments(1, 8, 1024, 1024)  # 32 items x 8 key tiles: K/V spans would drop below 4 tiles
    assert len(seg) == 64 and (seg[:, 4] <= -2).all()
    seg = segments(1, 1, 16384, 16384)  # 64 items x 128 tiles, 148 // 64 = 2 parts each
    assert len(seg) == 128 and (seg[:, 3] - seg[:, 2] == 64).all()
    seg = segments(1, 1, 128, 128)  # BASELINE config 1: one query tile, nothing to cut
    assert len(seg) == 1 and seg[0, 4] == -1


def test_bad_arguments():
    lib = fa.load_library()
    assert lib.fa_plan_segments(0, 1, 128, 128, 148, None, 0) == -1
    assert lib.fa_plan_segments(1, 1, 128, 128, 0, None, 0) == -1
