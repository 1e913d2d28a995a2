"""Host-side work decomposition of the attention kernel (fa_plan_segments, no device work): the stream-K cut of the
last wave must cover every (item, key tile) exactly once, give every tail CTA an equal span, and describe the
parts of a cut item consistently for the writer and the merging reader (csrc/fa_fwd.cuh: FaSegment).

The reference's grid is one CTA per (query block, b*h) (FlashAttention/flashattn_forward_wmma/flashattn_forward_wmma_v4.cu:452);
BASELINE config 3 head-sharded over 8 GPUs leaves 512 such items for 148 SMs, i.e. 3.46 waves."""
import ctypes

import numpy as np
import pytest

import flashattention_kernel_project_b200 as fa


def segments(B, H, Nq, Nkv, sms=148):
    lib = fa.load_library()
    n = lib.fa_plan_segments(B, H, Nq, Nkv, sms, None, 0)
    assert n > 0
    buf = (ctypes.c_int * (6 * n))()
    assert lib.fa_plan_segments(B, H, Nq, Nkv, sms, buf, n) == n
    return np.array(buf, dtype=np.int64).reshape(n, 6)


SHAPES = [(8, 32, 4096, 4096), (1, 32, 4096, 4096), (2, 32, 4096, 4096), (4, 16, 1024, 1024), (1, 1, 128, 128),
          (1, 1, 4096, 4096), (1, 32, 8192, 8192), (3, 5, 512, 512), (1, 2, 300, 300), (1, 7, 1000, 5000),
          (2, 3, 257, 129), (1, 1, 65536, 65536), (5, 31, 2048, 640)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("sms", [148, 132, 7])
def test_every_key_tile_of_every_item_is_covered_exactly_once(shape, sms):
    B, H, Nq, Nkv = shape
    seg = segments(B, H, Nq, Nkv, sms)
    items, tiles = B * H * ((Nq + 255) // 256), (Nkv + 127) // 128
    cover = np.zeros((items, tiles), dtype=np.int64)
    for cta, item, k0, k1, slot, tail_item in seg:
        assert 0 <= k0 < k1 <= tiles
        cover[item, k0:k1] += 1
        assert (slot == -1) == (k0 == 0 and k1 == tiles)  # whole items take the ordinary epilogue
    assert (cover == 1).all()
    # CTAs are numbered densely; whole-item CTAs come first, each with exactly one segment
    ctas = seg[:, 0]
    assert ctas.min() == 0 and set(ctas) == set(range(ctas.max() + 1))
    tail = seg[seg[:, 4] >= 0]
    if len(tail):
        first_tail_cta = tail[:, 0].min()
        full = seg[seg[:, 0] < first_tail_cta]
        assert (np.bincount(full[:, 0]) == 1).all() and (full[:, 4] == -1).all()
        assert first_tail_cta % sms == 0  # the cut applies to the partial last wave only
        # equal spans: the tail CTAs' work differs by at most one key tile
        work = np.bincount(seg[seg[:, 0] >= first_tail_cta, 0] - first_tail_cta,
                           weights=(seg[:, 3] - seg[:, 2])[seg[:, 0] >= first_tail_cta])
        assert work.max() - work.min() <= 1 and work.min() >= 2
        assert len(work) <= sms


@pytest.mark.parametrize("shape", SHAPES)
def test_parts_of_a_cut_item_use_distinct_slots_in_cta_order(shape):
    """The merging CTA reads part t of an item from slot 1 of tail CTA t only if that CTA's span began in an earlier
    item (then it is the item's first part), else from slot 0: no two parts may share a (cta, slot) pair, and the
    parts of an item must be consecutive tail CTAs in key order."""
    seg = segments(*shape)
    tail = seg[seg[:, 4] >= 0]
    assert len({(c, s) for c, _, _, _, s, _ in tail}) == len(tail)
    for item in np.unique(tail[:, 1]):
        parts = tail[tail[:, 1] == item]
        parts = parts[np.argsort(parts[:, 2])]
        assert (np.diff(parts[:, 0]) == 1).all()            # consecutive CTAs ...
        assert (parts[1:, 2] == parts[:-1, 3]).all()        # ... over consecutive key-tile ranges
        assert (parts[1:, 4] == 0).all()                    # later parts start their span inside the item
        assert len(np.unique(parts[:, 5])) == 1


def test_config3_shard_of_eight_gpus_needs_fewer_than_four_waves():
    """32 (b,h) of config 3 = 512 items on 148 SMs: 3 full waves + a cut tail of ceil(68 * 32 / 148) = 15 tiles."""
    seg = segments(1, 32, 4096, 4096, 148)
    work = np.bincount(seg[:, 0], weights=seg[:, 3] - seg[:, 2])
    assert (work[:444] == 32).all() and len(work) == 444 + 148 and work[444:].max() == 15


def test_bad_arguments():
    lib = fa.load_library()
    assert lib.fa_plan_segments(0, 1, 128, 128, 148, None, 0) == -1
    assert lib.fa_plan_segments(1, 1, 128, 128, 0, None, 0) == -1
