"""CPU: the zero-scratch arena behind `ops.new_stats` — slices are zero, disjoint, 16-byte aligned, never handed out twice,
and a chunk stays alive exactly as long as one of its slices is referenced."""
import gc
import weakref

import torch

from b200seg import ops


def test_slices_are_zero_disjoint_and_aligned():
    arena = ops._ZeroArena()
    a = arena.take(5, torch.float64, "cpu")
    b = arena.take(7, torch.float64, "cpu")
    c = arena.take(3, torch.float32, "cpu")
    assert a.numel() == 5 and b.numel() == 7 and c.numel() == 3
    assert float(a.abs().sum()) == 0 and float(b.abs().sum()) == 0 and float(c.abs().sum()) == 0
    a.fill_(1.0)
    assert float(b.abs().sum()) == 0                      # writing one slice never touches another
    assert (b.data_ptr() - a.data_ptr()) >= 5 * 8 and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0 and c.data_ptr() % 16 == 0
    assert a.untyped_storage().data_ptr() == b.untyped_storage().data_ptr()        # same chunk
    assert a.untyped_storage().data_ptr() != c.untyped_storage().data_ptr()        # one chunk per dtype


def test_chunk_rollover_and_lifetime():
    arena = ops._ZeroArena()
    n = arena.CHUNK_BYTES // 8
    first = arena.take(n - 2, torch.float64, "cpu")
    second = arena.take(16, torch.float64, "cpu")          # does not fit: a fresh, zeroed chunk
    assert first.untyped_storage().data_ptr() != second.untyped_storage().data_ptr()
    assert float(second.abs().sum()) == 0
    big = arena.take(3 * n, torch.float64, "cpu")          # larger than a chunk: its own allocation
    assert big.numel() == 3 * n and float(big.abs().sum()) == 0
    ref = weakref.ref(first.untyped_storage())
    del first
    gc.collect()
    assert ref() is None                                    # the exhausted chunk is freed with its last slice


def test_new_stats_shape_and_dtype():
    st = ops.new_stats(2, 5, "cpu")
    assert st.shape == (2, 5, 2) and st.dtype == torch.float64 and st.is_contiguous() and float(st.abs().sum()) == 0
    z = ops.zeros_scratch((3, 4), torch.float32, "cpu")
    assert z.shape == (3, 4) and z.dtype == torch.float32 and float(z.abs().sum()) == 0
