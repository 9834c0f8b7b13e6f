"""Tile enumeration of the a-trous level launches, checked on the CPU (no GPU, no device work).

`rmd_debug_level_cover` walks a launch's grid on the host with the SAME inline function the kernel uses to map
blockIdx to (column block, row phase, lattice tile) — csrc/svgf_atrous_tile.cu `tile_pos` / `level_grid` — and counts
how often every (row, column block) would be stored.  Properties (DESIGN.md "Band mode", reference border rule
src/filter.cu:38-39 is untouched by this: it is about which tiles run, not what they compute):
  * a whole-plane launch stores every row exactly once per column block, at every step and ragged size;
  * a band launch stores exactly the rows [row0, row0 + rows);
  * a band's boundary launch (split = 1) and interior launch (split = 2) are disjoint and together equal split = 0;
    the boundary launch enumerates only a few tile rows (that is its point);
  * every compiled variant (phase-major, lattice-row-major, serpentine) covers the same set."""
import ctypes

import numpy as np
import pytest

from raymarchdenoisercuda_b200 import _lib, build

WT = 128  # output columns per CTA


def cover(W, H, level, row0=0, rows=None, split=0, edges=((0, 0), (0, 0)), reverse=0, variant=-1):
    lib = _lib.load()
    rows = H - row0 if rows is None else rows
    nbx_max = (W + WT - 1) // WT
    buf = np.zeros(H * nbx_max, dtype=np.int32)
    nbx, tiles = ctypes.c_int(0), ctypes.c_int(0)
    grid = lib.rmd_debug_level_cover(W, H, level, row0, rows, split, edges[0][0], edges[0][1], edges[1][0], edges[1][1],
                                     reverse, variant, buf.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nbx), ctypes.byref(tiles))
    assert grid >= 0, grid
    assert nbx.value == nbx_max
    return buf.reshape(H, nbx_max), grid, tiles.value


@pytest.mark.parametrize("shape", [(1920, 1080), (333, 190), (130, 97), (64, 9), (3840, 2160)])
def test_whole_plane_launch_stores_every_row_once(shape):
    W, H = shape
    for level in range(5):
        c, grid, tiles = cover(W, H, level)
        assert (c == 1).all(), (shape, level)
        lattice_rows = -(-H // (1 << level))
        assert tiles <= grid and tiles >= c.shape[1] * min(1 << level, H) * (lattice_rows // 4) // 2


def test_argument_errors():
    lib = _lib.load()
    buf = (ctypes.c_int * 16)()
    assert lib.rmd_debug_level_cover(64, 4, 0, 0, 4, 0, 0, 0, 0, 0, 0, -1, None, None, None) == -1
    assert lib.rmd_debug_level_cover(64, 4, 5, 0, 4, 0, 0, 0, 0, 0, 0, -1, buf, None, None) == -3
    assert lib.rmd_debug_level_cover(64, 4, 0, 2, 4, 0, 0, 0, 0, 0, 0, -1, buf, None, None) == -2
    assert lib.rmd_debug_level_cover(64, 4, 0, 0, 4, 0, 0, 0, 0, 0, 0, 2, buf, None, None) == -3   # no such variant


@pytest.mark.parametrize("band", [(620, 40, 540), (2200, 0, 2160), (2200, 40, 2160), (1160, 40, 1080), (153, 40, 73)])
def test_band_launch_stores_exactly_its_rows(band):
    E, o0, own = band      # context rows, first owned row, owned rows (as rmd_svgf_band_configure takes them)
    W = 7680 if E > 200 else 512
    for level in range(5):
        c, _, _ = cover(W, E, level, o0, own)
        assert (c[o0:o0 + own] == 1).all() and c[:o0].sum() == 0 and c[o0 + own:].sum() == 0, (band, level)


@pytest.mark.parametrize("band", [(620, 40, 540, True, True), (2200, 0, 2160, False, True), (2200, 40, 2160, True, False),
                                  (1160, 40, 1080, True, True), (300, 40, 220, True, True)])
def test_boundary_and_interior_launches_partition_the_band(band):
    """What rmd_svgf_band_stage launches per level: edge ranges of nb = 2 * 2^(l+1) + 1 rows (21 at level 0) at the
    band's first / last owned rows, where a neighbour exists."""
    E, o0, own, up, down = band
    o1 = o0 + own
    for level in range(4):   # the last level is never split
        nb = 21 if level == 0 else 2 * (2 << level) + 1
        edges = ((o0, nb if up else 0), (o1 - nb, nb if down else 0))
        whole, grid0, _ = cover(7680, E, level, o0, own)
        edge, grid1, tiles1 = cover(7680, E, level, o0, own, split=1, edges=edges)
        rest, _, _ = cover(7680, E, level, o0, own, split=2, edges=edges)
        assert ((edge + rest) == whole).all() and (edge * rest == 0).all(), (band, level)
        for (e0, n) in edges:          # every edge row is in the boundary launch
            assert (edge[e0:e0 + n] == 1).all(), (band, level)
        if own >= 540:                 # and that launch is a fraction of the level
            assert grid1 <= grid0 // 4 and tiles1 <= grid1, (band, level, grid0, grid1)


def test_every_variant_covers_the_same_rows_in_either_direction():
    W, E, o0, own = 1000, 620, 40, 540
    for level in range(5):
        nb = 2 * (2 << level) + 1
        edges = ((o0, nb), (o0 + own - nb, nb))
        base = [cover(W, E, level, o0, own, split=s, edges=edges, variant=6)[0] for s in (0, 1, 2)]
        for v in build.ATROUS_VARIANTS:
            for rev in (0, 1):
                for s in (0, 1, 2):
                    got, _, _ = cover(W, E, level, o0, own, split=s, edges=edges, reverse=rev, variant=v)
                    assert np.array_equal(got, base[s]), (level, v, rev, s)
