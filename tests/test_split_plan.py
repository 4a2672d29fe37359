"""
CPU-only: the split decision of `launch(...; bc)` in the LIBRARY (chmy_launch_split_plan = the function chmy_launch's
region orchestration calls), for the BASELINE sizes and the slab-width preferences of every kernel family.

  * EXACT_SPLIT: the widths are the Launcher's outer_width, i.e. the regions are literally KernelLaunch.jl:56-87;
  * hint mode (default on connected ranks): whatever the widths, the inner region and the 2N slabs built from them with
    the reference's region formulas tile the launch range [0, n+1]^N exactly once, every slab is at least 3 wide (halo,
    boundary node and send plane of every batch lie inside the slabs, so the batches may run before the inner region has
    finished), and the inner region and the right x slab start on even x indices (the tuned kernels own aligned pairs);
  * no neighbour, an outer_width too small / too large, or the no-split policy: one full-range kernel.
"""
import ctypes as C
import itertools

import numpy as np
import pytest


class _NoArch:
    pass


@pytest.fixture(scope="module")
def ch():
    import chmy_b200
    chmy_b200.load_library()
    return chmy_b200


def desc(ch, n, ow, connected, exact=False):
    """update_thermal!-shaped launch (T Neumann everywhere, exchange T) on a grid whose `connected` sides have neighbours"""
    from chmy_b200 import _lib as L
    nd = len(n)
    topo = tuple(tuple(ch.Connected() if (D, S) in connected else ch.Bounded() for S in range(2)) for D in range(nd))
    g = ch.UniformGrid(_NoArch(), origin=(0.0,) * nd, extent=(1.0,) * nd, dims=n, topology=topo)
    T, To = ch.Field.shell(g), ch.Field.shell(g)
    q = ch.FieldTuple(**{"xyz"[D]: ch.Field.shell(g, ch.vector_location(D + 1, nd)) for D in range(nd)})
    la = ch.Launcher(_NoArch(), g, outer_width=ow, exact_split=exact)
    d = la.describe(None, g, (ch.update_thermal_, (T, To, q, 0.1, g)), bc=ch.batch(g, (T, ch.Neumann()), exchange=T))
    L.check(L.lib().chmy_validate_launch(C.byref(d)))
    return d, (T, To, q)


def plan(ch, d, pref, overlap=1):
    from chmy_b200 import _lib as L
    split, wl, wr = C.c_int32(), (C.c_int32 * 3)(), (C.c_int32 * 3)()
    p = None if pref is None else (C.c_int32 * 3)(*pref)
    L.check(L.lib().chmy_launch_split_plan(C.byref(d), p, overlap, C.byref(split), wl, wr))
    return bool(split.value), list(wl), list(wr)


def regions(n, wl, wr):
    """KernelLaunch.jl:60-87 with outer_width -> (wl, wr): inner + for D = N..1 the two slabs; boxes as (lo, size)"""
    N, ws = len(n), [x + 2 for x in n]
    out = [([wl[a] for a in range(N)], [ws[a] - wl[a] - wr[a] for a in range(N)])]
    for D in reversed(range(N)):
        for S in range(2):
            lo = [0 if a < D else ((0 if S == 0 else ws[a] - wr[a]) if a == D else wl[a]) for a in range(N)]
            sz = [ws[a] if a < D else ((wl[a] if S == 0 else wr[a]) if a == D else ws[a] - wl[a] - wr[a]) for a in range(N)]
            out.append((lo, sz))
    return out


def tiles_once(n, regs):
    """exact cover of [0, n+1]^N, checked per dimension-product without materialising 770^3 cells: total volume equals the
    range's volume and the boxes are pairwise disjoint"""
    ws = [x + 2 for x in n]
    vol = sum(int(np.prod(sz)) for _, sz in regs)
    if vol != int(np.prod(ws)):
        return False
    for (la, sa), (lb, sb) in itertools.combinations(regs, 2):
        if all(la[a] < lb[a] + sb[a] and lb[a] < la[a] + sa[a] for a in range(len(n))):
            return False
    return all(all(l >= 0 and l + s <= w and s > 0 for l, s, w in zip(lo, sz, ws)) for lo, sz in regs)


SIZES = [((767, 767, 767), (128, 8, 4)), ((766, 767, 765), (128, 8, 4)), ((8191, 8191), (128, 8)), ((16383, 16383), (128, 8)),
         ((256, 256), (16, 8)), ((30, 22, 14), (4, 3, 3)), ((125, 64, 20), (9, 4, 3)), ((62, 130), (7, 5))]
PREFS = [None, (60, 6, 0), (64, 0, 0), (60, 0, 0)]       # generic/tuned, a 60 x 6 tile, 3D thermal sweep, 2D sweeps


@pytest.mark.parametrize("n,ow", SIZES)
def test_hint_mode_tiles_the_range_with_even_x_starts(ch, n, ow):
    nd = len(n)
    for conn in ({(0, 1)}, {(0, 0), (0, 1), (1, 0)}, {(D, S) for D in range(nd) for S in range(2)}):
        d, keep = desc(ch, n, ow, conn)
        for pref in PREFS:
            split, wl, wr = plan(ch, d, pref if pref is None else list(pref)[:3])
            assert split, (n, ow, conn, pref)
            wl, wr = wl[:nd], wr[:nd]
            assert all(w >= 3 for w in wl + wr), (wl, wr)
            assert wl[0] % 2 == 0 and (n[0] + 2 - wr[0]) % 2 == 0, (wl, wr)
            assert tiles_once(n, regions(n, wl, wr)), (n, ow, pref, wl, wr)
            if pref is None:
                assert all(abs(a - b) <= 1 for a, b in zip(wl, ow)) and all(b - 1 <= a <= b + 2 for a, b in zip(wr, ow))


@pytest.mark.parametrize("n,ow", SIZES)
def test_exact_split_is_the_reference_region_algebra(ch, n, ow):
    d, keep = desc(ch, n, ow, set(), exact=True)                     # honoured even without a neighbour (tests, A/B)
    for pref in PREFS:
        split, wl, wr = plan(ch, d, pref)
        assert split and wl[:len(n)] == list(ow) and wr[:len(n)] == list(ow)
        assert tiles_once(n, regions(n, ow, ow))


def test_when_there_is_no_split(ch):
    from chmy_b200 import _lib as L
    n, ow = (767, 767, 767), (128, 8, 4)

    def split_of(n_, ow_, conn, pref):
        d_, keep_ = desc(ch, n_, ow_, conn)          # the fields must outlive the descriptor that points at them
        return plan(ch, d_, pref)[0]

    assert split_of(n, ow, set(), (60, 6, 0)) is False                   # no neighbour: nothing to overlap
    assert split_of(n, None, {(0, 1)}, None) is False                    # Launcher without outer_width
    assert split_of(n, (128, 2, 4), {(0, 1)}, None) is False             # a slab thinner than halo + node + send plane
    assert split_of((30, 22, 14), (17, 8, 4), {(0, 1)}, None) is False   # the two x slabs would overlap
    assert split_of((30, 22, 14), (16, 8, 4), {(0, 1)}, None) is True    # they just touch: an empty inner region
    d, keep = desc(ch, n, ow, {(0, 1)})
    assert plan(ch, d, (60, 6, 0), overlap=0)[0] is False                # bench.py --split off: one kernel, then the batches
    de, keep2 = desc(ch, n, ow, {(0, 1)}, exact=True)
    assert plan(ch, de, None, overlap=0)[0] is True                      # a literal split stays literal
    assert plan(ch, d, (60, 6, 0), overlap=1)[0] is True                 # the default: overlapped


def test_tiny_grid_whose_nudged_slabs_do_not_fit_runs_unsplit(ch):
    """n[0]=4, outer_width[0]=3 -> slabs 4 + 3 wide do not fit 6 cells; the plan must not fall back to slabs on
    odd x origins (the sweeps own aligned pairs of cells) -- it reports one full-range launch instead."""
    d, keep = desc(ch, (4, 40, 40), (3, 4, 4), {(0, 1)})
    split, wl, wr = plan(ch, d, None)
    assert split is False and wl == [0, 0, 0] and wr == [0, 0, 0]


def _tiles(ch, g, i0, i1, tail):
    from chmy_b200 import _lib as L
    n = g[0] * g[1] * g[2]
    out = (C.c_int32 * (4 * n))()
    L.check(L.lib().chmy_selftest_tile_order((C.c_int32 * 3)(*g), (C.c_int32 * 3)(*i0), (C.c_int32 * 3)(*i1), tail, out))
    return [tuple(out[4 * c:4 * c + 4]) for c in range(n)]


@pytest.mark.parametrize("g,i0,i1", [((13, 35, 13), (1, 1, 1), (12, 34, 12)), ((13, 35, 13), (0, 0, 0), (13, 35, 13)),
                                     ((3, 2, 1), (1, 1, 0), (2, 1, 0)), ((5, 4, 6), (2, 0, 1), (3, 4, 6)), ((1, 1, 1), (0, 0, 0), (0, 0, 0)),
                                     ((7, 9, 4), (1, 2, 1), (7, 8, 3)), ((6, 1, 5), (1, 0, 2), (5, 1, 4)), ((4, 4, 4), (2, 2, 2), (2, 2, 2)),
                                     ((13, 35, 2), (1, 1, 1), (12, 34, 1)), ((2, 2, 3), (0, 0, 1), (2, 2, 2))])
def test_tile_order_of_an_overlapped_sweep(ch, g, i0, i1):
    """the fused sweep's launch order (tile_decode, fused_sv.cuh): every tile exactly once in either order; boundary tiles --
    those outside the interior index box -- flagged as such (their CTAs feed the retire counter the boundary stream sleeps
    on: a tile missing from the count would let the batches start early, one too many would hang); and in the overlapped
    order no boundary tile is launched after the interior of the final layer (that stretch is what hides the exchange)"""
    allt = sorted(itertools.product(range(g[0]), range(g[1]), range(g[2])))
    inside = lambda x: all(i0[a] <= x[a] < i1[a] for a in range(3))
    nat = _tiles(ch, g, i0, i1, 0)
    assert [x[:3] for x in nat] == [(x, y, z) for z in range(g[2]) for y in range(g[1]) for x in range(g[0])]
    t = _tiles(ch, g, i0, i1, 1)
    for order in (nat, t):
        assert sorted(x[:3] for x in order) == allt
        assert all(bool(x[3]) == (not inside(x)) for x in order)
    flags = [x[3] for x in t]
    if any(flags):
        last_b = max(i for i, f in enumerate(flags) if f)
        tail_z = g[2] - 2 if g[2] >= 2 else 0              # the layer that runs last
        n_tail_int = sum(1 for x in t if not x[3] and x[2] == tail_z)
        assert last_b == len(t) - n_tail_int - 1, (last_b, len(t), n_tail_int)
        assert all(x[2] == tail_z for x in t[last_b + 1:])
    # everything but the final layer keeps the natural order (neighbours in time are neighbours in space)
    layer = g[0] * g[1]
    zs = [g[2] - 1] + list(range(g[2] - 1))
    for ci, z in enumerate(zs[:-1]):
        assert [x[:3] for x in t[ci * layer:(ci + 1) * layer]] == [(x, y, z) for y in range(g[1]) for x in range(g[0])]


# ---------------------------------------------------------------------------------------------- fuzz
def test_fuzz_any_plan_tiles_the_range_exactly_once(ch):
    """hypothesis over dimensionality, sizes, outer_width, kernel preference, connectivity and the exact flag: whatever the
    library decides, it never crashes, and a split it reports always tiles [0, n+1]^N exactly once with non-negative boxes
    (hint mode additionally: slabs >= 3 wide, even x starts) -- i.e. no cell is computed twice or skipped for ANY input."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @st.composite
    def cases(draw):
        nd = draw(st.integers(2, 3))                  # the op used by desc() (update_thermal!) exists in 2D and 3D
        n = tuple(draw(st.integers(1, 900)) for _ in range(nd))
        ow = tuple(draw(st.integers(0, x + 3)) for x in n)
        pref = draw(st.one_of(st.none(), st.tuples(st.integers(0, 130), st.integers(0, 20), st.integers(0, 9))))
        conn = {(D, S) for D in range(nd) for S in range(2) if draw(st.booleans())}
        return n, ow, pref, conn, draw(st.booleans())

    @settings(max_examples=400, deadline=None, suppress_health_check=list(HealthCheck))
    @given(cases())
    def check(case):
        n, ow, pref, conn, exact = case
        nd = len(n)
        d, keep = desc(ch, n, ow, conn, exact=exact)
        split, wl, wr = plan(ch, d, None if pref is None else list(pref))
        if not split:
            assert wl == [0, 0, 0] and wr == [0, 0, 0]
            return
        wl, wr = wl[:nd], wr[:nd]
        ws = [x + 2 for x in n]
        regs = [(lo, sz) for lo, sz in regions(n, wl, wr)]
        assert all(s >= 0 for _, sz in regs for s in sz), (case, wl, wr)
        assert all(l >= 0 and l + s <= w for lo, sz in regs for l, s, w in zip(lo, sz, ws)), (case, wl, wr)
        full = [(lo, sz) for lo, sz in regs if all(s > 0 for s in sz)]
        assert sum(int(np.prod(sz)) for _, sz in full) == int(np.prod(ws)), (case, wl, wr)
        for (la, sa), (lb, sb) in itertools.combinations(full, 2):
            assert not all(la[a] < lb[a] + sb[a] and lb[a] < la[a] + sa[a] for a in range(nd)), (case, wl, wr)
        if exact:
            assert wl == list(ow) and wr == list(ow)
        else:
            assert all(w >= 3 for w in wl + wr), (case, wl, wr)
            assert wl[0] % 2 == 0 and (ws[0] - wr[0]) % 2 == 0, (case, wl, wr)

    check()
