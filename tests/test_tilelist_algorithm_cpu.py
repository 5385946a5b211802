"""The tile-list construction of mtgs_b200/csrc/tilelists.cu, emulated level by level in numpy and compared with the
oracle's stable (tile, depth) sort (upstream isect_tiles + radix sort + isect_offset_encode).

This pins the ALGORITHM on the CPU, independent of CUDA: the same geometry rule (row groups / column groups of at
most 16 children, 32 for very large grids), the same four levels of order-preserving filters with chunked count ->
prefix -> fill, and the same child -> tile mapping.  The GPU parity tests then only have to show that the kernels
implement it."""
import numpy as np
import pytest

from mtgs_b200 import scenes


def tl_shift(n, lo):
    for s in range(lo, 6):
        if (n + (1 << s) - 1) >> s <= 16:
            return s
    return 5 if (n + 31) >> 5 <= 32 else -1


def bins(level, lo, hi, L, rg, cg, ncg):
    """children [a, b) of list L covered by an item with packed range [lo, hi) (tilelists.cu: tl_bins)."""
    if level == 1:
        return lo >> rg, (((hi - 1) >> rg) + 1 if hi > lo else lo >> rg)
    if level == 2:
        base = L << rg
        return max(lo - base, 0), min(hi - base, 1 << rg)
    if level == 3:
        return lo >> cg, (((hi - 1) >> cg) + 1 if hi > lo else lo >> cg)
    base = (L % ncg) << cg
    return max(lo - base, 0), min(hi - base, 1 << cg)


def run_level(level, lists, nb, geom, chunk):
    """lists: list of item lists [(gid, (x0, x1), (y0, y1)), ...] in stream order -> child lists, via chunked
    count / prefix / fill exactly as the kernels do it (so chunk boundaries and prefixes are exercised)."""
    rg, cg, ncg = geom
    out = [[] for _ in range(len(lists) * nb)]
    for L, items in enumerate(lists):
        chunks = [items[i:i + chunk] for i in range(0, len(items), chunk)]
        counts = np.zeros((len(chunks), nb), np.int64)
        for c, ch in enumerate(chunks):                      # count
            for (_, xr, yr) in ch:
                r = yr if level <= 2 else xr
                a, b = bins(level, r[0], r[1], L, rg, cg, ncg)
                counts[c, a:b] += 1
        base = np.cumsum(counts, 0) - counts                 # prefix over the chunks of the list
        filled = [[None] * int(counts[:, b].sum()) for b in range(nb)]
        for c, ch in enumerate(chunks):                      # fill (chunks may run in any order)
            cur = base[c].copy()
            for it in ch:
                r = it[2] if level <= 2 else it[1]
                a, b = bins(level, r[0], r[1], L, rg, cg, ncg)
                for k in range(a, b):
                    filled[k][cur[k]] = it
                    cur[k] += 1
        for b in range(nb):
            assert None not in filled[b]
            out[L * nb + b] = filled[b]
    return out


def tile_lists(order, rects, tile_w, tile_h, chunk=7):
    rg, cg = tl_shift(tile_h, 3), tl_shift(tile_w, 4)
    assert rg >= 0 and cg >= 0
    nrg, ncg = (tile_h + (1 << rg) - 1) >> rg, (tile_w + (1 << cg) - 1) >> cg
    geom = (rg, cg, ncg)
    l0 = [[(int(g), tuple(rects[g, 0]), tuple(rects[g, 1])) for g in order]]
    l1 = run_level(1, l0, nrg, geom, chunk)
    l2 = run_level(2, l1, 1 << rg, geom, chunk)
    l3 = run_level(3, l2, ncg, geom, chunk)
    l4 = run_level(4, l3, 1 << cg, geom, chunk)
    flat, offs = [], np.zeros(tile_h * tile_w, np.int64)
    for y in range(tile_h):
        for x in range(tile_w):
            L = y * ncg + (x >> cg)
            offs[y * tile_w + x] = len(flat)
            flat += [it[0] for it in l4[L * (1 << cg) + (x & ((1 << cg) - 1))]]
    # children that do not map to a tile (padding rows / columns) must be empty
    total = sum(len(c) for c in l4)
    assert total == len(flat)
    return np.array(flat, np.int64), offs


@pytest.mark.parametrize("scene,chunk", [(lambda: scenes.tiny(n=300, seed=3, width=64, height=48), 7),
                                         (lambda: scenes.tiny(n=257, seed=9, width=77, height=53), 3),
                                         (lambda: scenes.street(n=1500, seed=1, width=640, height=360), 64),
                                         (lambda: scenes.street(n=400, seed=4, width=4400, height=304), 16),
                                         (lambda: scenes.street(n=400, seed=5, width=304, height=4400), 16)])
def test_filter_hierarchy_equals_stable_sort(oracle, scene, chunk):
    s = scene()
    _, _, ref, _ = oracle.rasterization(s["means"], s["quats"], s["scales"], s["opacities"], s["colors"], s["viewmat"],
                                        s["K"], s["width"], s["height"], render_mode="RGB", rasterize_mode="classic")
    tile_w, tile_h = ref["tile_width"], ref["tile_height"]
    radii, m2, depths = ref["radii"], ref["means2d"], ref["depths"]
    vis = np.nonzero(radii > 0)[0]
    # tile rectangles as the projection kernel packs them (tile size 16)
    rects = np.zeros((len(radii), 2, 2), np.int64)
    r = radii[vis].astype(np.float32) / np.float32(16)
    tx, ty = m2[vis, 0] / np.float32(16), m2[vis, 1] / np.float32(16)
    rects[vis, 0, 0] = np.clip(np.floor(tx - r), 0, tile_w)
    rects[vis, 0, 1] = np.clip(np.ceil(tx + r), 0, tile_w)
    rects[vis, 1, 0] = np.clip(np.floor(ty - r), 0, tile_h)
    rects[vis, 1, 1] = np.clip(np.ceil(ty + r), 0, tile_h)
    np.testing.assert_array_equal((rects[:, 0, 1] - rects[:, 0, 0]) * (rects[:, 1, 1] - rects[:, 1, 0]),
                                  ref["tiles_per_gauss"])
    # stable depth order of the visible Gaussians (depthsort.cu): float bits of positive depths, ties by id
    keys = depths[vis].view(np.uint32)
    order = vis[np.argsort(keys, kind="stable")]
    flat, offs = tile_lists(order, rects, tile_w, tile_h, chunk)
    np.testing.assert_array_equal(flat, ref["flatten_ids"])
    np.testing.assert_array_equal(offs, ref["isect_offsets"].reshape(-1))


def test_geometry_rule():
    assert (tl_shift(68, 3), tl_shift(120, 4)) == (3, 4)      # 1080p: 9 row groups of 8, 8 column groups of 16
    assert (tl_shift(135, 3), tl_shift(240, 4)) == (4, 4)     # 4K: 9 row groups of 16, 15 column groups
    assert tl_shift(275, 3) == 5 and tl_shift(1024, 3) == 5 and tl_shift(1025, 3) == -1
