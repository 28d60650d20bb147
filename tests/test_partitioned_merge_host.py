"""CPU coverage of the multi-GPU Barnes-Hut path's host logic (no GPU needed): the merge of the
per-part trees of the key-range-partitioned build into one joined tree (particular_b200/csrc/
barneshut.cu: merge_top_tree, reached through the test hook pcuda_debug_merge_top_tree).

The per-part trees are made by the CPU statement of the tree specification (oracle.Octree built in
the frame of the whole cloud), packed exactly as the device kernels fill_pack / collect_boundary
pack them, merged by the library, and the joined tree is then checked structurally: every particle
is reached exactly once from the top root, every internal node carries the mass and centre of mass
of its children, and every merged cell equals the cell of the single tree over all particles."""
import ctypes as C

import numpy as np
import pytest

import oracle
from tests.conftest import plummer_cloud, uniform_cloud

BITS, LEVELS = 21, 22
NODE = np.dtype([("cm", np.float32, 4), ("first_child", np.uint32), ("nchild_level", np.uint32),
                 ("begin", np.uint32), ("count", np.uint32)])
PACK = np.dtype([("n_nodes", np.uint32), ("n_levels", np.uint32), ("level_begin", np.uint32, LEVELS + 2),
                 ("prefix", np.uint64, (LEVELS, 2)), ("mom", np.float64, (LEVELS, 2, 4))])
BOUND = np.dtype([("node", NODE), ("child", NODE, 8)])


def test_layouts_match_the_device_structs():
    assert NODE.itemsize == 32 and PACK.itemsize == 1864 and BOUND.itemsize == 9 * 32


def join(p, parts, nleaf=16):
    """Per-part trees in the common frame, joined arrays, packs and boundary records."""
    full = oracle.Octree(p, nleaf)
    frame = (full.origin, full.ext, full.inv)
    keys_in = oracle.morton_keys(p[:, :3], full.origin, full.inv)
    ks = np.sort(keys_in)
    split = [0] + [int(ks[q * len(p) // parts]) for q in range(1, parts)] + [1 << 64]
    nodes, keys, node_base, part_base = [], [], [], []
    packs = np.zeros(parts, PACK)
    per_part = []
    for q in range(parts):
        idx = np.flatnonzero((keys_in >= np.uint64(split[q])) & (keys_in.astype(object) < split[q + 1]))
        node_base.append(sum(len(x) for x in nodes))
        part_base.append(sum(len(x) for x in keys))
        if len(idx) == 0:
            per_part.append(None)
            continue
        t = oracle.Octree(p[idx], nleaf, frame=frame)
        assert np.array_equal(t.keys, np.sort(keys_in[idx]))
        rec = np.zeros(t.n_nodes, NODE)
        rec["cm"] = t.commass
        internal = t.n_child > 0
        rec["first_child"] = np.where(internal, t.first_child + node_base[q], 0)
        rec["nchild_level"] = t.n_child | (t.level << 8)
        rec["begin"] = t.begin + part_base[q]
        rec["count"] = t.count
        nodes.append(rec)
        keys.append(t.keys)
        lb = np.searchsorted(t.level, np.arange(LEVELS + 2)).astype(np.uint32)
        mom = t.moments()
        packs[q]["n_nodes"], packs[q]["n_levels"] = t.n_nodes, t.n_levels
        packs[q]["level_begin"] = lb
        for l in range(t.n_levels):
            for side, j in enumerate((lb[l], lb[l + 1] - 1)):
                packs[q]["prefix"][l][side] = int(t.keys[t.begin[j]]) >> (3 * (BITS - l))
                packs[q]["mom"][l][side] = mom[j]
        per_part.append((t, lb))
    nodes = np.concatenate(nodes) if nodes else np.zeros(0, NODE)
    keys = np.concatenate(keys) if keys else np.zeros(0, np.uint64)
    stage = np.zeros((parts, LEVELS, 2), BOUND)
    for q, pp in enumerate(per_part):
        if pp is None:
            continue
        t, lb = pp
        for l in range(t.n_levels):
            for side, j in enumerate((lb[l], lb[l + 1] - 1)):
                nd = nodes[node_base[q] + j]
                stage[q, l, side]["node"] = nd
                nc = int(nd["nchild_level"]) & 0xff
                stage[q, l, side]["child"][:nc] = nodes[nd["first_child"]: nd["first_child"] + nc]
    return full, nodes, keys, np.array(node_base, np.uint32), packs, stage


def merge(parts, packs, stage, node_base, top_base):
    from particular_b200 import _ffi
    top = np.zeros(4096, NODE)
    roots = np.zeros(1024, np.uint32)
    n_top, n_roots = C.c_uint32(), C.c_uint32()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    _ffi.check(_ffi.lib.pcuda_debug_merge_top_tree(parts, vp(packs), vp(stage), vp(node_base), top_base, vp(top),
                                                   len(top), C.byref(n_top), vp(roots), len(roots),
                                                   C.byref(n_roots)))
    return top[: n_top.value], roots[: n_roots.value]


def check_joined(p, parts, nleaf=16):
    full, nodes, keys, node_base, packs, stage = join(p, parts, nleaf)
    n = len(p)
    top, roots = merge(parts, packs, stage, node_base, len(nodes))
    assert len(roots) == 1
    joined = np.concatenate([nodes, top])
    ext = max(full.ext, 1e-30)
    single = {}
    for j in range(full.n_nodes):
        l = int(full.level[j])
        single[(l, int(full.keys[full.begin[j]]) >> (3 * (BITS - l)))] = j
    covered = np.zeros(n, np.int32)
    stack, visited, compared = [int(roots[0])], 0, 0
    while stack:
        i = stack.pop()
        nd = joined[i]
        visited += 1
        nc, lvl = int(nd["nchild_level"]) & 0xff, int(nd["nchild_level"]) >> 8
        assert nc <= 8
        if nc == 0:
            covered[nd["begin"]: nd["begin"] + nd["count"]] += 1
            continue
        ch = joined[nd["first_child"]: nd["first_child"] + nc]
        clv = ch["nchild_level"] >> 8
        assert ((clv == lvl) | (clv == lvl + 1)).all()
        m = ch["cm"][:, 3].astype(np.float64)
        assert np.isclose(m.sum(), nd["cm"][3], rtol=2e-6, atol=0), (i, m.sum(), nd["cm"][3])
        if m.sum() != 0:
            com = (ch["cm"][:, :3].astype(np.float64) * m[:, None]).sum(0) / m.sum()
            assert np.abs(com - nd["cm"][:3]).max() <= 2e-6 * ext + 1e-6 * np.abs(com).max(), (i, com, nd["cm"])
        assert int(ch["count"].sum()) == int(nd["count"])
        if i >= len(nodes):  # a node of the top tree: the same cell of the single tree, when it is one
            j = single.get((lvl, int(keys[nd["begin"]]) >> (3 * (BITS - lvl))))
            if j is not None and int(full.count[j]) == int(nd["count"]):
                compared += 1
                assert np.isclose(nd["cm"][3], full.commass[j, 3], rtol=1e-6, atol=0)
                assert np.abs(nd["cm"][:3] - full.commass[j, :3]).max() <= 1e-6 * ext + 1e-6 * np.abs(full.commass[j, :3]).max()
        stack.extend(range(int(nd["first_child"]), int(nd["first_child"]) + nc))
    assert (covered == 1).all(), (int((covered == 0).sum()), int((covered > 1).sum()))
    nonempty = int((packs["n_nodes"] > 0).sum())
    if nonempty > 1:
        assert len(top) >= 1 and compared >= 1 and int(top[0]["count"]) == n
    else:
        assert len(top) == 0
    return len(top), visited


@pytest.mark.parametrize("cloud", ["uniform", "plummer"])
@pytest.mark.parametrize("parts", [1, 2, 3, 8, 16])
def test_joined_tree_is_a_partition_with_consistent_moments(cloud, parts):
    p = uniform_cloud(6000, seed=11) if cloud == "uniform" else plummer_cloud(6000, seed=11)
    n_top, visited = check_joined(p, parts)
    assert n_top <= 1 + 8 * LEVELS * parts


@pytest.mark.parametrize("n,parts,nleaf", [(3, 8, 16), (40, 8, 16), (200, 16, 1), (2500, 5, 4), (1000, 2, 32)])
def test_small_and_ragged_parts(n, parts, nleaf):
    check_joined(plummer_cloud(n, seed=n), parts, nleaf)


def test_coincident_and_clustered_particles():
    """All keys equal (one part gets everything), and two tight clumps with deep single-child chains."""
    same = np.tile(np.array([[1.0, 2.0, 3.0, 5.0]], np.float32), (300, 1))
    check_joined(same, 4)
    rng = np.random.default_rng(5)
    two = np.concatenate([same[:150] + np.concatenate([rng.normal(0, 1e-4, (150, 3)), np.zeros((150, 1))], 1),
                          same[:150] + np.array([[10.0, 0, 0, 0]]) +
                          np.concatenate([rng.normal(0, 1e-4, (150, 3)), np.zeros((150, 1))], 1)]).astype(np.float32)
    check_joined(two, 3)
    check_joined(two, 8, nleaf=2)


def test_zero_mass_particles():
    p = uniform_cloud(3000, seed=3)
    p[::2, 3] = 0.0
    check_joined(p, 4)
    p[:, 3] = 0.0
    check_joined(p, 4)


def test_random_clouds_parts_and_leaf_sizes():
    """Property test: any cloud, any number of parts, any leaf size gives an exact partition."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(n=st.integers(1, 1500), parts=st.integers(1, 16), nleaf=st.sampled_from([1, 2, 8, 16, 32]),
           seed=st.integers(0, 10_000), clumpy=st.booleans())
    def run(n, parts, nleaf, seed, clumpy):
        p = plummer_cloud(n, seed=seed) if clumpy else uniform_cloud(n, seed=seed)
        if clumpy and n > 10:
            p[: n // 3, :3] = p[0, :3]  # a block of coincident particles: equal keys never straddle a cut
        check_joined(p, parts, nleaf)

    run()


# ---- the top tree of the OTHER parts (two-phase walk of the locally essential trees) --------------------
SHARE = 0x80000000


def stage_of(nodes, node_base, packs):
    """Boundary records as the device collects them (after the share marks were set in the node array)."""
    stage = np.zeros((len(packs), LEVELS, 2), BOUND)
    for q, pk in enumerate(packs):
        for l in range(int(pk["n_levels"]) if pk["n_nodes"] else 0):
            lb = pk["level_begin"]
            for side, j in enumerate((lb[l], lb[l + 1] - 1)):
                nd = nodes[node_base[q] + j]
                stage[q, l, side]["node"] = nd
                nc = int(nd["nchild_level"]) & 0xff
                stage[q, l, side]["child"][:nc] = nodes[nd["first_child"]: nd["first_child"] + nc]
    return stage


@pytest.mark.parametrize("cloud", ["uniform", "plummer"])
@pytest.mark.parametrize("parts", [2, 3, 8])
def test_top_tree_of_the_other_parts_keeps_the_share_marks(cloud, parts):
    """bh_multigpu.cu, sharded_let_dev with the two-phase walk: the rank's own part is left out of the merge
    (its pack says n_nodes = 0) and the other parts' nodes that are shares of a cell the own tree also holds
    (same level and prefix as a first / last node of an own level: let_mark_shares_kernel) carry bit 31 of
    nchild_level.  The merged tree must reach every particle of the other parts exactly once, none of the
    own, and every reachable node of such a cell — merged cells and continuation nodes included — must
    still carry the mark, with level and child count intact under it."""
    p = uniform_cloud(5000, seed=21) if cloud == "uniform" else plummer_cloud(5000, seed=21)
    full, nodes0, keys, node_base, packs0, _ = join(p, parts)
    part_of_particle = np.zeros(len(p), np.int32)
    start = 0
    for q in range(parts):  # particles of part q: the next run of the joined key array
        cnt = int(nodes0[node_base[q]]["count"]) if packs0[q]["n_nodes"] else 0
        part_of_particle[start: start + cnt] = q
        start += cnt
    assert start == len(p)
    for own in range(parts):
        if packs0[own]["n_nodes"] == 0:
            continue
        nodes, packs = nodes0.copy(), packs0.copy()
        own_cells = set()
        for l in range(int(packs[own]["n_levels"])):
            own_cells.update({(l, int(packs[own]["prefix"][l][0])), (l, int(packs[own]["prefix"][l][1]))})
        for q in range(parts):  # let_mark_shares_kernel
            if q == own or packs[q]["n_nodes"] == 0:
                continue
            lb = packs[q]["level_begin"]
            for l in range(int(packs[q]["n_levels"])):
                for side, j in enumerate((lb[l], lb[l + 1] - 1)):
                    if (l, int(packs[q]["prefix"][l][side])) in own_cells:
                        nodes[node_base[q] + j]["nchild_level"] |= SHARE
        stage = stage_of(nodes, node_base, packs)
        packs[own]["n_nodes"] = 0
        top, roots = merge(parts, packs, stage, node_base, len(nodes))
        others = int((packs["n_nodes"] > 0).sum())
        if others == 0:
            assert len(roots) == 0 and len(top) == 0
            continue
        assert len(roots) == 1 and (len(top) == 0) == (others == 1)
        joined = np.concatenate([nodes, top])
        covered = np.zeros(len(p), np.int32)
        stack, marked = [int(roots[0])], 0
        while stack:
            i = stack.pop()
            nd = joined[i]
            word = int(nd["nchild_level"])
            nc, lvl = word & 0xff, word >> 8 & 0xff
            assert nc <= 8 and lvl <= BITS and word & 0x7fff0000 == 0
            if nd["count"]:
                cell = (lvl, int(keys[nd["begin"]]) >> (3 * (BITS - lvl)))
                assert bool(word & SHARE) == (cell in own_cells), (own, i, cell)
                marked += bool(word & SHARE)
            if nc == 0:
                covered[nd["begin"]: nd["begin"] + nd["count"]] += 1
                continue
            ch = joined[nd["first_child"]: nd["first_child"] + nc]
            clv = ch["nchild_level"] >> 8 & 0xff
            assert ((clv == lvl) | (clv == lvl + 1)).all()
            assert int(ch["count"].sum()) == int(nd["count"])
            stack.extend(range(int(nd["first_child"]), int(nd["first_child"]) + nc))
        assert (covered[part_of_particle != own] == 1).all() and (covered[part_of_particle == own] == 0).all()
        assert marked >= 1  # the root cell at least is shared


# ---- world_size-2 (and 3) gloo run of the partitioned build's exchange + merge on CPU -----------------
def _free_port():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, n, q):
    """What sharded_forest_dev does after the raw all-gather, with the oracle standing in for the device
    build and gloo for NCCL: own part tree in the common frame -> pack + node slot -> all-gather into
    EQUAL slots (node / particle indices rebased to rank * slot) -> boundary records -> merge."""
    import hashlib
    import os

    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p = plummer_cloud(n, seed=77)
        full = oracle.Octree(p)
        keys_in = oracle.morton_keys(p[:, :3], full.origin, full.inv)
        ks = np.sort(keys_in)
        split = [0] + [int(ks[r * n // world]) for r in range(1, world)] + [1 << 64]
        counts = [int(((keys_in >= np.uint64(split[r])) & (keys_in.astype(object) < split[r + 1])).sum())
                  for r in range(world)]
        slot = max(max(counts), 1)
        idx = np.flatnonzero((keys_in >= np.uint64(split[rank])) & (keys_in.astype(object) < split[rank + 1]))
        t = oracle.Octree(p[idx], frame=(full.origin, full.ext, full.inv))
        lb = np.searchsorted(t.level, np.arange(LEVELS + 2)).astype(np.uint32)
        pack = np.zeros(1, PACK)
        pack["n_nodes"], pack["n_levels"], pack["level_begin"] = t.n_nodes, t.n_levels, lb
        mom = t.moments()
        for l in range(t.n_levels):
            for side, j in enumerate((lb[l], lb[l + 1] - 1)):
                pack["prefix"][0][l][side] = int(t.keys[t.begin[j]]) >> (3 * (BITS - l))
                pack["mom"][0][l][side] = mom[j]
        packs_t = [torch.empty(PACK.itemsize, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(packs_t, torch.from_numpy(pack.view(np.uint8).copy()))
        packs = np.concatenate([x.numpy().view(PACK) for x in packs_t])
        node_slot = int(packs["n_nodes"].max())
        rec = np.zeros(node_slot, NODE)  # own slot, indices rebased to the slot (copy_rebase_nodes)
        rec["cm"][: t.n_nodes] = t.commass
        rec["first_child"][: t.n_nodes] = np.where(t.n_child > 0, t.first_child + rank * node_slot, 0)
        rec["nchild_level"][: t.n_nodes] = t.n_child | (t.level << 8)
        rec["begin"][: t.n_nodes] = t.begin + rank * slot
        rec["count"][: t.n_nodes] = t.count
        nodes_t = [torch.empty(node_slot * NODE.itemsize, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(nodes_t, torch.from_numpy(rec.view(np.uint8).copy()))
        nodes = np.concatenate([x.numpy().view(NODE) for x in nodes_t])
        node_base = (np.arange(world) * node_slot).astype(np.uint32)
        stage = np.zeros((world, LEVELS, 2), BOUND)  # collect_boundary
        for r in range(world):
            for l in range(int(packs[r]["n_levels"])):
                b = packs[r]["level_begin"]
                for side, j in enumerate((b[l], b[l + 1] - 1)):
                    nd = nodes[node_base[r] + j]
                    stage[r, l, side]["node"] = nd
                    nc = int(nd["nchild_level"]) & 0xff
                    stage[r, l, side]["child"][:nc] = nodes[nd["first_child"]: nd["first_child"] + nc]
        top_base = world * node_slot
        top, roots = merge(world, packs, stage, node_base, top_base)
        joined = np.concatenate([nodes, top])
        covered = np.zeros(world * slot, np.int32)
        stack = [int(roots[0])]
        while stack:
            nd = joined[stack.pop()]
            nc = int(nd["nchild_level"]) & 0xff
            if nc == 0:
                covered[nd["begin"]: nd["begin"] + nd["count"]] += 1
            else:
                stack.extend(range(int(nd["first_child"]), int(nd["first_child"]) + nc))
        want = np.zeros(world * slot, np.int32)
        for r in range(world):
            want[r * slot: r * slot + counts[r]] = 1  # the rest of a slot is padding
        ok = bool(np.array_equal(covered, want)) and len(roots) == 1 and int(top[0]["count"]) == n
        ok = ok and np.isclose(top[0]["cm"][3], full.commass[0, 3], rtol=1e-6)
        q.put((rank, ok, hashlib.sha1(top.tobytes() + roots.tobytes()).hexdigest()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 5000), (3, 1200)])
def test_gloo_ranks_build_the_same_joined_tree(world, n):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, world, port, n, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert len({h for _, _, h in res}) == 1, res  # every rank merged the identical top tree
