"""Integer logic of the fused GEMM -> all-reduce protocol and of the split-K stage partition, restated from
csrc/gemm.cu / csrc/tp_reduce.cu and checked exhaustively on the CPU: the producer side (GEMM epilogue: tile -> owner rank,
owned index) and the consumer side (reducer: rank, owned index -> tile) must be inverse bijections for every tile count,
tp and CTA-group count, including a partial last block; the split-K slices must tile the stage list exactly."""
import itertools


def owner_of(tile, tp, rot_s):
    """gemm.cu, RS epilogue: own_idx = tile / tp; owner = (tile + (own_idx * tp) / rot_s) % tp."""
    own_idx = tile // tp
    return (tile + (own_idx * tp) // rot_s) % tp, own_idx


def tile_of(rank, own_idx, tp, rot_s):
    """tp_reduce.cu, tile_allreduce_kernel: tile = own_idx * TP + ((rank - (own_idx * TP) / rot_s) % TP + TP) % TP."""
    return own_idx * tp + ((rank - (own_idx * tp) // rot_s) % tp + tp) % tp


def rot_period(groups, tp):
    """gemm.cu, matmul_impl: rot_s = ceil(groups / tp) * tp."""
    return (groups + tp - 1) // tp * tp


def test_owner_mapping_is_a_bijection():
    for tp, groups, num_tiles in itertools.product((1, 2, 4, 8), (1, 7, 58, 66, 74, 148), (1, 2, 3, 15, 16, 17, 74, 75, 512, 513)):
        rot_s = rot_period(groups, tp)
        assert rot_s % tp == 0 and rot_s >= groups
        seen = {}
        for tile in range(num_tiles):
            owner, own_idx = owner_of(tile, tp, rot_s)
            assert 0 <= owner < tp
            assert tile_of(owner, own_idx, tp, rot_s) == tile       # the reducer finds exactly this tile
            assert (owner, own_idx) not in seen                     # one tile per (rank, owned index)
            seen[(owner, own_idx)] = tile
        blocks = (num_tiles + tp - 1) // tp                         # every rank walks this many owned indices ...
        for rank in range(tp):
            for own_idx in range(blocks):
                t = tile_of(rank, own_idx, tp, rot_s)
                assert (t < num_tiles) == ((rank, own_idx) in seen)  # ... and skips exactly the tiles that do not exist


def test_owner_rotation_spreads_a_cta_groups_tiles():
    """With tp = 2 and an even number of CTA groups, tile % tp would give a group the same owner for all its tiles."""
    tp, groups = 2, 74
    rot_s = rot_period(groups, tp)
    for g in range(groups):
        owners = [owner_of(t, tp, rot_s)[0] for t in range(g, 512, groups)]
        assert len(set(owners)) == tp and abs(owners.count(0) - owners.count(1)) <= 1


def splitk_slices(ktiles, sk):
    """gemm.cu: CTA r of the cluster takes stages [r * T / SK, (r + 1) * T / SK) of the concatenated stage list."""
    total = sum(ktiles)
    out = []
    for r in range(sk):
        lo, hi = r * total // sk, (r + 1) * total // sk
        mine, off = [], 0
        for s, kt in enumerate(ktiles):
            a, b = max(0, lo - off), min(kt, hi - off)
            mine += [(s, k) for k in range(a, b)]
            off += kt
        out.append(mine)
    return out


def test_splitk_slices_partition_the_stage_list():
    for ktiles in ((10, 8, 4), (35, 28, 14), (1, 1, 1), (16,), (0, 8, 0), (3, 0, 5), (11, 1, 1)):
        kt = [k for k in ktiles]
        total = sum(kt)
        every = [(s, k) for s, n in enumerate(kt) for k in range(n)]
        for sk in (2, 4, 8):
            if total < sk:
                continue
            slices = splitk_slices(kt, sk)
            assert [x for sl in slices for x in sl] == every       # exact cover, in order
            assert all(len(sl) >= 1 for sl in slices)              # every CTA issues its tile-complete commit
            assert max(len(sl) for sl in slices) - min(len(sl) for sl in slices) <= 1
