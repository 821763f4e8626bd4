"""The search planner (host code of librbod.so, no GPU needed): for every shape the tensor-core pass accepts, the work
decomposition it chooses respects the kernel's hard limits -- shared memory, candidate-list and merge capacities,
pipeline depth, the one-CTA-per-SM cooperative grid -- and covers the whole problem."""
import ctypes
import itertools

import pytest

from retrieval_based_object_detection_b200 import _native

SMS, SMEM_OPTIN = 148, 232448          # B200


def _plan(dim, rows, Q, k, variant=0):
    lib = _native.load()
    out = (ctypes.c_int64 * 13)()
    rc = lib.rbod_debug_plan(dim, rows, Q, k, variant, SMS, SMEM_OPTIN, out)
    return rc, dict(zip(("kc", "slices", "grid", "num_qt", "tiles", "stages", "kbs", "a_tmem_kb", "smem", "list_cap",
                         "list_stride", "final_cap", "n_cap"), list(out)))


DIMS = (3, 64, 100, 128, 320, 512, 640, 768)
ROWS = (1, 127, 128, 129, 4_000, 40_960, 41_000, 1_000_000, 12_500_000, 100_000_000)
QS = (1, 8, 9, 128, 129, 1_000, 10_000, 65_536)
KS = (1, 5, 10, 11, 40, 41, 100, 125, 128)


@pytest.mark.parametrize("variant", [0, 1, 2])
def test_every_accepted_shape_gets_a_plan_within_the_kernel_limits(variant):
    seen_kbs = set()
    for dim, rows, Q, k in itertools.product(DIMS, ROWS, QS, KS):
        rc, p = _plan(dim, rows, Q, k, variant)
        assert rc == 0, (dim, rows, Q, k, variant, _native.load().rbod_last_error())
        per_unit = 256 if variant == 2 else 128
        assert p["kc"] >= k and p["kc"] <= 128 and p["kc"] % 8 == 0
        assert p["tiles"] == (rows + 127) // 128 and p["num_qt"] == (Q + per_unit - 1) // per_unit
        assert 1 <= p["slices"] <= min(p["tiles"], 512) and p["slices"] * p["kc"] <= 8192   # the finish kernel holds <= 8192 keys
        # candidate lists: pruned at list_cap >= 2 kc back to ~kc, a tile may append 128 entries before the check;
        # what reaches the finish kernel fits its key buffer
        assert p["list_cap"] in (64, 128, 256) and p["list_cap"] >= 2 * p["kc"] and p["list_stride"] == p["list_cap"] + 128
        assert p["kc"] <= p["final_cap"] < p["list_stride"] and p["n_cap"] == min(8192, p["slices"] * p["final_cap"])
        assert p["slices"] * p["final_cap"] <= 8192 or p["final_cap"] == p["kc"]
        assert 1 <= p["grid"] <= SMS and (variant != 2 or p["grid"] % 2 == 0)            # cooperative launch: <= 1 CTA / SM
        assert p["grid"] <= p["slices"] * p["num_qt"] * (2 if variant == 2 else 1)
        assert p["smem"] <= SMEM_OPTIN
        assert p["stages"] >= (1 if variant == 1 else 2) and p["stages"] <= 8
        assert p["kbs"] in (2, 4) and (variant != 2 or p["kbs"] == 4) and (variant != 1 or p["kbs"] == 2)
        num_kb = (dim + 63) // 64
        assert 0 <= p["a_tmem_kb"] <= num_kb and (variant == 1) == (p["a_tmem_kb"] == 0)
        if variant != 1:                                   # query tile in TMEM next to >= 1 accumulator of 128 columns
            assert p["a_tmem_kb"] * 32 + 128 <= 512
        seen_kbs.add(p["kbs"])
    assert seen_kbs == ({2, 4} if variant == 0 else ({4} if variant == 2 else {2}))


def test_planner_policies():
    # tiny galleries keep k + 3 candidates (rounded to 8), larger ones the fixed lists
    assert _plan(768, 10_000, 10_000, 5)[1]["kc"] == 8
    assert _plan(768, 10_000, 10_000, 10)[1]["kc"] == 16
    assert _plan(768, 10_000, 10_000, 128)[1]["kc"] == 128
    assert _plan(768, 10_000_000, 10_000, 10)[1]["kc"] == 32
    assert _plan(768, 10_000_000, 10_000, 100)[1]["kc"] == 128
    # coarse stages from two query tiles up (never at the price of the hybrid layout), fine stages for a single tile
    head = _plan(768, 10_000_000, 10_000, 10)[1]
    assert head["kbs"] == 4 and head["stages"] == 2 and head["a_tmem_kb"] == 8 and head["slices"] == 11
    assert _plan(768, 10_000_000, 256, 10)[1]["kbs"] == 4
    assert _plan(768, 10_000_000, 128, 10)[1]["kbs"] == 2            # one query tile: HBM-bound
    shard = _plan(768, 1_250_000, 10_000, 10)[1]                     # the headline gallery's 8-GPU shard
    assert shard["kbs"] == 4 and shard["a_tmem_kb"] == 8
    assert _plan(512, 1_000_000, 10_000, 10)[1]["kbs"] == 4
    # candidate lists live in global memory: k = 100 gets the same pipeline as k = 10 (round 1: 128 KB of heaps in
    # shared memory cost it the hybrid layout and the second accumulator buffer)
    k100 = _plan(768, 10_000_000, 10_000, 100)[1]
    assert {x: k100[x] for x in ("kbs", "stages", "a_tmem_kb", "smem")} == {x: head[x] for x in ("kbs", "stages", "a_tmem_kb", "smem")}
    assert (head["list_cap"], k100["list_cap"], k100["list_stride"]) == (64, 256, 384)
    # left to itself (variant -1) the planner pairs CTAs (256 queries per unit, even grid) above 128 queries
    auto = _plan(768, 10_000_000, 10_000, 10, variant=-1)[1]
    assert auto["num_qt"] == 40 and auto["grid"] == 148 and auto["kbs"] == 4 and auto["a_tmem_kb"] == 8 and auto["stages"] >= 4
    assert _plan(768, 10_000_000, 128, 10, variant=-1)[1]["num_qt"] == 1 and _plan(768, 10_000_000, 129, 10, variant=-1)[1]["num_qt"] == 1
    assert _plan(768, 10_000_000, 128, 10, variant=-1)[1]["kbs"] == 2 and _plan(768, 10_000_000, 129, 10, variant=-1)[1]["kbs"] == 4
    # small batches spread one query tile over (almost) every SM
    assert _plan(768, 12_500_000, 1, 10)[1]["slices"] >= 140
    # what the tensor-core pass does not take is refused here (rbod_search routes it to the fp64 sweep)
    # rows wider than 768 columns stream the query tile through shared memory (variant 1) whatever variant was asked for
    for dim in (1024, 1280, 2048):
        wide = _plan(dim, 1_000_000, 10_000, 10, variant=0)[1]
        assert wide["a_tmem_kb"] == 0 and wide["kbs"] == 2 and wide["stages"] >= 2 and wide["smem"] <= SMEM_OPTIN
    assert _plan(2049, 1000, 1, 1)[0] == _native.RBOD_E_UNSUPPORTED
    assert _plan(512, 1000, 1, 129)[0] == _native.RBOD_E_UNSUPPORTED
