"""GPU parity of the tcgen05 tokeniser (fp16 candidate filter + exact fp32 re-rank) against the oracle, the
reference-generated fixtures and the exact CUDA-core kernel.  `pytest -m gpu`."""
import numpy as np
import pytest
import torch

import inputs as I
from oracle import rq_oracle as O
from parity import assert_ids_match, load_golden

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture(scope="module")
def ops():
    from rq_vae_recommender_b200 import ops as _ops
    return _ops


def run_tc(ops, x, cbs):
    stats = torch.zeros(4, dtype=torch.int32, device="cuda")
    ids = ops.rq_tokenize_tc(dev(x), [dev(c) for c in cbs], stats=stats)
    torch.cuda.synchronize()
    return ids.cpu().numpy(), stats.cpu().numpy()


def test_tc_supported(ops):
    assert ops.tc_supported(768, 256, 3) and ops.tc_supported(64, 256, 1)
    assert not ops.tc_supported(32, 256, 3) and not ops.tc_supported(768, 128, 3) and not ops.tc_supported(800, 256, 3)


def test_tc_ns_vs_reference(ops):
    g = load_golden("rq_ns2048")
    n, D, K, L = (int(v) for v in g["shape"])
    x, cbs = I.rq_problem(n, D, K, L, seed=1234)
    ids, stats = run_tc(ops, x, cbs)
    n_tie = assert_ids_match(ids, g["eval_ids"], x, cbs, "tc/ns2048")
    assert n_tie <= 2
    assert stats[0] < 0.2 * n * L, stats      # the filter decides most rows alone


@pytest.mark.parametrize("B,D,L", [(1, 768, 3), (100, 768, 3), (128, 768, 1), (129, 768, 2), (1000, 64, 3),
                                    (777, 128, 4), (5000, 256, 3), (4096, 512, 2), (20000, 768, 3)])
def test_tc_vs_oracle_shapes(ops, B, D, L):
    K = 256
    x, cbs = I.rq_problem(max(B, 1024), D, K, L, seed=B + D + L)
    x = x[:B]
    ids, stats = run_tc(ops, x, cbs)
    ref = O.rq_tokenize(x, cbs)
    n_tie = assert_ids_match(ids, ref, x, cbs, f"tc B={B} D={D} L={L}")
    assert n_tie <= max(2, B // 2000)
    exact = ops.rq_tokenize(dev(x), [dev(c) for c in cbs]).cpu().numpy()
    assert (ids != exact).any(1).sum() <= max(2, B // 2000)


def test_tc_scaled_and_extreme_inputs(ops):
    """Rows at very different scales, rows that overflow fp16, exact duplicates of codes, zero rows."""
    D, K, L = 768, 256, 3
    x, cbs = I.rq_problem(2048, D, K, L, seed=99)
    x = x[:512].copy()
    x[0:64] *= 1e-3
    x[64:128] *= 37.0
    x[128:132] *= 1e6            # fp16 overflow -> every code re-ranked exactly
    x[132:136] = 0.0
    x[136:140] = cbs[0][10:14]   # exact hits
    x[140, 5] = np.inf
    ids, stats = run_tc(ops, x[:140], cbs)
    ref = O.rq_tokenize(x[:140], cbs)
    assert_ids_match(ids, ref, x[:140], cbs, "tc/extreme")
    assert (ids[136:140, 0] == np.arange(10, 14)).all()
    ids2, _ = run_tc(ops, x[:141], cbs)          # a non-finite row must not disturb its neighbours
    assert (ids2[:140] == ids).all()


def test_tc_exact_ties_pick_first_index(ops):
    D, K, L = 128, 256, 2
    x, cbs = I.rq_problem(1024, D, K, L, seed=5)
    cbs[0][200] = cbs[0][17]
    cbs[0][90] = cbs[0][17]
    x = np.repeat(cbs[0][17:18], 256, axis=0) + 1e-4 * I.randn(4, 256, D)
    ids, stats = run_tc(ops, x, cbs)
    assert (ids[:, 0] == 17).all()
    assert stats[0] >= 256          # every row needed the exact re-rank at level 0


def test_tc_strided_rows_and_state_reuse(ops):
    D, K, L = 768, 256, 3
    x, cbs = I.rq_problem(1024, D, K, L, seed=21)
    state = ops.TcState([dev(c) for c in cbs])
    big = torch.zeros(1024, 1024, device="cuda")
    big[:, 128:128 + D] = dev(x)
    a = ops.rq_tokenize_tc(big[:, 128:128 + D], state=state).cpu().numpy()
    b = ops.rq_tokenize_tc(dev(x), state=state).cpu().numpy()
    assert np.array_equal(a, b)
    assert_ids_match(b, O.rq_tokenize(x, cbs), x, cbs)


def test_tc_misaligned_rows_take_the_scalar_load_path(ops):
    """x whose base is not 16-byte aligned (and whose row stride is not a multiple of 4) -> the converter's scalar loads."""
    D, K, L = 128, 256, 3
    x, cbs = I.rq_problem(1024, D, K, L, seed=33)
    big = torch.zeros(700, D + 7, device="cuda")
    big[:, 3:3 + D] = dev(x[:700])
    view = big[:, 3:3 + D]                         # offset 3 floats, row stride D+7
    assert view.data_ptr() % 16 != 0 and view.stride(0) % 4 != 0
    a = ops.rq_tokenize_tc(view, [dev(c) for c in cbs]).cpu().numpy()
    b = ops.rq_tokenize_tc(dev(x[:700]), [dev(c) for c in cbs]).cpu().numpy()
    assert np.array_equal(a, b)
    assert_ids_match(b, O.rq_tokenize(x[:700], cbs), x[:700], cbs)


def test_tc_max_levels(ops):
    """L = 8 (RQB_MAX_LEVELS): 28 Gram tables, the serial j >= 2 correction path, all 8 bytes of the packed ids."""
    D, K, L = 64, 256, 8
    x, cbs = I.rq_problem(1024, D, K, L, seed=44)
    ids, stats = run_tc(ops, x[:600], cbs)
    assert ids.shape == (600, 8)
    n_tie = assert_ids_match(ids, O.rq_tokenize(x[:600], cbs), x[:600], cbs, "tc L=8")
    assert n_tie <= 2
    assert not ops.tc_supported(64, 256, 9)


@pytest.mark.gpu
def test_cta_pair_variant_returns_the_same_ids(tmp_path):
    """The opt-in CTA-pair instantiation (RQB200_TC_PAIR=1: clusters of 2, tcgen05 cta_group::2, tensor-map TMA signalling the
    leader's mbarrier) must return exactly the ids of the default single-CTA kernel, including an odd tile count (the pair's
    second CTA then runs past the last tile) and a partial last tile.  The switch is read once per process, hence the
    subprocess, which writes its ids to disk."""
    import os, subprocess, sys
    from rq_vae_recommender_b200 import ops
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shapes = [(129, 768, 3), (513, 256, 4), (1000, 768, 3)]
    code = (
        "import sys, numpy as np, torch\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests', 'golden')!r})\n"
        "import inputs as I\n"
        "from rq_vae_recommender_b200 import ops\n"
        f"for (B, D, L) in {shapes!r}:\n"
        "    x, cbs = I.rq_problem(max(B, 1024), D, 256, L, seed=B + D); x = x[:B]\n"
        "    ids = ops.rq_tokenize_tc(torch.from_numpy(x).cuda(), [torch.from_numpy(c).cuda() for c in cbs])\n"
        f"    np.save({str(tmp_path)!r} + f'/pair_{{B}}_{{D}}_{{L}}.npy', ids.cpu().numpy())\n"
        "print('PAIR DONE')\n"
    )
    env = dict(os.environ, RQB200_TC_PAIR="1")
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "PAIR DONE" in res.stdout, res.stdout + res.stderr
    assert os.environ.get("RQB200_TC_PAIR", "0") != "1", "run this test with the default (single-CTA) kernel in the parent process"
    for (B, D, L) in shapes:
        x, cbs = I.rq_problem(max(B, 1024), D, 256, L, seed=B + D)
        x = x[:B]
        ids = ops.rq_tokenize_tc(torch.from_numpy(x).cuda(), [torch.from_numpy(c).cuda() for c in cbs]).cpu().numpy()
        pair = np.load(tmp_path / f"pair_{B}_{D}_{L}.npy")
        assert np.array_equal(ids, pair), (B, D, L, int((ids != pair).any(axis=1).sum()))


UNVALIDATED = {                      # kernel variants written without GPU access at the end of round 1
    "tc64": {"RQB200_TC_64": "1"},            # 64 rows per CTA, M=128 pair MMAs, x staged by TMA; clusters of 2
    "tc64x4": {"RQB200_TC_64": "4"},          # ... clusters of 4: two pairs share the codebook blocks by TMA multicast
    "tc64x8": {"RQB200_TC_64": "8"},          # ... clusters of 8
    "tc64_g2": {"RQB200_TC_64": "1", "RQB200_TC64_GROUPS": "2"},   # ... two epilogue groups on alternate tiles
    "tc64x4_g2": {"RQB200_TC_64": "4", "RQB200_TC64_GROUPS": "2"},
    "fast": {"RQB200_TC_FASTSCAN": "1"},      # 128-row kernel with rq_tc64_kernel's scan arithmetic (FFMA2, pair insertion, 1-LOP3 keys)
    "tma_fast": {"RQB200_TC_TMA": "1", "RQB200_TC_FASTSCAN": "1"},
    "tma": {"RQB200_TC_TMA": "1"},            # 128-row kernel, x through in-place TMA staging in the A slots
    "tma_pair": {"RQB200_TC_TMA": "1", "RQB200_TC_PAIR": "1"},
    "tma_pair_fast": {"RQB200_TC_TMA": "1", "RQB200_TC_PAIR": "1", "RQB200_TC_FASTSCAN": "1"},   # fewest bytes per row into the SM
}


@pytest.mark.gpu
@pytest.mark.skipif(__import__("os").environ.get("RQB200_TEST_UNVALIDATED", "0") != "1",
                    reason="these kernel variants were written without GPU access at the end of round 1 and have not run on "
                           "hardware yet: bring them up with tools/tc64_bringup.sh first, then set RQB200_TEST_UNVALIDATED=1")
@pytest.mark.parametrize("variant", sorted(UNVALIDATED))
def test_unvalidated_variants_return_the_same_ids(tmp_path, variant):
    """Every opt-in variant must return exactly the ids of the default kernel: odd tile counts (a pair's second CTA past the
    end), a partial last tile (TMA zero fill), inputs that force the `many` path (exact duplicates among the codes), L = 8."""
    import os, subprocess, sys
    from rq_vae_recommender_b200 import ops
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    shapes = [(1, 768, 3), (65, 768, 3), (129, 768, 3), (513, 256, 4), (1000, 768, 3), (20000, 768, 3), (600, 64, 8)]
    code = (
        "import sys, numpy as np, torch\n"
        f"sys.path.insert(0, {root!r}); sys.path.insert(0, {os.path.join(root, 'tests', 'golden')!r})\n"
        "import inputs as I\n"
        "from rq_vae_recommender_b200 import ops\n"
        f"for (B, D, L) in {shapes!r}:\n"
        "    x, cbs = I.rq_problem(max(B, 1024), D, 256, L, seed=B + D); x = x[:B]\n"
        "    if B == 513:\n"
        "        cbs[0][200] = cbs[0][17]; cbs[0][90] = cbs[0][17]; x[:64] = cbs[0][17] + 1e-4 * x[:64]\n"
        "    ids = ops.rq_tokenize_tc(torch.from_numpy(x).cuda(), [torch.from_numpy(c).cuda() for c in cbs])\n"
        f"    np.save({str(tmp_path)!r} + f'/v_{{B}}_{{D}}_{{L}}.npy', ids.cpu().numpy())\n"
        "print('VARIANT DONE')\n"
    )
    env = dict(os.environ, **UNVALIDATED[variant])
    res = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "VARIANT DONE" in res.stdout, res.stdout + res.stderr
    assert not any(os.environ.get(k, "0") != "0" for k in UNVALIDATED[variant]), "run the parent process with the default kernel"
    for (B, D, L) in shapes:
        x, cbs = I.rq_problem(max(B, 1024), D, 256, L, seed=B + D)
        x = x[:B]
        if B == 513:
            cbs[0][200] = cbs[0][17]; cbs[0][90] = cbs[0][17]; x[:64] = cbs[0][17] + 1e-4 * x[:64]
        ids = ops.rq_tokenize_tc(torch.from_numpy(x).cuda(), [torch.from_numpy(c).cuda() for c in cbs]).cpu().numpy()
        got = np.load(tmp_path / f"v_{B}_{D}_{L}.npy")
        assert np.array_equal(ids, got), (variant, B, D, L, int((ids != got).any(axis=1).sum()))
