"""Committed golden vectors (tests/golden/*.json, written by tests/golden/make_golden.py): the oracle must keep reproducing
them (CPU), and the CUDA path must reproduce them through the C ABI (GPU). Labels of pcs_refs.json are the ones the reference's
own generators print (src/types.rs:246-319), so a machine with cargo can diff them against real Plonky3."""
import hashlib
import json
import os

import numpy as np
import pytest

from tests import _oracle as orc

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with open(os.path.join(HERE, name)) as f:
        return json.load(f)


def limbs(d):
    return [int.from_bytes(bytes(d)[8 * i:8 * i + 8], "little") for i in range(4)]


def gen_refs_matrices():
    m0 = np.zeros((8, 2), dtype=np.uint64); m0[5] = [11, 12]
    m1 = np.zeros((4, 3), dtype=np.uint64); m1[2] = [107, 108, 109]
    m2 = np.zeros((2, 1), dtype=np.uint64); m2[1] = [202]
    return [m0, m1, m2]


def workload(case):
    import multi_stark_b200.system as mss
    return mss.u32_add_workload(1 << case["log_adds"])


# ---- CPU: the oracle against the fixtures -------------------------------------------------------------------------------
def test_golden_matches_survey_prediction():
    """SURVEY 8(c) quotes LEAF3 and COMPRESS computed independently with the `blake3` package"""
    g = load("pcs_refs.json")
    assert g["LEAF3"] == [4163513704854067712, 9384471110237386207, 13671380075168847140, 1533933974187331481]
    assert g["COMPRESS"] == [16432952784711837466, 12565756115161032165, 6915939387221618258, 11123773279136987111]


def test_oracle_reproduces_pcs_refs(oracle):
    g = load("pcs_refs.json")
    t = orc.MmcsTree(oracle, gen_refs_matrices())
    assert limbs(t.root) == g["COMMIT"]
    opened, proof = t.open(5)
    assert [int(x) for x in opened] == sum(g["OPENED"], [])
    for i, s in enumerate(proof):
        assert limbs(s) == g["SIB%d" % i]
    for n in (3, 17, 22, 20):
        row = np.arange(1, n + 1, dtype=np.uint64).reshape(1, n)
        assert limbs(orc.MmcsTree(oracle, [row]).root) == g["LEAF%d" % n]
    ch = np.zeros(8, dtype=np.uint64)
    oracle.orc_gen_challenger_refs(ch)
    assert [int(ch[0]), [int(ch[1]), int(ch[2])], [int(ch[3]), int(ch[4])], [int(ch[5]), int(ch[6])], int(ch[7])] == \
        [g["SAMPLE_BITS"], g["APCS"], g["AFRI"], g["BETA"], g["SAMPLE_BITS2"]]


def test_oracle_reproduces_commit_roots(oracle):
    for case in load("commits.json")["cases"]:
        rng = np.random.default_rng(case["seed"])
        mats = [orc.rand_matrix(rng, h, w) for h, w in case["shapes"]]
        root, h = orc.pcs_commit(oracle, mats, case["log_blowup"])
        oracle.orc_mmcs_free(h)
        assert bytes(root).hex() == case["root"]


@pytest.mark.parametrize("idx", range(5))
def test_oracle_reproduces_proof_digests(oracle, idx):
    case = load("proofs.json")["cases"][idx]
    if case["log_adds"] > 10 and idx == 2:
        pass  # 2^12 rows: BASELINE configs[0], ~1 s on the CPU
    S = orc.OracleSystem(oracle, case["kind"], **case["params"])
    byte, add, claims = workload(case)
    proof, _ = S.prove([byte, add], list(claims))
    assert len(proof) == case["proof_bytes"] and hashlib.sha256(proof).hexdigest() == case["proof_sha256"]
    assert S.preprocessed_commit().hex() == case["preprocessed_commit"]
    S.close()


# ---- GPU: the CUDA path against the fixtures -----------------------------------------------------------------------------
@pytest.fixture(scope="module")
def gpu():
    import multi_stark_b200 as ms
    ctx = ms.GpuContext(0)
    yield ms, ctx
    ctx.close()


@pytest.mark.gpu
def test_gpu_reproduces_pcs_refs(gpu):
    ms, ctx = gpu
    g = load("pcs_refs.json")
    root, pd = ms.GpuMmcs(ctx).commit(gen_refs_matrices())
    assert limbs(root) == g["COMMIT"]
    opened, proofs = pd.open_batch([5])
    assert [int(x) for x in opened[0]] == sum(g["OPENED"], [])
    for i, s in enumerate(proofs[0]):
        assert limbs(s) == g["SIB%d" % i]
    pd.free()
    for n in (3, 17, 22, 20):
        root, pd = ms.GpuMmcs(ctx).commit([np.arange(1, n + 1, dtype=np.uint64).reshape(1, n)])
        assert limbs(root) == g["LEAF%d" % n]
        pd.free()
    # the transcript object of the product (libmshost) cannot be given an empty seed: the challenger vectors are covered by
    # the oracle test above and, for the product, by the proof digests below (every challenge enters the proof bytes)


@pytest.mark.gpu
def test_gpu_reproduces_commit_roots(gpu):
    ms, ctx = gpu
    for case in load("commits.json")["cases"]:
        rng = np.random.default_rng(case["seed"])
        mats = [orc.rand_matrix(rng, h, w) for h, w in case["shapes"]]
        root, pd = ms.GpuPcs(ctx, case["log_blowup"]).commit(mats)
        pd.free()
        assert bytes(root).hex() == case["root"]


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(5))
def test_gpu_reproduces_proof_digests(gpu, idx):
    ms, ctx = gpu
    case = load("proofs.json")["cases"][idx]
    system = ms.System(case["kind"], **case["params"])
    prover = ms.Prover(ctx, system)
    byte, add, claims = workload(case)
    proof = prover.prove([byte, add], claims)
    assert len(proof) == case["proof_bytes"] and hashlib.sha256(proof).hexdigest() == case["proof_sha256"]
    assert prover.preprocessed_commit().hex() == case["preprocessed_commit"]
    n_act = int.from_bytes(proof[:8], "little")
    o = 8 + n_act
    assert proof[o:o + 32].hex() == case["stage_1_commit"] and proof[o + 64:o + 96].hex() == case["quotient_commit"]
    prover.close()
    system.close()
