"""The row-sharded prover over PEER MEMORY (multi_stark_b200/csrc/peer.cu): one rank per GPU over NCCL, matrices exchanged by
this library's kernels through CUDA-IPC windows (row blocks pushed into the owners' column blocks, row shards assembled from the
peers' LDE column blocks by the leaf-hash pass, subtree roots stored into every peer, flag barriers in peer memory). Needs at
least two GPUs: skipped on a single-GPU box, where tests/test_gpu_rowshard.py covers the same protocol over gloo and
tests/test_gpu_merkle.py::test_commit_from_column_blocks the assembling leaf kernels. The proof must be byte-identical to the
single-GPU proof, with peer memory and (MSGPU_P2P=0) with the NCCL exchange."""
import pytest

from tests.test_gpu_dist_prove import run_world

pytestmark = pytest.mark.gpu


def _gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


CASES = [
    # world, kind, log_heights, params, extra environment
    (2, "u32_add", [12], dict(log_blowup=1, num_queries=20), {}),
    (2, "u32_add", [12], dict(log_blowup=1, num_queries=20), {"MSGPU_P2P": "0"}),             # the same over NCCL collectives
    (2, "multi:5", [13, 13, 13, 12, 11], dict(log_blowup=1, num_queries=10), {}),               # several assembled matrices per height
    (2, "multi:3", [14, 12, 13], dict(log_blowup=1, num_queries=10), {"MSGPU_PEER_WINDOW_MB": "1"}),  # the heap grows mid-proof
    (2, "wide:16", [12], dict(log_blowup=2, num_queries=10), {}),                               # narrow stage-2 trace split too
    (4, "u32_add", [13], dict(log_blowup=1, num_queries=15), {}),                               # next rows fetched from another shard
    (4, "multi:3", [14, 12, 13], dict(log_blowup=1, num_queries=10), {}),
    (8, "multi:3", [15, 13, 14], dict(log_blowup=1, num_queries=10), {}),
]


@pytest.mark.parametrize("world,kind,log_heights,params,env", CASES)
def test_peer_memory_proof_is_byte_identical(tmp_path, world, kind, log_heights, params, env):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    proofs, single, infos = run_world(tmp_path, world, kind, log_heights, "rowshard", params, backend="nccl", extra_env=env)
    for r in range(world):
        assert proofs[r] == single, "rank %d's proof differs from the single-GPU proof" % r
    assert all(i["bytes_dev"] > 0 for i in infos)
