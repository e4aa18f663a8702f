"""BASELINE configs[3]: ONE proof of K U32-add circuits (+ the shared byte table) of different heights over the GPUs of one
box, one rank per GPU over NCCL (tests/_dist_worker.py is the rank). Checks that every rank's proof equals the single-GPU
proof byte for byte and prints the timings and the bytes exchanged.
usage: python tools/dist_prove.py --gpus 2 --log-heights 20,19 [--owners auto] [--reps 3] [--log-blowup 1] [--queries 100]
       --owners rowshard : every matrix split by ROWS over all ranks (host/rowshard_backend.hpp); --kind u32_add | wide:W | multi:K"""
import argparse
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=2)
    ap.add_argument("--log-heights", default="20,19")
    ap.add_argument("--owners", default="auto")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--log-blowup", type=int, default=1)
    ap.add_argument("--queries", type=int, default=100)
    ap.add_argument("--backend", default="nccl")
    ap.add_argument("--out", default="gpurun_out/dist_prove")
    ap.add_argument("--kind", default=None, help="system kind (default multi:K for K heights)")
    ap.add_argument("--no-single", action="store_true", help="skip the single-GPU proof on rank 0 (e.g. when it does not fit)")
    ap.add_argument("--blocks", action="store_true", help="rowshard + wide:W: every rank generates only the rows it reads (implies --no-single)")
    args = ap.parse_args()
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    k = len(args.log_heights.split(","))
    params = dict(log_blowup=args.log_blowup, num_queries=args.queries)
    kind = args.kind or "multi:%d" % k
    if args.blocks:
        args.no_single = True
    procs = []
    for r in range(args.gpus):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(args.gpus), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), DIST_REPS=str(args.reps), NCCL_DEBUG="WARN", DIST_SINGLE="0" if args.no_single else "1",
                   DIST_BLOCKS="1" if args.blocks else "0")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_dist_worker.py"), args.backend, kind,
                                       args.log_heights, args.owners, args.out, json.dumps(params)], env=env))
    rc = [p.wait() for p in procs]
    if any(rc):
        raise SystemExit("a rank failed: %s" % rc)
    proofs = [open("%s.rank%d.proof" % (args.out, r), "rb").read() for r in range(args.gpus)]
    infos = [json.load(open("%s.rank%d.json" % (args.out, r))) for r in range(args.gpus)]
    if args.no_single:
        single, same = proofs[0], all(p == proofs[0] for p in proofs)
        infos[0].setdefault("single_ms", [float("nan")])
        infos[0].setdefault("single_stages", None)
    else:
        single = open(args.out + ".single.proof", "rb").read()
        same = all(p == single for p in proofs)
    res = {"gpus": args.gpus, "kind": kind, "mode": "rowshard" if args.owners == "rowshard" else "circuits", "log_heights": args.log_heights,
           "owner": infos[0]["owner"], "identical_to_single_gpu": same, "all_ranks_same_proof": all(p == proofs[0] for p in proofs),
           "proof_bytes": len(single), "sharded_ms": [min(i["ms"]) for i in infos], "single_gpu_ms": min(infos[0]["single_ms"]),
           "sharded_stages_rank0": infos[0]["stages"], "single_stages": infos[0]["single_stages"],
           "device_bytes_exchanged_per_rank": [i["bytes_dev"] for i in infos], "host_bytes_per_rank": [i["bytes_host"] for i in infos], "comm_ms_per_proof": [i["comm_ms_per_proof"] for i in infos]}
    print(json.dumps(res))
    for r in range(args.gpus):
        os.remove("%s.rank%d.proof" % (args.out, r))
    if not args.no_single:
        os.remove(args.out + ".single.proof")
    if not same:
        raise SystemExit("sharded proof differs from the single-GPU proof")


if __name__ == "__main__":
    main()
