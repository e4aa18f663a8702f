"""BASELINE configs[2]: prove() of the wide AIR (256 columns x 2^log_rows rows, log_blowup 2) with its trace in column blocks,
one per GPU (one rank per GPU over NCCL; tests/_wide_prove_worker.py is the rank).
usage: python tools/wide_prove_sharded.py --gpus 4 [--log-rows 22] [--width 256] [--log-blowup 2] [--reps 3] [--single]"""
import argparse
import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ap = argparse.ArgumentParser()
ap.add_argument("--gpus", type=int, default=2)
ap.add_argument("--log-rows", type=int, default=22)
ap.add_argument("--width", type=int, default=256)
ap.add_argument("--log-blowup", type=int, default=2)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--single", action="store_true", help="rank 0 also proves the whole trace alone and compares the proof bytes")
ap.add_argument("--out", default="gpurun_out/wide_prove_sharded")
args = ap.parse_args()
s = socket.socket()
s.bind(("127.0.0.1", 0))
port = s.getsockname()[1]
s.close()
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
procs = []
for r in range(args.gpus):
    env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(args.gpus), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
               NCCL_DEBUG="WARN", WIDE_SINGLE="1" if args.single else "0")
    procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_wide_prove_worker.py"), "nccl", str(args.log_rows),
                                   str(args.width), str(args.log_blowup), args.out, str(args.reps), "100"], env=env))
rc = [p.wait() for p in procs]
if any(rc):
    raise SystemExit("a rank failed: %s" % rc)
infos = [json.load(open("%s.rank%d.json" % (args.out, r))) for r in range(args.gpus)]
res = {"gpus": args.gpus, "log_rows": args.log_rows, "width": args.width, "log_blowup": args.log_blowup,
       "prove_ms_from_host_blocks": min(infos[0]["ms"]), "proof_bytes": infos[0]["proof_bytes"], "phases_ms_rank0": infos[0]["timings"],
       "stages_rank0": infos[0]["stages"], "device_bytes_exchanged_rank0": infos[0]["bytes_dev"]}
if args.single:
    res.update(single_gpu_ms=min(infos[0]["single_ms"]), identical_to_single_gpu=infos[0]["identical"], single_stages=infos[0]["single_stages"])
print(json.dumps(res))
for f in (args.out + ".sharded.proof", args.out + ".single.proof"):
    if os.path.exists(f):
        os.remove(f)
