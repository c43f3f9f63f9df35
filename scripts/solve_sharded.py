"""torchrun --nproc-per-node N scripts/solve_sharded.py cfg5 [time limit] : whole solve of a BASELINE.json
configuration with the factor columns sharded over N GPUs (one process per GPU, NCCL over NVLink)."""
import ctypes as C, json, os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lorads_b200 import sdpa
from lorads_b200.capi import Solver, default_params, load_library

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lib = load_library()
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
limit = float(sys.argv[2]) if len(sys.argv) > 2 else 900.0
t = time.time()
if cfg == "cfg1": inst = sdpa.maxcut(800, 19176, 1)
elif cfg == "cfg2": inst = sdpa.maxcut(100_000, 500_000, 3)
elif cfg == "cfg4": inst = sdpa.matrix_completion(20_000, 20_000, 2_000_000, 3, 7)
elif cfg == "cfg5": inst = sdpa.maxcut(1_000_000, 5_000_000, 5)
else: raise SystemExit("unknown config")
if rank == 0: print(f"instance {inst.name} generated in {time.time()-t:.1f}s, world={world}", flush=True)
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    buf = (C.c_char * 128)()
    assert lib.lb2_comm_unique_id(buf) == 0
    uid = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
uid = uid.cuda(); dist.broadcast(uid, 0)
comm = (C.create_string_buffer(bytes(uid.cpu().numpy().tobytes()), 128), rank, world)
t = time.time()
S = Solver(inst, device=local, comm=comm)
dist.barrier()
if rank == 0: print(f"setup (presolve + upload) {time.time()-t:.1f}s rank={S.rank()} local ld={S.info(17)}", flush=True)
t = time.time()
res = S.solve(default_params(verbose=1 if rank == 0 else 0, timeSecLimit=limit))
dist.barrier()
if rank == 0:
    print(json.dumps({k: (float(f"{v:.10g}") if isinstance(v, float) else v) for k, v in res.items()}), f"wall {time.time()-t:.2f}s world={world}", flush=True)
S.close()
dist.barrier()
dist.destroy_process_group()
