"""Run under torchrun on N GPUs (tests/test_gpu_sharded.py): one sharded run, then the same run on one GPU."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from polychordlite_b200 import _capi as capi, mgpu  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
capi.set_option("device", local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
D, P, n, R = (int(x) for x in sys.argv[1:5])
s = capi.make_settings(D, P, nlive=n, num_repeats=R, seed=5)
info, dumps, nlike_total = mgpu.run_sharded(s, want_dump=True)
res = dict(rank=rank, logZ=info.logZ, ndead=int(info.ndead), nlike_local=int(info.nlike), nlike_total=nlike_total,
           nupdates=int(info.nupdates), ngen=int(info.ngenerations), device_ms=info.device_ms,
           dead_sum=float(dumps[-1]["dead"].sum()), ndumps=len(dumps))
if rank == 0:
    capi.set_option("batch_K", int(info.batch_K))   # the automatic batch size follows the number of devices
    single, sd = capi.run(s, want_dump=True)
    capi.set_option("batch_K", 0)
    res.update(single_logZ=single.logZ, single_ndead=int(single.ndead), single_nlike=int(single.nlike),
               single_ms=single.device_ms, single_dead_sum=float(sd[-1]["dead"].sum()),
               max_dead_diff=float(np.abs(sd[-1]["dead"] - dumps[-1]["dead"]).max())
               if sd[-1]["dead"].shape == dumps[-1]["dead"].shape else None)
gathered = [None] * world
dist.all_gather_object(gathered, res)
if rank == 0:
    print("SHARDED_RESULT " + json.dumps(gathered), flush=True)
dist.destroy_process_group()
