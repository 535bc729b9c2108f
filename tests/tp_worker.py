"""torchrun worker for tests/test_tp_gpu.py: loads a GGUF tensor-parallel across WORLD_SIZE GPUs, runs a prompt and
a few greedy steps, rank 0 saves logits / ids."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from ntransformer_b200.engine import Model  # noqa: E402

path, out, max_ctx = sys.argv[1], sys.argv[2], int(sys.argv[3])
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
m = Model.load(path, max_ctx, tp_rank=rank, tp_size=world)
m.init_tp()
print(f"[rank {rank}] tp comm ready", flush=True)
prompt = [1, 17, 300, 5, 44, 9]
logits = [m.forward(prompt, 0).copy()]
print(f"[rank {rank}] prefill done", flush=True)
ids, tok, pos = [], int(np.argmax(logits[0])), len(prompt)
for _ in range(12):
    ids.append(tok)
    l = m.forward([tok], pos).copy()
    assert m.argmax() == int(np.argmax(l))
    logits.append(l)
    tok, pos = int(np.argmax(l)), pos + 1
# every rank must hold identical logits (all-gathered) -> replicas stay in lock-step
t = torch.from_numpy(np.stack(logits)).cuda()
ref = t.clone()
dist.broadcast(ref, src=0)
assert torch.equal(t, ref), "ranks diverged"
if rank == 0:
    np.savez(out, logits=np.stack(logits), ids=np.array(ids))
dist.barrier()
torch.cuda.synchronize()
m.close()
print(f"[rank {rank}] closed", flush=True)
dist.destroy_process_group()
