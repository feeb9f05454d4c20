import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
local=int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local); dev=torch.device("cuda",local)
dist.init_process_group("nccl", device_id=dev)
print("backend", repr(dist.get_backend()), "world", dist.get_world_size(), "devcount", torch.cuda.device_count(), flush=True)
from unopose_b200 import peer as P
print("available", P.available(dev), flush=True)
dist.destroy_process_group()
