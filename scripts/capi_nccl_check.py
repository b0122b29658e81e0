"""The sharded standardisation through the C-ABI alone (no torch.distributed on the data path): every rank opens its own
NCCL communicator through ctypes (what a cgo / ctypes stub on the reference side would hold) and calls
asr_cmvn_partial_sums -> asr_cmvn_local_message -> asr_cmvn_exchange_nccl -> asr_cmvn_merge -> asr_cmvn_apply2.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 scripts/capi_nccl_check.py
torch is used for device memory only; the ncclUniqueId travels through a file.  Rank 0 compares with sklearn on the union."""
import ctypes as C
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import asr_b200 as A  # noqa: E402,F401
from asr_b200._lib import lib, check  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
torch.zeros(1, device=dev)                                   # CUDA context

# ---- NCCL through ctypes: the library torch has already loaded (same copy libasr_b200 resolves at run time) ----
import torch.cuda.nccl  # noqa: E402,F401  (makes sure libnccl is in the process)
nccl = C.CDLL("libnccl.so.2")


class UniqueId(C.Structure):
    _fields_ = [("internal", C.c_byte * 128)]


uid = UniqueId()
path = f"/tmp/asr_nccl_uid_{os.environ.get('MASTER_PORT', '0')}"
if rank == 0:
    assert nccl.ncclGetUniqueId(C.byref(uid)) == 0
    with open(path + ".tmp", "wb") as f:
        f.write(bytes(uid.internal))
    os.replace(path + ".tmp", path)
else:
    t0 = time.time()
    while not os.path.exists(path):
        assert time.time() - t0 < 60, "no ncclUniqueId"
        time.sleep(0.05)
    with open(path, "rb") as f:
        C.memmove(C.byref(uid), f.read(128), 128)
comm = C.c_void_p()
nccl.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, UniqueId, C.c_int]
assert nccl.ncclCommInitRank(C.byref(comm), world, uid, rank) == 0

# ---- this rank's rows (uneven shards) ----
D = 1313
rng = np.random.default_rng(5)
full = (rng.standard_normal((1500 * world + 11, D)) * rng.uniform(0.1, 30, D) + rng.uniform(-50, 50, D)).astype(np.float32)
full[:, 7] = 3.25                                            # a constant feature (sklearn: scale 1)
bounds = np.linspace(0, full.shape[0], world + 1).astype(int)
x = torch.from_numpy(full[bounds[rank]:bounds[rank + 1]]).to(dev)
n_loc = x.shape[0]
st = torch.cuda.current_stream().cuda_stream
ws = torch.empty(lib.asr_cmvn_workspace_bytes(D), dtype=torch.uint8, device=dev)
msg = torch.zeros(3 * D + 1, dtype=torch.float64, device=dev)
msgs = torch.zeros((world, 3 * D + 1), dtype=torch.float64, device=dev)
mean, var, scale = (torch.zeros(D, dtype=torch.float64, device=dev) for _ in range(3))
out = torch.empty((n_loc, D), dtype=torch.float64, device=dev)
n1 = lib.asr_cmvn_partial_sums(x.data_ptr(), 1, n_loc, D, D, None, 1, 0, n_loc, 0, ws.data_ptr(), ws.numel(), st)
assert n1 >= 0
n2 = lib.asr_cmvn_partial_sums(x.data_ptr(), 1, n_loc, D, D, None, 2, n1, n_loc, 0, ws.data_ptr(), ws.numel(), st)
assert n2 >= 0
check(lib.asr_cmvn_local_message(ws.data_ptr(), ws.numel(), n1, n2, n_loc, D, msg.data_ptr(), st), "local_message")
check(lib.asr_cmvn_exchange_nccl(comm, msg.data_ptr(), msgs.data_ptr(), D, st), "exchange_nccl")
check(lib.asr_cmvn_merge(msgs.data_ptr(), world, D, mean.data_ptr(), var.data_ptr(), scale.data_ptr(), None, st), "merge")
check(lib.asr_cmvn_apply2(x.data_ptr(), 1, n_loc, D, D, None, None, 0, n1, n2, full.shape[0], mean.data_ptr(), var.data_ptr(),
                          scale.data_ptr(), out.data_ptr(), 2, st), "apply2")
torch.cuda.synchronize()
from sklearn.preprocessing import StandardScaler  # noqa: E402
sc = StandardScaler().fit(full.astype(np.float64))
ref = sc.transform(full.astype(np.float64))[bounds[rank]:bounds[rank + 1]]
em = np.abs(mean.cpu().numpy() - sc.mean_).max() / np.abs(sc.mean_).max()
ev = (np.abs(var.cpu().numpy() - sc.var_) / np.maximum(sc.var_, 1e-300))[sc.var_ > 0].max()
er = np.abs(out.cpu().numpy() - ref).max()
print(f"rank {rank}/{world}: rows {n_loc}; vs sklearn on the union: mean rel {em:.2e}, var rel {ev:.2e}, rows max abs {er:.2e}", flush=True)
assert em <= 1e-12 and ev <= 1e-9 and er <= 1e-9
nccl.ncclCommDestroy.argtypes = [C.c_void_p]
nccl.ncclCommDestroy(comm)
if rank == 0 and os.path.exists(path):
    os.remove(path)
