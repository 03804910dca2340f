"""GPU, >= 2 devices: the multi-GPU path over NCCL -- organism shards on two ranks, statistics all-gathered out of
device memory (flemingo_b200.distributed.DeviceGather) -- reproduces the single-GPU result bit for bit on every rank."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parents[1]

WORKER = r"""
import json, os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from flemingo_b200 import synthetic
from flemingo_b200.distributed import evaluate_sharded, organism_costs, shard_by_cost
from flemingo_b200.engine import Engine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
device = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=device)
orgs = synthetic.mixed_organisms(301, 300, 3) + synthetic.pssm_organisms(20, 300, 4)
pos = synthetic.random_sequences(300, 300, 1)
neg = synthetic.random_sequences(260, 300, 2)
eng = Engine(local)
out = {}
for method, gamma in (("Welch", 0.2), ("Yuen", 0.2), ("Trim-left-M-SD", 0.5)):
    full = evaluate_sharded(eng, orgs, pos, neg, method, gamma, rank, world, device)
    # every rank holds the same gathered array
    t = torch.from_numpy(full.copy()).to(device)
    ref = t.clone()
    dist.broadcast(ref, src=0)
    assert torch.equal(t.view(torch.int64), ref.view(torch.int64)), "ranks disagree on the gathered statistics"
    if rank == 0:
        alone = eng.evaluate(orgs, pos, neg, method, gamma)
        out[method] = bool(np.array_equal(full.view(np.int64), alone.view(np.int64)))
# more ranks than organisms: an empty shard contributes a block of zeros
few = orgs[:1]
full = evaluate_sharded(eng, few, pos, neg, "Welch", 0.2, rank, world, device)
if rank == 0:
    out["one_organism"] = bool(np.array_equal(full, eng.evaluate(few, pos, neg, "Welch", 0.2)))
    shards = shard_by_cost(organism_costs(orgs, 300), world)
    out["shard_sizes"] = [int(len(s)) for s in shards]
    json.dump(out, open(sys.argv[2], "w"))
dist.destroy_process_group()
eng.close()
"""


def test_two_rank_nccl_sharded_equals_unsharded(tmp_path):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    result = tmp_path / "result.json"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", str(script), str(REPO), str(result)]
    run = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env, cwd=str(REPO))
    assert run.returncode == 0, run.stdout[-3000:] + run.stderr[-3000:]
    out = json.loads(result.read_text())
    assert out["Welch"] and out["Yuen"] and out["Trim-left-M-SD"] and out["one_organism"], out
    assert sum(out["shard_sizes"]) == 321 and min(out["shard_sizes"]) > 100


def test_single_process_multi_device_evaluator_equals_one_engine():
    """The GA driver's multi-GPU mode: ONE process, one engine per device (MultiDeviceEvaluator) -- same doubles as a
    single engine evaluating the whole population."""
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    from flemingo_b200 import synthetic
    from flemingo_b200.distributed import MultiDeviceEvaluator
    from flemingo_b200.engine import Engine
    orgs = synthetic.mixed_organisms(150, 300, 7) + synthetic.pssm_organisms(30, 300, 8)
    pos = synthetic.random_sequences(200, 300, 1)
    neg = synthetic.random_sequences(180, 300, 2)
    multi = MultiDeviceEvaluator([0, 1])
    one = Engine(0)
    try:
        for method, gamma in (("Welch", 0.2), ("Trim-left-M", 0.3)):
            a = multi(orgs, pos, neg, method, gamma)
            b = one.evaluate(orgs, pos, neg, method, gamma)
            assert np.array_equal(a.view(np.int64), b.view(np.int64))
        assert np.array_equal(multi(orgs[:1], pos, neg), one.evaluate(orgs[:1], pos, neg))   # fewer organisms than devices
    finally:
        multi.close()
        one.close()
